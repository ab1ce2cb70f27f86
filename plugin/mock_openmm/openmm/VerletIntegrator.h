/* STAND-IN for openmm/VerletIntegrator.h (the tests only construct one). */
#ifndef MOCK_OPENMM_VERLETINTEGRATOR_H_
#define MOCK_OPENMM_VERLETINTEGRATOR_H_
#include "Integrator.h"
namespace OpenMM { class VerletIntegrator : public Integrator { public: explicit VerletIntegrator(double stepSize) : stepSize(stepSize) {} double getStepSize() const { return stepSize; } private: double stepSize; }; }
#endif
