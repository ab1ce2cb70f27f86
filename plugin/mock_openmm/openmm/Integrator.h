/* STAND-IN for openmm/Integrator.h. */
#ifndef MOCK_OPENMM_INTEGRATOR_H_
#define MOCK_OPENMM_INTEGRATOR_H_
namespace OpenMM { class Integrator { public: virtual ~Integrator() {} }; }
#endif
