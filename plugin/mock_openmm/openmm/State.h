/* STAND-IN for openmm/State.h. */
#ifndef MOCK_OPENMM_STATE_H_
#define MOCK_OPENMM_STATE_H_
#include "Vec3.h"
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class State {
public:
    enum DataType { Positions = 1, Velocities = 2, Forces = 4, Energy = 8, Parameters = 16, ParameterDerivatives = 32 };
    State() : energy(0) {}
    const std::vector<Vec3>& getPositions() const { return positions; }
    const std::vector<Vec3>& getForces() const { return forces; }
    double getPotentialEnergy() const { return energy; }
    double getKineticEnergy() const { return 0.0; }
    const std::map<std::string, double>& getParameters() const { return parameters; }
    const std::map<std::string, double>& getEnergyParameterDerivatives() const { return derivatives; }
    void getPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box[0]; b = box[1]; c = box[2]; }
private:
    friend class Context;
    std::vector<Vec3> positions, forces;
    double energy;
    std::map<std::string, double> parameters, derivatives;
    Vec3 box[3];
};
}
#endif
