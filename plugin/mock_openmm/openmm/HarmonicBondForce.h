/* STAND-IN for openmm/HarmonicBondForce.h: E = k/2 (r - r0)^2 per bond, evaluated by the stand-in Context itself. */
#ifndef MOCK_OPENMM_HARMONICBONDFORCE_H_
#define MOCK_OPENMM_HARMONICBONDFORCE_H_
#include "Force.h"
#include <vector>
namespace OpenMM {
class HarmonicBondForce : public Force {
public:
    struct Bond { int p1, p2; double length, k; };
    int addBond(int particle1, int particle2, double length, double k) { Bond b = {particle1, particle2, length, k}; bonds.push_back(b); return (int) bonds.size()-1; }
    int getNumBonds() const { return (int) bonds.size(); }
    const Bond& getBond(int index) const { return bonds.at(index); }
protected:
    ForceImpl* createImpl() const;
private:
    std::vector<Bond> bonds;
};
}
#endif
