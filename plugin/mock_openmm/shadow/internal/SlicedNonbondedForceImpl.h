/* Shadows the reference's openmmapi/include/internal/SlicedNonbondedForceImpl.h for the link-and-run check of plugin/Makefile
 * (target run-check): the reference's own SlicedNonbondedForce.cpp is compiled UNCHANGED from where it lies and only needs the
 * Impl class to exist (createImpl, the three *InContext forwards); the real Impl drags in OpenMM's ForceImpl / Kernel /
 * ContextImpl machinery, which this image does not have.  The check drives the kernel directly, as the reference's Impl would
 * (openmmapi/src/SlicedNonbondedForceImpl.cpp:34,132-142,356-367). */
#ifndef MOCK_SLICED_NONBONDED_FORCE_IMPL_H_
#define MOCK_SLICED_NONBONDED_FORCE_IMPL_H_
#include "SlicedNonbondedForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/internal/ForceImpl.h"
namespace OpenMMLab {
class SlicedNonbondedForceImpl : public OpenMM::ForceImpl {
public:
    SlicedNonbondedForceImpl(const SlicedNonbondedForce& owner) : owner(owner) {}
    void getPMEParameters(double&, int&, int&, int&) const { throw OpenMM::OpenMMException("mock OpenMM: no Context"); }
    void getLJPMEParameters(double&, int&, int&, int&) const { throw OpenMM::OpenMMException("mock OpenMM: no Context"); }
    void updateParametersInContext(OpenMM::ContextImpl&) { throw OpenMM::OpenMMException("mock OpenMM: no Context"); }
private:
    const SlicedNonbondedForce& owner;
};
}
#endif
