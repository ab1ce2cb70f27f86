/* Takes the place of the reference's platforms/reference/include/ReferenceExtendedCustomCVForce.h (which needs OpenMM's Lepton)
 * for the run of the reference's OWN Reference-platform kernel file (plugin/Makefile, run-reference-tests-own): that file also
 * holds the ExtendedCustomCVForce kernel, which is outside the path and never created here; it only has to compile.  Force-included
 * (-include) ahead of everything: it defines the real header's include guard, so the real header -- found first, next to the
 * file that includes it -- comes out empty. */
#ifndef __ReferenceExtendedCustomCVForce_H__
#define __ReferenceExtendedCustomCVForce_H__
#include "ExtendedCustomCVForce.h"
#include "openmm/Vec3.h"
#include "openmm/internal/ContextImpl.h"
#include <map>
#include <string>
#include <vector>
namespace OpenMMLab {
class ReferenceExtendedCustomCVForce {
public:
    ReferenceExtendedCustomCVForce(const ExtendedCustomCVForce&) {}
    void updateTabulatedFunctions(const ExtendedCustomCVForce&) {}
    void calculateIxn(OpenMM::ContextImpl&, std::vector<OpenMM::Vec3>&, const std::map<std::string, double>&, std::vector<OpenMM::Vec3>&,
                      double*, std::map<std::string, double>&) const {}
};
}
#endif
