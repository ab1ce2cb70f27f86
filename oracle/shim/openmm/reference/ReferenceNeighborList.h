// Stand-in for openmm/reference/ReferenceNeighborList.h: the reference only
// iterates the list (ReferenceSlicedLJCoulombIxn.cpp:351,524); the list itself is
// built by the oracle harness (oracle/harness.cpp) in place of OpenMM's
// computeNeighborListVoxelHash (call site ReferenceOpenMMLabKernels.cpp:203).
#ifndef SNB_SHIM_NEIGHBORLIST_H_
#define SNB_SHIM_NEIGHBORLIST_H_
#include <set>
#include <utility>
#include <vector>
namespace OpenMM {
typedef std::pair<int, int> AtomPair;
class NeighborList : public std::vector<AtomPair> {};
}
#endif
