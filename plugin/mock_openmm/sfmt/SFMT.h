/* STAND-IN for OpenMM's sfmt/SFMT.h: the two calls the reference's tests make, on std::mt19937 (the tests draw random
 * coordinates and charges; no result depends on the particular sequence). */
#ifndef MOCK_OPENMM_SFMT_H_
#define MOCK_OPENMM_SFMT_H_
#include <random>
namespace OpenMM_SFMT {
class SFMT { public: std::mt19937 engine; };
static inline void init_gen_rand(unsigned int seed, SFMT& sfmt) { sfmt.engine.seed(seed); }
static inline double genrand_real2(SFMT& sfmt) { return (sfmt.engine()>>5)*(1.0/134217728.0); }       /* [0, 1) */
static inline unsigned int gen_rand32(SFMT& sfmt) { return (unsigned int) sfmt.engine(); }
}
#endif
