/*
 * Physical constants shared by the CUDA core and the CPU oracle.
 *
 * ONE_4PI_EPS0 is an OpenMM symbol (openmm/reference/SimTKOpenMMRealType.h, not
 * vendored in the reference tree).  OpenMM <= 8.0 used 138.935456; OpenMM 8.1
 * (the version the reference's CI pins, .github/workflows/Linux.yml:24,31) uses
 * the CODATA-2018 value below.  The reference only ever uses the symbol
 * (e.g. ReferenceSlicedLJCoulombIxn.cpp:186,371), so parity needs both sides
 * of a comparison to agree on one value: this header is that single place.
 */
#ifndef SNB_CONSTANTS_H_
#define SNB_CONSTANTS_H_

#define SNB_ONE_4PI_EPS0 138.93545764438198
#define SNB_PME_ORDER 5   /* ReferenceSlicedLJCoulombIxn.cpp:216 ; CudaOpenMMLabKernels.h:159 */

#endif
