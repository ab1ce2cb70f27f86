/*
 * slicednb.h -- C-ABI of the B200-native SlicedNonbondedForce kernel core.
 *
 * This is the drop-in boundary.  A platform plugin that implements
 * OpenMMLab::CalcSlicedNonbondedForceKernel
 * (reference: openmmapi/include/OpenMMLabKernels.h:29-87) forwards each of its
 * virtual methods to one entry point below:
 *
 *   initialize(system, force)            -> snb_create          (OpenMMLabKernels.h:50)
 *   execute(context, f, e, direct, recip)-> snb_execute         (OpenMMLabKernels.h:61)
 *   copyParametersToContext(ctx, force)  -> snb_update_parameters (OpenMMLabKernels.h:68)
 *   getPMEParameters(alpha,nx,ny,nz)     -> snb_get_pme_parameters   (OpenMMLabKernels.h:77)
 *   getLJPMEParameters(alpha,nx,ny,nz)   -> snb_get_ljpme_parameters (OpenMMLabKernels.h:86)
 *   ~KernelImpl                          -> snb_destroy
 *
 * The descriptor `snb_desc` is a flat snapshot of everything
 * ReferenceCalcSlicedNonbondedForceKernel::initialize reads from the
 * SlicedNonbondedForce and the System
 * (platforms/reference/src/ReferenceOpenMMLabKernels.cpp:65-191).
 *
 * Plain pointers and sizes only; no C++ or torch types; no exception crosses
 * this boundary (status codes + snb_last_error()).
 *
 * The SAME descriptor and execute signature are implemented by the CPU oracle
 * (oracle/, test infrastructure only) under the prefix snb_oracle_*, so parity
 * tests feed both sides byte-identical inputs.
 */
#ifndef SLICEDNB_H_
#define SLICEDNB_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* NonbondedMethod, same values as CalcSlicedNonbondedForceKernel::NonbondedMethod
 * (openmmapi/include/OpenMMLabKernels.h:31-38). */
enum {
    SNB_NO_CUTOFF = 0,
    SNB_CUTOFF_NON_PERIODIC = 1,
    SNB_CUTOFF_PERIODIC = 2,
    SNB_EWALD = 3,
    SNB_PME = 4,
    SNB_LJPME = 5
};

/* execute() flags = the four bools of CalcSlicedNonbondedForceKernel::execute
 * (openmmapi/include/OpenMMLabKernels.h:61). */
enum {
    SNB_INCLUDE_FORCES = 1,
    SNB_INCLUDE_ENERGY = 2,
    SNB_INCLUDE_DIRECT = 4,
    SNB_INCLUDE_RECIPROCAL = 8,
    SNB_EXECUTE_ALL = 15
};

/* status codes */
enum {
    SNB_OK = 0,
    SNB_ERR_INVALID = 1,        /* bad descriptor / argument                         */
    SNB_ERR_BOX_TOO_SMALL = 2,  /* ReferenceOpenMMLabKernels.cpp:208-210             */
    SNB_ERR_PARTICLE_COUNT = 3, /* ReferenceOpenMMLabKernels.cpp:264-265             */
    SNB_ERR_EXCEPTION_SET = 4,  /* ReferenceOpenMMLabKernels.cpp:290-291             */
    SNB_ERR_WRONG_METHOD = 5,   /* ReferenceOpenMMLabKernels.cpp:315-316,324-325     */
    SNB_ERR_CUDA = 6,           /* CUDA / cuFFT / NCCL failure                        */
    SNB_ERR_UNSUPPORTED = 7
};

/* creation flags (product only; ignored by the oracle).  They mirror the platform properties the
 * reference's CUDA platform reads (platforms/cuda/src/CudaOpenMMLabKernels.cpp:483,496):
 * Precision = mixed (default) | double, and DeterministicForces. */
enum {
    SNB_FLAG_DETERMINISTIC = 1,     /* PME spreading in 64-bit fixed point: bit-reproducible forces      */
    SNB_FLAG_DOUBLE_PRECISION = 2,  /* all pair / grid arithmetic in FP64 (implies deterministic)        */
    /* which direct-space kernel runs the PME path in an orthorhombic box with <= 2 subsets (both give the same pair
     * set and the same mixed-precision results; every other case has one kernel).  Neither bit: by system size --
     * the packed-FP32 kernel from SNB_PACKED_MIN_ATOMS atoms on, the general one below (platform property DirectKernel) */
    SNB_FLAG_DIRECT_PACKED = 4,
    SNB_FLAG_DIRECT_GENERAL = 8
};
#define SNB_PACKED_MIN_ATOMS 65536
    /* PME interpolation goes through shared-memory bricks (TMA) from this many atoms on; the per-atom kernel below */
#define SNB_PME_BRICK_MIN_ATOMS 250000

typedef struct snb_desc {
    int32_t struct_size;               /* = sizeof(snb_desc); guards against ABI drift */
    int32_t num_particles;
    int32_t num_subsets;               /* SlicedNonbondedForce::getNumSubsets          */
    int32_t method;                    /* SNB_NO_CUTOFF .. SNB_LJPME                    */
    double  cutoff;                    /* getCutoffDistance (nm)                        */
    int32_t use_switching_function;
    int32_t use_dispersion_correction;
    double  switching_distance;
    double  reaction_field_dielectric;
    int32_t exceptions_use_pbc;        /* getExceptionsUsePeriodicBoundaryConditions    */
    int32_t flags;                     /* SNB_FLAG_*                                    */
    double  ewald_error_tolerance;     /* used only when an alpha below is 0            */
    double  default_box[9];            /* System default periodic box, rows a,b,c       */
    double  ewald_alpha;               /* getPMEParameters; 0 => derive from tolerance  */
    int32_t pme_grid[3];
    int32_t device_index;              /* CUDA device of this handle                    */
    double  dispersion_alpha;          /* getLJPMEParameters; 0 => derive               */
    int32_t dispersion_grid[3];
    int32_t reserved0;

    const double*  charge;             /* [N] base particle parameters                  */
    const double*  sigma;              /* [N]                                           */
    const double*  epsilon;            /* [N]                                           */
    const int32_t* subset;             /* [N] getParticleSubset                         */

    int32_t num_exceptions;
    int32_t num_global_parameters;
    const int32_t* exception_particles;/* [X][2]                                        */
    const double*  exception_params;   /* [X][3] chargeProd, sigma, epsilon             */

    const double*  global_parameter_defaults; /* [G]                                    */

    int32_t num_particle_offsets;
    int32_t num_exception_offsets;
    const int32_t* particle_offset_index;  /* [P][2] global parameter, particle         */
    const double*  particle_offset_scale;  /* [P][3] chargeScale, sigmaScale, epsScale  */
    const int32_t* exception_offset_index; /* [E][2] global parameter, exception        */
    const double*  exception_offset_scale; /* [E][3]                                    */

    int32_t num_scaling_parameters;
    int32_t num_derivatives;
    const int32_t* scaling_parameters; /* [K][5] globalParam, subset1, subset2, includeCoulomb, includeLJ */
    const int32_t* derivative_parameters; /* [D] global-parameter index of each requested derivative      */
} snb_desc;

typedef struct snb_handle snb_handle;

/* slice index of a subset pair: SlicedNonbondedForce.h:22 */
static inline int snb_slice_index(int i, int j) { return i > j ? i*(i+1)/2 + j : j*(j+1)/2 + i; }

/* ---- lifetime ------------------------------------------------------------ */

/* initialize(): ReferenceOpenMMLabKernels.cpp:65-191 / CudaOpenMMLabKernels.cpp:254-886.
 * Copies everything it needs; `desc` and its arrays may be freed on return. */
int snb_create(const snb_desc* desc, snb_handle** out);
void snb_destroy(snb_handle* h);

/* copyParametersToContext(): ReferenceOpenMMLabKernels.cpp:263-312.
 * Particle count and the set of non-excluded exceptions must be unchanged. */
int snb_update_parameters(snb_handle* h, const snb_desc* desc);

/* ---- evaluation ---------------------------------------------------------- */

/* execute(): ReferenceOpenMMLabKernels.cpp:193-261.
 *   positions          [N][3] nm, host memory
 *   box                [9] rows a,b,c of the reduced triclinic box (ignored for non-periodic methods)
 *   global_parameters  [G] current Context parameter values (context.getParameter)
 *   flags              SNB_INCLUDE_*
 *   forces             [N][3] kJ/mol/nm, host; ADDED to (may be NULL when !FORCES)
 *   energy             out: sum_{slice,term} lambda*E  (0 if !ENERGY)
 *   slice_energies     [L][2] (Coulomb, vdW) raw slice energies, overwritten; may be NULL
 *   derivatives        [D] ADDED to, regardless of ENERGY (reference :252-258); may be NULL iff D==0
 */
int snb_execute(snb_handle* h, const double* positions, const double* box,
                const double* global_parameters, int flags,
                double* forces, double* energy, double* slice_energies, double* derivatives);

/* Same evaluation with inputs/outputs resident in device memory (no host copies):
 *   posq_dev    [N] float4 (x,y,z,unused) in ORIGINAL particle order
 *   forces_dev  [3][N] long long, 2^32 fixed point, component-major; ADDED to; may be 0
 * Energies/derivatives are read back with snb_read_energies (one small D2H). */
int snb_execute_device(snb_handle* h, const void* posq_dev, const double* box,
                       const double* global_parameters, int flags, void* forces_dev);
int snb_read_energies(snb_handle* h, int flags, double* energy, double* slice_energies,
                      double* derivatives);

/* The same, in the layout OpenMM's CUDA platform keeps its atoms in -- what a zero-copy binding to that platform hands over
 * (the reference: platforms/cuda/src/CudaOpenMMLabKernels.cpp:736-758 registers posq / parameters with
 * CudaNonbondedUtilities; force-buffer layout platforms/common/src/kernels/pme.cc:399-405).  Atoms sit in the platform's
 * REORDERED order, padded to a multiple of 32:
 *   posq             [padded] float4 (x, y, z, q), or double4 when positions_double (platform in double precision)
 *   posq_correction  [padded] float4 added to posq (platform in mixed precision), or NULL
 *   atom_index       [N] original index of the atom in slot i (CudaContext::getAtomIndexArray), or NULL = original order
 *   forces           [3][padded] long long, 2^32 fixed point: force[c*padded + slot]; ADDED to; may be NULL
 * All pointers are device pointers on the handle's device.  plugin/src/B200CudaOpenMMLabKernels.cpp is the binding. */
typedef struct snb_device_layout {
    int32_t struct_size;            /* = sizeof(snb_device_layout) */
    int32_t padded_num_atoms;
    int32_t positions_double;
    int32_t reserved;
    const void* posq;
    const void* posq_correction;
    const int32_t* atom_index;
    void* forces;
} snb_device_layout;
int snb_execute_device_ordered(snb_handle* h, const snb_device_layout* layout, const double* box,
                               const double* global_parameters, int flags);

int snb_get_pme_parameters(const snb_handle* h, double* alpha, int* nx, int* ny, int* nz);
int snb_get_ljpme_parameters(const snb_handle* h, double* alpha, int* nx, int* ny, int* nz);

/* ---- multi-GPU ----------------------------------------------------------
 * One process (one handle) per GPU of one box.  The tile list and the exceptions are
 * partitioned across the ranks, reciprocal space / self energy / dispersion correction stay
 * on rank 0 -- the split of the reference's multi-device kernel
 * (platforms/cuda/src/CudaParallelOpenMMLabKernels.cpp:19-53;
 * CudaOpenMMLabKernels.cpp:379,405,466,680-683,762-765).  Every rank calls snb_execute* with
 * the same box, global parameters and flags; rank 0's positions are broadcast (ncclBroadcast;
 * skipped when the caller declares them identical everywhere, snb_set_comm_options) and ONE
 * ncclAllReduce(sum, int64) per evaluation carries the 2^32 fixed-point forces, the slice
 * energies and the bookkeeping flags -- integer sums are exact, so the forces are bit-identical
 * to the single-GPU ones -- and leaves forces, energies and derivatives on every rank. */
int snb_nccl_unique_id(void* id128);             /* rank 0: fills a 128-byte ncclUniqueId; ship it to the other ranks */
int snb_comm_init(snb_handle* h, const void* id128, int rank, int world_size);
/* rank 0's share of the tile list relative to 1.0 for every other rank (it also runs reciprocal
 * space); negative = built-in cost model */
int snb_set_rank0_share(snb_handle* h, double share);
/* broadcast_positions = 0: every rank is given identical positions by its caller, no broadcast is needed */
int snb_set_comm_options(snb_handle* h, int broadcast_positions);
/* What snb_create derives from a descriptor on the host, without touching a GPU (tests compare it with the CPU oracle):
 * Ewald / PME separation parameter and grid (NonbondedForceImpl::calcPMEParameters, call sites
 * ReferenceOpenMMLabKernels.cpp:171,176,178), the LJPME pair, the classic-Ewald reciprocal vectors per axis
 * (calcEwaldParameters, :169-173) and the dispersion-correction coefficient of every slice
 * (SlicedNonbondedForceImpl::calcDispersionCorrections, SlicedNonbondedForceImpl.cpp:263-354).  Entries that do not
 * apply to desc->method come back 0; any output pointer may be null.  dispersion has num_subsets (num_subsets+1)/2 entries. */
int snb_host_parameters(const snb_desc* desc, double* alpha, int32_t grid[3], double* dispersion_alpha, int32_t dispersion_grid[3],
                        int32_t kmax[3], double* dispersion);
/* the partition rule itself (host only, no GPU): contiguous range [lo, hi) of n units for `rank` */
int snb_partition(int64_t n, int world_size, int rank, double rank0_share, int64_t* lo, int64_t* hi);

/* ---- introspection for tests and benches --------------------------------- */

/* Sorted list of interacting pairs (i<j, lexicographic) found by the device
 * neighbour search of the LAST execute; *pairs is malloc'ed, caller frees
 * with snb_free. */
int snb_dump_pairs(snb_handle* h, int32_t** pairs, int64_t* num_pairs);
/* number of interacting pairs of the last execute on this handle (this rank's share), without the list */
int snb_count_pairs(snb_handle* h, int64_t* num_pairs);
void snb_free(void* p);
/* test hook: set the capacity of the tile list (chunks of 32 entries).  A value that is too small makes the next
 * evaluation overflow, grow the buffers and repeat -- the path a locally dense system takes on its first evaluation. */
int snb_set_list_capacity(snb_handle* h, int64_t chunks);

/* Tile-list reuse (platform property ListSkin).  skin_nm > 0: the sorted order and the tile list are built with
 * cutoff + skin and KEPT between evaluations until some atom has moved more than skin/2 since the build (checked on
 * the device every evaluation), the box vectors change, or the list buffers / the multi-GPU partition do; block
 * boxes and local coordinates are refreshed every evaluation.  The exact cutoff predicate still decides the pair set
 * every evaluation, so energies, forces and the set of interacting pairs do not depend on when the list was built.
 * 0 (default): rebuilt every evaluation, as the Reference platform does
 * (platforms/reference/src/ReferenceOpenMMLabKernels.cpp:203); OpenMM's CUDA platform, from which the reference's
 * CUDA kernels inherit their neighbour list, pads and reuses its list the same way.  A periodic box clips the skin
 * to 0.9 (L_min/2 - cutoff). */
int snb_set_list_skin(snb_handle* h, double skin_nm);

/* per-phase device times (ms) of the last execute, when timing is enabled */
enum { SNB_PHASE_SORT = 0, SNB_PHASE_LIST, SNB_PHASE_DIRECT, SNB_PHASE_BONDED,
       SNB_PHASE_SPREAD, SNB_PHASE_FFT, SNB_PHASE_CONV, SNB_PHASE_INTERP,
       SNB_PHASE_REDUCE, SNB_PHASE_TOTAL, SNB_NUM_PHASES };
int snb_set_timing(snb_handle* h, int enabled);
int snb_get_phase_times(snb_handle* h, float* ms /*[SNB_NUM_PHASES]*/);
/* CUDA-event time (ms) of the last execute on the handle's stream: first memset .. last read-back copy
 * (excludes the host->device copy of the positions, which precedes it) */
int snb_get_last_device_ms(snb_handle* h, float* ms);
/* counters of the last execute: [0] 1 when the sort ran as the one cooperative kernel (k_sortFused), 0 for the separate
 * kernels, -1 when its launch was refused once and it is off for this handle,
 * [1] chunks of 32 tile-list entries, [2] kernel launches of this library, [3] k_computeParameters launches since
 * creation, [4] 1 when the evaluation was replayed from a captured CUDA graph, [5] 1 when the packed-FP32 direct-space
 * kernel (k_directPacked) took the launch, 0 for the general one (k_direct), [6] tile-list rebuilds and [7]
 * direct-space evaluations since list reuse was switched on (snb_set_list_skin) */
int snb_get_counters(snb_handle* h, int64_t* out /*[8]*/);

const char* snb_last_error(void);
const char* snb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SLICEDNB_H_ */
