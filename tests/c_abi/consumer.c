/* A plain-C consumer of the drop-in boundary (include/slicednb.h): fills snb_desc by hand -- no Python mirror, no ctypes --
 * creates a kernel, evaluates known-answer cases and checks them against their closed forms, the way the reference's own
 * tests do (tests/TestSlicedNonbondedForce.h: testCoulomb :73-95, testLJ :97-121, testCutoff :210-246; the slicing identities
 * of :1141-1165).  Built twice by tests/test_c_consumer.py: against the CPU oracle (-DSNB_ORACLE: the same entry points under
 * the prefix snb_oracle_) and against the CUDA library.  Exit code 0 and "OK" = every check held. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "slicednb.h"
#include "snb_constants.h"

#ifdef SNB_ORACLE
int snb_oracle_create(const snb_desc*, snb_handle**);
void snb_oracle_destroy(snb_handle*);
int snb_oracle_execute(snb_handle*, const double*, const double*, const double*, int, double*, double*, double*, double*);
const char* snb_oracle_last_error(void);
#define snb_create snb_oracle_create
#define snb_destroy snb_oracle_destroy
#define snb_execute snb_oracle_execute
#define snb_last_error snb_oracle_last_error
#endif

static int failures = 0;

static void expect(const char* what, double expected, double found, double tol) {
    /* assertEqualTo, openmmapi/include/internal/AssertionUtilities.h:7-14 */
    const double scale = fabs(expected) > 1.0 ? fabs(expected) : 1.0;
    if (!(fabs(expected-found)/scale <= tol)) {
        printf("FAIL %s: expected %.12g, found %.12g\n", what, expected, found);
        failures++;
    }
}

static void base_desc(snb_desc* d, int n, int subsets, int method) {
    memset(d, 0, sizeof(*d));
    d->struct_size = (int32_t) sizeof(*d);
    d->num_particles = n;
    d->num_subsets = subsets;
    d->method = method;
    d->cutoff = 1.0;
    d->switching_distance = -1.0;
    d->reaction_field_dielectric = 78.3;
    d->ewald_error_tolerance = 5e-4;
    d->default_box[0] = d->default_box[4] = d->default_box[8] = 4.0;
}

/* two particles in different subsets, no cutoff; lambda scales Coulomb + LJ of slice (0,1); its derivative is requested */
static void two_particles_no_cutoff(void) {
    const double q[2] = {1.0, -1.5}, sig[2] = {0.3, 0.34}, eps[2] = {0.5, 0.8};
    const int32_t subset[2] = {0, 1};
    const double defaults[1] = {1.0};
    const int32_t scaling[5] = {0, 0, 1, 1, 1};     /* global parameter 0 on subsets (0,1), Coulomb and LJ */
    const int32_t derivs[1] = {0};
    snb_desc d;
    base_desc(&d, 2, 2, SNB_NO_CUTOFF);
    d.charge = q; d.sigma = sig; d.epsilon = eps; d.subset = subset;
    d.num_global_parameters = 1; d.global_parameter_defaults = defaults;
    d.num_scaling_parameters = 1; d.scaling_parameters = scaling;
    d.num_derivatives = 1; d.derivative_parameters = derivs;
    snb_handle* h = NULL;
    if (snb_create(&d, &h) != SNB_OK) {
        printf("FAIL create: %s\n", snb_last_error());
        failures++;
        return;
    }
    const double r = 0.41, lambda = 0.37;
    const double pos[6] = {0.1, 0.2, 0.3, 0.1+r, 0.2, 0.3};
    const double globals[1] = {lambda};
    double forces[6] = {1.0, 2.0, 3.0, 0.0, 0.0, 0.0};     /* forces are ADDED to what is there */
    double energy = -1, slices[3*2], deriv[1] = {10.0};    /* derivatives are added, too */
    const int flags = SNB_INCLUDE_FORCES|SNB_INCLUDE_ENERGY|SNB_INCLUDE_DIRECT|SNB_INCLUDE_RECIPROCAL;
    if (snb_execute(h, pos, NULL, globals, flags, forces, &energy, slices, deriv) != SNB_OK) {
        printf("FAIL execute: %s\n", snb_last_error());
        failures++;
        snb_destroy(h);
        return;
    }
    const double s = 0.5*(sig[0]+sig[1]), e = sqrt(eps[0]*eps[1]);
    const double s6 = pow(s/r, 6.0);
    const double ec = SNB_ONE_4PI_EPS0*q[0]*q[1]/r, ev = 4.0*e*(s6*s6-s6);
    const double dEdr = -SNB_ONE_4PI_EPS0*q[0]*q[1]/(r*r)-4.0*e*(12.0*s6*s6-6.0*s6)/r;
    const int slice = snb_slice_index(0, 1);
    expect("energy", lambda*(ec+ev), energy, 1e-6);
    expect("slice (0,1) Coulomb", ec, slices[2*slice], 1e-6);
    expect("slice (0,1) vdW", ev, slices[2*slice+1], 1e-6);
    expect("slice (0,0) Coulomb", 0.0, slices[0], 1e-6);
    expect("slice (1,1) vdW", 0.0, slices[2*snb_slice_index(1, 1)+1], 1e-6);
    expect("dE/dlambda (added to 10)", 10.0+ec+ev, deriv[0], 1e-6);
    expect("F0x (added to 1)", 1.0+lambda*dEdr, forces[0], 1e-5);     /* particle 0 sits at the smaller x */
    expect("F0y (added to 2)", 2.0, forces[1], 1e-5);
    expect("F0z (added to 3)", 3.0, forces[2], 1e-5);
    expect("F1x", -lambda*dEdr, forces[3], 1e-5);
    /* derivatives regardless of includeEnergy, energy 0 without it (ReferenceOpenMMLabKernels.cpp:246-260) */
    deriv[0] = 0.0;
    energy = -1;
    snb_execute(h, pos, NULL, globals, SNB_INCLUDE_FORCES|SNB_INCLUDE_DIRECT|SNB_INCLUDE_RECIPROCAL, forces, &energy, slices, deriv);
    expect("energy without includeEnergy", 0.0, energy, 0.0);
    expect("dE/dlambda without includeEnergy", ec+ev, deriv[0], 1e-6);
    snb_destroy(h);
}

/* three particles in a periodic box, reaction-field cutoff: one pair inside the cutoff through the box face, one outside */
static void periodic_cutoff(void) {
    const double q[3] = {1.0, -1.0, 1.0}, sig[3] = {1.0, 1.0, 1.0}, eps[3] = {0.0, 0.0, 0.0};
    const int32_t subset[3] = {0, 0, 1};
    snb_desc d;
    base_desc(&d, 3, 2, SNB_CUTOFF_PERIODIC);
    d.cutoff = 1.0;
    d.reaction_field_dielectric = 50.0;
    d.charge = q; d.sigma = sig; d.epsilon = eps; d.subset = subset;
    snb_handle* h = NULL;
    if (snb_create(&d, &h) != SNB_OK) {
        printf("FAIL create (periodic): %s\n", snb_last_error());
        failures++;
        return;
    }
    /* 0-1: 0.6 nm apart through the x face of the 4 nm box; 0-2 and 1-2: beyond the cutoff */
    const double pos[9] = {0.2, 2.0, 2.0, 3.6, 2.0, 2.0, 2.0, 0.1, 2.0};
    const double box[9] = {4, 0, 0, 0, 4, 0, 0, 0, 4};
    double forces[9] = {0}, energy = 0, slices[6];
    const int flags = SNB_INCLUDE_FORCES|SNB_INCLUDE_ENERGY|SNB_INCLUDE_DIRECT|SNB_INCLUDE_RECIPROCAL;
    if (snb_execute(h, pos, box, NULL, flags, forces, &energy, slices, NULL) != SNB_OK) {
        printf("FAIL execute (periodic): %s\n", snb_last_error());
        failures++;
        snb_destroy(h);
        return;
    }
    const double rc = 1.0, r = 0.6, epsRF = 50.0;
    const double krf = (epsRF-1.0)/(2.0*epsRF+1.0)/(rc*rc*rc), crf = 3.0*epsRF/(2.0*epsRF+1.0)/rc;
    const double e01 = SNB_ONE_4PI_EPS0*q[0]*q[1]*(1.0/r+krf*r*r-crf);
    const double f01 = SNB_ONE_4PI_EPS0*q[0]*q[1]*(1.0/(r*r)-2.0*krf*r);    /* -dE/dr */
    expect("periodic energy", e01, energy, 1e-6);
    expect("periodic slice (0,0) Coulomb", e01, slices[0], 1e-6);
    expect("periodic slice (0,1) Coulomb", 0.0, slices[2*snb_slice_index(0, 1)], 1e-6);
    /* the minimum image of particle 1 seen from particle 0 is at x = -0.4: attraction pulls particle 0 towards -x */
    expect("periodic F0x", f01, forces[0], 1e-5);
    expect("periodic F1x", -f01, forces[3], 1e-5);
    expect("periodic F2", 0.0, fabs(forces[6])+fabs(forces[7])+fabs(forces[8]), 1e-9);
    /* a box smaller than twice the cutoff is refused with the reference's message (ReferenceOpenMMLabKernels.cpp:208-210) */
    const double small[9] = {1.9, 0, 0, 0, 4, 0, 0, 0, 4};
    const int status = snb_execute(h, pos, small, NULL, flags, forces, &energy, slices, NULL);
    if (status != SNB_ERR_BOX_TOO_SMALL || strstr(snb_last_error(), "twice the nonbonded cutoff") == NULL) {
        printf("FAIL box check: status %d, message '%s'\n", status, snb_last_error());
        failures++;
    }
    snb_destroy(h);
}

int main(void) {
    two_particles_no_cutoff();
    periodic_cutoff();
    if (failures) {
        printf("%d check(s) failed\n", failures);
        return 1;
    }
    printf("OK\n");
    return 0;
}
