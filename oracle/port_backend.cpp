/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle ("port" back end).
 * The product (openmm-lab_b200/) never links, imports or executes this file.
 *
 * A from-scratch restatement, in plain double-precision C++, of the arithmetic of
 * the reference's Reference platform for the SlicedNonbondedForce path.  Each
 * function names the reference lines it follows.  It is pinned (tests/) against
 *   (a) the reference's own known-answer tests (tests/TestSlicedNonbondedForce.h),
 *   (b) oracle/_ref (the reference's own sources compiled unchanged), when built,
 *   (c) golden vectors generated from (b) and committed under tests/golden/.
 *
 * Differences from the reference that do not change results:
 *   - FFT: own mixed-radix complex DFT instead of the vendored pocketfft
 *     (same unnormalised transform pair, ReferencePME.cpp:793-805);
 *   - B-spline moduli are computed once per call like the reference, but with a
 *     table of sines/cosines;
 *   - the stride bug of ReferencePME.cpp:682 (ngrid[2] used instead of ngrid[0]
 *     in the subset stride of the interpolation) is NOT reproduced; it only shows
 *     for multi-subset runs with nx != nz, which parity configs avoid.
 */
#include "backend.h"
#include "slicednb.h"
#include "snb_constants.h"

#include <cmath>
#include <complex>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace snb_oracle {

namespace {

typedef std::complex<double> cplx;
const double K_E = SNB_ONE_4PI_EPS0;
const int COUL = 0, VDW = 1;

inline int sliceOf(int a, int b) { return a > b ? a*(a+1)/2+b : b*(b+1)/2+a; }

// delta = pos[j] - pos[i]; periodic variant follows OpenMM's ReferenceForce::getDeltaRPeriodic
// (third party; call sites ReferenceSlicedLJCoulombIxn.cpp:360,446,565).
inline void delta(const double* pos, int i, int j, const PairSettings& s, bool periodic, double d[3]) {
    for (int k = 0; k < 3; k++)
        d[k] = pos[3*j+k]-pos[3*i+k];
    if (periodic) {
        double m = floor(d[2]/s.box[2][2]+0.5);
        for (int k = 0; k < 3; k++) d[k] -= s.box[2][k]*m;
        m = floor(d[1]/s.box[1][1]+0.5);
        for (int k = 0; k < 3; k++) d[k] -= s.box[1][k]*m;
        m = floor(d[0]/s.box[0][0]+0.5);
        for (int k = 0; k < 3; k++) d[k] -= s.box[0][k]*m;
    }
}

// ---- FFT -----------------------------------------------------------------------

struct Fft1d {
    int n;
    std::vector<cplx> twiddle;   // exp(-2 pi i k / n)
    std::vector<cplx> work;
    explicit Fft1d(int n) : n(n), twiddle(n), work(n) {
        for (int k = 0; k < n; k++)
            twiddle[k] = cplx(cos(2*M_PI*k/n), -sin(2*M_PI*k/n));
    }
    // out-of-place recursive decimation in time on a strided input; sign -1 forward, +1 backward
    void transform(const cplx* in, int stride, cplx* out, int len, int sign) {
        if (len == 1) {
            out[0] = in[0];
            return;
        }
        int p = 2;
        while (len%p != 0) p++;
        int m = len/p;
        for (int r = 0; r < p; r++)
            transform(in+r*stride, stride*p, out+r*m, m, sign);
        int step = n/len;       // twiddle[step*k] = exp(-2 pi i k/len)
        std::vector<cplx> tmp(p);
        for (int k = 0; k < m; k++) {
            for (int r = 0; r < p; r++) {
                cplx w = twiddle[(size_t) step*((r*k)%len)];
                if (sign > 0) w = std::conj(w);
                tmp[r] = out[r*m+k]*w;
            }
            for (int q = 0; q < p; q++) {
                cplx sum = 0;
                for (int r = 0; r < p; r++) {
                    cplx w = twiddle[(size_t) step*(((long long) r*q*m)%len)];
                    if (sign > 0) w = std::conj(w);
                    sum += tmp[r]*w;
                }
                work[q*m+k] = sum;
            }
        }
        // combine results back (work holds this level's output)
        for (int k = 0; k < len; k++)
            out[k] = work[k];
    }
};

// in-place 3-D transform of grid[nx][ny][nz], unnormalised
void fft3d(cplx* grid, const int n[3], int sign) {
    int strides[3] = {n[1]*n[2], n[2], 1};
    for (int axis = 0; axis < 3; axis++) {
        int len = n[axis];
        Fft1d plan(len);
        std::vector<cplx> line(len), result(len);
        int a1 = (axis+1)%3, a2 = (axis+2)%3;
        for (int i = 0; i < n[a1]; i++)
            for (int j = 0; j < n[a2]; j++) {
                cplx* base = grid+(size_t) i*strides[a1]+(size_t) j*strides[a2];
                for (int k = 0; k < len; k++)
                    line[k] = base[(size_t) k*strides[axis]];
                plan.transform(line.data(), 1, result.data(), len, sign);
                for (int k = 0; k < len; k++)
                    base[(size_t) k*strides[axis]] = result[k];
            }
    }
}

// ---- PME -------------------------------------------------------------------------

const int ORDER = SNB_PME_ORDER;

// Order-5 cardinal B-spline weights and derivatives at fractional offset w
// (Essmann recursion; ReferencePME.cpp:264-317).
void bspline(double w, double theta[ORDER], double dtheta[ORDER]) {
    theta[ORDER-1] = 0;
    theta[1] = w;
    theta[0] = 1-w;
    for (int k = 3; k < ORDER; k++) {
        double div = 1.0/(k-1.0);
        theta[k-1] = div*w*theta[k-2];
        for (int l = 1; l < k-1; l++)
            theta[k-l-1] = div*((w+l)*theta[k-l-2]+(k-l-w)*theta[k-l-1]);
        theta[0] = div*(1-w)*theta[0];
    }
    dtheta[0] = -theta[0];
    for (int k = 1; k < ORDER; k++)
        dtheta[k] = theta[k-1]-theta[k];
    double div = 1.0/(ORDER-1);
    theta[ORDER-1] = div*w*theta[ORDER-2];
    for (int l = 1; l < ORDER-1; l++)
        theta[ORDER-l-1] = div*((w+l)*theta[ORDER-l-2]+(ORDER-l-w)*theta[ORDER-l-1]);
    theta[0] = div*(1-w)*theta[0];
}

// |DFT of the integer-sampled order-5 spline|^2, with the < 1e-7 fix-up
// (ReferencePME.cpp:88-183).
std::vector<double> bsplineModuli(int n) {
    double theta[ORDER], dtheta[ORDER];
    bspline(0.0, theta, dtheta);
    std::vector<double> samples(std::max(n, ORDER+1), 0.0);
    for (int i = 1; i <= ORDER; i++)
        samples[i] = theta[i-1];
    std::vector<double> moduli(n);
    for (int i = 0; i < n; i++) {
        double sc = 0, ss = 0;
        for (int j = 0; j < n; j++) {
            double arg = (2.0*M_PI*i*j)/n;
            sc += samples[j]*cos(arg);
            ss += samples[j]*sin(arg);
        }
        moduli[i] = sc*sc+ss*ss;
    }
    for (int i = 0; i < n; i++)
        if (moduli[i] < 1.0e-7)
            moduli[i] = (moduli[(i-1+n)%n]+moduli[(i+1)%n])/2;
    return moduli;
}

struct AtomSpline {
    int cell[3];
    double theta[3][ORDER], dtheta[3][ORDER];
};

// One reciprocal-space pass (pme_exec / pme_exec_dpme, ReferencePME.cpp:754-871):
// spread "charges" per subset, FFT, convolve + slice energies, inverse FFT, interpolate.
// term = COUL uses the Coulomb kernel, term = VDW the dispersion kernel.
void reciprocalPass(const PairSettings& s, int term, double alpha, const int grid[3], int numAtoms,
                    const double* pos, int numSubsets, const int* subsets, const double* lambdas,
                    const std::vector<double>& charges, double* forces, double* sliceEnergies) {
    const int nx = grid[0], ny = grid[1], nz = grid[2];
    const size_t G = (size_t) nx*ny*nz;
    // reciprocal box vectors (ReferencePME.cpp:186-194)
    double det = s.box[0][0]*s.box[1][1]*s.box[2][2];
    double R[3][3] = {{0}};
    R[0][0] = s.box[1][1]*s.box[2][2]/det;
    R[1][0] = -s.box[1][0]*s.box[2][2]/det;
    R[1][1] = s.box[0][0]*s.box[2][2]/det;
    R[2][0] = (s.box[1][0]*s.box[2][1]-s.box[1][1]*s.box[2][0])/det;
    R[2][1] = -s.box[0][0]*s.box[2][1]/det;
    R[2][2] = s.box[0][0]*s.box[1][1]/det;

    // grid index, fraction and splines (ReferencePME.cpp:196-317)
    std::vector<AtomSpline> spl(numAtoms);
    for (int i = 0; i < numAtoms; i++)
        for (int d = 0; d < 3; d++) {
            double t = pos[3*i]*R[0][d]+pos[3*i+1]*R[1][d]+pos[3*i+2]*R[2][d];
            t = (t-floor(t))*grid[d];
            int ti = (int) t;
            spl[i].cell[d] = ti%grid[d];
            bspline(t-ti, spl[i].theta[d], spl[i].dtheta[d]);
        }

    // spreading (ReferencePME.cpp:320-396)
    std::vector<cplx> Q(G*numSubsets, cplx(0, 0));
    for (int i = 0; i < numAtoms; i++) {
        cplx* sub = Q.data()+G*subsets[i];
        for (int ix = 0; ix < ORDER; ix++) {
            int x = (spl[i].cell[0]+ix)%nx;
            for (int iy = 0; iy < ORDER; iy++) {
                int y = (spl[i].cell[1]+iy)%ny;
                for (int iz = 0; iz < ORDER; iz++) {
                    int z = (spl[i].cell[2]+iz)%nz;
                    sub[((size_t) x*ny+y)*nz+z] += charges[i]*spl[i].theta[0][ix]*spl[i].theta[1][iy]*spl[i].theta[2][iz];
                }
            }
        }
    }
    for (int j = 0; j < numSubsets; j++)
        fft3d(Q.data()+G*j, grid, -1);

    // convolution and slice energies (ReferencePME.cpp:400-496 Coulomb, :499-595 dispersion)
    std::vector<double> modX = bsplineModuli(nx), modY = bsplineModuli(ny), modZ = bsplineModuli(nz);
    double volume = s.box[0][0]*s.box[1][1]*s.box[2][2];
    int maxkx = (nx+1)/2, maxky = (ny+1)/2, maxkz = (nz+1)/2;
    for (int kx = 0; kx < nx; kx++) {
        double mx = kx < maxkx ? kx : kx-nx;
        double mhx = mx*R[0][0];
        for (int ky = 0; ky < ny; ky++) {
            double my = ky < maxky ? ky : ky-ny;
            double mhy = mx*R[1][0]+my*R[1][1];
            for (int kz = 0; kz < nz; kz++) {
                if (term == COUL && kx == 0 && ky == 0 && kz == 0)
                    continue;
                double mz = kz < maxkz ? kz : kz-nz;
                double mhz = mx*R[2][0]+my*R[2][1]+mz*R[2][2];
                double m2 = mhx*mhx+mhy*mhy+mhz*mhz;
                double eterm;
                if (term == COUL) {
                    double denom = m2*(M_PI*volume*modX[kx])*modY[ky]*modZ[kz];
                    eterm = K_E*exp(-(M_PI*M_PI/(alpha*alpha))*m2)/denom;
                }
                else {
                    double boxfactor = -2*M_PI*sqrt(M_PI)/(6.0*volume);
                    double denom = boxfactor/(modX[kx]*modY[ky]*modZ[kz]);
                    double m = sqrt(m2), b = (M_PI/alpha)*m;
                    eterm = (2.0*M_PI*M_PI*M_PI*sqrt(M_PI)*erfc(b)*m*m2+exp(-b*b)*(alpha*alpha*alpha-2.0*alpha*M_PI*M_PI*m2))*denom;
                }
                size_t idx = ((size_t) kx*ny+ky)*nz+kz;
                for (int j = 0; j < numSubsets; j++) {
                    cplx orig = Q[G*j+idx];
                    Q[G*j+idx] = orig*eterm;
                    sliceEnergies[2*(j*(j+3)/2)+term] += 0.5*eterm*std::norm(orig);
                    for (int i = 0; i < j; i++) {
                        cplx scaled = Q[G*i+idx];    // already multiplied by eterm
                        sliceEnergies[2*(j*(j+1)/2+i)+term] += orig.real()*scaled.real()+orig.imag()*scaled.imag();
                    }
                }
            }
        }
    }
    for (int j = 0; j < numSubsets; j++)
        fft3d(Q.data()+G*j, grid, +1);

    // interpolation (ReferencePME.cpp:598-702; subset stride nx, see file header)
    for (int i = 0; i < numAtoms; i++) {
        double fx = 0, fy = 0, fz = 0;
        int si = subsets[i];
        for (int ix = 0; ix < ORDER; ix++) {
            int x = (spl[i].cell[0]+ix)%nx;
            for (int iy = 0; iy < ORDER; iy++) {
                int y = (spl[i].cell[1]+iy)%ny;
                for (int iz = 0; iz < ORDER; iz++) {
                    int z = (spl[i].cell[2]+iz)%nz;
                    size_t idx = ((size_t) x*ny+y)*nz+z;
                    for (int sj = 0; sj < numSubsets; sj++) {
                        double v = lambdas[2*sliceOf(si, sj)+term]*Q[G*sj+idx].real();
                        fx += spl[i].dtheta[0][ix]*spl[i].theta[1][iy]*spl[i].theta[2][iz]*v;
                        fy += spl[i].theta[0][ix]*spl[i].dtheta[1][iy]*spl[i].theta[2][iz]*v;
                        fz += spl[i].theta[0][ix]*spl[i].theta[1][iy]*spl[i].dtheta[2][iz]*v;
                    }
                }
            }
        }
        double q = charges[i];
        forces[3*i] -= q*(fx*nx*R[0][0]);
        forces[3*i+1] -= q*(fx*nx*R[1][0]+fy*ny*R[1][1]);
        forces[3*i+2] -= q*(fx*nx*R[2][0]+fy*ny*R[2][1]+fz*nz*R[2][2]);
    }
}

// Classic Ewald reciprocal sum with per-subset structure factors
// (ReferenceSlicedLJCoulombIxn.cpp:240-342).
void ewaldReciprocal(const PairSettings& s, int numAtoms, const double* pos, int numSubsets, const int* subsets,
                     const double* params, const double* lambdas, double* forces, double* sliceEnergies) {
    int kmax = std::max(s.kmax[0], std::max(s.kmax[1], s.kmax[2]));
    if (kmax < 1)
        throw std::invalid_argument("kmax for Ewald summation < 1");
    double factorEwald = -1/(4*s.alpha*s.alpha);
    double recipCoeff = K_E*4*M_PI/(s.box[0][0]*s.box[1][1]*s.box[2][2]);
    double rbox[3] = {2*M_PI/s.box[0][0], 2*M_PI/s.box[1][1], 2*M_PI/s.box[2][2]};
    // eir[k][atom][dim] = exp(i k x_dim 2pi/L_dim)
    std::vector<cplx> eir((size_t) kmax*numAtoms*3);
    for (int i = 0; i < numAtoms; i++)
        for (int m = 0; m < 3; m++) {
            eir[((size_t) 0*numAtoms+i)*3+m] = cplx(1, 0);
            if (kmax > 1)
                eir[((size_t) 1*numAtoms+i)*3+m] = cplx(cos(pos[3*i+m]*rbox[m]), sin(pos[3*i+m]*rbox[m]));
            for (int j = 2; j < kmax; j++)
                eir[((size_t) j*numAtoms+i)*3+m] = eir[((size_t) (j-1)*numAtoms+i)*3+m]*eir[((size_t) numAtoms+i)*3+m];
        }
    std::vector<cplx> tabXY(numAtoms), tabQ(numAtoms);
    std::vector<double> cs(numSubsets), ss(numSubsets);
    int lowry = 0, lowrz = 1;
    for (int rx = 0; rx < s.kmax[0]; rx++) {
        double kx = rx*rbox[0];
        for (int ry = lowry; ry < s.kmax[1]; ry++) {
            double ky = ry*rbox[1];
            for (int n = 0; n < numAtoms; n++) {
                cplx ey = ry >= 0 ? eir[((size_t) ry*numAtoms+n)*3+1] : std::conj(eir[((size_t) (-ry)*numAtoms+n)*3+1]);
                tabXY[n] = eir[((size_t) rx*numAtoms+n)*3]*ey;
            }
            for (int rz = lowrz; rz < s.kmax[2]; rz++) {
                for (int n = 0; n < numAtoms; n++) {
                    cplx ez = rz >= 0 ? eir[((size_t) rz*numAtoms+n)*3+2] : std::conj(eir[((size_t) (-rz)*numAtoms+n)*3+2]);
                    tabQ[n] = params[3*n+2]*(tabXY[n]*ez);
                }
                std::fill(cs.begin(), cs.end(), 0.0);
                std::fill(ss.begin(), ss.end(), 0.0);
                for (int n = 0; n < numAtoms; n++) {
                    cs[subsets[n]] += tabQ[n].real();
                    ss[subsets[n]] += tabQ[n].imag();
                }
                double kz = rz*rbox[2];
                double k2 = kx*kx+ky*ky+kz*kz;
                double ak = exp(k2*factorEwald)/k2;
                for (int n = 0; n < numAtoms; n++) {
                    int i = subsets[n];
                    for (int j = 0; j < numSubsets; j++) {
                        double f = 2*recipCoeff*lambdas[2*sliceOf(i, j)+COUL]*ak*(cs[j]*tabQ[n].imag()-ss[j]*tabQ[n].real());
                        forces[3*n] += f*kx;
                        forces[3*n+1] += f*ky;
                        forces[3*n+2] += f*kz;
                    }
                }
                for (int j = 0; j < numSubsets; j++) {
                    for (int i = 0; i < j; i++)
                        sliceEnergies[2*(j*(j+1)/2+i)+COUL] += 2*recipCoeff*ak*(cs[i]*cs[j]+ss[i]*ss[j]);
                    sliceEnergies[2*(j*(j+3)/2)+COUL] += recipCoeff*ak*(cs[j]*cs[j]+ss[j]*ss[j]);
                }
                lowrz = 1-s.kmax[2];
            }
            lowry = 1-s.kmax[1];
        }
    }
}

// multiplicative-dispersion damping pieces used by LJPME direct and exclusion terms
inline double dampE(double dar2) { return 1.0-exp(-dar2)*(1.0+dar2+0.5*dar2*dar2); }
inline double dampF(double dar2) { return 1.0-exp(-dar2)*(1.0+dar2+0.5*dar2*dar2+dar2*dar2*dar2/6.0); }

// ReferenceSlicedLJCoulombIxn::calculateEwaldIxn (ReferenceSlicedLJCoulombIxn.cpp:177-491)
void ewaldFamily(const PairSettings& s, int numAtoms, const double* pos, int numSubsets, const int* subsets,
                 const double* params, const double* lambdas, const std::vector<std::set<int> >& exclusions,
                 const PairList& neighbors, double* forces, double* sliceEnergies, bool includeDirect,
                 bool includeReciprocal) {
    bool pme = s.method == SNB_PME || s.method == SNB_LJPME, ljpme = s.method == SNB_LJPME, ewald = s.method == SNB_EWALD;
    if (ljpme && s.useSwitch)
        throw std::invalid_argument("Switching cannot be used with LJPME");
    const double sqrtPi = sqrt(M_PI);
    if (includeReciprocal) {
        // self energy (:198-206)
        for (int i = 0; i < numAtoms; i++) {
            int sub = subsets[i], slice = sub*(sub+3)/2;
            double q = params[3*i+2];
            sliceEnergies[2*slice+COUL] -= K_E*q*q*s.alpha/sqrtPi;
            if (ljpme)
                sliceEnergies[2*slice+VDW] += pow(s.dispersionAlpha, 6.0)*64.0*pow(params[3*i], 6.0)*pow(params[3*i+1], 2.0)/12.0;
        }
        if (pme) {
            std::vector<double> charges(numAtoms);
            for (int i = 0; i < numAtoms; i++)
                charges[i] = params[3*i+2];
            reciprocalPass(s, COUL, s.alpha, s.grid, numAtoms, pos, numSubsets, subsets, lambdas, charges, forces, sliceEnergies);
            if (ljpme) {
                for (int i = 0; i < numAtoms; i++)
                    charges[i] = 8.0*pow(params[3*i], 3.0)*params[3*i+1];
                reciprocalPass(s, VDW, s.dispersionAlpha, s.dispersionGrid, numAtoms, pos, numSubsets, subsets, lambdas, charges, forces, sliceEnergies);
            }
        }
        else if (ewald)
            ewaldReciprocal(s, numAtoms, pos, numSubsets, subsets, params, lambdas, forces, sliceEnergies);
    }
    if (!includeDirect)
        return;

    // direct space (:351-429)
    for (const auto& pair : neighbors) {
        int ii = pair.first, jj = pair.second;
        int slice = sliceOf(subsets[ii], subsets[jj]);
        double d[3];
        delta(pos, jj, ii, s, true, d);            // d = pos[ii] - pos[jj]
        double r = sqrt(d[0]*d[0]+d[1]*d[1]+d[2]*d[2]), invR = 1.0/r;
        double sw = 1, dsw = 0;
        if (s.useSwitch && r > s.switchingDistance) {
            double t = (r-s.switchingDistance)/(s.cutoff-s.switchingDistance);
            sw = 1+t*t*t*(-10+t*(15-t*6));
            dsw = t*t*(-30+t*(60-t*30))/(s.cutoff-s.switchingDistance);
        }
        double alphaR = s.alpha*r;
        double qq = K_E*params[3*ii+2]*params[3*jj+2];
        double dEdRCoul = qq*invR*invR*invR*(erfc(alphaR)+2*alphaR*exp(-alphaR*alphaR)/sqrtPi);
        double sig = params[3*ii]+params[3*jj];
        double sig2 = (invR*sig)*(invR*sig), sig6 = sig2*sig2*sig2;
        double eps = params[3*ii+1]*params[3*jj+1];
        double dEdRvdW = sw*eps*(12.0*sig6-6.0)*sig6*invR*invR;
        double vdwEnergy = eps*(sig6-1.0)*sig6;
        if (ljpme) {
            double dar2 = (s.dispersionAlpha*r)*(s.dispersionAlpha*r);
            double invR2 = invR*invR, invR6 = invR2*invR2*invR2;
            double c6i = 8.0*pow(params[3*ii], 3.0)*params[3*ii+1];
            double c6j = 8.0*pow(params[3*jj], 3.0)*params[3*jj+1];
            double emult = c6i*c6j*invR6*dampE(dar2);
            dEdRvdW += 6.0*c6i*c6j*invR6*invR2*dampF(dar2);
            double invCut2 = 1.0/(s.cutoff*s.cutoff), invCut6 = invCut2*invCut2*invCut2;
            double s2 = sig*sig, s6 = s2*s2*s2;
            double shift = eps*(1.0-s6*invCut6)*s6*invCut6;
            double dcut2 = (s.dispersionAlpha*s.cutoff)*(s.dispersionAlpha*s.cutoff);
            shift -= c6i*c6j*invCut6*dampE(dcut2);
            vdwEnergy += emult+shift;
        }
        if (s.useSwitch) {
            dEdRvdW -= vdwEnergy*dsw*invR;
            vdwEnergy *= sw;
        }
        double factor = lambdas[2*slice+VDW]*dEdRvdW+lambdas[2*slice+COUL]*dEdRCoul;
        for (int k = 0; k < 3; k++) {
            forces[3*ii+k] += factor*d[k];
            forces[3*jj+k] -= factor*d[k];
        }
        sliceEnergies[2*slice+VDW] += vdwEnergy;
        sliceEnergies[2*slice+COUL] += qq*invR*erfc(alphaR);
    }

    // exclusion corrections (:433-490)
    const double twoOverSqrtPi = 2/sqrtPi;
    for (int i = 0; i < numAtoms; i++)
        for (int j : exclusions[i]) {
            if (j <= i)
                continue;
            int slice = sliceOf(subsets[i], subsets[j]);
            double d[3];
            delta(pos, j, i, s, s.periodicExceptions, d);
            double r = sqrt(d[0]*d[0]+d[1]*d[1]+d[2]*d[2]), invR = 1.0/r;
            double alphaR = s.alpha*r;
            double qq = K_E*params[3*i+2]*params[3*j+2];
            if (erf(alphaR) > 1e-6) {
                double dEdR = qq*invR*invR*invR*(erf(alphaR)-2*alphaR*exp(-alphaR*alphaR)/sqrtPi);
                double factor = lambdas[2*slice+COUL]*dEdR;
                for (int k = 0; k < 3; k++) {
                    forces[3*i+k] -= factor*d[k];
                    forces[3*j+k] += factor*d[k];
                }
                sliceEnergies[2*slice+COUL] -= qq*invR*erf(alphaR);
            }
            else
                sliceEnergies[2*slice+COUL] -= s.alpha*twoOverSqrtPi*qq;
            if (ljpme) {
                double dar2 = (s.dispersionAlpha*r)*(s.dispersionAlpha*r);
                double invR2 = invR*invR, invR6 = invR2*invR2*invR2;
                double c6i = 8.0*pow(params[3*i], 3.0)*params[3*i+1];
                double c6j = 8.0*pow(params[3*j], 3.0)*params[3*j+1];
                sliceEnergies[2*slice+VDW] += c6i*c6j*invR6*dampE(dar2);
                double dEdR = -6.0*c6i*c6j*invR6*invR2*dampF(dar2);
                double factor = lambdas[2*slice+VDW]*dEdR;
                for (int k = 0; k < 3; k++) {
                    forces[3*i+k] -= factor*d[k];
                    forces[3*j+k] += factor*d[k];
                }
            }
        }
}

// ReferenceSlicedLJCoulombIxn::calculateOneIxn (ReferenceSlicedLJCoulombIxn.cpp:553-611)
void oneIxn(const PairSettings& s, int ii, int jj, const double* pos, const int* subsets, const double* params,
            const double* lambdas, double* forces, double* sliceEnergies) {
    bool cutoff = s.method != SNB_NO_CUTOFF, periodic = s.method == SNB_CUTOFF_PERIODIC;
    int slice = sliceOf(subsets[ii], subsets[jj]);
    double d[3];
    delta(pos, jj, ii, s, periodic, d);
    double r2 = d[0]*d[0]+d[1]*d[1]+d[2]*d[2], r = sqrt(r2), invR = 1.0/r;
    double sw = 1, dsw = 0;
    if (s.useSwitch && r > s.switchingDistance) {
        double t = (r-s.switchingDistance)/(s.cutoff-s.switchingDistance);
        sw = 1+t*t*t*(-10+t*(15-t*6));
        dsw = t*t*(-30+t*(60-t*30))/(s.cutoff-s.switchingDistance);
    }
    double sig = params[3*ii]+params[3*jj];
    double sig2 = (invR*sig)*(invR*sig), sig6 = sig2*sig2*sig2;
    double eps = params[3*ii+1]*params[3*jj+1];
    double dEdRvdW = sw*eps*(12.0*sig6-6.0)*sig6*invR*invR;
    double qq = K_E*params[3*ii+2]*params[3*jj+2];
    double krf = pow(s.cutoff, -3.0)*(s.rfDielectric-1.0)/(2.0*s.rfDielectric+1.0);
    double crf = (1.0/s.cutoff)*(3.0*s.rfDielectric)/(2.0*s.rfDielectric+1.0);
    double dEdRCoul = invR*invR*qq*(cutoff ? invR-2.0*krf*r2 : invR);
    double energy = eps*(sig6-1.0)*sig6;
    if (s.useSwitch) {
        dEdRvdW -= energy*dsw*invR;
        energy *= sw;
    }
    sliceEnergies[2*slice+VDW] += energy;
    sliceEnergies[2*slice+COUL] += cutoff ? qq*(invR+krf*r2-crf) : qq*invR;
    double factor = lambdas[2*slice+VDW]*dEdRvdW+lambdas[2*slice+COUL]*dEdRCoul;
    for (int k = 0; k < 3; k++) {
        forces[3*ii+k] += factor*d[k];
        forces[3*jj+k] -= factor*d[k];
    }
}

}  // namespace

void backendPairIxn(const PairSettings& s, int numAtoms, const double* pos, int numSubsets,
                    const int* subsets, const double* params, const double* lambdas,
                    const std::vector<std::set<int> >& exclusions, const PairList& neighbors,
                    double* forces, double* sliceEnergies, bool includeDirect, bool includeReciprocal) {
    // ReferenceSlicedLJCoulombIxn::calculatePairIxn (:512-536)
    if (s.method >= SNB_EWALD) {
        ewaldFamily(s, numAtoms, pos, numSubsets, subsets, params, lambdas, exclusions, neighbors, forces,
                    sliceEnergies, includeDirect, includeReciprocal);
        return;
    }
    if (!includeDirect)
        return;
    if (s.method != SNB_NO_CUTOFF) {
        for (const auto& pair : neighbors)
            oneIxn(s, pair.first, pair.second, pos, subsets, params, lambdas, forces, sliceEnergies);
    }
    else {
        for (int ii = 0; ii < numAtoms; ii++)
            for (int jj = ii+1; jj < numAtoms; jj++)
                if (exclusions[jj].find(ii) == exclusions[jj].end())
                    oneIxn(s, ii, jj, pos, subsets, params, lambdas, forces, sliceEnergies);
    }
}

void backend14Ixn(const PairSettings& s, bool periodic, int i, int j, const double* pos,
                  const double* params, double* forces, const double* lambdas, double* sliceEnergies) {
    // ReferenceSlicedLJCoulomb14::calculateBondIxn (ReferenceSlicedLJCoulomb14.cpp:61-95)
    double d[3];
    delta(pos, j, i, s, periodic, d);                // d = pos[i] - pos[j]
    double invR = 1.0/sqrt(d[0]*d[0]+d[1]*d[1]+d[2]*d[2]);
    double sig2 = (invR*params[0])*(invR*params[0]), sig6 = sig2*sig2*sig2;
    double dEdR = lambdas[VDW]*params[1]*(12.0*sig6-6.0)*sig6;
    dEdR += lambdas[COUL]*K_E*params[2]*invR;
    dEdR *= invR*invR;
    for (int k = 0; k < 3; k++) {
        forces[3*i+k] += dEdR*d[k];
        forces[3*j+k] -= dEdR*d[k];
    }
    sliceEnergies[VDW] += params[1]*(sig6-1.0)*sig6;
    sliceEnergies[COUL] += K_E*params[2]*invR;
}

const char* backendName() { return "port"; }

}  // namespace snb_oracle
