// Device cost functors psi(x): the reference's factor cost functions as
// plain structs evaluated inside the fused sigma-point kernel.  Each functor provides
//   XD                 number of leading coordinates of x it reads (rows of sqrt(Sigma) needed)
//   Pending            state carried between the two halves of an evaluation
//   begin<FAST>(x, f)  first half: everything up to (and including) issuing the long-latency loads
//   finish(pending)    second half: psi up to the constant factor scale()
//   fast_ok(lo, hi)    per-factor, warp-uniform: may the FAST variant be used for sigma points whose
//                      leading coordinates lie in the box [lo, hi]?  (FAST must give identical results)
//   scale()            constant folded into the epilogue (e.g. the hinge weight sigma)
//   all_zero(lo, hi)   per-factor: is psi PROVABLY zero at every point of the box [lo, hi]?  (conservative: false when
//                      in doubt.)  A factor whose sigma-point box passes has moments that are exactly zero, so the
//                      sign-group kernel does not evaluate it at all (free-space culling; results are bit-identical).
// The split lets the kernel keep the gather of node i+32 in flight while it finishes node i.
#pragma once
#include <cuda_runtime.h>

namespace gvib200 {

// functors without long-latency loads evaluate everything in begin()
#define GVIB200_SIMPLE_FUNCTOR_INTERFACE()                                                        \
    static constexpr bool PREMAP = false;                                                         \
    static constexpr bool CULL = false;                                                           \
    struct Pending {                                                                              \
        double psi;                                                                               \
    };                                                                                            \
    template <bool FAST>                                                                          \
    __device__ __forceinline__ Pending begin(const double* x, int f) const {                     \
        Pending p;                                                                                \
        p.psi = eval(x, f);                                                                       \
        return p;                                                                                 \
    }                                                                                             \
    __device__ __forceinline__ double finish(const Pending& p) const { return p.psi; }            \
    __device__ __forceinline__ bool fast_ok(const double*, const double*) const { return false; } \
    __device__ __forceinline__ bool all_zero(const double*, const double*) const { return false; }

// src/1d_example.cpp:25-35 (and tests/test_GH.cpp:21-34 with y_offset = +0.05)
struct CostStereo1D {
    static constexpr int XD = 1;
    GVIB200_SIMPLE_FUNCTOR_INTERFACE()
    double mu_p, fb, sig_p_sq, sig_r_sq, y;  // fb = f*b, y = f*b/mu_p + y_offset
    __device__ __forceinline__ double eval(const double* x, int) const {
        const double a = x[0] - mu_p;
        const double r = y - fb / x[0];
        // (x-mu_p)^2 / sig_p_sq / 2 + (y - f b / x)^2 / sig_r_sq / 2, same operation order
        return a * a / sig_p_sq / 2.0 + r * r / sig_r_sq / 2.0;
    }
    __device__ __forceinline__ double scale() const { return 1.0; }
};

// CudaOperation_PlanarPR::cost_obstacle_planar (helpers/CudaOperation.h:491-508, n_balls = 1,
// slope = 1) over PlanarSDF::getSignedDistance (:51-103, :123-125):
//   psi(x) = sigma * max(0, eps + r - sd(x0, x1))^2,   sd = bilinear lookup, point clamped to the field.
// The field is stored as one 32-byte record per cell (upper indices clamped: the reference reads one past the edge with
// weight exactly 0 there), so a lookup is a single sector instead of four scattered doubles.  The record holds the
// bilinear form already arranged for the hinge: with v00 = v(r,c), v10 = v(r+1,c), v01 = v(r,c+1), v11 = v(r+1,c+1)
//   { thr - v00,  -(v10 - v00),  -(v01 - v00),  -((v11 - v01) - (v10 - v00)) }
// so that  thr - sd = t0 + fr * n_dr + fc * (n_dc + fr * n_drc)  is three FMAs.
struct CostPlanarHinge {
    static constexpr int XD = 2;
    static constexpr bool CULL = true;
    const double4* __restrict__ rec;  // [cols][rows]
    int rows, cols;
    double ox, oy, xmax, ymax, inv_cell, thr, sigma;
    double cx0, cy0;  // -ox*inv_cell, -oy*inv_cell
    struct Pending {
        double4 v;
        double fc, fr;
    };
    // FAST: the factor's whole sigma-point box lies inside the field, so the clamp is the identity
    __device__ __forceinline__ bool fast_ok(const double* lo, const double* hi) const {
        // one cell of margin keeps the cell coordinate safely positive / below the last node under rounding
        const double cell = 1.0 / inv_cell;
        return lo[0] >= ox + cell && hi[0] <= xmax - cell && lo[1] >= oy + cell && hi[1] <= ymax - cell;
    }
    // Free-space culling.  coarse[R + C * crows] = max over the 4 x 4 cells of block (R, C) of (thr - min of the cell's
    // four corner distances): bilinear interpolation stays between its corners, so thr - sd <= that bound anywhere in
    // the block.  A box whose blocks (one cell of slack on every side for rounding in the cell coordinates) are all
    // below -1e-9 has psi = max(0, thr - sd)^2 == 0 at every point: the hinge is exactly zero there.
    const double* __restrict__ coarse = nullptr;
    int crows = 0;
    __device__ __forceinline__ bool all_zero(const double* lo, const double* hi) const {
        if (coarse == nullptr) return false;
        const double c0 = fmin(fmax(fma(lo[0], inv_cell, cx0), 0.0), (double)(cols - 1));
        const double c1 = fmin(fmax(fma(hi[0], inv_cell, cx0), 0.0), (double)(cols - 1));
        const double r0 = fmin(fmax(fma(lo[1], inv_cell, cy0), 0.0), (double)(rows - 1));
        const double r1 = fmin(fmax(fma(hi[1], inv_cell, cy0), 0.0), (double)(rows - 1));
        if (!(c1 >= c0) || !(r1 >= r0)) return false;  // NaN / inverted box
        const int C0 = max((int)c0 - 1, 0) >> 2, C1 = min((int)c1 + 1, cols - 1) >> 2;
        const int R0 = max((int)r0 - 1, 0) >> 2, R1 = min((int)r1 + 1, rows - 1) >> 2;
        if ((C1 - C0 + 1) * (R1 - R0 + 1) > 64) return false;  // a very wide factor: not worth the scan
        double m = -1.0;
        for (int C = C0; C <= C1; ++C)
            for (int R = R0; R <= R1; ++R) m = fmax(m, __ldg(coarse + R + C * crows));
        return m < -1e-9;
    }
    // PREMAP: the sign-group kernel applies the affine map x -> cell coordinates once per factor (to mu and to the
    // columns of S) instead of once per sigma point, and hands cell coordinates to begin_mapped()
    static constexpr bool PREMAP = true;
    __device__ __forceinline__ double pre_mu(int r, double v) const { return fma(v, inv_cell, r == 0 ? cx0 : cy0); }
    __device__ __forceinline__ double pre_scale(int) const { return inv_cell; }
    template <bool FAST>
    __device__ __forceinline__ Pending begin(const double* x, int f) const {
        double xin = x[0], yin = x[1];
        if (!FAST) {
            xin = fmin(fmax(xin, ox), xmax);
            yin = fmin(fmax(yin, oy), ymax);
        }
        const double cr[2] = {fma(xin, inv_cell, cx0), fma(yin, inv_cell, cy0)};
        return begin_mapped<FAST>(cr, f);
    }
    template <bool FAST>
    __device__ __forceinline__ Pending begin_mapped(const double* cr, int) const {
        double col = cr[0], row = cr[1];
        if (!FAST) {  // clamp to the field (convertPoint2toCell, helpers/CudaOperation.h:61-81) in cell coordinates
            col = fmin(fmax(col, 0.0), (double)(cols - 1));
            row = fmin(fmax(row, 0.0), (double)(rows - 1));
        }
        // floor() by adding 1.5*2^52 with round-toward-minus-infinity (one DADD.RM); the low word of the
        // sum is the integer cell index, the difference back is floor(col) as a double.
        const double M = 6755399441055744.0;
        const double uc = __dadd_rd(col, M);
        const double ur = __dadd_rd(row, M);
        const unsigned lci = (unsigned)__double2loint(uc);
        const unsigned lri = (unsigned)__double2loint(ur);
        Pending p;
        p.fc = col - (uc - M);
        p.fr = row - (ur - M);
        // one 256-bit read-only load (SASS LDG.E.256.CONSTANT; records are 32-byte aligned)
        asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
                     : "=d"(p.v.x), "=d"(p.v.y), "=d"(p.v.z), "=d"(p.v.w)
                     : "l"(rec + (lci * (unsigned)rows + lri)));  // 32-bit cell index: one IMAD + one IMAD.WIDE
        return p;
    }
    __device__ __forceinline__ double finish(const Pending& p) const {
        // max(t, 0)^2 as (t + |t|)^2 / 4: t + |t| is exactly 2t or 0, so the square is exactly 4 max(t,0)^2 and the
        // factor 1/4 (a power of two) folds into scale() without changing a bit; two FP64 instructions instead of a
        // compare + select sequence.
        const double t = fma(p.fc, fma(p.fr, p.v.w, p.v.z), fma(p.fr, p.v.y, p.v.x));  // thr - sd
        const double u = t + fabs(t);
        return u * u;
    }
    __device__ __forceinline__ double scale() const { return 0.25 * sigma; }
};

// CudaOperation_Quad::cost_obstacle_planar (helpers/CudaOperation.h:565-605): a planar quadrotor x = (pos_x, pos_z, phi, ...)
// carries n_balls = 5 check points along its body axis (L = 5),
//   l = pos - (L - 1.5 r) (cos phi, sin phi) / 2,   pt_i = l + L (cos phi, sin phi) i / 5,   i = 0..4,
// each looked up in the PlanarSDF and hinged with slope 5:  psi = sigma * sum_i (5 max(0, eps + r - sd_i))^2.
// Reuses the 32-byte hinge records of CostPlanarHinge (thr - sd in three FMAs); the five points are clamped to the field
// one by one, as convertPoint2toCell does.
struct CostQuadHinge {
    static constexpr int XD = 3;
    static constexpr bool PREMAP = false;
    static constexpr bool CULL = false;
    CostPlanarHinge h;  // field, threshold, sigma
    double radius;
    struct Pending {
        CostPlanarHinge::Pending b[5];
    };
    __device__ __forceinline__ bool fast_ok(const double*, const double*) const { return false; }
    __device__ __forceinline__ bool all_zero(const double*, const double*) const { return false; }
    template <bool FAST>
    __device__ __forceinline__ Pending begin(const double* x, int f) const {
        constexpr double L = 5.0;
        double sn, cs;
        sincos(x[2], &sn, &cs);
        const double lx = x[0] - (L - radius * 1.5) * cs / 2.0;
        const double lz = x[1] - (L - radius * 1.5) * sn / 2.0;
        Pending p;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const double pt[2] = {lx + L * cs / 5 * i, lz + L * sn / 5 * i};
            p.b[i] = h.template begin<false>(pt, f);
        }
        return p;
    }
    __device__ __forceinline__ double finish(const Pending& p) const {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) s += h.finish(p.b[i]);
        return s;
    }
    __device__ __forceinline__ double scale() const { return 25.0 * h.scale(); }  // slope^2 = 25
};

// CudaOperation_3dpR::cost_obstacle_planar (helpers/CudaOperation.h:641-674, n_balls = 1, slope = 1) over the 3-D
// SignedDistanceField (:133-236): psi(x) = sigma * max(0, eps + r - sd(x0, x1, x2))^2 with the trilinear lookup of
// :219-236 on data[r + c * rows + z * rows * cols] (:299-301), the point clamped to the field (:176-205).
// Upper indices are clamped (the reference reads one past the edge with weight exactly 0 there).
struct CostHinge3D {
    static constexpr int XD = 3;
    static constexpr bool PREMAP = false;
    static constexpr bool CULL = false;
    const double* __restrict__ data;
    int rows, cols, nz;
    double ox, oy, oz, xmax, ymax, zmax, inv_cell, thr, sigma;
    struct Pending {
        double v[8];
        double fr, fc, fz;
    };
    __device__ __forceinline__ bool fast_ok(const double* lo, const double* hi) const {
        const double cell = 1.0 / inv_cell;
        return lo[0] >= ox + cell && hi[0] <= xmax - cell && lo[1] >= oy + cell && hi[1] <= ymax - cell &&
               lo[2] >= oz + cell && hi[2] <= zmax - cell;
    }
    __device__ __forceinline__ bool all_zero(const double*, const double*) const { return false; }
    template <bool FAST>
    __device__ __forceinline__ Pending begin(const double* x, int) const {
        double xin = x[0], yin = x[1], zin = x[2];
        if (!FAST) {
            xin = fmin(fmax(xin, ox), xmax);
            yin = fmin(fmax(yin, oy), ymax);
            zin = fmin(fmax(zin, oz), zmax);
        }
        const double col = (xin - ox) * inv_cell, row = (yin - oy) * inv_cell, zz = (zin - oz) * inv_cell;
        const double lr = floor(row), lc = floor(col), lz = floor(zz);
        const int lri = (int)lr, lci = (int)lc, lzi = (int)lz;
        const int hri = min(lri + 1, rows - 1), hci = min(lci + 1, cols - 1), hzi = min(lzi + 1, nz - 1);
        Pending p;
        p.fr = row - lr;
        p.fc = col - lc;
        p.fz = zz - lz;
        const size_t sl = (size_t)rows * cols;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = (k & 1) ? hri : lri, c = (k & 2) ? hci : lci, z = (k & 4) ? hzi : lzi;
            p.v[k] = __ldg(data + r + (size_t)c * rows + (size_t)z * sl);
        }
        return p;
    }
    __device__ __forceinline__ double finish(const Pending& p) const {
        // trilinear: along rows, then columns, then z
        const double c00 = fma(p.fr, p.v[1] - p.v[0], p.v[0]), c10 = fma(p.fr, p.v[3] - p.v[2], p.v[2]);
        const double c01 = fma(p.fr, p.v[5] - p.v[4], p.v[4]), c11 = fma(p.fr, p.v[7] - p.v[6], p.v[6]);
        const double c0 = fma(p.fc, c10 - c00, c00), c1 = fma(p.fc, c11 - c01, c01);
        const double sd = fma(p.fz, c1 - c0, c0);
        const double t = thr - sd;
        const double u = t + fabs(t);
        return u * u;
    }
    __device__ __forceinline__ double scale() const { return 0.25 * sigma; }
};

// CudaOperation_3dArm::cost_obstacle (helpers/CudaOperation.h:751-770) with ForwardKinematics (:325-410): a serial arm
// described by Denavit-Hartenberg parameters (a, alpha, d, theta_bias per joint) carries body spheres (frame, centre in
// that frame, radius); the state is (joint angles, joint velocities), the cost the sum over spheres of
// sigma * max(0, eps + r_i - sd(p_i))^2 with p_i = T_0 ... T_frame(i) [centre_i; 1] and sd the trilinear lookup of the
// 3-D SignedDistanceField (same lookup as CostHinge3D).  Two quirks of the reference are mirrored:
//  * n_balls = theta.size() (:752): the number of spheres evaluated is the dimension of the factor's state vector (angles
//    AND velocities), not the number of spheres the arm was given -- here min(that, n_spheres) so nothing is read out of
//    bounds;
//  * dh_matrix (:394-400) calls cosf / sinf: single-precision trigonometry of the double argument.  Mirrored as the
//    correctly rounded single-precision value, float(cos(double(float(theta)))): this is what an exact cosf returns, and it
//    is the same number on the device and in the CPU oracle (library cosf implementations differ in the last ulp).
constexpr int ARM_MAX_DOF = 3;
constexpr int ARM_MAX_SPHERES = 12;
template <int NDOF>
struct CostArm3D {
    static constexpr int XD = 2 * NDOF;
    GVIB200_SIMPLE_FUNCTOR_INTERFACE()
    CostHinge3D field;  // distance field; its thr / sigma members are not used
    double sigma, epsilon;
    int n_spheres;
    double a[ARM_MAX_DOF], ca[ARM_MAX_DOF], sa[ARM_MAX_DOF], d[ARM_MAX_DOF], bias[ARM_MAX_DOF];  // ca / sa: cosf / sinf of alpha
    int frame[ARM_MAX_SPHERES];
    double centre[ARM_MAX_SPHERES][3], radius[ARM_MAX_SPHERES];
    static __device__ __forceinline__ double f32(double v) { return (double)(float)v; }
    __device__ __forceinline__ double eval(const double* x, int f) const {
        // cumulative transforms T_0 ... T_j, rows 0..2 of the 4x4 (the last row stays (0, 0, 0, 1))
        double T[NDOF][12];
        double cur[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
#pragma unroll
        for (int j = 0; j < NDOF; ++j) {
            const double th = f32(x[j] + bias[j]);
            const double ct = f32(cos(th)), st = f32(sin(th));
            // dh = [[ct, -st ca, st sa, a ct], [st, ct ca, -ct sa, a st], [0, sa, ca, d], [0, 0, 0, 1]]
            const double m[12] = {ct, -st * ca[j], st * sa[j], a[j] * ct, st, ct * ca[j], -ct * sa[j], a[j] * st, 0.0, sa[j], ca[j], d[j]};
            double nx[12];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    double v = cur[4 * r + 0] * m[c] + cur[4 * r + 1] * m[4 + c] + cur[4 * r + 2] * m[8 + c];
                    if (c == 3) v += cur[4 * r + 3];
                    nx[4 * r + c] = v;
                }
            }
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                cur[e] = nx[e];
                T[j][e] = nx[e];
            }
        }
        const int nb = min(2 * NDOF, n_spheres);
        double cost = 0.0;
        for (int i = 0; i < nb; ++i) {
            const int fr = frame[i];
            double pt[3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
                pt[r] = T[fr][4 * r + 3] + (T[fr][4 * r + 0] * centre[i][0] + T[fr][4 * r + 1] * centre[i][1] + T[fr][4 * r + 2] * centre[i][2]);
            const CostHinge3D::Pending p = field.template begin<false>(pt, f);
            const double fr_ = p.fr, fc = p.fc, fz = p.fz;
            const double c00 = fma(fr_, p.v[1] - p.v[0], p.v[0]), c10 = fma(fr_, p.v[3] - p.v[2], p.v[2]);
            const double c01 = fma(fr_, p.v[5] - p.v[4], p.v[4]), c11 = fma(fr_, p.v[7] - p.v[6], p.v[6]);
            const double c0 = fma(fc, c10 - c00, c00), c1 = fma(fc, c11 - c01, c01);
            const double sd = fma(fz, c1 - c0, c0);
            const double t = epsilon + radius[i] - sd;
            if (t >= 0.0) cost = fma(t, t, cost);  // sd > eps + r: no contribution (:760-763)
        }
        return cost;
    }
    __device__ __forceinline__ double scale() const { return sigma; }
};

// cost_linear_gp (gp/cost_functions.h:36-39 -> MinimumAccGP::cost gp/minimum_acc_prior.h:103-106,
// LTV_GP::cost gp/LTV_prior.h:217-220): 1/2 (Phi th1 - th2)^T Qinv (Phi th1 - th2).
// Per-factor parameters: Phi[DS*DS], Qinv[DS*DS] column-major.
template <int DS>
struct CostLinearGP {
    static constexpr int XD = 2 * DS;
    GVIB200_SIMPLE_FUNCTOR_INTERFACE()
    const double* __restrict__ params;  // [n][2*DS*DS]
    __device__ __forceinline__ double eval(const double* x, int f) const {
        const double* Phi = params + (size_t)f * 2 * DS * DS;
        const double* Qi = Phi + DS * DS;
        double r[DS];
#pragma unroll
        for (int i = 0; i < DS; ++i) {
            double s = -x[DS + i];
#pragma unroll
            for (int k = 0; k < DS; ++k) s = fma(__ldg(Phi + i + k * DS), x[k], s);
            r[i] = s;
        }
        double q = 0.0;
#pragma unroll
        for (int j = 0; j < DS; ++j) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < DS; ++i) t = fma(__ldg(Qi + i + j * DS), r[i], t);
            q = fma(t, r[j], q);
        }
        return q;
    }
    __device__ __forceinline__ double scale() const { return 0.5; }
};

// cost_fixed_gp (gp/cost_functions.h:25-27 -> FixedPriorGP::fixed_factor_cost gp/fixed_prior.h:28-30):
// (x - mu)^T Kinv (x - mu).  Per-factor parameters: Kinv[DIM*DIM], mu[DIM].
template <int DIM>
struct CostFixedGP {
    static constexpr int XD = DIM;
    GVIB200_SIMPLE_FUNCTOR_INTERFACE()
    const double* __restrict__ params;  // [n][DIM*DIM + DIM]
    __device__ __forceinline__ double eval(const double* x, int f) const {
        const double* Ki = params + (size_t)f * (DIM * DIM + DIM);
        const double* mu = Ki + DIM * DIM;
        double r[DIM];
#pragma unroll
        for (int i = 0; i < DIM; ++i) r[i] = x[i] - __ldg(mu + i);
        double q = 0.0;
#pragma unroll
        for (int j = 0; j < DIM; ++j) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < DIM; ++i) t = fma(__ldg(Ki + i + j * DIM), r[i], t);
            q = fma(t, r[j], q);
        }
        return q;
    }
    __device__ __forceinline__ double scale() const { return 1.0; }
};

// x^T (c I) x: the integrand gx_1d of tests/test_gh_spgh.cpp:21-25 (known-answer tests).
template <int DIM>
struct CostQuadratic {
    static constexpr int XD = DIM;
    GVIB200_SIMPLE_FUNCTOR_INTERFACE()
    double c;
    __device__ __forceinline__ double eval(const double* x, int) const {
        double q = 0.0;
#pragma unroll
        for (int i = 0; i < DIM; ++i) q = fma(x[i], x[i], q);
        return q;
    }
    __device__ __forceinline__ double scale() const { return c; }
};

}  // namespace gvib200
