// Fixed-size dense helpers for the d x d blocks of the chain (d = state dim) and the
// dim x dim factor marginals.  Column-major (Eigen's default), everything unrolled so the
// matrices live in registers.  __host__ __device__ so the host test harness
// (tests/cpp/bt_host_emu.cpp) exercises exactly the arithmetic the kernels run.
#pragma once

#ifdef __CUDACC__
#define GVI_HD __host__ __device__ __forceinline__
#else
#define GVI_HD inline
#endif

#include <math.h>

namespace gvib200 {

template <int N>
struct Mat {
    double a[N * N];
    GVI_HD double& operator()(int i, int j) { return a[i + j * N]; }
    GVI_HD double operator()(int i, int j) const { return a[i + j * N]; }
};

template <int N>
struct Vec {
    double a[N];
    GVI_HD double& operator()(int i) { return a[i]; }
    GVI_HD double operator()(int i) const { return a[i]; }
};

template <int N>
GVI_HD void mat_zero(Mat<N>& A) {
#pragma unroll
    for (int i = 0; i < N * N; ++i) A.a[i] = 0.0;
}

template <int N>
GVI_HD void vec_zero(Vec<N>& v) {
#pragma unroll
    for (int i = 0; i < N; ++i) v.a[i] = 0.0;
}

template <int N>
GVI_HD void mat_load(Mat<N>& A, const double* __restrict__ p) {
#pragma unroll
    for (int i = 0; i < N * N; ++i) A.a[i] = p[i];
}

template <int N>
GVI_HD void mat_store(double* __restrict__ p, const Mat<N>& A) {
#pragma unroll
    for (int i = 0; i < N * N; ++i) p[i] = A.a[i];
}

template <int N>
GVI_HD void vec_load(Vec<N>& v, const double* __restrict__ p) {
#pragma unroll
    for (int i = 0; i < N; ++i) v.a[i] = p[i];
}

template <int N>
GVI_HD void vec_store(double* __restrict__ p, const Vec<N>& v) {
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = v.a[i];
}

// C = A * B
template <int N>
GVI_HD void mm(Mat<N>& C, const Mat<N>& A, const Mat<N>& B) {
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) s = fma(A(i, k), B(k, j), s);
            C(i, j) = s;
        }
}

// C = A^T * B
template <int N>
GVI_HD void mtm(Mat<N>& C, const Mat<N>& A, const Mat<N>& B) {
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) s = fma(A(k, i), B(k, j), s);
            C(i, j) = s;
        }
}

// C = A * B^T
template <int N>
GVI_HD void mmt(Mat<N>& C, const Mat<N>& A, const Mat<N>& B) {
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) s = fma(A(i, k), B(j, k), s);
            C(i, j) = s;
        }
}

// y = A x
template <int N>
GVI_HD void mv(Vec<N>& y, const Mat<N>& A, const Vec<N>& x) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(A(i, k), x(k), s);
        y(i) = s;
    }
}

// y = A^T x
template <int N>
GVI_HD void mtv(Vec<N>& y, const Mat<N>& A, const Vec<N>& x) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(A(k, i), x(k), s);
        y(i) = s;
    }
}

template <int N>
GVI_HD void mat_transpose(Mat<N>& B, const Mat<N>& A) {
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) B(i, j) = A(j, i);
}

template <int N>
GVI_HD void symmetrize(Mat<N>& A) {
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double s = 0.5 * (A(i, j) + A(j, i));
            A(i, j) = s;
            A(j, i) = s;
        }
}

// 1/sqrt(x): the device path uses the hardware-seeded rsqrt (no DSQRT + DDIV slow paths on the
// serial critical path of the chain engine); the host harness uses the plain expression.
GVI_HD double gvi_rsqrt(double x) {
#ifdef __CUDA_ARCH__
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// Running log-determinant as (mantissa product, binary exponent): one multiply per pivot on the
// serial path instead of one log(); log() is taken once per worker at the end.
struct LogDetAcc {
    double m;
    int e;
    GVI_HD LogDetAcc() : m(1.0), e(0) {}
    GVI_HD void mul(double s) { m *= s; }
    // renormalise the mantissa into [1, 2); call at least every few multiplies (each pivot may span
    // ~1e+-70 before the product leaves the double range)
    GVI_HD void normalize() {
#ifdef __CUDA_ARCH__
        const int hi = __double2hiint(m);
        const int ex = ((hi >> 20) & 0x7ff) - 1023;
        if (m > 0.0 && ex > -1022 && ex < 1024) {
            m = __hiloint2double(hi - (ex << 20), __double2loint(m));
            e += ex;
        }
#else
        if (m > 0.0 && isfinite(m)) {
            int ex;
            m = frexp(m, &ex) * 2.0;
            e += ex - 1;
        }
#endif
    }
    GVI_HD double value() const { return log(m) + (double)e * 0.69314718055994530942; }
};

// Cholesky factor A = L L^T of an SPD matrix (only the lower triangle of A is read) with rd[j] = 1 / L_jj (rsqrt of the
// pivot: division free).  The pivots are multiplied into `ld`; returns false when a pivot is not strictly positive.
template <int N>
GVI_HD bool chol_factor(Mat<N>& L, double* rd, const Mat<N>& A, LogDetAcc& ld) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double s = A(j, j);
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L(j, k), L(j, k), s);
        ok = ok && (s > 0.0);
        ld.mul(s);
        const double r = gvi_rsqrt(s);
        rd[j] = r;
        L(j, j) = s * r;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double t = A(i, j);
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-L(i, k), L(j, k), t);
            L(i, j) = t * r;
        }
    }
    ld.normalize();
    return ok;
}

// x = A^-1 b from the factor of chol_factor (forward then backward substitution)
template <int N>
GVI_HD void chol_solve(Vec<N>& x, const Mat<N>& L, const double* rd, const Vec<N>& b) {
    Vec<N> z;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double t = b(i);
#pragma unroll
        for (int k = 0; k < i; ++k) t = fma(-L(i, k), z(k), t);
        z(i) = t * rd[i];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        double t = z(i);
#pragma unroll
        for (int k = i + 1; k < N; ++k) t = fma(-L(k, i), x(k), t);
        x(i) = t * rd[i];
    }
}

// Inverse of an SPD matrix through its Cholesky factor: A = L L^T, Ainv = L^-T L^-1.
// The pivots are multiplied into `ld` (log det accumulator); returns false (and leaves Ainv
// unspecified, possibly NaN) when a pivot is not strictly positive -- the caller raises
// GVIB200_ENOTSPD.  Only the lower triangle of A is read.  Division free: every 1/L_jj is the
// rsqrt of the pivot.
template <int N>
GVI_HD bool spd_inverse(Mat<N>& Ainv, const Mat<N>& A, LogDetAcc& ld) {
    Mat<N> L;
    double rd[N];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double s = A(j, j);
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L(j, k), L(j, k), s);
        ok = ok && (s > 0.0);
        ld.mul(s);
        const double r = gvi_rsqrt(s);
        rd[j] = r;
        L(j, j) = s * r;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double t = A(i, j);
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-L(i, k), L(j, k), t);
            L(i, j) = t * r;
        }
    }
    ld.normalize();
    // M = L^-1 (lower triangular)
    Mat<N> M;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        M(j, j) = rd[j];
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double t = 0.0;
#pragma unroll
            for (int k = j; k < i; ++k) t = fma(-L(i, k), M(k, j), t);
            M(i, j) = t * rd[i];
        }
    }
    // Ainv = M^T M (symmetric)
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = j; i < N; ++i) {
            double t = 0.0;
#pragma unroll
            for (int k = i; k < N; ++k) t = fma(M(k, i), M(k, j), t);
            Ainv(i, j) = t;
            Ainv(j, i) = t;
        }
    return ok;
}

// Cyclic Jacobi eigen-decomposition of a symmetric N x N matrix (classical threshold variant).
// On return A's diagonal holds the eigenvalues (in lam) and V the orthonormal eigenvectors in its
// columns.  Used for the symmetric PSD square root S = V sqrt(lam) V^T that the reference takes with
// SelfAdjointEigenSolver::operatorSqrt() (quadrature/SparseGaussHermite.h:231-243) and for
// P_k = Sigma_k^-1 (gvibase/GVIFactorizedBase.h:111-114).
// Branch-free Jacobi rotation parameters for the pivot (app, aqq, apq): t = tan, c = cos, s = sin of the angle that
// annihilates apq.  One square root, one division, one reciprocal square root:
//   t = sign(d) b / (|d| + sqrt(d^2 + b^2)),  d = aqq - app, b = 2 apq   (the smaller root of t^2 + 2 theta t - 1 = 0)
GVI_HD void jacobi_angle(double app, double aqq, double apq, double& t, double& c, double& s) {
    const double d = aqq - app, b = 2.0 * apq;
    const bool tiny = fabs(apq) <= 1e-19 * (fabs(app) + fabs(aqq));  // the rotation would not change app / aqq at all
    const double den = fabs(d) + sqrt(fma(d, d, b * b));
    t = (tiny || den == 0.0) ? 0.0 : (d >= 0.0 ? b : -b) / den;
#ifdef __CUDA_ARCH__
    c = rsqrt(fma(t, t, 1.0));
#else
    c = 1.0 / sqrt(fma(t, t, 1.0));
#endif
    s = t * c;
}

template <int N>
GVI_HD void jacobi_rotate(Mat<N>& A, Mat<N>& V, int p, int q, double t, double c, double s) {
    const double apq = A(p, q);
    A(p, p) -= t * apq;
    A(q, q) += t * apq;
    A(p, q) = 0.0;
    A(q, p) = 0.0;
#pragma unroll
    for (int r = 0; r < N; ++r) {
        if (r != p && r != q) {
            const double arp = A(r, p), arq = A(r, q);
            const double nrp = fma(c, arp, -s * arq);
            const double nrq = fma(s, arp, c * arq);
            A(r, p) = nrp;
            A(p, r) = nrp;
            A(r, q) = nrq;
            A(q, r) = nrq;
        }
    }
#pragma unroll
    for (int r = 0; r < N; ++r) {
        const double vrp = V(r, p), vrq = V(r, q);
        V(r, p) = fma(c, vrp, -s * vrq);
        V(r, q) = fma(s, vrp, c * vrq);
    }
}

// N = 4: tournament ordering {(0,1),(2,3)}, {(0,2),(1,3)}, {(0,3),(1,2)} -- the two pivots of a round touch disjoint
// rows / columns, so their angles come from the same matrix and the two (long, division / square-root bound)
// dependency chains overlap inside one thread.  Same fixed point and convergence as the cyclic sweep.
GVI_HD void jacobi_eig4(Mat<4>& A, Mat<4>& V, Vec<4>& lam) {
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            diag += fabs(A(j, j));
#pragma unroll
            for (int i = 0; i < j; ++i) off += fabs(A(i, j));
        }
        // convergence is quadratic: a sweep entered with off <= 1e-17 diag would only move the result below rounding
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
#pragma unroll
        for (int round = 0; round < 3; ++round) {
            const int p0 = 0, q0 = round + 1;
            const int p1 = (round == 0) ? 2 : 1, q1 = (round == 2) ? 2 : 3;
            double t0, c0, s0, t1, c1, s1;
            jacobi_angle(A(p0, p0), A(q0, q0), A(p0, q0), t0, c0, s0);
            jacobi_angle(A(p1, p1), A(q1, q1), A(p1, q1), t1, c1, s1);
            jacobi_rotate<4>(A, V, p0, q0, t0, c0, s0);
            jacobi_rotate<4>(A, V, p1, q1, t1, c1, s1);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) lam(j) = A(j, j);
}

template <int N>
GVI_HD void jacobi_eig(Mat<N>& A, Mat<N>& V, Vec<N>& lam) {
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) V(i, j) = (i == j) ? 1.0 : 0.0;
    if (N == 1) {
        lam(0) = A(0, 0);
        return;
    }
    if constexpr (N == 4) {
        jacobi_eig4(A, V, lam);
        return;
    }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            diag += fabs(A(j, j));
#pragma unroll
            for (int i = 0; i < j; ++i) off += fabs(A(i, j));
        }
        if (off <= 1e-300 || off <= 1e-22 * diag) break;
#pragma unroll
        for (int p = 0; p < N - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                const double apq = A(p, q);
                const double app = A(p, p), aqq = A(q, q);
                // skip when the rotation would not change app/aqq at all
                if (fabs(apq) <= 1e-19 * (fabs(app) + fabs(aqq)) || apq == 0.0) {
                    A(p, q) = 0.0;
                    A(q, p) = 0.0;
                    continue;
                }
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0);
                const double s = t * c;
                const double tau = s / (1.0 + c);
                A(p, p) = app - t * apq;
                A(q, q) = aqq + t * apq;
                A(p, q) = 0.0;
                A(q, p) = 0.0;
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    if (r != p && r != q) {
                        const double arp = A(r, p), arq = A(r, q);
                        const double nrp = arp - s * (arq + tau * arp);
                        const double nrq = arq + s * (arp - tau * arq);
                        A(r, p) = nrp;
                        A(p, r) = nrp;
                        A(r, q) = nrq;
                        A(q, r) = nrq;
                    }
                }
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    const double vrp = V(r, p), vrq = V(r, q);
                    V(r, p) = vrp - s * (vrq + tau * vrp);
                    V(r, q) = vrq + s * (vrp - tau * vrq);
                }
            }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) lam(j) = A(j, j);
}

// S = V diag(f(lam)) V^T for f = sqrt and f = 1/sqrt: the PSD square root of Sigma and its inverse.
template <int N>
GVI_HD void sqrt_and_invsqrt(Mat<N>& S, Mat<N>& R, const Mat<N>& Sigma) {
    Mat<N> A = Sigma, V;
    Vec<N> lam;
    jacobi_eig<N>(A, V, lam);
    Vec<N> sq, isq;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        sq(k) = sqrt(lam(k));
        isq(k) = 1.0 / sq(k);
    }
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = j; i < N; ++i) {
            double s = 0.0, r = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double vv = V(i, k) * V(j, k);
                s = fma(vv, sq(k), s);
                r = fma(vv, isq(k), r);
            }
            S(i, j) = s;
            S(j, i) = s;
            R(i, j) = r;
            R(j, i) = r;
        }
}

// Bures-Wasserstein JKO step of one factor (BW_JKO, proxgd/ProxGVIFactorizedBaseGH.h:64-113 and
// proxgd/ProxGVIFactorizedLinear.h:118-157):
//   M = I - eta S_k;  H = M Sigma M^T;  Sigma+ = H/2 + eta I + sqrtm(H (H + 4 eta I))/2;  Vddmu = (inv(Sigma+) - P)/eta.
// H (H + 4 eta I) is a polynomial in the symmetric H, so with H = V diag(lam) V^T everything is diagonal in V:
//   inv(Sigma+) = V diag(1 / (lam/2 + eta + sqrt(lam^2 + 4 eta lam)/2)) V^T     (one Jacobi decomposition, no Schur form).
template <int N>
GVI_HD void bw_jko(Mat<N>& Vdd, const Mat<N>& Sigma, const Mat<N>& P, const Mat<N>& Sk, double eta) {
    Mat<N> M, T, H, V;
    Vec<N> lam;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) M(i, j) = (i == j ? 1.0 : 0.0) - eta * Sk(i, j);
    mm<N>(T, M, Sigma);
    mmt<N>(H, T, M);
    symmetrize<N>(H);
    jacobi_eig<N>(H, V, lam);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double l = lam(k);
        const double disc = fma(l, l, 4.0 * eta * l);
        lam(k) = 1.0 / (0.5 * l + eta + 0.5 * sqrt(disc > 0.0 ? disc : 0.0));
    }
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = j; i < N; ++i) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) v = fma(V(i, k) * V(j, k), lam(k), v);
            const double r = (v - 0.5 * (P(i, j) + P(j, i))) / eta;
            Vdd(i, j) = r;
            Vdd(j, i) = r;
        }
}

}  // namespace gvib200
