// Device-side set-up of the LTV GP prior: transition matrix Phi(dt) and Gramian Q(dt) of
//     Phi' = A(t) Phi,   Q' = A(t) Q + Q A(t)^T + B(t) B(t)^T,   Phi(0) = I, Q(0) = 0
// with A, B piece-wise constant on the four quarter intervals of a link (gp/LTV_prior.h:187-197 `A_function` /
// `system_param`), for a whole batch of links at once -- one thread per link.  The reference integrates both ODEs with
// GSL rkf45 at tolerance 1e-12 (gp/LTV_prior.h:123-152); on a piece-wise constant system each quarter has the closed form
// of Van Loan's block exponential
//     exp(h [[-A, B B^T], [0, A^T]]) = [[., E12], [0, E22]],   Phi_k = E22^T,   Q_k = Phi_k E12,
// which is what this kernel evaluates (scaling and squaring around a degree-18 Taylor polynomial, |Y| <= 1/4: truncation
// below 1e-30), followed by  Q <- Phi_k Q Phi_k^T + Q_k,  Phi <- Phi_k Phi.  Optionally also K^-1 = Q^-1 (Cholesky).
#pragma once
#include <cuda_runtime.h>

namespace gvib200 {

template <int N>
__device__ __forceinline__ void ltv_mm(double* __restrict__ C, const double* __restrict__ A, const double* __restrict__ B) {
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) {
            double s = 0.0;
            for (int k = 0; k < N; ++k) s = fma(A[i + k * N], B[k + j * N], s);
            C[i + j * N] = s;
        }
}

// NS: state dimension of a link end (4 for the planar point robot), NB: columns of B
template <int NS>
__global__ void __launch_bounds__(64) k_ltv_transition(int n, int nb, double delta_t, const double* __restrict__ A,
                                                       const double* __restrict__ B, double* __restrict__ Phi_out,
                                                       double* __restrict__ Q_out, double* __restrict__ Qinv_out) {
    constexpr int N2 = 2 * NS;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const double h = delta_t / 4.0;
    double Phi[NS * NS], Q[NS * NS];
    for (int e = 0; e < NS * NS; ++e) {
        Phi[e] = (e % NS == e / NS) ? 1.0 : 0.0;
        Q[e] = 0.0;
    }
    double Y[N2 * N2], E[N2 * N2], T[N2 * N2], W[N2 * N2];
    for (int k = 0; k < 4; ++k) {
        const double* Ak = A + ((size_t)f * 4 + k) * NS * NS;
        const double* Bk = B + ((size_t)f * 4 + k) * NS * nb;
        for (int e = 0; e < N2 * N2; ++e) Y[e] = 0.0;
        for (int j = 0; j < NS; ++j)
            for (int i = 0; i < NS; ++i) {
                double bb = 0.0;
                for (int q = 0; q < nb; ++q) bb = fma(Bk[i + q * NS], Bk[j + q * NS], bb);
                Y[i + j * N2] = -Ak[i + j * NS] * h;
                Y[i + (NS + j) * N2] = bb * h;
                Y[(NS + i) + (NS + j) * N2] = Ak[j + i * NS] * h;
            }
        // scaling: max row sum <= 1/4
        double nrm = 0.0;
        for (int i = 0; i < N2; ++i) {
            double r = 0.0;
            for (int j = 0; j < N2; ++j) r += fabs(Y[i + j * N2]);
            nrm = fmax(nrm, r);
        }
        int s = 0;
        while (nrm > 0.25 && s < 60) {
            nrm *= 0.5;
            ++s;
        }
        const double sc = ldexp(1.0, -s);
        for (int e = 0; e < N2 * N2; ++e) {
            Y[e] *= sc;
            E[e] = ((e % N2 == e / N2) ? 1.0 : 0.0) + Y[e];
            T[e] = Y[e];
        }
        for (int m = 2; m <= 18; ++m) {
            ltv_mm<N2>(W, T, Y);
            const double inv = 1.0 / m;
            for (int e = 0; e < N2 * N2; ++e) {
                T[e] = W[e] * inv;
                E[e] += T[e];
            }
        }
        for (int q = 0; q < s; ++q) {
            ltv_mm<N2>(W, E, E);
            for (int e = 0; e < N2 * N2; ++e) E[e] = W[e];
        }
        // Phi_k = E22^T, Q_k = sym(Phi_k E12)
        double Pk[NS * NS], Qk[NS * NS], t1[NS * NS], t2[NS * NS];
        for (int j = 0; j < NS; ++j)
            for (int i = 0; i < NS; ++i) {
                Pk[i + j * NS] = E[(NS + j) + (NS + i) * N2];
                t1[i + j * NS] = E[i + (NS + j) * N2];
            }
        ltv_mm<NS>(Qk, Pk, t1);
        // Q <- Pk Q Pk^T + sym(Qk)
        ltv_mm<NS>(t1, Pk, Q);
        for (int j = 0; j < NS; ++j)
            for (int i = 0; i < NS; ++i) {
                double sum = 0.0;
                for (int q = 0; q < NS; ++q) sum = fma(t1[i + q * NS], Pk[j + q * NS], sum);
                t2[i + j * NS] = sum + 0.5 * (Qk[i + j * NS] + Qk[j + i * NS]);
            }
        for (int e = 0; e < NS * NS; ++e) Q[e] = t2[e];
        ltv_mm<NS>(t1, Pk, Phi);
        for (int e = 0; e < NS * NS; ++e) Phi[e] = t1[e];
    }
    for (int j = 0; j < NS; ++j)
        for (int i = 0; i < NS; ++i) {
            Phi_out[(size_t)f * NS * NS + i + j * NS] = Phi[i + j * NS];
            Q_out[(size_t)f * NS * NS + i + j * NS] = 0.5 * (Q[i + j * NS] + Q[j + i * NS]);
        }
    if (Qinv_out != nullptr) {
        // K^-1 = Q^-1 through the Cholesky factor of the symmetrised Gramian, symmetric by construction
        double L[NS * NS], Li[NS * NS];
        for (int e = 0; e < NS * NS; ++e) L[e] = Li[e] = 0.0;
        for (int j = 0; j < NS; ++j) {
            double dsum = 0.5 * (Q[j + j * NS] + Q[j + j * NS]);
            for (int q = 0; q < j; ++q) dsum -= L[j + q * NS] * L[j + q * NS];
            const double djj = sqrt(dsum);
            L[j + j * NS] = djj;
            for (int i = j + 1; i < NS; ++i) {
                double v = 0.5 * (Q[i + j * NS] + Q[j + i * NS]);
                for (int q = 0; q < j; ++q) v -= L[i + q * NS] * L[j + q * NS];
                L[i + j * NS] = v / djj;
            }
        }
        for (int j = 0; j < NS; ++j) {  // Li = L^-1 (lower), column by column
            Li[j + j * NS] = 1.0 / L[j + j * NS];
            for (int i = j + 1; i < NS; ++i) {
                double v = 0.0;
                for (int q = j; q < i; ++q) v -= L[i + q * NS] * Li[q + j * NS];
                Li[i + j * NS] = v / L[i + i * NS];
            }
        }
        for (int j = 0; j < NS; ++j)
            for (int i = 0; i <= j; ++i) {
                double v = 0.0;
                for (int q = j; q < NS; ++q) v = fma(Li[q + i * NS], Li[q + j * NS], v);
                Qinv_out[(size_t)f * NS * NS + i + j * NS] = v;
                Qinv_out[(size_t)f * NS * NS + j + i * NS] = v;
            }
    }
}

}  // namespace gvib200
