// K1S: the fused sigma-point / cost / moment kernel for factors of dimension <= 4, restructured around the
// SIGN-GROUP structure of the sparse Gauss-Hermite rule.
//
// nwspgr's symmetric rule (quadrature/GH/SparseGH/nwspgr.m:108-133 mirrors the positive orthant) consists of groups of
// 2^k nodes (+-a_1, ..., +-a_k on k coordinates, 0 elsewhere) that all carry the same weight: (4, 6) has 953 nodes in
// 190 groups.  For one group
//   x_sigma = mu + sum_i sigma_i d_i,   d_i = S[:, c_i] a_i              (Gray code: one add per row and node)
//   sum_sigma w psi            = w A_0                                    A_T = sum_sigma (prod_{i in T} sigma_i) psi_sigma
//   sum_sigma w psi xi_c       = (w a_c) A_{c}                            (Walsh-Hadamard butterflies on the 2^k values)
//   sum_sigma w psi xi_c xi_d  = (w a_c a_d) A_{c,d},  (w a_c^2) A_0 on the diagonal
// so the per-node cost of the moment accumulation drops from ~20 FP64 instructions to ~4 and the sigma point costs XD
// adds instead of XD*DIM FMAs; the sums are the same numbers as SparseGaussHermite::Integrate's
// (quadrature/SparseGaussHermite.h:197-221), in a different summation order.
//
// Mapping: one THREAD per factor, one WARP per part of the group list (the groups are split into NPART parts of
// equal node count; the 32 lanes of a warp work on 32 consecutive factors, so every table access is warp uniform and
// comes from the kernel-parameter constant bank).  A CTA = NPART warps = 32 factors; the partial sums of the parts
// meet in shared memory in a fixed order (reproducible), then the CTA runs the Vdmu / Vddmu epilogue
// (ngd/NGDFactorizedBaseGH.h:61-73).  Groups with four non-zero coordinates are processed as two 8-node units with
// the sign of the fourth coordinate fixed, which bounds the live psi values to 8.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cost_functors.cuh"

namespace gvib200 {

constexpr int K1S_NPART = 8;
constexpr int K1S_THREADS = 32 * K1S_NPART;
constexpr int K1S_MAX_DATA = 3400;  // doubles; the whole table travels as a kernel parameter (< 32 KB)

// entries of one group with K non-zero coordinates c_0 < ... < c_{K-1}:
//   a[K], w, (w a_i)[K], (w a_i^2)[K], (w a_i a_j)[pairs i < j in lexicographic order]
__host__ __device__ constexpr int k1s_stride(int K) { return 3 * K + 1 + K * (K - 1) / 2; }
__host__ __device__ constexpr int k1s_popc(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1); }
__host__ __device__ constexpr int k1s_pair(int K, int i, int j) {  // index of pair (i < j)
    return i * K - i * (i + 1) / 2 + (j - i - 1);
}

// i-th set bit of mask m
__host__ __device__ constexpr int k1s_coord(int m, int i) {
    int q = 0;
    for (int b = 0; b < 4; ++b)
        if (m & (1 << b)) {
            if (q == i) return b;
            ++q;
        }
    return 0;
}

struct SymTable {
    int dim;
    int n_nodes;                 // nodes of the rule (for bookkeeping)
    double w0;                   // weight of the node at the origin (0 when the rule has none); handled by part 0
    int moff[16];                // first double of mask m's group list in data
    int pbeg[K1S_NPART + 1];     // part p works on glist[pbeg[p] .. pbeg[p+1])
    unsigned short glist[1024];  // (mask << 10) | group index within the mask; parts balanced by node count
    double data[K1S_MAX_DATA];
};

template <class Cost>
struct SymArgs {
    int n;                 // factors
    int state_dim;
    const int* start;      // [n]
    const double* mu;      // joint mean
    const double* SR;      // [n][2*DIM*DIM]: S = Sigma^1/2, R = Sigma^-1/2 (column-major)
    const double* T;       // [n]
    double* fcost;         // [n]
    double* fVdmu;         // [n][DIM]
    double* fVdd;          // [n][DIM*DIM]
    double* raw;           // optional [n][1 + DIM + DIM*DIM]
    double ximax[4];
    Cost cost;
};

template <int DIM>
struct SymAcc {
    static constexpr int NE2 = DIM * (DIM + 1) / 2;
    static constexpr int N = 1 + DIM + NE2;
    double e0;
    double e1[DIM];
    double e2[NE2];  // packed upper triangle
};
template <int DIM>
__host__ __device__ constexpr int k1s_e2(int a, int b) { return a * DIM - a * (a - 1) / 2 + (b - a); }  // a <= b

// Walsh-Hadamard butterflies in place: v[T] <- sum_p (-1)^{|p & T|} v[p]
template <int KF>
__device__ __forceinline__ void k1s_wht(double (&v)[1 << KF]) {
#pragma unroll
    for (int b = 0; b < KF; ++b)
#pragma unroll
        for (int p = 0; p < (1 << KF); ++p)
            if (!(p & (1 << b))) {
                const double u = v[p], w = v[p | (1 << b)];
                v[p] = u + w;
                v[p | (1 << b)] = u - w;
            }
}

// psi at the 2^KF sign patterns of one unit.  Pattern bit i set = coordinate c_i negative.  cidx: the unit's
// coordinates (warp uniform), sS: this thread's S rows in shared memory, sS[(r * DIM + c) * K1S_THREADS].
template <int DIM, class Cost, bool FAST, int KF, bool HASFIX>
__device__ __forceinline__ void k1s_eval_unit(double (&psi)[1 << KF], const double* __restrict__ t, const int (&cidx)[4],
                                              double sfix, const double* __restrict__ sS, const double (&mu)[Cost::XD],
                                              const Cost& cost, int f) {
    constexpr int XD = Cost::XD;
    constexpr int NP = 1 << KF;
    double x[XD], d2[XD][KF > 0 ? KF : 1];
#pragma unroll
    for (int r = 0; r < XD; ++r) x[r] = mu[r];
    if (HASFIX) {
        const double af = sfix * t[KF];  // the fixed coordinate is the last one of the group
#pragma unroll
        for (int r = 0; r < XD; ++r) x[r] = fma(sS[(r * DIM + cidx[KF]) * K1S_THREADS], af, x[r]);
    }
#pragma unroll
    for (int i = 0; i < KF; ++i) {
        const double a = t[i];
#pragma unroll
        for (int r = 0; r < XD; ++r) {
            const double d = sS[(r * DIM + cidx[i]) * K1S_THREADS] * a;
            x[r] += d;
            d2[r][i] = d + d;
        }
    }
    // Gray-code walk over the patterns, evaluated in batches so that a few gathers are in flight
    constexpr int NB = (NP < 4) ? NP : 4;
    int p = 0;
#pragma unroll
    for (int s0 = 0; s0 < NP; s0 += NB) {
        typename Cost::Pending pend[NB];
        int pat[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int s = s0 + q;
            if (s > 0) {
                int b = 0;
                while (!((s >> b) & 1)) ++b;  // compile-time after unrolling
                p ^= (1 << b);
                if (p & (1 << b)) {
#pragma unroll
                    for (int r = 0; r < XD; ++r) x[r] -= d2[r][b];
                } else {
#pragma unroll
                    for (int r = 0; r < XD; ++r) x[r] += d2[r][b];
                }
            }
            pat[q] = p;
            pend[q] = cost.template begin<FAST>(x, f);
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) psi[pat[q]] = cost.finish(pend[q]);
    }
}

// moment accumulation of one unit of mask M (compile-time coordinates) from the Walsh sums A[]
template <int DIM, int M, bool HASFIX>
__device__ __forceinline__ void k1s_acc_mask(SymAcc<DIM>& acc, const double* __restrict__ t, const double* A, double sfix) {
    constexpr int K = k1s_popc(M);
    constexpr int KF = HASFIX ? K - 1 : K;
    constexpr int c[4] = {k1s_coord(M, 0), k1s_coord(M, 1), k1s_coord(M, 2), k1s_coord(M, 3)};
    const double A0 = A[0];
    acc.e0 = fma(t[K], A0, acc.e0);
#pragma unroll
    for (int i = 0; i < KF; ++i) {
        acc.e1[c[i]] = fma(t[K + 1 + i], A[1 << i], acc.e1[c[i]]);
        acc.e2[k1s_e2<DIM>(c[i], c[i])] = fma(t[2 * K + 1 + i], A0, acc.e2[k1s_e2<DIM>(c[i], c[i])]);
#pragma unroll
        for (int j = i + 1; j < KF; ++j)
            acc.e2[k1s_e2<DIM>(c[i], c[j])] =
                fma(t[3 * K + 1 + k1s_pair(K, i, j)], A[(1 << i) | (1 << j)], acc.e2[k1s_e2<DIM>(c[i], c[j])]);
    }
    if (HASFIX) {
        constexpr int jf = K - 1;
        acc.e1[c[jf]] = fma(sfix * t[K + 1 + jf], A0, acc.e1[c[jf]]);
        acc.e2[k1s_e2<DIM>(c[jf], c[jf])] = fma(t[2 * K + 1 + jf], A0, acc.e2[k1s_e2<DIM>(c[jf], c[jf])]);
#pragma unroll
        for (int i = 0; i < KF; ++i)
            acc.e2[k1s_e2<DIM>(c[i], c[jf])] =
                fma(sfix * t[3 * K + 1 + k1s_pair(K, i, jf)], A[1 << i], acc.e2[k1s_e2<DIM>(c[i], c[jf])]);
    }
}

template <int DIM, bool HASFIX>
__device__ __forceinline__ void k1s_acc_switch(SymAcc<DIM>& acc, int mask, const double* __restrict__ t, const double* A,
                                               double sfix) {
    switch (mask) {  // warp uniform
#define K1S_CASE(M_)                                                                         \
    case M_:                                                                                 \
        if constexpr ((M_) < (1 << DIM) && (k1s_popc(M_) == 4) == HASFIX && k1s_popc(M_) >= 1) \
            k1s_acc_mask<DIM, M_, HASFIX>(acc, t, A, sfix);                                  \
        break;
        K1S_CASE(1) K1S_CASE(2) K1S_CASE(3) K1S_CASE(4) K1S_CASE(5) K1S_CASE(6) K1S_CASE(7) K1S_CASE(8)
        K1S_CASE(9) K1S_CASE(10) K1S_CASE(11) K1S_CASE(12) K1S_CASE(13) K1S_CASE(14) K1S_CASE(15)
#undef K1S_CASE
        default: break;
    }
}

// all groups [gb, ge) of one mask with K non-zero coordinates
template <int DIM, class Cost, bool FULL, bool FAST, int K>
__device__ __forceinline__ void k1s_run_mask(SymAcc<DIM>& acc, const SymTable& tab, int mask, int gb, int ge,
                                             const double* __restrict__ sS, const double (&mu)[Cost::XD], const Cost& cost,
                                             int f) {
    constexpr bool HASFIX = (K == 4);
    constexpr int KF = HASFIX ? 3 : K;
    int cidx[4] = {0, 0, 0, 0};
    {
        int q = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (mask & (1 << b)) cidx[q++] = b;
    }
    for (int g = gb; g < ge; ++g) {
        const double* t = tab.data + tab.moff[mask] + g * k1s_stride(K);
#pragma unroll 1
        for (int half = 0; half < (HASFIX ? 2 : 1); ++half) {
            const double sfix = half ? -1.0 : 1.0;
            double psi[1 << KF];
            k1s_eval_unit<DIM, Cost, FAST, KF, HASFIX>(psi, t, cidx, sfix, sS, mu, cost, f);
            if (FULL) {
                k1s_wht<KF>(psi);
                k1s_acc_switch<DIM, HASFIX>(acc, mask, t, psi, sfix);
            } else {
                double s = psi[0];
#pragma unroll
                for (int q = 1; q < (1 << KF); ++q) s += psi[q];
                acc.e0 = fma(t[K], s, acc.e0);
            }
        }
    }
}

template <int DIM, class Cost, bool FULL, bool FAST>
__device__ __forceinline__ void k1s_run_part(SymAcc<DIM>& acc, const SymTable& tab, int part, const double* __restrict__ sS,
                                             const double (&mu)[Cost::XD], const Cost& cost, int f) {
    if (part == 0 && tab.w0 != 0.0) {  // the node at the origin
        typename Cost::Pending pd = cost.template begin<FAST>(mu, f);
        acc.e0 = fma(tab.w0, cost.finish(pd), acc.e0);
    }
#pragma unroll 1
    for (int idx = tab.pbeg[part]; idx < tab.pbeg[part + 1]; ++idx) {
        const int ent = tab.glist[idx];
        const int mask = ent >> 10, gb = ent & 1023;
        const int K = __popc(mask);
        if (K == 1) k1s_run_mask<DIM, Cost, FULL, FAST, 1>(acc, tab, mask, gb, gb + 1, sS, mu, cost, f);
        if constexpr (DIM >= 2)
            if (K == 2) k1s_run_mask<DIM, Cost, FULL, FAST, 2>(acc, tab, mask, gb, gb + 1, sS, mu, cost, f);
        if constexpr (DIM >= 3)
            if (K == 3) k1s_run_mask<DIM, Cost, FULL, FAST, 3>(acc, tab, mask, gb, gb + 1, sS, mu, cost, f);
        if constexpr (DIM >= 4)
            if (K == 4) k1s_run_mask<DIM, Cost, FULL, FAST, 4>(acc, tab, mask, gb, gb + 1, sS, mu, cost, f);
    }
}

template <int DIM, class Cost, bool FULL>
__global__ void __launch_bounds__(K1S_THREADS, 2)
    k_moments_sym(const __grid_constant__ SymTable tab, const __grid_constant__ SymArgs<Cost> a) {
    constexpr int XD = Cost::XD;
    constexpr int NE2 = DIM * (DIM + 1) / 2;
    constexpr int NACC = 1 + DIM + NE2;
    constexpr int NOUT = 1 + DIM + DIM * DIM;
    // S rows [(r*DIM + c)][thread] during the node loop, afterwards the partial sums of the parts [part][e][lane]
    constexpr int NBUF = (XD * DIM * K1S_THREADS > K1S_NPART * NACC * 32) ? XD * DIM * K1S_THREADS : K1S_NPART * NACC * 32;
    __shared__ double buf[NBUF];
    __shared__ double tot[NOUT][33];                        // totals per factor: e0, e1, e2 (full, mirrored)
    double* sSall = buf;
    double (*red)[NACC][32] = reinterpret_cast<double (*)[NACC][32]>(buf);
    const int lane = threadIdx.x & 31;
    const int part = threadIdx.x >> 5;
    const int f0 = blockIdx.x * 32;
    const int f = min(f0 + lane, a.n - 1);  // tail lanes recompute the last factor and do not store
    double mu[XD];
    const double* sS = sSall + threadIdx.x;
    bool fast;
    {
        const double* Sp = a.SR + (size_t)f * 2 * DIM * DIM;
        const double* mp = a.mu + (size_t)a.start[f] * a.state_dim;
        double lo[XD], hi[XD];
#pragma unroll
        for (int r = 0; r < XD; ++r) {
            mu[r] = __ldg(mp + r);
            double rad = 0.0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                const double s = __ldg(Sp + r + c * DIM);
                sSall[(r * DIM + c) * K1S_THREADS + threadIdx.x] = s;
                rad = fma(fabs(s), a.ximax[c], rad);
            }
            lo[r] = mu[r] - rad;
            hi[r] = mu[r] + rad;
        }
        fast = __all_sync(0xffffffffu, a.cost.fast_ok(lo, hi));
    }
    SymAcc<DIM> acc;
    acc.e0 = 0.0;
#pragma unroll
    for (int c = 0; c < DIM; ++c) acc.e1[c] = 0.0;
#pragma unroll
    for (int c = 0; c < NE2; ++c) acc.e2[c] = 0.0;
    if (fast) k1s_run_part<DIM, Cost, FULL, true>(acc, tab, part, sS, mu, a.cost, f);
    else k1s_run_part<DIM, Cost, FULL, false>(acc, tab, part, sS, mu, a.cost, f);

    // ---- parts meet in shared memory, summed in part order ----
    __syncthreads();  // every warp is done with its S rows: the buffer is reused
    red[part][0][lane] = acc.e0;
    if (FULL) {
#pragma unroll
        for (int c = 0; c < DIM; ++c) red[part][1 + c][lane] = acc.e1[c];
#pragma unroll
        for (int c = 0; c < NE2; ++c) red[part][1 + DIM + c][lane] = acc.e2[c];
    }
    __syncthreads();
    const double sc = a.cost.scale();
    for (int idx = threadIdx.x; idx < (FULL ? NACC : 1) * 32; idx += K1S_THREADS) {
        const int e = idx >> 5, l = idx & 31;
        double s = red[0][e][l];
#pragma unroll
        for (int p = 1; p < K1S_NPART; ++p) s += red[p][e][l];
        s *= sc;
        if (e <= DIM) {
            tot[e][l] = s;
        } else {  // unpack the upper triangle into the full symmetric matrix
            int q = e - 1 - DIM, ra = 0;
            while (q >= DIM - ra) {
                q -= DIM - ra;
                ++ra;
            }
            const int rb = ra + q;
            tot[1 + DIM + ra + rb * DIM][l] = s;
            tot[1 + DIM + rb + ra * DIM][l] = s;
        }
    }
    __syncthreads();
    // ---- epilogue: Vdmu = R e1 / T, Vddmu = R (e2 - e0 I) R / T (upper triangle mirrored), cost = e0 / T ----
    const int nf = min(32, a.n - f0);
    if (!FULL) {
        if (threadIdx.x < nf) a.fcost[f0 + threadIdx.x] = tot[0][threadIdx.x] / __ldg(a.T + f0 + threadIdx.x);
        return;
    }
    constexpr int NEP = DIM * DIM + DIM + 1;
    for (int idx = threadIdx.x; idx < nf * NEP; idx += K1S_THREADS) {
        const int l = idx / NEP, e = idx - l * NEP;
        const int ff = f0 + l;
        const double invT = 1.0 / __ldg(a.T + ff);
        const double* R = a.SR + (size_t)ff * 2 * DIM * DIM + DIM * DIM;
        const double e0 = tot[0][l];
        if (e < DIM * DIM) {
            int i = e % DIM, j = e / DIM;
            if (i > j) {  // upper triangle mirrored (ngd/NGDFactorizedBaseGH.h:71-72)
                const int tt = i;
                i = j;
                j = tt;
            }
            double v = 0.0;
            for (int b = 0; b < DIM; ++b) {
                double t = 0.0;
                for (int aa = 0; aa < DIM; ++aa) {
                    const double m = tot[1 + DIM + aa + b * DIM][l] - (aa == b ? e0 : 0.0);
                    t = fma(__ldg(R + aa + i * DIM), m, t);
                }
                v = fma(t, __ldg(R + b + j * DIM), v);
            }
            a.fVdd[(size_t)ff * DIM * DIM + e] = v * invT;
        } else if (e < DIM * DIM + DIM) {
            const int i = e - DIM * DIM;
            double v = 0.0;
            for (int aa = 0; aa < DIM; ++aa) v = fma(__ldg(R + i + aa * DIM), tot[1 + aa][l], v);
            a.fVdmu[(size_t)ff * DIM + i] = v * invT;
        } else {
            a.fcost[ff] = e0 * invT;
        }
    }
    if (a.raw != nullptr) {
        for (int idx = threadIdx.x; idx < nf * NOUT; idx += K1S_THREADS) {
            const int l = idx / NOUT, e = idx - l * NOUT;
            a.raw[(size_t)(f0 + l) * NOUT + e] = tot[e][l];
        }
    }
}

}  // namespace gvib200
