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
// Mapping: 8 lanes ("parts") per factor, 4 factors per warp, 32 factors per CTA.  For every coordinate mask the groups
// are dealt round-robin to the 8 parts (the last round is padded with zero-weight groups), so all lanes of a warp run
// the same code on a group of the same shape while the 8 lanes of one factor gather neighbouring cells of the distance
// field (a warp-wide gather touches ~10 sectors instead of 32).  The table (per mask: [round][entry][part]) is staged
// once per CTA into shared memory with a TMA bulk copy.  The partial sums of the 8 parts meet in a fixed xor-butterfly
// (reproducible), then the CTA runs the Vdmu / Vddmu epilogue (ngd/NGDFactorizedBaseGH.h:61-73).  Groups with four
// non-zero coordinates are processed as two 8-node units with the sign of the fourth coordinate fixed, which bounds
// the live psi values to 8.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cost_functors.cuh"
#include "smallmat.h"

namespace gvib200 {

constexpr int K1S_NPART = 8;          // lanes per factor
constexpr int K1S_THREADS = 256;      // 8 warps x 4 factors per CTA (10 x 2 and 6 x 3 CTAs per SM spill and run ~10 % slower)
constexpr int K1S_FPC = K1S_THREADS / K1S_NPART;  // factors per CTA
constexpr int K1S_MAX_DATA = 6144;    // doubles of the staged table (48 KB of shared memory)

// entries of one group with K non-zero coordinates c_0 < ... < c_{K-1}:
//   a[K], w, (w a_i)[K], (w a_i^2)[K], (w a_i a_j)[pairs i < j in lexicographic order]
__host__ __device__ constexpr int k1s_stride(int K) { return 3 * K + 1 + K * (K - 1) / 2; }
__host__ __device__ constexpr int k1s_popc(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1); }
__host__ __device__ constexpr int k1s_pair(int K, int i, int j) {  // index of pair (i < j)
    return i * K - i * (i + 1) / 2 + (j - i - 1);
}

// tuning (measured on B200 at the headline shape): 4 pending gathers per batch and 2 resident CTAs per SM (128
// registers); 3 CTAs at 80 registers spill and run 25-35 % slower, 8 pending gathers change nothing
__host__ __device__ constexpr int k1s_nb(int) { return 4; }
__host__ __device__ constexpr int k1s_minb(int) { return 2; }

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
    int n_nodes;      // nodes of the rule (bookkeeping)
    int ndata;        // doubles in data (multiple of 2)
    int moff[16];     // first double of mask m's block in data; mask 0 = the node at the origin (one pseudo entry: w)
    int rounds[16];   // rounds of mask m: round r hands group 8 r + p to part p
    const double* data;  // device: per mask [round][entry][part]
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
    unsigned long long* evaluated;  // optional counter: factors whose sigma points were evaluated (not culled)
    const double* covD;    // marginal covariance blocks of the sweep's state (culling pass only)
    const double* covO;    // off-diagonal covariance blocks (s, s+1) (fused culling + prologue pass only)
    double* SR_out;        // fused culling + prologue pass: S, R of the factors that are kept (same array as SR)
    double xinorm;         // max ||xi||_2 over the rule's nodes
    // free-space culling (k_cull_sym): the factors to evaluate, compacted; null = all n factors in order
    int* active;           // [n] factor indices
    int* n_active;         // their number; reset (together with done) by the last CTA of the moment kernel
    unsigned* done;        // CTAs of the moment kernel that have finished
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
template <int DIM, class Cost, bool FAST, int KF, bool HASFIX, int VAR>
__device__ __forceinline__ void k1s_eval_unit(double (&psi)[1 << KF], const double* __restrict__ t, const int (&cidx)[4],
                                              double sfix, const double* __restrict__ sS, const double (&mu)[Cost::XD],
                                              const Cost& cost, int f) {
    constexpr int XD = Cost::XD;
    constexpr int NP = 1 << KF;
    double x[XD], d2[XD][KF > 0 ? KF : 1];
#pragma unroll
    for (int r = 0; r < XD; ++r) x[r] = mu[r];
    if (HASFIX) {
        const double af = sfix * t[KF * K1S_NPART];  // the fixed coordinate is the last one of the group
#pragma unroll
        for (int r = 0; r < XD; ++r) x[r] = fma(sS[(r * DIM + cidx[KF]) * K1S_THREADS], af, x[r]);
    }
#pragma unroll
    for (int i = 0; i < KF; ++i) {
        const double a = t[i * K1S_NPART];
#pragma unroll
        for (int r = 0; r < XD; ++r) {
            const double d = sS[(r * DIM + cidx[i]) * K1S_THREADS] * a;
            x[r] += d;
            d2[r][i] = d + d;
        }
    }
    // Gray-code walk over the patterns, evaluated in batches so that a few gathers are in flight
    constexpr int NB = (NP < k1s_nb(VAR)) ? NP : k1s_nb(VAR);
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
            if constexpr (Cost::PREMAP) pend[q] = cost.template begin_mapped<FAST>(x, f);
            else pend[q] = cost.template begin<FAST>(x, f);
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
    acc.e0 = fma(t[K * K1S_NPART], A0, acc.e0);
#pragma unroll
    for (int i = 0; i < KF; ++i) {
        acc.e1[c[i]] = fma(t[(K + 1 + i) * K1S_NPART], A[1 << i], acc.e1[c[i]]);
        acc.e2[k1s_e2<DIM>(c[i], c[i])] = fma(t[(2 * K + 1 + i) * K1S_NPART], A0, acc.e2[k1s_e2<DIM>(c[i], c[i])]);
#pragma unroll
        for (int j = i + 1; j < KF; ++j)
            acc.e2[k1s_e2<DIM>(c[i], c[j])] =
                fma(t[(3 * K + 1 + k1s_pair(K, i, j)) * K1S_NPART], A[(1 << i) | (1 << j)], acc.e2[k1s_e2<DIM>(c[i], c[j])]);
    }
    if (HASFIX) {
        constexpr int jf = K - 1;
        acc.e1[c[jf]] = fma(sfix * t[(K + 1 + jf) * K1S_NPART], A0, acc.e1[c[jf]]);
        acc.e2[k1s_e2<DIM>(c[jf], c[jf])] = fma(t[(2 * K + 1 + jf) * K1S_NPART], A0, acc.e2[k1s_e2<DIM>(c[jf], c[jf])]);
#pragma unroll
        for (int i = 0; i < KF; ++i)
            acc.e2[k1s_e2<DIM>(c[i], c[jf])] =
                fma(sfix * t[(3 * K + 1 + k1s_pair(K, i, jf)) * K1S_NPART], A[1 << i], acc.e2[k1s_e2<DIM>(c[i], c[jf])]);
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

// all rounds of one mask with K non-zero coordinates; tm: this lane's column of the mask's block in shared memory
template <int DIM, class Cost, bool FULL, bool FAST, int K, int VAR>
__device__ __forceinline__ void k1s_run_mask(SymAcc<DIM>& acc, const double* __restrict__ tm, int rounds, int mask,
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
#pragma unroll 1
    for (int r = 0; r < rounds; ++r) {
        const double* t = tm + r * (k1s_stride(K) * K1S_NPART);
#pragma unroll 1
        for (int half = 0; half < (HASFIX ? 2 : 1); ++half) {
            const double sfix = half ? -1.0 : 1.0;
            double psi[1 << KF];
            k1s_eval_unit<DIM, Cost, FAST, KF, HASFIX, VAR>(psi, t, cidx, sfix, sS, mu, cost, f);
            if (FULL) {
                k1s_wht<KF>(psi);
                k1s_acc_switch<DIM, HASFIX>(acc, mask, t, psi, sfix);
            } else {
                double s = psi[0];
#pragma unroll
                for (int q = 1; q < (1 << KF); ++q) s += psi[q];
                acc.e0 = fma(t[K * K1S_NPART], s, acc.e0);
            }
        }
    }
}

template <int DIM, class Cost, bool FULL, bool FAST, int VAR>
__device__ __forceinline__ void k1s_run_part(SymAcc<DIM>& acc, const SymTable& tab, const double* __restrict__ stab, int part,
                                             const double* __restrict__ sS, const double (&mu)[Cost::XD], const Cost& cost,
                                             int f) {
    {  // the node at the origin: weight w0 for part 0, zero for the other parts
        typename Cost::Pending pd;
        if constexpr (Cost::PREMAP) pd = cost.template begin_mapped<FAST>(mu, f);
        else pd = cost.template begin<FAST>(mu, f);
        acc.e0 = fma(stab[tab.moff[0] + part], cost.finish(pd), acc.e0);
    }
#pragma unroll 1
    for (int mask = 1; mask < (1 << DIM); ++mask) {
        const int rounds = tab.rounds[mask];
        if (rounds == 0) continue;
        const double* tm = stab + tab.moff[mask] + part;
        const int K = __popc(mask);
        if (K == 1) k1s_run_mask<DIM, Cost, FULL, FAST, 1, VAR>(acc, tm, rounds, mask, sS, mu, cost, f);
        if constexpr (DIM >= 2)
            if (K == 2) k1s_run_mask<DIM, Cost, FULL, FAST, 2, VAR>(acc, tm, rounds, mask, sS, mu, cost, f);
        if constexpr (DIM >= 3)
            if (K == 3) k1s_run_mask<DIM, Cost, FULL, FAST, 3, VAR>(acc, tm, rounds, mask, sS, mu, cost, f);
        if constexpr (DIM >= 4)
            if (K == 4) k1s_run_mask<DIM, Cost, FULL, FAST, 4, VAR>(acc, tm, rounds, mask, sS, mu, cost, f);
    }
}

// TMA bulk copy helpers (cp.async.bulk + mbarrier), as in kernels.cuh
__device__ __forceinline__ uint32_t k1s_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// last CTA out resets the compacted-list counters for the next sweep
template <class Args>
__device__ __forceinline__ void k1s_finish_cta(const Args& a) {
    if (a.active == nullptr) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.done, 1u) == gridDim.x - 1) {
            *a.n_active = 0;
            *a.done = 0u;
            __threadfence();
        }
    }
}

// Free-space culling pass (one thread per factor, before the moment kernel).  Every sigma point x = mu + S xi has
//   |x_r - mu_r| = |sum_c S_rc xi_c| <= ||S_r||_2 ||xi||_2 = sqrt(Sigma_rr) ||xi||_2      (S = Sigma^1/2 is symmetric),
// so the box mu_r +- sqrt(Sigma_rr) max_i ||xi_i||_2 holds them all and needs only the diagonal of the factor's marginal
// covariance.  A factor whose cost functor proves psi == 0 on that box gets its (exactly zero) outputs written here; the
// others are appended to the compacted list the moment kernel works on (warp-aggregated append: the order of the list
// varies from run to run, the per-factor arithmetic and therefore every result does not).
template <int DIM, class Cost, bool FULL>
__global__ void __launch_bounds__(256) k_cull_sym(const __grid_constant__ SymArgs<Cost> a) {
    constexpr int XD = Cost::XD;
    constexpr int NOUT = 1 + DIM + DIM * DIM;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (f < a.n) {
        const int s = a.start[f], sd = a.state_dim;
        const double* mp = a.mu + (size_t)s * sd;
        double lo[XD], hi[XD];
#pragma unroll
        for (int r = 0; r < XD; ++r) {
            const int blk = r / sd, q = r - blk * sd;  // coordinate r lives in state s + blk
            const double var = __ldg(a.covD + (size_t)(s + blk) * sd * sd + q + q * sd);
            const double m = __ldg(mp + r);
            const double rad = sqrt(fmax(var, 0.0)) * a.xinorm * (1.0 + 1e-12);
            lo[r] = m - rad;
            hi[r] = m + rad;
        }
        keep = !a.cost.all_zero(lo, hi);
        if (!keep) a.fcost[f] = 0.0;
    }
    const int lane_ = threadIdx.x & 31;
    if (FULL) {
        // the warp zeroes the outputs of its culled factors together: whole lines instead of 8-byte pieces
        unsigned z = __ballot_sync(0xffffffffu, f < a.n && !keep);
        const int fw = f - lane_;  // first factor of the warp
        while (z != 0u) {
            const int l = __ffs(z) - 1;
            z &= z - 1u;
            const size_t ff = (size_t)(fw + l);
            for (int e = lane_; e < DIM * DIM; e += 32) a.fVdd[ff * DIM * DIM + e] = 0.0;
            if (lane_ < DIM) a.fVdmu[ff * DIM + lane_] = 0.0;
            if (a.raw != nullptr)
                for (int e = lane_; e < NOUT; e += 32) a.raw[ff * NOUT + e] = 0.0;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m != 0u) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(a.n_active, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (keep) a.active[base + __popc(m & ((1u << lane) - 1u))] = f;
    }
}

// Free-space culling FUSED with the per-factor prologue (K2): the factors that survive the box test are compacted inside
// the CTA and the first threads of the CTA -- dense warps -- form S = Sigma^1/2, R = Sigma^-1/2 for them (the same
// extraction and the same sqrt_and_invsqrt as k_prologue: identical bits); culled factors get their exactly-zero outputs
// and NO square root (nothing reads it: their raw moments are zero).  At the headline shape half of the Jacobi work
// disappears and the other half no longer competes with the latency-bound solve pass for the SMs.  One atomic per CTA
// appends the CTA's survivors to the global list.
template <int DIM, class Cost, bool FULL>
__global__ void __launch_bounds__(256) k_cull_prologue_sym(const __grid_constant__ SymArgs<Cost> a) {
    constexpr int XD = Cost::XD;
    constexpr int NOUT = 1 + DIM + DIM * DIM;
    __shared__ int lst[256];
    __shared__ int wcnt[8];
    __shared__ int gbase;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
    bool keep = false;
    if (f < a.n) {
        const int s = a.start[f], sd = a.state_dim;
        const double* mp = a.mu + (size_t)s * sd;
        double lo[XD], hi[XD];
#pragma unroll
        for (int r = 0; r < XD; ++r) {
            const int blk = r / sd, q = r - blk * sd;  // coordinate r lives in state s + blk
            const double var = __ldg(a.covD + (size_t)(s + blk) * sd * sd + q + q * sd);
            const double m = __ldg(mp + r);
            const double rad = sqrt(fmax(var, 0.0)) * a.xinorm * (1.0 + 1e-12);
            lo[r] = m - rad;
            hi[r] = m + rad;
        }
        keep = !a.cost.all_zero(lo, hi);
        if (!keep) a.fcost[f] = 0.0;
    }
    if (FULL) {
        // the warp zeroes the outputs of its culled factors together: whole lines instead of 8-byte pieces
        unsigned z = __ballot_sync(0xffffffffu, f < a.n && !keep);
        const int fw = f - lane_;  // first factor of the warp
        while (z != 0u) {
            const int l = __ffs(z) - 1;
            z &= z - 1u;
            const size_t ff = (size_t)(fw + l);
            for (int e = lane_; e < DIM * DIM; e += 32) a.fVdd[ff * DIM * DIM + e] = 0.0;
            if (lane_ < DIM) a.fVdmu[ff * DIM + lane_] = 0.0;
            if (a.raw != nullptr)
                for (int e = lane_; e < NOUT; e += 32) a.raw[ff * NOUT + e] = 0.0;
        }
    }
    // compaction inside the CTA (factor order preserved), one atomic per CTA for its place in the global list
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane_ == 0) wcnt[warp_] = __popc(m);
    __syncthreads();
    int before = 0, nk = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int c = wcnt[w];
        if (w < warp_) before += c;
        nk += c;
    }
    if (keep) lst[before + __popc(m & ((1u << lane_) - 1u))] = f;
    if (threadIdx.x == 0) gbase = (nk > 0) ? atomicAdd(a.n_active, nk) : 0;
    __syncthreads();
    if ((int)threadIdx.x >= nk) return;
    const int fk = lst[threadIdx.x];
    a.active[gbase + threadIdx.x] = fk;
    // ---- prologue of the kept factor fk (k_prologue: gvibase/GVIFactorizedBase.h:111-114, quadrature/SparseGaussHermite.h:231-233)
    const int s = a.start[fk], sd = a.state_dim;
    const size_t bs = (size_t)sd * sd;
    Mat<DIM> Sig, S, R;
#pragma unroll
    for (int j = 0; j < DIM; ++j)
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            const int bi = i / sd, ii = i - bi * sd, bj = j / sd, jj = j - bj * sd;
            double v;
            if (bi == bj) v = a.covD[(size_t)(s + bi) * bs + ii + jj * sd];
            else if (bi < bj) v = a.covO[(size_t)s * bs + ii + jj * sd];  // block (s, s+1)
            else v = a.covO[(size_t)s * bs + jj + ii * sd];               // its transpose
            Sig(i, j) = v;
        }
    sqrt_and_invsqrt<DIM>(S, R, Sig);
    double* out = a.SR_out + (size_t)fk * 2 * DIM * DIM;
#pragma unroll
    for (int e = 0; e < DIM * DIM; ++e) {
        out[e] = S.a[e];
        out[DIM * DIM + e] = R.a[e];
    }
}

template <int DIM, class Cost, bool FULL, int VAR = 0>
__global__ void __launch_bounds__(K1S_THREADS, k1s_minb(VAR))
    k_moments_sym(const __grid_constant__ SymTable tab, const __grid_constant__ SymArgs<Cost> a) {
    constexpr int XD = Cost::XD;
    constexpr int NE2 = DIM * (DIM + 1) / 2;
    constexpr int NACC = 1 + DIM + NE2;
    constexpr int NOUT = 1 + DIM + DIM * DIM;
    extern __shared__ __align__(16) double stab[];          // the staged table, tab.ndata doubles
    __shared__ double sSall[XD * DIM * K1S_THREADS];        // S rows, [(r*DIM + c)][thread]
    __shared__ double tot[NOUT][K1S_FPC + 1];                        // totals per factor: e0, e1, e2 (full, mirrored)
    __shared__ double sR[DIM * DIM + 1][K1S_FPC + 1];                // R = Sigma^-1/2 and 1/T of the CTA's factors, fetched up
                                                                     // front so that the epilogue does not wait on HBM
    __shared__ __align__(8) uint64_t mbar;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int part = lane & (K1S_NPART - 1);
    const int fl = warp * 4 + (lane >> 3);  // factor of this lane within the CTA, 0..31
    const int f0 = blockIdx.x * K1S_FPC;
    // with culling the CTA works on the compacted list; CTAs beyond it leave at once
    const int nwork = (a.active != nullptr) ? *a.n_active : a.n;
    if (f0 >= nwork) {
        k1s_finish_cta(a);
        return;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0 && a.evaluated != nullptr) atomicAdd(a.evaluated, (unsigned long long)nwork);
    const int slot = min(f0 + fl, nwork - 1);  // tail lanes recompute the last factor and do not store
    const int f = (a.active != nullptr) ? a.active[slot] : slot;
    // ---- stage the table: one TMA bulk copy, completion on an mbarrier ----
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k1s_smem_u32(&mbar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned bytes = (unsigned)tab.ndata * sizeof(double);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k1s_smem_u32(&mbar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         k1s_smem_u32(stab)),
                     "l"(tab.data), "r"(bytes), "r"(k1s_smem_u32(&mbar))
                     : "memory");
    }
    double mu[XD];
    const double* sS = sSall + threadIdx.x;
    bool fast;
    {
        const double* Sp = a.SR + (size_t)f * 2 * DIM * DIM;
        if (FULL) {  // the 8 lanes of a factor share the loads of R (and 1/T)
            for (int e = part; e < DIM * DIM; e += K1S_NPART) sR[e][fl] = __ldg(Sp + DIM * DIM + e);
        }
        if (part == 0) sR[DIM * DIM][fl] = 1.0 / __ldg(a.T + f);
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
        if constexpr (Cost::PREMAP) {  // from here on mu and S live in the functor's mapped coordinates
#pragma unroll
            for (int r = 0; r < XD; ++r) {
                mu[r] = a.cost.pre_mu(r, mu[r]);
#pragma unroll
                for (int c = 0; c < DIM; ++c) sSall[(r * DIM + c) * K1S_THREADS + threadIdx.x] *= a.cost.pre_scale(r);
            }
        }
    }
    __syncthreads();  // mbarrier initialised (and visible) before anybody waits on it
    {
        uint32_t ok;
        do {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(k1s_smem_u32(&mbar)), "r"(0)
                : "memory");
        } while (!ok);
    }
    SymAcc<DIM> acc;
    acc.e0 = 0.0;
#pragma unroll
    for (int c = 0; c < DIM; ++c) acc.e1[c] = 0.0;
#pragma unroll
    for (int c = 0; c < NE2; ++c) acc.e2[c] = 0.0;
    if (fast) k1s_run_part<DIM, Cost, FULL, true, VAR>(acc, tab, stab, part, sS, mu, a.cost, f);
    else k1s_run_part<DIM, Cost, FULL, false, VAR>(acc, tab, stab, part, sS, mu, a.cost, f);

    // ---- the 8 parts of a factor meet in a fixed xor-butterfly over lane bits 0..2 ----
#pragma unroll
    for (int o = 1; o < K1S_NPART; o <<= 1) {
        acc.e0 += __shfl_xor_sync(0xffffffffu, acc.e0, o);
        if (FULL) {
#pragma unroll
            for (int c = 0; c < DIM; ++c) acc.e1[c] += __shfl_xor_sync(0xffffffffu, acc.e1[c], o);
#pragma unroll
            for (int c = 0; c < NE2; ++c) acc.e2[c] += __shfl_xor_sync(0xffffffffu, acc.e2[c], o);
        }
    }
    const double sc = a.cost.scale();
    if (part == 0) {
        tot[0][fl] = acc.e0 * sc;
        if (FULL) {
#pragma unroll
            for (int c = 0; c < DIM; ++c) tot[1 + c][fl] = acc.e1[c] * sc;
#pragma unroll
            for (int ra = 0; ra < DIM; ++ra)
#pragma unroll
                for (int rb = ra; rb < DIM; ++rb) {
                    const double v = acc.e2[k1s_e2<DIM>(ra, rb)] * sc;
                    tot[1 + DIM + ra + rb * DIM][fl] = v;
                    tot[1 + DIM + rb + ra * DIM][fl] = v;
                }
        }
    }
    __syncthreads();
    // ---- epilogue: Vdmu = R e1 / T, Vddmu = R (e2 - e0 I) R / T (upper triangle mirrored), cost = e0 / T ----
    const int nf = min(K1S_FPC, nwork - f0);
    const int* act = a.active;
    if (!FULL) {
        if (threadIdx.x < nf) {
            const int ff = act ? act[f0 + threadIdx.x] : f0 + (int)threadIdx.x;
            a.fcost[ff] = tot[0][threadIdx.x] * sR[DIM * DIM][threadIdx.x];
        }
        k1s_finish_cta(a);
        return;
    }
    constexpr int NEP = DIM * DIM + DIM + 1;
    for (int idx = threadIdx.x; idx < nf * NEP; idx += K1S_THREADS) {
        const int l = idx / NEP, e = idx - l * NEP;
        const int ff = act ? act[f0 + l] : f0 + l;
        const double invT = sR[DIM * DIM][l];
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
                    t = fma(sR[aa + i * DIM][l], m, t);
                }
                v = fma(t, sR[b + j * DIM][l], v);
            }
            a.fVdd[(size_t)ff * DIM * DIM + e] = v * invT;
        } else if (e < DIM * DIM + DIM) {
            const int i = e - DIM * DIM;
            double v = 0.0;
            for (int aa = 0; aa < DIM; ++aa) v = fma(sR[i + aa * DIM][l], tot[1 + aa][l], v);
            a.fVdmu[(size_t)ff * DIM + i] = v * invT;
        } else {
            a.fcost[ff] = e0 * invT;
        }
    }
    if (a.raw != nullptr) {
        for (int idx = threadIdx.x; idx < nf * NOUT; idx += K1S_THREADS) {
            const int l = idx / NOUT, e = idx - l * NOUT;
            a.raw[(size_t)(act ? act[f0 + l] : f0 + l) * NOUT + e] = tot[e][l];
        }
    }
    k1s_finish_cta(a);
}

}  // namespace gvib200
