// K1G: the fused sigma-point / cost / moment kernel for factors of dimension > 4 (two-state factors of dim 8 / 12, the
// robot functors at dim 6), built on the SPARSITY and SIGN-GROUP structure of the sparse Gauss-Hermite rule.
//
// A node of nwspgr's rule (quadrature/GH/SparseGH/nwspgr.m:108-133) at accuracy level k has at most k - 1 non-zero
// coordinates whatever the dimension -- (12, 4): 2649 nodes, none with more than 3 of its 12 coordinates set -- and the
// nodes come in groups of 2^k sign combinations (+-a_1 .. +-a_k on coordinates c_1 < .. < c_k) that share one weight:
// (12, 4) = 1 + 48 + 198 + 220 groups of 1 / 2 / 4 / 8 nodes.  The generic node loop (k_moments) pays
//   x = mu + S xi             2 DIM^2 flops            here:  y = y0 + sum_i (+-) Y[:, c_i] a_i      YD adds per node (Gray code)
//   e1 += w psi xi            2 DIM                           one FMA per group and non-zero coordinate
//   e2 += w psi xi xi^T       DIM (DIM + 1)                   one FMA per group and pair of non-zero coordinates
// per node (633 flops at DIM = 12 with cost_linear_gp, 72 shared-memory loads: LSU bound at 4 % of the FP64 roof); with the
// groups the sigma point costs YD adds, the three moment sums become Walsh-Hadamard butterflies on the 2^k values of psi
// followed by <= 10 accumulator updates per GROUP.  The sums are the same numbers as SparseGaussHermite::Integrate's
// (quadrature/SparseGaussHermite.h:197-221) in a different order.
//
// "y" are REDUCED coordinates: a cost functor that depends on x only through a linear map y = L x (cost_linear_gp:
// r = Phi th1 - th2, gp/cost_functions.h:36-39; the hinge costs: the leading position coordinates) folds L into the
// per-factor columns Y = L S and y0 = L mu once per factor (GrpReduce below), so the per-node work never touches the
// full state dimension.
//
// Mapping: one warp per factor, groups dealt round-robin to the 32 lanes in order of decreasing size (so that the lanes of
// a warp run groups of the same shape almost always); every lane keeps PRIVATE accumulators for all 1 + DIM + DIM (DIM+1) / 2
// xi-space sums in shared memory ([entry][lane]: conflict-free, dynamically indexed by the group's coordinates), which meet
// in a fixed-order sum at the end (reproducible), followed by the epilogue shared with k_moments.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cost_functors.cuh"

namespace gvib200 {

constexpr int K1G_KMAX = 3;      // non-zero coordinates per group this kernel handles (rules beyond fall back to k_moments)
constexpr int K1G_WARPS = 8;     // factors in flight per CTA

struct GrpTable {
    int dim;
    int n_groups;            // groups with k >= 1, sorted by decreasing k
    int has_origin;          // the rule has a node at xi = 0
    double w0;               // its weight
    const int* hdr;          // device [n_groups]: k | c0 << 4 | c1 << 12 | c2 << 20   (coordinates ascending)
    const double* val;       // device [n_groups][4]: a_0, a_1, a_2 (0 beyond k), w
};

// ---- reduced coordinates of a cost functor: y = L x, YD of them; psi / scale = eval(y) --------------------------------
// default: the functor reads the leading XD coordinates of x
template <class Cost, int DIM>
struct GrpReduce {
    static constexpr int YD = Cost::XD;
    // lanes of the warp fill y0[YD] and Y[c * YD + r] (c < DIM) from mu (factor mean, DIM) and S (column-major DIM x DIM)
    static __device__ __forceinline__ void build(const Cost&, int, const double* __restrict__ mu, const double* __restrict__ S,
                                                 double* y0, double* Y, int lane) {
        for (int e = lane; e < YD * DIM; e += 32) {
            const int c = e / YD, r = e - c * YD;
            Y[e] = __ldg(S + r + c * DIM);
        }
        if (lane < YD) y0[lane] = __ldg(mu + lane);
    }
    struct Params {};
    static __device__ __forceinline__ Params params(const Cost&, int) { return Params(); }
    template <int NP>
    static __device__ __forceinline__ void eval(const Cost& cost, const Params&, int f, const double (&y)[NP][YD], double (&psi)[NP]) {
        typename Cost::Pending pend[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) pend[p] = cost.template begin<false>(y[p], f);
#pragma unroll
        for (int p = 0; p < NP; ++p) psi[p] = cost.finish(pend[p]);
    }
};

// cost_linear_gp: psi = 1/2 r^T Qinv r with r = Phi th1 - th2 = [Phi, -I] x  ->  y = r (DS values instead of 2 DS)
template <int DS>
struct GrpReduce<CostLinearGP<DS>, 2 * DS> {
    static constexpr int YD = DS;
    static constexpr int DIM = 2 * DS;
    using Cost = CostLinearGP<DS>;
    static __device__ __forceinline__ void build(const Cost& cost, int f, const double* __restrict__ mu, const double* __restrict__ S,
                                                 double* y0, double* Y, int lane) {
        const double* Phi = cost.params + (size_t)f * 2 * DS * DS;
        for (int e = lane; e < YD * DIM; e += 32) {
            const int c = e / YD, r = e - c * YD;
            double s = -__ldg(S + (DS + r) + c * DIM);
#pragma unroll
            for (int k = 0; k < DS; ++k) s = fma(__ldg(Phi + r + k * DS), __ldg(S + k + c * DIM), s);
            Y[e] = s;
        }
        if (lane < YD) {
            double s = -__ldg(mu + DS + lane);
#pragma unroll
            for (int k = 0; k < DS; ++k) s = fma(__ldg(Phi + lane + k * DS), __ldg(mu + k), s);
            y0[lane] = s;
        }
    }
    struct Params {
        double q[DS * (DS + 1) / 2];  // Qinv: diagonal entries, and the SUM of the two mirrored off-diagonal entries
    };
    static __device__ __forceinline__ Params params(const Cost& cost, int f) {
        const double* Qi = cost.params + (size_t)f * 2 * DS * DS + DS * DS;
        Params p;
        int idx = 0;
#pragma unroll
        for (int i = 0; i < DS; ++i)
#pragma unroll
            for (int j = i; j < DS; ++j) p.q[idx++] = (i == j) ? __ldg(Qi + i + i * DS) : __ldg(Qi + i + j * DS) + __ldg(Qi + j + i * DS);
        return p;
    }
    template <int NP>
    static __device__ __forceinline__ void eval(const Cost&, const Params& pr, int, const double (&y)[NP][YD], double (&psi)[NP]) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            double q = 0.0;
            int idx = 0;
#pragma unroll
            for (int i = 0; i < DS; ++i) {
                double t = 0.0;
#pragma unroll
                for (int j = i; j < DS; ++j) t = fma(pr.q[idx++], y[p][j], t);
                q = fma(t, y[p][i], q);
            }
            psi[p] = q;
        }
    }
};

template <class Cost>
struct GrpArgs {
    int n;
    int state_dim;
    const int* start;
    const double* mu;
    const double* SR;
    const double* T;
    double* fcost;
    double* fVdmu;
    double* fVdd;
    double* raw;
    Cost cost;
};

template <int DIM>
__host__ __device__ constexpr int k1g_e2(int a, int b) { return 1 + DIM + a * DIM - a * (a - 1) / 2 + (b - a); }  // a <= b

// one group with K non-zero coordinates: psi at its 2^K sign patterns, Walsh sums, accumulator updates
template <int DIM, class Cost, bool FULL, int K>
__device__ __forceinline__ void k1g_group(const Cost& cost, const typename GrpReduce<Cost, DIM>::Params& prm, int f, int hdr,
                                          const double* __restrict__ gv, const double* __restrict__ y0, const double* __restrict__ Y,
                                          double* __restrict__ acc, bool& nz) {
    using Red = GrpReduce<Cost, DIM>;
    constexpr int YD = Red::YD;
    constexpr int NP = 1 << K;
    int c[K];
    double a[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        c[i] = (hdr >> (4 + 8 * i)) & 0xff;
        a[i] = gv[i];
    }
    const double w = gv[3];
    // sign pattern p: bit i set = coordinate c_i negative.  y[p] = y0 + sum_i (+-) Y[:, c_i] a_i
    double y[NP][YD];
    {
        double d[K][YD];
#pragma unroll
        for (int r = 0; r < YD; ++r) {
            double s = y0[r];
#pragma unroll
            for (int i = 0; i < K; ++i) {
                d[i][r] = Y[c[i] * YD + r] * a[i];
                s += d[i][r];
            }
            y[0][r] = s;
        }
#pragma unroll
        for (int p = 1; p < NP; ++p) {
            // flip the lowest set bit of p relative to p with that bit cleared
            int b = 0;
            while (!((p >> b) & 1)) ++b;  // compile-time after unrolling
            const int q = p & ~(1 << b);
#pragma unroll
            for (int r = 0; r < YD; ++r) y[p][r] = y[q][r] - (d[b][r] + d[b][r]);
        }
    }
    double psi[NP];
    Red::template eval<NP>(cost, prm, f, y, psi);
    if (!FULL) {
        double s = psi[0];
#pragma unroll
        for (int p = 1; p < NP; ++p) s += psi[p];
        if (s != 0.0) nz = true;
        acc[0] = fma(w, s, acc[0]);
        return;
    }
    // Walsh-Hadamard butterflies: psi[T] <- sum_p (-1)^{|p & T|} psi[p]
#pragma unroll
    for (int b = 0; b < K; ++b)
#pragma unroll
        for (int p = 0; p < NP; ++p)
            if (!(p & (1 << b))) {
                const double u = psi[p], v = psi[p | (1 << b)];
                psi[p] = u + v;
                psi[p | (1 << b)] = u - v;
            }
    bool any = false;
#pragma unroll
    for (int p = 0; p < NP; ++p) any = any || (psi[p] != 0.0);
    if (!any) return;  // every psi of the group vanished (free space): nothing to add
    nz = true;
    const double A0 = psi[0];
    acc[0] = fma(w, A0, acc[0]);
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const double wa = w * a[i];
        acc[(1 + c[i]) * 32] = fma(wa, psi[1 << i], acc[(1 + c[i]) * 32]);
        acc[k1g_e2<DIM>(c[i], c[i]) * 32] = fma(wa * a[i], A0, acc[k1g_e2<DIM>(c[i], c[i]) * 32]);
#pragma unroll
        for (int j = i + 1; j < K; ++j)
            acc[k1g_e2<DIM>(c[i], c[j]) * 32] = fma(wa * a[j], psi[(1 << i) | (1 << j)], acc[k1g_e2<DIM>(c[i], c[j]) * 32]);
    }
}

template <int DIM>
struct K1GCfg {
    static constexpr int NE2 = DIM * (DIM + 1) / 2;
    static constexpr int NACC = 1 + DIM + NE2;
    static constexpr int NOUT = 1 + DIM + DIM * DIM;
    // per warp: lane-private accumulators [NACC][32], then y0 / Y (YDMAX = DIM), then the totals `my`
    static constexpr int WARP_DOUBLES = NACC * 32 + DIM + DIM * DIM + NOUT + 1;
};

// shared epilogue (same arithmetic as k_moments): my[] = e0, e1[DIM], e2 full (xi-space, scaled) -> raw, Vdmu, Vddmu, cost
template <int DIM, class Args>
__device__ __forceinline__ void k1_epilogue(const Args& a, int f, const double* my, double invT, int lane) {
    constexpr int NOUT = 1 + DIM + DIM * DIM;
    if (a.raw != nullptr)
        for (int e = lane; e < NOUT; e += 32) a.raw[(size_t)f * NOUT + e] = my[e];
    const double* R = a.SR + (size_t)f * 2 * DIM * DIM + DIM * DIM;
    const double e0 = my[0];
    for (int e = lane; e < DIM * DIM + DIM + 1; e += 32) {
        if (e < DIM * DIM) {
            int i = e % DIM, j = e / DIM;
            if (i > j) {  // upper triangle mirrored (ngd/NGDFactorizedBaseGH.h:71-72)
                const int t = i;
                i = j;
                j = t;
            }
            double v = 0.0;
            for (int b = 0; b < DIM; ++b) {
                double t = 0.0;
                for (int aa = 0; aa < DIM; ++aa) {
                    const double m = my[1 + DIM + aa + b * DIM] - (aa == b ? e0 : 0.0);
                    t = fma(__ldg(R + aa + i * DIM), m, t);
                }
                v = fma(t, __ldg(R + b + j * DIM), v);
            }
            a.fVdd[(size_t)f * DIM * DIM + e] = v * invT;
        } else if (e < DIM * DIM + DIM) {
            const int i = e - DIM * DIM;
            double v = 0.0;
            for (int aa = 0; aa < DIM; ++aa) v = fma(__ldg(R + i + aa * DIM), my[1 + aa], v);
            a.fVdmu[(size_t)f * DIM + i] = v * invT;
        } else {
            a.fcost[f] = e0 * invT;
        }
    }
}

template <int DIM, class Cost, bool FULL>
__global__ void __launch_bounds__(K1G_WARPS * 32, 1) k_moments_grp(const __grid_constant__ GrpTable tab, const __grid_constant__ GrpArgs<Cost> a) {
    using Red = GrpReduce<Cost, DIM>;
    using Cfg = K1GCfg<DIM>;
    constexpr int YD = Red::YD;
    constexpr int NACC = Cfg::NACC;
    extern __shared__ __align__(16) double smem[];
    // staged table: values [n_groups][4] doubles, then headers (ints)
    double* sval = smem;
    int* shdr = reinterpret_cast<int*>(sval + (size_t)4 * tab.n_groups);
    double* warp0 = sval + (size_t)4 * tab.n_groups + ((tab.n_groups + 1) / 2);
    for (int i = threadIdx.x; i < 4 * tab.n_groups; i += blockDim.x) sval[i] = __ldg(tab.val + i);
    for (int i = threadIdx.x; i < tab.n_groups; i += blockDim.x) shdr[i] = __ldg(tab.hdr + i);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double* acc_all = warp0 + (size_t)warp * Cfg::WARP_DOUBLES;  // [NACC][32]
    double* acc = acc_all + lane;
    double* y0 = acc_all + NACC * 32;
    double* Y = y0 + DIM;
    double* my = Y + DIM * DIM;
    const double sc = a.cost.scale();
    for (int f = blockIdx.x * nwarps + warp; f < a.n; f += gridDim.x * nwarps) {
        const double* Sp = a.SR + (size_t)f * 2 * DIM * DIM;
        const double* mp = a.mu + (size_t)a.start[f] * a.state_dim;
        __syncwarp();
        Red::build(a.cost, f, mp, Sp, y0, Y, lane);
        const typename Red::Params prm = Red::params(a.cost, f);
        if (FULL) {
#pragma unroll 4
            for (int e = 0; e < NACC; ++e) acc[e * 32] = 0.0;
        } else {
            acc[0] = 0.0;
        }
        __syncwarp();
        bool nz = false;
        if (tab.has_origin && lane == 0) {  // the node at xi = 0
            double yy[1][YD], ps[1];
#pragma unroll
            for (int r = 0; r < YD; ++r) yy[0][r] = y0[r];
            Red::template eval<1>(a.cost, prm, f, yy, ps);
            if (ps[0] != 0.0) nz = true;
            acc[0] = fma(tab.w0, ps[0], acc[0]);
        }
        for (int g = lane; g < tab.n_groups; g += 32) {
            const int hdr = shdr[g];
            const double* gv = sval + 4 * g;
            switch (hdr & 0xf) {
                case 1: k1g_group<DIM, Cost, FULL, 1>(a.cost, prm, f, hdr, gv, y0, Y, acc, nz); break;
                case 2: k1g_group<DIM, Cost, FULL, 2>(a.cost, prm, f, hdr, gv, y0, Y, acc, nz); break;
                default: k1g_group<DIM, Cost, FULL, 3>(a.cost, prm, f, hdr, gv, y0, Y, acc, nz); break;
            }
        }
        nz = __any_sync(0xffffffffu, nz);
        __syncwarp();
        const double invT = 1.0 / __ldg(a.T + f);
        // ---- the 32 private copies of every sum meet in a fixed order (rotated start: conflict-free reads) ----
        if (!FULL) {
            double s = acc[0];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) a.fcost[f] = s * sc * invT;
            continue;
        }
        if (!nz && a.raw == nullptr) {  // free space: psi vanished at every node, every moment is exactly zero
            for (int e = lane; e < DIM * DIM + DIM + 1; e += 32) {
                if (e < DIM * DIM) a.fVdd[(size_t)f * DIM * DIM + e] = 0.0;
                else if (e < DIM * DIM + DIM) a.fVdmu[(size_t)f * DIM + (e - DIM * DIM)] = 0.0;
                else a.fcost[f] = 0.0;
            }
            continue;
        }
        for (int e = lane; e < NACC; e += 32) {
            const double* row = acc_all + e * 32;
            double s = 0.0;
#pragma unroll 8
            for (int j = 0; j < 32; ++j) s += row[(j + lane) & 31];
            s *= sc;
            if (e <= DIM) {
                my[e] = s;
            } else {  // packed upper triangle (ra <= rb) -> full mirrored matrix
                int idx = e - 1 - DIM, ra = 0;
                while (idx >= DIM - ra) {
                    idx -= DIM - ra;
                    ++ra;
                }
                const int rb = ra + idx;
                my[1 + DIM + ra + rb * DIM] = s;
                my[1 + DIM + rb + ra * DIM] = s;
            }
        }
        __syncwarp();
        k1_epilogue<DIM>(a, f, my, invT, lane);
    }
}

}  // namespace gvib200
