// CUDA kernels of the factorized NGD-GVI iteration (sm_100a, FP64).
//   k_prologue   K2  per-factor marginal extraction + PSD sqrt / inverse sqrt (Jacobi in registers)
//   k_moments    K1  fused sigma points + cost functor + moment reduction + Vdmu/Vddmu epilogue
//   k_linear         closed-form linear-Gaussian factors (gradient + cost)
//   k_assemble   K3  deterministic gather of factor blocks into the block-tridiagonal joint
//   k_cr_*       K4  tile-wise block cyclic reduction: block-tridiagonal Cholesky / solve / selected inverse / log det
//   k_candidate      line-search candidate (mu + a dmu, Lambda + a dLambda)
//   k_total_cost     sum of factor costs + 1/2 log det, fixed summation order
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bt_cr.h"
#include "cost_functors.cuh"
#include "smallmat.h"

namespace gvib200 {

// ------------------------------------------------------------------------------------------
// TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// ------------------------------------------------------------------------------------------
// K2: per-factor prologue.  Extract Sigma_k from the block-tridiagonal covariance
// (GVIFactorizedBase::update_precision_from_joint, gvibase/GVIFactorizedBase.h:111-114 +
// TrajectoryBlock::extract helpers/MatrixHelper.h:132-134) and form S = Sigma^1/2 (symmetric PSD root,
// quadrature/SparseGaussHermite.h:231-233) and R = Sigma^-1/2 (so that P_k = R R).
// One thread per factor; SR[f] = {S[DIM*DIM], R[DIM*DIM]} column-major.
// ------------------------------------------------------------------------------------------
template <int DIM, int SD>
__global__ void k_prologue(int n, const int* __restrict__ start, const double* __restrict__ covD,
                           const double* __restrict__ covO, double* __restrict__ SR) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    constexpr int NS = DIM / SD;  // states spanned (1 or 2)
    static_assert(NS == 1 || NS == 2, "factors span one or two consecutive states");
    const int s = start[f];
    Mat<DIM> Sig, S, R;
#pragma unroll
    for (int j = 0; j < SD; ++j)
#pragma unroll
        for (int i = 0; i < SD; ++i) {
            Sig(i, j) = covD[(size_t)s * SD * SD + i + j * SD];
            if (NS == 2) {
                Sig(SD + i, SD + j) = covD[(size_t)(s + 1) * SD * SD + i + j * SD];
                const double o = covO[(size_t)s * SD * SD + i + j * SD];  // block (s, s+1)
                Sig(i, SD + j) = o;
                Sig(SD + j, i) = o;
            }
        }
    sqrt_and_invsqrt<DIM>(S, R, Sig);
    double* out = SR + (size_t)f * 2 * DIM * DIM;
#pragma unroll
    for (int e = 0; e < DIM * DIM; ++e) {
        out[e] = S.a[e];
        out[DIM * DIM + e] = R.a[e];
    }
}

// Group-cooperative symmetric eigendecomposition (parallel Jacobi, round-robin ordering): JG lanes work on one matrix, a warp
// holds 32 / JG matrices.  A, V are N x N column-major in shared memory (one copy per group), on return V holds the
// eigenvectors and the diagonal of A the eigenvalues.  A round applies N/2 rotations on disjoint index pairs at once: the
// angles come from one state of the matrix, then the group's lanes split the column update A J (and V J) and the row update
// J^T (A J).  For the factor dimensions above 4 (two-state factors: 8, 12; robot states: 6) one thread per factor spends its
// time spilling a 12 x 12 matrix and leaves most of the machine idle (10^4 factors = 2 warps per SM); with 8 lanes per
// factor the matrices live in shared memory and every lane has work in every phase.
constexpr int JG = 32;  // lanes per matrix (measured at dim 12, 10^4 factors: 32 lanes 0.65 ms, 8 lanes 0.80 ms, one thread 1.26 ms per launch)
template <int N>
__device__ __forceinline__ void group_jacobi_eig(double* __restrict__ A, double* __restrict__ V, double* __restrict__ rot, int gl) {
    constexpr int M = (N + 1) / 2;      // pairs per round
    constexpr int NR = 2 * M - 1;       // rounds per sweep (odd N: one index sits out per round)
    for (int e = gl; e < N * N; e += JG) V[e] = (e % N == e / N) ? 1.0 : 0.0;
    __syncwarp();
    // Quadratic convergence: a 12 x 12 SPD matrix is at rounding level after 5-6 sweeps; past that point rounding noise
    // (~1e-17 relative) keeps re-filling the annihilated entries above jacobi_angle's `tiny` threshold, so the loop is
    // bounded by a sweep count, with the exact exits for the easy cases (already diagonal / a sweep of identity rotations)
    for (int sweep = 0; sweep < 8; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int e = gl; e < N * N; e += JG) {
            const int i = e % N, j = e / N;
            const double v = fabs(A[e]);
            if (i == j) diag += v;
            else if (i < j) off += v;
        }
#pragma unroll
        for (int o = JG / 2; o > 0; o >>= 1) {
            off += __shfl_xor_sync(0xffffffffu, off, o);
            diag += __shfl_xor_sync(0xffffffffu, diag, o);
        }
        // the groups of a warp move in lock step: leave when all of them have converged (a converged matrix only sees
        // identity rotations)
        if (__all_sync(0xffffffffu, off <= 1e-300 || off <= 4e-17 * diag)) break;
        bool rotated = false;  // some pair of this sweep was not already negligible (jacobi_angle's `tiny`)
        for (int r = 0; r < NR; ++r) {
            // circle method on 2 M positions (position 2 M - 1 is fixed; for odd N it is the idle slot)
            for (int k = gl; k < M; k += JG) {
                int p, q;
                if (k == 0) {
                    p = 2 * M - 1;
                    q = r;
                } else {
                    p = (r + k) % NR;
                    q = (r - k + NR) % NR;
                }
                if (p > q) {
                    const int t = p;
                    p = q;
                    q = t;
                }
                double t = 0.0, c = 1.0, sn = 0.0;
                if (q < N) jacobi_angle(A[p + p * N], A[q + q * N], A[p + q * N], t, c, sn);
                else p = q = -1;  // idle pair
                rotated = rotated || (t != 0.0);
                rot[4 * k + 0] = (double)p;
                rot[4 * k + 1] = (double)q;
                rot[4 * k + 2] = c;
                rot[4 * k + 3] = sn;
            }
            __syncwarp();
            // column update of A and V: (x_jp, x_jq) <- (c x_jp - s x_jq, s x_jp + c x_jq)
            for (int task = gl; task < M * N; task += JG) {
                const int k = task / N, j = task - k * N;
                const int p = (int)rot[4 * k], q = (int)rot[4 * k + 1];
                if (p < 0) continue;
                const double c = rot[4 * k + 2], sn = rot[4 * k + 3];
                const double ap = A[j + p * N], aq = A[j + q * N];
                A[j + p * N] = fma(c, ap, -sn * aq);
                A[j + q * N] = fma(sn, ap, c * aq);
                const double vp = V[j + p * N], vq = V[j + q * N];
                V[j + p * N] = fma(c, vp, -sn * vq);
                V[j + q * N] = fma(sn, vp, c * vq);
            }
            __syncwarp();
            // row update of A; the annihilated entries are set to exactly zero
            for (int task = gl; task < M * N; task += JG) {
                const int k = task / N, j = task - k * N;
                const int p = (int)rot[4 * k], q = (int)rot[4 * k + 1];
                if (p < 0) continue;
                const double c = rot[4 * k + 2], sn = rot[4 * k + 3];
                const double ap = A[p + j * N], aq = A[q + j * N];
                double np_ = fma(c, ap, -sn * aq), nq = fma(sn, ap, c * aq);
                if (j == q) np_ = 0.0;
                if (j == p) nq = 0.0;
                A[p + j * N] = np_;
                A[q + j * N] = nq;
            }
            __syncwarp();
        }
        // converged: a whole sweep of identity rotations (the off-diagonal part sits at rounding level and stays there)
        if (!__any_sync(0xffffffffu, rotated)) break;
        // keep A exactly symmetric (the two one-sided updates round differently)
        for (int e = gl; e < N * N; e += JG) {
            const int i = e % N, j = e / N;
            if (i < j) {
                const double v = 0.5 * (A[i + j * N] + A[j + i * N]);
                A[i + j * N] = v;
                A[j + i * N] = v;
            }
        }
        __syncwarp();
    }
}

// K2 for factor dimensions above 4: JG lanes per factor (group_jacobi_eig), same outputs as k_prologue
template <int DIM, int SD>
__global__ void __launch_bounds__(256) k_prologue_warp(int n, const int* __restrict__ start, const double* __restrict__ covD,
                                                       const double* __restrict__ covO, double* __restrict__ SR) {
    constexpr int NS = DIM / SD;
    static_assert(NS == 1 || NS == 2, "factors span one or two consecutive states");
    constexpr int WD = 2 * DIM * DIM + 4 * ((DIM + 1) / 2) + 2 * DIM + 1;  // odd stride: the groups of a warp hit different banks
    constexpr int GPB = 256 / JG;                                           // factors per CTA
    extern __shared__ __align__(16) double sm[];
    const int gl = threadIdx.x & (JG - 1), grp = threadIdx.x / JG;
    // tail groups recompute the last factor and do not store (the groups of a warp run in lock step)
    const int fraw = blockIdx.x * GPB + grp;
    const int f = fraw < n ? fraw : n - 1;
    double* A = sm + (size_t)grp * WD;
    double* V = A + DIM * DIM;
    double* rot = V + DIM * DIM;
    double* sq = rot + 4 * ((DIM + 1) / 2);
    double* isq = sq + DIM;
    const int s = start[f];
    for (int e = gl; e < DIM * DIM; e += JG) {
        const int i = e % DIM, j = e / DIM;
        const int bi = i / SD, bj = j / SD, li = i - bi * SD, lj = j - bj * SD;
        double v;
        if (bi == bj) v = covD[(size_t)(s + bi) * SD * SD + li + lj * SD];
        else if (bi < bj) v = covO[(size_t)s * SD * SD + li + lj * SD];   // block (s, s+1)
        else v = covO[(size_t)s * SD * SD + lj + li * SD];                 // its transpose
        A[e] = v;
    }
    __syncwarp();
    // the Jacobi iteration works on the symmetric part
    for (int e = gl; e < DIM * DIM; e += JG) {
        const int i = e % DIM, j = e / DIM;
        if (i < j) {
            const double v = 0.5 * (A[i + j * DIM] + A[j + i * DIM]);
            A[i + j * DIM] = v;
            A[j + i * DIM] = v;
        }
    }
    __syncwarp();
    group_jacobi_eig<DIM>(A, V, rot, gl);
    for (int k = gl; k < DIM; k += JG) {
        const double l = sqrt(A[k + k * DIM]);
        sq[k] = l;
        isq[k] = 1.0 / l;
    }
    __syncwarp();
    if (fraw >= n) return;
    double* out = SR + (size_t)f * 2 * DIM * DIM;
    for (int e = gl; e < DIM * DIM; e += JG) {
        const int i = e % DIM, j = e / DIM;
        const int lo = i < j ? i : j, hi = i < j ? j : i;  // one arithmetic for (i, j) and (j, i): exactly symmetric
        double sv = 0.0, rv = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            const double vv = V[hi + k * DIM] * V[lo + k * DIM];
            sv = fma(vv, sq[k], sv);
            rv = fma(vv, isq[k], rv);
        }
        out[e] = sv;
        out[DIM * DIM + e] = rv;
    }
}

// ------------------------------------------------------------------------------------------
// K1: fused sigma-point / cost / moment kernel.  One warp per factor, lanes stride the nodes.
//   x_i = mu_k + S_k xi_i                      quadrature/SparseGaussHermite.h:242
//   e0 = sum w psi, e1 = sum w psi xi, e2 = sum w psi xi xi^T   (xi-space; E1 = S e1, E2 = S e2 S,
//        identical to the x-space sums of ngd/NGDFactorizedBaseGH.h:46-48)
//   Vdmu = P E1 / T = R e1 / T;  Vddmu = (P E2 P - P E0)/T = R (e2 - e0 I) R / T  (:61-73)
//   cost = E0 / T                                                                    (:122-129)
// The node table (xi | w rows) is staged into shared memory with a TMA bulk copy; when it does not fit
// it is streamed in chunks.  The warp reduction is a fixed xor-butterfly, so results are reproducible.
// ------------------------------------------------------------------------------------------
template <class Cost>
struct MomentArgs {
    int n;                 // factors in this group
    int n_nodes;           // quadrature nodes, padded with zero-weight nodes to a multiple of 32
    int chunk;             // nodes per shared-memory chunk (multiple of 32; >= n_nodes: single stage)
    int state_dim;
    const double* table;   // planes: NP x [n_nodes] double2 (xi_2p, xi_2p+1), then [n_nodes] weights
    const int* start;      // [n] start state
    const double* mu;      // joint mean
    const double* SR;      // [n][2*DIM*DIM]
    const double* T;       // [n] temperatures
    double* fcost;         // [n] E0 / T
    double* fVdmu;         // [n][DIM]
    double* fVdd;          // [n][DIM*DIM]
    double* raw;           // optional [n][1 + DIM + DIM*DIM]: e0, e1, e2 (xi-space, times scale)
    double ximax[12];      // max |xi_c| over the table (bounding box of the sigma points)
    Cost cost;
};

template <int DIM>
struct MomentAcc {
    static constexpr int NE2 = DIM * (DIM + 1) / 2;
    double e0;
    double e1[DIM];
    double e2[NE2];  // packed upper triangle, (a, b) a <= b at index a*DIM - a(a-1)/2 + (b - a)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// threads per CTA / minimum resident CTAs: small factors keep everything in registers
template <int DIM>
struct K1Cfg {
    static constexpr int THREADS = (DIM <= 4) ? 256 : 128;
    static constexpr int MIN_BLOCKS = (DIM <= 4) ? 2 : 1;
};
// per-warp scratch: epilogue totals (+ the S rows when they do not fit in registers)
template <int DIM, int XD>
struct K1Scratch {
    static constexpr bool S_IN_SMEM = (XD * DIM > 32);
    static constexpr int NOUT = 1 + DIM + DIM * DIM;
    static constexpr int DOUBLES = NOUT + (S_IN_SMEM ? XD * DIM : 0);
};

// One node: xi (from the shared planes), x = mu + S xi, first half of the cost evaluation.
template <int DIM, class Cost, bool FAST, bool S_IN_SMEM, int NSREG>
__device__ __forceinline__ typename Cost::Pending k1_begin(const double2* __restrict__ planes, const double* __restrict__ wts,
                                                           int plane_stride, int i, const double (&mu)[Cost::XD],
                                                           const double (&S)[NSREG][DIM], const double* __restrict__ sS,
                                                           const Cost& cost, int f, double (&xi)[DIM], double& w) {
    constexpr int NP = (DIM + 1) / 2;
    constexpr int XD = Cost::XD;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const double2 v = planes[(size_t)p * plane_stride + i];
        xi[2 * p] = v.x;
        if (2 * p + 1 < DIM) xi[2 * p + 1] = v.y;
    }
    w = wts[i];
    double x[XD];
#pragma unroll
    for (int r = 0; r < XD; ++r) {
        double s = mu[r];
#pragma unroll
        for (int c = 0; c < DIM; ++c) s = fma(S_IN_SMEM ? sS[r * DIM + c] : S[S_IN_SMEM ? 0 : r][c], xi[c], s);
        x[r] = s;
    }
    return cost.template begin<FAST>(x, f);
}

template <int DIM, bool FULL>
__device__ __forceinline__ void k1_accumulate(MomentAcc<DIM>& acc, bool& nz, const double (&xi)[DIM], double w,
                                              double psi) {
    const double p = w * psi;
    acc.e0 += p;
    if (FULL) {
        // whole-warp skip of free-space nodes (psi == 0 for every lane): exact, the terms are zeros
        if (__any_sync(0xffffffffu, p != 0.0)) {
            nz = true;
            int idx = 0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                const double q = p * xi[c];
                acc.e1[c] += q;
#pragma unroll
                for (int d2 = c; d2 < DIM; ++d2) {
                    acc.e2[idx] = fma(q, xi[d2], acc.e2[idx]);
                    ++idx;
                }
            }
        }
    }
}

// Software-pipelined node loop over one shared-memory chunk: the gather of node i+32 is issued
// before node i is finished.  cn is a multiple of 32, so the loop is warp-uniform.
template <int DIM, class Cost, bool FULL, bool FAST, bool S_IN_SMEM, int NSREG>
__device__ __forceinline__ void k1_node_loop(MomentAcc<DIM>& acc, bool& nz, const double2* __restrict__ planes,
                                             const double* __restrict__ wts, int plane_stride, int cn, int lane,
                                             const double (&mu)[Cost::XD], const double (&S)[NSREG][DIM],
                                             const double* __restrict__ sS, const Cost& cost, int f) {
    double xiA[DIM], wA;
    typename Cost::Pending pa =
        k1_begin<DIM, Cost, FAST, S_IN_SMEM, NSREG>(planes, wts, plane_stride, lane, mu, S, sS, cost, f, xiA, wA);
    for (int i = lane + 32; i < cn; i += 32) {
        double xiB[DIM], wB;
        typename Cost::Pending pb =
            k1_begin<DIM, Cost, FAST, S_IN_SMEM, NSREG>(planes, wts, plane_stride, i, mu, S, sS, cost, f, xiB, wB);
        k1_accumulate<DIM, FULL>(acc, nz, xiA, wA, cost.finish(pa));
#pragma unroll
        for (int c = 0; c < DIM; ++c) xiA[c] = xiB[c];
        wA = wB;
        pa = pb;
    }
    k1_accumulate<DIM, FULL>(acc, nz, xiA, wA, cost.finish(pa));
}

template <int DIM, class Cost, bool FULL>
__global__ void __launch_bounds__(K1Cfg<DIM>::THREADS, K1Cfg<DIM>::MIN_BLOCKS) k_moments(const MomentArgs<Cost> a) {
    constexpr int NP = (DIM + 1) / 2;
    constexpr int XD = Cost::XD;
    constexpr int NE2 = DIM * (DIM + 1) / 2;
    constexpr int NOUT = 1 + DIM + DIM * DIM;
    constexpr bool S_IN_SMEM = K1Scratch<DIM, XD>::S_IN_SMEM;
    constexpr int NSREG = S_IN_SMEM ? 1 : XD;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t mbar;
    const double2* planes = reinterpret_cast<const double2*>(smem);   // [NP][chunk]
    const double* wts = smem + (size_t)2 * NP * a.chunk;              // [chunk]
    double* scratch = smem + (size_t)(2 * NP + 1) * a.chunk;          // [warps][K1Scratch::DOUBLES]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    double* my = scratch + warp * K1Scratch<DIM, XD>::DOUBLES;
    double* sS = my + NOUT;  // [XD][DIM] row r at sS[r*DIM + c] (only when S_IN_SMEM)

    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    __syncthreads();
    unsigned parity = 0;
    const bool single = a.n_nodes <= a.chunk;
    // stage nodes [c0, c0 + cn) of every plane: NP + 1 bulk copies completing on one mbarrier
    auto stage_chunk = [&](int c0, int cn) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&mbar, (unsigned)(cn * (2 * NP + 1) * sizeof(double)));
#pragma unroll
            for (int p = 0; p < NP; ++p)
                bulk_copy_g2s(smem + (size_t)2 * p * a.chunk, a.table + (size_t)2 * p * a.n_nodes + (size_t)2 * c0,
                              (unsigned)(cn * 2 * sizeof(double)), &mbar);
            bulk_copy_g2s(smem + (size_t)2 * NP * a.chunk, a.table + (size_t)2 * NP * a.n_nodes + c0,
                          (unsigned)(cn * sizeof(double)), &mbar);
        }
        mbar_wait(&mbar, parity);
        parity ^= 1;
    };
    if (single) stage_chunk(0, a.n_nodes);

    const int stride = gridDim.x * nwarps;
    for (int base = blockIdx.x * nwarps; base < a.n; base += stride) {
        const int f = base + warp;
        const bool active = f < a.n;
        // factor data: mu_k and the first XD rows of S_k
        double mu[XD], S[NSREG][DIM];
        bool fast = false;
        if (active) {
            const double* Sp = a.SR + (size_t)f * 2 * DIM * DIM;
            const double* mp = a.mu + (size_t)a.start[f] * a.state_dim;
#pragma unroll
            for (int r = 0; r < XD; ++r) mu[r] = __ldg(mp + r);
            if (S_IN_SMEM) {
                __syncwarp();
                for (int e = lane; e < XD * DIM; e += 32) sS[e] = __ldg(Sp + (e / DIM) + (e % DIM) * DIM);
                __syncwarp();
            } else {
                double lo[XD], hi[XD];
#pragma unroll
                for (int r = 0; r < NSREG; ++r) {
                    double rad = 0.0;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) {
                        S[r][c] = __ldg(Sp + r + c * DIM);
                        rad = fma(fabs(S[r][c]), a.ximax[c], rad);
                    }
                    lo[r] = mu[r] - rad;
                    hi[r] = mu[r] + rad;
                }
                fast = a.cost.fast_ok(lo, hi);
            }
        }
        MomentAcc<DIM> acc;
        bool nz = false;  // warp-uniform: some node of this factor had psi != 0
        acc.e0 = 0.0;
#pragma unroll
        for (int c = 0; c < DIM; ++c) acc.e1[c] = 0.0;
#pragma unroll
        for (int c = 0; c < NE2; ++c) acc.e2[c] = 0.0;

        for (int c0 = 0; c0 < a.n_nodes; c0 += a.chunk) {
            const int cn = min(a.chunk, a.n_nodes - c0);
            if (!single) {
                __syncthreads();  // every warp is done with the previous chunk
                stage_chunk(c0, cn);
            }
            if (active) {
                if (fast)
                    k1_node_loop<DIM, Cost, FULL, true, S_IN_SMEM, NSREG>(acc, nz, planes, wts, a.chunk, cn, lane, mu, S, sS,
                                                                           a.cost, f);
                else
                    k1_node_loop<DIM, Cost, FULL, false, S_IN_SMEM, NSREG>(acc, nz, planes, wts, a.chunk, cn, lane, mu, S, sS,
                                                                            a.cost, f);
            }
        }
        if (!active) continue;  // whole warp
        // ---- fixed-order butterfly reduction ----
        const double sc = a.cost.scale();
        acc.e0 = warp_sum(acc.e0) * sc;
        const double invT = 1.0 / __ldg(a.T + f);
        if (!FULL) {
            if (lane == 0) a.fcost[f] = acc.e0 * invT;
            continue;
        }
        if (!nz && a.raw == nullptr) {
            // free space: psi vanished at every node, so every moment is exactly zero
            for (int e = lane; e < DIM * DIM + DIM + 1; e += 32) {
                if (e < DIM * DIM) a.fVdd[(size_t)f * DIM * DIM + e] = 0.0;
                else if (e < DIM * DIM + DIM) a.fVdmu[(size_t)f * DIM + (e - DIM * DIM)] = 0.0;
                else a.fcost[f] = 0.0;
            }
            continue;
        }
#pragma unroll
        for (int c = 0; c < DIM; ++c) acc.e1[c] = warp_sum(acc.e1[c]) * sc;
#pragma unroll
        for (int c = 0; c < NE2; ++c) acc.e2[c] = warp_sum(acc.e2[c]) * sc;
        // ---- epilogue: lane 0 publishes the totals, lanes split the small products ----
        if (lane == 0) {
            my[0] = acc.e0;
            int idx = 0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                my[1 + c] = acc.e1[c];
#pragma unroll
                for (int d2 = c; d2 < DIM; ++d2) {
                    my[1 + DIM + c + d2 * DIM] = acc.e2[idx];
                    my[1 + DIM + d2 + c * DIM] = acc.e2[idx];
                    ++idx;
                }
            }
        }
        __syncwarp();
        if (a.raw != nullptr) {
            for (int e = lane; e < NOUT; e += 32) a.raw[(size_t)f * NOUT + e] = my[e];
        }
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
                // (R (e2 - e0 I) R)_{ij} = sum_b (sum_a R_ai M_ab) R_bj
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
        __syncwarp();
    }
}

// x-space moments for the parity probe (gvib200_moments): E1 = S e1, E2 = S e2 S.
template <int DIM>
__global__ void k_raw_to_x(int n, const double* __restrict__ raw, const double* __restrict__ SR, double* __restrict__ E0,
                           double* __restrict__ E1, double* __restrict__ E2) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    constexpr int NOUT = 1 + DIM + DIM * DIM;
    const double* r = raw + (size_t)f * NOUT;
    const double* S = SR + (size_t)f * 2 * DIM * DIM;
    if (E0) E0[f] = r[0];
    if (E1) {
        for (int i = 0; i < DIM; ++i) {
            double s = 0.0;
            for (int k = 0; k < DIM; ++k) s = fma(S[i + k * DIM], r[1 + k], s);
            E1[(size_t)f * DIM + i] = s;
        }
    }
    if (E2) {
        for (int j = 0; j < DIM; ++j)
            for (int i = 0; i < DIM; ++i) {
                double v = 0.0;
                for (int b = 0; b < DIM; ++b) {
                    double t = 0.0;
                    for (int aa = 0; aa < DIM; ++aa) t = fma(S[i + aa * DIM], r[1 + DIM + aa + b * DIM], t);
                    v = fma(t, S[b + j * DIM], v);
                }
                E2[(size_t)f * DIM * DIM + i + j * DIM] = v;
            }
    }
}

// ------------------------------------------------------------------------------------------
// Closed-form linear-Gaussian factors (ngd/NGDFactorizedLinear.h:93-129).  One thread per factor.
//   r = Lambda mu_k - Psi mu_t;  Vdmu = 2 C Lambda^T Kinv r / T
//   cost = C (tr(A Sigma_k) + r^T Kinv r) / T,  A = Lambda^T Kinv Lambda
// Vddmu = 2 C A / T is state independent (the reference evaluates it through a 4-th moment loop that is
// algebraically the same, :107-119) and is pre-assembled once into the constant block-tridiagonal Klin.
// ------------------------------------------------------------------------------------------
constexpr int LIN_MAX_DIM = 12;
// per-factor arrays are element-major ("structure of arrays": element e of factor f at [e * n + f]) so that the
// thread-per-factor kernel reads consecutive addresses across a warp
struct LinearArgs {
    int n, dim, m, state_dim;
    const int* start;
    const double* Lambda;  // [m*dim][n]   (column-major element order within a factor)
    const double* psi;     // [m][n] = Psi mu_t
    const double* Kinv;    // [m*m][n]
    const double* A;       // [dim*(dim+1)/2][n]: upper triangle of Lambda^T Kinv Lambda, packed row by row
    const double* C;       // [n]
    const double* T;       // [n]
    const double* mu;      // joint mean
    const double* covD;
    const double* covO;
    double* fcost;   // [n]
    double* fVdmu;   // [n][dim] or null (cost only)
    // The closed form has a covariance part, (C/T) tr(A Sigma_k), and a mean part, (C/T) r^T Kinv r with r = Lambda mu - psi
    // (and Vdmu).  part = 1 writes the covariance part into fcost, part = 2 adds the mean part to it and writes Vdmu: the
    // two halves need different inputs (candidate covariance / candidate mean), which become available at different
    // times of an iteration.  Every caller runs part 1 then part 2, so all paths share one arithmetic.
    int part;
};

// DIM_, M_, SD_ > 0: compile-time shapes (fully unrolled); 0: taken from the arguments
template <int DIM_, int M_, int SD_>
__global__ void __launch_bounds__(128) k_linear(const LinearArgs a) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.n) return;
    const int dim = DIM_ ? DIM_ : a.dim, m = M_ ? M_ : a.m, sd = SD_ ? SD_ : a.state_dim;
    constexpr int MAXD = DIM_ ? DIM_ : LIN_MAX_DIM, MAXM = M_ ? M_ : LIN_MAX_DIM;
    const size_t n = (size_t)a.n;
    const int s = a.start[f];
    const double c_over_t = a.C[f] / a.T[f];
    double cov_part = 0.0;  // part 3: both halves in one pass, the covariance part stays in a register (same arithmetic:
                            // the mean part is added to the ROUNDED product tr * c_over_t either way)
    if (a.part == 1 || a.part == 3) {
        // tr(A Sigma_k) = sum_i A_ii S_ii + 2 sum_{i<j} A_ij S_ij, Sigma_k assembled from the covariance blocks
        const double* A = a.A + f;
        double tr = 0.0;
        int e = 0;
#pragma unroll
        for (int i = 0; i < MAXD; ++i)
#pragma unroll
            for (int j = i; j < MAXD; ++j)
                if (i < dim && j < dim) {
                    const int bi = i / sd, ii = i % sd, bj = j / sd, jj = j % sd;
                    const double sij = (bi == bj) ? a.covD[(size_t)(s + bi) * sd * sd + ii + jj * sd]
                                                  : a.covO[(size_t)s * sd * sd + ii + jj * sd];  // bi < bj: block (s, s+1)
                    const double av = A[(size_t)e * n];
                    tr = fma(i == j ? av : 2.0 * av, sij, tr);
                    ++e;
                }
        cov_part = tr * c_over_t;
        if (a.part == 1) {
            a.fcost[f] = cov_part;
            return;
        }
    }
    const double* L = a.Lambda + f;
    const double* Ki = a.Kinv + f;
    double mu[MAXD], r[MAXM], kr[MAXM];
#pragma unroll
    for (int k = 0; k < MAXD; ++k)
        if (k < dim) mu[k] = a.mu[(size_t)s * sd + k];
#pragma unroll
    for (int i = 0; i < MAXM; ++i)
        if (i < m) {
            double v = -a.psi[(size_t)i * n + f];
#pragma unroll
            for (int k = 0; k < MAXD; ++k)
                if (k < dim) v = fma(L[(size_t)(i + k * m) * n], mu[k], v);
            r[i] = v;
        }
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < MAXM; ++i)
        if (i < m) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < MAXM; ++k)
                if (k < m) v = fma(Ki[(size_t)(i + k * m) * n], r[k], v);
            kr[i] = v;
            q = fma(v, r[i], q);
        }
    if (a.fVdmu != nullptr) {
#pragma unroll
        for (int k = 0; k < MAXD; ++k)
            if (k < dim) {
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < MAXM; ++i)
                    if (i < m) v = fma(L[(size_t)(i + k * m) * n], kr[i], v);
                a.fVdmu[(size_t)f * dim + k] = 2.0 * v * c_over_t;
            }
    }
    a.fcost[f] = fma(q, c_over_t, a.part == 3 ? cov_part : a.fcost[f]);
}

// ------------------------------------------------------------------------------------------
// Prox-GVI (proxgd/): per-factor Bures-Wasserstein JKO epilogues.  One thread per factor.
//   GH factors   (ProxGVIFactorizedBaseGH.h:153-161): b_k = P E1 = R e1,  S_k = P E2 P - P E0 = R (e2 - e0 I) R
//                from the xi-space moments K1 left in `raw`; cost = E0 (not divided by the temperature, :250-257)
//   linear       (ProxGVIFactorizedLinear.h:95-104): b_k = Lambda^T Kinv (Lambda mu - Psi mu_t), S_k = Lambda^T Kinv Lambda
//   both:        Vdmu = -b_k,  Vddmu = (inv(Sigma+) - P)/eta  (bw_jko, smallmat.h)
// ------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(64) k_prox_gh(int n, double eta, const double* __restrict__ raw, const double* __restrict__ SR,
                                                double* __restrict__ fcost, double* __restrict__ fVdmu,
                                                double* __restrict__ fVdd) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    constexpr int NOUT = 1 + DIM + DIM * DIM;
    const double* r = raw + (size_t)f * NOUT;
    Mat<DIM> S, R, E, Sig, P, Sk, T, Vdd;
    mat_load<DIM>(S, SR + (size_t)f * 2 * DIM * DIM);
    mat_load<DIM>(R, SR + (size_t)f * 2 * DIM * DIM + DIM * DIM);
    const double e0 = r[0];
#pragma unroll
    for (int j = 0; j < DIM; ++j)
#pragma unroll
        for (int i = 0; i < DIM; ++i) E(i, j) = r[1 + DIM + i + j * DIM] - (i == j ? e0 : 0.0);
    mm<DIM>(T, R, E);
    mm<DIM>(Sk, T, R);
    symmetrize<DIM>(Sk);
    mm<DIM>(Sig, S, S);
    mm<DIM>(P, R, R);
    bw_jko<DIM>(Vdd, Sig, P, Sk, eta);
    mat_store<DIM>(fVdd + (size_t)f * DIM * DIM, Vdd);
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
        double b = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) b = fma(R(i, k), r[1 + k], b);
        fVdmu[(size_t)f * DIM + i] = -b;
    }
    fcost[f] = e0;
}

// The same epilogue for factor dimensions above 4 with one WARP per factor (the thread-per-factor form keeps eight 12 x 12
// matrices per thread in local memory: 0.66 ms per launch at 10^4 dim-12 factors): the matrices live in shared memory,
// the lanes split the entries of every product, the eigen-decomposition of H is the warp-parallel Jacobi of the prologue
// (group_jacobi_eig).  Same formulas, the products in the same k order.
template <int DIM>
__global__ void __launch_bounds__(256) k_prox_gh_warp(int n, double eta, const double* __restrict__ raw, const double* __restrict__ SR,
                                                      double* __restrict__ fcost, double* __restrict__ fVdmu,
                                                      double* __restrict__ fVdd) {
    constexpr int DD = DIM * DIM;
    constexpr int NOUT = 1 + DIM + DD;
    constexpr int WD = 5 * DD + 4 * ((DIM + 1) / 2) + DIM + 1;  // odd stride
    constexpr int GPB = 256 / JG;
    extern __shared__ __align__(16) double sm[];
    const int gl = threadIdx.x & (JG - 1), grp = threadIdx.x / JG;
    const int fraw = blockIdx.x * GPB + grp;
    const int f = fraw < n ? fraw : n - 1;  // tail groups recompute the last factor and do not store
    double* S = sm + (size_t)grp * WD;
    double* R = S + DD;
    double* W1 = R + DD;
    double* W2 = W1 + DD;
    double* W3 = W2 + DD;
    double* rot = W3 + DD;
    double* lam = rot + 4 * ((DIM + 1) / 2);
    const double* r = raw + (size_t)f * NOUT;
    const double e0 = r[0];
    for (int e = gl; e < DD; e += JG) {
        S[e] = SR[(size_t)f * 2 * DD + e];
        R[e] = SR[(size_t)f * 2 * DD + DD + e];
        W1[e] = r[1 + DIM + e] - ((e % DIM) == (e / DIM) ? e0 : 0.0);  // E = e2 - e0 I
    }
    __syncwarp();
    auto mul = [&](double* C, const double* A, const double* B, bool bt) {  // C = A B (bt: A B^T), lanes split the entries
        for (int e = gl; e < DD; e += JG) {
            const int i = e % DIM, j = e / DIM;
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) v = fma(A[i + k * DIM], bt ? B[j + k * DIM] : B[k + j * DIM], v);
            C[e] = v;
        }
        __syncwarp();
    };
    auto sym = [&](double* A) {
        for (int e = gl; e < DD; e += JG) {
            const int i = e % DIM, j = e / DIM;
            if (i > j) {
                const double v = 0.5 * (A[i + j * DIM] + A[j + i * DIM]);
                A[i + j * DIM] = v;
                A[j + i * DIM] = v;
            }
        }
        __syncwarp();
    };
    mul(W2, R, W1, false);   // R E
    mul(W3, W2, R, false);   // S_k = R E R
    sym(W3);
    for (int e = gl; e < DD; e += JG) W2[e] = ((e % DIM) == (e / DIM) ? 1.0 : 0.0) - eta * W3[e];  // M = I - eta S_k
    __syncwarp();
    mul(W1, S, S, false);    // Sigma = S S
    mul(W3, W2, W1, false);  // T = M Sigma
    mul(W1, W3, W2, true);   // H = T M^T
    sym(W1);
    group_jacobi_eig<DIM>(W1, W3, rot, gl);  // eigenvalues on the diagonal of W1, eigenvectors in W3
    for (int k = gl; k < DIM; k += JG) {
        const double l = W1[k + k * DIM];
        const double disc = fma(l, l, 4.0 * eta * l);
        lam[k] = 1.0 / (0.5 * l + eta + 0.5 * sqrt(disc > 0.0 ? disc : 0.0));
    }
    __syncwarp();
    if (fraw >= n) return;
    for (int e = gl; e < DD; e += JG) {
        const int i0 = e % DIM, j0 = e / DIM;
        const int i = i0 > j0 ? i0 : j0, j = i0 > j0 ? j0 : i0;  // one arithmetic for (i, j) and (j, i)
        double v = 0.0, pij = 0.0, pji = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            v = fma(W3[i + k * DIM] * W3[j + k * DIM], lam[k], v);
            pij = fma(R[i + k * DIM], R[k + j * DIM], pij);  // P = R R
            pji = fma(R[j + k * DIM], R[k + i * DIM], pji);
        }
        fVdd[(size_t)f * DD + e] = (v - 0.5 * (pij + pji)) / eta;
    }
    for (int i = gl; i < DIM; i += JG) {
        double b = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) b = fma(R[i + k * DIM], r[1 + k], b);
        fVdmu[(size_t)f * DIM + i] = -b;
    }
    if (gl == 0) fcost[f] = e0;
}

template <int DIM, int M, int SD>
__global__ void __launch_bounds__(64) k_prox_linear(const LinearArgs a, double eta, double* __restrict__ fVdd) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.n) return;
    const size_t n = (size_t)a.n;
    const int s = a.start[f];
    const double* L = a.Lambda + f;
    const double* Ki = a.Kinv + f;
    double mu[DIM], r[M], kr[M];
#pragma unroll
    for (int k = 0; k < DIM; ++k) mu[k] = a.mu[(size_t)s * SD + k];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double v = -a.psi[(size_t)i * n + f];
#pragma unroll
        for (int k = 0; k < DIM; ++k) v = fma(L[(size_t)(i + k * M) * n], mu[k], v);
        r[i] = v;
    }
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) v = fma(Ki[(size_t)(i + k * M) * n], r[k], v);
        kr[i] = v;
    }
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) v = fma(L[(size_t)(i + k * M) * n], kr[i], v);
        a.fVdmu[(size_t)f * DIM + k] = -v;  // Vdmu = -b_k
    }
    Mat<DIM> Sk, Sig, P, Vdd;
    {
        const double* A = a.A + f;
        int e = 0;
#pragma unroll
        for (int i = 0; i < DIM; ++i)
#pragma unroll
            for (int j = i; j < DIM; ++j) {
                const double av = A[(size_t)e * n];
                Sk(i, j) = av;
                Sk(j, i) = av;
                ++e;
            }
    }
#pragma unroll
    for (int j = 0; j < DIM; ++j)
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            const int bi = i / SD, ii = i % SD, bj = j / SD, jj = j % SD;
            Sig(i, j) = (bi == bj)  ? a.covD[(size_t)(s + bi) * SD * SD + ii + jj * SD]
                        : (bi < bj) ? a.covO[(size_t)s * SD * SD + ii + jj * SD]
                                    : a.covO[(size_t)s * SD * SD + jj + ii * SD];
        }
    symmetrize<DIM>(Sig);
    LogDetAcc ld;
    spd_inverse<DIM>(P, Sig, ld);
    bw_jko<DIM>(Vdd, Sig, P, Sk, eta);
    mat_store<DIM>(fVdd + (size_t)f * DIM * DIM, Vdd);
}

// Prox candidate (ProxGVIGH::onestep_linesearch, proxgd/ProxGVI-GH-impl.h:24-43): mu' = mu + a dmu, Lambda' = Lambda + a dLambda
__global__ void k_candidate_add(size_t nmu, size_t nD, size_t nO, double alpha, const double* __restrict__ mu,
                                const double* __restrict__ dmu, const double* __restrict__ LD, const double* __restrict__ LO,
                                const double* __restrict__ dD, const double* __restrict__ dO, double* __restrict__ mu_c,
                                double* __restrict__ LD_c, double* __restrict__ LO_c) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nmu) mu_c[i] = mu[i] + alpha * dmu[i];
    if (i < nD) LD_c[i] = LD[i] + alpha * dD[i];
    if (i < nO) LO_c[i] = LO[i] + alpha * dO[i];
}

// ------------------------------------------------------------------------------------------
// K3: assembly (local2joint_* + the sums of NGDGH::compute_gradients, ngd/NGD-GH-impl.h:36-57).
// Every output element gathers, in a fixed order, the factor blocks that touch its state:
//   Vdmu[s]   = sum of fVdmu pieces;   VD[s] = KlinD[s] + sum of diag pieces;   VO[s] = KlinO[s] + off pieces
// ------------------------------------------------------------------------------------------
// One thread per (state, block element): consecutive threads read consecutive doubles of a factor's block, so the
// gather is coalesced.  Threads e < D of a state also gather Vdmu.
template <int D>
__global__ void __launch_bounds__(256) k_assemble(int S, const int* __restrict__ vptr, const int* __restrict__ voff,
                           const int* __restrict__ dptr, const int* __restrict__ doff, const int* __restrict__ dld,
                           const int* __restrict__ optr, const int* __restrict__ ooff, const int* __restrict__ old,
                           const double* __restrict__ fVdmu, const double* __restrict__ fVdd,
                           const double* __restrict__ KlinD, const double* __restrict__ KlinO,
                           double* __restrict__ Vdmu, double* __restrict__ VD, double* __restrict__ VO,
                           double* __restrict__ rhs) {
    constexpr int DD = D * D;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = (int)(gid / DD), e = (int)(gid % DD);
    if (s >= S) return;
    const int i = e % D, j = e / D;
    double m = KlinD[gid];
    for (int q = dptr[s]; q < dptr[s + 1]; ++q) m += fVdd[doff[q] + i + j * dld[q]];
    VD[gid] = m;
    if (s < S - 1) {
        double o = KlinO[gid];
        for (int q = optr[s]; q < optr[s + 1]; ++q) o += fVdd[ooff[q] + i + j * old[q]];
        VO[gid] = o;
    }
    if (e < D) {
        double v = 0.0;
        for (int q = vptr[s]; q < vptr[s + 1]; ++q) v += fVdmu[voff[q] + e];
        Vdmu[(size_t)s * D + e] = v;
        rhs[(size_t)s * D + e] = -v;
    }
}

// The same assembly over the ELL form of the adjacency ([k][state], -1 padded): one thread per (state, block column),
// one coalesced index load per contribution instead of the pointer chase, D elements per thread.  Same summation order
// as k_assemble (contributions in factor-id order on top of the constant block), so the results are bit-identical.
template <int D>
__global__ void __launch_bounds__(128) k_assemble_ell(int S, int nv, int nd, int no, const int* __restrict__ ev,
                                                      const int* __restrict__ ed, const int* __restrict__ edl,
                                                      const int* __restrict__ eo, const int* __restrict__ eol,
                                                      const double* __restrict__ fVdmu, const double* __restrict__ fVdd,
                                                      const double* __restrict__ KlinD, const double* __restrict__ KlinO,
                                                      double* __restrict__ Vdmu, double* __restrict__ VD, double* __restrict__ VO,
                                                      double* __restrict__ rhs) {
    constexpr int DD = D * D;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = (int)(gid / D), j = (int)(gid - (long long)s * D);
    if (s >= S) return;
    const size_t cb = (size_t)s * DD + (size_t)j * D;  // first element of block column j of state s
    double m[D], o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) m[i] = KlinD[cb + i];
    const bool has_off = (s < S - 1);
    if (has_off) {
#pragma unroll
        for (int i = 0; i < D; ++i) o[i] = KlinO[cb + i];
    }
    for (int k = 0; k < nd; ++k) {
        const int off = ed[(size_t)k * S + s];
        if (off >= 0) {
            const double* src = fVdd + off + (size_t)j * edl[(size_t)k * S + s];
#pragma unroll
            for (int i = 0; i < D; ++i) m[i] += src[i];
        }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) VD[cb + i] = m[i];
    if (has_off) {
        for (int k = 0; k < no; ++k) {
            const int off = eo[(size_t)k * S + s];
            if (off >= 0) {
                const double* src = fVdd + off + (size_t)j * eol[(size_t)k * S + s];
#pragma unroll
                for (int i = 0; i < D; ++i) o[i] += src[i];
            }
        }
#pragma unroll
        for (int i = 0; i < D; ++i) VO[cb + i] = o[i];
    }
    // thread j of the state gathers element j of Vdmu
    double v = 0.0;
    for (int k = 0; k < nv; ++k) {
        const int off = ev[(size_t)k * S + s];
        if (off >= 0) v += fVdmu[off + j];
    }
    Vdmu[(size_t)s * D + j] = v;
    rhs[(size_t)s * D + j] = -v;
}

// The same assembly with compile-time widths (NV <= 4 mean contributions, ND <= 1 diagonal and no off-diagonal block
// contribution per state: every chain with single-state nonlinear factors, e.g. the headline shape): all index loads are
// issued first, then all data loads as 16-byte vectors, so that a thread has its whole working set in flight at once
// instead of one dependent pair after the other.  Same operands in the same order: bit-identical to k_assemble_ell.
template <int D, int NV, int ND>
__global__ void __launch_bounds__(128) k_assemble_ell_fast(int S, const int* __restrict__ ev, const int* __restrict__ ed,
                                                           const int* __restrict__ edl, const double* __restrict__ fVdmu,
                                                           const double* __restrict__ fVdd, const double* __restrict__ KlinD,
                                                           const double* __restrict__ KlinO, double* __restrict__ Vdmu,
                                                           double* __restrict__ VD, double* __restrict__ VO,
                                                           double* __restrict__ rhs) {
    static_assert(D % 2 == 0, "vector loads need an even block dimension");
    constexpr int DD = D * D;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = (int)(gid / D), j = (int)(gid - (long long)s * D);
    if (s >= S) return;
    const size_t cb = (size_t)s * DD + (size_t)j * D;
    int iv[NV > 0 ? NV : 1], id[ND > 0 ? ND : 1], il[ND > 0 ? ND : 1];
#pragma unroll
    for (int k = 0; k < NV; ++k) iv[k] = __ldg(ev + (size_t)k * S + s);
#pragma unroll
    for (int k = 0; k < ND; ++k) {
        id[k] = __ldg(ed + (size_t)k * S + s);
        il[k] = __ldg(edl + (size_t)k * S + s);
    }
    // VO == nullptr: no factor contributes an off-diagonal block and the caller reads the constant KlinO in place of VO
    const bool has_off = (s < S - 1) && (VO != nullptr);
    double2 m[D / 2], o[D / 2];
#pragma unroll
    for (int i = 0; i < D / 2; ++i) {
        m[i] = __ldg(reinterpret_cast<const double2*>(KlinD + cb) + i);
        o[i] = has_off ? __ldg(reinterpret_cast<const double2*>(KlinO + cb) + i) : make_double2(0.0, 0.0);
    }
    double c[ND > 0 ? ND : 1][D];
#pragma unroll
    for (int k = 0; k < ND; ++k) {
        // a factor block may start at an odd double (factor outputs are packed): scalar loads, predicated
        const double* src = fVdd + (id[k] >= 0 ? (size_t)id[k] + (size_t)j * il[k] : 0);
#pragma unroll
        for (int i = 0; i < D; ++i) c[k][i] = id[k] >= 0 ? __ldg(src + i) : 0.0;
    }
    double vv[NV > 0 ? NV : 1];
#pragma unroll
    for (int k = 0; k < NV; ++k) vv[k] = iv[k] >= 0 ? __ldg(fVdmu + iv[k] + j) : 0.0;
#pragma unroll
    for (int k = 0; k < ND; ++k)
        if (id[k] >= 0) {
#pragma unroll
            for (int i = 0; i < D / 2; ++i) {
                m[i].x += c[k][2 * i];
                m[i].y += c[k][2 * i + 1];
            }
        }
#pragma unroll
    for (int i = 0; i < D / 2; ++i) reinterpret_cast<double2*>(VD + cb)[i] = m[i];
    if (has_off) {
#pragma unroll
        for (int i = 0; i < D / 2; ++i) reinterpret_cast<double2*>(VO + cb)[i] = o[i];
    }
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k)
        if (iv[k] >= 0) v += vv[k];
    Vdmu[(size_t)s * D + j] = v;
    rhs[(size_t)s * D + j] = -v;
}

// Line-search candidate (NGDGH::onestep_linesearch, ngd/NGD-GH-impl.h:129-148):
//   mu' = mu + a dmu,  Lambda' = Lambda + a (Vddmu - Lambda)
__global__ void k_candidate(size_t nmu, size_t nD, size_t nO, double alpha, const double* __restrict__ mu,
                            const double* __restrict__ dmu, const double* __restrict__ LD,
                            const double* __restrict__ LO, const double* __restrict__ VD,
                            const double* __restrict__ VO, double* __restrict__ mu_c, double* __restrict__ LD_c,
                            double* __restrict__ LO_c) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nmu) mu_c[i] = mu[i] + alpha * dmu[i];
    if (i < nD) LD_c[i] = LD[i] + alpha * (VD[i] - LD[i]);
    if (i < nO) LO_c[i] = LO[i] + alpha * (VO[i] - LO[i]);
}

// dprecision = Vddmu - Lambda (ngd/NGD-GH-impl.h:57), only materialised for gvib200_gradients()
__global__ void k_sub(size_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}

// Deterministic one-launch total: block b reduces a fixed contiguous slice of v into partial[b]; the block that
// finishes last (device counter) adds the partials in a fixed tree: out[0] = sum(v) + half * extra[0].
// fixed-tree sum of one double per thread over a 256-thread block (shuffles inside a warp, then across the 8 warps);
// the result is valid in thread 0
__device__ __forceinline__ double block_sum_256(double v, double* sh8) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh8[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 8) {
        t = sh8[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_down_sync(0x000000ffu, t, o);
    }
    return t;
}

// ------------------------------------------------------------------------------------------
// Multi-GPU exchange through peer memory (NVLink / NVSwitch): every rank owns a "mailbox" in its device memory that its
// peers map over CUDA IPC.  A rank PUSHES its record into every peer's mailbox with plain stores, publishes it with a
// release store of the pass's epoch to the peer's flag word, and waits for the P flags of its own mailbox -- no NCCL call,
// no host involvement, no extra launch: the exchange sits inside the single-CTA kernel that needs it.
//   records     [slot][depth][rank][MBOX_REC]   boundary records of the chain passes (slot = workspace slot of the pass)
//   rec flags   [slot][depth][rank]             epoch of the record held in that cell
//   cost        [depth][rank][4]                (cost, flag0, flag1, 0) of a cost evaluation
//   cost flags  [depth][rank]
// Epochs only grow; a cell is reused every MBOX_DEPTH epochs.  Between two passes on the same slot lies at least one
// cost exchange in which every rank waits for every other, so a writer is never more than one epoch ahead of a reader.
// ------------------------------------------------------------------------------------------
constexpr int MBOX_RANKS = 16;
constexpr int MBOX_DEPTH = 4;
constexpr int MBOX_REC = 256;  // doubles per boundary record cell (5 d^2 + 4 d = 204 at d = 6)
constexpr size_t MBOX_OFF_REC = 0;
constexpr size_t MBOX_OFF_RECFLAG = MBOX_OFF_REC + (size_t)2 * MBOX_DEPTH * MBOX_RANKS * MBOX_REC;
constexpr size_t MBOX_OFF_COST = MBOX_OFF_RECFLAG + (size_t)2 * MBOX_DEPTH * MBOX_RANKS;
constexpr size_t MBOX_OFF_COSTFLAG = MBOX_OFF_COST + (size_t)MBOX_DEPTH * MBOX_RANKS * 4;
constexpr size_t MBOX_DOUBLES = MBOX_OFF_COSTFLAG + (size_t)MBOX_DEPTH * MBOX_RANKS;

struct MboxPeers {
    double* p[MBOX_RANKS];  // every rank's mailbox as mapped on THIS device (p[rank] is the local one)
    int world, rank;
};

__device__ __forceinline__ void mbox_publish(double* cell, unsigned long long epoch) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(cell), "l"(epoch) : "memory");
}
// spin until the flag word carries `epoch`; gives up after ~2 s so that a dead peer cannot hang the device
__device__ __forceinline__ bool mbox_wait(const double* cell, unsigned long long epoch) {
    const long long t0 = clock64();
    while (true) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(cell) : "memory");
        if (v >= epoch) return true;
        if (clock64() - t0 > 4000000000ll) return false;
        __nanosleep(64);
    }
}

// block_sum_256 leaves its result in thread 0: hand it to every thread (one more barrier)
__device__ __forceinline__ double tot_bcast(double v, double* sh8) {
    __syncthreads();
    if (threadIdx.x == 0) sh8[0] = v;
    __syncthreads();
    return sh8[0];
}

__global__ void __launch_bounds__(256) k_total(size_t n, const double* __restrict__ v, double* __restrict__ partial,
                                               unsigned* __restrict__ counter, const double* __restrict__ extra, double half,
                                               double* __restrict__ out, int* __restrict__ dflag, double* zc, int which,
                                               double* __restrict__ red, const MboxPeers peers, unsigned long long epoch) {
    __shared__ double sh8[8];
    __shared__ bool is_last;
    // block b owns the fixed slice [b * per, (b + 1) * per); a thread takes (at most four) elements at stride 256, all
    // loads in flight at once
    const size_t per = (n + gridDim.x - 1) / gridDim.x;
    const size_t lo = (size_t)blockIdx.x * per;
    const size_t hi = (lo + per < n) ? lo + per : n;
    double s = 0.0;
    for (size_t i = lo + threadIdx.x; i < hi; i += 1024) {
        const double a0 = v[i];
        const double a1 = (i + 256 < hi) ? v[i + 256] : 0.0;
        const double a2 = (i + 512 < hi) ? v[i + 512] : 0.0;
        const double a3 = (i + 768 < hi) ? v[i + 768] : 0.0;
        s += (a0 + a1) + (a2 + a3);
    }
    const double bs = block_sum_256(s, sh8);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = bs;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double t = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 256) t += __ldcg(partial + i);
    const double tot = block_sum_256(t, sh8);
    if (peers.world > 1) {
        // multi-GPU: this rank's (cost, flags) go to every peer's mailbox, the P records come back, every rank forms the
        // same rank-ordered sum; the host gets the result through mapped memory like on one GPU
        __shared__ double rec[MBOX_RANKS][4];
        __shared__ int bad;
        const int P = peers.world, dq = (int)(epoch % MBOX_DEPTH);
        if (threadIdx.x == 0) bad = 0;
        const double total = tot_bcast(tot, sh8) + (extra ? half * extra[0] : 0.0);  // (barriers inside: all threads)
        if ((int)threadIdx.x < P) {
            double* dst = peers.p[threadIdx.x] + MBOX_OFF_COST + ((size_t)dq * MBOX_RANKS + peers.rank) * 4;
            dst[0] = total;
            dst[1] = (double)dflag[0];
            dst[2] = (double)dflag[1];
            dst[3] = 0.0;
            __threadfence_system();
            mbox_publish(peers.p[threadIdx.x] + MBOX_OFF_COSTFLAG + (size_t)dq * MBOX_RANKS + peers.rank, epoch);
            const double* mine = peers.p[peers.rank];
            if (!mbox_wait(mine + MBOX_OFF_COSTFLAG + (size_t)dq * MBOX_RANKS + threadIdx.x, epoch)) bad = 1;
            const double* src = mine + MBOX_OFF_COST + ((size_t)dq * MBOX_RANKS + threadIdx.x) * 4;
            for (int e = 0; e < 4; ++e) rec[threadIdx.x][e] = __ldcg(src + e);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double c = 0.0, f0 = bad ? 1.0 : 0.0, f1 = 0.0;
            for (int r = 0; r < P; ++r) {
                c += rec[r][0];
                f0 += rec[r][1];
                f1 += rec[r][2];
            }
            out[0] = c;
            dflag[0] = f0 > 0.0 ? 1 : 0;
            dflag[1] = f1 > 0.0 ? 1 : 0;
            *counter = 0u;
            if (zc != nullptr) {
                zc[which] = c;
                zc[2] = f0 > 0.0 ? 1.0 : 0.0;
                zc[3] = f1 > 0.0 ? 1.0 : 0.0;
                __threadfence_system();
            }
        }
        return;
    }
    if (threadIdx.x == 0) {
        const double total = tot + (extra ? half * extra[0] : 0.0);
        out[0] = total;
        *counter = 0u;  // ready for the next launch
        if (zc != nullptr) {  // mapped host memory: the host reads the cost and the not-SPD flags without a copy
            zc[which] = total;
            zc[2] = (double)dflag[0];
            zc[3] = (double)dflag[1];
            __threadfence_system();
        }
        if (red != nullptr) {  // multi-GPU: staging of the cost / flag all-reduce (k_red_pack folded in)
            red[0] = total;
            red[1] = (double)dflag[0];
            red[2] = (double)dflag[1];
            red[3] = 0.0;
        }
    }
}

// Deterministic single-block sum: out[0] = sum(v[0..n)) (+ half * extra[0] if extra != null)
__global__ void k_sum(size_t n, const double* __restrict__ v, const double* __restrict__ extra, double half,
                      double* __restrict__ out) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0] + (extra ? half * extra[0] : 0.0);
}

// ------------------------------------------------------------------------------------------
// K4: tile-wise block cyclic reduction, see bt_cr.h.  Three launches per solve:
//   k_cr_tile_forward   one CTA per tile: load the tile into shared memory, eliminate its interior level by level,
//                       stream the elimination records to HBM, hand the separator Schur complements to the top
//   k_cr_top            ONE CTA: cyclic reduction of the separator chain (records stay in shared memory), 2-node
//                       solve, expansion back to every separator (or, K == 0, the whole chain at once)
//   k_cr_tile_backward  one CTA per tile: back substitution and / or Takahashi selected inverse, level by level
// ------------------------------------------------------------------------------------------
constexpr int CR_THREADS = 512;

// development aid (GVIB200_CHAIN_CLOCKS=1): CTA 0 of the chain kernels stamps clock64() at its phase boundaries
// (compile with -DGVIB200_CHAIN_CLOCKS; the production build has no stamps)
__device__ long long* g_cr_clk = nullptr;
__device__ __forceinline__ void cr_stamp(int slot) {
#ifdef GVIB200_CHAIN_CLOCKS
    if (blockIdx.x == 0 && threadIdx.x == 0 && g_cr_clk != nullptr) g_cr_clk[slot] = clock64();
#endif
}

// deterministic block sum of one double per thread (fixed tree), result valid in thread 0
__device__ __forceinline__ double cr_block_sum(double v, double* red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    return red[0];
}

// D workers per node (bt_cr.h): thread tid works on node tid / D of the round as worker tid % D
template <int D, bool RHS, bool BATCH = false>
__device__ __forceinline__ bool cr_forward_levels(const CrView<D>& v, const CrRec<D>& rec, size_t rec_base, const CrGeom& gm,
                                                  LogDetAcc& ld, int clk0 = 0) {
    bool ok = true;
    const int per_round = blockDim.x / D;
    const int tn = threadIdx.x / D, c = threadIdx.x - tn * D;
    for (int l = 0; l < gm.levels; ++l) {
        const int cnt = cr_count(gm.T, l);
        for (int base = 0; base < cnt; base += per_round) {
            const int t = base + tn;
            const bool on = (tn < per_round) && (t < cnt);
            CrElim<D> el;
            if (on) ok = cr_fwd_A<D, RHS, BATCH>(v, rec, rec_base, gm, l, t, c, el, ld) && ok;
            __syncthreads();
            if (on) cr_fwd_B<D, RHS>(v, el, c);
            __syncthreads();
        }
        cr_stamp(clk0 + 2 + l);
    }
    return ok;
}

template <int D, bool RHS, bool SELINV>
__device__ __forceinline__ void cr_backward_levels(const CrView<D>& v, const CrRec<D>& rec, size_t rec_base,
                                                   const CrGeom& gm, int clk0 = 0) {
    static_assert(!(RHS && SELINV), "a pass is either a solve or a selected inverse (the forward phase of a solve keeps no pivot inverse)");
    const int per_round = blockDim.x / D;
    const int tn = threadIdx.x / D, c = threadIdx.x - tn * D;
    for (int l = gm.levels - 1; l >= 0; --l) {
        const int cnt = cr_count(gm.T, l);
        for (int base = 0; base < cnt; base += per_round) {
            const int t = base + tn;
            const bool on = (tn < per_round) && (t < cnt);
            if (SELINV) {
                CrSel<D> o;
                if (on) cr_bwd_selinv_compute<D>(v, rec, rec_base, gm, l, t, c, o);
                __syncthreads();  // the coupling of the left neighbour is overwritten
                if (on) cr_bwd_selinv_store<D>(v, o, c);
            }
            if (RHS && on) cr_bwd_solve<D>(v, rec, rec_base, gm, l, t, c);
        }
        __syncthreads();
        cr_stamp(clk0 + 2 + (gm.levels - 1 - l));
    }
}

// The three kernel bodies as device functions over caller-provided shared memory (dynamic `smem`, geometry `gm`,
// reduction scratch `red`), so that the multi-GPU pass can chain several of them inside one single-CTA launch.
template <int D, bool RHS, bool BATCH = false>
__device__ __forceinline__ void cr_dev_tile_forward(const CrArgs<D>& a, int tile, double* smem, CrGeom& gm, double* red) {
    const int n0 = tile * a.T;
    const int Tk = min(a.T, a.n - 1 - n0);
    if (threadIdx.x == 0) cr_make_geom(gm, Tk);
    __syncthreads();
    CrView<D> v = cr_make_view<D>(smem, a.T + 1);
    if constexpr (BATCH) {
        v.ldnode = a.ldnode;
        v.ld_base = n0;
        v.ld_mul = 1;
        v.ld_nmax = (long long)a.n - 1;
        v.ld_stride = a.ld_stride;
        v.ld_max = a.ld_max;
    }
    const int clk0 = RHS ? 0 : 32;
    cr_stamp(clk0);
    cr_tile_load<D, RHS, BATCH>(a, v, gm, n0, threadIdx.x, blockDim.x);
    __syncthreads();
    cr_stamp(clk0 + 1);
    LogDetAcc ld;
    const bool ok = cr_forward_levels<D, RHS, BATCH>(v, a.rec, (size_t)tile * (a.T - 1), gm, ld, clk0);
    cr_tile_store_reduced<D, RHS, BATCH>(a, v, gm, tile, n0, threadIdx.x, blockDim.x);
    const double s = cr_block_sum(ld.value(), red);
    if (threadIdx.x == 0) a.ld[tile] = s;
    if (!ok) *a.notspd = 1;
    cr_stamp(clk0 + 20);
}

template <int D, bool RHS, bool SELINV, bool BATCH = false>
__device__ __forceinline__ void cr_dev_top(const CrArgs<D>& a, double* smem, CrGeom& gm, double* red) {
    const int nt = (a.K == 0) ? a.n : a.K + 1;  // nodes of the top chain
    if (threadIdx.x == 0) cr_make_geom(gm, nt - 1);
    __syncthreads();
    CrView<D> v = cr_make_view<D>(smem, nt);
    if constexpr (BATCH) {
        v.ldnode = a.ldnode;
        v.ld_base = 0;
        v.ld_mul = (a.K == 0) ? 1 : a.T;  // the top chain of a tiled level consists of its separators k T (the last one is n - 1)
        v.ld_nmax = (long long)a.n - 1;
        v.ld_stride = a.ld_stride;
        v.ld_max = a.ld_max;
    }
    // elimination records of the top stay in shared memory
    const size_t nrec = nt > 2 ? (size_t)(nt - 2) : 0;
    CrRec<D> rec;
    rec.G = v.g + (size_t)D * v.NS;
    rec.H = rec.G + cr_rec_capacity(nrec, D * D);
    rec.Dinv = rec.H + cr_rec_capacity(nrec, D * D);
    rec.y = rec.Dinv + cr_rec_capacity(nrec, D * D);
    cr_top_load<D, RHS, BATCH>(a, v, gm, threadIdx.x, blockDim.x);
    __syncthreads();
    LogDetAcc ld;
    bool ok = cr_forward_levels<D, RHS, BATCH>(v, rec, 0, gm, ld);
    if (threadIdx.x == 0) ok = cr_top2<D, RHS, SELINV, BATCH>(v, gm.T, ld) && ok;
    __syncthreads();
    cr_backward_levels<D, RHS, SELINV>(v, rec, 0, gm);
    if (a.K == 0)
        cr_store_results<D, RHS, SELINV>(v, gm, nt, nt - 1, a.x, a.cD, a.cO, 0, threadIdx.x, blockDim.x, a.xbase, a.xalpha, a.xout);
    else cr_store_results<D, RHS, SELINV>(v, gm, nt, nt - 1, a.tx, a.tD, a.tO, 0, threadIdx.x, blockDim.x);
    const double s = cr_block_sum(ld.value(), red);
    if (threadIdx.x == 0) a.ld[a.K] = s;
    if (a.ldout != nullptr) {  // total log determinant: the tiles' partial sums are complete (previous launch)
        __syncthreads();
        double t = 0.0;
        for (int i = threadIdx.x; i < a.K; i += blockDim.x) t += a.ld[i];
        const double tot = cr_block_sum(t, red);
        if (threadIdx.x == 0) a.ldout[0] = tot + s;
    }
    if (!ok) *a.notspd = 1;
}

template <int D, bool RHS, bool SELINV>
__device__ __forceinline__ void cr_dev_tile_backward(const CrArgs<D>& a, int tile, double* smem, CrGeom& gm) {
    const int n0 = tile * a.T;
    const int Tk = min(a.T, a.n - 1 - n0);
    if (threadIdx.x == 0) cr_make_geom(gm, Tk);
    __syncthreads();
    const CrView<D> v = cr_make_view<D>(smem, a.T + 1);
    const int clk0 = RHS ? 64 : 96;
    cr_stamp(clk0);
    cr_tile_seed<D, RHS, SELINV>(a, v, tile, threadIdx.x, blockDim.x);
    __syncthreads();
    cr_stamp(clk0 + 1);
    cr_backward_levels<D, RHS, SELINV>(v, a.rec, (size_t)tile * (a.T - 1), gm, clk0);
    const bool last = (tile == a.K - 1);
    cr_store_results<D, RHS, SELINV>(v, gm, Tk + (last ? 1 : 0), Tk, a.x, a.cD, a.cO, (size_t)n0, threadIdx.x, blockDim.x,
                                     a.xbase, a.xalpha, a.xout);
    __syncthreads();
    cr_stamp(clk0 + 20);
}

template <int D, bool RHS, bool BATCH = false>
__global__ void __launch_bounds__(CR_THREADS, 1) k_cr_tile_forward(const CrArgs<D> a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ CrGeom gm;
    __shared__ double red[CR_THREADS];
    cr_dev_tile_forward<D, RHS, BATCH>(a, blockIdx.x, smem, gm, red);
}

template <int D, bool RHS, bool SELINV, bool BATCH = false>
__global__ void __launch_bounds__(CR_THREADS, 1) k_cr_top(const CrArgs<D> a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ CrGeom gm;
    __shared__ double red[CR_THREADS];
    cr_dev_top<D, RHS, SELINV, BATCH>(a, smem, gm, red);
}

template <int D, bool RHS, bool SELINV>
__global__ void __launch_bounds__(CR_THREADS, 1) k_cr_tile_backward(const CrArgs<D> a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ CrGeom gm;
    cr_dev_tile_backward<D, RHS, SELINV>(a, blockIdx.x, smem, gm);
}

// Multi-GPU pass, the two single-CTA stages around the boundary all-gather, each ONE launch:
//   k_cr_mid_forward  separator system of this rank's tiles summed (cr_sum_level) -> the "mid" tile over it eliminated ->
//                     this rank's boundary record packed for the all-gather
//   k_cr_dist_top     all boundary records -> chain of rank boundaries (cr_build_global) solved redundantly -> seeds of
//                     the mid tile -> the mid tile walked back down (its results seed the real tiles) -> this rank's
//                     share of log det
// A stage only consumes what the same CTA wrote before the preceding __syncthreads().
template <int D, bool RHS>
__global__ void __launch_bounds__(CR_THREADS, 1) k_cr_mid_forward(const CrArgs<D> a, const CrArgs<D> mid, double* D1, double* O1,
                                                                  double* g1, double* send) {
    extern __shared__ __align__(16) double smem[];
    __shared__ CrGeom gm;
    __shared__ double red[CR_THREADS];
    cr_sum_level<D, RHS>(a, D1, O1, g1, threadIdx.x, blockDim.x);
    __syncthreads();
    cr_dev_tile_forward<D, RHS>(mid, 0, smem, gm, red);
    __syncthreads();
    cr_pack_boundary<D, RHS>(mid, send, threadIdx.x, blockDim.x);
}

template <int D, bool RHS, bool SELINV>
__global__ void __launch_bounds__(CR_THREADS, 1) k_cr_dist_top(int P, int rank, const double* recs, double* Dt, double* Ot, double* gt,
                                                               const CrArgs<D> top, const CrArgs<D> mid, const double* tile_ld,
                                                               int n_tile_ld, double* ldout) {
    extern __shared__ __align__(16) double smem[];
    __shared__ CrGeom gm;
    __shared__ double red[CR_THREADS];
    cr_build_global<D>(P, recs, Dt, Ot, gt, threadIdx.x, blockDim.x);
    __syncthreads();
    cr_dev_top<D, RHS, SELINV>(top, smem, gm, red);
    __syncthreads();
    cr_seed_mid<D, RHS, SELINV>(mid, rank, top.x, top.cD, top.cO, threadIdx.x, blockDim.x);
    __syncthreads();
    cr_dev_tile_backward<D, RHS, SELINV>(mid, 0, smem, gm);
    if (ldout != nullptr) {
        // this rank's share of log det: its tiles + its mid tile; the chain of rank boundaries is counted by rank 0 only
        __syncthreads();
        double t = 0.0;
        for (int i = threadIdx.x; i < n_tile_ld; i += blockDim.x) t += tile_ld[i];
        const double tot = cr_block_sum(t, red);
        if (threadIdx.x == 0) ldout[0] = tot + mid.ld[0] + (rank == 0 ? top.ld[0] : 0.0);
    }
}

// Multi-GPU pass, the single-CTA stage between the two tile launches, ONE launch that lives in shared memory like k_cr_top:
//   separator system of this rank's tiles gathered (cr_top_load) and reduced by cyclic reduction down to its two end
//   nodes, records in shared memory -> what is left on the end nodes IS this rank's boundary record: pushed into every
//   peer's mailbox over NVLink -> all records in -> chain of rank boundaries (P + 1 nodes) solved redundantly in a second,
//   small shared-memory region -> its results seed the two end nodes -> back substitution / Takahashi recursion up the
//   separator chain -> separator results for the tiles' backward launch, this rank's share of log det.
template <int D, bool RHS, bool SELINV>
__global__ void __launch_bounds__(CR_THREADS, 1)
    k_cr_dist_mid(const CrArgs<D> a, const CrArgs<D> top, double* send, double* recv, double* Dt, double* Ot, double* gt,
                  const MboxPeers peers, int slot, unsigned long long epoch, double* ldout) {
    extern __shared__ __align__(16) double smem[];
    __shared__ CrGeom gm, gm2;
    __shared__ double red[CR_THREADS];
    __shared__ int timed_out;
    constexpr int DD = D * D;
    constexpr int NB = cr_boundary_doubles<D>();
    const int P = peers.world, rank = peers.rank, dq = (int)(epoch % MBOX_DEPTH);
    const int nt = a.K + 1;  // nodes of the separator chain
    if (threadIdx.x == 0) {
        cr_make_geom(gm, nt - 1);
        timed_out = 0;
    }
    __syncthreads();
    const CrView<D> v = cr_make_view<D>(smem, nt);
    const size_t nrec = nt > 2 ? (size_t)(nt - 2) : 0;
    CrRec<D> rec;
    rec.G = v.g + (size_t)D * v.NS;
    rec.H = rec.G + cr_rec_capacity(nrec, DD);
    rec.Dinv = rec.H + cr_rec_capacity(nrec, DD);
    rec.y = rec.Dinv + cr_rec_capacity(nrec, DD);
    // everything about the chain of rank boundaries stays in shared memory too: the received records, the chain itself,
    // its results, and the working arrays of its solve
    double* srecv = rec.y + cr_rec_capacity(nrec, D);
    double* ssend = srecv + (((size_t)P * NB + 1) & ~size_t(1));
    double* sDt = ssend + ((NB + 1) & ~1);
    double* sOt = sDt + (size_t)(P + 1) * DD;
    double* sgt = sOt + (size_t)(P + 1) * DD;
    double* sxt = sgt + (((size_t)(P + 1) * D + 1) & ~size_t(1));
    double* scD = sxt + (((size_t)(P + 1) * D + 1) & ~size_t(1));
    double* scO = scD + (size_t)(P + 1) * DD;
    double* sld = scO + (size_t)(P + 1) * DD;
    double* smem_top = sld + 2;
    (void)send;
    (void)recv;
    (void)Dt;
    (void)Ot;
    (void)gt;
    send = ssend;
    recv = srecv;
    Dt = sDt;
    Ot = sOt;
    gt = sgt;
    CrArgs<D> tp = top;
    tp.Dg = sDt;
    tp.Og = sOt;
    tp.g = sgt;
    tp.x = sxt;
    tp.cD = scD;
    tp.cO = scO;
    tp.ld = sld;
    cr_top_load<D, RHS>(a, v, gm, threadIdx.x, blockDim.x);
    __syncthreads();
    LogDetAcc ld;
    bool ok = cr_forward_levels<D, RHS>(v, rec, 0, gm, ld);
    // boundary record [Dfirst | Dlast | CL | CR | O | gfirst | glast | gl | gr]: the reduced end nodes carry everything
    for (int e = threadIdx.x; e < DD; e += blockDim.x) {
        send[e] = v.Dn[(size_t)e * v.NS + 0];
        send[DD + e] = v.Dn[(size_t)e * v.NS + 1];
        send[2 * DD + e] = 0.0;
        send[3 * DD + e] = 0.0;
        send[4 * DD + e] = v.P[(size_t)e * v.NS + 0];
    }
    for (int e = threadIdx.x; e < D; e += blockDim.x) {
        double* gv = send + 5 * DD;
        gv[e] = RHS ? v.g[(size_t)e * v.NS + 0] : 0.0;
        gv[D + e] = RHS ? v.g[(size_t)e * v.NS + 1] : 0.0;
        gv[2 * D + e] = 0.0;
        gv[3 * D + e] = 0.0;
    }
    __syncthreads();
    // push the record into every rank's mailbox (the own one included), then publish it
    const size_t cell = ((size_t)(slot * MBOX_DEPTH + dq) * MBOX_RANKS + rank);
    for (int idx = threadIdx.x; idx < P * NB; idx += blockDim.x) {
        const int r = idx / NB, e = idx - r * NB;
        peers.p[r][MBOX_OFF_REC + cell * MBOX_REC + e] = send[e];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < P) {
        mbox_publish(peers.p[threadIdx.x] + MBOX_OFF_RECFLAG + cell, epoch);
        const double* mine = peers.p[rank];
        if (!mbox_wait(mine + MBOX_OFF_RECFLAG + (size_t)(slot * MBOX_DEPTH + dq) * MBOX_RANKS + threadIdx.x, epoch)) timed_out = 1;
    }
    __syncthreads();
    {
        const double* mine = peers.p[rank] + MBOX_OFF_REC + (size_t)(slot * MBOX_DEPTH + dq) * MBOX_RANKS * MBOX_REC;
        for (int idx = threadIdx.x; idx < P * NB; idx += blockDim.x) {
            const int r = idx / NB, e = idx - r * NB;
            recv[idx] = __ldcg(mine + (size_t)r * MBOX_REC + e);  // written by the peers: not through L1
        }
    }
    __syncthreads();
    if (timed_out) {
        if (threadIdx.x == 0) *a.notspd = 1;  // surfaces as an error on the host instead of a hang
        return;
    }
    cr_build_global<D>(P, recv, Dt, Ot, gt, threadIdx.x, blockDim.x);
    __syncthreads();
    cr_dev_top<D, RHS, SELINV>(tp, smem_top, gm2, red);  // results: tp.x / tp.cD / tp.cO, tp.ld[0]
    __syncthreads();
    // seed the two end nodes of the separator chain with the results on this rank's boundaries
    if (SELINV)
        for (int e = threadIdx.x; e < DD; e += blockDim.x) {
            v.Dn[(size_t)e * v.NS + 0] = tp.cD[(size_t)rank * DD + e];
            v.Dn[(size_t)e * v.NS + 1] = tp.cD[(size_t)(rank + 1) * DD + e];
            v.P[(size_t)e * v.NS + 0] = tp.cO[(size_t)rank * DD + e];
        }
    if (RHS)
        for (int e = threadIdx.x; e < D; e += blockDim.x) {
            v.g[(size_t)e * v.NS + 0] = tp.x[(size_t)rank * D + e];
            v.g[(size_t)e * v.NS + 1] = tp.x[(size_t)(rank + 1) * D + e];
        }
    __syncthreads();
    cr_backward_levels<D, RHS, SELINV>(v, rec, 0, gm);
    cr_store_results<D, RHS, SELINV>(v, gm, nt, nt - 1, a.tx, a.tD, a.tO, 0, threadIdx.x, blockDim.x);
    const double s = cr_block_sum(ld.value(), red);
    if (ldout != nullptr) {
        // this rank's share of log det: its tiles + its separator chain; the chain of rank boundaries is counted by rank 0
        __syncthreads();
        double t = 0.0;
        for (int i = threadIdx.x; i < a.K; i += blockDim.x) t += a.ld[i];
        const double tot = cr_block_sum(t, red);
        if (threadIdx.x == 0) ldout[0] = tot + s + (rank == 0 ? tp.ld[0] : 0.0);
    }
    if (!ok) *a.notspd = 1;
}

// multi-GPU glue kernels (single small CTAs; the arithmetic is in bt_cr.h)
template <int D, bool RHS>
__global__ void k_cr_sum_level(const CrArgs<D> a, double* D1, double* O1, double* g1) {
    cr_sum_level<D, RHS>(a, D1, O1, g1, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
// out[0] = sum(v[0..n)) + a[0] + (b ? b[0] : 0), fixed order
__global__ void k_sum3(size_t n, const double* __restrict__ v, const double* __restrict__ a, const double* __restrict__ b,
                       double* __restrict__ out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0] + a[0] + (b ? b[0] : 0.0);
}
// staging of the cost / flag exchange: buf = {cost, flag0, flag1, 0}; after the all-gather cost and flags are written back
// ------------------------------------------------------------------------------------------
// Batches of independent problems with a line search PER PROBLEM (gvib200_batch_iterate): the problems are concatenated
// block-diagonally into one chain, problem q owns the states [soff[q], soff[q+1]).
// ------------------------------------------------------------------------------------------
// per-state step size from the per-problem one
__global__ void k_batch_alpha(int S, const int* __restrict__ sprob, const double* __restrict__ step_p, double* __restrict__ alpha_node) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) alpha_node[i] = step_p[sprob[i]];
}
// candidate mean mu' = mu + alpha(state) dmu
__global__ void k_batch_candidate_mu(size_t n, int d, const double* __restrict__ alpha_node, const double* __restrict__ mu,
                                     const double* __restrict__ dmu, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = mu[i] + alpha_node[i / d] * dmu[i];
}
// One warp per problem: sum of the problem's factor costs (group by group, the factors of a problem are a contiguous range
// [seg[g][q], seg[g][q+1]) of every group) + half the sum of its nodes' log pivots.  Lane l takes the elements l, l + 32, ...
// of each range relative to the range's start and the lanes meet in a fixed xor tree, so the value depends on the problem
// alone -- not on where in the batch it sits.  fcost == nullptr: the log-pivot sum only (not finite = some pivot was not
// positive: that problem's matrix is not SPD).
__global__ void k_problem_costs(int P, int G, const int* __restrict__ seg, const int* __restrict__ gfirst,
                                const double* __restrict__ fcost, const int* __restrict__ soff, const double* __restrict__ ldnode,
                                double half, double* __restrict__ out) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= P) return;
    double s = 0.0;
    if (fcost != nullptr)
        for (int g = 0; g < G; ++g) {
            const int lo = seg[(size_t)g * (P + 1) + q], hi = seg[(size_t)g * (P + 1) + q + 1];
            const double* fc = fcost + gfirst[g];
            for (int i = lo + lane; i < hi; i += 32) s += fc[i];
        }
    double l = 0.0;
    if (ldnode != nullptr)
        for (int i = soff[q] + lane; i < soff[q + 1]; i += 32) l += ldnode[i];
    s = fma(half, l, s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[q] = s;
}

// gvib200_set_state_async: the not-SPD flags of its selected inverse go to mapped host memory (no copy-engine transfer: a
// small device-to-host copy would queue behind another handle's bulk download on the shared copy engine)
__global__ void k_flags_to_host(const int* flags, double* zc_slot) {
    zc_slot[0] = (double)(flags[0] | flags[1]);
    __threadfence_system();
}
__global__ void k_red_pack(const double* cost, const int* flags, double* buf) {
    buf[0] = cost[0];
    buf[1] = (double)flags[0];
    buf[2] = (double)flags[1];
    buf[3] = 0.0;
}
__global__ void k_red_unpack(int world, const double* all, double* cost, int* flags, double* zc, int which) {
    // all: [world][4] = every rank's (cost, flag0, flag1, 0); summed in rank order, identically on every rank
    double c = 0.0, f0 = 0.0, f1 = 0.0;
    for (int r = 0; r < world; ++r) {
        c += all[4 * r];
        f0 += all[4 * r + 1];
        f1 += all[4 * r + 2];
    }
    cost[0] = c;
    flags[0] = f0 > 0.0 ? 1 : 0;
    flags[1] = f1 > 0.0 ? 1 : 0;
    if (zc != nullptr) {  // mapped host memory, as in k_total
        zc[which] = c;
        zc[2] = f0 > 0.0 ? 1.0 : 0.0;
        zc[3] = f1 > 0.0 ? 1.0 : 0.0;
        __threadfence_system();
    }
}

// cell records for CostPlanarHinge from the column-major field (layout: cost_functors.cuh)
__global__ void k_build_sdf_records(int rows, int cols, double thr, const double* __restrict__ data, double4* __restrict__ rec) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int r = idx % rows, c = idx / rows;
    const int hr = min(r + 1, rows - 1), hc = min(c + 1, cols - 1);
    const double v00 = data[r + (size_t)c * rows], v10 = data[hr + (size_t)c * rows];
    const double v01 = data[r + (size_t)hc * rows], v11 = data[hr + (size_t)hc * rows];
    double4 v;
    v.x = thr - v00;
    v.y = -(v10 - v00);
    v.z = -(v01 - v00);
    v.w = -((v11 - v01) - (v10 - v00));
    rec[idx] = v;
}

// free-space bound for CostPlanarHinge::all_zero: per 4 x 4 cell block the max of thr - (min corner distance of a cell)
__global__ void k_build_sdf_coarse(int rows, int cols, double thr, const double* __restrict__ data, double* __restrict__ coarse) {
    const int crows = (rows + 3) / 4, ccols = (cols + 3) / 4;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= crows * ccols) return;
    const int R = idx % crows, C = idx / crows;
    double lo = 1e300;  // min distance over the corners of the block's cells (upper indices clamped like the records)
    for (int c = 4 * C; c <= min(4 * C + 4, cols - 1); ++c)
        for (int r = 4 * R; r <= min(4 * R + 4, rows - 1); ++r) lo = fmin(lo, data[r + (size_t)c * rows]);
    coarse[idx] = thr - lo;
}

// FP64 FMA micro-benchmark: the FP64 roofline denominator (not in MEASURED_PEAKS.json)
__global__ void k_fp64_peak(int iters, double* out) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}

}  // namespace gvib200
