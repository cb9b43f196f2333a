// Block-tridiagonal SPD engine, second generation: tile-wise block cyclic reduction.
//
// Replaces, for the joint precision pattern fixed at gvibase/GVI-GH.h:214-230:
//   GVIGH::inverse_GBP + calculate_factor_message   gvibase/GVI-GH-GBP-impl.h:245-342
//   EigenWrapper::inv_sparse (Takahashi on LDLT)     helpers/EigenWrapper.h:336-381
//   SparseLDLT(Precision).vectorD().log().sum()      gvibase/GVI-GH-GBP-impl.h:234-238
//   ConjugateGradient(Vddmu).solve(-Vdmu)            ngd/NGD-GH-impl.h:59-60 (direct solve instead)
//
// The chain of n nodes is cut into K tiles of T links; tile k owns the nodes n0 = kT .. n0 + Tk (local 0 .. Tk), its two
// end nodes are separators shared with the neighbouring tiles.  One CTA holds a tile in shared memory and eliminates
// its Tk - 1 interior nodes by block cyclic reduction: at level l every node j = (2t+1) 2^l < Tk is eliminated against
// its alive neighbours i = j - 2^l and k = min(j + 2^l, Tk) -- a nested-dissection Cholesky, log2(T) dependent steps
// instead of T.  What is left (Schur complements on the separators and their coupling) is a block-tridiagonal system of
// K + 1 nodes which ONE CTA reduces the same way (the "top"), solves, and expands again; a last launch walks every tile
// back down (solution and / or the Takahashi selected inverse: diagonal and first off-diagonal blocks of the inverse).
// Three launches per solve, every global access coalesced.
//
// Storage inside a tile: node j lives in slot(j) -- 0 and Tk in slots 0 and 1, the nodes eliminated at level l
// contiguously from off[l] -- so that the threads of one level touch consecutive slots; element e of slot s sits at
// [e * NS + s] (structure of arrays: conflict-free shared-memory accesses).  The elimination record of node j (G, H,
// Dinv, y) goes to global memory at record index base + slot(j) - 2 in a 32-record interleaved layout, written and
// read back by the same lane pattern.
//
// Everything here is __host__ __device__ and free of CUDA intrinsics: tests/cpp/bt_host_emu.cpp runs the very same
// arithmetic on the CPU against the oracle.
#pragma once
#include "smallmat.h"

namespace gvib200 {

constexpr int CR_MAX_LEVELS = 16;

struct CrGeom {
    int T;                       // local nodes 0 .. T
    int levels;                  // cyclic-reduction levels until only {0, T} are alive
    int off[CR_MAX_LEVELS + 1];  // first slot (minus 2) of the nodes eliminated at level l; off[levels] = T - 1
};

GVI_HD int cr_count(int T, int l) { return (((T - 1) >> l) + 1) >> 1; }  // nodes eliminated at level l

GVI_HD void cr_make_geom(CrGeom& g, int T) {
    g.T = T;
    int lv = 0;
    while ((1 << lv) < T) ++lv;
    g.levels = lv;
    int o = 0;
    for (int l = 0; l <= CR_MAX_LEVELS; ++l) {
        g.off[l] = o;
        if (l < lv) o += cr_count(T, l);
    }
}

GVI_HD int cr_ctz(int j) {  // j > 0
#ifdef __CUDA_ARCH__
    return __ffs(j) - 1;
#else
    return __builtin_ctz((unsigned)j);
#endif
}

GVI_HD int cr_slot(const CrGeom& g, int j) {
    if (j == 0) return 0;
    if (j == g.T) return 1;
    const int l = cr_ctz(j);
    return 2 + g.off[l] + (j >> (l + 1));
}

// record r, element e of a [.. x NE] record array in the 32-record interleaved layout
GVI_HD size_t cr_rec(size_t r, int e, int NE) { return (r >> 5) * (size_t)(32 * NE) + (size_t)e * 32 + (r & 31); }
GVI_HD size_t cr_rec_capacity(size_t nrec, int NE) { return ((nrec + 31) >> 5) * (size_t)(32 * NE); }

// working arrays of one tile (shared memory on the device), slot indexed, structure of arrays with stride NS
template <int D>
struct CrView {
    int NS;
    double* Dn;  // [D*D][NS] Schur-updated diagonal blocks; after the forward pass reused for the covariance diagonal
    double* P;   // [D*D][NS] coupling of a node to its next alive node; reused for the covariance couplings
    double* g;   // [D][NS] right-hand side; reused for the solution
    // optional per-node log det of the pivots (batches of independent problems need log det PER PROBLEM): local node j of
    // this view is node min(ld_base + j ld_mul, ld_nmax) of its CrArgs level, which is chain node min(that ld_stride, ld_max)
    double* ldnode = nullptr;
    long long ld_base = 0, ld_mul = 1, ld_nmax = 0, ld_stride = 1, ld_max = 0;
};
template <int D>
GVI_HD size_t cr_ld_index(const CrView<D>& v, int j) {
    long long a = v.ld_base + (long long)j * v.ld_mul;
    if (a > v.ld_nmax) a = v.ld_nmax;
    a *= v.ld_stride;
    if (a > v.ld_max) a = v.ld_max;
    return (size_t)a;
}

template <int D>
struct CrRec {
    double* G;     // Dinv * A[j,k]
    double* H;     // Dinv * A[i,j]^T
    double* Dinv;  // inverse of the pivot block
    double* y;     // Dinv * g_j
};

template <int D>
GVI_HD void cr_ld(Mat<D>& A, const double* base, int NS, int s) {
#pragma unroll
    for (int e = 0; e < D * D; ++e) A.a[e] = base[(size_t)e * NS + s];
}
template <int D>
GVI_HD void cr_st(double* base, int NS, int s, const Mat<D>& A) {
#pragma unroll
    for (int e = 0; e < D * D; ++e) base[(size_t)e * NS + s] = A.a[e];
}
template <int D>
GVI_HD void cr_ldv(Vec<D>& v, const double* base, int NS, int s) {
#pragma unroll
    for (int e = 0; e < D; ++e) v.a[e] = base[(size_t)e * NS + s];
}
template <int D>
GVI_HD void cr_stv(double* base, int NS, int s, const Vec<D>& v) {
#pragma unroll
    for (int e = 0; e < D; ++e) base[(size_t)e * NS + s] = v.a[e];
}
template <int D>
GVI_HD void cr_rec_ld(Mat<D>& A, const double* base, size_t r) {
#pragma unroll
    for (int e = 0; e < D * D; ++e) A.a[e] = base[cr_rec(r, e, D * D)];
}
template <int D>
GVI_HD void cr_rec_st(double* base, size_t r, const Mat<D>& A) {
#pragma unroll
    for (int e = 0; e < D * D; ++e) base[cr_rec(r, e, D * D)] = A.a[e];
}
template <int D>
GVI_HD void cr_rec_ldv(Vec<D>& v, const double* base, size_t r) {
#pragma unroll
    for (int e = 0; e < D; ++e) v.a[e] = base[cr_rec(r, e, D)];
}
template <int D>
GVI_HD void cr_rec_stv(double* base, size_t r, const Vec<D>& v) {
#pragma unroll
    for (int e = 0; e < D; ++e) base[cr_rec(r, e, D)] = v.a[e];
}

// ------------------------------------------------------------------------------------------------------------------
// Node arithmetic, D workers per node.  Worker c (0 <= c < D) of a node produces column c (forward) or row c (backward)
// of every block the elimination writes; what it needs beyond that (pivot factor, the full neighbouring blocks) it
// reads from shared memory or recomputes -- the D workers of a node never exchange values, so the only ordering the
// caller has to provide are the two phase boundaries documented below.  Per worker this is ~1/D of the block products
// of a one-thread-per-node elimination: the levels of the reduction are latency bound, so the depth per level is what
// counts.
// ------------------------------------------------------------------------------------------------------------------
template <int D>
GVI_HD void cr_ld_col(Vec<D>& x, const double* base, int NS, int s, int c) {
#pragma unroll
    for (int i = 0; i < D; ++i) x.a[i] = base[(size_t)(i + c * D) * NS + s];
}
template <int D>
GVI_HD void cr_ld_row(Vec<D>& x, const double* base, int NS, int s, int c) {
#pragma unroll
    for (int m = 0; m < D; ++m) x.a[m] = base[(size_t)(c + m * D) * NS + s];
}

// what worker c of an elimination keeps between its two phases
template <int D>
struct CrElim {
    int si;
    Mat<D> Pi;         // coupling (i, j)
    Vec<D> Pirow;      // its row c
    Vec<D> Gc, Hc, y;  // column c of G and H, all of y
};

// Forward elimination of node j = (2t+1) 2^l, phase A: pivot factor, column c of the record, column c of the Schur
// update of the RIGHT neighbour (every right neighbour is updated by exactly one elimination of the level and phase A
// only reads blocks of j and the coupling of i, so phase A is race free).
// BATCH (batches of independent problems, gvib200_batch_iterate): the node's own log det goes to v.ldnode and a factor
// that failed is replaced by zeros; the default instantiation is the plain elimination.
template <int D, bool RHS, bool BATCH = false>
GVI_HD bool cr_fwd_A(const CrView<D>& v, const CrRec<D>& rec, size_t rec_base, const CrGeom& gm, int l, int t, int c,
                     CrElim<D>& el, LogDetAcc& ld) {
    const int j = (2 * t + 1) << l;
    const int i = j - (1 << l);
    const int k = (j + (1 << l) < gm.T) ? j + (1 << l) : gm.T;
    const int sj = 2 + gm.off[l] + t;
    const int sk = cr_slot(gm, k);
    el.si = cr_slot(gm, i);
    const size_t r = rec_base + (size_t)(sj - 2);
    Mat<D> Dt, Pj, L;
    Vec<D> Pjc, ec, Dic;
    double rd[D];
    cr_ld<D>(Dt, v.Dn, v.NS, sj);
    cr_ld<D>(Pj, v.P, v.NS, sj);
    cr_ld<D>(el.Pi, v.P, v.NS, el.si);
    cr_ld_col<D>(Pjc, v.P, v.NS, sj, c);
    cr_ld_row<D>(el.Pirow, v.P, v.NS, el.si, c);
    symmetrize<D>(Dt);
    bool ok;
    if constexpr (BATCH) {
        // the pivots count once per node: worker 0 merges the node's product into its running one and stores the node's own
        // log det
        LogDetAcc nd;
        ok = chol_factor<D>(L, rd, Dt, nd);
        if (c == 0) {
            ld.m *= nd.m;
            ld.e += nd.e;
            ld.normalize();
            if (v.ldnode != nullptr) v.ldnode[cr_ld_index<D>(v, j)] = ok ? nd.value() : nan("");
        }
        if (!ok) {
            // A pivot that is not positive has filled the factor with NaN.  A NaN would not stay inside its problem -- the
            // exactly-zero coupling blocks between two problems turn 0 * NaN into NaN on the other side -- so the factor is
            // replaced by zeros: everything downstream of this node is finite garbage, the node's log det is NaN and its
            // problem's cost with it.
#pragma unroll
            for (int e = 0; e < D * D; ++e) L.a[e] = 0.0;
#pragma unroll
            for (int m = 0; m < D; ++m) rd[m] = 0.0;
        }
    } else {
        LogDetAcc other;  // the pivots count once per node: worker 0 carries them
        ok = chol_factor<D>(L, rd, Dt, c == 0 ? ld : other);
    }
#pragma unroll
    for (int m = 0; m < D; ++m) ec.a[m] = (m == c) ? 1.0 : 0.0;
    chol_solve<D>(el.Gc, L, rd, Pjc);       // G = Dinv Pj
    chol_solve<D>(el.Hc, L, rd, el.Pirow);  // H = Dinv Pi^T
#pragma unroll
    for (int m = 0; m < D; ++m) {
        rec.G[cr_rec(r, m + c * D, D * D)] = el.Gc.a[m];
        rec.H[cr_rec(r, m + c * D, D * D)] = el.Hc.a[m];
    }
    // A pass is either a solve (RHS) or a selected inverse: the pivot inverse itself is only read by the Takahashi
    // recursion (cr_bwd_selinv_compute), so a solve pass neither computes nor stores it.
    if (!RHS) {
        chol_solve<D>(Dic, L, rd, ec);
#pragma unroll
        for (int m = 0; m < D; ++m) rec.Dinv[cr_rec(r, m + c * D, D * D)] = Dic.a[m];
    }
    // right neighbour: Dn[k] -= Pj^T G
#pragma unroll
    for (int q = 0; q < D; ++q) {
        double sum = 0.0;
#pragma unroll
        for (int m = 0; m < D; ++m) sum = fma(Pj(m, q), el.Gc.a[m], sum);
        v.Dn[(size_t)(q + c * D) * v.NS + sk] -= sum;
    }
    if (RHS) {
        Vec<D> gv;
        cr_ldv<D>(gv, v.g, v.NS, sj);
        chol_solve<D>(el.y, L, rd, gv);
        double yc = 0.0, sum = 0.0;
#pragma unroll
        for (int m = 0; m < D; ++m) {
            yc = (m == c) ? el.y.a[m] : yc;
            sum = fma(Pjc.a[m], el.y.a[m], sum);
        }
        rec.y[cr_rec(r, c, D)] = yc;
        v.g[(size_t)c * v.NS + sk] -= sum;
    }
    return ok;
}

// phase B: column c of the Schur update of the LEFT neighbour and of its new coupling to k (again one writer per
// element; the old coupling was read in phase A)
template <int D, bool RHS>
GVI_HD void cr_fwd_B(const CrView<D>& v, const CrElim<D>& el, int c) {
#pragma unroll
    for (int q = 0; q < D; ++q) {
        double sh = 0.0, sg = 0.0;
#pragma unroll
        for (int m = 0; m < D; ++m) {
            sh = fma(el.Pi(q, m), el.Hc.a[m], sh);
            sg = fma(el.Pi(q, m), el.Gc.a[m], sg);
        }
        v.Dn[(size_t)(q + c * D) * v.NS + el.si] -= sh;
        v.P[(size_t)(q + c * D) * v.NS + el.si] = -sg;
    }
    if (RHS) {
        double sum = 0.0;
#pragma unroll
        for (int m = 0; m < D; ++m) sum = fma(el.Pirow.a[m], el.y.a[m], sum);
        v.g[(size_t)c * v.NS + el.si] -= sum;
    }
}

// Back substitution of node j at level l, component c: x_j = y_j - G x_k - H x_i (x lives in v.g)
template <int D>
GVI_HD void cr_bwd_solve(const CrView<D>& v, const CrRec<D>& rec, size_t rec_base, const CrGeom& gm, int l, int t, int c) {
    const int j = (2 * t + 1) << l;
    const int i = j - (1 << l);
    const int k = (j + (1 << l) < gm.T) ? j + (1 << l) : gm.T;
    const int sj = 2 + gm.off[l] + t, si = cr_slot(gm, i), sk = cr_slot(gm, k);
    const size_t r = rec_base + (size_t)(sj - 2);
    Vec<D> xi, xk;
    cr_ldv<D>(xi, v.g, v.NS, si);
    cr_ldv<D>(xk, v.g, v.NS, sk);
    double x = rec.y[cr_rec(r, c, D)];
#pragma unroll
    for (int m = 0; m < D; ++m) {
        x = fma(-rec.G[cr_rec(r, c + m * D, D * D)], xk.a[m], x);
        x = fma(-rec.H[cr_rec(r, c + m * D, D * D)], xi.a[m], x);
    }
    v.g[(size_t)c * v.NS + sj] = x;
}

// Takahashi recursion for node j at level l, row c.  On entry v.Dn holds Sigma_ii, Sigma_kk and v.P[si] = Sigma_{i,k};
// cr_bwd_selinv_compute only reads; after a barrier cr_bwd_selinv_store writes row c of v.Dn[sj] = Sigma_jj and of
// v.P[sj] = Sigma_{j,k}, and column c of v.P[si] = Sigma_{i,j} (the coupling of i is overwritten: hence the barrier).
template <int D>
struct CrSel {
    int si, sj;
    Vec<D> jj, jk, ji;  // row c of Sigma_jj, Sigma_jk, Sigma_ji
};

template <int D>
GVI_HD void cr_bwd_selinv_compute(const CrView<D>& v, const CrRec<D>& rec, size_t rec_base, const CrGeom& gm, int l, int t,
                                  int c, CrSel<D>& o) {
    const int j = (2 * t + 1) << l;
    const int i = j - (1 << l);
    const int k = (j + (1 << l) < gm.T) ? j + (1 << l) : gm.T;
    const int sk = cr_slot(gm, k);
    o.sj = 2 + gm.off[l] + t;
    o.si = cr_slot(gm, i);
    const size_t r = rec_base + (size_t)(o.sj - 2);
    Mat<D> G, H, Sii, Skk, Sik;
    Vec<D> Gr, Hr;
    cr_rec_ld<D>(G, rec.G, r);
    cr_rec_ld<D>(H, rec.H, r);
#pragma unroll
    for (int m = 0; m < D; ++m) {
        Gr.a[m] = rec.G[cr_rec(r, c + m * D, D * D)];
        Hr.a[m] = rec.H[cr_rec(r, c + m * D, D * D)];
        o.jj.a[m] = rec.Dinv[cr_rec(r, c + m * D, D * D)];
    }
    cr_ld<D>(Sii, v.Dn, v.NS, o.si);
    cr_ld<D>(Skk, v.Dn, v.NS, sk);
    cr_ld<D>(Sik, v.P, v.NS, o.si);
    // Sigma_{j,k} = -(G Sigma_kk + H Sigma_ik);  Sigma_{j,i} = -(G Sigma_ik^T + H Sigma_ii)
#pragma unroll
    for (int q = 0; q < D; ++q) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int m = 0; m < D; ++m) {
            a = fma(Gr.a[m], Skk(m, q), a);
            a = fma(Hr.a[m], Sik(m, q), a);
            b = fma(Gr.a[m], Sik(q, m), b);
            b = fma(Hr.a[m], Sii(m, q), b);
        }
        o.jk.a[q] = -a;
        o.ji.a[q] = -b;
    }
    // Sigma_jj = Dinv - Sigma_jk G^T - Sigma_ji H^T
#pragma unroll
    for (int q = 0; q < D; ++q) {
        double a = o.jj.a[q];
#pragma unroll
        for (int m = 0; m < D; ++m) {
            a = fma(-o.jk.a[m], G(q, m), a);
            a = fma(-o.ji.a[m], H(q, m), a);
        }
        o.jj.a[q] = a;
    }
}

template <int D>
GVI_HD void cr_bwd_selinv_store(const CrView<D>& v, const CrSel<D>& o, int c) {
#pragma unroll
    for (int q = 0; q < D; ++q) {
        v.Dn[(size_t)(c + q * D) * v.NS + o.sj] = o.jj.a[q];
        v.P[(size_t)(c + q * D) * v.NS + o.sj] = o.jk.a[q];
        v.P[(size_t)(q + c * D) * v.NS + o.si] = o.ji.a[q];
    }
}

// The system left after all levels: nodes 0 and T (slots 0, 1) coupled by P[slot 0]; T == 0: a single node.
// Solves it in place: v.g <- x (RHS), v.Dn <- Sigma diagonal blocks and v.P[0] <- Sigma_{0,T} (SELINV).
template <int D, bool RHS, bool SELINV, bool BATCH = false>
GVI_HD bool cr_top2(const CrView<D>& v, int T, LogDetAcc& ld) {
    Mat<D> A0, A1, P, I0, S1, G, Tm;
    cr_ld<D>(A0, v.Dn, v.NS, 0);
    symmetrize<D>(A0);
    bool ok;
    if (BATCH && v.ldnode != nullptr) {
        LogDetAcc nd;
        ok = spd_inverse<D>(I0, A0, nd);
        v.ldnode[cr_ld_index<D>(v, 0)] = ok ? nd.value() : nan("");
        if (!ok) mat_zero<D>(I0);  // finite garbage instead of NaN (see cr_fwd_A)
        ld.m *= nd.m;
        ld.e += nd.e;
        ld.normalize();
    } else {
        ok = spd_inverse<D>(I0, A0, ld);
    }
    Vec<D> g0, g1, y0, x1, tv;
    if (RHS) {
        cr_ldv<D>(g0, v.g, v.NS, 0);
        mv<D>(y0, I0, g0);
    }
    if (T == 0) {
        if (RHS) cr_stv<D>(v.g, v.NS, 0, y0);
        if (SELINV) cr_st<D>(v.Dn, v.NS, 0, I0);
        return ok;
    }
    cr_ld<D>(A1, v.Dn, v.NS, 1);
    cr_ld<D>(P, v.P, v.NS, 0);
    mm<D>(G, I0, P);  // A0^-1 P
    mtm<D>(Tm, P, G);
#pragma unroll
    for (int e = 0; e < D * D; ++e) A1.a[e] -= Tm.a[e];
    symmetrize<D>(A1);
    if (BATCH && v.ldnode != nullptr) {
        LogDetAcc nd;
        const bool ok1 = spd_inverse<D>(S1, A1, nd);  // Sigma_TT
        ok = ok1 && ok;
        v.ldnode[cr_ld_index<D>(v, T)] = ok1 ? nd.value() : nan("");
        if (!ok1) mat_zero<D>(S1);
        ld.m *= nd.m;
        ld.e += nd.e;
        ld.normalize();
    } else {
        ok = spd_inverse<D>(S1, A1, ld) && ok;  // Sigma_TT
    }
    if (RHS) {
        cr_ldv<D>(g1, v.g, v.NS, 1);
        mtv<D>(tv, P, y0);
#pragma unroll
        for (int e = 0; e < D; ++e) g1.a[e] -= tv.a[e];
        mv<D>(x1, S1, g1);
        mv<D>(tv, G, x1);
#pragma unroll
        for (int e = 0; e < D; ++e) y0.a[e] -= tv.a[e];
        cr_stv<D>(v.g, v.NS, 0, y0);
        cr_stv<D>(v.g, v.NS, 1, x1);
    }
    if (SELINV) {
        Mat<D> S01, S00;
        mm<D>(Tm, G, S1);
#pragma unroll
        for (int e = 0; e < D * D; ++e) S01.a[e] = -Tm.a[e];  // Sigma_{0,T} = -G Sigma_TT
        mmt<D>(Tm, S01, G);
#pragma unroll
        for (int e = 0; e < D * D; ++e) S00.a[e] = I0.a[e] - Tm.a[e];
        symmetrize<D>(S00);
        cr_st<D>(v.Dn, v.NS, 0, S00);
        cr_st<D>(v.Dn, v.NS, 1, S1);
        cr_st<D>(v.P, v.NS, 0, S01);
    }
    return ok;
}

}  // namespace gvib200

// ------------------------------------------------------------------------------------------------------------------
// Glue shared by the CUDA kernels (kernels.cuh) and the host harness: a "thread" tid of nthreads strides the copies.
// ------------------------------------------------------------------------------------------------------------------
namespace gvib200 {

template <int D>
struct CrArgs {
    int n;  // nodes of the chain
    int T;  // links per tile (tile k owns nodes kT .. min((k+1)T, n-1))
    int K;  // tiles; K == 0: the whole chain is handled by the top kernel alone
    // the system: diag[n][D*D], off[n-1][D*D] (block (i, i+1)), rhs[n][D] (null without a right-hand side)
    const double* Dg;
    const double* Og;
    const double* g;
    CrRec<D> rec;  // elimination records of the tiles, K * (T - 1) records
    // reduced system on the separators (input of the top), node k = separator kT:
    //   diag(k) = rDn[k] + rCL[k] (k < K) + rCR[k-1] (k > 0), coupling rO[k], rhs rg + rgl + rgr likewise
    double *rDn, *rCL, *rCR, *rO, *rg, *rgl, *rgr;
    double *tD, *tO, *tx;  // results of the top on the separators: [K+1][DD], [K][DD], [K+1][D]
    double *x, *cD, *cO;   // outputs: solution [n][D]; selected inverse diag [n][DD], off [n-1][DD]
    double* ld;            // [K + 1] partial log determinants (tiles, then the top)
    double* ldout;         // optional: the top kernel also writes the total log determinant (sum of ld[0..K]) here
    int* notspd;
    // optional fusions of the line-search candidate (NGDGH::onestep_linesearch, ngd/NGD-GH-impl.h:129-148):
    //   Dg2 != null: the system is Dg + alpha (Dg2 - Dg) (same for the off-diagonal blocks) and is also written to
    //                Dout / Oout while it is loaded            (Lambda' = Lambda + a (Vddmu - Lambda))
    //   xout != null: xout = xbase + xalpha * x is written next to the solution   (mu' = mu + a dmu)
    const double *Dg2, *Og2;
    double alpha;
    double *Dout, *Oout;
    const double* xbase;
    double xalpha;
    double* xout;
    // batches of independent problems (optional): alpha_node[i] replaces alpha for the blocks of node i of this level
    // (diagonal block i and coupling (i, i+1)); ldnode receives the log det of every node's pivot block, node i of this
    // level being chain node min(i ld_stride, ld_max)
    const double* alpha_node;
    double* ldnode;
    long long ld_stride, ld_max;
};

template <int D, bool BATCH = false>
GVI_HD double cr_sys_diag(const CrArgs<D>& a, size_t idx) {
    double v = a.Dg[idx];
    if (a.Dg2 != nullptr) {
        const double al = (BATCH && a.alpha_node) ? a.alpha_node[idx / (D * D)] : a.alpha;
        v = v + al * (a.Dg2[idx] - v);
        a.Dout[idx] = v;
    }
    return v;
}
template <int D, bool BATCH = false>
GVI_HD double cr_sys_off(const CrArgs<D>& a, size_t idx) {
    double v = a.Og[idx];
    if (a.Og2 != nullptr) {
        const double al = (BATCH && a.alpha_node) ? a.alpha_node[idx / (D * D)] : a.alpha;
        v = v + al * (a.Og2[idx] - v);
        a.Oout[idx] = v;
    }
    return v;
}

GVI_HD int cr_pad_slots(int nodes) { return nodes | 1; }  // odd stride: conflict-free transposing copies

// shared-memory doubles needed by one tile CTA / by the top CTA
template <int D>
GVI_HD size_t cr_tile_doubles(int T) { return (size_t)(2 * D * D + D) * cr_pad_slots(T + 1); }
template <int D>
GVI_HD size_t cr_top_doubles(int n_top) {
    const size_t nrec = n_top > 2 ? (size_t)(n_top - 2) : 0;
    return (size_t)(2 * D * D + D) * cr_pad_slots(n_top) + 3 * cr_rec_capacity(nrec, D * D) + cr_rec_capacity(nrec, D);
}

template <int D>
GVI_HD CrView<D> cr_make_view(double* sm, int nodes) {
    CrView<D> v;
    v.NS = cr_pad_slots(nodes);
    v.Dn = sm;
    v.P = v.Dn + (size_t)D * D * v.NS;
    v.g = v.P + (size_t)D * D * v.NS;
    return v;
}

// tile -> working arrays.  Separator slots start from zero: they only collect this tile's Schur contributions.
// The copies are unrolled CR_UNROLL deep (loads first, then the scattered shared-memory stores) so that every thread
// keeps several global loads in flight: a tile CTA is alone on its SM and would otherwise pay one DRAM latency per
// element.
constexpr int CR_UNROLL = 8;

template <int D, bool RHS, bool BATCH = false>
GVI_HD void cr_tile_load(const CrArgs<D>& a, const CrView<D>& v, const CrGeom& gm, int n0, int tid, int nthreads) {
    constexpr int DD = D * D;
    const int Tk = gm.T;
    const int N = (Tk + 1) * DD;
    const bool fused = (a.Dg2 != nullptr);
    const size_t g0 = (size_t)n0 * DD;
    for (int base = tid; base < N; base += nthreads * CR_UNROLL) {
        // all loads first (the candidate stores below may alias them as far as the compiler knows)
        double dv[CR_UNROLL], ov[CR_UNROLL], d2[CR_UNROLL], o2[CR_UNROLL];
#pragma unroll
        for (int u = 0; u < CR_UNROLL; ++u) {
            const int idx = base + u * nthreads;
            dv[u] = ov[u] = d2[u] = o2[u] = 0.0;
            if (idx < N) {
                const int node = idx / DD;
                if (node != 0 && node != Tk) {
                    dv[u] = a.Dg[g0 + idx];
                    if (fused) d2[u] = a.Dg2[g0 + idx];
                }
                if (node < Tk) {
                    ov[u] = a.Og[g0 + idx];
                    if (fused) o2[u] = a.Og2[g0 + idx];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < CR_UNROLL; ++u) {
            const int idx = base + u * nthreads;
            if (idx < N) {
                const int node = idx / DD, e = idx - node * DD;
                const int s = cr_slot(gm, node);
                if (fused) {
                    const double al = (BATCH && a.alpha_node) ? a.alpha_node[n0 + node] : a.alpha;
                    dv[u] = dv[u] + al * (d2[u] - dv[u]);
                    ov[u] = ov[u] + al * (o2[u] - ov[u]);
                    if (node != 0 && node != Tk) a.Dout[g0 + idx] = dv[u];
                    if (node < Tk) a.Oout[g0 + idx] = ov[u];
                }
                v.Dn[(size_t)e * v.NS + s] = (node == 0 || node == Tk) ? 0.0 : dv[u];
                if (node < Tk) v.P[(size_t)e * v.NS + s] = ov[u];
            }
        }
    }
    if (RHS) {
        const int NV = (Tk + 1) * D;
        for (int base = tid; base < NV; base += nthreads * CR_UNROLL) {
            double gv[CR_UNROLL];
#pragma unroll
            for (int u = 0; u < CR_UNROLL; ++u) {
                const int idx = base + u * nthreads;
                gv[u] = 0.0;
                if (idx < NV) {
                    const int node = idx / D;
                    if (node != 0 && node != Tk) gv[u] = a.g[(size_t)n0 * D + idx];
                }
            }
#pragma unroll
            for (int u = 0; u < CR_UNROLL; ++u) {
                const int idx = base + u * nthreads;
                if (idx < NV) {
                    const int node = idx / D, e = idx - node * D;
                    v.g[(size_t)e * v.NS + cr_slot(gm, node)] = gv[u];
                }
            }
        }
    }
}

// what the tile hands to the top: its contributions to the two separators and their coupling
template <int D, bool RHS, bool BATCH = false>
GVI_HD void cr_tile_store_reduced(const CrArgs<D>& a, const CrView<D>& v, const CrGeom& gm, int tile, int n0, int tid,
                                  int nthreads) {
    constexpr int DD = D * D;
    const bool last = (tile == a.K - 1);
    for (int e = tid; e < DD; e += nthreads) {
        a.rDn[(size_t)tile * DD + e] = cr_sys_diag<D, BATCH>(a, (size_t)n0 * DD + e);
        if (last) a.rDn[(size_t)a.K * DD + e] = cr_sys_diag<D, BATCH>(a, (size_t)(a.n - 1) * DD + e);
        a.rCL[(size_t)tile * DD + e] = v.Dn[(size_t)e * v.NS + 0];
        a.rCR[(size_t)tile * DD + e] = v.Dn[(size_t)e * v.NS + 1];
        a.rO[(size_t)tile * DD + e] = v.P[(size_t)e * v.NS + 0];
    }
    if (RHS) {
        for (int e = tid; e < D; e += nthreads) {
            a.rg[(size_t)tile * D + e] = a.g[(size_t)n0 * D + e];
            if (last) a.rg[(size_t)a.K * D + e] = a.g[(size_t)(a.n - 1) * D + e];
            a.rgl[(size_t)tile * D + e] = v.g[(size_t)e * v.NS + 0];
            a.rgr[(size_t)tile * D + e] = v.g[(size_t)e * v.NS + 1];
        }
    }
}

// top: gather the separator system (or, K == 0, the chain itself) into the working arrays
template <int D, bool RHS, bool BATCH = false>
GVI_HD void cr_top_load(const CrArgs<D>& a, const CrView<D>& v, const CrGeom& gm, int tid, int nthreads) {
    constexpr int DD = D * D;
    const int nt = gm.T + 1;  // nodes of the top
    for (int idx = tid; idx < nt * DD; idx += nthreads) {
        const int node = idx / DD, e = idx - node * DD;
        const int s = cr_slot(gm, node);
        double dv;
        if (a.K == 0) {
            dv = cr_sys_diag<D, BATCH>(a, (size_t)node * DD + e);
            if (node < nt - 1) v.P[(size_t)e * v.NS + s] = cr_sys_off<D, BATCH>(a, (size_t)node * DD + e);
        } else {
            dv = a.rDn[(size_t)node * DD + e];
            if (node < a.K) dv += a.rCL[(size_t)node * DD + e];
            if (node > 0) dv += a.rCR[(size_t)(node - 1) * DD + e];
            if (node < nt - 1) v.P[(size_t)e * v.NS + s] = a.rO[(size_t)node * DD + e];
        }
        v.Dn[(size_t)e * v.NS + s] = dv;
    }
    if (RHS) {
        for (int idx = tid; idx < nt * D; idx += nthreads) {
            const int node = idx / D, e = idx - node * D;
            double gv;
            if (a.K == 0) {
                gv = a.g[(size_t)node * D + e];
            } else {
                gv = a.rg[(size_t)node * D + e];
                if (node < a.K) gv += a.rgl[(size_t)node * D + e];
                if (node > 0) gv += a.rgr[(size_t)(node - 1) * D + e];
            }
            v.g[(size_t)e * v.NS + cr_slot(gm, node)] = gv;
        }
    }
}

// working arrays -> AoS results for nodes [0, count) (diag / solution) and couplings [0, ncoup), written at node offset n0
template <int D, bool RHS, bool SELINV>
GVI_HD void cr_store_results(const CrView<D>& v, const CrGeom& gm, int count, int ncoup, double* x, double* cD, double* cO,
                             size_t n0, int tid, int nthreads, const double* xbase = nullptr, double xalpha = 0.0,
                             double* xout = nullptr) {
    constexpr int DD = D * D;
    if (SELINV) {
        for (int idx = tid; idx < count * DD; idx += nthreads) {
            const int node = idx / DD, e = idx - node * DD;
            const int s = cr_slot(gm, node);
            const int et = (e % D) * D + e / D;  // the rows of a diagonal block come from different workers: symmetrize
            cD[n0 * DD + idx] = 0.5 * (v.Dn[(size_t)e * v.NS + s] + v.Dn[(size_t)et * v.NS + s]);
            if (node < ncoup) cO[n0 * DD + idx] = v.P[(size_t)e * v.NS + s];
        }
    }
    if (RHS) {
        const int NV = count * D;
        for (int base = tid; base < NV; base += nthreads * CR_UNROLL) {
            double bv[CR_UNROLL];
#pragma unroll
            for (int u = 0; u < CR_UNROLL; ++u) {
                const int idx = base + u * nthreads;
                bv[u] = (xout != nullptr && idx < NV) ? xbase[n0 * D + idx] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < CR_UNROLL; ++u) {
                const int idx = base + u * nthreads;
                if (idx < NV) {
                    const int node = idx / D, e = idx - node * D;
                    const double xv = v.g[(size_t)e * v.NS + cr_slot(gm, node)];
                    x[n0 * D + idx] = xv;
                    if (xout != nullptr) xout[n0 * D + idx] = bv[u] + xalpha * xv;
                }
            }
        }
    }
}

// tile, backward: seed the two separators from the results of the top
template <int D, bool RHS, bool SELINV>
GVI_HD void cr_tile_seed(const CrArgs<D>& a, const CrView<D>& v, int tile, int tid, int nthreads) {
    constexpr int DD = D * D;
    if (SELINV) {
        for (int e = tid; e < DD; e += nthreads) {
            v.Dn[(size_t)e * v.NS + 0] = a.tD[(size_t)tile * DD + e];
            v.Dn[(size_t)e * v.NS + 1] = a.tD[(size_t)(tile + 1) * DD + e];
            v.P[(size_t)e * v.NS + 0] = a.tO[(size_t)tile * DD + e];
        }
    }
    if (RHS) {
        for (int e = tid; e < D; e += nthreads) {
            v.g[(size_t)e * v.NS + 0] = a.tx[(size_t)tile * D + e];
            v.g[(size_t)e * v.NS + 1] = a.tx[(size_t)(tile + 1) * D + e];
        }
    }
}

}  // namespace gvib200

// ------------------------------------------------------------------------------------------------------------------
// Multi-GPU: the chain is cut along the time axis, rank r owns the links [r m, (r+1) m) and both of its end nodes (the
// end nodes are shared with the neighbouring ranks; each rank holds only ITS share of a shared diagonal block / rhs).
// Per pass every rank reduces its segment to its two end nodes with the same tile kernels applied twice (tiles ->
// separator chain -> one "tile" over the separator chain), the P boundary records meet in ONE all-gather, every rank
// solves the (P+1)-node chain of rank boundaries redundantly and walks back down.  Boundary record of a rank:
//   [ Dfirst | Dlast | CL | CR | O | gfirst | glast | gl | gr ]  = 5 d^2 + 4 d doubles.
// ------------------------------------------------------------------------------------------------------------------
namespace gvib200 {

template <int D>
GVI_HD constexpr int cr_boundary_doubles() { return 5 * D * D + 4 * D; }

// level-1 separator system (three-term) -> plain arrays D1[K+1], O1[K], g1[K+1]
template <int D, bool RHS>
GVI_HD void cr_sum_level(const CrArgs<D>& a, double* D1, double* O1, double* g1, int tid, int nthreads) {
    constexpr int DD = D * D;
    const int nt = a.K + 1;
    for (int idx = tid; idx < nt * DD; idx += nthreads) {
        const int node = idx / DD;
        double dv = a.rDn[idx];
        if (node < a.K) {
            dv += a.rCL[idx];
            O1[idx] = a.rO[idx];
        }
        if (node > 0) dv += a.rCR[idx - DD];
        D1[idx] = dv;
    }
    if (RHS) {
        for (int idx = tid; idx < nt * D; idx += nthreads) {
            const int node = idx / D;
            double gv = a.rg[idx];
            if (node < a.K) gv += a.rgl[idx];
            if (node > 0) gv += a.rgr[idx - D];
            g1[idx] = gv;
        }
    }
}

// reduced arrays of the single mid tile -> this rank's boundary record
template <int D, bool RHS>
GVI_HD void cr_pack_boundary(const CrArgs<D>& mid, double* rec, int tid, int nthreads) {
    constexpr int DD = D * D;
    for (int e = tid; e < DD; e += nthreads) {
        rec[e] = mid.rDn[e];
        rec[DD + e] = mid.rDn[DD + e];
        rec[2 * DD + e] = mid.rCL[e];
        rec[3 * DD + e] = mid.rCR[e];
        rec[4 * DD + e] = mid.rO[e];
    }
    for (int e = tid; e < D; e += nthreads) {
        double* gv = rec + 5 * DD;
        gv[e] = RHS ? mid.rg[e] : 0.0;
        gv[D + e] = RHS ? mid.rg[D + e] : 0.0;
        gv[2 * D + e] = RHS ? mid.rgl[e] : 0.0;
        gv[3 * D + e] = RHS ? mid.rgr[e] : 0.0;
    }
}

// all boundary records -> the chain of rank boundaries (P + 1 nodes)
template <int D>
GVI_HD void cr_build_global(int P, const double* recs, double* Dt, double* Ot, double* gt, int tid, int nthreads) {
    constexpr int DD = D * D;
    constexpr int NB = 5 * D * D + 4 * D;
    for (int idx = tid; idx < (P + 1) * DD; idx += nthreads) {
        const int node = idx / DD, e = idx - node * DD;
        double dv = 0.0;
        if (node > 0) dv += recs[(size_t)(node - 1) * NB + DD + e] + recs[(size_t)(node - 1) * NB + 3 * DD + e];  // Dlast + CR
        if (node < P) {
            dv += recs[(size_t)node * NB + e] + recs[(size_t)node * NB + 2 * DD + e];  // Dfirst + CL
            Ot[idx] = recs[(size_t)node * NB + 4 * DD + e];
        }
        Dt[idx] = dv;
    }
    for (int idx = tid; idx < (P + 1) * D; idx += nthreads) {
        const int node = idx / D, e = idx - node * D;
        double gv = 0.0;
        if (node > 0) gv += recs[(size_t)(node - 1) * NB + 5 * DD + D + e] + recs[(size_t)(node - 1) * NB + 5 * DD + 3 * D + e];
        if (node < P) gv += recs[(size_t)node * NB + 5 * DD + e] + recs[(size_t)node * NB + 5 * DD + 2 * D + e];
        gt[idx] = gv;
    }
}

// results on the rank boundaries -> seeds of this rank's mid tile
template <int D, bool RHS, bool SELINV>
GVI_HD void cr_seed_mid(const CrArgs<D>& mid, int rank, const double* xt, const double* cDt, const double* cOt, int tid,
                        int nthreads) {
    constexpr int DD = D * D;
    if (SELINV)
        for (int e = tid; e < DD; e += nthreads) {
            mid.tD[e] = cDt[(size_t)rank * DD + e];
            mid.tD[DD + e] = cDt[(size_t)(rank + 1) * DD + e];
            mid.tO[e] = cOt[(size_t)rank * DD + e];
        }
    if (RHS)
        for (int e = tid; e < D; e += nthreads) {
            mid.tx[e] = xt[(size_t)rank * D + e];
            mid.tx[D + e] = xt[(size_t)(rank + 1) * D + e];
        }
}

}  // namespace gvib200
