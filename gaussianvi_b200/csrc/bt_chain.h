// Block-tridiagonal SPD engine: partitioned (nested-dissection) block Cholesky, solve, log det
// and Takahashi selected inverse (diagonal + first off-diagonal blocks of the inverse).
//
// Replaces, for the joint precision pattern fixed at gvibase/GVI-GH.h:214-230:
//   GVIGH::inverse_GBP + calculate_factor_message   gvibase/GVI-GH-GBP-impl.h:245-342
//   EigenWrapper::inv_sparse (Takahashi on LDLT)     helpers/EigenWrapper.h:336-381
//   SparseLDLT(Precision).vectorD().log().sum()      gvibase/GVI-GH-GBP-impl.h:234-238
//   ConjugateGradient(Vddmu).solve(-Vdmu)            ngd/NGD-GH-impl.h:59-60 (direct solve instead)
//
// One level:  n nodes, separators at indices 0, L, 2L, ... and n-1; the interior of every segment
// (a = kL, b = min((k+1)L, n-1)) is eliminated left to right by ONE worker (a CUDA thread, or a loop
// iteration of the host test harness), carrying the fill block W = A[a, j] of the left separator.
// The Schur complement on the separators is again block tridiagonal with n' = K + 1 nodes and is
// handed to the next level; the last level is solved serially.  Going back down, each segment
// recovers its interior solution / selected-inverse blocks from its two separators.
//
// Everything here is __host__ __device__ and free of CUDA intrinsics so that
// tests/cpp/bt_host_emu.cpp can run the very same arithmetic on the CPU against the oracle.
#pragma once
#include "smallmat.h"

namespace gvib200 {

template <int D>
struct BtLevel {
    int n;  // nodes of this level
    int L;  // segment length (>= 1); K = ceil((n-1)/L) segments; L == 0 marks the serial top level
    int K;
    // system of this level: diag(i) = Dn[i] + CL[i] (i < n-1) + CR[i-1] (i > 0); CL/CR may be null
    const double* Dn;
    const double* CL;
    const double* CR;
    const double* O;  // [n-1] block (i, i+1)
    const double* g;  // rhs, same three-term structure; all null when no rhs
    const double* gl;
    const double* gr;
    // elimination record (indexed by node of this level)
    double* G;     // Dinv * O_j
    double* H;     // Dinv * W_j^T
    double* Dinv;  // inverse of the eliminated pivot block
    double* y;     // Dinv * g_j (rhs mode)
    double* ld;    // [max(K,1)] partial log det of the pivots eliminated by worker k
    // reduced system written for the next level (n' = K + 1)
    double* rDn;
    double* rCL;
    double* rCR;
    double* rO;
    double* rg;
    double* rgl;
    double* rgr;
    int* notspd;  // set to 1 when a pivot block is not positive definite
};

template <int D>
GVI_HD void bt_load_diag(Mat<D>& A, const BtLevel<D>& lv, int i) {
    mat_load<D>(A, lv.Dn + (size_t)i * D * D);
    if (lv.CL != nullptr) {
        if (i < lv.n - 1) {
            const double* p = lv.CL + (size_t)i * D * D;
#pragma unroll
            for (int e = 0; e < D * D; ++e) A.a[e] += p[e];
        }
        if (i > 0) {
            const double* p = lv.CR + (size_t)(i - 1) * D * D;
#pragma unroll
            for (int e = 0; e < D * D; ++e) A.a[e] += p[e];
        }
    }
}

template <int D>
GVI_HD void bt_load_rhs(Vec<D>& v, const BtLevel<D>& lv, int i) {
    vec_load<D>(v, lv.g + (size_t)i * D);
    if (lv.gl != nullptr) {
        if (i < lv.n - 1) {
            const double* p = lv.gl + (size_t)i * D;
#pragma unroll
            for (int e = 0; e < D; ++e) v.a[e] += p[e];
        }
        if (i > 0) {
            const double* p = lv.gr + (size_t)(i - 1) * D;
#pragma unroll
            for (int e = 0; e < D; ++e) v.a[e] += p[e];
        }
    }
}

// Forward elimination of the interior of segment k.
template <int D, bool RHS>
GVI_HD void bt_forward_segment(const BtLevel<D>& lv, int k) {
    const int a = k * lv.L;
    const int b = (a + lv.L < lv.n - 1) ? a + lv.L : lv.n - 1;
    constexpr int DD = D * D;
    Mat<D> W, CLacc, Dt, Oj, Dinv, G, H, T;
    Vec<D> glacc, gcur, yv, tv;
    mat_zero<D>(CLacc);
    vec_zero<D>(glacc);
    LogDetAcc ld;
    bool ok = true;
    // reduced node k keeps the (summed) diagonal / rhs of separator a
    bt_load_diag<D>(T, lv, a);
    mat_store<D>(lv.rDn + (size_t)k * DD, T);
    if (RHS) {
        bt_load_rhs<D>(tv, lv, a);
        vec_store<D>(lv.rg + (size_t)k * D, tv);
    }
    if (k == lv.K - 1) {  // the last separator n-1 becomes reduced node K
        bt_load_diag<D>(T, lv, lv.n - 1);
        mat_store<D>(lv.rDn + (size_t)lv.K * DD, T);
        if (RHS) {
            bt_load_rhs<D>(tv, lv, lv.n - 1);
            vec_store<D>(lv.rg + (size_t)lv.K * D, tv);
        }
    }
    mat_load<D>(W, lv.O + (size_t)a * DD);  // block (a, a+1)
    if (b == a + 1) {                         // empty interior: the coupling is untouched
        mat_store<D>(lv.rO + (size_t)k * DD, W);
        mat_store<D>(lv.rCL + (size_t)k * DD, CLacc);
        mat_store<D>(lv.rCR + (size_t)k * DD, CLacc);
        if (RHS) {
            vec_store<D>(lv.rgl + (size_t)k * D, glacc);
            vec_store<D>(lv.rgr + (size_t)k * D, glacc);
        }
        lv.ld[k] = 0.0;
        return;
    }
    bt_load_diag<D>(Dt, lv, a + 1);
    if (RHS) bt_load_rhs<D>(gcur, lv, a + 1);
    for (int j = a + 1; j < b; ++j) {
        ok = spd_inverse<D>(Dinv, Dt, ld) && ok;
        mat_load<D>(Oj, lv.O + (size_t)j * DD);  // block (j, j+1)
        mm<D>(G, Dinv, Oj);
        mmt<D>(H, Dinv, W);  // Dinv * W^T
        mat_store<D>(lv.G + (size_t)j * DD, G);
        mat_store<D>(lv.H + (size_t)j * DD, H);
        mat_store<D>(lv.Dinv + (size_t)j * DD, Dinv);
        // left separator: CL -= W Dinv W^T
        mm<D>(T, W, H);
#pragma unroll
        for (int e = 0; e < DD; ++e) CLacc.a[e] -= T.a[e];
        if (RHS) {
            mv<D>(yv, Dinv, gcur);
            vec_store<D>(lv.y + (size_t)j * D, yv);
            mv<D>(tv, W, yv);
#pragma unroll
            for (int e = 0; e < D; ++e) glacc.a[e] -= tv.a[e];
        }
        // fill towards the next node: A[a, j+1] = -W G
        mm<D>(T, W, G);
#pragma unroll
        for (int e = 0; e < DD; ++e) W.a[e] = -T.a[e];
        // Schur update of the next node: -O_j^T G
        mtm<D>(T, Oj, G);
        if (RHS) mtv<D>(tv, Oj, yv);
        if (j + 1 < b) {
            bt_load_diag<D>(Dt, lv, j + 1);
#pragma unroll
            for (int e = 0; e < DD; ++e) Dt.a[e] -= T.a[e];
            symmetrize<D>(Dt);
            if (RHS) {
                bt_load_rhs<D>(gcur, lv, j + 1);
#pragma unroll
                for (int e = 0; e < D; ++e) gcur.a[e] -= tv.a[e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < DD; ++e) T.a[e] = -T.a[e];
            symmetrize<D>(T);
            mat_store<D>(lv.rCR + (size_t)k * DD, T);
            if (RHS) {
#pragma unroll
                for (int e = 0; e < D; ++e) tv.a[e] = -tv.a[e];
                vec_store<D>(lv.rgr + (size_t)k * D, tv);
            }
        }
    }
    symmetrize<D>(CLacc);
    mat_store<D>(lv.rCL + (size_t)k * DD, CLacc);
    mat_store<D>(lv.rO + (size_t)k * DD, W);  // block (a, b) of the reduced system
    if (RHS) vec_store<D>(lv.rgl + (size_t)k * D, glacc);
    lv.ld[k] = ld.value();
    if (!ok) *lv.notspd = 1;
}

// Back-substitution: x on the separators comes from the reduced level (xr[K+1][D]).
template <int D>
GVI_HD void bt_backsolve_segment(const BtLevel<D>& lv, int k, const double* __restrict__ xr, double* __restrict__ x) {
    const int a = k * lv.L;
    const int b = (a + lv.L < lv.n - 1) ? a + lv.L : lv.n - 1;
    constexpr int DD = D * D;
    Vec<D> xa, xn, yv, t1, t2;
    Mat<D> G, H;
    vec_load<D>(xa, xr + (size_t)k * D);
    vec_load<D>(xn, xr + (size_t)(k + 1) * D);
    vec_store<D>(x + (size_t)a * D, xa);
    if (k == lv.K - 1) vec_store<D>(x + (size_t)(lv.n - 1) * D, xn);
    for (int j = b - 1; j > a; --j) {
        mat_load<D>(G, lv.G + (size_t)j * DD);
        mat_load<D>(H, lv.H + (size_t)j * DD);
        vec_load<D>(yv, lv.y + (size_t)j * D);
        mv<D>(t1, G, xn);
        mv<D>(t2, H, xa);
#pragma unroll
        for (int e = 0; e < D; ++e) xn.a[e] = yv.a[e] - t1.a[e] - t2.a[e];
        vec_store<D>(x + (size_t)j * D, xn);
    }
}

// Takahashi recursion on segment k.  cDr[K+1], cOr[K]: selected inverse of the reduced level.
// Writes cD[j] = Sigma_jj and cO[j] = Sigma_{j,j+1} for every node j in [a, b) (and cD[n-1]).
template <int D>
GVI_HD void bt_selinv_segment(const BtLevel<D>& lv, int k, const double* __restrict__ cDr,
                              const double* __restrict__ cOr, double* __restrict__ cD, double* __restrict__ cO) {
    const int a = k * lv.L;
    const int b = (a + lv.L < lv.n - 1) ? a + lv.L : lv.n - 1;
    constexpr int DD = D * D;
    Mat<D> Saa, Snn, Sna, G, H, Dinv, Sjn, Sja, T1, T2;
    mat_load<D>(Saa, cDr + (size_t)k * DD);
    mat_load<D>(Snn, cDr + (size_t)(k + 1) * DD);
    mat_load<D>(T1, cOr + (size_t)k * DD);  // Sigma_{a,b}
    mat_transpose<D>(Sna, T1);              // Sigma_{b,a}
    mat_store<D>(cD + (size_t)a * DD, Saa);
    if (k == lv.K - 1) mat_store<D>(cD + (size_t)(lv.n - 1) * DD, Snn);
    for (int j = b - 1; j > a; --j) {
        mat_load<D>(G, lv.G + (size_t)j * DD);
        mat_load<D>(H, lv.H + (size_t)j * DD);
        mat_load<D>(Dinv, lv.Dinv + (size_t)j * DD);
        // Sigma_{j,j+1} = -(G Snn + H Sna^T)
        mm<D>(T1, G, Snn);
        mmt<D>(T2, H, Sna);
#pragma unroll
        for (int e = 0; e < DD; ++e) Sjn.a[e] = -(T1.a[e] + T2.a[e]);
        // Sigma_{j,a} = -(G Sna + H Saa)
        mm<D>(T1, G, Sna);
        mm<D>(T2, H, Saa);
#pragma unroll
        for (int e = 0; e < DD; ++e) Sja.a[e] = -(T1.a[e] + T2.a[e]);
        // Sigma_jj = Dinv - Sigma_{j,j+1} G^T - Sigma_{j,a} H^T
        mmt<D>(T1, Sjn, G);
        mmt<D>(T2, Sja, H);
#pragma unroll
        for (int e = 0; e < DD; ++e) Snn.a[e] = Dinv.a[e] - T1.a[e] - T2.a[e];
        symmetrize<D>(Snn);
        mat_store<D>(cD + (size_t)j * DD, Snn);
        mat_store<D>(cO + (size_t)j * DD, Sjn);
        Sna = Sja;
    }
    // Sigma_{a,a+1} = Sigma_{a+1,a}^T  (for an empty interior Sna is still Sigma_{b,a})
    mat_transpose<D>(T1, Sna);
    mat_store<D>(cO + (size_t)a * DD, T1);
}

// Serial top level (n small): plain block Thomas + selected inverse + solve by one worker.
// x, cD, cO may be null (skipped).  Uses lv.G / lv.Dinv / lv.y as scratch; writes lv.ld[0].
template <int D, bool RHS>
GVI_HD void bt_serial_top(const BtLevel<D>& lv, double* __restrict__ x, double* __restrict__ cD,
                          double* __restrict__ cO) {
    constexpr int DD = D * D;
    const int n = lv.n;
    Mat<D> Dt, Dinv, Oj, G, T, Snn, Sjn;
    Vec<D> gcur, yv, tv, xn;
    LogDetAcc ld;
    bool ok = true;
    bt_load_diag<D>(Dt, lv, 0);
    if (RHS) bt_load_rhs<D>(gcur, lv, 0);
    for (int j = 0; j < n; ++j) {
        ok = spd_inverse<D>(Dinv, Dt, ld) && ok;
        mat_store<D>(lv.Dinv + (size_t)j * DD, Dinv);
        if (RHS) {
            mv<D>(yv, Dinv, gcur);
            vec_store<D>(lv.y + (size_t)j * D, yv);
        }
        if (j + 1 < n) {
            mat_load<D>(Oj, lv.O + (size_t)j * DD);
            mm<D>(G, Dinv, Oj);
            mat_store<D>(lv.G + (size_t)j * DD, G);
            mtm<D>(T, Oj, G);
            bt_load_diag<D>(Dt, lv, j + 1);
#pragma unroll
            for (int e = 0; e < DD; ++e) Dt.a[e] -= T.a[e];
            symmetrize<D>(Dt);
            if (RHS) {
                mtv<D>(tv, Oj, yv);
                bt_load_rhs<D>(gcur, lv, j + 1);
#pragma unroll
                for (int e = 0; e < D; ++e) gcur.a[e] -= tv.a[e];
            }
        }
    }
    lv.ld[0] = ld.value();
    if (!ok) *lv.notspd = 1;
    if (RHS && x != nullptr) {
        vec_load<D>(xn, lv.y + (size_t)(n - 1) * D);
        vec_store<D>(x + (size_t)(n - 1) * D, xn);
        for (int j = n - 2; j >= 0; --j) {
            mat_load<D>(G, lv.G + (size_t)j * DD);
            vec_load<D>(yv, lv.y + (size_t)j * D);
            mv<D>(tv, G, xn);
#pragma unroll
            for (int e = 0; e < D; ++e) xn.a[e] = yv.a[e] - tv.a[e];
            vec_store<D>(x + (size_t)j * D, xn);
        }
    }
    if (cD != nullptr) {
        mat_load<D>(Snn, lv.Dinv + (size_t)(n - 1) * DD);
        mat_store<D>(cD + (size_t)(n - 1) * DD, Snn);
        for (int j = n - 2; j >= 0; --j) {
            mat_load<D>(G, lv.G + (size_t)j * DD);
            mat_load<D>(Dinv, lv.Dinv + (size_t)j * DD);
            mm<D>(T, G, Snn);
#pragma unroll
            for (int e = 0; e < DD; ++e) Sjn.a[e] = -T.a[e];
            mmt<D>(T, Sjn, G);
#pragma unroll
            for (int e = 0; e < DD; ++e) Snn.a[e] = Dinv.a[e] - T.a[e];
            symmetrize<D>(Snn);
            mat_store<D>(cD + (size_t)j * DD, Snn);
            mat_store<D>(cO + (size_t)j * DD, Sjn);
        }
    }
}

}  // namespace gvib200
