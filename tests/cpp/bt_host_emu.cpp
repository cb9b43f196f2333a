// TEST HARNESS (not part of libgvib200.so): runs the block-tridiagonal engine's
// __host__ __device__ arithmetic (gaussianvi_b200/csrc/bt_cr.h) on the CPU, one loop iteration
// per CUDA thread, so tests/test_capi_host.py can check plan + math against the oracle without
// a GPU.  Also exposes the Jacobi sqrt / inverse-sqrt helper.
#include <cstring>
#include <vector>

#include "../../gaussianvi_b200/csrc/bt_cr_plan.h"

using namespace gvib200;

template <int N>
static void sq(const double* Sigma, double* S, double* R) {
    Mat<N> A, s, r;
    mat_load<N>(A, Sigma);
    sqrt_and_invsqrt<N>(s, r, A);
    mat_store<N>(S, s);
    mat_store<N>(R, r);
}

extern "C" int emu_sqrt_invsqrt(int n, const double* Sigma, double* S, double* R) {
    switch (n) {
        case 1: sq<1>(Sigma, S, R); return 0;
        case 2: sq<2>(Sigma, S, R); return 0;
        case 4: sq<4>(Sigma, S, R); return 0;
        case 8: sq<8>(Sigma, S, R); return 0;
        case 12: sq<12>(Sigma, S, R); return 0;
        default: return -1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// chain engine (bt_cr.h): the three kernels of kernels.cuh replayed sequentially, "threads" of a CTA in
// rounds of NT with both phases of a round separated exactly where the kernels place __syncthreads().
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr int NT = 512;

template <int D, bool RHS>
bool emu_forward_levels(const CrView<D>& v, const CrRec<D>& rec, size_t rec_base, const CrGeom& gm, double& ldsum) {
    bool ok = true;
    const int per_round = NT / D;
    std::vector<LogDetAcc> lds(NT);
    for (int l = 0; l < gm.levels; ++l) {
        const int cnt = cr_count(gm.T, l);
        for (int base = 0; base < cnt; base += per_round) {
            std::vector<CrElim<D>> ctx(NT);
            for (int tid = 0; tid < per_round * D; ++tid)
                if (base + tid / D < cnt)
                    ok = cr_fwd_A<D, RHS>(v, rec, rec_base, gm, l, base + tid / D, tid % D, ctx[tid], lds[tid]) && ok;
            for (int tid = 0; tid < per_round * D; ++tid)
                if (base + tid / D < cnt) cr_fwd_B<D, RHS>(v, ctx[tid], tid % D);
        }
    }
    for (auto& l : lds) ldsum += l.value();
    return ok;
}

template <int D, bool RHS, bool SELINV>
void emu_backward_levels(const CrView<D>& v, const CrRec<D>& rec, size_t rec_base, const CrGeom& gm) {
    const int per_round = NT / D;
    for (int l = gm.levels - 1; l >= 0; --l) {
        const int cnt = cr_count(gm.T, l);
        for (int base = 0; base < cnt; base += per_round) {
            if (SELINV) {
                std::vector<CrSel<D>> ctx(NT);
                for (int tid = 0; tid < per_round * D; ++tid)
                    if (base + tid / D < cnt) cr_bwd_selinv_compute<D>(v, rec, rec_base, gm, l, base + tid / D, tid % D, ctx[tid]);
                for (int tid = 0; tid < per_round * D; ++tid)
                    if (base + tid / D < cnt) cr_bwd_selinv_store<D>(v, ctx[tid], tid % D);
            }
            if (RHS)
                for (int tid = 0; tid < per_round * D; ++tid)
                    if (base + tid / D < cnt) cr_bwd_solve<D>(v, rec, rec_base, gm, l, base + tid / D, tid % D);
        }
    }
}

template <int D, bool RHS>
void emu_tile_forward(const CrArgs<D>& a, bool& ok);
template <int D, bool RHS, bool SELINV>
void emu_tile_backward(const CrArgs<D>& a);

// k_cr_top replayed: a.K == 0 works on the chain itself, else on the separator system of `a`
template <int D, bool RHS, bool SELINV>
void emu_top(const CrArgs<D>& a, bool& ok) {
    const int nt = (a.K == 0) ? a.n : a.K + 1;
    CrGeom gm;
    cr_make_geom(gm, nt - 1);
    std::vector<double> sm(cr_top_doubles<D>(nt), 0.0);
    CrView<D> v = cr_make_view<D>(sm.data(), nt);
    const size_t nrec = nt > 2 ? (size_t)(nt - 2) : 0;
    CrRec<D> rec;
    rec.G = v.g + (size_t)D * v.NS;
    rec.H = rec.G + cr_rec_capacity(nrec, D * D);
    rec.Dinv = rec.H + cr_rec_capacity(nrec, D * D);
    rec.y = rec.Dinv + cr_rec_capacity(nrec, D * D);
    cr_top_load<D, RHS>(a, v, gm, 0, 1);
    double ld = 0.0;
    ok = emu_forward_levels<D, RHS>(v, rec, 0, gm, ld) && ok;
    LogDetAcc l2;
    ok = cr_top2<D, RHS, SELINV>(v, gm.T, l2) && ok;
    ld += l2.value();
    emu_backward_levels<D, RHS, SELINV>(v, rec, 0, gm);
    if (a.K == 0) cr_store_results<D, RHS, SELINV>(v, gm, nt, nt - 1, a.x, a.cD, a.cO, 0, 0, 1);
    else cr_store_results<D, RHS, SELINV>(v, gm, nt, nt - 1, a.tx, a.tD, a.tO, 0, 0, 1);
    a.ld[a.K] = ld;
}

template <int D, bool RHS, bool SELINV>
int emu_cr_pass(int n, const double* D0, const double* O0, const double* rhs, double* x, double* cD, double* cO,
                double* logdet, int force_T, size_t smem_bytes) {
    constexpr int DD = D * D;
    CrPlan plan, pmid;
    bool three = false;
    if (!cr_make_plan<D>(plan, n, 148, smem_bytes, force_T)) {  // same fallback as gvib200_problem_finalize
        if (force_T != 0) return -1;
        if (!cr_make_plan<D>(plan, n, 148, smem_bytes, 0, true) || !cr_make_plan<D>(pmid, plan.K + 1, 148, smem_bytes)) return -1;
        three = true;
    }
    std::vector<double> ws(plan.ws_doubles + 16, 0.0);
    int notspd = 0;
    CrArgs<D> a = cr_bind<D>(plan, ws.data(), D0, O0, rhs, x, cD, cO, &notspd);
    bool ok = true;
    double ld = 0.0;
    emu_tile_forward<D, RHS>(a, ok);
    if (three) {
        std::vector<double> wsm(pmid.ws_doubles + 16, 0.0), D1((size_t)(plan.K + 1) * DD), O1((size_t)(plan.K + 1) * DD),
            g1((size_t)(plan.K + 1) * D);
        cr_sum_level<D, RHS>(a, D1.data(), O1.data(), g1.data(), 0, 1);
        CrArgs<D> mid = cr_bind<D>(pmid, wsm.data(), D1.data(), O1.data(), RHS ? g1.data() : nullptr, a.tx, a.tD, a.tO, &notspd);
        emu_tile_forward<D, RHS>(mid, ok);
        emu_top<D, RHS, SELINV>(mid, ok);
        emu_tile_backward<D, RHS, SELINV>(mid);
        for (int i = 0; i < pmid.ld_count; ++i) ld += mid.ld[i];
        for (int i = 0; i < plan.K; ++i) ld += a.ld[i];
    } else {
        emu_top<D, RHS, SELINV>(a, ok);
        for (int i = 0; i < plan.ld_count; ++i) ld += a.ld[i];
    }
    emu_tile_backward<D, RHS, SELINV>(a);
    if (logdet) *logdet = ld;
    return (ok && !notspd) ? 0 : -4;
}

template <int D>
int emu_cr(int n, const double* D0, const double* O0, const double* rhs, double* x, double* cD, double* cO, double* logdet,
           int force_T, size_t smem_bytes) {
    int rc = 0;
    if (rhs) rc = emu_cr_pass<D, true, false>(n, D0, O0, rhs, x, nullptr, nullptr, logdet, force_T, smem_bytes);
    if (rc == 0 && cD) rc = emu_cr_pass<D, false, true>(n, D0, O0, nullptr, nullptr, cD, cO, logdet, force_T, smem_bytes);
    return rc;
}
}  // namespace

// force_T: > 0 tile size, 0 automatic, < 0 top-only.  smem_bytes: the per-CTA shared-memory budget the plan assumes.
extern "C" int emu_cr_blocktri(int n, int d, const double* D0, const double* O0, const double* rhs, double* x, double* cD,
                               double* cO, double* logdet, int force_T, size_t smem_bytes) {
    switch (d) {
        case 1: return emu_cr<1>(n, D0, O0, rhs, x, cD, cO, logdet, force_T, smem_bytes);
        case 2: return emu_cr<2>(n, D0, O0, rhs, x, cD, cO, logdet, force_T, smem_bytes);
        case 3: return emu_cr<3>(n, D0, O0, rhs, x, cD, cO, logdet, force_T, smem_bytes);
        case 4: return emu_cr<4>(n, D0, O0, rhs, x, cD, cO, logdet, force_T, smem_bytes);
        case 6: return emu_cr<6>(n, D0, O0, rhs, x, cD, cO, logdet, force_T, smem_bytes);
        default: return -1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// multi-GPU pass replayed in one process: P "ranks", each with its own segment, workspace and share of the shared
// end blocks; the all-gather is a concatenation.  Dloc / Oloc / gloc: per rank [m+1] / [m] / [m+1] local blocks.
// ---------------------------------------------------------------------------------------------------------------
namespace {

template <int D, bool RHS>
void emu_tile_forward(const CrArgs<D>& a, bool& ok) {
    for (int tile = 0; tile < a.K; ++tile) {
        const int n0 = tile * a.T;
        const int Tk = (a.T < a.n - 1 - n0) ? a.T : a.n - 1 - n0;
        CrGeom gm;
        cr_make_geom(gm, Tk);
        std::vector<double> sm(cr_tile_doubles<D>(a.T), 0.0);
        CrView<D> v = cr_make_view<D>(sm.data(), a.T + 1);
        cr_tile_load<D, RHS>(a, v, gm, n0, 0, 1);
        double ld = 0.0;
        ok = emu_forward_levels<D, RHS>(v, a.rec, (size_t)tile * (a.T - 1), gm, ld) && ok;
        cr_tile_store_reduced<D, RHS>(a, v, gm, tile, n0, 0, 1);
        a.ld[tile] = ld;
    }
}

template <int D, bool RHS, bool SELINV>
void emu_tile_backward(const CrArgs<D>& a) {
    for (int tile = 0; tile < a.K; ++tile) {
        const int n0 = tile * a.T;
        const int Tk = (a.T < a.n - 1 - n0) ? a.T : a.n - 1 - n0;
        CrGeom gm;
        cr_make_geom(gm, Tk);
        std::vector<double> sm(cr_tile_doubles<D>(a.T), 0.0);
        CrView<D> v = cr_make_view<D>(sm.data(), a.T + 1);
        cr_tile_seed<D, RHS, SELINV>(a, v, tile, 0, 1);
        emu_backward_levels<D, RHS, SELINV>(v, a.rec, (size_t)tile * (a.T - 1), gm);
        const bool last = (tile == a.K - 1);
        cr_store_results<D, RHS, SELINV>(v, gm, Tk + (last ? 1 : 0), Tk, a.x, a.cD, a.cO, (size_t)n0, 0, 1);
    }
}

template <int D, bool RHS, bool SELINV>
int emu_dist_pass(int P, int m, const double* Dloc, const double* Oloc, const double* gloc, double* xloc, double* cDloc,
                  double* cOloc, double* logdet, int T) {
    constexpr int DD = D * D;
    constexpr int NB = cr_boundary_doubles<D>();
    const int n = m + 1;
    struct Rank {
        CrPlan p1, p2;
        std::vector<double> ws1, ws2, D1, O1, g1;
        CrArgs<D> a1, a2;
    };
    std::vector<Rank> R((size_t)P);
    std::vector<double> recs((size_t)P * NB, 0.0);
    int notspd = 0;
    bool ok = true;
    for (int r = 0; r < P; ++r) {
        Rank& k = R[(size_t)r];
        if (!cr_make_plan<D>(k.p1, n, 148, 220 * 1024, T)) return -1;
        if (!cr_make_plan<D>(k.p2, k.p1.K + 1, 1, 220 * 1024, k.p1.K < 2 ? 2 : k.p1.K)) return -1;
        k.ws1.assign(k.p1.ws_doubles + 16, 0.0);
        k.ws2.assign(k.p2.ws_doubles + 16, 0.0);
        k.D1.assign((size_t)(k.p1.K + 1) * DD, 0.0);
        k.O1.assign((size_t)(k.p1.K + 1) * DD, 0.0);
        k.g1.assign((size_t)(k.p1.K + 1) * D, 0.0);
        k.a1 = cr_bind<D>(k.p1, k.ws1.data(), Dloc + (size_t)r * n * DD, Oloc + (size_t)r * m * DD,
                          RHS ? gloc + (size_t)r * n * D : nullptr, RHS ? xloc + (size_t)r * n * D : nullptr,
                          SELINV ? cDloc + (size_t)r * n * DD : nullptr, SELINV ? cOloc + (size_t)r * m * DD : nullptr, &notspd);
        // the mid tile works on the summed separator chain and writes its results into the seeds of the real tiles
        k.a2 = cr_bind<D>(k.p2, k.ws2.data(), k.D1.data(), k.O1.data(), RHS ? k.g1.data() : nullptr, k.a1.tx, k.a1.tD, k.a1.tO,
                          &notspd);
        emu_tile_forward<D, RHS>(k.a1, ok);
        cr_sum_level<D, RHS>(k.a1, k.D1.data(), k.O1.data(), k.g1.data(), 0, 1);
        emu_tile_forward<D, RHS>(k.a2, ok);
        cr_pack_boundary<D, RHS>(k.a2, recs.data() + (size_t)r * NB, 0, 1);  // + all-gather
    }
    // every rank: the chain of rank boundaries, solved redundantly (here once)
    std::vector<double> Dt((size_t)(P + 1) * DD), Ot((size_t)(P + 1) * DD), gt((size_t)(P + 1) * D), xt((size_t)(P + 1) * D),
        cDt((size_t)(P + 1) * DD), cOt((size_t)(P + 1) * DD);
    cr_build_global<D>(P, recs.data(), Dt.data(), Ot.data(), gt.data(), 0, 1);
    double ld_top = 0.0;
    {
        CrPlan pt;
        if (!cr_make_plan<D>(pt, P + 1, 1, 220 * 1024, -1)) return -1;
        std::vector<double> wst(pt.ws_doubles + 16, 0.0);
        CrArgs<D> at = cr_bind<D>(pt, wst.data(), Dt.data(), Ot.data(), RHS ? gt.data() : nullptr, xt.data(), cDt.data(), cOt.data(),
                                  &notspd);
        const int nt = P + 1;
        CrGeom gm;
        cr_make_geom(gm, nt - 1);
        std::vector<double> sm(cr_top_doubles<D>(nt), 0.0);
        CrView<D> v = cr_make_view<D>(sm.data(), nt);
        const size_t nrec = nt > 2 ? (size_t)(nt - 2) : 0;
        CrRec<D> rec;
        rec.G = v.g + (size_t)D * v.NS;
        rec.H = rec.G + cr_rec_capacity(nrec, D * D);
        rec.Dinv = rec.H + cr_rec_capacity(nrec, D * D);
        rec.y = rec.Dinv + cr_rec_capacity(nrec, D * D);
        cr_top_load<D, RHS>(at, v, gm, 0, 1);
        ok = emu_forward_levels<D, RHS>(v, rec, 0, gm, ld_top) && ok;
        LogDetAcc l2;
        ok = cr_top2<D, RHS, SELINV>(v, gm.T, l2) && ok;
        ld_top += l2.value();
        emu_backward_levels<D, RHS, SELINV>(v, rec, 0, gm);
        cr_store_results<D, RHS, SELINV>(v, gm, nt, nt - 1, at.x, at.cD, at.cO, 0, 0, 1);
    }
    double ld = ld_top;
    for (int r = 0; r < P; ++r) {
        Rank& k = R[(size_t)r];
        cr_seed_mid<D, RHS, SELINV>(k.a2, r, xt.data(), cDt.data(), cOt.data(), 0, 1);
        emu_tile_backward<D, RHS, SELINV>(k.a2);
        emu_tile_backward<D, RHS, SELINV>(k.a1);
        for (int i = 0; i < k.p1.K; ++i) ld += k.a1.ld[i];
        for (int i = 0; i < k.p2.K; ++i) ld += k.a2.ld[i];
    }
    if (logdet) *logdet = ld;
    return (ok && !notspd) ? 0 : -4;
}
}  // namespace

extern "C" int emu_cr_distributed(int P, int m, int d, const double* Dloc, const double* Oloc, const double* gloc, double* xloc,
                                  double* cDloc, double* cOloc, double* logdet, int T) {
    int rc = -1;
#define DIST_CASE(D_)                                                                                                   \
    case D_:                                                                                                            \
        rc = emu_dist_pass<D_, true, false>(P, m, Dloc, Oloc, gloc, xloc, nullptr, nullptr, logdet, T);                 \
        if (rc == 0) rc = emu_dist_pass<D_, false, true>(P, m, Dloc, Oloc, nullptr, nullptr, cDloc, cOloc, logdet, T);  \
        break;
    switch (d) {
        DIST_CASE(1) DIST_CASE(2) DIST_CASE(4) DIST_CASE(6)
        default: break;
    }
#undef DIST_CASE
    return rc;
}
