// TEST HARNESS (not part of libgvib200.so): runs the block-tridiagonal engine's
// __host__ __device__ arithmetic (gaussianvi_b200/csrc/bt_chain.h) on the CPU, one loop iteration
// per CUDA thread, so tests/test_bt_chain_host.py can check plan + math against the oracle without
// a GPU.  Also exposes the Jacobi sqrt / inverse-sqrt helper.
#include <cstring>
#include <vector>

#include "../../gaussianvi_b200/csrc/bt_plan.h"

using namespace gvib200;

template <int D>
static int run(int S, const double* D0, const double* O0, const double* rhs, double* x, double* cD, double* cO,
               double* logdet, int seg, int nserial) {
    BtPlan plan = bt_make_plan(S, D, seg, nserial, false);
    std::vector<double> ws(plan.ws_doubles + 16, 0.0);
    int notspd = 0;
    const size_t nl = plan.levels.size();
    for (size_t l = 0; l + 1 < nl; ++l) {
        BtLevel<D> lv = bt_bind_level<D>(plan, l, ws.data(), D0, O0, rhs, &notspd);
        for (int k = 0; k < lv.K; ++k) {
            if (rhs) bt_forward_segment<D, true>(lv, k);
            else bt_forward_segment<D, false>(lv, k);
        }
    }
    {
        BtLevel<D> top = bt_bind_level<D>(plan, nl - 1, ws.data(), D0, O0, rhs, &notspd);
        double* tx = (nl == 1) ? x : ws.data() + plan.levels[nl - 1].x;
        double* tD = (nl == 1) ? cD : ws.data() + plan.levels[nl - 1].cD;
        double* tO = (nl == 1) ? cO : ws.data() + plan.levels[nl - 1].cO;
        if (rhs) bt_serial_top<D, true>(top, tx, cD ? tD : nullptr, tO);
        else bt_serial_top<D, false>(top, nullptr, cD ? tD : nullptr, tO);
    }
    for (int l = (int)nl - 2; l >= 0; --l) {
        BtLevel<D> lv = bt_bind_level<D>(plan, l, ws.data(), D0, O0, rhs, &notspd);
        const auto& up = plan.levels[l + 1];
        double* lx = (l == 0) ? x : ws.data() + plan.levels[l].x;
        double* lD = (l == 0) ? cD : ws.data() + plan.levels[l].cD;
        double* lO = (l == 0) ? cO : ws.data() + plan.levels[l].cO;
        for (int k = 0; k < lv.K; ++k) {
            if (rhs) bt_backsolve_segment<D>(lv, k, ws.data() + up.x, lx);
            if (cD) bt_selinv_segment<D>(lv, k, ws.data() + up.cD, ws.data() + up.cO, lD, lO);
        }
    }
    double ld = 0.0;
    for (size_t i = 0; i < plan.ld_count; ++i) ld += ws[plan.ld_offset + i];
    if (logdet) *logdet = ld;
    return notspd ? -4 : 0;
}

extern "C" int emu_blocktri(int S, int d, const double* D0, const double* O0, const double* rhs, double* x,
                            double* cD, double* cO, double* logdet, int seg, int nserial) {
    switch (d) {
        case 1: return run<1>(S, D0, O0, rhs, x, cD, cO, logdet, seg, nserial);
        case 2: return run<2>(S, D0, O0, rhs, x, cD, cO, logdet, seg, nserial);
        case 3: return run<3>(S, D0, O0, rhs, x, cD, cO, logdet, seg, nserial);
        case 4: return run<4>(S, D0, O0, rhs, x, cD, cO, logdet, seg, nserial);
        case 6: return run<6>(S, D0, O0, rhs, x, cD, cO, logdet, seg, nserial);
        default: return -1;
    }
}

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
