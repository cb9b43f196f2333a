// Host-side plan of the tile-wise cyclic-reduction engine (bt_cr.h): tile size, tile count and the offsets (in
// doubles) of the records / separator system / top results inside one workspace allocation.  Shared by the CUDA
// driver (gvib200.cu) and the host test harness.
#pragma once
#include <cstddef>

#include "bt_cr.h"

namespace gvib200 {

struct CrPlan {
    int n = 0, D = 0, T = 0, K = 0;
    size_t recG = 0, recH = 0, recDinv = 0, recy = 0;
    size_t rDn = 0, rCL = 0, rCR = 0, rO = 0, rg = 0, rgl = 0, rgr = 0;
    size_t tD = 0, tO = 0, tx = 0, ld = 0;
    size_t ws_doubles = 0;
    size_t tile_smem_bytes = 0, top_smem_bytes = 0;
    int ld_count = 0;
};

template <int D>
inline int cr_max_top_nodes(size_t smem_bytes) {
    int lo = 1, hi = 1 << 16;
    while (lo < hi) {  // largest n with cr_top_doubles(n) * 8 <= smem_bytes
        const int mid = (lo + hi + 1) / 2;
        if (cr_top_doubles<D>(mid) * sizeof(double) <= smem_bytes) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
template <int D>
inline int cr_max_tile_links(size_t smem_bytes) {
    int lo = 1, hi = 1 << 16;
    while (lo < hi) {
        const int mid = (lo + hi + 1) / 2;
        if (cr_tile_doubles<D>(mid) * sizeof(double) <= smem_bytes) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// tiles_hint: how many tile CTAs the device runs at once (SM count); smem_bytes: dynamic shared memory per CTA.
// force_T > 0 fixes the tile size (tests); force_T < 0 forces the top-only path.  Returns false when the chain is too
// long for a two-level plan on this device.
// allow_mid: the separator chain may exceed what the top kernel holds (the caller runs a mid level over it).
template <int D>
inline bool cr_make_plan(CrPlan& p, int n, int tiles_hint, size_t smem_bytes, int force_T = 0, bool allow_mid = false) {
    constexpr size_t DD = (size_t)D * D;
    p = CrPlan();
    p.n = n;
    p.D = D;
    const int top_max = cr_max_top_nodes<D>(smem_bytes);
    const int t_max = cr_max_tile_links<D>(smem_bytes);
    if (force_T < 0 || (force_T == 0 && n <= top_max)) {
        if (n > top_max) return false;
        p.K = 0;
        p.T = 0;
    } else {
        int T;
        if (force_T > 0) {
            T = force_T;
        } else {
            int K = tiles_hint < top_max - 1 ? tiles_hint : top_max - 1;
            if (K < 1) K = 1;
            T = (n - 1 + K - 1) / K;
            if (T < 32) T = 32;
            if (T > t_max) T = t_max;
            if (allow_mid && (n - 1 + T - 1) / T + 1 > top_max) T = t_max;  // long chain: the largest tiles that fit
        }
        if (T < 2) T = 2;
        if (T > t_max) return false;
        p.T = T;
        p.K = (n - 1 + T - 1) / T;
        if (p.K + 1 > top_max && !allow_mid) return false;
    }
    size_t off = 0;
    auto take = [&](size_t cnt) {
        size_t o = off;
        off += (cnt + 1) & ~size_t(1);
        return o;
    };
    const size_t nrec = p.K > 0 ? (size_t)p.K * (p.T - 1) : 0;
    p.recG = take(cr_rec_capacity(nrec, (int)DD));
    p.recH = take(cr_rec_capacity(nrec, (int)DD));
    p.recDinv = take(cr_rec_capacity(nrec, (int)DD));
    p.recy = take(cr_rec_capacity(nrec, D));
    const size_t K1 = (size_t)p.K + 1;
    p.rDn = take(K1 * DD);
    p.rCL = take(K1 * DD);
    p.rCR = take(K1 * DD);
    p.rO = take(K1 * DD);
    p.rg = take(K1 * D);
    p.rgl = take(K1 * D);
    p.rgr = take(K1 * D);
    p.tD = take(K1 * DD);
    p.tO = take(K1 * DD);
    p.tx = take(K1 * D);
    p.ld = take(K1);
    p.ld_count = p.K + 1;
    p.ws_doubles = off;
    p.tile_smem_bytes = p.K > 0 ? cr_tile_doubles<D>(p.T) * sizeof(double) : 0;
    p.top_smem_bytes = (p.K + 1 > top_max) ? 0 : cr_top_doubles<D>(p.K > 0 ? p.K + 1 : n) * sizeof(double);
    return true;
}

template <int D>
inline CrArgs<D> cr_bind(const CrPlan& p, double* ws, const double* Dg, const double* Og, const double* g, double* x,
                         double* cD, double* cO, int* notspd) {
    CrArgs<D> a;
    a.n = p.n;
    a.T = p.T;
    a.K = p.K;
    a.Dg = Dg;
    a.Og = Og;
    a.g = g;
    a.rec.G = ws + p.recG;
    a.rec.H = ws + p.recH;
    a.rec.Dinv = ws + p.recDinv;
    a.rec.y = ws + p.recy;
    a.rDn = ws + p.rDn;
    a.rCL = ws + p.rCL;
    a.rCR = ws + p.rCR;
    a.rO = ws + p.rO;
    a.rg = ws + p.rg;
    a.rgl = ws + p.rgl;
    a.rgr = ws + p.rgr;
    a.tD = ws + p.tD;
    a.tO = ws + p.tO;
    a.tx = ws + p.tx;
    a.x = x;
    a.cD = cD;
    a.cO = cO;
    a.ld = ws + p.ld;
    a.ldout = nullptr;
    a.notspd = notspd;
    a.Dg2 = nullptr;
    a.Og2 = nullptr;
    a.alpha = 0.0;
    a.Dout = nullptr;
    a.Oout = nullptr;
    a.xbase = nullptr;
    a.xalpha = 0.0;
    a.xout = nullptr;
    a.alpha_node = nullptr;
    a.ldnode = nullptr;
    a.ld_stride = 1;
    a.ld_max = (long long)p.n - 1;
    return a;
}

}  // namespace gvib200
