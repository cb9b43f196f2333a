// Host-side plan of the partitioned block-tridiagonal elimination: level sizes and workspace
// offsets (in doubles).  Shared by the CUDA driver (gvib200.cu) and the host test harness.
#pragma once
#include <cstddef>
#include <vector>

#include "bt_chain.h"

namespace gvib200 {

struct BtPlanLevel {
    int n = 0, L = 0, K = 0;  // L == 0: serial top level
    // record of this level
    size_t G = 0, H = 0, Dinv = 0, y = 0, ld = 0;
    // system of this level (levels >= 1 live in the workspace; level 0 is external)
    size_t Dn = 0, CL = 0, CR = 0, O = 0, g = 0, gl = 0, gr = 0;
    // results of this level (levels >= 1)
    size_t x = 0, cD = 0, cO = 0;
};

struct BtPlan {
    int D = 0;
    std::vector<BtPlanLevel> levels;
    size_t ws_doubles = 0;
    size_t ld_offset = 0, ld_count = 0;  // all partial log dets are contiguous
    bool two_node_top = false;           // multi-GPU: stop at the 2-node system (first, last)
};

// seg: segment length of the parallel levels; n_serial: size at or below which the remaining chain
// is solved serially.  With two_node_top the recursion continues until exactly 2 nodes remain and no
// serial level is appended (the caller owns the reduced 2-node system: rank boundary exchange).
inline BtPlan bt_make_plan(int n0, int D, int seg, int n_serial, bool two_node_top = false) {
    BtPlan p;
    p.D = D;
    p.two_node_top = two_node_top;
    const size_t DD = (size_t)D * D;
    size_t off = 0;
    auto take = [&](size_t cnt) {
        size_t o = off;
        off += (cnt + 1) & ~size_t(1);  // keep 16-byte alignment
        return o;
    };
    if (n_serial < 2) n_serial = 2;
    // first pass: level sizes
    int n = n0;
    while (true) {
        BtPlanLevel lv;
        lv.n = n;
        bool top = two_node_top ? (n <= 2) : (n <= n_serial);
        if (top) {
            lv.L = 0;
            lv.K = 0;
            p.levels.push_back(lv);
            break;
        }
        lv.L = seg;
        if (two_node_top && (n - 1) <= seg * 1) lv.L = n - 1;
        lv.K = (n - 1 + lv.L - 1) / lv.L;
        p.levels.push_back(lv);
        n = lv.K + 1;
    }
    // second pass: offsets
    p.ld_offset = 0;
    size_t ldc = 0;
    for (auto& lv : p.levels) ldc += (lv.L == 0) ? 1 : (size_t)lv.K;
    p.ld_count = ldc;
    p.ld_offset = take(ldc);
    size_t ldo = p.ld_offset;
    for (size_t l = 0; l < p.levels.size(); ++l) {
        auto& lv = p.levels[l];
        lv.ld = ldo;
        ldo += (lv.L == 0) ? 1 : (size_t)lv.K;
        lv.G = take((size_t)lv.n * DD);
        lv.H = take((size_t)lv.n * DD);
        lv.Dinv = take((size_t)lv.n * DD);
        lv.y = take((size_t)lv.n * D);
        if (l > 0) {
            const auto& prev = p.levels[l - 1];
            lv.Dn = take((size_t)lv.n * DD);
            lv.CL = take((size_t)prev.K * DD);
            lv.CR = take((size_t)prev.K * DD);
            lv.O = take((size_t)prev.K * DD);
            lv.g = take((size_t)lv.n * D);
            lv.gl = take((size_t)prev.K * D);
            lv.gr = take((size_t)prev.K * D);
            lv.x = take((size_t)lv.n * D);
            lv.cD = take((size_t)lv.n * DD);
            lv.cO = take((size_t)lv.n * DD);
        }
    }
    p.ws_doubles = off;
    return p;
}

// Bind level l of a plan to pointers.  D0/O0/g0: the level-0 system (external); ws: workspace base.
template <int D>
inline BtLevel<D> bt_bind_level(const BtPlan& p, size_t l, double* ws, const double* D0, const double* O0,
                                const double* g0, int* notspd) {
    BtLevel<D> b{};
    const auto& lv = p.levels[l];
    b.n = lv.n;
    b.L = lv.L;
    b.K = lv.K;
    if (l == 0) {
        b.Dn = D0;
        b.CL = nullptr;
        b.CR = nullptr;
        b.O = O0;
        b.g = g0;
        b.gl = nullptr;
        b.gr = nullptr;
    } else {
        b.Dn = ws + lv.Dn;
        b.CL = ws + lv.CL;
        b.CR = ws + lv.CR;
        b.O = ws + lv.O;
        b.g = (g0 != nullptr) ? ws + lv.g : nullptr;
        b.gl = (g0 != nullptr) ? ws + lv.gl : nullptr;
        b.gr = (g0 != nullptr) ? ws + lv.gr : nullptr;
    }
    b.G = ws + lv.G;
    b.H = ws + lv.H;
    b.Dinv = ws + lv.Dinv;
    b.y = ws + lv.y;
    b.ld = ws + lv.ld;
    if (l + 1 < p.levels.size()) {
        const auto& nx = p.levels[l + 1];
        b.rDn = ws + nx.Dn;
        b.rCL = ws + nx.CL;
        b.rCR = ws + nx.CR;
        b.rO = ws + nx.O;
        b.rg = ws + nx.g;
        b.rgl = ws + nx.gl;
        b.rgr = ws + nx.gr;
    }
    b.notspd = notspd;
    return b;
}

}  // namespace gvib200
