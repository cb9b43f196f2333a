// Sparse Gauss-Hermite (Smolyak) node/weight table generator -- host C++.
//
// Produces the table the reference loads from quadrature/SparseGHQuadratureWeights_cereal.bin
// (quadrature/SparseGaussHermite.h:138-166), which the reference itself generates by calling
// a MATLAB-compiled nwspgr('GQN', dim, k, 1) (quadrature/generateSpGHWeights.h:23-84).  This
// is a from-scratch C++ implementation of the published Heiss-Winschel construction
// (quadrature/GH/SparseGH/nwspgr.m:32-134 is the reference's statement of it); the 1-D rules
// come from gqn_table.inc.  Row order = lexicographic, duplicates merged only when bit-equal,
// weights renormalised to sum 1 -- so tables are interchangeable with the reference's.
#include "spgh_table.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <vector>

namespace gvib200 {
namespace {
#include "gqn_table.inc"

struct Rule1D {
    const double* n;
    const double* w;
    int size;
};

Rule1D rule1d(int level) {
    return Rule1D{GQN_NODES + GQN_OFFSET[level], GQN_WEIGHTS + GQN_OFFSET[level],
                  GQN_OFFSET[level + 1] - GQN_OFFSET[level]};
}

double binom(int n, int k) {
    if (k < 0 || k > n) return 0.0;
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return std::round(r);
}

// All multi-indices in N^d with entries >= 1 and sum == norm, enumerated in the order that
// fixes the sequence in which equal nodes' weights are added (same order as nwspgr.m:147-169,
// i.e. reverse-lexicographic descent starting from (norm-d+1, 1, ..., 1)).
void enumerate_levels(int d, int norm, std::vector<std::vector<int>>& out) {
    std::vector<int> seq(d, 0);
    const int a = norm - d;
    seq[0] = a;
    out.push_back(seq);
    int c = 0;
    while (seq[d - 1] < a) {
        if (c == d - 1) {
            for (int i = c - 1; i >= 0; --i) {
                c = i;
                if (seq[i] != 0) break;
            }
        }
        seq[c] -= 1;
        c += 1;
        int partial = 0;
        for (int i = 0; i < c; ++i) partial += seq[i];
        seq[c] = a - partial;
        for (int i = c + 1; i < d; ++i) seq[i] = 0;
        out.push_back(seq);
    }
    for (auto& s : out)
        for (auto& v : s) v += 1;
}

struct Grid {
    int dim = 0;
    std::vector<double> nodes;  // row-major [n][dim]
    std::vector<double> w;
    int size() const { return static_cast<int>(w.size()); }
};

void stable_sort_rows(Grid& g) {
    const int n = g.size(), dim = g.dim;
    std::vector<int> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    const double* p = g.nodes.data();
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
        const double* ra = p + static_cast<size_t>(a) * dim;
        const double* rb = p + static_cast<size_t>(b) * dim;
        for (int j = 0; j < dim; ++j) {
            if (ra[j] < rb[j]) return true;
            if (ra[j] > rb[j]) return false;
        }
        return false;
    });
    std::vector<double> nn(g.nodes.size()), ww(n);
    for (int i = 0; i < n; ++i) {
        std::copy(p + static_cast<size_t>(idx[i]) * dim, p + static_cast<size_t>(idx[i] + 1) * dim,
                  nn.begin() + static_cast<size_t>(i) * dim);
        ww[i] = g.w[idx[i]];
    }
    g.nodes.swap(nn);
    g.w.swap(ww);
}

void merge_equal_rows(Grid& g) {
    const int n = g.size(), dim = g.dim;
    if (n == 0) return;
    int last = 0;
    for (int j = 1; j < n; ++j) {
        bool same = true;
        for (int c = 0; c < dim; ++c)
            if (g.nodes[static_cast<size_t>(j) * dim + c] != g.nodes[static_cast<size_t>(last) * dim + c]) {
                same = false;
                break;
            }
        if (same) {
            g.w[last] += g.w[j];
        } else {
            ++last;
            if (last != j) {
                std::copy(g.nodes.begin() + static_cast<size_t>(j) * dim,
                          g.nodes.begin() + static_cast<size_t>(j + 1) * dim,
                          g.nodes.begin() + static_cast<size_t>(last) * dim);
                g.w[last] = g.w[j];
            }
        }
    }
    g.w.resize(last + 1);
    g.nodes.resize(static_cast<size_t>(last + 1) * dim);
}

// tensor product of 1-D rules, first dimension slowest
void append_product(Grid& g, const std::vector<int>& levels, double scale) {
    const int dim = g.dim;
    std::vector<Rule1D> r(dim);
    size_t total = 1;
    for (int j = 0; j < dim; ++j) {
        r[j] = rule1d(levels[j]);
        total *= r[j].size;
    }
    std::vector<int> ctr(dim, 0);
    for (size_t t = 0; t < total; ++t) {
        // weight = kron(w_1, ..., w_dim): multiply left to right like repeated kron()
        double w = r[0].w[ctr[0]];
        for (int j = 1; j < dim; ++j) w = w * r[j].w[ctr[j]];
        for (int j = 0; j < dim; ++j) g.nodes.push_back(r[j].n[ctr[j]]);
        g.w.push_back(scale * w);
        for (int j = dim - 1; j >= 0; --j) {
            if (++ctr[j] < r[j].size) break;
            ctr[j] = 0;
        }
    }
}
}  // namespace

void generate_spgh_table(int dim, int k, std::vector<double>& nodes_rowmajor, std::vector<double>& weights) {
    if (dim < 1 || k < 1 || k > GQN_MAX_LEVEL) throw std::invalid_argument("spgh table: need dim >= 1, 1 <= deg <= 25");
    Grid g;
    g.dim = dim;
    const int minq = std::max(0, k - dim), maxq = k - 1;
    for (int q = minq; q <= maxq; ++q) {
        const double bq = (((maxq - q) % 2) ? -1.0 : 1.0) * binom(dim - 1, dim + q - k);
        std::vector<std::vector<int>> seqs;
        enumerate_levels(dim, dim + q, seqs);
        for (const auto& s : seqs) append_product(g, s, bq);
        stable_sort_rows(g);
        merge_equal_rows(g);
    }
    // mirror the positive orthant into the others, one coordinate at a time
    const double m = rule1d(1).n[0];
    for (int j = 0; j < dim; ++j) {
        const int nr = g.size();
        for (int r = 0; r < nr; ++r) {
            if (g.nodes[static_cast<size_t>(r) * dim + j] != m) {
                for (int c = 0; c < dim; ++c) {
                    double v = g.nodes[static_cast<size_t>(r) * dim + c];
                    g.nodes.push_back(c == j ? 2 * m - v : v);
                }
                g.w.push_back(g.w[r]);
            }
        }
    }
    stable_sort_rows(g);
    double sum = 0.0;
    for (double v : g.w) sum += v;  // left-to-right, like sum(weights)
    for (double& v : g.w) v /= sum;
    nodes_rowmajor.swap(g.nodes);
    weights.swap(g.w);
}

}  // namespace gvib200
