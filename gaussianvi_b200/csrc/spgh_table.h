// Host-side sparse Gauss-Hermite table generator (see spgh_table.cpp).
#pragma once
#include <vector>

namespace gvib200 {
// nodes_rowmajor: [n][dim]; weights: [n]; throws std::invalid_argument on bad (dim, deg).
void generate_spgh_table(int dim, int deg, std::vector<double>& nodes_rowmajor, std::vector<double>& weights);
}  // namespace gvib200
