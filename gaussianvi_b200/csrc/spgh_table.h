// Host-side sparse Gauss-Hermite table generator (see spgh_table.cpp).
#pragma once
#include <vector>

namespace gvib200 {
// nodes_rowmajor: [n][dim]; weights: [n]; throws std::invalid_argument on bad (dim, deg).
void generate_spgh_table(int dim, int deg, std::vector<double>& nodes_rowmajor, std::vector<double>& weights);

// one (dim, deg) rule of a table file (spgh_table_io.cpp: the reference's cereal binary format)
struct SpghTableEntry {
    int dim = 0, deg = 0;
    std::vector<double> nodes_rowmajor;  // [n][dim]
    std::vector<double> weights;         // [n]
};
void write_spgh_table_file(const char* path, const std::vector<SpghTableEntry>& entries);  // throws on I/O errors
void read_spgh_table_file(const char* path, std::vector<SpghTableEntry>& entries);         // throws on malformed files
}  // namespace gvib200
