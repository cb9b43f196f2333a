// Sparse-GH table file, wire compatible with the reference's quadrature/SparseGHQuadratureWeights_cereal.bin.
//
// The reference writes `cereal::BinaryOutputArchive(ofs)(map)` with
//   map : std::unordered_map<std::tuple<double,double>, std::tuple<Eigen::MatrixXd, Eigen::VectorXd>>
// (quadrature/saveSparseGHWeightMap.h:43-51, quadrature/SparseGHQuadratureWeights.h:14-16).  cereal's portable binary
// framing of that type is (little endian, no padding, no names):
//   uint64  number of entries                    cereal/types/concepts/pair_associative_container.hpp (size tag, size_type)
//   per entry, in the map's iteration order:
//     double dim, double deg                     key tuple, elements in order (cereal/types/tuple.hpp)
//     int32 rows, int32 cols                     helpers/SerializeEigenMaps.h:195-199
//     rows*cols doubles, ROW-major element order helpers/SerializeEigenMaps.h:205-209 (loops i over rows, j over cols)
//     int32 size, size doubles                   helpers/SerializeEigenMaps.h:212-224
// A reader must not depend on the entry order (an unordered_map's iteration order); the writer here emits the keys in
// the order the caller gives them, which every cereal reader accepts.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "spgh_table.h"

namespace gvib200 {

namespace {
struct File {
    FILE* f = nullptr;
    explicit File(const char* path, const char* mode) : f(std::fopen(path, mode)) {}
    ~File() {
        if (f) std::fclose(f);
    }
};
template <class T>
void put(FILE* f, const T& v) {
    if (std::fwrite(&v, sizeof(T), 1, f) != 1) throw std::runtime_error("table file: write failed");
}
template <class T>
void get(FILE* f, T& v) {
    if (std::fread(&v, sizeof(T), 1, f) != 1) throw std::runtime_error("table file: truncated");
}
}  // namespace

void write_spgh_table_file(const char* path, const std::vector<SpghTableEntry>& entries) {
    File file(path, "wb");
    if (!file.f) throw std::runtime_error(std::string("table file: cannot open for writing: ") + path);
    put<uint64_t>(file.f, (uint64_t)entries.size());
    for (const auto& e : entries) {
        const int32_t rows = (int32_t)e.weights.size(), cols = (int32_t)e.dim;
        if ((size_t)rows * cols != e.nodes_rowmajor.size()) throw std::invalid_argument("table file: node array has the wrong size");
        put<double>(file.f, (double)e.dim);
        put<double>(file.f, (double)e.deg);
        put<int32_t>(file.f, rows);
        put<int32_t>(file.f, cols);
        if (!e.nodes_rowmajor.empty() &&
            std::fwrite(e.nodes_rowmajor.data(), sizeof(double), e.nodes_rowmajor.size(), file.f) != e.nodes_rowmajor.size())
            throw std::runtime_error("table file: write failed");
        put<int32_t>(file.f, rows);
        if (rows > 0 && std::fwrite(e.weights.data(), sizeof(double), (size_t)rows, file.f) != (size_t)rows)
            throw std::runtime_error("table file: write failed");
    }
}

void read_spgh_table_file(const char* path, std::vector<SpghTableEntry>& entries) {
    File file(path, "rb");
    if (!file.f) throw std::runtime_error(std::string("table file: cannot open: ") + path);
    uint64_t count = 0;
    get(file.f, count);
    if (count > (1u << 20)) throw std::runtime_error("table file: implausible entry count (not a cereal table file?)");
    entries.clear();
    entries.reserve((size_t)count);
    for (uint64_t k = 0; k < count; ++k) {
        double dim = 0, deg = 0;
        int32_t rows = 0, cols = 0, size = 0;
        get(file.f, dim);
        get(file.f, deg);
        get(file.f, rows);
        get(file.f, cols);
        if (rows < 0 || cols < 0 || dim != (double)(int)dim || deg != (double)(int)deg || (rows > 0 && cols != (int)dim) ||
            (uint64_t)rows * (uint64_t)cols > (1ull << 32))
            throw std::runtime_error("table file: malformed entry header");
        SpghTableEntry e;
        e.dim = (int)dim;
        e.deg = (int)deg;
        e.nodes_rowmajor.resize((size_t)rows * cols);
        if (!e.nodes_rowmajor.empty() &&
            std::fread(e.nodes_rowmajor.data(), sizeof(double), e.nodes_rowmajor.size(), file.f) != e.nodes_rowmajor.size())
            throw std::runtime_error("table file: truncated");
        get(file.f, size);
        if (size != rows) throw std::runtime_error("table file: weight count differs from the node count");
        e.weights.resize((size_t)size);
        if (size > 0 && std::fread(e.weights.data(), sizeof(double), (size_t)size, file.f) != (size_t)size)
            throw std::runtime_error("table file: truncated");
        entries.push_back(std::move(e));
    }
    if (std::fgetc(file.f) != EOF) throw std::runtime_error("table file: trailing bytes");
}

}  // namespace gvib200
