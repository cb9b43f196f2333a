// Fixture generator / cross-checker for the sparse-GH table wire format (tests/test_table_file.py).
// Built ONLY where the reference tree is present (it needs the cereal headers vendored there):
//   g++ -std=c++17 -I/root/reference/include/cereal/include -I<repo>/gaussianvi_b200/csrc \
//       cereal_table_tool.cpp <repo>/gaussianvi_b200/csrc/spgh_table.cpp -o cereal_table_tool
// The map type and the two serialize() functions restate helpers/SerializeEigenMaps.h:28-44,195-224 and
// quadrature/SparseGHQuadratureWeights.h:14-16 with stand-ins for Eigen::MatrixXd / VectorXd (Eigen is not in the
// image); the framing itself is produced / consumed by the real cereal library.
//   write <out.bin>  : archive(map) of the rules (1,3), (2,2), (3,2) -> the committed golden file
//   read  <in.bin>   : load a file with cereal and print every entry as text (sorted by key)
#include <cereal/archives/binary.hpp>
#include <cereal/types/tuple.hpp>
#include <cereal/types/unordered_map.hpp>

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "spgh_table.h"

struct MatrixXd {  // row/col sizes + column-major storage, element access like Eigen's
    int r = 0, c = 0;
    std::vector<double> a;
    int rows() const { return r; }
    int cols() const { return c; }
    void resize(int rr, int cc) { r = rr; c = cc; a.assign((size_t)rr * cc, 0.0); }
    double& operator()(int i, int j) { return a[(size_t)i + (size_t)j * r]; }
};
struct VectorXd {
    std::vector<double> a;
    int size() const { return (int)a.size(); }
    void resize(int n) { a.assign((size_t)n, 0.0); }
    double& operator()(int i) { return a[(size_t)i]; }
};

namespace std {
template <>
struct hash<std::tuple<double, double>> {
    size_t operator()(const std::tuple<double, double>& key) const {
        size_t hash1 = std::hash<double>{}(std::get<0>(key));
        size_t hash2 = std::hash<double>{}(std::get<1>(key));
        return hash1 ^ (hash2 << 1);
    }
};
}  // namespace std

namespace cereal {
template <class Archive>
void serialize(Archive& archive, MatrixXd& matrix) {
    int rows = matrix.rows();
    int cols = matrix.cols();
    archive(rows, cols);
    if (Archive::is_loading::value) matrix.resize(rows, cols);
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) archive(matrix(i, j));
}
template <class Archive>
void serialize(Archive& archive, VectorXd& vector) {
    int size = vector.size();
    archive(size);
    if (Archive::is_loading::value) vector.resize(size);
    for (int i = 0; i < size; ++i) archive(vector(i));
}
}  // namespace cereal

using Map = std::unordered_map<std::tuple<double, double>, std::tuple<MatrixXd, VectorXd>>;

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    const std::string mode = argv[1];
    if (mode == "write") {
        Map map;
        const int keys[3][2] = {{1, 3}, {2, 2}, {3, 2}};
        for (auto& k : keys) {
            std::vector<double> nodes, w;
            gvib200::generate_spgh_table(k[0], k[1], nodes, w);
            MatrixXd M;
            VectorXd V;
            M.resize((int)w.size(), k[0]);
            V.resize((int)w.size());
            for (int i = 0; i < (int)w.size(); ++i) {
                V(i) = w[i];
                for (int j = 0; j < k[0]; ++j) M(i, j) = nodes[(size_t)i * k[0] + j];
            }
            map[std::make_tuple((double)k[0], (double)k[1])] = std::make_tuple(M, V);
        }
        std::ofstream ofs(argv[2], std::ios::binary);
        cereal::BinaryOutputArchive archive(ofs);
        archive(map);
        return 0;
    }
    if (mode == "read") {
        Map map;
        std::ifstream ifs(argv[2], std::ios::binary);
        if (!ifs) return 3;
        cereal::BinaryInputArchive archive(ifs);
        archive(map);
        std::vector<std::tuple<double, double>> keys;
        for (auto& kv : map) keys.push_back(kv.first);
        std::sort(keys.begin(), keys.end());
        for (auto& k : keys) {
            auto& v = map[k];
            MatrixXd& M = std::get<0>(v);
            VectorXd& V = std::get<1>(v);
            std::printf("%d %d %d %d\n", (int)std::get<0>(k), (int)std::get<1>(k), M.rows(), M.cols());
            for (int i = 0; i < M.rows(); ++i) {
                for (int j = 0; j < M.cols(); ++j) std::printf("%.17g ", M(i, j));
                std::printf("%.17g\n", V(i));
            }
        }
        return 0;
    }
    return 2;
}
