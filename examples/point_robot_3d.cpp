// A 3-D point robot (state = position + velocity, d = 6) written against the reference's GPU-path names
// (NGDFactorizedBaseGH_Cuda<CudaOperation_3dpR>, NGDFactorizedLinear_Cuda; gp/factorized_opts_linear_Cuda.h,
// helpers/CudaOperation.h:610-676): fixed priors at both ends, a minimum-acceleration GP prior between consecutive
// states and one 3-D hinge factor (trilinear SignedDistanceField lookup) per interior state.  The field is the signed
// distance to one ball, attached with set_sdf() (the reference reads maps/3dpR/pRSDF3D.bin from its source tree).
//   usage: point_robot_3d S n_iters   -> prints "cost <it> <value>" and "mean <i> <value>" lines
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "gp/factorized_opts_linear_Cuda.h"
#include "helpers/CudaOperation.h"
#include "ngd/NGD-GH-Cuda.h"
#include "ngd/NGDFactorizedBaseGH_Cuda.h"

using namespace gvi;

int main(int argc, char** argv) {
    const int S = argc > 1 ? std::atoi(argv[1]) : 12;
    const int n_iters = argc > 2 ? std::atoi(argv[2]) : 5;
    const int d = 6, gh_degree = 3;
    const double delta_t = 0.3, T = 1.0, Th = 10.0;
    const double start[6] = {-3.5, -2.5, -1.5, 0, 0, 0}, goal[6] = {3.0, 2.0, 1.0, 0, 0, 0};

    // signed distance to a ball (centre (0, 0.3, 0), radius 0.8) on a 40 x 60 x 80 grid, origin (-4, -3, -2), cell 0.1
    auto sdf = std::make_shared<SignedDistanceField>();
    sdf->origin_x = -4;
    sdf->origin_y = -3;
    sdf->origin_z = -2;
    sdf->cell_size = 0.1;
    const int nz = 40, rows = 60, cols = 80;
    for (int z = 0; z < nz; ++z) {
        MatrixXd slice = MatrixXd::Zero(rows, cols);
        for (int c = 0; c < cols; ++c)
            for (int r = 0; r < rows; ++r) {
                const double x = -4 + 0.1 * c, y = -3 + 0.1 * r, zz = -2 + 0.1 * z;
                slice(r, c) = std::sqrt(x * x + (y - 0.3) * (y - 0.3) + zz * zz) - 0.8;
            }
        sdf->data.push_back(slice);
    }
    auto cuda = std::make_shared<CudaOperation_3dpR>(0.3, 0.5, 1.0);
    cuda->set_sdf(sdf);

    VectorXd mu0 = VectorXd::Zero(S * d);
    for (int i = 0; i < S; ++i) {
        const double t = S > 1 ? (double)i / (S - 1) : 0.0;
        for (int k = 0; k < 3; ++k) {
            mu0(i * d + k) = start[k] * (1 - t) + goal[k] * t;
            mu0(i * d + 3 + k) = (goal[k] - start[k]) / ((S - 1) * delta_t);
        }
    }
    using Base = GVIFactorizedBase;
    using Collision = NGDFactorizedBaseGH_Cuda<CudaOperation_3dpR>;
    std::vector<std::shared_ptr<Base>> factors;
    auto map = std::make_shared<QuadratureWeightsMap>();
    MatrixXd Qc = 0.8 * MatrixXd::Identity(3, 3), K0 = 1e-4 * MatrixXd::Identity(d, d);
    VectorXd vs = VectorXd::Zero(d), vg = VectorXd::Zero(d);
    for (int k = 0; k < d; ++k) {
        vs(k) = start[k];
        vg(k) = goal[k];
    }
    for (int i = 0; i < S; ++i) {
        if (i == 0) factors.emplace_back(new FixedGpPrior(d, d, cost_fixed_gp, FixedPriorGP(K0, vs), S, 0, T, Th));
        if (i == S - 1) factors.emplace_back(new FixedGpPrior(d, d, cost_fixed_gp, FixedPriorGP(K0, vg), S, S - 1, T, Th));
        if (i < S - 1)
            factors.emplace_back(new LinearGpPrior(2 * d, d, cost_linear_gp, MinimumAccGP(Qc, i, delta_t, vs), S, i, T, Th));
        if (i > 0 && i < S - 1) factors.emplace_back(new Collision(d, d, gh_degree, S, i, 0.3, 0.5, 1.0, T, Th, map, cuda));
    }
    NGDGH<Base> opt{factors, d, S, n_iters};
    opt.classify_factors();
    opt.set_initial_values(mu0, 10.0 * MatrixXd::Identity(S * d, S * d));
    opt.optimize();
    const auto& st = opt.iteration_stats();
    for (size_t it = 0; it < st.size(); ++it) std::printf("cost %zu %.15g\n", it, st[it].cost);
    VectorXd m = opt.mean();
    for (int i = 0; i < S * d; ++i) std::printf("mean %d %.15g\n", i, m(i));
    return 0;
}
