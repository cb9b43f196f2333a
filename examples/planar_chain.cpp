// A 2-D point robot (d = 4) on the gvib200 facade, built the way VIMP builds its planners on the reference: fixed priors
// at both ends, a minimum-acceleration GP prior between consecutive states, weak anchors, and one planar hinge-SDF
// collision factor per interior state -- interleaved in the order a planner creates them.  Mirrors
// gaussianvi_b200.problems.make_cfg2(S) + a hinge group; tests/test_gpu_facade.py compares the two.
//   usage: planar_chain S n_iters   -> prints "cost <it> <value>" and "mean <i> <value>" lines
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "gp/factorized_opts_linear.h"
#include "ngd/NGD-GH.h"
#include "ngd/NGDFactorizedBaseGH.h"

using namespace gvi;

int main(int argc, char** argv) {
    const int S = argc > 1 ? std::atoi(argv[1]) : 60;
    const int n_iters = argc > 2 ? std::atoi(argv[2]) : 10;
    const int d = 4, gh_degree = 6;
    const double delta_t = 0.1, T = 1.0, Th = 10.0;
    double start[4] = {-15, -5, 0, 0}, goal[4] = {15, 14, 0, 0};

    // signed distance to one disc (centre (0, 4), radius 2.5) on a 120 x 160 grid, origin (-20, -10), cell 0.25
    auto sdf = std::make_shared<PlanarSDF>();
    sdf->origin_x = -20;
    sdf->origin_y = -10;
    sdf->cell_size = 0.25;
    sdf->data = MatrixXd::Zero(120, 160);
    for (int c = 0; c < 160; ++c)
        for (int r = 0; r < 120; ++r)
            sdf->data(r, c) = std::hypot(-20 + 0.25 * c - 0.0, -10 + 0.25 * r - 4.0) - 2.5;
    PlanarHingeCost hinge;
    hinge.sdf = sdf;
    hinge.sigma = 0.1;
    hinge.epsilon = 0.5;
    hinge.radius = 1.0;

    // initial mean: straight line with constant velocity
    VectorXd mu0 = VectorXd::Zero(S * d);
    for (int i = 0; i < S; ++i) {
        const double t = S > 1 ? (double)i / (S - 1) : 0.0;
        for (int k = 0; k < 2; ++k) {
            mu0(i * d + k) = start[k] * (1 - t) + goal[k] * t;
            mu0(i * d + 2 + k) = (goal[k] - start[k]) / ((S - 1) * delta_t);
        }
    }
    auto state = [&](int i) {
        VectorXd v = VectorXd::Zero(d);
        for (int k = 0; k < d; ++k) v(k) = mu0(i * d + k);
        return v;
    };

    using Base = GVIFactorizedBase;
    std::vector<std::shared_ptr<Base>> factors;
    MatrixXd Qc = 0.8 * MatrixXd::Identity(2, 2);
    MatrixXd K0 = 1e-4 * MatrixXd::Identity(d, d), K1 = MatrixXd::Identity(d, d);
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
        if (i > 0 && i < S - 1 && i % 10 == 0)
            factors.emplace_back(new FixedGpPrior(d, d, cost_fixed_gp, FixedPriorGP(K1, state(i)), S, i, T, Th));
        if (i > 0 && i < S - 1)
            factors.emplace_back(new NGDFactorizedBaseGH<PlanarHingeCost>(d, d, gh_degree, nullptr, hinge, S, i, T, Th));
    }

    NGDGH<Base> opt{factors, d, S, n_iters};
    opt.set_initial_values(mu0, 10.0 * MatrixXd::Identity(S * d, S * d));
    opt.optimize();
    const auto& st = opt.iteration_stats();
    for (size_t it = 0; it < st.size(); ++it) std::printf("cost %zu %.15g\n", it, st[it].cost);
    VectorXd m = opt.mean();
    for (int i = 0; i < S * d; ++i) std::printf("mean %d %.15g\n", i, m(i));
    VectorXd fc = opt.factor_cost_vector();
    double s = 0;
    for (long i = 0; i < (long)fc.size(); ++i) s += fc(i);
    std::printf("sumfc %.15g nfactors %ld\n", s, (long)fc.size());
    return 0;
}
