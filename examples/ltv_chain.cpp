// A 2-D point robot (d = 4) under a linear time-varying GP prior (gp/LTV_prior.h) on the gvib200 facade, with the knobs of
// the reference's GPU path: planar hinge factors through NGDFactorizedBaseGH_Cuda<CudaOperation_PlanarPR>, LTV_GP links
// (damped oscillator A(t) = [[0, I], [-w^2 I, -c I]], B = [0; I], piece-wise constant per quarter interval), the EMA
// update set_alpha (gvibase/GVI-GH-Cuda-impl.h:112-114), set_temperature / switch_to_high_temperature, the per-factor
// expectations E_Phis / E_xMuPhis / E_xMuxMuTPhis and SparseGaussHermite::update_parameters / sigmapts.
// tests/test_facade.py drives the same problem through the ctypes mirror and the oracle.
//   usage: ltv_chain S n_iters alpha
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "gp/factorized_opts_LTV.h"
#include "gp/factorized_opts_linear.h"
#include "helpers/CudaOperation.h"
#include "ngd/NGD-GH.h"
#include "ngd/NGDFactorizedBaseGH_Cuda.h"
#include "quadrature/SparseGaussHermite.h"

using namespace gvi;

int main(int argc, char** argv) {
    const int S = argc > 1 ? std::atoi(argv[1]) : 30;
    const int n_iters = argc > 2 ? std::atoi(argv[2]) : 6;
    const double alpha = argc > 3 ? std::atof(argv[3]) : 0.8;
    const int d = 4, gh_degree = 6;
    const double delta_t = 0.2, T = 1.0, Th = 10.0;

    // distance field: one disc (centre (1, 5), radius 2) on a 120 x 160 grid, origin (-20, -10), cell 0.25
    auto sdf = std::make_shared<PlanarSDF>();
    sdf->origin_x = -20;
    sdf->origin_y = -10;
    sdf->cell_size = 0.25;
    sdf->data = MatrixXd::Zero(120, 160);
    for (int c = 0; c < 160; ++c)
        for (int r = 0; r < 120; ++r) sdf->data(r, c) = std::hypot(-20 + 0.25 * c - 1.0, -10 + 0.25 * r - 5.0) - 2.0;
    auto cuda = std::make_shared<CudaOperation_PlanarPR>(0.1, 0.5, 1.0);
    cuda->set_sdf(sdf);

    // nominal trajectory (initial mean, and -target_mean of the LTV prior: sign quirk of gp/LTV_prior.h:87-94)
    std::vector<VectorXd> nominal((size_t)S), target((size_t)S);
    VectorXd mu0 = VectorXd::Zero(S * d);
    for (int i = 0; i < S; ++i) {
        VectorXd v = VectorXd::Zero(d);
        v(0) = -6.0 + 12.0 * i / (S - 1);
        v(1) = 5.0 + 3.5 * std::sin(3.0 * i / (S - 1));
        v(2) = 12.0 / ((S - 1) * delta_t);
        v(3) = 3.5 * 3.0 / ((S - 1) * delta_t) * std::cos(3.0 * i / (S - 1));
        nominal[(size_t)i] = v;
        target[(size_t)i] = -1.0 * v;
        for (int k = 0; k < d; ++k) mu0(i * d + k) = v(k);
    }
    // quarter-interval dynamics
    const int nq = 4 * (S - 1) + 1;
    std::vector<MatrixXd> hA((size_t)nq), hB((size_t)nq);
    for (int q = 0; q < nq; ++q) {
        const double w = 1.5 + 0.4 * std::sin(0.37 * q), c = 1.4 + 0.3 * std::cos(0.21 * q);
        MatrixXd A = MatrixXd::Zero(d, d), B = MatrixXd::Zero(d, 2);
        for (int k = 0; k < 2; ++k) {
            A(k, 2 + k) = 1.0;
            A(2 + k, k) = -w * w;
            A(2 + k, 2 + k) = -c;
            B(2 + k, k) = 1.0;
        }
        hA[(size_t)q] = A;
        hB[(size_t)q] = B;
    }

    using Base = GVIFactorizedBase;
    using Collision = NGDFactorizedBaseGH_Cuda<CudaOperation_PlanarPR>;
    std::vector<std::shared_ptr<Base>> factors;
    auto map = std::make_shared<QuadratureWeightsMap>();
    const MatrixXd K0 = 1e-4 * MatrixXd::Identity(d, d), Qc = MatrixXd::Identity(2, 2);
    for (int i = 0; i < S; ++i) {
        if (i == 0) factors.emplace_back(new FixedGpPrior(d, d, cost_fixed_gp, FixedPriorGP(K0, nominal[0]), S, 0, T, Th));
        if (i == S - 1) factors.emplace_back(new FixedGpPrior(d, d, cost_fixed_gp, FixedPriorGP(K0, nominal[(size_t)S - 1]), S, S - 1, T, Th));
        if (i < S - 1)
            factors.emplace_back(new LTVGpPrior(2 * d, d, nullptr, LTV_GP(Qc, i, delta_t, nominal[0], S, hA, hB, target), S, i, T, Th));
        if (i > 0 && i < S - 1) factors.emplace_back(new Collision(d, d, gh_degree, S, i, 0.1, 0.5, 1.0, T, Th, map, cuda));
    }

    NGDGH<Base> opt{factors, d, S, n_iters};
    opt.classify_factors();
    opt.set_alpha(alpha);
    opt.set_temperature(T);
    opt.set_high_temperature(Th);
    opt.set_stop_err(1e-5);
    opt.set_niter_low_temperature(1 << 30);
    opt.set_mu(mu0);
    opt.initilize_precision_matrix(100.0);
    opt.optimize();
    const auto& st = opt.iteration_stats();
    for (size_t it = 0; it < st.size(); ++it) std::printf("cost %zu %.15g\n", it, st[it].cost);
    VectorXd m = opt.mean();
    for (int i = 0; i < S * d; ++i) std::printf("mean %d %.15g\n", i, m(i));
    // per-factor expectations at the final state, then the high temperature by hand
    const std::vector<double> e0 = opt.E_Phis();
    const std::vector<MatrixXd> e1 = opt.E_xMuPhis(), e2 = opt.E_xMuxMuTPhis();
    for (size_t i = 0; i < e0.size(); ++i)
        std::printf("ephi %zu %.15g %.15g %.15g\n", i, e0[i], e1[i].rows() ? e1[i](0, 0) : 0.0, e2[i].rows() ? e2[i](1, 1) : 0.0);
    const double c_low = opt.cost_value();
    opt.switch_to_high_temperature();
    std::printf("costs %.15g %.15g temperature %.3g\n", c_low, opt.cost_value(), opt.temperature());

    // SparseGaussHermite: parameters switched in place, sigma points rebuilt on the host
    PlanarHingeCost hinge = cuda->cost_class();
    MatrixXd P2 = MatrixXd::Identity(2, 2);
    VectorXd m2 = VectorXd::Zero(2);
    SparseGaussHermite<PlanarHingeCost> gh(4, 2, m2, P2, hinge);
    MatrixXd P4 = MatrixXd::Zero(4, 4);
    VectorXd m4 = VectorXd::Zero(4);
    for (int i = 0; i < 4; ++i) {
        m4(i) = i == 0 ? 1.0 : (i == 1 ? 3.5 : 0.2 * i);
        for (int j = 0; j < 4; ++j) P4(i, j) = (i == j ? 0.5 + 0.1 * i : 0.05);
    }
    gh.update_parameters(gh_degree, 4, m4, P4);
    const auto mom = gh.Integrate();
    const MatrixXd X = gh.sigmapts();
    const VectorXd w = gh.weights();
    double sx = 0.0, sxx = 0.0;  // first two moments of the sigma points reproduce (mean, P)
    for (long i = 0; i < X.rows(); ++i) {
        sx += w(i) * X(i, 1);
        sxx += w(i) * (X(i, 1) - m4(1)) * (X(i, 2) - m4(2));
    }
    std::printf("gh %ld %.15g %.15g %.15g %.15g\n", (long)X.rows(), mom.E_phi, mom.E_xmu_phi(0), sx, sxx);
    return 0;
}
