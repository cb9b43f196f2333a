// The reference's src/1d_example_proxGVI.cpp on the gvib200 facade (Prox-GVI, 10 iterations); prints mean / covariance /
// precision / cost per iteration in the format of data/1d_proxgvi/*.csv.
#include <cstdio>

#include "proxgd/ProxGVI-GH.h"
#include "proxgd/ProxGVIFactorizedBaseGH.h"

using namespace gvi;

int main() {
    const int n_iters = 10;
    using Factor = ProxGVIFactorizedBaseGH<Stereo1DCost>;
    std::vector<std::shared_ptr<Factor>> vec_opt_fact;
    vec_opt_fact.emplace_back(new Factor(1, 1, 10, nullptr, Stereo1DCost(), 1, 0, 1.0, 10.0));
    ProxGVIGH<Factor> opt{vec_opt_fact, 1, 1, 1};
    opt.set_niter_low_temperature(n_iters);
    opt.set_initial_values(VectorXd::Constant(1, 20.0), MatrixXd::Constant(1, 1, 1.0 / 9.0));
    opt.set_step_size_base(0.75);
    for (int it = 0; it < n_iters; ++it) {
        const double mean = opt.mean()(0), cov = opt.covariance()(0, 0), prec = opt.precision()(0, 0);
        opt.optimize();
        std::printf("%d %.15g %.15g %.15g %.15g\n", it, mean, cov, prec, opt.iteration_stats()[0].cost);
    }
    return 0;
}
