// The reference's src/1d_example.cpp on the gvib200 facade: same construction, same knobs, the iteration runs on the
// GPU.  The only source change a user makes: the cost is named by a device cost class (gvi::Stereo1DCost) instead of
// a host std::function with gvi::NoneType -- the function argument is still accepted and ignored.
// Prints mean / covariance / precision / cost per iteration; with an argument <prefix> it instead runs the 10 iterations
// in one optimize() call and writes <prefix>{mean,cov,precision,cost,factor_costs,...}.csv like the reference does.
//   g++ -std=c++17 -I gaussianvi_b200/cpp examples/1d_example.cpp -L gaussianvi_b200 -lgvib200 -Wl,-rpath,$PWD/gaussianvi_b200
#include <cstdio>

#include "ngd/NGD-GH.h"
#include "ngd/NGDFactorizedBaseGH.h"

using namespace gvi;

double cost_function(const VectorXd& vec_x, const Stereo1DCost& c) {  // src/1d_example.cpp:25-35 (never called)
    const double x = vec_x(0);
    const double y = c.f * c.b / c.mu_p + c.y_offset;
    return (x - c.mu_p) * (x - c.mu_p) / c.sig_p_sq / 2 + (y - c.f * c.b / x) * (y - c.f * c.b / x) / c.sig_r_sq / 2;
}

int main(int argc, char** argv) {
    const int dim_state = 1, num_states = 1, dim_factor = 1, start_index = 0, gh_degree = 10, n_iters = 10;
    const double temperature = 1.0, high_temperature = 10.0;
    using Factor = NGDFactorizedBaseGH<Stereo1DCost>;
    std::vector<std::shared_ptr<Factor>> vec_opt_fact;
    vec_opt_fact.emplace_back(new Factor(dim_factor, dim_state, gh_degree, cost_function, Stereo1DCost(), num_states,
                                         start_index, temperature, high_temperature));
    VectorXd init_mu = VectorXd::Constant(1, 20.0);
    MatrixXd init_prec = MatrixXd::Constant(1, 1, 1.0 / 9.0);

    if (argc > 1) {
        // src/1d_example.cpp:53-82 as written: n_iters iterations in one optimize() call, results saved under a prefix
        // (the reference writes data/1d/{mean,cov,precision,cost,factor_costs}.csv)
        NGDGH<Factor> rec{vec_opt_fact, dim_state, num_states, n_iters};
        rec.set_niter_low_temperature(n_iters);
        rec.set_initial_values(init_mu, init_prec);
        rec.set_step_size_base(0.75);
        rec.update_file_names(argv[1]);
        rec.optimize(false);
        return 0;
    }
    NGDGH<Factor> opt{vec_opt_fact, dim_state, num_states, 1};
    opt.set_niter_low_temperature(n_iters);
    opt.set_initial_values(init_mu, init_prec);
    opt.set_step_size_base(0.75);
    for (int it = 0; it < n_iters; ++it) {
        const double mean = opt.mean()(0), cov = opt.covariance()(0, 0), prec = opt.precision()(0, 0);
        opt.optimize();  // one iteration per call (niterations = 1): the schedule state lives in the problem
        std::printf("%d %.15g %.15g %.15g %.15g\n", it, mean, cov, prec, opt.iteration_stats()[0].cost);
    }
    return 0;
}
