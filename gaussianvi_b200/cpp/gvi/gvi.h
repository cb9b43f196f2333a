// gvi.h -- header-only C++ facade that keeps the reference's operator API (namespace gvi) on top of the gvib200 C-ABI.
//
// A user of hzyu17/GaussianVI builds factors and an optimizer exactly as before:
//     gvi::NGDFactorizedBaseGH<CostClass>(dimension, state_dim, gh_degree, function, cost_class, num_states, start_index,
//                                         temperature, high_temperature)            ngd/NGDFactorizedBaseGH.h:37-41
//     gvi::NGDFactorizedLinear<Factor>(dimension, dim_state, function, linear_factor, num_states, start_indx,
//                                      temperature, high_temperature)               ngd/NGDFactorizedLinear.h:28-35
//     gvi::NGDGH<F>(vec_factors, dim_state, num_states, niters, T, T_high)          ngd/NGD-GH.h:40-53
//     set_initial_values / set_mu / set_precision / set_step_size_base / set_max_iter_backtrack /
//     set_niter_low_temperature / optimize() / mean() / covariance() / precision() / cost_value() /
//     factor_cost_vector() / E_Phis() ...                                          gvibase/GVI-GH-GBP.h:149-378
//     gvi::MinimumAccGP / FixedPriorGP / LTV_GP and the aliases FixedGpPrior / LinearGpPrior
//                                                                                   gp/*.h, gp/factorized_opts_linear.h:9-14
// and the whole iteration runs on the GPU.  The `function` argument stays in the signatures for source compatibility
// but is never called: the device path dispatches on CostClass through DeviceCostTraits<CostClass>; a cost class without
// a specialisation is a compile-time error (there is no CPU fallback).
//
// Host code only: matrices are Eigen's when <Eigen/Dense> is found, else the stand-in of gvi/matrix.h.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/gvib200.h"
#include "matrix.h"

namespace gvi {

struct NoneType {};

inline void gvib200_check(int rc, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string(what) + ": gvib200 error " + std::to_string(rc) + ": " + gvib200_last_error());
}

// one context per process and device (the reference has no notion of a device; GVIB200_DEVICE selects it)
inline gvib200_ctx* default_context() {
    static gvib200_ctx* ctx = nullptr;
    if (!ctx) {
        const char* dev = std::getenv("GVIB200_DEVICE");
        gvib200_check(gvib200_ctx_create(dev ? std::atoi(dev) : 0, &ctx), "gvib200_ctx_create");
    }
    return ctx;
}

// ------------------------------------------------------------------------------------------------ cost classes
// 1-D stereo-camera cost of src/1d_example.cpp:25-35
struct Stereo1DCost {
    double mu_p = 20, f = 400, b = 0.1, sig_r_sq = 0.09, sig_p_sq = 9, y_offset = -0.8;
};
// PlanarSDF(origin, cell_size, data), helpers/CudaOperation.h:27-128
struct PlanarSDF {
    double origin_x = 0, origin_y = 0, cell_size = 1;
    MatrixXd data;  // rows x cols signed distances
};
// cost_obstacle_planar of CudaOperation_PlanarPR, helpers/CudaOperation.h:491-508 (defaults :456)
struct PlanarHingeCost {
    std::shared_ptr<PlanarSDF> sdf;
    double sigma = 15.5, epsilon = 0.5, radius = 1.0;
};
// x^T (c I) x, the integrand of tests/test_gh_spgh.cpp:21-25
struct QuadraticCost {
    double c = 1.0;
};
// SignedDistanceField(origin, cell_size, data), helpers/CudaOperation.h:133-160: z slices of rows x cols matrices
struct SignedDistanceField {
    double origin_x = 0, origin_y = 0, origin_z = 0, cell_size = 1;
    std::vector<MatrixXd> data;  // data[z] is rows x cols
};
// cost_obstacle_planar of CudaOperation_3dpR, helpers/CudaOperation.h:641-674 (3-D point robot, x(0:3) = position)
struct Hinge3DCost {
    std::shared_ptr<SignedDistanceField> sdf;
    double sigma = 15.5, epsilon = 0.5, radius = 1.0;
};
// cost_obstacle_planar of CudaOperation_Quad, helpers/CudaOperation.h:565-605 (planar quadrotor, x = (x, z, phi, ...))
struct QuadHingeCost {
    std::shared_ptr<PlanarSDF> sdf;
    double sigma = 15.5, epsilon = 0.5, radius = 1.0;
};

// cost_obstacle of CudaOperation_3dArm over ForwardKinematics, helpers/CudaOperation.h:325-410,751-770: a serial arm in
// Denavit-Hartenberg form with body spheres; state = (joint angles, joint velocities), at most 3 joints (state blocks of the
// chain engine go up to 6 x 6)
struct Arm3DCost {
    std::shared_ptr<SignedDistanceField> sdf;
    VectorXd a, alpha, d, theta_bias, radii;
    std::vector<int> frames;
    MatrixXd centers;  // one row per sphere
    double sigma = 15.5, epsilon = 0.5;
};

struct DeviceCostSpec {
    int kind = 0;
    std::vector<unsigned char> params;        // one struct (shared by the group)
    std::shared_ptr<PlanarSDF> sdf;           // planar hinge / quadrotor hinge
    std::shared_ptr<SignedDistanceField> sdf3d;  // 3-D hinge
    bool operator<(const DeviceCostSpec& o) const {
        if (kind != o.kind) return kind < o.kind;
        if (sdf.get() != o.sdf.get()) return sdf.get() < o.sdf.get();
        if (sdf3d.get() != o.sdf3d.get()) return sdf3d.get() < o.sdf3d.get();
        return params < o.params;
    }
};
// hand the field(s) of a cost spec to a problem (before the factors that use them are added)
inline void upload_cost_fields(gvib200_problem* prob, const DeviceCostSpec& c) {
    if (c.sdf) {
        const PlanarSDF& s = *c.sdf;
        gvib200_check(gvib200_set_planar_sdf(prob, (int)s.data.rows(), (int)s.data.cols(), s.origin_x, s.origin_y, s.cell_size,
                                             s.data.data()), "set_planar_sdf");
    }
    if (c.sdf3d && !c.sdf3d->data.empty()) {
        const SignedDistanceField& s = *c.sdf3d;
        const int rows = (int)s.data[0].rows(), cols = (int)s.data[0].cols(), nz = (int)s.data.size();
        std::vector<double> flat((size_t)rows * cols * nz);  // data[r + c * rows + z * rows * cols], CudaOperation.h:299-301
        for (int z = 0; z < nz; ++z)
            for (int cc = 0; cc < cols; ++cc)
                for (int r = 0; r < rows; ++r) flat[(size_t)r + (size_t)cc * rows + (size_t)z * rows * cols] = s.data[(size_t)z](r, cc);
        gvib200_check(gvib200_set_sdf3d(prob, rows, cols, nz, s.origin_x, s.origin_y, s.origin_z, s.cell_size, flat.data()),
                      "set_sdf3d");
    }
}
template <class P>
inline std::vector<unsigned char> pod_bytes(const P& p) {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(&p);
    return std::vector<unsigned char>(b, b + sizeof(P));
}

template <class CostClass>
struct DeviceCostTraits;  // no definition: an unspecialised cost class cannot run (compile-time error)
template <>
struct DeviceCostTraits<Stereo1DCost> {
    static DeviceCostSpec spec(const Stereo1DCost& c) {
        gvib200_stereo1d_params p{c.mu_p, c.f, c.b, c.sig_r_sq, c.sig_p_sq, c.y_offset};
        return DeviceCostSpec{GVIB200_COST_STEREO_1D, pod_bytes(p), nullptr, nullptr};
    }
};
template <>
struct DeviceCostTraits<PlanarHingeCost> {
    static DeviceCostSpec spec(const PlanarHingeCost& c) {
        gvib200_hinge_params p{c.sigma, c.epsilon, c.radius};
        return DeviceCostSpec{GVIB200_COST_PLANAR_HINGE, pod_bytes(p), c.sdf, nullptr};
    }
};
template <>
struct DeviceCostTraits<Hinge3DCost> {
    static DeviceCostSpec spec(const Hinge3DCost& c) {
        gvib200_hinge_params p{c.sigma, c.epsilon, c.radius};
        return DeviceCostSpec{GVIB200_COST_HINGE_3D, pod_bytes(p), nullptr, c.sdf};
    }
};
template <>
struct DeviceCostTraits<Arm3DCost> {
    static DeviceCostSpec spec(const Arm3DCost& c) {
        gvib200_arm_params p;
        std::memset(&p, 0, sizeof(p));
        p.sigma = c.sigma;
        p.epsilon = c.epsilon;
        p.n_dof = (int)c.a.size();
        p.n_spheres = (int)c.frames.size();
        if (p.n_dof > GVIB200_ARM_MAX_DOF || p.n_spheres > GVIB200_ARM_MAX_SPHERES)
            throw std::invalid_argument("Arm3DCost: at most 3 joints and 12 body spheres");
        for (int j = 0; j < p.n_dof; ++j) {
            p.a[j] = c.a(j);
            p.alpha[j] = c.alpha(j);
            p.d[j] = c.d(j);
            p.theta_bias[j] = c.theta_bias(j);
        }
        for (int i = 0; i < p.n_spheres; ++i) {
            p.frames[i] = c.frames[(size_t)i];
            p.radii[i] = c.radii(i);
            for (int k = 0; k < 3; ++k) p.centers[i][k] = c.centers(i, k);
        }
        return DeviceCostSpec{GVIB200_COST_ARM_3D, pod_bytes(p), nullptr, c.sdf};
    }
};
template <>
struct DeviceCostTraits<QuadHingeCost> {
    static DeviceCostSpec spec(const QuadHingeCost& c) {
        gvib200_hinge_params p{c.sigma, c.epsilon, c.radius};
        return DeviceCostSpec{GVIB200_COST_QUAD_HINGE, pod_bytes(p), c.sdf, nullptr};
    }
};
template <>
struct DeviceCostTraits<QuadraticCost> {
    static DeviceCostSpec spec(const QuadraticCost& c) { return DeviceCostSpec{GVIB200_COST_QUADRATIC, pod_bytes(c.c), nullptr, nullptr}; }
};

// ------------------------------------------------------------------------------------------------ linear priors
// gp/linear_factor.h:18-31
class LinearFactor {
public:
    virtual ~LinearFactor() {}
    virtual VectorXd get_mu() const = 0;
    virtual MatrixXd get_covariance() const = 0;
    virtual MatrixXd get_precision() const = 0;
    virtual MatrixXd get_Lambda() const = 0;
    virtual MatrixXd get_Psi() const = 0;
    virtual double get_Constant() const = 0;
};

// gp/fixed_prior.h:19-50
class FixedPriorGP : public LinearFactor {
public:
    FixedPriorGP() {}
    FixedPriorGP(const MatrixXd& Covariance, const VectorXd& mu) : _K(Covariance), _invK(Covariance.inverse()), _dim((int)mu.size()), _mu(mu) {}
    VectorXd get_mu() const override { return _mu; }
    MatrixXd get_precision() const override { return _invK; }
    MatrixXd get_covariance() const override { return _K; }
    MatrixXd get_Lambda() const override { return MatrixXd::Identity(_dim, _dim); }
    MatrixXd get_Psi() const override { return MatrixXd::Identity(_dim, _dim); }
    double get_Constant() const override { return 1.0; }

private:
    MatrixXd _K, _invK;
    int _dim = 0;
    VectorXd _mu;
};

namespace detail {
inline MatrixXd lambda_from_phi(const MatrixXd& Phi) {  // [-Phi, I]
    const long n = Phi.rows();
    MatrixXd L = MatrixXd::Zero(n, 2 * n);
    for (long j = 0; j < n; ++j)
        for (long i = 0; i < n; ++i) L(i, j) = -Phi(i, j);
    for (long i = 0; i < n; ++i) L(i, n + i) = 1.0;
    return L;
}
// exp of a small square matrix: scaling and squaring with a Taylor series
inline MatrixXd expm(const MatrixXd& X) {
    const long n = X.rows();
    double nrm = 0.0;
    for (long i = 0; i < n; ++i) {
        double s = 0.0;
        for (long j = 0; j < n; ++j) s += std::fabs(X(i, j));
        nrm = nrm > s ? nrm : s;
    }
    int sq = 0;
    while (nrm > 0.25) {
        nrm *= 0.5;
        ++sq;
    }
    MatrixXd Y = X * (1.0 / std::pow(2.0, sq));
    MatrixXd E = MatrixXd::Identity(n, n), term = MatrixXd::Identity(n, n);
    for (int k = 1; k < 19; ++k) {
        term = (term * Y) * (1.0 / k);
        E = E + term;
    }
    for (int s = 0; s < sq; ++s) E = E * E;
    return E;
}
// symmetric square root V sqrt(D) V^T of a symmetric PSD matrix (SelfAdjointEigenSolver::operatorSqrt,
// quadrature/SparseGaussHermite.h:232-233) by cyclic Jacobi rotations -- host-side, for the sigmapts() accessor only
inline MatrixXd sym_sqrt(const MatrixXd& A) {
    const long n = A.rows();
    MatrixXd M = A, V = MatrixXd::Identity(n, n);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (long q = 1; q < n; ++q)
            for (long p = 0; p < q; ++p) off += M(p, q) * M(p, q);
        if (off < 1e-300) break;
        for (long q = 1; q < n; ++q)
            for (long p = 0; p < q; ++p) {
                if (M(p, q) == 0.0) continue;
                const double theta = (M(q, q) - M(p, p)) / (2.0 * M(p, q));
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (long k = 0; k < n; ++k) {
                    const double a = M(k, p), b = M(k, q);
                    M(k, p) = c * a - sn * b;
                    M(k, q) = sn * a + c * b;
                }
                for (long k = 0; k < n; ++k) {
                    const double a = M(p, k), b = M(q, k);
                    M(p, k) = c * a - sn * b;
                    M(q, k) = sn * a + c * b;
                }
                for (long k = 0; k < n; ++k) {
                    const double a = V(k, p), b = V(k, q);
                    V(k, p) = c * a - sn * b;
                    V(k, q) = sn * a + c * b;
                }
            }
    }
    MatrixXd S = MatrixXd::Zero(n, n);
    for (long j = 0; j < n; ++j)
        for (long i = 0; i < n; ++i) {
            double v = 0.0;
            for (long k = 0; k < n; ++k) v += V(i, k) * std::sqrt(M(k, k) > 0.0 ? M(k, k) : 0.0) * V(j, k);
            S(i, j) = v;
        }
    return S;
}
}  // namespace detail

// gp/minimum_acc_prior.h:26-130
class MinimumAccGP : public LinearFactor {
public:
    MinimumAccGP() {}
    MinimumAccGP(const MatrixXd& Qc, double start_index, const double& delta_t, const VectorXd& mu_0)
        : _dim((int)Qc.cols()), _dim_state(2 * (int)Qc.cols()), _delta_t(delta_t), _Qc(Qc) {
        (void)start_index;
        (void)mu_0;
        const int n = _dim, ns = _dim_state;
        _Phi = MatrixXd::Identity(ns, ns);
        for (int i = 0; i < n; ++i) _Phi(i, n + i) = delta_t;
        const MatrixXd iq = Qc.inverse();
        _invQ = MatrixXd::Zero(ns, ns);
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                _invQ(i, j) = 12 * iq(i, j) / std::pow(delta_t, 3);
                _invQ(i, n + j) = -6 * iq(i, j) / std::pow(delta_t, 2);
                _invQ(n + i, j) = -6 * iq(i, j) / std::pow(delta_t, 2);
                _invQ(n + i, n + j) = 4 * iq(i, j) / delta_t;
            }
        _Lambda = detail::lambda_from_phi(_Phi);
        _Psi = MatrixXd::Zero(ns, 2 * ns);  // a(t) = 0: eliminated (:76-79)
        _target_mu = VectorXd::Zero(2 * ns);
    }
    MatrixXd Phi() const { return _Phi; }
    MatrixXd Qc() const { return _Qc; }
    int dim_posvel() const { return 2 * _dim; }
    VectorXd get_mu() const override { return _target_mu; }
    MatrixXd get_precision() const override { return _invQ; }
    MatrixXd get_covariance() const override { return _invQ.inverse(); }
    MatrixXd get_Lambda() const override { return _Lambda; }
    MatrixXd get_Psi() const override { return _Psi; }
    double get_Constant() const override { return 0.5; }

private:
    int _dim = 0, _dim_state = 0;
    double _delta_t = 0;
    MatrixXd _Qc, _invQ, _Phi, _Lambda, _Psi;
    VectorXd _target_mu;
};

// gp/LTV_prior.h:28-247.  Phi' = A Phi, Q' = A Q + Q A^T + B B^T with A, B piece-wise constant on the four quarter
// intervals (:187-197); the reference integrates with GSL rkf45 at tolerance 1e-12, here each quarter is integrated
// exactly with Van Loan's block exponential.
class LTV_GP : public LinearFactor {
public:
    LTV_GP() {}
    LTV_GP(const MatrixXd& Qc, int start_index, const double& delta_t, const VectorXd& mu_0, int n_states,
           const std::vector<MatrixXd>& hA, const std::vector<MatrixXd>& hB, const std::vector<VectorXd>& target_mean)
        : _dim((int)Qc.cols()), _dim_state(2 * (int)Qc.cols()) {
        (void)mu_0;
        (void)n_states;
        const int ns = _dim_state;
        const double h = delta_t / 4.0;
        MatrixXd Phi = MatrixXd::Identity(ns, ns), Q = MatrixXd::Zero(ns, ns);
        for (int k = 0; k < 4; ++k) {
            const MatrixXd& A = hA[4 * start_index + k];
            const MatrixXd& B = hB[4 * start_index + k];
            const MatrixXd BBt = B * B.transpose();
            MatrixXd M = MatrixXd::Zero(2 * ns, 2 * ns);
            for (int j = 0; j < ns; ++j)
                for (int i = 0; i < ns; ++i) {
                    M(i, j) = -A(i, j) * h;
                    M(i, ns + j) = BBt(i, j) * h;
                    M(ns + i, ns + j) = A(j, i) * h;
                }
            const MatrixXd E = detail::expm(M);
            MatrixXd Phik(ns, ns), E12(ns, ns);
            for (int j = 0; j < ns; ++j)
                for (int i = 0; i < ns; ++i) {
                    Phik(i, j) = E(ns + j, ns + i);
                    E12(i, j) = E(i, ns + j);
                }
            MatrixXd Qk = Phik * E12;
            Qk = 0.5 * (Qk + Qk.transpose());
            Q = Phik * Q * Phik.transpose() + Qk;
            Phi = Phik * Phi;
        }
        _Phi = Phi;
        _Q = 0.5 * (Q + Q.transpose());
        MatrixXd iq = _Q.inverse();
        _invQ = 0.5 * (iq + iq.transpose());
        _Lambda = detail::lambda_from_phi(_Phi);
        _Psi = -1.0 * _Lambda;  // [Phi, -I] (:91-94; note the sign quirk: the prior is centred on -target_mean)
        _target_mu = VectorXd::Zero(2 * ns);
        for (int i = 0; i < ns; ++i) {
            _target_mu(i) = target_mean[start_index](i);
            _target_mu(ns + i) = target_mean[start_index + 1](i);
        }
    }
    MatrixXd Phi() const { return _Phi; }
    MatrixXd Q() const { return _Q; }
    VectorXd get_mu() const override { return _target_mu; }
    MatrixXd get_precision() const override { return _invQ; }
    MatrixXd get_covariance() const override { return _Q; }
    MatrixXd get_Lambda() const override { return _Lambda; }
    MatrixXd get_Psi() const override { return _Psi; }
    double get_Constant() const override { return 0.5; }

private:
    int _dim = 0, _dim_state = 0;
    MatrixXd _Phi, _Q, _invQ, _Lambda, _Psi;
    VectorXd _target_mu;
};

// ------------------------------------------------------------------------------------------------ factors
// GVIFactorizedBase (gvibase/GVIFactorizedBase.h:49-71): what the joint optimizer needs to know about one factor
class GVIFactorizedBase {
public:
    virtual ~GVIFactorizedBase() {}
    GVIFactorizedBase(int dimension, int state_dimension, int num_states, int start_index, double temperature,
                      double high_temperature)
        : _dim(dimension), _state_dim(state_dimension), _num_states(num_states), _start_index(start_index),
          _temperature(temperature), _high_temperature(high_temperature) {}
    int dimension() const { return _dim; }
    int start_index() const { return _start_index; }
    double temperature() const { return _temperature; }
    double high_temperature() const { return _high_temperature; }
    virtual bool is_linear() const = 0;
    // what the joint optimizer needs to batch the factor onto the device (type-erased)
    virtual const DeviceCostSpec* cost_spec() const { return nullptr; }
    virtual int gh_degree() const { return 0; }
    virtual const MatrixXd* lin_Lambda() const { return nullptr; }
    virtual const MatrixXd* lin_Psi() const { return nullptr; }
    virtual const MatrixXd* lin_Kinv() const { return nullptr; }
    virtual const VectorXd* lin_mu() const { return nullptr; }
    virtual double lin_constant() const { return 0.0; }

    int _dim, _state_dim, _num_states, _start_index;
    double _temperature, _high_temperature;
};

// ngd/NGDFactorizedBaseGH.h:37-129
template <class CostClass>
class NGDFactorizedBaseGH : public GVIFactorizedBase {
public:
    using GHFunction = std::function<double(const VectorXd&, const CostClass&)>;
    NGDFactorizedBaseGH(int dimension, int state_dim, int gh_degree, const GHFunction& function, const CostClass& cost_class,
                        int num_states, int start_index, double temperature = 1.0, double high_temperature = 10.0)
        : GVIFactorizedBase(dimension, state_dim, num_states, start_index, temperature, high_temperature),
          _gh_degree(gh_degree), _cost(DeviceCostTraits<CostClass>::spec(cost_class)) {
        (void)function;  // kept for source compatibility; the device functor is selected by CostClass
    }
    bool is_linear() const override { return false; }
    const DeviceCostSpec* cost_spec() const override { return &_cost; }
    int gh_degree() const override { return _gh_degree; }
    int _gh_degree;
    DeviceCostSpec _cost;
};
// NGDFactorizedSimpleGH = NGDFactorizedBaseGH<NoneType> in the reference (an arbitrary host function): it has no device
// functor, so instantiating it fails to compile -- use a cost class with DeviceCostTraits instead.
using NGDFactorizedSimpleGH = NGDFactorizedBaseGH<NoneType>;

// ngd/NGDFactorizedLinear.h:28-129
template <class Factor>
class NGDFactorizedLinear : public GVIFactorizedBase {
public:
    using CostFunction = std::function<double(const VectorXd&, const Factor&)>;
    NGDFactorizedLinear(const int& dimension, int dim_state, const CostFunction& function, const Factor& linear_factor,
                        int num_states, int start_indx, double temperature = 1.0, double high_temperature = 10.0)
        : GVIFactorizedBase(dimension, dim_state, num_states, start_indx, temperature, high_temperature),
          _target_mean(linear_factor.get_mu()), _target_precision(linear_factor.get_precision()),
          _Lambda(linear_factor.get_Lambda()), _Psi(linear_factor.get_Psi()), _constant(linear_factor.get_Constant()) {
        (void)function;
    }
    bool is_linear() const override { return true; }
    const MatrixXd* lin_Lambda() const override { return &_Lambda; }
    const MatrixXd* lin_Psi() const override { return &_Psi; }
    const MatrixXd* lin_Kinv() const override { return &_target_precision; }
    const VectorXd* lin_mu() const override { return &_target_mean; }
    double lin_constant() const override { return _constant; }
    VectorXd _target_mean;
    MatrixXd _target_precision, _Lambda, _Psi;
    double _constant;
};
using FixedGpPrior = NGDFactorizedLinear<FixedPriorGP>;    // gp/factorized_opts_linear.h:9-10
using LinearGpPrior = NGDFactorizedLinear<MinimumAccGP>;
using LTVGpPrior = NGDFactorizedLinear<LTV_GP>;            // gp/factorized_opts_LTV.h

// ------------------------------------------------------------------------------------------------ "_Cuda" aliases
// The reference's own GPU path names its classes ..._Cuda and hands every nonlinear factor a CudaOperation_* object that
// owns the distance field and the hinge parameters (helpers/CudaOperation.h:413-779, ngd/NGDFactorizedBaseGH_Cuda.h:25-50,
// gp/factorized_opts_linear_Cuda.h:9-14).  Callers written against those names (VIMP's GPU planners) compile against the
// stand-ins below; everything runs on the one device path of this library.  The reference's CudaOperation_* constructors
// read their map from a file of the source tree (:463-468, :612-616); here the field is attached with set_sdf().
struct QuadratureWeightsMap {};  // quadrature/SparseGHQuadratureWeights.h:16 -- the table lives in the library (gvib200_table_*)
struct CudaOperation_PlanarPR {  // helpers/CudaOperation.h:452-531
    CudaOperation_PlanarPR(double cost_sigma = 15.5, double epsilon = 0.5, double radius = 1) : _sigma(cost_sigma), _epsilon(epsilon), _radius(radius) {}
    void set_sdf(const std::shared_ptr<PlanarSDF>& sdf) { _sdf = sdf; }
    using CostClass = PlanarHingeCost;
    CostClass cost_class() const { return CostClass{_sdf, _sigma, _epsilon, _radius}; }
    double _sigma, _epsilon, _radius;
    std::shared_ptr<PlanarSDF> _sdf;
};
struct CudaOperation_Quad {  // helpers/CudaOperation.h:534-607
    CudaOperation_Quad(double cost_sigma = 15.5, double epsilon = 0.5, double radius = 1) : _sigma(cost_sigma), _epsilon(epsilon), _radius(radius) {}
    void set_sdf(const std::shared_ptr<PlanarSDF>& sdf) { _sdf = sdf; }
    using CostClass = QuadHingeCost;
    CostClass cost_class() const { return CostClass{_sdf, _sigma, _epsilon, _radius}; }
    double _sigma, _epsilon, _radius;
    std::shared_ptr<PlanarSDF> _sdf;
};
struct CudaOperation_3dpR {  // helpers/CudaOperation.h:610-676
    CudaOperation_3dpR(double cost_sigma = 15.5, double epsilon = 0.5, double radius = 1) : _sigma(cost_sigma), _epsilon(epsilon), _radius(radius) {}
    void set_sdf(const std::shared_ptr<SignedDistanceField>& sdf) { _sdf = sdf; }
    using CostClass = Hinge3DCost;
    CostClass cost_class() const { return CostClass{_sdf, _sigma, _epsilon, _radius}; }
    double _sigma, _epsilon, _radius;
    std::shared_ptr<SignedDistanceField> _sdf;
};

struct CudaOperation_3dArm {  // helpers/CudaOperation.h:680-793 (the distance field is attached with set_sdf, not read from a file)
    CudaOperation_3dArm(const VectorXd& a, const VectorXd& alpha, const VectorXd& d, const VectorXd& theta_bias, const VectorXd& radii,
                        const std::vector<int>& frames, const MatrixXd& centers, double cost_sigma = 15.5, double epsilon = 0.5)
        : _sigma(cost_sigma), _epsilon(epsilon), _radius(0.0) {
        _cost.a = a;
        _cost.alpha = alpha;
        _cost.d = d;
        _cost.theta_bias = theta_bias;
        _cost.radii = radii;
        _cost.frames = frames;
        _cost.centers = centers;
    }
    void set_sdf(const std::shared_ptr<SignedDistanceField>& sdf) { _cost.sdf = sdf; }
    // The reference's factor constructor hands (cost_sigma, epsilon, radius) to every CudaOperation class; the arm takes its
    // radii from the body spheres, so the `radius` member written by NGDFactorizedBaseGH_Cuda is accepted and unused.
    struct CostClass : Arm3DCost {
        double radius = 0.0;
    };
    CostClass cost_class() const {
        CostClass k;
        static_cast<Arm3DCost&>(k) = _cost;
        k.sigma = _sigma;
        k.epsilon = _epsilon;
        return k;
    }
    double _sigma, _epsilon, _radius;
    Arm3DCost _cost;
};
template <>
struct DeviceCostTraits<CudaOperation_3dArm::CostClass> {
    static DeviceCostSpec spec(const CudaOperation_3dArm::CostClass& c) { return DeviceCostTraits<Arm3DCost>::spec(c); }
};

// NGDFactorizedBaseGH_Cuda<CudaClass>(dimension, state_dim, gh_degree, num_states, start_index, cost_sigma, epsilon, radius,
//                                     temperature, high_temperature, weight_sigpts_map_option, cuda_ptr)
// ngd/NGDFactorizedBaseGH_Cuda.h:35-50.  The hinge parameters of the constructor win over the CudaClass object's, as in
// the reference (the factor keeps its own _sigma / _epsilon / _radius).
template <class CudaClass>
class NGDFactorizedBaseGH_Cuda : public NGDFactorizedBaseGH<typename CudaClass::CostClass> {
    using Cost = typename CudaClass::CostClass;
    static Cost with_params(const CudaClass& c, double sigma, double epsilon, double radius) {
        Cost k = c.cost_class();
        k.sigma = sigma;
        k.epsilon = epsilon;
        k.radius = radius;
        return k;
    }

public:
    NGDFactorizedBaseGH_Cuda(int dimension, int state_dim, int gh_degree, int num_states, int start_index, double cost_sigma,
                             double epsilon, double radius, double temperature, double high_temperature,
                             std::shared_ptr<QuadratureWeightsMap> /*weight_sigpts_map_option*/, std::shared_ptr<CudaClass> cuda_ptr)
        : NGDFactorizedBaseGH<Cost>(dimension, state_dim, gh_degree, nullptr, with_params(*cuda_ptr, cost_sigma, epsilon, radius),
                                    num_states, start_index, temperature, high_temperature) {}
    // hooks of the reference's GPU path (gvibase/GVIFactorizedBase_Cuda.h:153-189): nothing to do, the state is device resident
    void cuda_init() {}
    void cuda_free() {}
};
template <class Factor>
using NGDFactorizedLinear_Cuda = NGDFactorizedLinear<Factor>;  // ngd/NGDFactorizedLinear_Cuda.h:28-41 (same constructor)

// cost_fixed_gp / cost_linear_gp (gp/cost_functions.h:25-39): signature placeholders for source compatibility
inline double cost_fixed_gp(const VectorXd&, const FixedPriorGP&) { return 0.0; }
inline double cost_linear_gp(const VectorXd&, const MinimumAccGP&) { return 0.0; }

// block-tridiagonal result (GVIGH::covariance() / precision() return exactly this pattern as an SpMat in the reference)
struct BlockTridiagonal {
    int num_states = 0, dim_state = 0;
    std::vector<double> diag, off;  // [S][d*d], [S-1][d*d] (block (i, i+1)), column-major blocks
    double operator()(long r, long c) const {
        const int d = dim_state;
        const long bi = r / d, bj = c / d, i = r % d, j = c % d;
        if (bi == bj) return diag[(size_t)bi * d * d + i + j * d];
        if (bj == bi + 1) return off[(size_t)bi * d * d + i + j * d];
        if (bi == bj + 1) return off[(size_t)bj * d * d + j + i * d];
        return 0.0;
    }
    MatrixXd toDense() const {
        const long n = (long)num_states * dim_state;
        MatrixXd M = MatrixXd::Zero(n, n);
        for (long c = 0; c < n; ++c)
            for (long r = (c >= dim_state ? c - 2 * dim_state + 1 : 0); r < n && r < c + 2 * dim_state; ++r)
                if (r >= 0) M(r, c) = (*this)(r, c);
        return M;
    }
};

// ------------------------------------------------------------------------------------------------ joint optimizer
// GVIGH (gvibase/GVI-GH-GBP.h:28-433) -- the control skeleton stays on the host, the state lives on the device
template <class FactorizedOptimizer>
class GVIGH {
public:
    GVIGH(const std::vector<std::shared_ptr<FactorizedOptimizer>>& vec_fact_optimizers, int dim_state, int num_states,
          int niterations = 5, double temperature = 1.0, double high_temperature = 100.0)
        : _dim_state(dim_state), _num_states(num_states), _dim(dim_state * num_states), _niters(niterations),
          _temperature(temperature), _high_temperature(high_temperature) {
        for (auto& f : vec_fact_optimizers) _factors.push_back(f);
        gvib200_default_opts(&_opts);
    }
    virtual ~GVIGH() {
        if (_prob) gvib200_problem_destroy(_prob);
    }
    GVIGH(const GVIGH&) = delete;
    GVIGH& operator=(const GVIGH&) = delete;

    // factors of other concrete types can be appended before the first use (a joint problem mixes factor types)
    void add_factor(const std::shared_ptr<GVIFactorizedBase>& f) {
        if (_prob) throw std::logic_error("add_factor after the problem was built");
        _factors.push_back(f);
    }

    // ---- knobs (gvibase/GVI-GH-GBP.h:168-248) ----
    void set_step_size_base(double v) { _opts.step_size_base = v; }
    void set_max_iter_backtrack(int v) { _opts.max_backtrack = v; }
    void set_niter_low_temperature(int v) { _opts.niters_lowtemp = v; }
    void set_niterations(int v) { _niters = v; }
    void set_reuse_accepted_sweep(bool v) { _opts.reuse_accepted_sweep = v ? 1 : 0; }
    // hooks of the reference's GPU path (gvibase/GVI-GH-Cuda.h:140,223,392): the factors are classified and resident on the
    // device by construction; the EMA of update_proposal (GVI-GH-Cuda-impl.h:112-114) is the identity at its default 1
    void classify_factors() {}
    // EMA of an accepted proposal (gvibase/GVI-GH-Cuda.h:223, GVI-GH-Cuda-impl.h:112-114): alpha * new + (1 - alpha) * current
    void set_alpha(double alpha) {
        if (!(alpha > 0.0) || alpha > 1.0) throw std::invalid_argument("set_alpha: alpha must be in (0, 1]");
        _opts.ema_alpha = alpha;
    }
    // gvibase/GVI-GH-GBP.h:187-194,218-224,261-267: bookkeeping members of the optimizer.  As in the reference the
    // temperatures that enter the arithmetic are the factors' own (constructor arguments of every factor); the step size
    // of set_step_size is not used by the back-tracking loop (step_size_base is), and stop_err is never read.
    void set_step_size(double step_size) { _step_size = step_size; }
    void set_stop_err(double stop_err) { _stop_err = stop_err; }
    void set_temperature(double temperature) { _temperature = temperature; }
    void set_high_temperature(double high_temp) { _high_temperature = high_temp; }
    double temperature() const { return _temperature; }
    void set_initial_precision_factor(double f) { _initial_precision_factor = f; }
    void initilize_precision_matrix() { initilize_precision_matrix(_initial_precision_factor); }
    // gvibase/GVI-GH-GBP-impl.h:18-27 (public, gvibase/GVI-GH-GBP.h:361): all factors take their high temperature now
    void switch_to_high_temperature() {
        build();
        gvib200_check(gvib200_switch_to_high_temperature(_prob), "switch_to_high_temperature");
        _temperature = _high_temperature;
        _is_high_T = true;
    }
    void time_test() {
        build();
        float ms[4] = {0, 0, 0, 0};
        for (int stage = 0; stage < 4; ++stage) gvib200_check(gvib200_time_stage(_prob, stage, 10, &_opts, &ms[stage], nullptr), "time_test");
        std::printf("time_test (ms / call): moment sweep %.4f  cost sweep %.4f  assemble + dmu solve %.4f  candidate + selected inverse %.4f\n",
                    ms[0], ms[1], ms[2], ms[3]);
    }

    void set_mu(const VectorXd& mean) {
        _mu0.assign(mean.data(), mean.data() + mean.size());
        if (_prob && _has_state) gvib200_check(gvib200_set_state(_prob, _mu0.data(), nullptr, nullptr), "set_mu");
    }
    // precision given as its block-tridiagonal part: any matrix type with operator()(row, col)
    template <class Mat>
    void set_precision(const Mat& P) {
        const int d = _dim_state, S = _num_states;
        _pd.assign((size_t)S * d * d, 0.0);
        _po.assign((size_t)(S > 1 ? S - 1 : 1) * d * d, 0.0);
        for (int s = 0; s < S; ++s)
            for (int j = 0; j < d; ++j)
                for (int i = 0; i < d; ++i) {
                    _pd[(size_t)s * d * d + i + j * d] = P(s * d + i, s * d + j);
                    if (s + 1 < S) _po[(size_t)s * d * d + i + j * d] = P(s * d + i, (s + 1) * d + j);
                }
        push_state();
    }
    template <class Mat>
    void set_initial_values(const VectorXd& mean, const Mat& P) {
        _mu0.assign(mean.data(), mean.data() + mean.size());
        set_precision(P);
    }
    void initilize_precision_matrix(double initial_precision_factor) {  // (sic) gvibase/GVI-GH.h:201-204
        _initial_precision_factor = initial_precision_factor;
        const int d = _dim_state, S = _num_states;
        _pd.assign((size_t)S * d * d, 0.0);
        _po.assign((size_t)(S > 1 ? S - 1 : 1) * d * d, 0.0);
        for (int s = 0; s < S; ++s)
            for (int i = 0; i < d; ++i) _pd[(size_t)s * d * d + i + i * d] = initial_precision_factor;
        push_state();
    }

    // ---- the loop (gvibase/GVI-GH-GBP-impl.h:33-130) ----
    void optimize(std::optional<bool> verbose = std::nullopt) { run_loop(false, verbose.value_or(false)); }

    // ---- result files (gvibase/GVI-GH.h:282-329, helpers/DataRecorder.h:154-224) ----
    // Naming the files switches the recorder on: optimize() then records mean / marginal covariance / marginal precision /
    // cost / factor costs at the start of every iteration and writes <prefix>{mean,cov,precision,cost,factor_costs,zk_sdf,
    // Sk_sdf}[_afterfix].csv when it ends (the dense joint_cov / joint_precision files only for joint dimension <= 64).
    void update_file_names(const std::string& prefix = "", const std::string& afterfix = "") {
        _file_prefix = prefix;
        _file_afterfix = afterfix;
        _record = true;
    }
    void save_data(bool verbose = true) {
        if (_trace.n_recorded < 1) return;
        if (verbose) std::printf("Saving data to: %s*.csv\n", _file_prefix.c_str());
        gvib200_check(gvib200_trace_save(&_trace, _num_states, _dim_state, (int)_factors.size(), _file_prefix.c_str(),
                                         _file_afterfix.c_str(), 64), "save_data");
    }
    const std::vector<gvib200_iter_stats>& iteration_stats() const { return _stats; }

    // ---- results (gvibase/GVI-GH-GBP.h:163-167) ----
    VectorXd mean() {
        build();
        VectorXd m = VectorXd::Zero(_dim);
        gvib200_check(gvib200_get_mean(_prob, m.data()), "mean");
        return m;
    }
    BlockTridiagonal covariance() { return blocks(true); }
    BlockTridiagonal precision() { return blocks(false); }

    double cost_value() {
        build();
        double c = 0.0;
        gvib200_check(gvib200_cost(_prob, nullptr, nullptr, nullptr, &c, nullptr), "cost_value");
        return c;
    }
    template <class Mat>
    double cost_value(const VectorXd& mean, const Mat& P) {
        build();
        std::vector<double> pd, po;
        pack_precision(P, pd, po);
        double c = 0.0;
        gvib200_check(gvib200_cost(_prob, mean.data(), pd.data(), _num_states > 1 ? po.data() : nullptr, &c, nullptr), "cost_value");
        return c;
    }
    // factor costs in the order the factors were handed to the constructor
    VectorXd factor_cost_vector() {
        build();
        std::vector<double> fc(_factors.size());
        double c = 0.0;
        gvib200_check(gvib200_cost(_prob, nullptr, nullptr, nullptr, &c, fc.data()), "factor_cost_vector");
        VectorXd out = VectorXd::Zero((long)_factors.size());
        for (size_t i = 0; i < _factors.size(); ++i) out(i) = fc[(size_t)_id_of_factor[i]];
        return out;
    }
    gvib200_problem* handle() {
        build();
        return _prob;
    }

    // ---- per-factor expectations (gvibase/GVI-GH-GBP.h:367-396), in the order the factors were handed over.  Nonlinear
    // (GH) factors: the three integrals of the fused moment kernel at the current state.  Closed-form linear factors take
    // no quadrature: E_Phis() returns their expected cost E_q[phi] = T * fact_cost_value (ngd/NGDFactorizedLinear.h:
    // 122-129 stores exactly that in _E_Phi), the two matrix-valued lists hold an empty matrix for them.
    std::vector<double> E_Phis() {
        std::vector<double> e0;
        std::vector<MatrixXd> e1, e2;
        expectations(e0, e1, e2);
        return e0;
    }
    std::vector<MatrixXd> E_xMuPhis() {
        std::vector<double> e0;
        std::vector<MatrixXd> e1, e2;
        expectations(e0, e1, e2);
        return e1;
    }
    std::vector<MatrixXd> E_xMuxMuTPhis() {
        std::vector<double> e0;
        std::vector<MatrixXd> e1, e2;
        expectations(e0, e1, e2);
        return e2;
    }

    // ---- 1-D cost surface (gvibase/GVI-GH-GBP.h:402-431): Z(j, i) = cost_value(mean_i, precision_j)
    MatrixXd cost_map(const double& x_start, const double& x_end, const double& y_start, const double& y_end, const int& nmesh) {
        if (_dim != 1) throw std::logic_error("cost_map: 1-D problems only");
        build();
        const double res_x = (x_end - x_start) / nmesh, res_y = (y_end - y_start) / nmesh;
        MatrixXd Z = MatrixXd::Zero(nmesh, nmesh);
        for (int i = 0; i < nmesh; ++i) {
            const double m = x_start + i * res_x;
            for (int j = 0; j < nmesh; ++j) {
                const double prec = y_start + j * res_y;
                double c = 0.0;
                gvib200_check(gvib200_cost(_prob, &m, &prec, nullptr, &c, nullptr), "cost_map");
                Z(j, i) = c;
            }
        }
        return Z;
    }
    void save_costmap(std::string filename = "costmap.csv") {
        const MatrixXd Z = cost_map(18, 25, 0.05, 1, 40);
        gvib200_check(gvib200_csv_write(filename.c_str(), (int)Z.rows(), (int)Z.cols(), Z.data()), "save_costmap");
    }

protected:
    void expectations(std::vector<double>& e0, std::vector<MatrixXd>& e1, std::vector<MatrixXd>& e2) {
        build();
        const size_t nf = _factors.size();
        e0.assign(nf, 0.0);
        e1.assign(nf, MatrixXd());
        e2.assign(nf, MatrixXd());
        // the library returns the GH factors' moments group by group in id order
        std::vector<std::pair<int, size_t>> gh;  // (library id, caller index)
        size_t n1 = 0, n2 = 0;
        for (size_t i = 0; i < nf; ++i)
            if (!_factors[i]->is_linear()) {
                gh.emplace_back(_id_of_factor[i], i);
                n1 += (size_t)_factors[i]->_dim;
                n2 += (size_t)_factors[i]->_dim * _factors[i]->_dim;
            }
        std::sort(gh.begin(), gh.end());
        if (!gh.empty()) {
            std::vector<double> m0(gh.size()), m1(n1), m2(n2);
            gvib200_check(gvib200_moments(_prob, m0.data(), m1.data(), m2.data()), "moments");
            size_t o1 = 0, o2 = 0;
            for (size_t k = 0; k < gh.size(); ++k) {
                const size_t i = gh[k].second;
                const int dim = _factors[i]->_dim;
                e0[i] = m0[k];
                e1[i] = MatrixXd::Zero(dim, 1);
                e2[i] = MatrixXd::Zero(dim, dim);
                for (int r = 0; r < dim; ++r) e1[i](r, 0) = m1[o1 + (size_t)r];
                for (int c = 0; c < dim; ++c)
                    for (int r = 0; r < dim; ++r) e2[i](r, c) = m2[o2 + (size_t)r + (size_t)c * dim];
                o1 += (size_t)dim;
                o2 += (size_t)dim * dim;
            }
        }
        bool any_linear = false;
        for (size_t i = 0; i < nf; ++i) any_linear = any_linear || _factors[i]->is_linear();
        if (any_linear) {
            std::vector<double> fc(nf);
            double c = 0.0;
            gvib200_check(gvib200_cost(_prob, nullptr, nullptr, nullptr, &c, fc.data()), "factor costs");
            for (size_t i = 0; i < nf; ++i)
                if (_factors[i]->is_linear())
                    e0[i] = fc[(size_t)_id_of_factor[i]] * (_is_high_T ? _factors[i]->_high_temperature : _factors[i]->_temperature);
        }
    }
    template <class Mat>
    void pack_precision(const Mat& P, std::vector<double>& pd, std::vector<double>& po) const {
        const int d = _dim_state, S = _num_states;
        pd.assign((size_t)S * d * d, 0.0);
        po.assign((size_t)(S > 1 ? S - 1 : 1) * d * d, 0.0);
        for (int s = 0; s < S; ++s)
            for (int j = 0; j < d; ++j)
                for (int i = 0; i < d; ++i) {
                    pd[(size_t)s * d * d + i + j * d] = P(s * d + i, s * d + j);
                    if (s + 1 < S) po[(size_t)s * d * d + i + j * d] = P(s * d + i, (s + 1) * d + j);
                }
    }
    BlockTridiagonal blocks(bool cov) {
        build();
        BlockTridiagonal b;
        b.num_states = _num_states;
        b.dim_state = _dim_state;
        const size_t dd = (size_t)_dim_state * _dim_state;
        b.diag.assign((size_t)_num_states * dd, 0.0);
        b.off.assign((size_t)(_num_states > 1 ? _num_states - 1 : 1) * dd, 0.0);
        gvib200_check((cov ? gvib200_get_cov_blocks : gvib200_get_prec_blocks)(_prob, b.diag.data(), b.off.data()), "blocks");
        if (_num_states == 1) b.off.clear();
        return b;
    }
    void push_state() {
        if (!_prob) return;  // applied when the problem is built
        if (_mu0.empty()) _mu0.assign((size_t)_dim, 0.0);
        gvib200_check(gvib200_set_state(_prob, _mu0.data(), _pd.data(), _num_states > 1 ? _po.data() : nullptr), "set_state");
        _has_state = true;
    }

    // Bucket the factors by (type, shape, cost parameters): every bucket becomes ONE batched group of the C-ABI, whatever
    // order the caller interleaved them in; _id_of_factor maps the caller's order to the library's factor ids.
    void build() {
        if (_prob) return;
        gvib200_ctx* ctx = default_context();
        gvib200_check(gvib200_problem_create(ctx, _num_states, _dim_state, &_prob), "problem_create");
        if (_prox) gvib200_check(gvib200_problem_set_option(_prob, "prox", 1), "set_option prox");
        struct GhKey {
            DeviceCostSpec cost;
            int dim, deg;
            bool operator<(const GhKey& o) const {
                if (dim != o.dim) return dim < o.dim;
                if (deg != o.deg) return deg < o.deg;
                return cost < o.cost;
            }
        };
        struct LinKey {
            int dim, m, kdim;
            bool operator<(const LinKey& o) const {
                if (dim != o.dim) return dim < o.dim;
                if (m != o.m) return m < o.m;
                return kdim < o.kdim;
            }
        };
        std::map<GhKey, std::vector<size_t>> gh;
        std::map<LinKey, std::vector<size_t>> lin;
        for (size_t i = 0; i < _factors.size(); ++i) {
            GVIFactorizedBase* f = _factors[i].get();
            if (!f->is_linear()) {
                gh[GhKey{*f->cost_spec(), f->_dim, f->gh_degree()}].push_back(i);
            } else {
                lin[LinKey{f->_dim, (int)f->lin_Lambda()->rows(), (int)f->lin_Psi()->cols()}].push_back(i);
            }
        }
        _id_of_factor.assign(_factors.size(), -1);
        for (auto& kv : gh) {
            const auto& idx = kv.second;
            const int n = (int)idx.size();
            std::vector<int32_t> start((size_t)n);
            std::vector<double> T((size_t)n), Th((size_t)n);
            for (int k = 0; k < n; ++k) {
                GVIFactorizedBase* f = _factors[idx[(size_t)k]].get();
                start[(size_t)k] = f->_start_index;
                T[(size_t)k] = f->_temperature;
                Th[(size_t)k] = f->_high_temperature;
            }
            const DeviceCostSpec& c = kv.first.cost;
            upload_cost_fields(_prob, c);
            int first = 0;
            gvib200_check(gvib200_add_gh_factors(_prob, c.kind, kv.first.dim, kv.first.deg, n, start.data(), T.data(), Th.data(),
                                                 c.params.data(), c.params.size(), &first), "add_gh_factors");
            for (int k = 0; k < n; ++k) _id_of_factor[idx[(size_t)k]] = first + k;
        }
        for (auto& kv : lin) {
            const auto& idx = kv.second;
            const int n = (int)idx.size(), dim = kv.first.dim, m = kv.first.m, kdim = kv.first.kdim;
            std::vector<int32_t> start((size_t)n);
            std::vector<double> L((size_t)n * m * dim), P((size_t)n * m * kdim), mt((size_t)n * kdim), K((size_t)n * m * m),
                C((size_t)n), T((size_t)n), Th((size_t)n);
            for (int k = 0; k < n; ++k) {
                GVIFactorizedBase* f = _factors[idx[(size_t)k]].get();
                const MatrixXd &La = *f->lin_Lambda(), &Ps = *f->lin_Psi(), &Ki = *f->lin_Kinv();
                const VectorXd& mu = *f->lin_mu();
                start[(size_t)k] = f->_start_index;
                T[(size_t)k] = f->_temperature;
                Th[(size_t)k] = f->_high_temperature;
                C[(size_t)k] = f->lin_constant();
                for (int j = 0; j < dim; ++j)
                    for (int i = 0; i < m; ++i) L[((size_t)k * dim + j) * m + i] = La(i, j);
                for (int j = 0; j < kdim; ++j)
                    for (int i = 0; i < m; ++i) P[((size_t)k * kdim + j) * m + i] = Ps(i, j);
                for (int j = 0; j < kdim; ++j) mt[(size_t)k * kdim + j] = mu(j);
                for (int j = 0; j < m; ++j)
                    for (int i = 0; i < m; ++i) K[((size_t)k * m + j) * m + i] = Ki(i, j);
            }
            int first = 0;
            gvib200_check(gvib200_add_linear_factors(_prob, dim, m, kdim, n, start.data(), L.data(), P.data(), mt.data(), K.data(),
                                                     C.data(), T.data(), Th.data(), &first), "add_linear_factors");
            for (int k = 0; k < n; ++k) _id_of_factor[idx[(size_t)k]] = first + k;
        }
        gvib200_check(gvib200_problem_finalize(_prob), "finalize");
        if (!_pd.empty()) push_state();
    }

    void run_loop(bool prox, bool verbose) {
        build();
        _stats.assign((size_t)_niters, gvib200_iter_stats());
        int done = 0;
        if (_record) {
            const size_t dd = (size_t)_dim_state * _dim_state, n = (size_t)_niters;
            _t_mean.assign(n * _dim, 0.0);
            _t_cov.assign(n * _num_states * dd, 0.0);
            _t_prec.assign(n * _num_states * dd, 0.0);
            _t_cov_off.assign(n * (size_t)(_num_states > 1 ? _num_states - 1 : 1) * dd, 0.0);
            _t_prec_off.assign(_t_cov_off.size(), 0.0);
            _t_cost.assign(n, 0.0);
            _t_fac.assign(n * _factors.size(), 0.0);
            const bool joint = _dim <= 64;
            _trace = gvib200_trace{_niters, 0, _t_mean.data(), _t_cov.data(), _t_prec.data(), joint ? _t_cov_off.data() : nullptr,
                                   joint ? _t_prec_off.data() : nullptr, _t_cost.data(), _t_fac.data()};
            gvib200_check(gvib200_optimize_traced(_prob, &_opts, _niters, prox ? 1 : 0, _stats.data(), &done, &_trace), "optimize");
            save_data(verbose);
        } else if (prox) {
            gvib200_check(gvib200_prox_optimize(_prob, &_opts, _niters, _stats.data(), &done), "prox optimize");
        } else {
            gvib200_check(gvib200_optimize(_prob, &_opts, _niters, _stats.data(), &done, nullptr, nullptr), "optimize");
        }
        _stats.resize((size_t)done);
        for (auto& st : _stats)
            if (st.switched_high_T) {
                _is_high_T = true;
                _temperature = _high_temperature;
            }
        if (verbose)
            for (int i = 0; i < done; ++i) std::printf("iteration %d cost %.15g\n", i, _stats[(size_t)i].cost);
    }

    int _dim_state, _num_states, _dim, _niters;
    bool _record = false;
    std::string _file_prefix, _file_afterfix;
    gvib200_trace _trace{};
    std::vector<double> _t_mean, _t_cov, _t_prec, _t_cov_off, _t_prec_off, _t_cost, _t_fac;
    double _temperature, _high_temperature;
    double _step_size = 0.9, _stop_err = 1e-5, _initial_precision_factor = 100.0;  // gvibase/GVI-GH-GBP.h:54,93
    bool _is_high_T = false;
    std::vector<std::shared_ptr<GVIFactorizedBase>> _factors;
    std::vector<int> _id_of_factor;
    gvib200_opts _opts;
    gvib200_problem* _prob = nullptr;
    bool _has_state = false;
    bool _prox = false;  // set by ProxGVIGH before the problem is built
    std::vector<double> _mu0, _pd, _po;
    std::vector<gvib200_iter_stats> _stats;
};

// NGDGH (ngd/NGD-GH.h:28-110)
template <class FactorizedOptimizer>
class NGDGH : public GVIGH<FactorizedOptimizer> {
    using Base = GVIGH<FactorizedOptimizer>;

public:
    NGDGH(const std::vector<std::shared_ptr<FactorizedOptimizer>>& vec_fact_optimizers, int dim_state, int num_states,
          int niterations = 5, double temperature = 1.0, double high_temperature = 100.0)
        : Base(vec_fact_optimizers, dim_state, num_states, niterations, temperature, high_temperature) {}

    // compute_gradients (ngd/NGD-GH-impl.h:20-63): (dmu, dprecision)
    std::pair<VectorXd, BlockTridiagonal> compute_gradients() {
        Base::build();
        VectorXd dmu = VectorXd::Zero(Base::_dim);
        BlockTridiagonal dp;
        dp.num_states = Base::_num_states;
        dp.dim_state = Base::_dim_state;
        const size_t dd = (size_t)Base::_dim_state * Base::_dim_state;
        dp.diag.assign((size_t)Base::_num_states * dd, 0.0);
        dp.off.assign((size_t)(Base::_num_states > 1 ? Base::_num_states - 1 : 1) * dd, 0.0);
        gvib200_check(gvib200_gradients(Base::_prob, dmu.data(), dp.diag.data(), dp.off.data()), "compute_gradients");
        return {dmu, dp};
    }
    VectorXd Vdmu() {  // ngd/NGD-GH.h:90
        VectorXd v = VectorXd::Zero(Base::_dim);
        gvib200_check(gvib200_get_V(Base::_prob, v.data(), nullptr, nullptr), "Vdmu");
        return v;
    }
    BlockTridiagonal Vddmu() {  // ngd/NGD-GH.h:92
        BlockTridiagonal b;
        b.num_states = Base::_num_states;
        b.dim_state = Base::_dim_state;
        const size_t dd = (size_t)Base::_dim_state * Base::_dim_state;
        b.diag.assign((size_t)Base::_num_states * dd, 0.0);
        b.off.assign((size_t)(Base::_num_states > 1 ? Base::_num_states - 1 : 1) * dd, 0.0);
        gvib200_check(gvib200_get_V(Base::_prob, nullptr, b.diag.data(), b.off.data()), "Vddmu");
        return b;
    }
};

// ------------------------------------------------------------------------------------------------ Prox-GVI
// proxgd/ProxGVIFactorizedBaseGH.h, ProxGVIFactorizedLinear.h: the factor classes carry the same data as their NGD
// counterparts; the Bures-Wasserstein JKO step (BW_JKO) runs on the device.
template <class CostClass>
using ProxGVIFactorizedBaseGH = NGDFactorizedBaseGH<CostClass>;
template <class Factor>
using ProxGVIFactorizedLinear = NGDFactorizedLinear<Factor>;

// ProxGVIGH (proxgd/ProxGVI-GH.h, -impl.h:124-205)
template <class FactorizedOptimizer>
class ProxGVIGH : public GVIGH<FactorizedOptimizer> {
    using Base = GVIGH<FactorizedOptimizer>;

public:
    ProxGVIGH(const std::vector<std::shared_ptr<FactorizedOptimizer>>& vec_fact_optimizers, int dim_state, int num_states,
              int niterations = 5, double temperature = 1.0, double high_temperature = 100.0)
        : Base(vec_fact_optimizers, dim_state, num_states, niterations, temperature, high_temperature) {
        Base::_prox = true;
    }
    void optimize(std::optional<bool> verbose = std::nullopt) { Base::run_loop(true, verbose.value_or(false)); }
};

// SparseGaussHermite (quadrature/SparseGaussHermite.h:38-277) for a DEVICE cost class: the three integrals the factor
// optimizers take -- E[phi], E[(x-mu) phi], E[(x-mu)(x-mu)^T phi] -- evaluated by the fused moment kernel.
template <class CostClass>
class SparseGaussHermite {
public:
    SparseGaussHermite(int deg, int dim, const VectorXd& mean, const MatrixXd& P, const CostClass& cost_class)
        : _deg(deg), _dim(dim), _mean(mean), _P(P), _cost(DeviceCostTraits<CostClass>::spec(cost_class)) {}
    ~SparseGaussHermite() {
        if (_prob) gvib200_problem_destroy(_prob);
    }
    void update_mean(const VectorXd& mean) { _mean = mean; _dirty = true; }
    void update_P(const MatrixXd& P) { _P = P; _dirty = true; }
    void set_polynomial_deg(int deg) { _deg = deg; reset(); }
    // quadrature/SparseGaussHermite.h:254-272
    void update_dimension(int dim) { _dim = dim; reset(); }
    void update_parameters(int deg, int dim, const VectorXd& mean, const MatrixXd& P) {
        _deg = deg;
        _dim = dim;
        _mean = mean;
        _P = P;
        reset();
    }
    VectorXd mean() const { return _mean; }
    // sigma points X = Z sqrtm(P)^T + 1 mean^T (quadrature/SparseGaussHermite.h:231-243,275): the device kernel forms them
    // on the fly and never stores them; this accessor rebuilds them on the host for inspection
    MatrixXd sigmapts() const {
        const MatrixXd Z = zeromeanpts();
        const MatrixXd S = detail::sym_sqrt(_P);
        MatrixXd X = Z * S.transpose();
        for (long i = 0; i < X.rows(); ++i)
            for (int c = 0; c < _dim; ++c) X(i, c) += _mean(c);
        return X;
    }
    // nodes of the rule (zero-mean sigma points) and weights, straight from the table generator
    MatrixXd zeromeanpts() const {
        const int n = gvib200_table_size(_dim, _deg);
        gvib200_check(n, "table_size");
        std::vector<double> nodes((size_t)n * _dim), w((size_t)n);
        gvib200_check(gvib200_table_generate(_dim, _deg, nodes.data(), w.data(), n), "table_generate");
        MatrixXd Z = MatrixXd::Zero(n, _dim);
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < _dim; ++c) Z(i, c) = nodes[(size_t)i * _dim + c];
        return Z;
    }
    VectorXd weights() const {
        const int n = gvib200_table_size(_dim, _deg);
        gvib200_check(n, "table_size");
        std::vector<double> nodes((size_t)n * _dim), w((size_t)n);
        gvib200_check(gvib200_table_generate(_dim, _deg, nodes.data(), w.data(), n), "table_generate");
        VectorXd out = VectorXd::Zero(n);
        for (int i = 0; i < n; ++i) out(i) = w[(size_t)i];
        return out;
    }
    struct Moments {
        double E_phi;
        VectorXd E_xmu_phi;
        MatrixXd E_xmuxmuT_phi;
    };
    // Integrate (quadrature/SparseGaussHermite.h:197-221) of phi, (x-mu) phi and (x-mu)(x-mu)^T phi at (mean, P)
    Moments Integrate() {
        ensure();
        Moments m;
        m.E_xmu_phi = VectorXd::Zero(_dim);
        m.E_xmuxmuT_phi = MatrixXd::Zero(_dim, _dim);
        gvib200_check(gvib200_moments(_prob, &m.E_phi, m.E_xmu_phi.data(), m.E_xmuxmuT_phi.data()), "Integrate");
        return m;
    }

private:
    void reset() {
        if (_prob) gvib200_problem_destroy(_prob);
        _prob = nullptr;
        _dirty = true;
    }
    void ensure() {
        if (!_prob) {
            gvib200_check(gvib200_problem_create(default_context(), 1, _dim, &_prob), "problem_create");
            upload_cost_fields(_prob, _cost);
            const int32_t start = 0;
            gvib200_check(gvib200_add_gh_factors(_prob, _cost.kind, _dim, _deg, 1, &start, nullptr, nullptr, _cost.params.data(),
                                                 _cost.params.size(), nullptr), "add_gh_factors");
            gvib200_check(gvib200_problem_finalize(_prob), "finalize");
        }
        if (_dirty) {
            const MatrixXd prec = _P.inverse();  // the library state is a precision; P is the covariance (:231-233)
            gvib200_check(gvib200_set_state(_prob, _mean.data(), prec.data(), nullptr), "set_state");
            _dirty = false;
        }
    }
    int _deg, _dim;
    VectorXd _mean;
    MatrixXd _P;
    DeviceCostSpec _cost;
    gvib200_problem* _prob = nullptr;
    bool _dirty = true;
};

}  // namespace gvi
