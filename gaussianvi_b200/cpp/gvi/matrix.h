// Dense matrix types of the facade: Eigen's when <Eigen/Dense> is available (the reference is Eigen-facing), otherwise
// a minimal column-major stand-in with the subset of the Eigen API the facade and its callers use.  Either way the
// facade only relies on rows(), cols(), size(), data() and operator().
#pragma once
#if defined(GVIB200_USE_EIGEN) || (__has_include(<Eigen/Dense>) && !defined(GVIB200_NO_EIGEN))
#include <Eigen/Dense>
namespace gvi {
using MatrixXd = Eigen::MatrixXd;
using VectorXd = Eigen::VectorXd;
}  // namespace gvi
#else
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <vector>
namespace gvi {
class MatrixXd {
public:
    MatrixXd() : r_(0), c_(0) {}
    MatrixXd(long r, long c) : r_(r), c_(c), a_((size_t)(r * c), 0.0) {}
    static MatrixXd Zero(long r, long c) { return MatrixXd(r, c); }
    static MatrixXd Constant(long r, long c, double v) {
        MatrixXd m(r, c);
        for (auto& x : m.a_) x = v;
        return m;
    }
    static MatrixXd Identity(long r, long c) {
        MatrixXd m(r, c);
        for (long i = 0; i < (r < c ? r : c); ++i) m(i, i) = 1.0;
        return m;
    }
    long rows() const { return r_; }
    long cols() const { return c_; }
    long size() const { return r_ * c_; }
    double* data() { return a_.data(); }
    const double* data() const { return a_.data(); }
    double& operator()(long i, long j) { return a_[(size_t)(i + j * r_)]; }
    double operator()(long i, long j) const { return a_[(size_t)(i + j * r_)]; }
    void setZero() {
        for (auto& x : a_) x = 0.0;
    }
    MatrixXd transpose() const {
        MatrixXd t(c_, r_);
        for (long j = 0; j < c_; ++j)
            for (long i = 0; i < r_; ++i) t(j, i) = (*this)(i, j);
        return t;
    }
    MatrixXd block(long i0, long j0, long nr, long nc) const {
        MatrixXd b(nr, nc);
        for (long j = 0; j < nc; ++j)
            for (long i = 0; i < nr; ++i) b(i, j) = (*this)(i0 + i, j0 + j);
        return b;
    }
    void setBlock(long i0, long j0, const MatrixXd& b) {
        for (long j = 0; j < b.cols(); ++j)
            for (long i = 0; i < b.rows(); ++i) (*this)(i0 + i, j0 + j) = b(i, j);
    }
    MatrixXd inverse() const {  // Gauss-Jordan with partial pivoting
        if (r_ != c_) throw std::invalid_argument("inverse: not square");
        const long n = r_;
        MatrixXd A(*this), I = Identity(n, n);
        for (long c = 0; c < n; ++c) {
            long p = c;
            for (long r = c + 1; r < n; ++r)
                if (std::fabs(A(r, c)) > std::fabs(A(p, c))) p = r;
            if (A(p, c) == 0.0) throw std::runtime_error("inverse: singular");
            for (long j = 0; j < n; ++j) {
                std::swap(A(c, j), A(p, j));
                std::swap(I(c, j), I(p, j));
            }
            const double inv = 1.0 / A(c, c);
            for (long j = 0; j < n; ++j) {
                A(c, j) *= inv;
                I(c, j) *= inv;
            }
            for (long r = 0; r < n; ++r) {
                if (r == c) continue;
                const double f = A(r, c);
                for (long j = 0; j < n; ++j) {
                    A(r, j) -= f * A(c, j);
                    I(r, j) -= f * I(c, j);
                }
            }
        }
        return I;
    }
    friend MatrixXd operator*(const MatrixXd& A, const MatrixXd& B) {
        if (A.c_ != B.r_) throw std::invalid_argument("matrix product: shape mismatch");
        MatrixXd C(A.r_, B.c_);
        for (long j = 0; j < B.c_; ++j)
            for (long k = 0; k < A.c_; ++k)
                for (long i = 0; i < A.r_; ++i) C(i, j) += A(i, k) * B(k, j);
        return C;
    }
    friend MatrixXd operator*(double s, const MatrixXd& A) {
        MatrixXd C(A);
        for (auto& x : C.a_) x *= s;
        return C;
    }
    friend MatrixXd operator*(const MatrixXd& A, double s) { return s * A; }
    friend MatrixXd operator/(const MatrixXd& A, double s) { return (1.0 / s) * A; }
    friend MatrixXd operator+(const MatrixXd& A, const MatrixXd& B) {
        MatrixXd C(A);
        for (size_t i = 0; i < C.a_.size(); ++i) C.a_[i] += B.a_[i];
        return C;
    }
    friend MatrixXd operator-(const MatrixXd& A, const MatrixXd& B) {
        MatrixXd C(A);
        for (size_t i = 0; i < C.a_.size(); ++i) C.a_[i] -= B.a_[i];
        return C;
    }
    MatrixXd operator-() const { return -1.0 * (*this); }

protected:
    long r_, c_;
    std::vector<double> a_;
};

class VectorXd : public MatrixXd {
public:
    VectorXd() : MatrixXd() {}
    explicit VectorXd(long n) : MatrixXd(n, 1) {}
    VectorXd(const MatrixXd& m) : MatrixXd(m) {
        if (m.cols() != 1 && m.size() != 0) throw std::invalid_argument("VectorXd: not a column");
    }
    static VectorXd Zero(long n) { return VectorXd(n); }
    static VectorXd Constant(long n, double v) { return VectorXd(MatrixXd::Constant(n, 1, v)); }
    double& operator()(long i) { return a_[(size_t)i]; }
    double operator()(long i) const { return a_[(size_t)i]; }
    double& operator[](long i) { return a_[(size_t)i]; }
    double operator[](long i) const { return a_[(size_t)i]; }
    VectorXd segment(long i0, long n) const { return VectorXd(block(i0, 0, n, 1)); }
};
}  // namespace gvi
#endif
