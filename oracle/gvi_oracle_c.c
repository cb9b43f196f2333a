/* CPU oracle, C restatement (C99 + OpenMP) of the factorized NGD-GVI path of hzyu17/GaussianVI.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Nothing under gaussianvi_b200/ links or loads this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.  It exists because the NumPy oracle
 * (oracle/gvi_oracle.py) is too slow to serve as a CPU baseline and the reference itself cannot be compiled here (no
 * Eigen / GSL / MATLAB runtime: SURVEY.md 0.10).  It is validated against the NumPy oracle and, through it, against the
 * reference's golden vectors (tests/test_oracle_c.py).  Parity status: pinned by the reference's fixtures for the 1-D
 * trace, the tables and the quadrature KATs; for multi-state chains "parity unpinned" (no reference fixture exists).
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).  Matrices are
 * column-major (Eigen's default); block-tridiagonal matrices are diag[S][d*d], off[S-1][d*d] (block (i, i+1)).
 *
 * Two schedules of one NGD iteration:
 *   schedule 0 ("reference"): exactly what GVIGH::optimize executes (gvibase/GVI-GH-GBP-impl.h:33-130):
 *       cost_value() + factor_cost_vector() (two inverse_GBP + two cost sweeps), compute_gradients() with three
 *       separate integrals per factor (psi re-evaluated three times, ngd/NGDFactorizedBaseGH.h:53-74) and the O(dim^4)
 *       fourth-moment loop of the linear factors (ngd/NGDFactorizedLinear.h:107-119), one full cost_value per
 *       line-search trial, set_precision() (a fourth inverse_GBP) on acceptance;
 *   schedule 1 ("lean"): the same arithmetic with every redundant re-evaluation removed (one fused moment sweep, the
 *       accepted trial's covariance / costs carried over).
 * The mean step is solved by a direct block Cholesky (oracle policy); the reference calls Eigen CG.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 24
#define MAX_LIN_GROUPS 4

enum { ORC_COST_STEREO_1D = 1, ORC_COST_PLANAR_HINGE = 2, ORC_COST_QUADRATIC = 5 };

typedef struct {
    int n, dim, m, kdim;
    const int* start;
    const double *Lambda, *Psi, *mu_t, *Kinv, *C, *T; /* per factor, column-major */
} orc_lin_group;

typedef struct {
    int S, d;
    /* one group of sparse-GH factors */
    int n_gh, gh_dim, n_nodes, cost_kind;
    const int* gh_start;
    const double* Z; /* [n_nodes][gh_dim] row-major */
    const double* w;
    const double* gh_T;
    double cp[8]; /* stereo: mu_p, f, b, sig_r_sq, sig_p_sq, y_offset; hinge: sigma, eps, radius; quadratic: c */
    int rows, cols;
    double ox, oy, cell;
    const double* sdf; /* column-major rows x cols */
    int n_lin_groups;
    orc_lin_group lin[MAX_LIN_GROUPS];
} orc_problem;

typedef struct {
    double cost, new_cost, step;
    int n_backtrack, accepted, status;
    int n_psi_sweeps; /* quadrature sweeps over all GH factors executed (each evaluates psi at every node) */
    int n_inversions; /* block-tridiagonal selected inversions executed */
} orc_stats;

/* ---------------------------------------------------------------------------------------------- small dense */
/* Gauss-Jordan inverse with partial pivoting (Eigen's dense inverse(): gvibase/GVIFactorizedBase.h:113,
   GVI-GH-GBP-impl.h:292,300,337).  Returns 0 on success. */
static int mat_inv(int n, const double* A, double* Ai) {
    double M[MAXN * 2 * MAXN];
    const int w = 2 * n;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            M[i * w + j] = A[i + j * n];
            M[i * w + n + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < n; ++c) {
        int p = c;
        double best = fabs(M[c * w + c]);
        for (int r = c + 1; r < n; ++r)
            if (fabs(M[r * w + c]) > best) best = fabs(M[r * w + c]), p = r;
        if (best == 0.0) return -1;
        if (p != c)
            for (int j = 0; j < w; ++j) {
                double t = M[c * w + j];
                M[c * w + j] = M[p * w + j];
                M[p * w + j] = t;
            }
        const double inv = 1.0 / M[c * w + c];
        for (int j = 0; j < w; ++j) M[c * w + j] *= inv;
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = M[r * w + c];
            if (f == 0.0) continue;
            for (int j = 0; j < w; ++j) M[r * w + j] -= f * M[c * w + j];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Ai[i + j * n] = M[i * w + n + j];
    return 0;
}

/* C = op(A) * op(B), all n x n column-major; ta / tb: transpose flags */
static void mm(int n, const double* A, int ta, const double* B, int tb, double* C) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += (ta ? A[k + i * n] : A[i + k * n]) * (tb ? B[j + k * n] : B[k + j * n]);
            C[i + j * n] = s;
        }
}

/* in-place lower Cholesky; returns 0 if SPD.  *logdet += 2 sum log L_ii */
static int chol(int n, double* A, double* logdet) {
    for (int j = 0; j < n; ++j) {
        double s = A[j + j * n];
        for (int k = 0; k < j; ++k) s -= A[j + k * n] * A[j + k * n];
        if (!(s > 0.0)) return -1;
        const double l = sqrt(s);
        A[j + j * n] = l;
        if (logdet) *logdet += 2.0 * log(l);
        for (int i = j + 1; i < n; ++i) {
            double t = A[i + j * n];
            for (int k = 0; k < j; ++k) t -= A[i + k * n] * A[j + k * n];
            A[i + j * n] = t / l;
        }
        for (int i = 0; i < j; ++i) A[i + j * n] = 0.0;
    }
    return 0;
}

/* symmetric PSD square root S = V sqrt(D) V^T: SelfAdjointEigenSolver::operatorSqrt()
   (quadrature/SparseGaussHermite.h:231-233), here by cyclic Jacobi */
static void sqrtm_psd(int n, const double* Sig, double* S) {
    double A[MAXN * MAXN], V[MAXN * MAXN];
    memcpy(A, Sig, sizeof(double) * n * n);
    for (int i = 0; i < n * n; ++i) V[i] = 0.0;
    for (int i = 0; i < n; ++i) V[i + i * n] = 1.0;
    for (int sweep = 0; sweep < 60 && n > 1; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int j = 0; j < n; ++j) {
            diag += fabs(A[j + j * n]);
            for (int i = 0; i < j; ++i) off += fabs(A[i + j * n]);
        }
        if (off <= 1e-300 || off <= 1e-22 * diag) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p + q * n];
                if (apq == 0.0) continue;
                const double app = A[p + p * n], aqq = A[q + q * n];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) { /* A <- A J */
                    const double akp = A[k + p * n], akq = A[k + q * n];
                    A[k + p * n] = c * akp - s * akq;
                    A[k + q * n] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) { /* A <- J^T A */
                    const double apk = A[p + k * n], aqk = A[q + k * n];
                    A[p + k * n] = c * apk - s * aqk;
                    A[q + k * n] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[k + p * n], vkq = V[k + q * n];
                    V[k + p * n] = c * vkp - s * vkq;
                    V[k + q * n] = s * vkp + c * vkq;
                }
            }
    }
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += V[i + k * n] * sqrt(A[k + k * n]) * V[j + k * n];
            S[i + j * n] = s;
        }
}

/* ---------------------------------------------------------------------------------------------- chain */
/* GVIGH::inverse_GBP + calculate_factor_message (gvibase/GVI-GH-GBP-impl.h:245-342): forward / backward messages,
   then one 2d x 2d inverse per consecutive pair; later pairs overwrite the shared diagonal block (:296-302). */
int orc_inverse_gbp(int S, int d, const double* D, const double* O, double* cD, double* cO) {
    const int dd = d * d;
    if (S == 1) return mat_inv(d, D, cD);
    double* F = (double*)calloc((size_t)S * dd, sizeof(double));
    double* B = (double*)calloc((size_t)S * dd, sizeof(double));
    int rc = 0;
    double T[MAXN * MAXN], Ti[MAXN * MAXN], U[MAXN * MAXN];
    for (int i = 0; i < S - 1 && !rc; ++i) {
        for (int e = 0; e < dd; ++e) T[e] = D[(size_t)i * dd + e] + F[(size_t)i * dd + e];
        rc |= mat_inv(d, T, Ti);
        mm(d, Ti, 0, O + (size_t)i * dd, 0, U);               /* (D_i + F_i)^-1 O_i */
        mm(d, O + (size_t)i * dd, 1, U, 0, T);                /* O_i^T ...          */
        for (int e = 0; e < dd; ++e) F[(size_t)(i + 1) * dd + e] = -T[e];
        const int j = S - 1 - i;
        for (int e = 0; e < dd; ++e) T[e] = D[(size_t)j * dd + e] + B[(size_t)j * dd + e];
        rc |= mat_inv(d, T, Ti);
        mm(d, Ti, 0, O + (size_t)(j - 1) * dd, 1, U);         /* (D_j + B_j)^-1 O_{j-1}^T */
        mm(d, O + (size_t)(j - 1) * dd, 0, U, 0, T);
        for (int e = 0; e < dd; ++e) B[(size_t)(j - 1) * dd + e] = -T[e];
    }
    const int n2 = 2 * d;
    for (int i = 0; i < S - 1 && !rc; ++i) {
        double J[MAXN * MAXN], Ji[MAXN * MAXN];
        for (int c = 0; c < d; ++c)
            for (int r = 0; r < d; ++r) {
                J[r + c * n2] = D[(size_t)i * dd + r + c * d] + F[(size_t)i * dd + r + c * d];
                J[(d + r) + (d + c) * n2] = D[(size_t)(i + 1) * dd + r + c * d] + B[(size_t)(i + 1) * dd + r + c * d];
                J[r + (d + c) * n2] = O[(size_t)i * dd + r + c * d];
                J[(d + c) + r * n2] = O[(size_t)i * dd + r + c * d];
            }
        rc |= mat_inv(n2, J, Ji);
        for (int c = 0; c < d; ++c)
            for (int r = 0; r < d; ++r) {
                cD[(size_t)i * dd + r + c * d] = Ji[r + c * n2];
                cD[(size_t)(i + 1) * dd + r + c * d] = Ji[(d + r) + (d + c) * n2];
                cO[(size_t)i * dd + r + c * d] = Ji[r + (d + c) * n2];
            }
    }
    free(F);
    free(B);
    return rc;
}

/* block Cholesky of a block-tridiagonal SPD matrix: log det (1/2 sum log D of SparseLDLT,
   gvibase/GVI-GH-GBP-impl.h:234-238) and, with rhs, the direct solve replacing ConjugateGradient
   (ngd/NGD-GH-impl.h:59-60).  Returns 0 if SPD. */
int orc_block_solve(int S, int d, const double* D, const double* O, const double* rhs, double* x, double* logdet) {
    const int dd = d * d;
    double* Ld = (double*)malloc(sizeof(double) * (size_t)S * dd);
    double* Lo = (double*)malloc(sizeof(double) * (size_t)(S > 1 ? S - 1 : 1) * dd);
    double* y = (double*)malloc(sizeof(double) * (size_t)S * d);
    double ld = 0.0;
    int rc = 0;
    double Dt[MAXN * MAXN];
    memcpy(Dt, D, sizeof(double) * dd);
    for (int i = 0; i < S; ++i) {
        double* L = Ld + (size_t)i * dd;
        memcpy(L, Dt, sizeof(double) * dd);
        if (chol(d, L, &ld)) { rc = -4; break; }
        if (i < S - 1) {
            /* Lo_i = O_i^T L^-T: solve L X = O_i (forward), Lo = X^T */
            double X[MAXN * MAXN];
            const double* Oi = O + (size_t)i * dd;
            for (int c = 0; c < d; ++c)
                for (int r = 0; r < d; ++r) {
                    double t = Oi[r + c * d];
                    for (int k = 0; k < r; ++k) t -= L[r + k * d] * X[k + c * d];
                    X[r + c * d] = t / L[r + r * d];
                }
            double* lo = Lo + (size_t)i * dd;
            for (int c = 0; c < d; ++c)
                for (int r = 0; r < d; ++r) lo[r + c * d] = X[c + r * d];
            for (int c = 0; c < d; ++c)
                for (int r = 0; r < d; ++r) {
                    double s = 0.0;
                    for (int k = 0; k < d; ++k) s += lo[r + k * d] * lo[c + k * d];
                    Dt[r + c * d] = D[(size_t)(i + 1) * dd + r + c * d] - s;
                }
        }
    }
    if (!rc && rhs && x) {
        for (int i = 0; i < S; ++i) {
            const double* L = Ld + (size_t)i * dd;
            double t[MAXN];
            for (int r = 0; r < d; ++r) {
                t[r] = rhs[(size_t)i * d + r];
                if (i > 0) {
                    const double* lo = Lo + (size_t)(i - 1) * dd;
                    for (int k = 0; k < d; ++k) t[r] -= lo[r + k * d] * y[(size_t)(i - 1) * d + k];
                }
            }
            for (int r = 0; r < d; ++r) {
                double s = t[r];
                for (int k = 0; k < r; ++k) s -= L[r + k * d] * y[(size_t)i * d + k];
                y[(size_t)i * d + r] = s / L[r + r * d];
            }
        }
        for (int i = S - 1; i >= 0; --i) {
            const double* L = Ld + (size_t)i * dd;
            double t[MAXN];
            for (int r = 0; r < d; ++r) {
                t[r] = y[(size_t)i * d + r];
                if (i < S - 1) {
                    const double* lo = Lo + (size_t)i * dd;
                    for (int k = 0; k < d; ++k) t[r] -= lo[k + r * d] * x[(size_t)(i + 1) * d + k];
                }
            }
            for (int r = d - 1; r >= 0; --r) {
                double s = t[r];
                for (int k = r + 1; k < d; ++k) s -= L[k + r * d] * x[(size_t)i * d + k];
                x[(size_t)i * d + r] = s / L[r + r * d];
            }
        }
    }
    if (logdet) *logdet = ld;
    free(Ld);
    free(Lo);
    free(y);
    return rc;
}

/* ---------------------------------------------------------------------------------------------- cost functions */
/* PlanarSDF::getSignedDistance (helpers/CudaOperation.h:51-103,123-125): clamp to the field, bilinear lookup */
static double sdf_lookup(const orc_problem* p, double x, double y) {
    const double xmax = p->ox + (p->cols - 1.0) * p->cell, ymax = p->oy + (p->rows - 1.0) * p->cell;
    const double xr = x < p->ox ? p->ox : (x > xmax ? xmax : x);
    const double yr = y < p->oy ? p->oy : (y > ymax ? ymax : y);
    const double col = (xr - p->ox) / p->cell, row = (yr - p->oy) / p->cell;
    const double lr = floor(row), lc = floor(col), hr = lr + 1.0, hc = lc + 1.0;
    const int lri = (int)lr, lci = (int)lc;
    int hri = (int)hr, hci = (int)hc;
    if (hri > p->rows - 1) hri = p->rows - 1; /* the reference reads one past the edge with weight exactly 0 */
    if (hci > p->cols - 1) hci = p->cols - 1;
    const double* D = p->sdf;
    const size_t R = (size_t)p->rows;
    return (hr - row) * (hc - col) * D[lri + lci * R] + (row - lr) * (hc - col) * D[hri + lci * R] +
           (hr - row) * (col - lc) * D[lri + hci * R] + (row - lr) * (col - lc) * D[hri + hci * R];
}

static double psi_eval(const orc_problem* p, const double* x) {
    switch (p->cost_kind) {
        case ORC_COST_STEREO_1D: { /* src/1d_example.cpp:25-35 */
            const double mu_p = p->cp[0], f = p->cp[1], b = p->cp[2], sig_r_sq = p->cp[3], sig_p_sq = p->cp[4];
            const double y = f * b / mu_p + p->cp[5];
            return (x[0] - mu_p) * (x[0] - mu_p) / sig_p_sq / 2 + (y - f * b / x[0]) * (y - f * b / x[0]) / sig_r_sq / 2;
        }
        case ORC_COST_PLANAR_HINGE: { /* cost_obstacle_planar, helpers/CudaOperation.h:491-508 */
            const double sd = sdf_lookup(p, x[0], x[1]);
            const double thr = p->cp[1] + p->cp[2];
            const double err = sd > thr ? 0.0 : (thr - sd);
            return err * err * p->cp[0];
        }
        case ORC_COST_QUADRATIC: {
            double q = 0.0;
            for (int i = 0; i < p->gh_dim; ++i) q += x[i] * x[i];
            return p->cp[0] * q;
        }
    }
    return 0.0;
}

/* ---------------------------------------------------------------------------------------------- factors */
/* TrajectoryBlock::extract (helpers/MatrixHelper.h:132-138): dim x dim marginal of a factor spanning nst states */
static void extract_cov(int d, int nst, int s, const double* cD, const double* cO, double* Sig) {
    const int dim = d * nst, dd = d * d;
    for (int c = 0; c < d; ++c)
        for (int r = 0; r < d; ++r) {
            Sig[r + c * dim] = cD[(size_t)s * dd + r + c * d];
            if (nst == 2) {
                Sig[(d + r) + (d + c) * dim] = cD[(size_t)(s + 1) * dd + r + c * d];
                Sig[r + (d + c) * dim] = cO[(size_t)s * dd + r + c * d];
                Sig[(d + c) + r * dim] = cO[(size_t)s * dd + r + c * d];
            }
        }
}

/* update_sigmapoints (quadrature/SparseGaussHermite.h:231-243): X = Z sqrtm(Sigma)^T + mu, materialised like _sigmapts */
static void sigma_points(const orc_problem* p, const double* mu, const double* Sig, double* X) {
    const int dim = p->gh_dim;
    double S[MAXN * MAXN];
    sqrtm_psd(dim, Sig, S);
    for (int i = 0; i < p->n_nodes; ++i)
        for (int r = 0; r < dim; ++r) {
            double s = 0.0;
            for (int c = 0; c < dim; ++c) s += p->Z[(size_t)i * dim + c] * S[r + c * dim];
            X[(size_t)i * dim + r] = s + mu[r];
        }
}

/* E0, E1, E2 of one GH factor (ngd/NGDFactorizedBaseGH.h:46-48,57-67), x-space sums in node order
   (SparseGaussHermite::Integrate, quadrature/SparseGaussHermite.h:197-221).  faithful: three separate integrals. */
static void gh_moments_one(const orc_problem* p, const double* mu, const double* Sig, double* X, int faithful, double* E0,
                           double* E1, double* E2) {
    const int dim = p->gh_dim, n = p->n_nodes;
    sigma_points(p, mu, Sig, X);
    double e0 = 0.0;
    for (int r = 0; r < dim; ++r) E1[r] = 0.0;
    for (int e = 0; e < dim * dim; ++e) E2[e] = 0.0;
    if (faithful) {
        for (int i = 0; i < n; ++i) { /* Integrate(_func_Vmu) */
            const double ps = psi_eval(p, X + (size_t)i * dim);
            for (int r = 0; r < dim; ++r) E1[r] += (X[(size_t)i * dim + r] - mu[r]) * ps * p->w[i];
        }
        for (int i = 0; i < n; ++i) e0 += psi_eval(p, X + (size_t)i * dim) * p->w[i]; /* Integrate(_func_phi) */
        for (int i = 0; i < n; ++i) { /* Integrate(_func_Vmumu) */
            const double ps = psi_eval(p, X + (size_t)i * dim);
            for (int c = 0; c < dim; ++c)
                for (int r = 0; r < dim; ++r)
                    E2[r + c * dim] += (X[(size_t)i * dim + r] - mu[r]) * (X[(size_t)i * dim + c] - mu[c]) * ps * p->w[i];
        }
    } else {
        for (int i = 0; i < n; ++i) {
            const double ps = psi_eval(p, X + (size_t)i * dim);
            const double wp = ps * p->w[i];
            e0 += wp;
            if (wp == 0.0) continue;
            for (int r = 0; r < dim; ++r) E1[r] += (X[(size_t)i * dim + r] - mu[r]) * wp;
            for (int c = 0; c < dim; ++c)
                for (int r = 0; r < dim; ++r)
                    E2[r + c * dim] += (X[(size_t)i * dim + r] - mu[r]) * (X[(size_t)i * dim + c] - mu[c]) * wp;
        }
    }
    *E0 = e0;
}

int orc_gh_moments(const orc_problem* p, const double* mu, const double* cD, const double* cO, int faithful, double* E0,
                   double* E1, double* E2) {
    const int dim = p->gh_dim, nst = dim / p->d;
#pragma omp parallel
    {
        double* X = (double*)malloc(sizeof(double) * (size_t)p->n_nodes * dim);
#pragma omp for schedule(dynamic, 16)
        for (int f = 0; f < p->n_gh; ++f) {
            double Sig[MAXN * MAXN];
            extract_cov(p->d, nst, p->gh_start[f], cD, cO, Sig);
            gh_moments_one(p, mu + (size_t)p->gh_start[f] * p->d, Sig, X, faithful, E0 + f, E1 + (size_t)f * dim,
                           E2 + (size_t)f * dim * dim);
        }
        free(X);
    }
    return 0;
}

/* closed-form pieces of one linear factor (ngd/NGDFactorizedLinear.h:93-129) */
static void lin_common(const orc_lin_group* g, int f, const double* mu_k, double* r, double* Kr, double* A) {
    const int dim = g->dim, m = g->m;
    const double* L = g->Lambda + (size_t)f * m * dim;
    const double* P = g->Psi + (size_t)f * m * g->kdim;
    const double* mt = g->mu_t + (size_t)f * g->kdim;
    const double* K = g->Kinv + (size_t)f * m * m;
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int k = 0; k < dim; ++k) s += L[i + k * m] * mu_k[k];
        for (int k = 0; k < g->kdim; ++k) s -= P[i + k * m] * mt[k];
        r[i] = s;
    }
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int k = 0; k < m; ++k) s += K[i + k * m] * r[k];
        Kr[i] = s;
    }
    if (A) {
        double KL[MAXN * MAXN];
        for (int j = 0; j < dim; ++j)
            for (int i = 0; i < m; ++i) {
                double s = 0.0;
                for (int k = 0; k < m; ++k) s += K[i + k * m] * L[k + j * m];
                KL[i + j * m] = s;
            }
        for (int j = 0; j < dim; ++j)
            for (int i = 0; i < dim; ++i) {
                double s = 0.0;
                for (int k = 0; k < m; ++k) s += L[k + i * m] * KL[k + j * m];
                A[i + j * dim] = s;
            }
    }
}

/* factor_cost_vector (gvibase/GVI-GH-GBP-impl.h:188-212); costs are written per group: fc_gh[n_gh], fc_lin[group][n] */
static void factor_costs(const orc_problem* p, const double* mu, const double* cD, const double* cO, double* fc_gh,
                         double** fc_lin) {
    const int d = p->d;
    if (p->n_gh > 0) {
        const int dim = p->gh_dim, nst = dim / d;
#pragma omp parallel
        {
            double* X = (double*)malloc(sizeof(double) * (size_t)p->n_nodes * dim);
#pragma omp for schedule(dynamic, 16)
            for (int f = 0; f < p->n_gh; ++f) { /* NGDFactorizedBaseGH::fact_cost_value :122-129 */
                double Sig[MAXN * MAXN];
                extract_cov(d, nst, p->gh_start[f], cD, cO, Sig);
                const double* mk = mu + (size_t)p->gh_start[f] * d;
                sigma_points(p, mk, Sig, X);
                double e0 = 0.0;
                for (int i = 0; i < p->n_nodes; ++i) e0 += psi_eval(p, X + (size_t)i * dim) * p->w[i];
                fc_gh[f] = e0 / p->gh_T[f];
            }
            free(X);
        }
    }
    for (int gi = 0; gi < p->n_lin_groups; ++gi) {
        const orc_lin_group* g = &p->lin[gi];
        const int dim = g->dim, nst = dim / d;
#pragma omp parallel for schedule(static)
        for (int f = 0; f < g->n; ++f) { /* NGDFactorizedLinear::fact_cost_value :122-129 */
            double r[MAXN], Kr[MAXN], A[MAXN * MAXN], Sig[MAXN * MAXN];
            lin_common(g, f, mu + (size_t)g->start[f] * d, r, Kr, A);
            extract_cov(d, nst, g->start[f], cD, cO, Sig);
            double tr = 0.0, q = 0.0;
            for (int j = 0; j < dim; ++j)
                for (int i = 0; i < dim; ++i) tr += A[i + j * dim] * Sig[j + i * dim];
            for (int i = 0; i < g->m; ++i) q += r[i] * Kr[i];
            fc_lin[gi][f] = (tr + q) * g->C[f] / g->T[f];
        }
    }
}

int orc_factor_costs(const orc_problem* p, const double* mu, const double* cD, const double* cO, double* fc) {
    double* fl[MAX_LIN_GROUPS];
    /* output order: the GH group, then the linear groups in group order */
    size_t off = (size_t)p->n_gh;
    for (int g = 0; g < p->n_lin_groups; ++g) {
        fl[g] = fc + off;
        off += (size_t)p->lin[g].n;
    }
    factor_costs(p, mu, cD, cO, fc, fl);
    return 0;
}

/* NGDGH::compute_gradients (ngd/NGD-GH-impl.h:20-63) up to the linear solve: per-factor calculate_partial_V, then the
   local2joint scatter (ngd/NGDFactorizedBaseGH.h:91-106) in factor order.  Outputs joint Vdmu[S*d], VD, VO. */
int orc_assemble_V(const orc_problem* p, const double* mu, const double* cD, const double* cO, int faithful, double* Vdmu,
                   double* VD, double* VO) {
    const int S = p->S, d = p->d, dd = d * d;
    memset(Vdmu, 0, sizeof(double) * (size_t)S * d);
    memset(VD, 0, sizeof(double) * (size_t)S * dd);
    memset(VO, 0, sizeof(double) * (size_t)(S > 1 ? S - 1 : 1) * dd);
    int rc = 0;
    /* ---- GH factors: NGDFactorizedBaseGH::calculate_partial_V :53-74 ---- */
    if (p->n_gh > 0) {
        const int dim = p->gh_dim, nst = dim / d;
        double* fV = (double*)malloc(sizeof(double) * (size_t)p->n_gh * dim);
        double* fM = (double*)malloc(sizeof(double) * (size_t)p->n_gh * dim * dim);
#pragma omp parallel
        {
            double* X = (double*)malloc(sizeof(double) * (size_t)p->n_nodes * dim);
#pragma omp for schedule(dynamic, 16)
            for (int f = 0; f < p->n_gh; ++f) {
                double Sig[MAXN * MAXN], P[MAXN * MAXN], E0, E1[MAXN], E2[MAXN * MAXN], T1[MAXN * MAXN], M[MAXN * MAXN];
                extract_cov(d, nst, p->gh_start[f], cD, cO, Sig);
                if (mat_inv(dim, Sig, P)) {
#pragma omp atomic write
                    rc = -1;
                }
                gh_moments_one(p, mu + (size_t)p->gh_start[f] * d, Sig, X, faithful, &E0, E1, E2);
                const double T = p->gh_T[f];
                for (int r = 0; r < dim; ++r) {
                    double s = 0.0;
                    for (int k = 0; k < dim; ++k) s += P[r + k * dim] * E1[k];
                    fV[(size_t)f * dim + r] = s / T;
                }
                mm(dim, P, 0, E2, 0, T1);
                mm(dim, T1, 0, P, 0, M);
                for (int e = 0; e < dim * dim; ++e) M[e] -= P[e] * E0;
                double* out = fM + (size_t)f * dim * dim;
                for (int c = 0; c < dim; ++c)
                    for (int r = 0; r <= c; ++r) { /* upper triangle mirrored :71-72 */
                        out[r + c * dim] = M[r + c * dim] / T;
                        out[c + r * dim] = M[r + c * dim] / T;
                    }
            }
            free(X);
        }
        for (int f = 0; f < p->n_gh; ++f) {
            const int s = p->gh_start[f];
            const double* M = fM + (size_t)f * dim * dim;
            for (int r = 0; r < dim; ++r) Vdmu[(size_t)s * d + r] += fV[(size_t)f * dim + r];
            for (int c = 0; c < d; ++c)
                for (int r = 0; r < d; ++r) {
                    VD[(size_t)s * dd + r + c * d] += M[r + c * dim];
                    if (nst == 2) {
                        VD[(size_t)(s + 1) * dd + r + c * d] += M[(d + r) + (d + c) * dim];
                        VO[(size_t)s * dd + r + c * d] += M[r + (d + c) * dim];
                    }
                }
        }
        free(fV);
        free(fM);
    }
    /* ---- linear factors: NGDFactorizedLinear::calculate_partial_V :93-120 ---- */
    for (int gi = 0; gi < p->n_lin_groups; ++gi) {
        const orc_lin_group* g = &p->lin[gi];
        const int dim = g->dim, nst = dim / d, m = g->m;
        double* fV = (double*)malloc(sizeof(double) * (size_t)g->n * dim);
        double* fM = (double*)malloc(sizeof(double) * (size_t)g->n * dim * dim);
#pragma omp parallel for schedule(static)
        for (int f = 0; f < g->n; ++f) {
            double r[MAXN], Kr[MAXN], A[MAXN * MAXN];
            lin_common(g, f, mu + (size_t)g->start[f] * d, r, Kr, A);
            const double* L = g->Lambda + (size_t)f * m * dim;
            const double ct = g->C[f] / g->T[f];
            for (int k = 0; k < dim; ++k) {
                double s = 0.0;
                for (int i = 0; i < m; ++i) s += L[i + k * m] * Kr[i];
                fV[(size_t)f * dim + k] = 2.0 * s * ct;
            }
            double* out = fM + (size_t)f * dim * dim;
            if (faithful) { /* the O(dim^4) fourth-moment loop :107-119 */
                double Sig[MAXN * MAXN], P[MAXN * MAXN], tmp[MAXN * MAXN], T1[MAXN * MAXN], M[MAXN * MAXN];
                extract_cov(d, nst, g->start[f], cD, cO, Sig);
                mat_inv(dim, Sig, P);
                double tr = 0.0;
                for (int j = 0; j < dim; ++j)
                    for (int i = 0; i < dim; ++i) tr += A[i + j * dim] * Sig[j + i * dim];
                for (int i = 0; i < dim; ++i)
                    for (int j = 0; j < dim; ++j) {
                        double s = 0.0;
                        for (int k = 0; k < dim; ++k)
                            for (int l = 0; l < dim; ++l)
                                s += (Sig[i + j * dim] * Sig[k + l * dim] + Sig[i + k * dim] * Sig[j + l * dim] +
                                      Sig[i + l * dim] * Sig[j + k * dim]) * A[k + l * dim];
                        tmp[i + j * dim] = s;
                    }
                mm(dim, P, 0, tmp, 0, T1);
                mm(dim, T1, 0, P, 0, M);
                for (int e = 0; e < dim * dim; ++e) out[e] = (M[e] - P[e] * tr) * ct;
            } else {
                for (int e = 0; e < dim * dim; ++e) out[e] = 2.0 * A[e] * ct;
            }
        }
        for (int f = 0; f < g->n; ++f) {
            const int s = g->start[f];
            const double* M = fM + (size_t)f * dim * dim;
            for (int r = 0; r < dim; ++r) Vdmu[(size_t)s * d + r] += fV[(size_t)f * dim + r];
            for (int c = 0; c < d; ++c)
                for (int r = 0; r < d; ++r) {
                    VD[(size_t)s * dd + r + c * d] += M[r + c * dim];
                    if (nst == 2) {
                        VD[(size_t)(s + 1) * dd + r + c * d] += M[(d + r) + (d + c) * dim];
                        VO[(size_t)s * dd + r + c * d] += M[r + (d + c) * dim];
                    }
                }
        }
        free(fV);
        free(fM);
    }
    return rc;
}

/* GVIGH::cost_value(mean, Precision) (gvibase/GVI-GH-GBP-impl.h:217-239): inverse_GBP, sum of factor costs in factor
   order (GH group first here, then linear groups -- the sum is what matters), + 1/2 log det. */
static int cost_value(const orc_problem* p, const double* mu, const double* LD, const double* LO, double* cD, double* cO,
                      double* fc, double* cost, orc_stats* st) {
    /* a precision that is not positive definite: the reference's sum of log(LDLT pivots) is NaN and the comparison
       `new_cost < cost_iter` (:99) is false -- the trial is rejected.  Same outcome (status -4 -> rejected by the caller)
       without pushing the garbage covariances of an indefinite matrix through the factors (NaN sigma points). */
    double ld = 0.0;
    if (orc_block_solve(p->S, p->d, LD, LO, NULL, NULL, &ld)) return -4;
    if (orc_inverse_gbp(p->S, p->d, LD, LO, cD, cO)) return -4;
    st->n_inversions++;
    orc_factor_costs(p, mu, cD, cO, fc);
    if (p->n_gh > 0) st->n_psi_sweeps++;
    size_t nf = (size_t)p->n_gh;
    for (int g = 0; g < p->n_lin_groups; ++g) nf += (size_t)p->lin[g].n;
    double v = 0.0;
    for (size_t i = 0; i < nf; ++i) v += fc[i];
    *cost = v + ld / 2;
    return 0;
}

int orc_cost_value(const orc_problem* p, const double* mu, const double* LD, const double* LO, double* cost, double* fc) {
    const int S = p->S, dd = p->d * p->d;
    double* cD = (double*)malloc(sizeof(double) * (size_t)S * dd);
    double* cO = (double*)malloc(sizeof(double) * (size_t)S * dd);
    orc_stats st;
    memset(&st, 0, sizeof(st));
    int rc = cost_value(p, mu, LD, LO, cD, cO, fc, cost, &st);
    free(cD);
    free(cO);
    return rc;
}

/* One iteration of GVIGH::optimize (gvibase/GVI-GH-GBP-impl.h:33-130) with NGDGH::compute_gradients /
   onestep_linesearch / update_proposal (ngd/NGD-GH-impl.h:20-63,129-156).  State in/out: mu[S*d], LD, LO (precision),
   cD, cO (covariance; must hold inverse(LD, LO) on entry when schedule == 1 and *have_cov != 0).
   schedule: 0 reference, 1 lean (see file header).  No temperature switching (callers keep niters_lowtemp > niters). */
int orc_ngd_iterate(const orc_problem* p, double* mu, double* LD, double* LO, double* cD, double* cO, double step_size_base,
                    double backtrack_ratio, int max_backtrack, int schedule, double* carried_cost, int* have_carry,
                    orc_stats* st) {
    const int S = p->S, d = p->d, dd = d * d;
    const size_t nmu = (size_t)S * d, nD = (size_t)S * dd;
    size_t nf = (size_t)p->n_gh;
    for (int g = 0; g < p->n_lin_groups; ++g) nf += (size_t)p->lin[g].n;
    memset(st, 0, sizeof(*st));
    double* fc = (double*)malloc(sizeof(double) * (nf ? nf : 1));
    double* Vdmu = (double*)malloc(sizeof(double) * nmu);
    double* VD = (double*)malloc(sizeof(double) * nD);
    double* VO = (double*)malloc(sizeof(double) * nD);
    double* dmu = (double*)malloc(sizeof(double) * nmu);
    double* mu_c = (double*)malloc(sizeof(double) * nmu);
    double* LD_c = (double*)malloc(sizeof(double) * nD);
    double* LO_c = (double*)malloc(sizeof(double) * nD);
    double* cD_c = (double*)malloc(sizeof(double) * nD);
    double* cO_c = (double*)malloc(sizeof(double) * nD);
    int rc = 0;
    double cost_iter = 0.0;
    const int faithful = (schedule == 0);
    if (schedule == 1 && *have_carry) {
        cost_iter = *carried_cost; /* cost and covariance of the accepted trial */
    } else {
        rc = cost_value(p, mu, LD, LO, cD, cO, fc, &cost_iter, st); /* cost_iter = cost_value() :61 */
        if (!rc && schedule == 0) {
            double dummy; /* fact_costs = factor_cost_vector() :69 -- a second inversion + cost sweep */
            rc = cost_value(p, mu, LD, LO, cD, cO, fc, &dummy, st);
        }
    }
    st->cost = cost_iter;
    if (!rc) {
        rc = orc_assemble_V(p, mu, cD, cO, faithful, Vdmu, VD, VO);
        if (p->n_gh > 0) st->n_psi_sweeps += faithful ? 3 : 1;
        for (size_t i = 0; i < nmu; ++i) Vdmu[i] = -Vdmu[i];
        if (!rc) rc = orc_block_solve(S, d, VD, VO, Vdmu, dmu, NULL);
    }
    if (!rc) {
        int cnt = 0;
        double step = step_size_base;
        for (;;) {
            step *= backtrack_ratio;
            for (size_t i = 0; i < nmu; ++i) mu_c[i] = mu[i] + step * dmu[i];
            for (size_t i = 0; i < nD; ++i) LD_c[i] = LD[i] + step * (VD[i] - LD[i]);
            for (size_t i = 0; i < (size_t)(S - 1) * dd; ++i) LO_c[i] = LO[i] + step * (VO[i] - LO[i]);
            double new_cost = 0.0;
            int trc = cost_value(p, mu_c, LD_c, LO_c, cD_c, cO_c, fc, &new_cost, st);
            st->new_cost = new_cost;
            if (trc == 0 && new_cost < cost_iter) {
                memcpy(mu, mu_c, sizeof(double) * nmu);
                memcpy(LD, LD_c, sizeof(double) * nD);
                if (S > 1) memcpy(LO, LO_c, sizeof(double) * (size_t)(S - 1) * dd);
                if (schedule == 0) { /* set_precision: inverse_GBP again (GVI-GH-GBP-impl.h:169-183) */
                    orc_inverse_gbp(S, d, LD, LO, cD, cO);
                    st->n_inversions++;
                } else {
                    memcpy(cD, cD_c, sizeof(double) * nD);
                    if (S > 1) memcpy(cO, cO_c, sizeof(double) * (size_t)(S - 1) * dd);
                    *carried_cost = new_cost;
                    *have_carry = 1;
                }
                st->accepted = 1;
                st->step = step;
                st->n_backtrack = cnt;
                break;
            }
            cnt++;
            if (cnt > max_backtrack) {
                st->n_backtrack = cnt;
                break;
            }
        }
    }
    st->status = rc;
    free(fc); free(Vdmu); free(VD); free(VO); free(dmu); free(mu_c); free(LD_c); free(LO_c); free(cD_c); free(cO_c);
    return rc;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
