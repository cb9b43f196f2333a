/* gvib200 -- C-ABI of the B200-native factorized NGD-GVI hot path (libgvib200.so).
 *
 * The reference (hzyu17/GaussianVI) has no FFI boundary: callers are C++ translation units
 * that include its header-only templates.  This header is therefore the
 * boundary a maintainer would bind to; each entry point names the reference interface it
 * replaces (paths relative to the reference root).  The C++ facade in
 * gaussianvi_b200/cpp/gvi/ re-exposes the reference's class names on top of these calls.
 *
 * Conventions
 *   - every function returns 0 on success or a negative GVIB200_E* code; the message of the last
 *     failure on the calling thread is gvib200_last_error().
 *   - all array arguments are caller-owned HOST pointers; doubles, column-major for matrices
 *     (Eigen's default), int32 for indices.  Copies are synchronous.  The library owns all device
 *     memory behind the opaque handles.
 *   - block-tridiagonal matrices (the sparsity pattern fixed at gvibase/GVI-GH.h:214-230) are
 *     passed as  diag[S][d*d]  (block (i,i)) and  off[S-1][d*d]  (block (i,i+1), column-major).
 *   - one ctx per GPU, one CUDA stream per problem; handles are not thread-safe, distinct
 *     handles may be driven from distinct host threads.
 *   - there is no CPU fallback: without a usable CUDA device ctx_create fails with GVIB200_ECUDA.
 */
#ifndef GVIB200_H
#define GVIB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVIB200_OK 0
#define GVIB200_EINVAL (-1)   /* bad argument                                         */
#define GVIB200_ECUDA (-2)    /* CUDA runtime error / no device                        */
#define GVIB200_ESTATE (-3)   /* call order violated (e.g. iterate before set_state)   */
#define GVIB200_ENOTSPD (-4)  /* a block-tridiagonal matrix had a non-positive pivot   */
#define GVIB200_ENOTABLE (-5) /* (dim, deg) quadrature table unavailable               */
#define GVIB200_ENCCL (-6)    /* collective failure                                    */

typedef struct gvib200_ctx gvib200_ctx;
typedef struct gvib200_problem gvib200_problem;

/* ---- cost functors psi(x) evaluated at sigma points ------------- */
enum {
    /* src/1d_example.cpp:25-35 (params: gvib200_stereo1d_params; tests/test_GH.cpp:21-34 uses y_offset=+0.05) */
    GVIB200_COST_STEREO_1D = 1,
    /* CudaOperation_PlanarPR::cost_obstacle_planar, helpers/CudaOperation.h:491-508 over PlanarSDF :27-128
       (params: gvib200_hinge_params; field set with gvib200_set_planar_sdf) */
    GVIB200_COST_PLANAR_HINGE = 2,
    /* cost_linear_gp, gp/cost_functions.h:36-39 / gp/cost_functions_LTV.h:34-37:
       1/2 (Phi th1 - th2)^T Qinv (Phi th1 - th2); per-factor params: Phi[ds*ds], Qinv[ds*ds] column-major */
    GVIB200_COST_LINEAR_GP = 3,
    /* cost_fixed_gp, gp/cost_functions.h:25-27: (x-mu)^T Kinv (x-mu); per-factor params: Kinv[dim*dim], mu[dim] */
    GVIB200_COST_FIXED_GP = 4,
    /* x^T (c I) x  -- the test integrand gx_1d of tests/test_gh_spgh.cpp:21-25 (params: one double c) */
    GVIB200_COST_QUADRATIC = 5,
    /* CudaOperation_3dpR::cost_obstacle_planar, helpers/CudaOperation.h:641-674 over the 3-D SignedDistanceField
       :133-236 (trilinear lookup; params: gvib200_hinge_params; field set with gvib200_set_sdf3d); reads x[0:3] */
    GVIB200_COST_HINGE_3D = 6,
    /* CudaOperation_Quad::cost_obstacle_planar, helpers/CudaOperation.h:565-605: planar quadrotor (x, z, phi, ...), five
       check points along the body axis (L = 5), slope 5, over the PlanarSDF (params: gvib200_hinge_params) */
    GVIB200_COST_QUAD_HINGE = 7,
    /* CudaOperation_3dArm::cost_obstacle, helpers/CudaOperation.h:751-770, over ForwardKinematics :325-410 (Denavit-Hartenberg
       chain with body spheres) and the 3-D SignedDistanceField (params: gvib200_arm_params; field set with gvib200_set_sdf3d).
       The state is (joint angles, joint velocities): factor dim = state dim = 2 n_dof, n_dof <= 3 (the block-tridiagonal
       engine holds state blocks up to 6 x 6, so a 7-DOF arm with 14-dimensional states is out of reach -- DESIGN.md 7).
       Mirrored quirks: the number of spheres evaluated is min(2 n_dof, n_spheres) (n_balls = theta.size(), :752) and
       dh_matrix's single-precision cosf / sinf (:394-400), as the correctly rounded single-precision values. */
    GVIB200_COST_ARM_3D = 8
};

typedef struct {
    double mu_p, f, b, sig_r_sq, sig_p_sq, y_offset; /* y = f*b/mu_p + y_offset; 1d_example: -0.8 */
} gvib200_stereo1d_params;

typedef struct {
    double sigma, epsilon, radius; /* helpers/CudaOperation.h:456 defaults 15.5, 0.5, 1 */
} gvib200_hinge_params;

#define GVIB200_ARM_MAX_DOF 3
#define GVIB200_ARM_MAX_SPHERES 12
typedef struct {
    double sigma, epsilon;                     /* helpers/CudaOperation.h:683-684 defaults 15.5, 0.5 */
    int n_dof, n_spheres;
    double a[GVIB200_ARM_MAX_DOF], alpha[GVIB200_ARM_MAX_DOF], d[GVIB200_ARM_MAX_DOF], theta_bias[GVIB200_ARM_MAX_DOF];
    int frames[GVIB200_ARM_MAX_SPHERES];       /* joint frame of every body sphere */
    double centers[GVIB200_ARM_MAX_SPHERES][3];/* centre in that frame (row per sphere, helpers/CudaOperation.h:343) */
    double radii[GVIB200_ARM_MAX_SPHERES];
} gvib200_arm_params;

/* ---- context ---------------------------------------------------------------------- */
int gvib200_ctx_create(int device, gvib200_ctx** out);
int gvib200_ctx_destroy(gvib200_ctx* ctx);
const char* gvib200_last_error(void);
/* library build identification: "gvib200 <version> sm_100a" */
const char* gvib200_version(void);

/* Multi-GPU (one process per GPU): attach an initialised NCCL communicator (ncclComm_t passed as void*),
   this rank's index and the world size, and the path of the libnccl the communicator was created with (the library
   resolves ncclAllGather / ncclAllReduce from it at run time; NULL = "libnccl.so.2").  Every problem created afterwards
   on this ctx is ONE TIME SEGMENT of a chain that is cut along the time axis: rank r owns the links [r m, (r+1) m) and
   both of its end states, i.e. num_states = m + 1 local states of which the first / last are shared with the
   neighbouring ranks.  A shared state's precision diagonal block and its factors are split between the two ranks (each
   passes ITS share; only the sums matter), its mean is passed identically to both.  Per block-tridiagonal pass the ranks
   exchange 5 d^2 + 4 d doubles in one all-gather; per cost evaluation 4 doubles in one all-reduce.  All ranks must make
   the same sequence of calls. */
int gvib200_ctx_set_comm(gvib200_ctx* ctx, void* nccl_comm, int rank, int world, const char* libnccl_path);
/* Peer-memory exchange (optional, after gvib200_ctx_set_comm; one box, GPUs connected by NVLink / NVSwitch): every rank
   creates a mailbox in its device memory and hands out its CUDA IPC handle (64 bytes), the handles of all ranks are
   exchanged by the caller (any transport: the ctypes mirror uses torch.distributed) and given to _connect in rank order.
   From then on the boundary records of the chain passes and the cost / flag exchange of a line-search trial are pushed
   into the peers' mailboxes by the kernels themselves (plain NVLink stores + a release flag, polled by the consumer):
   no NCCL call and no extra launch on the iteration path -- a multi-GPU block-tridiagonal pass is 3 launches like a
   single-GPU one.  Without a connected mailbox the library falls back to the NCCL all-gathers.  Returns the handle size. */
int gvib200_ctx_mailbox_create(gvib200_ctx* ctx, void* handle_out, size_t handle_capacity);
int gvib200_ctx_mailbox_connect(gvib200_ctx* ctx, int world, int rank, const void* handles, size_t handle_stride);

/* ---- sparse Gauss-Hermite tables (replaces the cereal map consumed at
        quadrature/SparseGaussHermite.h:138-166 and its MATLAB generator
        quadrature/generateSpGHWeights.h:23-84) -------------------------------------------- */
/* number of nodes of the (dim, deg) rule, or a negative error */
int gvib200_table_size(int dim, int deg);
/* host-side generation into caller buffers: nodes_rowmajor[n*dim], weights[n] */
int gvib200_table_generate(int dim, int deg, double* nodes_rowmajor, double* weights, int capacity);
/* override the generated table of (dim, deg) with an externally loaded one (e.g. read from the
   reference's SparseGHQuadratureWeights_cereal.bin) */
int gvib200_table_set(gvib200_ctx* ctx, int dim, int deg, int n, const double* nodes_rowmajor, const double* weights);
/* Table file in the reference's wire format: cereal BinaryOutputArchive of
   unordered_map<tuple<double,double>, tuple<MatrixXd,VectorXd>> (quadrature/saveSparseGHWeightMap.h:43-51,
   helpers/SerializeEigenMaps.h:195-224), the file SparseGaussHermite reads at construction
   (quadrature/SparseGaussHermite.h:58-72).
   _write: generate the n_keys rules (dims[i], degs[i]) and write them (replaces save_pointweightmaps()).
   _load:  read a file and register every rule in the context (as gvib200_table_set does); *n_loaded = rules read.
   _query: without a context: number of rules in the file; if dims/degs/sizes are non-null, up to `capacity` of them. */
int gvib200_table_file_write(const char* path, int n_keys, const int32_t* dims, const int32_t* degs);
int gvib200_table_file_load(gvib200_ctx* ctx, const char* path, int* n_loaded);
int gvib200_table_file_query(const char* path, int capacity, int32_t* dims, int32_t* degs, int32_t* sizes);
/* copy the rule (dim, deg) the context currently holds (generated or loaded) into caller buffers; returns the number of
   nodes (both buffers null: size query only) */
int gvib200_table_get(gvib200_ctx* ctx, int dim, int deg, double* nodes_rowmajor, double* weights, int capacity);

/* ---- problem definition (replaces the construction of GVIGH<Factor>, gvibase/GVI-GH-GBP.h:41-64,
        from a vector of factor optimizers) --------------------------------------------------- */
int gvib200_problem_create(gvib200_ctx* ctx, int num_states, int dim_state, gvib200_problem** out);
int gvib200_problem_destroy(gvib200_problem* prob);

/* signed-distance field used by GVIB200_COST_PLANAR_HINGE: PlanarSDF(origin, cell_size, data)
   helpers/CudaOperation.h:40-45; data column-major rows x cols */
int gvib200_set_planar_sdf(gvib200_problem* prob, int rows, int cols, double origin_x, double origin_y,
                           double cell_size, const double* data_colmajor);

/* 3-D signed-distance field used by GVIB200_COST_HINGE_3D: SignedDistanceField(origin, cell_size, data)
   helpers/CudaOperation.h:151-160; data[r + c * rows + z * rows * cols] (:299-301), x along columns, y along rows */
int gvib200_set_sdf3d(gvib200_problem* prob, int rows, int cols, int nz, double origin_x, double origin_y, double origin_z,
                      double cell_size, const double* data);

/* n nonlinear factors NGDFactorizedBaseGH<CostClass>(dimension, state_dim, gh_degree, function, cost_class,
   num_states, start_index, temperature, high_temperature) -- ngd/NGDFactorizedBaseGH.h:37-41.
   dim must be a multiple of dim_state (1 or 2 consecutive states).  temperature/high_temperature: n
   values each, or NULL for the reference defaults 1.0 / 10.0.  cost_params: one struct for
   STEREO_1D / PLANAR_HINGE / QUADRATIC, n packed per-factor records for LINEAR_GP / FIXED_GP.
   Returns (through first_id) the global id of the first factor added; ids are consecutive and fix
   the order of factor_costs. */
int gvib200_add_gh_factors(gvib200_problem* prob, int cost_kind, int dim, int deg, int n, const int32_t* start_index,
                           const double* temperature, const double* high_temperature, const void* cost_params,
                           size_t cost_params_bytes, int* first_id);

/* n closed-form linear-Gaussian factors NGDFactorizedLinear<Factor>(dimension, dim_state, function,
   linear_factor, num_states, start_indx, temperature, high_temperature) -- ngd/NGDFactorizedLinear.h:28-47,
   where linear_factor supplies get_Lambda() [m x dim], get_Psi() [m x kdim], get_mu() [kdim],
   get_precision() [m x m], get_Constant() (gp/linear_factor.h:18-31).  Per-factor packed, column-major. */
int gvib200_add_linear_factors(gvib200_problem* prob, int dim, int m, int kdim, int n, const int32_t* start_index,
                               const double* Lambda, const double* Psi, const double* mu_t, const double* Kinv,
                               const double* C, const double* temperature, const double* high_temperature,
                               int* first_id);

/* freeze the factor set, build the state<->factor adjacency, upload tables */
int gvib200_problem_finalize(gvib200_problem* prob);

/* ---- state (GVIGH::set_initial_values / set_mu / set_precision, gvibase/GVI-GH-GBP.h:201-233,
        GVI-GH-GBP-impl.h:169-183: also recomputes the covariance blocks and every factor marginal) */
int gvib200_set_state(gvib200_problem* prob, const double* mu, const double* prec_diag, const double* prec_off);
int gvib200_get_mean(gvib200_problem* prob, double* mu);                                 /* GVIGH::mean()       */
int gvib200_get_prec_blocks(gvib200_problem* prob, double* diag, double* off);           /* GVIGH::precision()  */
int gvib200_get_cov_blocks(gvib200_problem* prob, double* diag, double* off);            /* GVIGH::covariance() */
/* Asynchronous variants of the same accessors for pipelines of independent problems (one handle each): the transfers use
   PINNED host memory and are enqueued on the handle's stream, the call returns without waiting, so the upload / download
   of one handle overlaps the iteration of another (both PCIe directions and the SMs busy at once).  set_state_async also
   enqueues the selected inverse and the factor marginals; a precision that is not positive definite is reported
   (GVIB200_ENOTSPD) by the next gvib200_ngd_iterate / gvib200_prox_iterate / synchronous accessor / gvib200_sync on that
   handle.  Host buffers handed to the *_async getters are valid after gvib200_sync.  The reference has no counterpart: its
   optimizer object owns host-side Eigen state (gvibase/GVI-GH-GBP.h:201-233) and its GPU path copies synchronously
   (helpers/CudaOperation.cu, cudaMemcpy + cudaDeviceSynchronize per call). */
int gvib200_set_state_async(gvib200_problem* prob, const double* mu, const double* prec_diag, const double* prec_off);
int gvib200_get_mean_async(gvib200_problem* prob, double* mu);
int gvib200_get_prec_blocks_async(gvib200_problem* prob, double* diag, double* off);
int gvib200_get_cov_blocks_async(gvib200_problem* prob, double* diag, double* off);
int gvib200_sync(gvib200_problem* prob);

/* ---- per-factor quadrature moments at the current state: E_Phis / E_xMuPhis / E_xMuxMuTPhis
        (gvibase/GVI-GH-GBP.h:348-378; SparseGaussHermite::Integrate quadrature/SparseGaussHermite.h:197-221).
        Output order: GH factors in id order; E1 packed [dim] and E2 packed [dim*dim] per factor using each
        factor's own dim.  Any pointer may be NULL. ------------------------------------------------------- */
int gvib200_moments(gvib200_problem* prob, double* E0, double* E1, double* E2);

/* GVIGH::cost_value(mean, Precision) and factor_cost_vector(mean, Precision), GVI-GH-GBP-impl.h:188-239.
   NULL mu/diag/off = the current state.  fac_costs (n_factors, id order) may be NULL. */
int gvib200_cost(gvib200_problem* prob, const double* mu, const double* prec_diag, const double* prec_off,
                 double* cost, double* fac_costs);

/* NGDGH::compute_gradients, ngd/NGD-GH-impl.h:20-63: dmu[S*d], dprecision as blocks.  The mean step is
   solved by a direct block-tridiagonal Cholesky (the reference calls Eigen CG capped at 2n iterations). */
int gvib200_gradients(gvib200_problem* prob, double* dmu, double* dprec_diag, double* dprec_off);
/* joint Vdmu / Vddmu of the last gradients call (NGDGH::Vdmu()/Vddmu(), ngd/NGD-GH.h:90-92) */
int gvib200_get_V(gvib200_problem* prob, double* Vdmu, double* Vddmu_diag, double* Vddmu_off);

typedef struct {
    double step_size_base;   /* GVIGH::_step_size_base, default 0.55 (gvibase/GVI-GH-GBP.h:94) */
    double backtrack_ratio;  /* 0.75, hard-coded at gvibase/GVI-GH-GBP-impl.h:89                */
    int max_backtrack;       /* _niters_backtrack, default 10                                   */
    int niters_lowtemp;      /* _niters_lowtemp, default 10                                     */
    int reuse_accepted_sweep;/* 0: faithful schedule (1 moment sweep + T_ls cost sweeps / iteration);
                                1: line-search sweeps compute full moments and the accepted trial's
                                   moments seed the next iteration (identical results, fewer sweeps) */
    double ema_alpha;        /* GVIGH::_alpha of the reference's GPU path, default 1 (gvibase/GVI-GH-Cuda.h:78,223): an
                                accepted trial takes alpha * new + (1 - alpha) * current for mean and precision
                                (gvibase/GVI-GH-Cuda-impl.h:112-114)                              */
} gvib200_opts;
void gvib200_default_opts(gvib200_opts* opts);

typedef struct {
    double cost;          /* cost_iter at the start of the iteration (recorded by the reference) */
    double new_cost;      /* cost of the accepted trial (or of the last rejected one)            */
    double step;          /* accepted step size                                                   */
    int n_backtrack;      /* rejected trials                                                      */
    int accepted;         /* 1 if a trial was accepted                                            */
    int switched_high_T;  /* 1 if the optimizer switched to the high temperature this iteration   */
    int converged;        /* 1 if back-tracking was exhausted at high temperature                 */
    int status;           /* 0 or GVIB200_ENOTSPD                                                 */
    int n_moment_sweeps;  /* quadrature sweeps executed (full moments)                            */
    int n_cost_sweeps;    /* quadrature sweeps executed (cost only)                               */
} gvib200_iter_stats;

/* the whole GVIGH::optimize loop (gvibase/GVI-GH-GBP-impl.h:33-130) with state resident in HBM:
   runs up to n_iters iterations, fills stats[0..n_done).  fac_costs_trace (n_iters x n_factors) and
   the per-iteration snapshots the reference's recorder keeps are optional (NULL to skip). */
int gvib200_optimize(gvib200_problem* prob, const gvib200_opts* opts, int n_iters, gvib200_iter_stats* stats,
                     int* n_done, double* fac_costs_trace, double* mean_trace);
/* one iteration (same code path; iteration index is kept inside the problem for the temperature switch) */
int gvib200_ngd_iterate(gvib200_problem* prob, const gvib200_opts* opts, gvib200_iter_stats* stats);

/* Result recorder (replaces VIMPResults::update_data / save_data, helpers/DataRecorder.h:96-118,177-224, and
   GVIGH::update_file_names / save_data, gvibase/GVI-GH.h:282-329) in BANDED form: the joint covariance / precision are
   recorded as their d x d diagonal blocks, exactly the blocks the reference's joint2marginals() extracts
   (helpers/DataRecorder.h:124-131); the dense joint_cov / joint_precision files are written only when the joint
   dimension is <= dense_limit (from the diagonal and first off-diagonal blocks: the precision is exactly block
   tridiagonal and the reference's covariance() holds only the block-tridiagonal part of the inverse,
   gvibase/GVI-GH-GBP-impl.h:245-305).  What is recorded at iteration i is the state BEFORE the step, its cost and its factor
   costs (gvibase/GVI-GH-GBP-impl.h:61-72).  All buffers are caller-owned host memory; null members are skipped. */
typedef struct gvib200_trace {
    int capacity;        /* iterations the buffers hold */
    int n_recorded;      /* out */
    double* mean;        /* [capacity][S*d] */
    double* cov_diag;    /* [capacity][S*d*d]  column-major d x d blocks */
    double* prec_diag;   /* [capacity][S*d*d] */
    double* cov_off;     /* [capacity][(S-1)*d*d]  blocks (i, i+1); only needed for the dense joint files */
    double* prec_off;    /* [capacity][(S-1)*d*d] */
    double* cost;        /* [capacity] */
    double* fac_costs;   /* [capacity][n_factors] */
} gvib200_trace;
/* prox != 0 runs gvib200_prox_iterate instead of gvib200_ngd_iterate */
int gvib200_optimize_traced(gvib200_problem* prob, const gvib200_opts* opts, int n_iters, int prox, gvib200_iter_stats* stats,
                            int* n_done, gvib200_trace* trace);
/* MatrixIO::saveData (helpers/MatrixHelper.h:52-61) with CSVFormat (helpers/CommonDefinitions.h:32): rows on lines,
   ", " between coefficients, 15 significant digits.  data is column-major rows x cols. */
int gvib200_csv_write(const char* path, int rows, int cols, const double* data_colmajor);
/* writes <prefix>mean[_afterfix].csv, cov, precision, cost, factor_costs, zk_sdf, Sk_sdf (and joint_cov / joint_precision
   when S*d <= dense_limit) with the reference's shapes: one column per iteration */
int gvib200_trace_save(const gvib200_trace* trace, int num_states, int dim_state, int n_factors, const char* prefix,
                       const char* afterfix, int dense_limit);
/* Prox-GVI (proxgd/ProxGVI-GH-impl.h:124-205 over ProxGVIFactorizedBaseGH / ProxFactorizedLinear): per-factor
   Bures-Wasserstein JKO steps instead of natural gradients, no linear solve, step eta = step_size_base^B, the last
   candidate is accepted when back-tracking is exhausted.  The problem must have been given
   gvib200_problem_set_option(prob, "prox", 1) BEFORE finalize.  Single GPU. */
int gvib200_prox_iterate(gvib200_problem* prob, const gvib200_opts* opts, gvib200_iter_stats* stats);
int gvib200_prox_optimize(gvib200_problem* prob, const gvib200_opts* opts, int n_iters, gvib200_iter_stats* stats, int* n_done);
/* reset the iteration counter / temperature phase (keeps the state) */
int gvib200_reset_schedule(gvib200_problem* prob);
/* GVIGH::switch_to_high_temperature (gvibase/GVI-GH-GBP-impl.h:18-27, public at gvibase/GVI-GH-GBP.h:361): every factor takes
   its high temperature now; the optimizer will not switch again by itself */
int gvib200_switch_to_high_temperature(gvib200_problem* prob);

/* ---- batches of INDEPENDENT problems with a line search per problem.  The reference runs one optimizer object per problem
        (GVIGH::optimize, gvibase/GVI-GH-GBP-impl.h:33-130: cost_iter, back-tracking count and step size, temperature phase
        and the converged flag belong to the object); here the problems are concatenated block-diagonally into ONE chain
        (problem q owns the states [state_offsets[q], state_offsets[q+1]), no factor and no coupling block crosses a
        boundary, every group's factors are ordered by problem) so that sweeps, assembly and chain passes run over the whole
        batch in single launches, and gvib200_batch_iterate keeps what is per object per problem: its cost (factor costs +
        log det / 2 of ITS precision block, from the per-node log pivots of the chain engine), its step size and
        back-tracking count, its temperature phase, its converged flag.  One call = one iteration of every problem that has
        not converged; stats[q] is problem q's iteration record (status GVIB200_ENOTSPD: that problem's Vddmu has no Cholesky
        factor, its state is unchanged); n_trials = trial sweeps the batch needed (the largest T_ls over the problems, +1
        when some problem exhausted its back-tracking).  gvib200_ngd_iterate on the same handle remains the joint line
        search (one step size, the summed cost).  ema_alpha != 1, Prox-GVI and multi-GPU chains are not available here. */
int gvib200_set_batch(gvib200_problem* prob, int n_problems, const int32_t* state_offsets /* [n_problems + 1] */);
int gvib200_batch_iterate(gvib200_problem* prob, const gvib200_opts* opts, gvib200_iter_stats* stats /* [n_problems] */,
                          int* n_trials);
/* GVIGH::optimize (gvibase/GVI-GH-GBP-impl.h:33-130) of every problem: up to n_iters lock-step iterations, stops when every
   problem has converged; stats[it * n_problems + q] is problem q's record of iteration it, *n_done the iterations run */
int gvib200_batch_optimize(gvib200_problem* prob, const gvib200_opts* opts, int n_iters, gvib200_iter_stats* stats, int* n_done);
int gvib200_batch_costs(gvib200_problem* prob, double* cost_per_problem /* [n_problems] */);

/* ---- device-side set-up of the LTV GP prior (gp/LTV_prior.h:123-197 compute_Phi_gsl / compute_Q_gsl with the piece-wise
        constant A_function / system_param of :187-197), batched over links: one launch integrates
        Phi' = A Phi, Q' = A Q + Q A^T + B B^T over [0, delta_t] for n_links links, A / B constant on each quarter interval.
        Where the reference runs GSL rkf45 at tolerance 1e-12 per link on the host, each quarter is integrated exactly by
        Van Loan's block exponential.  A: [n_links][4][dim_state x dim_state], B: [n_links][4][dim_state x n_inputs],
        outputs [n_links][dim_state x dim_state], all column-major; Qinv (optional, may be NULL) = Q^-1 = the K^-1 of the
        LTV_GP linear factor.  dim_state in {2, 4, 6}. */
int gvib200_ltv_transition(gvib200_ctx* ctx, int n_links, int dim_state, int n_inputs, const double* A, const double* B,
                           double delta_t, double* Phi, double* Q, double* Qinv);

/* ---- stand-alone block-tridiagonal engine (GVIGH::inverse_GBP, GVI-GH-GBP-impl.h:245-305;
        EigenWrapper::inv_sparse helpers/EigenWrapper.h:336-381; SparseLDLT log det :234-238) -------- */
int gvib200_selected_inverse(gvib200_ctx* ctx, int S, int d, const double* diag, const double* off, double* cov_diag,
                             double* cov_off, double* logdet);
int gvib200_blocktri_solve(gvib200_ctx* ctx, int S, int d, const double* diag, const double* off, const double* rhs,
                           double* x, double* logdet);

/* ---- measurement hooks (bench.py): device-resident timing of the stages with CUDA events on the
        problem's stream.  stage: 0 = moment sweep (K2+K1), 1 = cost sweep, 2 = assemble + dmu solve,
        3 = candidate + selected inverse + log det, 5 = the fused moment kernel K1 alone (full moments), 6 = K1 alone
        (cost only).  Returns milliseconds per repetition. */
int gvib200_time_stage(gvib200_problem* prob, int stage, int reps, const gvib200_opts* opts, float* ms_per_rep,
                       long long* kernel_launches);
int gvib200_fp64_peak(gvib200_ctx* ctx, double* tflops);   /* DFMA micro-benchmark: the FP64 roofline denominator */
long long gvib200_launch_count(gvib200_ctx* ctx);            /* kernels launched through this ctx so far */

/* CUDA-event stopwatch on the problem's stream (the stream every kernel of the problem is launched on) */
int gvib200_timer_start(gvib200_problem* prob);
int gvib200_timer_stop(gvib200_problem* prob, float* ms);

/* per-launch profile: between _begin and _end every kernel launch of the problem is bracketed by a CUDA event
   pair; _end returns launch counts and summed device time per kernel class (gvib200_kernel_class_name). */
typedef struct {
    int n_classes;
    long long count[16];
    double ms[16];
} gvib200_profile;
int gvib200_profile_begin(gvib200_problem* prob);
int gvib200_profile_end(gvib200_problem* prob, gvib200_profile* out);
const char* gvib200_kernel_class_name(int kernel_class);
/* Statistics of the free-space culling (problem option "cull", default on): GH factors whose sigma points were actually
   evaluated by the sign-group kernel since the counter was last reset.  A factor is culled only when its cost functor
   proves psi == 0 on the factor's whole sigma-point box (CostPlanarHinge: a conservative bound from the distance field);
   its moments are then exactly zero, which is also what the evaluation would return -- results are bit-identical. */
int gvib200_evaluated_factors(gvib200_problem* prob, long long* count, int reset);

typedef struct {
    int num_states, dim_state, n_factors, n_gh_factors, n_linear_factors, chain_levels, chain_tiles, chain_tile_links;
    long long sigma_points_per_sweep; /* sum over GH factors of the nodes of their rule */
} gvib200_info;
int gvib200_problem_info(gvib200_problem* prob, gvib200_info* out);

/* switches.  "prox" = 1 (before finalize): a Prox-GVI problem.  "generic_k1" = 1: run the generic node-loop moment kernel even where the sign-group kernel
   (dimension <= 4) applies; both must give the same moments to rounding. */
int gvib200_problem_set_option(gvib200_problem* prob, const char* name, int value);

/* device-side snapshot / rewind of the optimizer state (mean, precision, covariance, factor marginals,
   iteration counter); no host traffic */
int gvib200_snapshot_save(gvib200_problem* prob);
int gvib200_snapshot_restore(gvib200_problem* prob);

#ifdef __cplusplus
}
#endif
#endif /* GVIB200_H */
