// libgvib200.so -- C-ABI implementation (include/gvib200.h) and host control of the device-resident
// NGD-GVI iteration.  Host code mirrors GVIGH::optimize (gvibase/GVI-GH-GBP-impl.h:33-130) and
// NGDGH (ngd/NGD-GH-impl.h); all arithmetic of the hot path runs in the kernels of kernels.cuh.
#include "../../include/gvib200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <set>
#include <stdexcept>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <cstdlib>
#include <vector>

#include "bt_cr_plan.h"
#include "k1_grp.cuh"
#include "k1_sym.cuh"
#include "kernels.cuh"
#include "ltv_setup.cuh"
#include "spgh_table.h"

using namespace gvib200;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(GVIB200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));             \
    } while (0)
#define TRY(expr)                 \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != 0) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct Table {
    int dim = 0, deg = 0, n = 0, n_pad = 0;
    std::vector<double> nodes, w;  // host
    double ximax[12] = {};         // max |xi_c|
    double xinorm = 0.0;           // max ||xi||_2 over the nodes
    double* d_rows = nullptr;      // device planes: NP x [n_pad] double2, then [n_pad] weights
    bool sym_ok = false;           // the rule has the sign-group structure K1S needs (dim <= 4, fits shared memory)
    SymTable sym;
    std::vector<double> sym_data;  // host copy of sym.data
    double* d_sym = nullptr;
    bool grp_ok = false;           // the rule consists of full sign groups with <= K1G_KMAX non-zero coordinates (K1G)
    GrpTable grp{};
    int* d_grp_hdr = nullptr;
    double* d_grp_val = nullptr;
};

// Sparse sign-group view of a rule for K1G (k1_grp.cuh): groups of 2^k sign combinations of (a_1 .. a_k) on k <= K1G_KMAX
// coordinates with one common weight, sorted by decreasing k.  Returns false when the rule does not have that structure.
static bool build_grp_table(const Table& t, std::vector<int>& hdr, std::vector<double>& val, GrpTable& gt) {
    const int dim = t.dim, n = t.n;
    if (dim > 255) return false;
    struct Grp {
        std::vector<int> c;
        std::vector<double> a;
        double w = 0.0;
        int count = 0;
        unsigned seen = 0;
    };
    std::map<std::vector<double>, Grp> groups;
    gt = GrpTable();
    gt.dim = dim;
    for (int i = 0; i < n; ++i) {
        std::vector<double> key(dim);
        int k = 0, pat = 0;
        for (int c = 0; c < dim; ++c) {
            const double v = t.nodes[(size_t)i * dim + c];
            key[c] = std::fabs(v);
            if (v != 0.0) {
                if (v < 0.0) pat |= 1 << k;
                ++k;
            }
        }
        if (k == 0) {
            if (gt.has_origin) return false;
            gt.has_origin = 1;
            gt.w0 = t.w[i];
            continue;
        }
        if (k > K1G_KMAX) return false;
        auto it = groups.find(key);
        if (it == groups.end()) {
            Grp g;
            for (int c = 0; c < dim; ++c)
                if (key[c] != 0.0) {
                    g.c.push_back(c);
                    g.a.push_back(key[c]);
                }
            g.w = t.w[i];
            it = groups.emplace(key, g).first;
        }
        Grp& g = it->second;
        if (g.w != t.w[i] || (g.seen & (1u << pat))) return false;
        g.seen |= 1u << pat;
        g.count++;
    }
    std::vector<const Grp*> order;
    for (auto& kv : groups) {
        if (kv.second.count != (1 << kv.second.c.size())) return false;
        order.push_back(&kv.second);
    }
    std::stable_sort(order.begin(), order.end(), [](const Grp* x, const Grp* y) { return x->c.size() > y->c.size(); });
    hdr.clear();
    val.clear();
    for (const Grp* g : order) {
        int h = (int)g->c.size();
        for (size_t i = 0; i < g->c.size(); ++i) h |= g->c[i] << (4 + 8 * (int)i);
        hdr.push_back(h);
        for (int i = 0; i < 3; ++i) val.push_back(i < (int)g->a.size() ? g->a[(size_t)i] : 0.0);
        val.push_back(g->w);
    }
    gt.n_groups = (int)order.size();
    return gt.n_groups > 0;
}

// Sign-group view of a sparse-GH rule for K1S (k1_sym.cuh): every set of nodes sharing |xi| must be a full group of
// 2^k sign combinations with one common weight.  Returns false when the rule does not have that structure.
static bool build_sym_table(const Table& t, SymTable& st) {
    const int dim = t.dim, n = t.n;
    if (dim < 1 || dim > 4) return false;
    std::memset(&st, 0, sizeof(st));
    st.dim = dim;
    st.n_nodes = n;
    struct Grp {
        std::vector<double> a;  // magnitudes of the non-zero coordinates
        double w;
        int count;
        unsigned sign_seen;     // bitset over the 2^k sign patterns
    };
    double w0 = 0.0;                            // weight of the node at the origin, if any
    std::map<std::vector<double>, Grp> groups;  // key: |xi| (all dim coordinates)
    for (int i = 0; i < n; ++i) {
        std::vector<double> key(dim);
        int mask = 0, pat = 0, q = 0;
        for (int c = 0; c < dim; ++c) {
            const double v = t.nodes[(size_t)i * dim + c];
            key[c] = std::fabs(v);
            if (v != 0.0) {
                mask |= 1 << c;
                if (v < 0.0) pat |= 1 << q;
                ++q;
            }
        }
        if (mask == 0) {
            if (w0 != 0.0) return false;  // two nodes at the origin
            w0 = t.w[i];
            continue;
        }
        auto it = groups.find(key);
        if (it == groups.end()) {
            Grp g;
            for (int c = 0; c < dim; ++c)
                if (key[c] != 0.0) g.a.push_back(key[c]);
            g.w = t.w[i];
            g.count = 0;
            g.sign_seen = 0;
            it = groups.emplace(key, g).first;
        }
        Grp& g = it->second;
        if (g.w != t.w[i]) return false;           // weights must be bit-equal inside a group
        if (g.sign_seen & (1u << pat)) return false;
        g.sign_seen |= 1u << pat;
        g.count++;
    }
    std::vector<std::vector<const Grp*>> by_mask(16);
    for (auto& kv : groups) {
        int mask = 0;
        for (int c = 0; c < dim; ++c)
            if (kv.first[c] != 0.0) mask |= 1 << c;
        const int k = (int)kv.second.a.size();
        if (kv.second.count != (1 << k)) return false;  // not a full sign group
        by_mask[mask].push_back(&kv.second);
    }
    // device layout: mask 0 = the origin node ([part]: w0 for part 0); mask m: [round][entry][part], group 8 r + p of the
    // mask goes to part p in round r, the last round is padded with zero-weight groups at xi = 0
    std::vector<double>& data = const_cast<Table&>(t).sym_data;
    data.assign(K1S_NPART, 0.0);
    data[0] = w0;
    st.moff[0] = 0;
    st.rounds[0] = 1;
    for (int m = 1; m < 16; ++m) {
        st.moff[m] = (int)data.size();
        const int K = k1s_popc(m), stride = k1s_stride(K);
        const int G = (int)by_mask[m].size();
        const int rounds = (G + K1S_NPART - 1) / K1S_NPART;
        st.rounds[m] = rounds;
        data.resize(data.size() + (size_t)rounds * stride * K1S_NPART, 0.0);
        double* blk = data.data() + st.moff[m];
        for (int gi = 0; gi < G; ++gi) {
            const Grp* g = by_mask[m][gi];
            const int r = gi / K1S_NPART, p = gi % K1S_NPART;
            auto put = [&](int e, double v) { blk[((size_t)r * stride + e) * K1S_NPART + p] = v; };
            for (int i = 0; i < K; ++i) put(i, g->a[i]);
            put(K, g->w);
            for (int i = 0; i < K; ++i) put(K + 1 + i, g->w * g->a[i]);
            for (int i = 0; i < K; ++i) put(2 * K + 1 + i, g->w * g->a[i] * g->a[i]);
            for (int i = 0; i < K; ++i)
                for (int j = i + 1; j < K; ++j) put(3 * K + 1 + k1s_pair(K, i, j), g->w * g->a[i] * g->a[j]);
        }
    }
    if (data.size() > (size_t)K1S_MAX_DATA) return false;
    st.ndata = (int)data.size();
    return true;
}

struct gvib200_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    std::map<std::pair<int, int>, std::unique_ptr<Table>> tables;
    // multi-GPU: NCCL entry points resolved at run time from the library the caller initialised the communicator with
    void* nccl_comm = nullptr;
    void* nccl_comm2 = nullptr;  // split of nccl_comm for the side stream (the two chain passes of an iteration overlap)
    void* nccl_lib = nullptr;
    int (*ncclAllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int rank = 0, world = 1;
    // peer-memory mailbox (kernels.cuh "Multi-GPU exchange through peer memory"): when connected, the boundary records of
    // the chain passes and the cost / flag exchange travel as NVLink stores issued by the kernels themselves -- no NCCL call
    // on the iteration path
    double* mbox = nullptr;
    void* peer_mapped[MBOX_RANKS] = {};  // mappings opened with cudaIpcOpenMemHandle (closed on destroy)
    MboxPeers peers{};
    bool mbox_on = false;
    unsigned long long ep_rec[2] = {0, 0}, ep_cost = 0;  // epochs of the next exchange per slot / of the next cost exchange
    long long launches = 0;  // kernels launched through this ctx
    // kernels whose function attributes (dynamic shared memory opt-in) were set on THIS context's device: the attributes
    // are per device, so a process driving several contexts configures each of them
    std::set<const void*> configured;
    bool need_config(const void* kern) { return configured.insert(kern).second; }
};

struct GhGroup {
    int kind = 0, dim = 0, deg = 0, n = 0, first_id = 0;
    size_t voff = 0, moff = 0;  // offsets (doubles) of this group in fVdmu / fVdd
    std::vector<int> start;
    std::vector<double> T, Thigh;
    std::vector<char> params;
    const Table* table = nullptr;
    int* d_start = nullptr;
    double* d_T = nullptr;
    double* d_Thigh = nullptr;
    void* d_params = nullptr;
    double* d_SR[2] = {nullptr, nullptr};
    bool sr_full[2] = {false, false};  // d_SR[b] holds S, R of EVERY factor at the state of buffer b (the fused culling +
                                       // prologue pass only writes the factors it keeps)
    double* d_raw = nullptr;
    mutable int* d_active = nullptr;   // free-space culling: compacted list of the factors to evaluate
    mutable int* d_nactive = nullptr;  // [2]: list length, finished-CTA counter
};

struct LinGroup {
    int dim = 0, m = 0, kdim = 0, n = 0, first_id = 0;
    size_t voff = 0, moff = 0;  // moff: offset of the Vddmu blocks in fVdd (Prox mode only)
    std::vector<int> start;
    std::vector<double> Lambda, psi, Kinv, A, C, T, Thigh;
    int* d_start = nullptr;
    double *d_Lambda = nullptr, *d_psi = nullptr, *d_Kinv = nullptr, *d_A = nullptr, *d_C = nullptr, *d_T = nullptr;
};

// kernel classes for the per-launch profile (gvib200_profile_begin / _end)
enum { KC_MOMENTS_FULL = 0, KC_MOMENTS_COST, KC_PROLOGUE, KC_LINEAR, KC_ASSEMBLE, KC_BT_FORWARD, KC_BT_TOP, KC_BT_BACK,
       KC_SUM, KC_CANDIDATE, KC_OTHER, KC_CULL, KC_COUNT };
static const char* const KC_NAMES[KC_COUNT] = {"k_moments<full>", "k_moments<cost>", "k_prologue", "k_linear", "k_assemble",
                                               "k_bt_forward",   "k_bt_top",        "k_bt_back",  "k_sum",    "k_candidate",
                                               "other",          "k_cull"};
struct ProfRec {
    int kc;
    cudaEvent_t a, b;
    bool side;  // launched on the side stream
};

struct FactorRef {
    bool linear;
    int group;
    int index;
};

struct gvib200_problem {
    gvib200_ctx* ctx = nullptr;
    int S = 0, d = 0;
    cudaStream_t stream = nullptr;
    bool finalized = false, has_state = false;
    std::vector<GhGroup> gh;
    std::vector<LinGroup> lin;
    std::vector<FactorRef> factors;  // id order
    int n_factors = 0;
    size_t nV = 0, nM = 0;  // sizes of fVdmu / fVdd
    // SDF
    double4* d_sdf_rec = nullptr;   // hinge records, built for the threshold sdf_thr (epsilon + radius)
    double* d_sdf_data = nullptr;   // the raw field (column-major), kept to rebuild the records
    double* d_sdf_coarse = nullptr; // free-space culling: per 8 x 8 cell block, max of thr - (min corner distance)
    bool cull = true;               // option "cull"
    unsigned long long* d_evaluated = nullptr;  // factors evaluated by the sign-group kernel (statistics)
    double sdf_thr = 0.0;
    bool sdf_rec_valid = false;
    int sdf_rows = 0, sdf_cols = 0;
    double sdf_ox = 0, sdf_oy = 0, sdf_cell = 0;
    // 3-D SDF (GVIB200_COST_HINGE_3D)
    double* d_sdf3 = nullptr;
    int sdf3_rows = 0, sdf3_cols = 0, sdf3_nz = 0;
    double sdf3_o[3] = {0, 0, 0}, sdf3_cell = 0;
    // state, double buffered (cur / candidate)
    int cur = 0;
    double *mu[2] = {}, *LD[2] = {}, *LO[2] = {}, *CD[2] = {}, *CO[2] = {};
    double *fcost[2] = {}, *fVdmu[2] = {}, *fVdd[2] = {};
    double* partial = nullptr; // per-block partial sums of the factor costs
    double* scal = nullptr;    // device scalars: [0..1] logdet cur/cand slots, [2..3] cost slots, [4] tmp
    double* h_scal = nullptr;  // pinned mirror
    double* zc = nullptr;      // mapped pinned memory written by k_total: [0..1] total cost of buffer 0 / 1, [2..3] flags
    double* zc_dev = nullptr;  // its device address
    bool zc_ok[2] = {false, false};  // zc[which] is the cost of buffer `which` as it stands on the device
    int* d_flag = nullptr;     // not-SPD flag
    unsigned* d_counter = nullptr;  // arrival counter of k_total (zero between launches)
    int* h_flag = nullptr;
    double *Vdmu = nullptr, *VD = nullptr, *VO = nullptr, *rhs = nullptr, *dmu = nullptr;
    // second set of assembled gradients: the assembly of a trial's sweep is launched speculatively (before the host
    // knows whether the trial is accepted) so that the host round trip hides underneath it; swapped in on acceptance
    double *Vdmu2 = nullptr, *VD2 = nullptr, *VO2 = nullptr, *rhs2 = nullptr;
    bool vo_alias = false;  // no factor contributes an off-diagonal block of Vddmu: VO and VO2 alias the constant KlinO
    bool asm_valid = false;  // Vdmu / VD / VO / rhs hold the assembly of the sweep at the current state
    cudaEvent_t ev_host = nullptr;
    cudaEvent_t ev_pending = nullptr;  // gvib200_set_state_async: upload + selected inverse + factor marginals are complete
    double *KlinD = nullptr, *KlinO = nullptr;
    // adjacency
    // the same adjacency in ELL form (fixed width, -1 padded, [k][state]) when every state has few contributors: the
    // assembly then needs one index load per contribution instead of a pointer chase
    int *ell_v = nullptr, *ell_d = nullptr, *ell_dl = nullptr, *ell_o = nullptr, *ell_ol = nullptr;
    int ell_nv = -1, ell_nd = 0, ell_no = 0;  // widths; ell_nv < 0: no ELL form
    int *vptr = nullptr, *voff = nullptr, *dptr = nullptr, *doff = nullptr, *dld = nullptr, *optr = nullptr,
        *ooff = nullptr, *old = nullptr;
    // chain engine (bt_cr.h): one plan, two workspaces so that the dmu solve and the candidate's selected inverse
    // can run concurrently on the two streams
    CrPlan plan;
    int tile_threads = CR_THREADS;  // threads of a tile CTA (two co-resident tile CTAs per SM run 256 each)
    double* ws[2] = {nullptr, nullptr};
    double* ldsum[2] = {nullptr, nullptr};
    // multi-GPU (ctx->world > 1): the mid level (one "tile" over this rank's separator chain), the boundary exchange
    // buffers and the redundantly solved chain of rank boundaries -- one set per workspace slot
    CrPlan plan_mid, plan_top;
    bool three_level = false;                  // single GPU, very long chain: plan_mid tiles the separator chain
    double* ws_mid[2] = {nullptr, nullptr};
    double* ws_top[2] = {nullptr, nullptr};
    double* dist_buf[2] = {nullptr, nullptr};  // D1 | O1 | g1 | send | recv | Dt | Ot | gt | xt | cDt | cOt
    double* red_buf = nullptr;                 // [4 + 4 * world]: this rank's (cost, flag0, flag1, 0), then every rank's
    cudaStream_t stream2 = nullptr;  // side stream of the fork / join inside one iteration
    cudaStream_t ls = nullptr;       // stream the LAUNCH macro currently targets
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_pro = nullptr, ev_mu = nullptr, ev_k1 = nullptr;
    // schedule (GVIGH::optimize locals)
    int iter = 0;
    bool is_lowtemp = true, converged = false;
    bool sweep_valid = false;  // fcost/fVdmu/fVdd[cur] hold a full moment sweep at the current state
    const double* sweep_cO = nullptr;  // off-diagonal covariance blocks of the sweep being launched (fused culling + prologue)
    // batch of independent problems with a line search per problem (gvib200_set_batch / gvib200_batch_iterate)
    struct Batch {
        int P = 0, G = 0;
        std::vector<int> soff;                 // [P + 1] first state of every problem
        std::vector<int> sprob_h;              // [S] problem of every state
        int *d_sprob = nullptr, *d_soff = nullptr, *d_seg = nullptr, *d_gfirst = nullptr;
        double *d_step = nullptr, *d_alpha_node = nullptr, *d_ldn[2] = {nullptr, nullptr}, *d_ldn_solve = nullptr, *d_pcost = nullptr;
        double* h_pcost = nullptr;             // pinned [P]
        double *d_pcost2 = nullptr, *h_pcost2 = nullptr;  // per-problem log-pivot sums of the solve pass (SPD check of Vddmu)
        double* h_step = nullptr;              // pinned [P]
        bool ldn_valid[2] = {false, false};    // d_ldn[b] holds the log pivots of buffer b's precision
        bool cost_valid = false;               // cost_cur is the per-problem cost at the current state
        std::vector<double> cost_cur;
        std::vector<char> lowtemp, converged;
    } batch;
    bool pending_check = false;  // gvib200_set_state_async: the not-SPD flag of its selected inverse has not been read yet
    bool grads_valid = false;
    bool prox = false;              // Prox-GVI problem (option "prox" before finalize): linear factors get per-iteration
                                    // Vddmu blocks, no constant Klin, GH costs are not divided by a temperature
    bool flags_synced = false;      // multi-GPU: the not-SPD flags were all-reduced since the last chain pass
    bool force_generic_k1 = false;  // tests: run the generic node-loop kernel even where K1S applies
    // profiling: one CUDA event pair per launch while enabled
    bool profile = false;
    std::vector<ProfRec> prof;
    // timer (gvib200_timer_start / _stop)
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    // snapshot (gvib200_snapshot_save / _restore)
    double* snap = nullptr;
    size_t snap_doubles = 0;
    int snap_iter = 0, snap_cur = 0;
    bool snap_lowtemp = true, snap_sweep_valid = false;
    std::vector<char> snap_sr_full;  // per GH group: sr_full of the snapshot's buffer
};

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
template <class T>
static int dev_alloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    CUDA_TRY(cudaMalloc((void**)p, count * sizeof(T)));
    return 0;
}
template <class T>
static int dev_upload(T** p, const std::vector<T>& v, cudaStream_t st) {
    TRY(dev_alloc(p, v.size()));
    if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return 0;
}
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

static inline void prof_begin(gvib200_problem* p, int kc);
static inline void prof_end(gvib200_problem* p);
#define LAUNCH(prob, kc, kern, grid, block, smem, ...)                               \
    do {                                                                             \
        if ((prob)->profile) prof_begin((prob), (kc));                               \
        kern<<<(grid), (block), (smem), (prob)->ls>>>(__VA_ARGS__);              \
        if ((prob)->profile) prof_end((prob));                                       \
        (prob)->ctx->launches++;                                                     \
    } while (0)

static inline void prof_begin(gvib200_problem* p, int kc) {
    ProfRec r;
    r.kc = kc;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, p->ls);
    r.side = (p->ls != p->stream);
    p->prof.push_back(r);
}
static inline void prof_end(gvib200_problem* p) { cudaEventRecord(p->prof.back().b, p->ls); }

static int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GVIB200_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------------
static int get_table(gvib200_ctx* ctx, int dim, int deg, const Table** out) {
    auto key = std::make_pair(dim, deg);
    auto it = ctx->tables.find(key);
    if (it == ctx->tables.end()) {
        std::unique_ptr<Table> t(new Table);
        t->dim = dim;
        t->deg = deg;
        try {
            generate_spgh_table(dim, deg, t->nodes, t->w);
        } catch (const std::exception& e) {
            return fail(GVIB200_ENOTABLE, std::string("sparse GH table: ") + e.what());
        }
        t->n = (int)t->w.size();
        it = ctx->tables.emplace(key, std::move(t)).first;
    }
    Table* t = it->second.get();
    if (t->d_rows == nullptr) {
        // padded with zero-weight nodes at xi = 0 to a multiple of 32: the node loop needs no tail handling
        t->n_pad = (t->n + 31) & ~31;
        const int NP = (dim + 1) / 2;
        std::vector<double> rows((size_t)t->n_pad * (2 * NP + 1), 0.0);
        for (int c = 0; c < 12; ++c) t->ximax[c] = 0.0;
        for (int i = 0; i < t->n; ++i) {
            double nrm2 = 0.0;
            for (int c = 0; c < dim; ++c) {
                const double v = t->nodes[(size_t)i * dim + c];
                rows[(size_t)(c / 2) * 2 * t->n_pad + (size_t)2 * i + (c & 1)] = v;
                if (c < 12) t->ximax[c] = std::max(t->ximax[c], std::fabs(v));
                nrm2 += v * v;
            }
            t->xinorm = std::max(t->xinorm, std::sqrt(nrm2));
            rows[(size_t)2 * NP * t->n_pad + i] = t->w[i];
        }
        CUDA_TRY(cudaMalloc((void**)&t->d_rows, rows.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpy(t->d_rows, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice));
        if (dim > 4) {
            std::vector<int> gh;
            std::vector<double> gvv;
            t->grp_ok = build_grp_table(*t, gh, gvv, t->grp);
            if (t->grp_ok) {
                CUDA_TRY(cudaMalloc((void**)&t->d_grp_hdr, gh.size() * sizeof(int)));
                CUDA_TRY(cudaMalloc((void**)&t->d_grp_val, gvv.size() * sizeof(double)));
                CUDA_TRY(cudaMemcpy(t->d_grp_hdr, gh.data(), gh.size() * sizeof(int), cudaMemcpyHostToDevice));
                CUDA_TRY(cudaMemcpy(t->d_grp_val, gvv.data(), gvv.size() * sizeof(double), cudaMemcpyHostToDevice));
                t->grp.hdr = t->d_grp_hdr;
                t->grp.val = t->d_grp_val;
            }
        }
        t->sym_ok = build_sym_table(*t, t->sym);
        if (t->sym_ok) {
            CUDA_TRY(cudaMalloc((void**)&t->d_sym, t->sym_data.size() * sizeof(double)));
            CUDA_TRY(cudaMemcpy(t->d_sym, t->sym_data.data(), t->sym_data.size() * sizeof(double), cudaMemcpyHostToDevice));
            t->sym.data = t->d_sym;
        }
    }
    *out = t;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// chain engine drivers (tile-wise block cyclic reduction, bt_cr.h): three launches per pass
// ------------------------------------------------------------------------------------------------
template <class K>
static int cr_allow_smem(gvib200_ctx* ctx, K kern, size_t bytes) {
    if (!ctx->need_config(reinterpret_cast<const void*>(kern))) return 0;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return 0;
}

// layout of dist_buf (doubles) for P ranks, K local tiles
struct DistLayout {
    size_t D1, O1, g1, send, recv, Dt, Ot, gt, xt, cDt, cOt, total;
};
static DistLayout dist_layout(int D, int K, int P) {
    const size_t DD = (size_t)D * D, NB = 5 * DD + 4 * D;
    DistLayout L;
    size_t off = 0;
    auto take = [&](size_t n) {
        size_t o = off;
        off += (n + 1) & ~size_t(1);
        return o;
    };
    L.D1 = take((size_t)(K + 1) * DD);
    L.O1 = take((size_t)(K + 1) * DD);
    L.g1 = take((size_t)(K + 1) * D);
    L.send = take(NB);
    L.recv = take((size_t)P * NB);
    L.Dt = take((size_t)(P + 1) * DD);
    L.Ot = take((size_t)(P + 1) * DD);
    L.gt = take((size_t)(P + 1) * D);
    L.xt = take((size_t)(P + 1) * D);
    L.cDt = take((size_t)(P + 1) * DD);
    L.cOt = take((size_t)(P + 1) * DD);
    L.total = off;
    return L;
}

// Multi-GPU chain pass (bt_cr.h, "Multi-GPU"): tiles -> separator chain -> mid tile -> ONE all-gather of the boundary
// records -> chain of rank boundaries solved redundantly on every rank -> back down.  Everything on p->ls.
template <int D, bool RHS, bool SELINV>
static int chain_pass_dist(gvib200_problem* p, int slot, const CrArgs<D>& a, double* d_logdet) {
    gvib200_ctx* ctx = p->ctx;
    const CrPlan& pl = p->plan;
    const int P = ctx->world;
    const DistLayout L = dist_layout(D, pl.K, P);
    double* buf = p->dist_buf[slot];
    constexpr int NB = cr_boundary_doubles<D>();
    // the mid tile works on the summed separator chain and writes its results into the seeds of the real tiles
    CrArgs<D> mid = cr_bind<D>(p->plan_mid, p->ws_mid[slot], buf + L.D1, buf + L.O1, RHS ? buf + L.g1 : nullptr, a.tx, a.tD,
                               a.tO, a.notspd);
    CrArgs<D> top = cr_bind<D>(p->plan_top, p->ws_top[slot], buf + L.Dt, buf + L.Ot, RHS ? buf + L.gt : nullptr, buf + L.xt,
                               buf + L.cDt, buf + L.cOt, a.notspd);
    p->flags_synced = false;
    TRY(cr_allow_smem(ctx, k_cr_mid_forward<D, RHS>, ctx->smem_optin - 5120));
    TRY(cr_allow_smem(ctx, k_cr_dist_top<D, RHS, SELINV>, ctx->smem_optin - 5120));
    if (ctx->mbox_on) {
        // 3 launches, like on one GPU: tiles | ONE single-CTA kernel in shared memory (separator chain reduced to this
        // rank's two end nodes, boundary record pushed to the peers' mailboxes over NVLink, chain of rank boundaries,
        // back up the separator chain, log det) | tiles back
        TRY(cr_allow_smem(ctx, k_cr_dist_mid<D, RHS, SELINV>, ctx->smem_optin - 5120));
        constexpr size_t DDc = (size_t)D * D;
        const size_t smem_mid = (cr_top_doubles<D>(pl.K + 1) + cr_top_doubles<D>(P + 1) + (size_t)(P + 1) * NB + 4 * (P + 1) * DDc +
                                 2 * (size_t)(P + 1) * D + 32) * sizeof(double);
        if (smem_mid > ctx->smem_optin - 5120) return fail(GVIB200_EINVAL, "chain_pass_dist: separator chain too long for one CTA");
        const unsigned long long epoch = ++ctx->ep_rec[slot & 1];
        LAUNCH(p, KC_BT_FORWARD, (k_cr_tile_forward<D, RHS>), pl.K, p->tile_threads, pl.tile_smem_bytes, a);
        LAUNCH(p, KC_BT_TOP, (k_cr_dist_mid<D, RHS, SELINV>), 1, CR_THREADS, smem_mid, a, top, buf + L.send, buf + L.recv, buf + L.Dt,
               buf + L.Ot, buf + L.gt, ctx->peers, slot & 1, epoch, d_logdet);
        LAUNCH(p, KC_BT_BACK, (k_cr_tile_backward<D, RHS, SELINV>), pl.K, p->tile_threads, pl.tile_smem_bytes, a);
        return check_launch("chain_pass_dist (mailbox)");
    }
    // 4 launches + one all-gather: tiles | separator sum + mid tile + boundary record | all-gather | chain of rank
    // boundaries + seeds + mid tile back (+ log det) | tiles back
    LAUNCH(p, KC_BT_FORWARD, (k_cr_tile_forward<D, RHS>), pl.K, p->tile_threads, pl.tile_smem_bytes, a);
    LAUNCH(p, KC_BT_TOP, (k_cr_mid_forward<D, RHS>), 1, CR_THREADS, p->plan_mid.tile_smem_bytes, a, mid, buf + L.D1, buf + L.O1,
           buf + L.g1, buf + L.send);
    void* comm = (p->ls == p->stream2 && ctx->nccl_comm2) ? ctx->nccl_comm2 : ctx->nccl_comm;
    if (ctx->ncclAllGather(buf + L.send, buf + L.recv, (size_t)NB, /*ncclFloat64*/ 8, comm, p->ls) != 0)
        return fail(GVIB200_ENCCL, "chain pass: ncclAllGather failed");
    LAUNCH(p, KC_BT_TOP, (k_cr_dist_top<D, RHS, SELINV>), 1, CR_THREADS,
           std::max(p->plan_top.top_smem_bytes, p->plan_mid.tile_smem_bytes), P, ctx->rank, buf + L.recv, buf + L.Dt, buf + L.Ot,
           buf + L.gt, top, mid, a.ld, pl.K, d_logdet);
    LAUNCH(p, KC_BT_BACK, (k_cr_tile_backward<D, RHS, SELINV>), pl.K, p->tile_threads, pl.tile_smem_bytes, a);
    return check_launch("chain_pass_dist");
}

// Single GPU, chains too long for two levels: the separator chain (K + 1 nodes) is itself tiled (plan_mid), its top
// results seed the tiles of the first level.
template <int D, bool RHS, bool SELINV>
static int chain_pass_3level(gvib200_problem* p, int slot, const CrArgs<D>& a, double* d_logdet) {
    const CrPlan &pl = p->plan, &pm = p->plan_mid;
    const DistLayout L = dist_layout(D, pl.K, 1);
    double* buf = p->dist_buf[slot];
    CrArgs<D> mid = cr_bind<D>(pm, p->ws_mid[slot], buf + L.D1, buf + L.O1, RHS ? buf + L.g1 : nullptr, a.tx, a.tD, a.tO, a.notspd);
    if (a.ldnode != nullptr) {  // node i of the mid level is separator i of the first level = chain node min(i T, n - 1)
        mid.ldnode = a.ldnode;
        mid.ld_stride = (long long)pl.T * a.ld_stride;
        mid.ld_max = a.ld_max;
    }
    const bool batch = (a.ldnode != nullptr || a.alpha_node != nullptr);  // per-node log det / step sizes (gvib200_batch_iterate)
    if (batch) LAUNCH(p, KC_BT_FORWARD, (k_cr_tile_forward<D, RHS, true>), pl.K, p->tile_threads, pl.tile_smem_bytes, a);
    else LAUNCH(p, KC_BT_FORWARD, (k_cr_tile_forward<D, RHS>), pl.K, p->tile_threads, pl.tile_smem_bytes, a);
    LAUNCH(p, KC_OTHER, (k_cr_sum_level<D, RHS>), std::min(64, cdiv(pl.K + 1, 16)), 256, 0, a, buf + L.D1, buf + L.O1, buf + L.g1);
    if (batch) {
        if (pm.K > 0) LAUNCH(p, KC_BT_FORWARD, (k_cr_tile_forward<D, RHS, true>), pm.K, CR_THREADS, pm.tile_smem_bytes, mid);
        LAUNCH(p, KC_BT_TOP, (k_cr_top<D, RHS, SELINV, true>), 1, CR_THREADS, pm.top_smem_bytes, mid);
    } else {
        if (pm.K > 0) LAUNCH(p, KC_BT_FORWARD, (k_cr_tile_forward<D, RHS>), pm.K, CR_THREADS, pm.tile_smem_bytes, mid);
        LAUNCH(p, KC_BT_TOP, (k_cr_top<D, RHS, SELINV>), 1, CR_THREADS, pm.top_smem_bytes, mid);
    }
    if (pm.K > 0) LAUNCH(p, KC_BT_BACK, (k_cr_tile_backward<D, RHS, SELINV>), pm.K, CR_THREADS, pm.tile_smem_bytes, mid);
    LAUNCH(p, KC_BT_BACK, (k_cr_tile_backward<D, RHS, SELINV>), pl.K, p->tile_threads, pl.tile_smem_bytes, a);
    if (d_logdet) {
        LAUNCH(p, KC_SUM, k_sum, 1, 256, 0, (size_t)pm.ld_count, mid.ld, nullptr, 0.0, p->scal + 6);
        LAUNCH(p, KC_SUM, k_sum3, 1, 256, 0, (size_t)pl.K, a.ld, p->scal + 6, nullptr, d_logdet);
    }
    return check_launch("chain_pass_3level");
}

// optional candidate fusions of a chain pass (see CrArgs)
struct ChainFuse {
    const double *Dg2 = nullptr, *Og2 = nullptr;
    double alpha = 0.0;
    double *Dout = nullptr, *Oout = nullptr;
    const double* xbase = nullptr;
    double xalpha = 0.0;
    double* xout = nullptr;
    const double* alpha_node = nullptr;  // per-state step size (batches of independent problems)
    double* ldnode = nullptr;            // per-state log det of the pivot blocks
};

template <int D, bool RHS, bool SELINV>
static int chain_pass(gvib200_problem* p, int slot, const double* Dg, const double* Og, const double* rhs, double* x,
                      double* cD, double* cO, double* d_logdet, int* d_flag, const ChainFuse* fuse = nullptr) {
    const CrPlan& pl = p->plan;
    CrArgs<D> a = cr_bind<D>(pl, p->ws[slot], Dg, Og, rhs, x, cD, cO, d_flag);
    if (fuse) {
        a.Dg2 = fuse->Dg2;
        a.Og2 = fuse->Og2;
        a.alpha = fuse->alpha;
        a.Dout = fuse->Dout;
        a.Oout = fuse->Oout;
        a.xbase = fuse->xbase;
        a.xalpha = fuse->xalpha;
        a.xout = fuse->xout;
        a.alpha_node = fuse->alpha_node;
        a.ldnode = fuse->ldnode;
    }
    TRY(cr_allow_smem(p->ctx, k_cr_tile_forward<D, RHS>, p->ctx->smem_optin - 5120));
    TRY(cr_allow_smem(p->ctx, k_cr_top<D, RHS, SELINV>, p->ctx->smem_optin - 5120));
    const bool batch = (a.ldnode != nullptr || a.alpha_node != nullptr);  // per-node log det / step sizes (gvib200_batch_iterate)
    if (batch) {
        TRY(cr_allow_smem(p->ctx, k_cr_tile_forward<D, RHS, true>, p->ctx->smem_optin - 5120));
        TRY(cr_allow_smem(p->ctx, k_cr_top<D, RHS, SELINV, true>, p->ctx->smem_optin - 5120));
    }
    TRY(cr_allow_smem(p->ctx, k_cr_tile_backward<D, RHS, SELINV>, p->ctx->smem_optin - 5120));
    if (p->ctx->world > 1) return chain_pass_dist<D, RHS, SELINV>(p, slot, a, d_logdet);
    if (p->three_level) return chain_pass_3level<D, RHS, SELINV>(p, slot, a, d_logdet);
    a.ldout = d_logdet;  // the top kernel adds up the partial log determinants itself
    if (batch) {
        if (pl.K > 0) LAUNCH(p, KC_BT_FORWARD, (k_cr_tile_forward<D, RHS, true>), pl.K, p->tile_threads, pl.tile_smem_bytes, a);
        LAUNCH(p, KC_BT_TOP, (k_cr_top<D, RHS, SELINV, true>), 1, CR_THREADS, pl.top_smem_bytes, a);
    } else {
        if (pl.K > 0) LAUNCH(p, KC_BT_FORWARD, (k_cr_tile_forward<D, RHS>), pl.K, p->tile_threads, pl.tile_smem_bytes, a);
        LAUNCH(p, KC_BT_TOP, (k_cr_top<D, RHS, SELINV>), 1, CR_THREADS, pl.top_smem_bytes, a);
    }
    if (pl.K > 0) LAUNCH(p, KC_BT_BACK, (k_cr_tile_backward<D, RHS, SELINV>), pl.K, p->tile_threads, pl.tile_smem_bytes, a);
    return check_launch("chain_pass");
}

// selected inverse + log det of the block-tridiagonal (Dg, Og) -> (cD, cO), logdet scalar (device)
template <int D>
static int chain_selinv(gvib200_problem* p, int slot, const double* Dg, const double* Og, double* cD, double* cO,
                        double* d_logdet, int* d_flag, const ChainFuse* fuse = nullptr) {
    return chain_pass<D, false, true>(p, slot, Dg, Og, nullptr, nullptr, cD, cO, d_logdet, d_flag, fuse);
}

template <int D>
static int chain_solve(gvib200_problem* p, int slot, const double* Dg, const double* Og, const double* rhs, double* x,
                       double* d_logdet, int* d_flag, const ChainFuse* fuse = nullptr) {
    return chain_pass<D, true, false>(p, slot, Dg, Og, rhs, x, nullptr, nullptr, d_logdet, d_flag, fuse);
}

#define DISPATCH_D(d, CALL)                                                              \
    switch (d) {                                                                         \
        case 1: { constexpr int D_ = 1; CALL; } break;                                   \
        case 2: { constexpr int D_ = 2; CALL; } break;                                   \
        case 3: { constexpr int D_ = 3; CALL; } break;                                   \
        case 4: { constexpr int D_ = 4; CALL; } break;                                   \
        case 6: { constexpr int D_ = 6; CALL; } break;                                   \
        default: return fail(GVIB200_EINVAL, "unsupported state dimension (1, 2, 3, 4, 6)"); \
    }

// slot: which workspace (0: main stream, 1: side stream); the not-SPD flag of slot s is d_flag[s]
static int do_selinv(gvib200_problem* p, const double* Dg, const double* Og, double* cD, double* cO, double* d_logdet,
                     int slot = 0, const ChainFuse* fuse = nullptr) {
    int rc = 0;
    DISPATCH_D(p->d, rc = chain_selinv<D_>(p, slot, Dg, Og, cD, cO, d_logdet, p->d_flag + slot, fuse));
    return rc;
}
static int do_solve(gvib200_problem* p, const double* Dg, const double* Og, const double* rhs, double* x, double* d_logdet,
                    int slot = 0, const ChainFuse* fuse = nullptr) {
    int rc = 0;
    DISPATCH_D(p->d, rc = chain_solve<D_>(p, slot, Dg, Og, rhs, x, d_logdet, p->d_flag + slot, fuse));
    return rc;
}

// ------------------------------------------------------------------------------------------------
// quadrature sweeps
// ------------------------------------------------------------------------------------------------
template <int DIM, int SD>
static int launch_prologue(gvib200_problem* p, const GhGroup& g, const double* cD, const double* cO, double* SR) {
    if constexpr (DIM > 4) {  // JG lanes per factor: parallel Jacobi in shared memory
        constexpr int WD = 2 * DIM * DIM + 4 * ((DIM + 1) / 2) + 2 * DIM + 1;
        constexpr int GPB = 256 / JG;
        const size_t smem = (size_t)GPB * WD * sizeof(double);
        if (p->ctx->need_config(reinterpret_cast<const void*>(k_prologue_warp<DIM, SD>)))
            CUDA_TRY(cudaFuncSetAttribute(k_prologue_warp<DIM, SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(p, KC_PROLOGUE, (k_prologue_warp<DIM, SD>), cdiv(g.n, GPB), 256, smem, g.n, g.d_start, cD, cO, SR);
        return 0;
    }
    const int block = (DIM <= 4) ? 128 : 32;
    LAUNCH(p, KC_PROLOGUE, (k_prologue<DIM, SD>), cdiv(g.n, block), block, 0, g.n, g.d_start, cD, cO, SR);
    return 0;
}

// K1S (k1_sym.cuh): factors of dimension <= 4 on a rule with the sign-group structure
template <int DIM, class Cost>
static int launch_moments_sym(gvib200_problem* p, const GhGroup& g, const Cost& cost, const double* mu, const double* SR,
                              double* fcost, double* fVdmu, double* fVdd, double* raw, bool full, const double* covD) {
    SymArgs<Cost> a;
    a.n = g.n;
    a.state_dim = p->d;
    a.start = g.d_start;
    a.mu = mu;
    a.SR = SR;
    a.T = g.d_T;
    a.fcost = fcost + g.first_id;
    a.fVdmu = fVdmu + g.voff;
    a.fVdd = fVdd + g.moff;
    a.raw = raw;
    a.evaluated = p->d_evaluated;
    a.covD = covD;
    a.covO = p->sweep_cO;
    a.SR_out = const_cast<double*>(SR);
    a.xinorm = g.table->xinorm;
    a.active = nullptr;
    a.n_active = nullptr;
    a.done = nullptr;
    if constexpr (Cost::CULL) {
        if (p->cull) {
            if (g.d_active == nullptr) {
                TRY(dev_alloc(&g.d_active, (size_t)g.n));
                TRY(dev_alloc(&g.d_nactive, 2));  // [0] compacted count, [1] finished CTAs (both reset by the moment kernel)
                CUDA_TRY(cudaMemsetAsync(g.d_nactive, 0, 2 * sizeof(int), p->ls));
            }
            a.active = g.d_active;
            a.n_active = g.d_nactive;
            a.done = reinterpret_cast<unsigned*>(g.d_nactive + 1);
        }
    }
    for (int c = 0; c < 4; ++c) a.ximax[c] = g.table->ximax[c];
    a.cost = cost;
    const int grid = cdiv(g.n, K1S_FPC);
    const size_t smem = (size_t)g.table->sym.ndata * sizeof(double);
    if (p->ctx->need_config(reinterpret_cast<const void*>(k_moments_sym<DIM, Cost, true>))) {
        CUDA_TRY(cudaFuncSetAttribute(k_moments_sym<DIM, Cost, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      K1S_MAX_DATA * (int)sizeof(double)));
        CUDA_TRY(cudaFuncSetAttribute(k_moments_sym<DIM, Cost, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      K1S_MAX_DATA * (int)sizeof(double)));
    }
    if constexpr (Cost::CULL) if (a.active != nullptr) {
        if (!g.sr_full[(SR == g.d_SR[1]) ? 1 : 0]) {  // S, R of the kept factors come from the culling pass itself
            if (full) LAUNCH(p, KC_CULL, (k_cull_prologue_sym<DIM, Cost, true>), cdiv(g.n, 256), 256, 0, a);
            else LAUNCH(p, KC_CULL, (k_cull_prologue_sym<DIM, Cost, false>), cdiv(g.n, 256), 256, 0, a);
        } else {
            if (full) LAUNCH(p, KC_CULL, (k_cull_sym<DIM, Cost, true>), cdiv(g.n, 256), 256, 0, a);
            else LAUNCH(p, KC_CULL, (k_cull_sym<DIM, Cost, false>), cdiv(g.n, 256), 256, 0, a);
        }
    }
    if (p->profile) prof_begin(p, full ? KC_MOMENTS_FULL : KC_MOMENTS_COST);
    if (full) k_moments_sym<DIM, Cost, true><<<grid, K1S_THREADS, smem, p->ls>>>(g.table->sym, a);
    else k_moments_sym<DIM, Cost, false><<<grid, K1S_THREADS, smem, p->ls>>>(g.table->sym, a);
    if (p->profile) prof_end(p);
    p->ctx->launches++;
    return check_launch("k_moments_sym");
}

// K1G (k1_grp.cuh): factors of dimension > 4 on a rule made of sparse sign groups
template <int DIM, class Cost>
static int launch_moments_grp(gvib200_problem* p, const GhGroup& g, const Cost& cost, const double* mu, const double* SR,
                              double* fcost, double* fVdmu, double* fVdd, double* raw, bool full) {
    GrpArgs<Cost> a;
    a.n = g.n;
    a.state_dim = p->d;
    a.start = g.d_start;
    a.mu = mu;
    a.SR = SR;
    a.T = g.d_T;
    a.fcost = fcost + g.first_id;
    a.fVdmu = fVdmu + g.voff;
    a.fVdd = fVdd + g.moff;
    a.raw = raw;
    a.cost = cost;
    const GrpTable& tab = g.table->grp;
    const size_t table_doubles = (size_t)4 * tab.n_groups + (size_t)(tab.n_groups + 1) / 2;
    const size_t limit = std::min<size_t>(p->ctx->smem_optin, 227 * 1024) - 1024;
    int warps = K1G_WARPS;
    while (warps > 1 && (table_doubles + (size_t)warps * K1GCfg<DIM>::WARP_DOUBLES) * sizeof(double) > limit) --warps;
    const size_t smem = (table_doubles + (size_t)warps * K1GCfg<DIM>::WARP_DOUBLES) * sizeof(double);
    if (smem > limit) return fail(GVIB200_EINVAL, "K1G: the rule does not fit shared memory");
    if (p->ctx->need_config(reinterpret_cast<const void*>(k_moments_grp<DIM, Cost, true>))) {
        CUDA_TRY(cudaFuncSetAttribute(k_moments_grp<DIM, Cost, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
        CUDA_TRY(cudaFuncSetAttribute(k_moments_grp<DIM, Cost, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    }
    const int grid = std::max(1, std::min(cdiv(g.n, warps), p->ctx->sm_count));
    if (p->profile) prof_begin(p, full ? KC_MOMENTS_FULL : KC_MOMENTS_COST);
    if (full) k_moments_grp<DIM, Cost, true><<<grid, warps * 32, smem, p->ls>>>(tab, a);
    else k_moments_grp<DIM, Cost, false><<<grid, warps * 32, smem, p->ls>>>(tab, a);
    if (p->profile) prof_end(p);
    p->ctx->launches++;
    return check_launch("k_moments_grp");
}

template <int DIM, class Cost>
static int launch_moments(gvib200_problem* p, const GhGroup& g, const Cost& cost, const double* mu, const double* SR,
                          double* fcost, double* fVdmu, double* fVdd, double* raw, bool full, const double* covD) {
    if constexpr (DIM <= 4) {
        if (g.table->sym_ok && !p->force_generic_k1)
            return launch_moments_sym<DIM, Cost>(p, g, cost, mu, SR, fcost, fVdmu, fVdd, raw, full, covD);
    } else {
        if (g.table->grp_ok && !p->force_generic_k1)
            return launch_moments_grp<DIM, Cost>(p, g, cost, mu, SR, fcost, fVdmu, fVdd, raw, full);
    }
    constexpr int XD = Cost::XD;
    constexpr int ROW = 2 * ((DIM + 1) / 2) + 1;  // doubles per node over all planes
    constexpr int THREADS = K1Cfg<DIM>::THREADS;
    constexpr int WARPS = THREADS / 32;
    const size_t scratch = (size_t)WARPS * K1Scratch<DIM, XD>::DOUBLES * sizeof(double);
    // leave room for K1Cfg::MIN_BLOCKS resident CTAs when the whole table fits, else stream it in chunks
    const size_t budget = (std::min<size_t>(p->ctx->smem_optin, 227 * 1024) - 2048) / K1Cfg<DIM>::MIN_BLOCKS;
    int chunk = g.table->n_pad;
    size_t need = (size_t)chunk * ROW * sizeof(double) + scratch;
    if (need > budget) {
        chunk = (int)((budget - scratch) / (ROW * sizeof(double)));
        chunk &= ~31;
        need = (size_t)chunk * ROW * sizeof(double) + scratch;
    }
    MomentArgs<Cost> a;
    a.n = g.n;
    a.n_nodes = g.table->n_pad;
    a.chunk = chunk;
    for (int c = 0; c < 12; ++c) a.ximax[c] = g.table->ximax[c];
    a.state_dim = p->d;
    a.table = g.table->d_rows;
    a.start = g.d_start;
    a.mu = mu;
    a.SR = SR;
    a.T = g.d_T;
    a.fcost = fcost + g.first_id;
    a.fVdmu = fVdmu + g.voff;
    a.fVdd = fVdd + g.moff;
    a.raw = raw;
    a.cost = cost;
    auto kfull = k_moments<DIM, Cost, true>;
    auto kcost = k_moments<DIM, Cost, false>;
    auto kern = full ? kfull : kcost;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    int per_sm = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, need));
    if (per_sm < 1) per_sm = 1;
    int grid = std::min(cdiv(g.n, WARPS), p->ctx->sm_count * per_sm);
    if (grid < 1) grid = 1;
    if (p->profile) prof_begin(p, full ? KC_MOMENTS_FULL : KC_MOMENTS_COST);
    kern<<<grid, THREADS, need, p->ls>>>(a);
    if (p->profile) prof_end(p);
    p->ctx->launches++;
    return check_launch("k_moments");
}

// hinge records + the coarse free-space bound for the threshold thr (rebuilt when the threshold changes)
static int ensure_sdf_records(gvib200_problem* p, double thr) {
    if (p->sdf_rec_valid && p->sdf_thr == thr) return 0;
    const long long ncell = (long long)p->sdf_rows * p->sdf_cols;
    LAUNCH(p, KC_OTHER, k_build_sdf_records, cdiv(ncell, 256), 256, 0, p->sdf_rows, p->sdf_cols, thr, p->d_sdf_data, p->d_sdf_rec);
    const int crows = (p->sdf_rows + 3) / 4, ccols = (p->sdf_cols + 3) / 4;
    LAUNCH(p, KC_OTHER, k_build_sdf_coarse, cdiv((long long)crows * ccols, 128), 128, 0, p->sdf_rows, p->sdf_cols, thr, p->d_sdf_data,
           p->d_sdf_coarse);
    p->sdf_thr = thr;
    p->sdf_rec_valid = true;
    return check_launch("sdf records");
}

struct SweepTarget {
    const double* mu;
    const double* cD;
    const double* cO;
    int which;  // buffer index for SR / fcost / fVdmu / fVdd
};

template <int DIM, int SD>
static int gh_group_run(gvib200_problem* p, GhGroup& g, const SweepTarget& t, bool prologue, bool sweep, bool full,
                        double* raw) {
    double* SR = g.d_SR[t.which];
    // Hinge factors on the sign-group kernel with free-space culling: the culling pass of the sweep forms S, R itself, for the
    // factors it keeps only (k_cull_prologue_sym) -- no separate prologue launch.  d_SR is then incomplete (sr_full), which
    // only matters to a later sweep that runs without the fusion: it runs the full prologue first.
    static const bool no_fused_env = getenv("GVIB200_NO_FUSED_PROLOGUE") != nullptr;  // development switch
    const bool fused = g.kind == GVIB200_COST_PLANAR_HINGE && DIM >= 2 && DIM <= 4 && p->cull && !p->prox && g.table != nullptr &&
                       g.table->sym_ok && !p->force_generic_k1 && !no_fused_env;
    if (prologue) {
        if (fused) {
            g.sr_full[t.which] = false;
        } else {
            TRY((launch_prologue<DIM, SD>(p, g, t.cD, t.cO, SR)));
            g.sr_full[t.which] = true;
        }
    } else if (sweep && !fused && !g.sr_full[t.which]) {
        TRY((launch_prologue<DIM, SD>(p, g, t.cD, t.cO, SR)));
        g.sr_full[t.which] = true;
    }
    if (!sweep) return check_launch("k_prologue");
    p->sweep_cO = t.cO;
    double* fc = p->fcost[t.which];
    double* fv = p->fVdmu[t.which];
    double* fm = p->fVdd[t.which];
    switch (g.kind) {
        case GVIB200_COST_STEREO_1D: {
            if constexpr (DIM == 1) {
                const auto* hp = reinterpret_cast<const gvib200_stereo1d_params*>(g.params.data());
                CostStereo1D c;
                c.mu_p = hp->mu_p;
                c.fb = hp->f * hp->b;
                c.sig_p_sq = hp->sig_p_sq;
                c.sig_r_sq = hp->sig_r_sq;
                c.y = hp->f * hp->b / hp->mu_p + hp->y_offset;
                return launch_moments<DIM>(p, g, c, t.mu, SR, fc, fv, fm, raw, full, t.cD);
            }
            break;
        }
        case GVIB200_COST_PLANAR_HINGE: {
            if constexpr (DIM >= 2 && DIM <= 4) {
                if (p->d_sdf_rec == nullptr) return fail(GVIB200_ESTATE, "planar hinge cost needs gvib200_set_planar_sdf");
                const auto* hp = reinterpret_cast<const gvib200_hinge_params*>(g.params.data());
                const double thr = hp->epsilon + hp->radius;
                TRY(ensure_sdf_records(p, thr));
                CostPlanarHinge c;
                c.coarse = p->d_sdf_coarse;
                c.crows = (p->sdf_rows + 3) / 4;
                c.rec = p->d_sdf_rec;
                c.rows = p->sdf_rows;
                c.cols = p->sdf_cols;
                c.ox = p->sdf_ox;
                c.oy = p->sdf_oy;
                c.xmax = p->sdf_ox + (p->sdf_cols - 1.0) * p->sdf_cell;
                c.ymax = p->sdf_oy + (p->sdf_rows - 1.0) * p->sdf_cell;
                c.inv_cell = 1.0 / p->sdf_cell;
                c.cx0 = -p->sdf_ox * c.inv_cell;
                c.cy0 = -p->sdf_oy * c.inv_cell;
                c.thr = hp->epsilon + hp->radius;
                c.sigma = hp->sigma;
                return launch_moments<DIM>(p, g, c, t.mu, SR, fc, fv, fm, raw, full, t.cD);
            }
            break;
        }
        case GVIB200_COST_QUAD_HINGE: {
            if constexpr (DIM == 6) {  // planar quadrotor state (x, z, phi, vx, vz, vphi)
                if (p->d_sdf_rec == nullptr) return fail(GVIB200_ESTATE, "quadrotor hinge cost needs gvib200_set_planar_sdf");
                const auto* hp = reinterpret_cast<const gvib200_hinge_params*>(g.params.data());
                const double thr = hp->epsilon + hp->radius;
                TRY(ensure_sdf_records(p, thr));
                CostQuadHinge c;
                c.h.rec = p->d_sdf_rec;
                c.h.rows = p->sdf_rows;
                c.h.cols = p->sdf_cols;
                c.h.ox = p->sdf_ox;
                c.h.oy = p->sdf_oy;
                c.h.xmax = p->sdf_ox + (p->sdf_cols - 1.0) * p->sdf_cell;
                c.h.ymax = p->sdf_oy + (p->sdf_rows - 1.0) * p->sdf_cell;
                c.h.inv_cell = 1.0 / p->sdf_cell;
                c.h.cx0 = -p->sdf_ox * c.h.inv_cell;
                c.h.cy0 = -p->sdf_oy * c.h.inv_cell;
                c.h.thr = thr;
                c.h.sigma = hp->sigma;
                c.radius = hp->radius;
                return launch_moments<DIM>(p, g, c, t.mu, SR, fc, fv, fm, raw, full, t.cD);
            }
            break;
        }
        case GVIB200_COST_HINGE_3D: {
            if constexpr (DIM == 3 || DIM == 6) {
                if (p->d_sdf3 == nullptr) return fail(GVIB200_ESTATE, "3-D hinge cost needs gvib200_set_sdf3d");
                const auto* hp = reinterpret_cast<const gvib200_hinge_params*>(g.params.data());
                CostHinge3D c;
                c.data = p->d_sdf3;
                c.rows = p->sdf3_rows;
                c.cols = p->sdf3_cols;
                c.nz = p->sdf3_nz;
                c.ox = p->sdf3_o[0];
                c.oy = p->sdf3_o[1];
                c.oz = p->sdf3_o[2];
                c.xmax = c.ox + (c.cols - 1.0) * p->sdf3_cell;
                c.ymax = c.oy + (c.rows - 1.0) * p->sdf3_cell;
                c.zmax = c.oz + (c.nz - 1.0) * p->sdf3_cell;
                c.inv_cell = 1.0 / p->sdf3_cell;
                c.thr = hp->epsilon + hp->radius;
                c.sigma = hp->sigma;
                return launch_moments<DIM>(p, g, c, t.mu, SR, fc, fv, fm, raw, full, t.cD);
            }
            break;
        }
        case GVIB200_COST_ARM_3D: {
            if constexpr (DIM == SD && (DIM == 4 || DIM == 6)) {
                if (p->d_sdf3 == nullptr) return fail(GVIB200_ESTATE, "arm cost needs gvib200_set_sdf3d");
                const auto* ap = reinterpret_cast<const gvib200_arm_params*>(g.params.data());
                CostArm3D<DIM / 2> c;
                c.field.data = p->d_sdf3;
                c.field.rows = p->sdf3_rows;
                c.field.cols = p->sdf3_cols;
                c.field.nz = p->sdf3_nz;
                c.field.ox = p->sdf3_o[0];
                c.field.oy = p->sdf3_o[1];
                c.field.oz = p->sdf3_o[2];
                c.field.xmax = c.field.ox + (c.field.cols - 1.0) * p->sdf3_cell;
                c.field.ymax = c.field.oy + (c.field.rows - 1.0) * p->sdf3_cell;
                c.field.zmax = c.field.oz + (c.field.nz - 1.0) * p->sdf3_cell;
                c.field.inv_cell = 1.0 / p->sdf3_cell;
                c.field.thr = 0.0;
                c.field.sigma = 0.0;
                c.sigma = ap->sigma;
                c.epsilon = ap->epsilon;
                c.n_spheres = ap->n_spheres;
                for (int j = 0; j < ARM_MAX_DOF; ++j) {
                    const bool on = j < ap->n_dof;
                    c.a[j] = on ? ap->a[j] : 0.0;
                    c.d[j] = on ? ap->d[j] : 0.0;
                    c.bias[j] = on ? ap->theta_bias[j] : 0.0;
                    // cosf / sinf of alpha (helpers/CudaOperation.h:394-400), correctly rounded single precision
                    c.ca[j] = on ? (double)(float)std::cos((double)(float)ap->alpha[j]) : 1.0;
                    c.sa[j] = on ? (double)(float)std::sin((double)(float)ap->alpha[j]) : 0.0;
                }
                for (int i = 0; i < ARM_MAX_SPHERES; ++i) {
                    const bool on = i < ap->n_spheres;
                    c.frame[i] = on ? ap->frames[i] : 0;
                    c.radius[i] = on ? ap->radii[i] : 0.0;
                    for (int k = 0; k < 3; ++k) c.centre[i][k] = on ? ap->centers[i][k] : 0.0;
                }
                return launch_moments<DIM>(p, g, c, t.mu, SR, fc, fv, fm, raw, full, t.cD);
            }
            break;
        }
        case GVIB200_COST_LINEAR_GP: {
            if constexpr (DIM % 2 == 0 && DIM == 2 * SD) {
                CostLinearGP<DIM / 2> c;
                c.params = reinterpret_cast<const double*>(g.d_params);
                return launch_moments<DIM>(p, g, c, t.mu, SR, fc, fv, fm, raw, full, t.cD);
            }
            break;
        }
        case GVIB200_COST_FIXED_GP: {
            if constexpr (DIM == SD) {
                CostFixedGP<DIM> c;
                c.params = reinterpret_cast<const double*>(g.d_params);
                return launch_moments<DIM>(p, g, c, t.mu, SR, fc, fv, fm, raw, full, t.cD);
            }
            break;
        }
        case GVIB200_COST_QUADRATIC: {
            if constexpr (DIM <= 4) {
                CostQuadratic<DIM> c;
                c.c = *reinterpret_cast<const double*>(g.params.data());
                return launch_moments<DIM>(p, g, c, t.mu, SR, fc, fv, fm, raw, full, t.cD);
            }
            break;
        }
        default: break;
    }
    return fail(GVIB200_EINVAL, "unsupported (cost kind, factor dim, state dim) combination");
}

static int gh_group_dispatch(gvib200_problem* p, GhGroup& g, const SweepTarget& t, bool prologue, bool sweep, bool full,
                             double* raw) {
    const int dim = g.dim, sd = p->d;
#define GH_CASE(DIM_, SD_) \
    if (dim == DIM_ && sd == SD_) return gh_group_run<DIM_, SD_>(p, g, t, prologue, sweep, full, raw);
    GH_CASE(1, 1)
    GH_CASE(2, 1)
    GH_CASE(2, 2)
    GH_CASE(3, 3)
    GH_CASE(4, 2)
    GH_CASE(4, 4)
    GH_CASE(6, 6)
    GH_CASE(8, 4)
    GH_CASE(12, 6)
#undef GH_CASE
    return fail(GVIB200_EINVAL, "unsupported (factor dim, state dim) combination");
}

template <int DIM>
static void launch_raw_to_x(gvib200_problem* p, const GhGroup& g, const double* SR, double* E0, double* E1, double* E2) {
    LAUNCH(p, KC_OTHER, (k_raw_to_x<DIM>), cdiv(g.n, 128), 128, 0, g.n, g.d_raw, SR, E0, E1, E2);
}

static LinearArgs linear_args(gvib200_problem* p, const LinGroup& g, const SweepTarget& t, bool full) {
    LinearArgs a;
    a.n = g.n;
    a.dim = g.dim;
    a.m = g.m;
    a.state_dim = p->d;
    a.start = g.d_start;
    a.Lambda = g.d_Lambda;
    a.psi = g.d_psi;
    a.Kinv = g.d_Kinv;
    a.A = g.d_A;
    a.C = g.d_C;
    a.T = g.d_T;
    a.mu = t.mu;
    a.covD = t.cD;
    a.covO = t.cO;
    a.fcost = p->fcost[t.which] + g.first_id;
    a.fVdmu = full ? p->fVdmu[t.which] + g.voff : nullptr;
    a.part = 3;
    return a;
}

// parts: 1 = covariance part, 2 = mean part, 3 = both (one launch per group, same arithmetic)
static int run_linear(gvib200_problem* p, const SweepTarget& t, bool full, int parts = 3) {
  for (int part = 1; part <= 3; ++part) {
    if (parts == 3 ? part != 3 : !(parts & part) || part == 3) continue;
    for (auto& g : p->lin) {
        LinearArgs a;
        a.part = part;
        a.n = g.n;
        a.dim = g.dim;
        a.m = g.m;
        a.state_dim = p->d;
        a.start = g.d_start;
        a.Lambda = g.d_Lambda;
        a.psi = g.d_psi;
        a.Kinv = g.d_Kinv;
        a.A = g.d_A;
        a.C = g.d_C;
        a.T = g.d_T;
        a.mu = t.mu;
        a.covD = t.cD;
        a.covO = t.cO;
        a.fcost = p->fcost[t.which] + g.first_id;
        a.fVdmu = full ? p->fVdmu[t.which] + g.voff : nullptr;
        const int sd = p->d;
        const int grid = cdiv(g.n, 128);
        if (g.dim == 8 && g.m == 4 && sd == 4) LAUNCH(p, KC_LINEAR, (k_linear<8, 4, 4>), grid, 128, 0, a);
        else if (g.dim == 4 && g.m == 4 && sd == 4) LAUNCH(p, KC_LINEAR, (k_linear<4, 4, 4>), grid, 128, 0, a);
        else if (g.dim == 12 && g.m == 6 && sd == 6) LAUNCH(p, KC_LINEAR, (k_linear<12, 6, 6>), grid, 128, 0, a);
        else if (g.dim == 6 && g.m == 6 && sd == 6) LAUNCH(p, KC_LINEAR, (k_linear<6, 6, 6>), grid, 128, 0, a);
        else LAUNCH(p, KC_LINEAR, (k_linear<0, 0, 0>), grid, 128, 0, a);
    }
  }
    return check_launch("k_linear");
}

// prologue + sweep over every factor at (mu, cov) of buffer `which`
static int run_sweep(gvib200_problem* p, int which, bool prologue, bool full, bool want_raw, bool with_linear = true) {
    SweepTarget t{p->mu[which], p->CD[which], p->CO[which], which};
    for (auto& g : p->gh) {
        double* raw = nullptr;
        if (want_raw) {
            if (g.d_raw == nullptr) TRY(dev_alloc(&g.d_raw, (size_t)g.n * (1 + g.dim + g.dim * g.dim)));
            raw = g.d_raw;
        }
        TRY(gh_group_dispatch(p, g, t, prologue, true, full, raw));
    }
    if (with_linear) TRY(run_linear(p, t, full));
    return 0;
}

static int run_prologue_only(gvib200_problem* p, int which) {
    SweepTarget t{p->mu[which], p->CD[which], p->CO[which], which};
    for (auto& g : p->gh) TRY(gh_group_dispatch(p, g, t, true, false, false, nullptr));
    return 0;
}

// multi-GPU: sum the cost of buffer `which` over the ranks and make the not-SPD flags global (one small all-reduce)
static int dist_reduce(gvib200_problem* p, double* d_cost) {
    gvib200_ctx* ctx = p->ctx;
    if (ctx->world <= 1) return 0;
    LAUNCH(p, KC_OTHER, k_red_pack, 1, 1, 0, d_cost, p->d_flag, p->red_buf);
    // an all-gather of the four doubles (cheaper than an all-reduce at this size) + a sum in rank order on every rank
    if (ctx->ncclAllGather(p->red_buf, p->red_buf + 4, 4, /*ncclFloat64*/ 8, ctx->nccl_comm, p->ls) != 0)
        return fail(GVIB200_ENCCL, "ncclAllGather (cost) failed");
    LAUNCH(p, KC_OTHER, k_red_unpack, 1, 1, 0, ctx->world, p->red_buf + 4, d_cost, p->d_flag, (double*)nullptr, 0);
    p->flags_synced = true;
    return check_launch("dist_reduce");
}

// total cost of buffer `which`: sum of factor costs + logdet/2 (GVI-GH-GBP-impl.h:217-239) -> scal[2 + which]
static void run_total(gvib200_problem* p, int which) {
    const size_t n = (size_t)p->n_factors;
    if (n > 8192) {
        const int nb = (int)std::min<size_t>(1024, (n + 1023) / 1024);  // <= 4 elements per thread
        const bool mailbox = (p->ctx->world > 1 && p->ctx->mbox_on);
        const bool single = (p->ctx->world == 1) || mailbox;
        MboxPeers peers = p->ctx->peers;
        unsigned long long epoch = 0;
        if (mailbox) epoch = ++p->ctx->ep_cost;
        else peers.world = 1;
        LAUNCH(p, KC_SUM, k_total, nb, 256, 0, n, p->fcost[which], p->partial, p->d_counter, p->scal + which, 0.5,
               p->scal + 2 + which, p->d_flag, single ? p->zc_dev : nullptr, which, single ? nullptr : p->red_buf, peers, epoch);
        p->zc_ok[which] = true;
        if (mailbox) p->flags_synced = true;  // the cost exchange made the not-SPD flags global
        if (!single) {  // sum over the ranks, flags made global; the unpack kernel hands the result to the host (mapped memory)
            gvib200_ctx* ctx = p->ctx;
            if (ctx->ncclAllGather(p->red_buf, p->red_buf + 4, 4, /*ncclFloat64*/ 8, ctx->nccl_comm, p->ls) != 0) {
                fail(GVIB200_ENCCL, "ncclAllGather (cost) failed");
                p->zc_ok[which] = false;
                return;
            }
            LAUNCH(p, KC_OTHER, k_red_unpack, 1, 1, 0, ctx->world, p->red_buf + 4, p->scal + 2 + which, p->d_flag, p->zc_dev, which);
            p->flags_synced = true;
        }
        return;
    } else {
        LAUNCH(p, KC_SUM, k_sum, 1, 1024, 0, n, p->fcost[which], p->scal + which, 0.5, p->scal + 2 + which);
        p->zc_ok[which] = false;
    }
    dist_reduce(p, p->scal + 2 + which);
}

template <int D>
static void launch_assemble(gvib200_problem* p, int which, bool alt) {
    if constexpr (D % 2 == 0) {
        static const bool no_fast_env = getenv("GVIB200_NO_FAST_ASSEMBLE") != nullptr;  // development switch
        if (p->ell_nv >= 0 && p->ell_nv <= 4 && p->ell_nd <= 1 && p->ell_no == 0 && !no_fast_env) {
            double *oV = alt ? p->Vdmu2 : p->Vdmu, *oD = alt ? p->VD2 : p->VD, *oO = alt ? p->VO2 : p->VO, *oR = alt ? p->rhs2 : p->rhs;
            if (p->vo_alias) oO = nullptr;  // VO / VO2 ARE KlinO (finalize): nothing to copy
            const int grid = cdiv((long long)p->S * D, 128);
            if (p->ell_nd == 1)
                LAUNCH(p, KC_ASSEMBLE, (k_assemble_ell_fast<D, 4, 1>), grid, 128, 0, p->S, p->ell_v, p->ell_d, p->ell_dl, p->fVdmu[which],
                       p->fVdd[which], p->KlinD, p->KlinO, oV, oD, oO, oR);
            else
                LAUNCH(p, KC_ASSEMBLE, (k_assemble_ell_fast<D, 4, 0>), grid, 128, 0, p->S, p->ell_v, p->ell_d, p->ell_dl, p->fVdmu[which],
                       p->fVdd[which], p->KlinD, p->KlinO, oV, oD, oO, oR);
            return;
        }
    }
    if (p->ell_nv >= 0) {
        LAUNCH(p, KC_ASSEMBLE, (k_assemble_ell<D>), cdiv((long long)p->S * D, 128), 128, 0, p->S, p->ell_nv, p->ell_nd, p->ell_no, p->ell_v,
               p->ell_d, p->ell_dl, p->ell_o, p->ell_ol, p->fVdmu[which], p->fVdd[which], p->KlinD, p->KlinO, alt ? p->Vdmu2 : p->Vdmu,
               alt ? p->VD2 : p->VD, alt ? p->VO2 : p->VO, alt ? p->rhs2 : p->rhs);
        return;
    }
    LAUNCH(p, KC_ASSEMBLE, (k_assemble<D>), cdiv((long long)p->S * D * D, 256), 256, 0, p->S, p->vptr, p->voff, p->dptr, p->doff, p->dld, p->optr,
           p->ooff, p->old, p->fVdmu[which], p->fVdd[which], p->KlinD, p->KlinO, alt ? p->Vdmu2 : p->Vdmu, alt ? p->VD2 : p->VD,
           alt ? p->VO2 : p->VO, alt ? p->rhs2 : p->rhs);
}

// d_flag[0] / d_flag[1]: not-SPD flags of the chain passes run in workspace slot 0 / 1
static int dist_reduce(gvib200_problem* p, double* d_cost);
static int dispatch_assemble(gvib200_problem* p, int which, bool alt = false);
// spec_which >= 0: the assembly of that buffer's sweep is enqueued behind the copies; the host waits for the copies only
static int read_flags(gvib200_problem* p, int* flag0, int* flag1, int spec_which = -1, bool zc = false) {
    if (p->ctx->world > 1 && !p->flags_synced) TRY(dist_reduce(p, p->scal + 7));  // scal[7]: scratch
    if (!zc) CUDA_TRY(cudaMemcpyAsync(p->h_flag, p->d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    if (spec_which >= 0) {
        CUDA_TRY(cudaEventRecord(p->ev_host, p->stream));
        TRY(dispatch_assemble(p, spec_which, true));
        CUDA_TRY(cudaEventSynchronize(p->ev_host));
    } else {
        CUDA_TRY(cudaStreamSynchronize(p->stream));
    }
    if (zc) {
        p->h_flag[0] = p->zc[2] != 0.0;
        p->h_flag[1] = p->zc[3] != 0.0;
    }
    *flag0 = p->h_flag[0];
    *flag1 = p->h_flag[1];
    return 0;
}
static int read_flag(gvib200_problem* p, int* flag) {
    int f0 = 0, f1 = 0;
    TRY(read_flags(p, &f0, &f1));
    *flag = f0 | f1;
    return 0;
}
static int clear_flag(gvib200_problem* p) {
    CUDA_TRY(cudaMemsetAsync(p->d_flag, 0, 2 * sizeof(int), p->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C-ABI: context / tables
// ------------------------------------------------------------------------------------------------
extern "C" const char* gvib200_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* gvib200_version(void) { return "gvib200 0.1 sm_100a"; }

extern "C" int gvib200_ctx_create(int device, gvib200_ctx** out) {
    if (!out) return fail(GVIB200_EINVAL, "ctx_create: null out");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(GVIB200_ECUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(GVIB200_EINVAL, "ctx_create: bad device index");
    CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<gvib200_ctx> c(new gvib200_ctx);
    c->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    *out = c.release();
    return 0;
}

extern "C" int gvib200_ctx_destroy(gvib200_ctx* ctx) {
    if (!ctx) return 0;
    for (auto& kv : ctx->tables)
        if (kv.second->d_rows) cudaFree(kv.second->d_rows);
    for (auto& kv : ctx->tables)
        if (kv.second->d_sym) cudaFree(kv.second->d_sym);
    for (auto& kv : ctx->tables) {
        if (kv.second->d_grp_hdr) cudaFree(kv.second->d_grp_hdr);
        if (kv.second->d_grp_val) cudaFree(kv.second->d_grp_val);
    }
    for (int r = 0; r < MBOX_RANKS; ++r)
        if (ctx->peer_mapped[r]) cudaIpcCloseMemHandle(ctx->peer_mapped[r]);
    if (ctx->mbox) cudaFree(ctx->mbox);
    delete ctx;
    return 0;
}

extern "C" int gvib200_ctx_set_comm(gvib200_ctx* ctx, void* nccl_comm, int rank, int world, const char* libnccl_path) {
    if (!ctx || world < 1 || rank < 0 || rank >= world) return fail(GVIB200_EINVAL, "ctx_set_comm: bad arguments");
    if (world > 1) {
        if (!nccl_comm) return fail(GVIB200_EINVAL, "ctx_set_comm: null communicator");
        const char* path = (libnccl_path && libnccl_path[0]) ? libnccl_path : "libnccl.so.2";
        void* lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return fail(GVIB200_ENCCL, std::string("ctx_set_comm: dlopen(") + path + "): " + dlerror());
        ctx->nccl_lib = lib;
        ctx->ncclAllGather = reinterpret_cast<decltype(ctx->ncclAllGather)>(dlsym(lib, "ncclAllGather"));
        ctx->ncclAllReduce = reinterpret_cast<decltype(ctx->ncclAllReduce)>(dlsym(lib, "ncclAllReduce"));
        if (!ctx->ncclAllGather || !ctx->ncclAllReduce) return fail(GVIB200_ENCCL, "ctx_set_comm: NCCL symbols not found");
        // a second communicator over the same ranks (collective call: every rank is inside ctx_set_comm) so that the dmu
        // solve and the candidate's selected inverse can issue their boundary all-gathers from two streams
        using split_fn = int (*)(void*, int, int, void**, void*);
        split_fn split = reinterpret_cast<split_fn>(dlsym(lib, "ncclCommSplit"));
        ctx->nccl_comm2 = nullptr;
        if (split && !getenv("GVIB200_NO_FORK")) {
            void* c2 = nullptr;
            if (split(nccl_comm, 0, rank, &c2, nullptr) == 0) ctx->nccl_comm2 = c2;
        }
    }
    ctx->nccl_comm = nccl_comm;
    ctx->rank = rank;
    ctx->world = world;
    return 0;
}

extern "C" int gvib200_ctx_mailbox_create(gvib200_ctx* ctx, void* handle_out, size_t handle_capacity) {
    if (!ctx || !handle_out || handle_capacity < sizeof(cudaIpcMemHandle_t))
        return fail(GVIB200_EINVAL, "mailbox_create: bad arguments (the handle needs 64 bytes)");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!ctx->mbox) {
        CUDA_TRY(cudaMalloc((void**)&ctx->mbox, MBOX_DOUBLES * sizeof(double)));
        CUDA_TRY(cudaMemset(ctx->mbox, 0, MBOX_DOUBLES * sizeof(double)));
        CUDA_TRY(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, ctx->mbox));
    std::memcpy(handle_out, &h, sizeof(h));
    return (int)sizeof(h);
}

extern "C" int gvib200_ctx_mailbox_connect(gvib200_ctx* ctx, int world, int rank, const void* handles, size_t handle_stride) {
    if (!ctx || !handles || world < 2 || world > MBOX_RANKS || rank < 0 || rank >= world ||
        handle_stride < sizeof(cudaIpcMemHandle_t))
        return fail(GVIB200_EINVAL, "mailbox_connect: bad arguments (2..16 ranks)");
    if (!ctx->mbox) return fail(GVIB200_ESTATE, "mailbox_connect: call gvib200_ctx_mailbox_create first");
    if (ctx->world != world || ctx->rank != rank) return fail(GVIB200_ESTATE, "mailbox_connect: call gvib200_ctx_set_comm first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            ctx->peers.p[r] = ctx->mbox;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char*)handles + (size_t)r * handle_stride, sizeof(h));
        void* q = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(GVIB200_ECUDA, std::string("mailbox_connect: cudaIpcOpenMemHandle (rank ") + std::to_string(r) +
                                           "): " + cudaGetErrorString(e));
        }
        ctx->peer_mapped[r] = q;
        ctx->peers.p[r] = (double*)q;
    }
    ctx->peers.world = world;
    ctx->peers.rank = rank;
    ctx->mbox_on = !getenv("GVIB200_NO_MAILBOX");
    return 0;
}

extern "C" int gvib200_table_size(int dim, int deg) {
    std::vector<double> n, w;
    try {
        generate_spgh_table(dim, deg, n, w);
    } catch (const std::exception& e) {
        return fail(GVIB200_ENOTABLE, e.what());
    }
    return (int)w.size();
}

extern "C" int gvib200_table_generate(int dim, int deg, double* nodes_rowmajor, double* weights, int capacity) {
    std::vector<double> n, w;
    try {
        generate_spgh_table(dim, deg, n, w);
    } catch (const std::exception& e) {
        return fail(GVIB200_ENOTABLE, e.what());
    }
    if ((int)w.size() > capacity) return fail(GVIB200_EINVAL, "table_generate: capacity too small");
    if (nodes_rowmajor) std::memcpy(nodes_rowmajor, n.data(), n.size() * sizeof(double));
    if (weights) std::memcpy(weights, w.data(), w.size() * sizeof(double));
    return (int)w.size();
}

extern "C" int gvib200_table_set(gvib200_ctx* ctx, int dim, int deg, int n, const double* nodes_rowmajor,
                                 const double* weights) {
    if (!ctx || dim < 1 || n < 1 || !nodes_rowmajor || !weights) return fail(GVIB200_EINVAL, "table_set: bad arguments");
    std::unique_ptr<Table> t(new Table);
    t->dim = dim;
    t->deg = deg;
    t->n = n;
    t->nodes.assign(nodes_rowmajor, nodes_rowmajor + (size_t)n * dim);
    t->w.assign(weights, weights + n);
    auto key = std::make_pair(dim, deg);
    auto it = ctx->tables.find(key);
    if (it != ctx->tables.end() && it->second->d_rows) cudaFree(it->second->d_rows);
    if (it != ctx->tables.end() && it->second->d_sym) cudaFree(it->second->d_sym);
    if (it != ctx->tables.end() && it->second->d_grp_hdr) cudaFree(it->second->d_grp_hdr);
    if (it != ctx->tables.end() && it->second->d_grp_val) cudaFree(it->second->d_grp_val);
    ctx->tables[key] = std::move(t);
    return 0;
}

extern "C" int gvib200_table_file_write(const char* path, int n_keys, const int32_t* dims, const int32_t* degs) {
    if (!path || n_keys < 0 || (n_keys > 0 && (!dims || !degs))) return fail(GVIB200_EINVAL, "table_file_write: bad arguments");
    try {
        std::vector<SpghTableEntry> entries((size_t)n_keys);
        for (int i = 0; i < n_keys; ++i) {
            entries[i].dim = dims[i];
            entries[i].deg = degs[i];
            generate_spgh_table(dims[i], degs[i], entries[i].nodes_rowmajor, entries[i].weights);
        }
        write_spgh_table_file(path, entries);
    } catch (const std::invalid_argument& e) {
        return fail(GVIB200_ENOTABLE, e.what());
    } catch (const std::exception& e) {
        return fail(GVIB200_EINVAL, e.what());
    }
    return 0;
}

extern "C" int gvib200_table_file_load(gvib200_ctx* ctx, const char* path, int* n_loaded) {
    if (!ctx || !path) return fail(GVIB200_EINVAL, "table_file_load: bad arguments");
    std::vector<SpghTableEntry> entries;
    try {
        read_spgh_table_file(path, entries);
    } catch (const std::exception& e) {
        return fail(GVIB200_EINVAL, e.what());
    }
    for (auto& e : entries) {
        if (e.weights.empty()) continue;
        TRY(gvib200_table_set(ctx, e.dim, e.deg, (int)e.weights.size(), e.nodes_rowmajor.data(), e.weights.data()));
    }
    if (n_loaded) *n_loaded = (int)entries.size();
    return 0;
}

extern "C" int gvib200_table_file_query(const char* path, int capacity, int32_t* dims, int32_t* degs, int32_t* sizes) {
    if (!path) return fail(GVIB200_EINVAL, "table_file_query: null path");
    std::vector<SpghTableEntry> entries;
    try {
        read_spgh_table_file(path, entries);
    } catch (const std::exception& e) {
        return fail(GVIB200_EINVAL, e.what());
    }
    for (int i = 0; i < (int)entries.size() && i < capacity; ++i) {
        if (dims) dims[i] = entries[i].dim;
        if (degs) degs[i] = entries[i].deg;
        if (sizes) sizes[i] = (int32_t)entries[i].weights.size();
    }
    return (int)entries.size();
}

extern "C" int gvib200_table_get(gvib200_ctx* ctx, int dim, int deg, double* nodes_rowmajor, double* weights, int capacity) {
    if (!ctx) return fail(GVIB200_EINVAL, "table_get: null context");
    auto it = ctx->tables.find(std::make_pair(dim, deg));
    std::vector<double> n, w;
    if (it != ctx->tables.end()) {
        n = it->second->nodes;
        w = it->second->w;
    } else {
        try {
            generate_spgh_table(dim, deg, n, w);
        } catch (const std::exception& e) {
            return fail(GVIB200_ENOTABLE, e.what());
        }
    }
    if (!nodes_rowmajor && !weights) return (int)w.size();  // size query
    if ((int)w.size() > capacity) return fail(GVIB200_EINVAL, "table_get: capacity too small");
    if (nodes_rowmajor) std::memcpy(nodes_rowmajor, n.data(), n.size() * sizeof(double));
    if (weights) std::memcpy(weights, w.data(), w.size() * sizeof(double));
    return (int)w.size();
}

// ------------------------------------------------------------------------------------------------
// C-ABI: device-side LTV prior set-up (gp/LTV_prior.h:123-197), ltv_setup.cuh
// ------------------------------------------------------------------------------------------------
extern "C" int gvib200_ltv_transition(gvib200_ctx* ctx, int n_links, int dim_state, int n_inputs, const double* A,
                                      const double* B, double delta_t, double* Phi, double* Q, double* Qinv) {
    if (!ctx || n_links < 1 || n_inputs < 1 || !A || !B || !Phi || !Q || !(delta_t > 0))
        return fail(GVIB200_EINVAL, "ltv_transition: bad arguments");
    if (!(dim_state == 2 || dim_state == 4 || dim_state == 6))
        return fail(GVIB200_EINVAL, "ltv_transition: state dimension must be 2, 4 or 6");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t dd = (size_t)dim_state * dim_state, nA = (size_t)n_links * 4 * dd,
                 nB = (size_t)n_links * 4 * dim_state * n_inputs, nO = (size_t)n_links * dd;
    double *dA = nullptr, *dB = nullptr, *dO = nullptr;
    auto cleanup = [&]() {
        if (dA) cudaFree(dA);
        if (dB) cudaFree(dB);
        if (dO) cudaFree(dO);
    };
    auto body = [&]() -> int {
        TRY(dev_alloc(&dA, nA));
        TRY(dev_alloc(&dB, nB));
        TRY(dev_alloc(&dO, 3 * nO));
        CUDA_TRY(cudaMemcpy(dA, A, nA * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(dB, B, nB * sizeof(double), cudaMemcpyHostToDevice));
        const int block = 64, grid = cdiv(n_links, block);
        double* dQi = Qinv ? dO + 2 * nO : nullptr;
        switch (dim_state) {
            case 2: k_ltv_transition<2><<<grid, block>>>(n_links, n_inputs, delta_t, dA, dB, dO, dO + nO, dQi); break;
            case 4: k_ltv_transition<4><<<grid, block>>>(n_links, n_inputs, delta_t, dA, dB, dO, dO + nO, dQi); break;
            default: k_ltv_transition<6><<<grid, block>>>(n_links, n_inputs, delta_t, dA, dB, dO, dO + nO, dQi); break;
        }
        ctx->launches++;
        TRY(check_launch("k_ltv_transition"));
        CUDA_TRY(cudaMemcpy(Phi, dO, nO * sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(Q, dO + nO, nO * sizeof(double), cudaMemcpyDeviceToHost));
        if (Qinv) CUDA_TRY(cudaMemcpy(Qinv, dO + 2 * nO, nO * sizeof(double), cudaMemcpyDeviceToHost));
        return 0;
    };
    const int rc = body();
    cleanup();
    return rc;
}

// ------------------------------------------------------------------------------------------------
// C-ABI: problem definition
// ------------------------------------------------------------------------------------------------
extern "C" int gvib200_problem_create(gvib200_ctx* ctx, int num_states, int dim_state, gvib200_problem** out) {
    if (!ctx || !out || num_states < 1) return fail(GVIB200_EINVAL, "problem_create: bad arguments");
    if (!(dim_state == 1 || dim_state == 2 || dim_state == 3 || dim_state == 4 || dim_state == 6))
        return fail(GVIB200_EINVAL, "problem_create: state dimension must be 1, 2, 3, 4 or 6");
    CUDA_TRY(cudaSetDevice(ctx->device));
    std::unique_ptr<gvib200_problem> p(new gvib200_problem);
    p->ctx = ctx;
    p->S = num_states;
    p->d = dim_state;
    CUDA_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    {
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&p->stream2, cudaStreamNonBlocking, prio_hi));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_pro, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_k1, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_host, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_pending, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_mu, cudaEventDisableTiming));
    p->ls = p->stream;
    *out = p.release();
    return 0;
}

static void free_batch(gvib200_problem* p);
static void free_problem(gvib200_problem* p) {
    // speculative work (the assembly of the last trial's sweep) may still be in flight on either stream
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->stream2) cudaStreamSynchronize(p->stream2);
    free_batch(p);
    auto F = [](void* q) {
        if (q) cudaFree(q);
    };
    for (auto& g : p->gh) {
        F(g.d_start); F(g.d_T); F(g.d_Thigh); F(g.d_params); F(g.d_SR[0]); F(g.d_SR[1]); F(g.d_raw); F(g.d_active); F(g.d_nactive);
    }
    for (auto& g : p->lin) {
        F(g.d_start); F(g.d_Lambda); F(g.d_psi); F(g.d_Kinv); F(g.d_A); F(g.d_C); F(g.d_T);
    }
    F(p->d_sdf_rec);
    F(p->d_sdf_data);
    F(p->d_sdf3);
    F(p->d_sdf_coarse);
    F(p->d_evaluated);
    for (int i = 0; i < 2; ++i) {
        F(p->mu[i]); F(p->LD[i]); F(p->LO[i]); F(p->CD[i]); F(p->CO[i]); F(p->fcost[i]); F(p->fVdmu[i]); F(p->fVdd[i]);
    }
    F(p->scal); F(p->partial); F(p->d_flag); F(p->d_counter); F(p->Vdmu); F(p->VD); if (!p->vo_alias) { F(p->VO); F(p->VO2); } F(p->rhs); F(p->Vdmu2); F(p->VD2); F(p->rhs2); F(p->dmu); F(p->KlinD); F(p->KlinO);
    F(p->ell_v); F(p->ell_d); F(p->ell_dl); F(p->ell_o); F(p->ell_ol); F(p->vptr); F(p->voff); F(p->dptr); F(p->doff); F(p->dld); F(p->optr); F(p->ooff); F(p->old); F(p->ws[0]); F(p->ws[1]);
    for (int i = 0; i < 2; ++i) {
        F(p->ws_mid[i]); F(p->ws_top[i]); F(p->dist_buf[i]);
    }
    F(p->red_buf);
    F(p->snap);
    for (auto& r : p->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    if (p->t0) cudaEventDestroy(p->t0);
    if (p->t1) cudaEventDestroy(p->t1);
    if (p->h_scal) cudaFreeHost(p->h_scal);
    if (p->zc) cudaFreeHost(p->zc);
    if (p->h_flag) cudaFreeHost(p->h_flag);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    if (p->ev_pro) cudaEventDestroy(p->ev_pro);
    if (p->ev_k1) cudaEventDestroy(p->ev_k1);
    if (p->ev_host) cudaEventDestroy(p->ev_host);
    if (p->ev_pending) cudaEventDestroy(p->ev_pending);
    if (p->ev_mu) cudaEventDestroy(p->ev_mu);
    if (p->stream2) cudaStreamDestroy(p->stream2);
    if (p->stream) cudaStreamDestroy(p->stream);
}

extern "C" int gvib200_problem_destroy(gvib200_problem* prob) {
    if (!prob) return 0;
    cudaSetDevice(prob->ctx->device);
    free_problem(prob);
    delete prob;
    return 0;
}

extern "C" int gvib200_set_planar_sdf(gvib200_problem* p, int rows, int cols, double ox, double oy, double cell,
                                      const double* data) {
    if (!p || rows < 2 || cols < 2 || !(cell > 0) || !data) return fail(GVIB200_EINVAL, "set_planar_sdf: bad arguments");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    if (p->d_sdf_rec) cudaFree(p->d_sdf_rec);
    if (p->d_sdf_data) cudaFree(p->d_sdf_data);
    p->d_sdf_rec = nullptr;
    p->d_sdf_data = nullptr;
    const size_t n = (size_t)rows * cols;
    CUDA_TRY(cudaMalloc((void**)&p->d_sdf_data, n * sizeof(double)));
    CUDA_TRY(cudaMalloc((void**)&p->d_sdf_rec, n * sizeof(double4)));
    if (p->d_sdf_coarse) cudaFree(p->d_sdf_coarse);
    p->d_sdf_coarse = nullptr;
    CUDA_TRY(cudaMalloc((void**)&p->d_sdf_coarse, (size_t)((rows + 3) / 4) * ((cols + 3) / 4) * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(p->d_sdf_data, data, n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->sdf_rec_valid = false;
    p->sdf_rows = rows;
    p->sdf_cols = cols;
    p->sdf_ox = ox;
    p->sdf_oy = oy;
    p->sdf_cell = cell;
    // everything computed with the previous field is stale
    p->sweep_valid = false;
    p->asm_valid = false;
    p->grads_valid = false;
    p->zc_ok[0] = p->zc_ok[1] = false;
    return 0;
}

extern "C" int gvib200_set_sdf3d(gvib200_problem* p, int rows, int cols, int nz, double ox, double oy, double oz, double cell,
                                 const double* data) {
    if (!p || rows < 2 || cols < 2 || nz < 2 || !(cell > 0) || !data) return fail(GVIB200_EINVAL, "set_sdf3d: bad arguments");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    if (p->d_sdf3) cudaFree(p->d_sdf3);
    p->d_sdf3 = nullptr;
    const size_t n = (size_t)rows * cols * nz;
    CUDA_TRY(cudaMalloc((void**)&p->d_sdf3, n * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(p->d_sdf3, data, n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->sdf3_rows = rows;
    p->sdf3_cols = cols;
    p->sdf3_nz = nz;
    p->sdf3_o[0] = ox;
    p->sdf3_o[1] = oy;
    p->sdf3_o[2] = oz;
    p->sdf3_cell = cell;
    p->sweep_valid = false;
    p->asm_valid = false;
    p->grads_valid = false;
    p->zc_ok[0] = p->zc_ok[1] = false;
    return 0;
}

static size_t cost_param_record(int kind, int dim) {
    switch (kind) {
        case GVIB200_COST_LINEAR_GP: return (size_t)2 * (dim / 2) * (dim / 2) * sizeof(double);
        case GVIB200_COST_FIXED_GP: return (size_t)(dim * dim + dim) * sizeof(double);
        default: return 0;
    }
}

extern "C" int gvib200_add_gh_factors(gvib200_problem* p, int kind, int dim, int deg, int n, const int32_t* start,
                                      const double* T, const double* Thigh, const void* params, size_t bytes,
                                      int* first_id) {
    if (!p || n < 1 || !start) return fail(GVIB200_EINVAL, "add_gh_factors: bad arguments");
    if (p->finalized) return fail(GVIB200_ESTATE, "add_gh_factors: problem already finalized");
    if (dim % p->d != 0 || (dim / p->d != 1 && dim / p->d != 2))
        return fail(GVIB200_EINVAL, "add_gh_factors: factor dim must span 1 or 2 states");
    const int nst = dim / p->d;
    for (int i = 0; i < n; ++i)
        if (start[i] < 0 || start[i] + nst > p->S) return fail(GVIB200_EINVAL, "add_gh_factors: start_index out of range");
    size_t want = 0;
    switch (kind) {
        case GVIB200_COST_STEREO_1D: want = sizeof(gvib200_stereo1d_params); break;
        case GVIB200_COST_PLANAR_HINGE:
        case GVIB200_COST_HINGE_3D:
        case GVIB200_COST_QUAD_HINGE: want = sizeof(gvib200_hinge_params); break;
        case GVIB200_COST_QUADRATIC: want = sizeof(double); break;
        case GVIB200_COST_ARM_3D: {
            want = sizeof(gvib200_arm_params);
            if (params && bytes == want) {
                const auto* ap = reinterpret_cast<const gvib200_arm_params*>(params);
                if (ap->n_dof < 1 || ap->n_dof > GVIB200_ARM_MAX_DOF || 2 * ap->n_dof != dim || dim != p->d || ap->n_spheres < 1 ||
                    ap->n_spheres > GVIB200_ARM_MAX_SPHERES)
                    return fail(GVIB200_EINVAL, "add_gh_factors: arm cost needs dim = state dim = 2 n_dof, n_dof <= 3, 1..12 spheres");
                for (int i = 0; i < ap->n_spheres; ++i)
                    if (ap->frames[i] < 0 || ap->frames[i] >= ap->n_dof)
                        return fail(GVIB200_EINVAL, "add_gh_factors: arm sphere frame out of range");
            }
            break;
        }
        case GVIB200_COST_LINEAR_GP:
        case GVIB200_COST_FIXED_GP: want = cost_param_record(kind, dim) * n; break;
        default: return fail(GVIB200_EINVAL, "add_gh_factors: unknown cost kind");
    }
    if (bytes != want || !params) return fail(GVIB200_EINVAL, "add_gh_factors: cost_params size mismatch");
    GhGroup g;
    g.kind = kind;
    g.dim = dim;
    g.deg = deg;
    g.n = n;
    g.first_id = p->n_factors;
    g.start.assign(start, start + n);
    g.T.assign(n, 1.0);
    g.Thigh.assign(n, 10.0);
    if (T) g.T.assign(T, T + n);
    if (Thigh) g.Thigh.assign(Thigh, Thigh + n);
    g.params.assign((const char*)params, (const char*)params + bytes);
    TRY(get_table(p->ctx, dim, deg, &g.table));
    for (int i = 0; i < n; ++i) p->factors.push_back(FactorRef{false, (int)p->gh.size(), i});
    p->n_factors += n;
    if (first_id) *first_id = g.first_id;
    p->gh.push_back(std::move(g));
    return 0;
}

extern "C" int gvib200_add_linear_factors(gvib200_problem* p, int dim, int m, int kdim, int n, const int32_t* start,
                                          const double* Lambda, const double* Psi, const double* mu_t,
                                          const double* Kinv, const double* C, const double* T, const double* Thigh,
                                          int* first_id) {
    if (!p || n < 1 || !start || !Lambda || !Psi || !mu_t || !Kinv || !C)
        return fail(GVIB200_EINVAL, "add_linear_factors: bad arguments");
    if (p->finalized) return fail(GVIB200_ESTATE, "add_linear_factors: problem already finalized");
    if (dim % p->d != 0 || (dim / p->d != 1 && dim / p->d != 2) || dim > LIN_MAX_DIM || m > LIN_MAX_DIM || m < 1 ||
        kdim < 1)
        return fail(GVIB200_EINVAL, "add_linear_factors: unsupported dimensions");
    const int nst = dim / p->d;
    for (int i = 0; i < n; ++i)
        if (start[i] < 0 || start[i] + nst > p->S) return fail(GVIB200_EINVAL, "add_linear_factors: start_index out of range");
    LinGroup g;
    g.dim = dim;
    g.m = m;
    g.kdim = kdim;
    g.n = n;
    g.first_id = p->n_factors;
    g.start.assign(start, start + n);
    g.Lambda.assign(Lambda, Lambda + (size_t)n * m * dim);
    g.Kinv.assign(Kinv, Kinv + (size_t)n * m * m);
    g.C.assign(C, C + n);
    g.T.assign(n, 1.0);
    g.Thigh.assign(n, 10.0);
    if (T) g.T.assign(T, T + n);
    if (Thigh) g.Thigh.assign(Thigh, Thigh + n);
    g.psi.assign((size_t)n * m, 0.0);
    g.A.assign((size_t)n * dim * dim, 0.0);
    std::vector<double> KL((size_t)m * dim);
    for (int f = 0; f < n; ++f) {
        const double* L = Lambda + (size_t)f * m * dim;
        const double* P = Psi + (size_t)f * m * kdim;
        const double* mt = mu_t + (size_t)f * kdim;
        const double* K = Kinv + (size_t)f * m * m;
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            for (int k = 0; k < kdim; ++k) s += P[i + (size_t)k * m] * mt[k];
            g.psi[(size_t)f * m + i] = s;
        }
        // A = Lambda^T Kinv Lambda
        for (int j = 0; j < dim; ++j)
            for (int i = 0; i < m; ++i) {
                double s = 0.0;
                for (int k = 0; k < m; ++k) s += K[i + (size_t)k * m] * L[k + (size_t)j * m];
                KL[i + (size_t)j * m] = s;
            }
        double* A = g.A.data() + (size_t)f * dim * dim;
        for (int j = 0; j < dim; ++j)
            for (int i = 0; i < dim; ++i) {
                double s = 0.0;
                for (int k = 0; k < m; ++k) s += L[k + (size_t)i * m] * KL[k + (size_t)j * m];
                A[i + (size_t)j * dim] = s;
            }
    }
    for (int i = 0; i < n; ++i) p->factors.push_back(FactorRef{true, (int)p->lin.size(), i});
    p->n_factors += n;
    if (first_id) *first_id = g.first_id;
    p->lin.push_back(std::move(g));
    return 0;
}

// constant part of Vddmu contributed by the linear factors: sum_k scatter(2 C_k A_k / T_k)
static int upload_klin(gvib200_problem* p) {
    const int d = p->d, S = p->S;
    const size_t dd = (size_t)d * d;
    std::vector<double> KD((size_t)S * dd, 0.0), KO((size_t)std::max(S - 1, 1) * dd, 0.0);
    for (auto& g : p->lin) {
        if (p->prox) break;  // Prox-GVI: the linear factors' Vddmu depends on the state (BW_JKO), nothing is constant
        for (int f = 0; f < g.n; ++f) {
            const double sc = 2.0 * g.C[f] / g.T[f];
            const double* A = g.A.data() + (size_t)f * g.dim * g.dim;
            const int s = g.start[f];
            const int nst = g.dim / d;
            for (int bj = 0; bj < nst; ++bj)
                for (int bi = 0; bi <= bj; ++bi) {
                    double* dst = (bi == bj) ? KD.data() + (size_t)(s + bi) * dd : KO.data() + (size_t)s * dd;
                    for (int j = 0; j < d; ++j)
                        for (int i = 0; i < d; ++i)
                            dst[i + (size_t)j * d] += sc * A[(bi * d + i) + (size_t)(bj * d + j) * g.dim];
                }
        }
    }
    CUDA_TRY(cudaMemcpyAsync(p->KlinD, KD.data(), (size_t)S * dd * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    if (S > 1)
        CUDA_TRY(cudaMemcpyAsync(p->KlinO, KO.data(), (size_t)(S - 1) * dd * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" int gvib200_problem_finalize(gvib200_problem* p) {
    if (!p) return fail(GVIB200_EINVAL, "finalize: null problem");
    if (p->finalized) return 0;
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    const int S = p->S, d = p->d;
    const size_t dd = (size_t)d * d;
    // offsets of per-factor outputs
    size_t nV = 0, nM = 0;
    for (auto& g : p->gh) {
        g.voff = nV;
        g.moff = nM;
        nV += (size_t)g.n * g.dim;
        nM += (size_t)g.n * g.dim * g.dim;
    }
    for (auto& g : p->lin) {
        g.voff = nV;
        nV += (size_t)g.n * g.dim;
        if (p->prox) {
            g.moff = nM;
            nM += (size_t)g.n * g.dim * g.dim;
        }
    }
    if (p->prox)  // ProxGVIFactorizedBaseGH::fact_cost_value does not divide by the temperature
        for (auto& g : p->gh) {
            g.T.assign(g.T.size(), 1.0);
            g.Thigh.assign(g.Thigh.size(), 1.0);
        }
    p->nV = nV;
    p->nM = nM;
    // adjacency (id order within each state => fixed summation order)
    std::vector<std::vector<int>> vl(S), dl(S), dll(S), ol(S), oll(S);
    for (auto& fr : p->factors) {
        if (!fr.linear) {
            auto& g = p->gh[fr.group];
            const int s = g.start[fr.index], nst = g.dim / d;
            const size_t vo = g.voff + (size_t)fr.index * g.dim, mo = g.moff + (size_t)fr.index * g.dim * g.dim;
            for (int b = 0; b < nst; ++b) {
                vl[s + b].push_back((int)(vo + (size_t)b * d));
                dl[s + b].push_back((int)(mo + (size_t)b * d + (size_t)b * d * g.dim));
                dll[s + b].push_back(g.dim);
            }
            if (nst == 2) {
                ol[s].push_back((int)(mo + (size_t)d * g.dim));
                oll[s].push_back(g.dim);
            }
        } else {
            auto& g = p->lin[fr.group];
            const int s = g.start[fr.index], nst = g.dim / d;
            const size_t vo = g.voff + (size_t)fr.index * g.dim;
            for (int b = 0; b < nst; ++b) vl[s + b].push_back((int)(vo + (size_t)b * d));
            if (p->prox) {
                const size_t mo = g.moff + (size_t)fr.index * g.dim * g.dim;
                for (int b = 0; b < nst; ++b) {
                    dl[s + b].push_back((int)(mo + (size_t)b * d + (size_t)b * d * g.dim));
                    dll[s + b].push_back(g.dim);
                }
                if (nst == 2) {
                    ol[s].push_back((int)(mo + (size_t)d * g.dim));
                    oll[s].push_back(g.dim);
                }
            }
        }
    }
    if (nV > 0x7fffffffULL || nM > 0x7fffffffULL) return fail(GVIB200_EINVAL, "finalize: problem too large for int32 offsets");
    auto flatten = [&](const std::vector<std::vector<int>>& ll, std::vector<int>& ptr, std::vector<int>& val) {
        ptr.assign(S + 1, 0);
        val.clear();
        for (int s = 0; s < S; ++s) {
            ptr[s] = (int)val.size();
            val.insert(val.end(), ll[s].begin(), ll[s].end());
        }
        ptr[S] = (int)val.size();
    };
    std::vector<int> ptr, val, val2;
    flatten(vl, ptr, val);
    TRY(dev_upload(&p->vptr, ptr, p->stream));
    TRY(dev_upload(&p->voff, val, p->stream));
    flatten(dl, ptr, val);
    flatten(dll, ptr, val2);
    TRY(dev_upload(&p->dptr, ptr, p->stream));
    TRY(dev_upload(&p->doff, val, p->stream));
    TRY(dev_upload(&p->dld, val2, p->stream));
    flatten(ol, ptr, val);
    flatten(oll, ptr, val2);
    TRY(dev_upload(&p->optr, ptr, p->stream));
    TRY(dev_upload(&p->ooff, val, p->stream));
    TRY(dev_upload(&p->old, val2, p->stream));
    {   // ELL form
        auto width = [&](const std::vector<std::vector<int>>& ll) {
            size_t w = 0;
            for (auto& l : ll) w = std::max(w, l.size());
            return (int)w;
        };
        const int nv = width(vl), nd = width(dl), no = width(ol);
        if (nv <= 8 && nd <= 4 && no <= 4) {
            auto ell = [&](const std::vector<std::vector<int>>& ll, int w, std::vector<int>& out) {
                out.assign((size_t)std::max(w, 1) * S, -1);
                for (int st = 0; st < S; ++st)
                    for (size_t k = 0; k < ll[st].size(); ++k) out[k * S + st] = ll[st][k];
            };
            std::vector<int> e;
            ell(vl, std::max(nv, 4), e);  // at least the widths k_assemble_ell_fast is compiled for (-1 padded)
            TRY(dev_upload(&p->ell_v, e, p->stream));
            ell(dl, std::max(nd, 1), e);
            TRY(dev_upload(&p->ell_d, e, p->stream));
            ell(dll, std::max(nd, 1), e);
            TRY(dev_upload(&p->ell_dl, e, p->stream));
            ell(ol, no, e);
            TRY(dev_upload(&p->ell_o, e, p->stream));
            ell(oll, no, e);
            TRY(dev_upload(&p->ell_ol, e, p->stream));
            p->ell_nv = nv;
            p->ell_nd = nd;
            p->ell_no = no;
        }
    }
    // factor groups
    for (auto& g : p->gh) {
        TRY(dev_upload(&g.d_start, g.start, p->stream));
        TRY(dev_upload(&g.d_T, g.T, p->stream));
        TRY(dev_upload(&g.d_Thigh, g.Thigh, p->stream));
        if (g.kind == GVIB200_COST_LINEAR_GP || g.kind == GVIB200_COST_FIXED_GP) {
            CUDA_TRY(cudaMalloc(&g.d_params, g.params.size()));
            CUDA_TRY(cudaMemcpyAsync(g.d_params, g.params.data(), g.params.size(), cudaMemcpyHostToDevice, p->stream));
        }
        for (int i = 0; i < 2; ++i) {
            TRY(dev_alloc(&g.d_SR[i], (size_t)g.n * 2 * g.dim * g.dim));
            // finite everywhere from the start: the entries of culled factors are never written by the fused culling + prologue
            // pass and meet exactly-zero raw moments in k_raw_to_x
            CUDA_TRY(cudaMemsetAsync(g.d_SR[i], 0, (size_t)g.n * 2 * g.dim * g.dim * sizeof(double), p->stream));
        }
    }
    for (auto& g : p->lin) {
        TRY(dev_upload(&g.d_start, g.start, p->stream));
        // element-major device copies (see LinearArgs); A as its packed upper triangle
        auto soa = [&](const std::vector<double>& aos, int ne) {
            std::vector<double> out(aos.size());
            for (int f = 0; f < g.n; ++f)
                for (int e = 0; e < ne; ++e) out[(size_t)e * g.n + f] = aos[(size_t)f * ne + e];
            return out;
        };
        std::vector<double> Apk((size_t)g.n * (g.dim * (g.dim + 1) / 2));
        for (int f = 0; f < g.n; ++f) {
            int e = 0;
            for (int i = 0; i < g.dim; ++i)
                for (int j = i; j < g.dim; ++j) Apk[(size_t)(e++) * g.n + f] = g.A[(size_t)f * g.dim * g.dim + i + (size_t)j * g.dim];
        }
        TRY(dev_upload(&g.d_Lambda, soa(g.Lambda, g.m * g.dim), p->stream));
        TRY(dev_upload(&g.d_psi, soa(g.psi, g.m), p->stream));
        TRY(dev_upload(&g.d_Kinv, soa(g.Kinv, g.m * g.m), p->stream));
        TRY(dev_upload(&g.d_A, Apk, p->stream));
        TRY(dev_upload(&g.d_C, g.C, p->stream));
        TRY(dev_upload(&g.d_T, g.T, p->stream));
    }
    // state
    for (int i = 0; i < 2; ++i) {
        TRY(dev_alloc(&p->mu[i], (size_t)S * d));
        TRY(dev_alloc(&p->LD[i], (size_t)S * dd));
        TRY(dev_alloc(&p->LO[i], (size_t)S * dd));
        TRY(dev_alloc(&p->CD[i], (size_t)S * dd));
        TRY(dev_alloc(&p->CO[i], (size_t)S * dd));
        TRY(dev_alloc(&p->fcost[i], (size_t)p->n_factors));
        TRY(dev_alloc(&p->fVdmu[i], nV));
        TRY(dev_alloc(&p->fVdd[i], nM));
        CUDA_TRY(cudaMemsetAsync(p->LO[i], 0, (size_t)S * dd * sizeof(double), p->stream));
        CUDA_TRY(cudaMemsetAsync(p->CO[i], 0, (size_t)S * dd * sizeof(double), p->stream));
    }
    TRY(dev_alloc(&p->scal, 8));
    TRY(dev_alloc(&p->partial, 1024));
    CUDA_TRY(cudaMemsetAsync(p->scal, 0, 8 * sizeof(double), p->stream));
    CUDA_TRY(cudaMallocHost((void**)&p->h_scal, 8 * sizeof(double)));
    CUDA_TRY(cudaHostAlloc((void**)&p->zc, 8 * sizeof(double), cudaHostAllocMapped));
    CUDA_TRY(cudaHostGetDevicePointer((void**)&p->zc_dev, p->zc, 0));
    TRY(dev_alloc(&p->d_flag, 2));
    CUDA_TRY(cudaMemsetAsync(p->d_flag, 0, 2 * sizeof(int), p->stream));
    TRY(dev_alloc(&p->d_evaluated, 1));
    CUDA_TRY(cudaMemsetAsync(p->d_evaluated, 0, sizeof(unsigned long long), p->stream));
    TRY(dev_alloc(&p->d_counter, 1));
    CUDA_TRY(cudaMemsetAsync(p->d_counter, 0, sizeof(unsigned), p->stream));
    CUDA_TRY(cudaMallocHost((void**)&p->h_flag, 2 * sizeof(int)));
    TRY(dev_alloc(&p->Vdmu, (size_t)S * d));
    TRY(dev_alloc(&p->rhs, (size_t)S * d));
    TRY(dev_alloc(&p->Vdmu2, (size_t)S * d));
    TRY(dev_alloc(&p->rhs2, (size_t)S * d));
    TRY(dev_alloc(&p->dmu, (size_t)S * d));
    TRY(dev_alloc(&p->VD, (size_t)S * dd));
    TRY(dev_alloc(&p->VD2, (size_t)S * dd));
    TRY(dev_alloc(&p->KlinD, (size_t)S * dd));
    TRY(dev_alloc(&p->KlinO, (size_t)S * dd));
    // No factor contributes an off-diagonal block of Vddmu beyond the constant linear part (the case of every chain whose
    // GH factors sit on single states) and the vectorised assembly kernel applies: the off-diagonal blocks of Vddmu ARE KlinO
    // -- the assembly neither reads nor writes them (25.6 MB less traffic per assembly at the headline shape)
    p->vo_alias = !p->prox && (d % 2 == 0) && p->ell_nv >= 0 && p->ell_nv <= 4 && p->ell_nd <= 1 && p->ell_no == 0 &&
                  getenv("GVIB200_NO_FAST_ASSEMBLE") == nullptr && getenv("GVIB200_NO_VO_ALIAS") == nullptr;
    if (p->vo_alias) {
        p->VO = p->VO2 = p->KlinO;
    } else {
        TRY(dev_alloc(&p->VO, (size_t)S * dd));
        TRY(dev_alloc(&p->VO2, (size_t)S * dd));
    }
    TRY(upload_klin(p));
    // chain plan
    {
        // dynamic shared memory left for the chain kernels next to their static arrays
        const size_t smem = p->ctx->smem_optin - 5120;
        const int P = p->ctx->world;
        bool ok = false;
        int tiles_per_sm = 1;
        if (const char* e = getenv("GVIB200_TILES_PER_SM")) tiles_per_sm = std::max(1, atoi(e));
        if (const char* e = getenv("GVIB200_TILE_THREADS")) p->tile_threads = std::max(32, std::min(CR_THREADS, atoi(e)));
        int long_tiles = 2;  // measured on cfg5 (1024 x 1002 states): 1 -> 3.79, 2 -> 3.64, 3 -> 3.89 ms per iteration
        if (const char* e = getenv("GVIB200_LONG_TILES_PER_SM")) long_tiles = std::max(1, std::min(4, atoi(e)));
#define PLAN_CASE(D_)                                                                                          \
    case D_: {                                                                                                 \
        int force_T = 0;                                                                                       \
        if (P > 1) { /* multi-GPU: always tiled, the separator chain goes through the mid level */            \
            /* two SMs stay free of tile CTAs, as on one GPU: the single-CTA stage of one pass must not queue behind  \
               the tile CTAs of the other pass */                                                              \
            int K = std::min(std::max(1, p->ctx->sm_count - 2), cr_max_top_nodes<D_>(smem) - 1);               \
            K = std::max(1, std::min(K, S - 1));                                                               \
            force_T = std::max(2, std::min((S - 1 + K - 1) / K, cr_max_tile_links<D_>(smem)));                 \
        }                                                                                                      \
        /* two SMs are left free of tile CTAs: the single-CTA top kernels of the two concurrent passes run there */ \
        ok = cr_make_plan<D_>(p->plan, S, tiles_per_sm * std::max(1, p->ctx->sm_count - 2), smem, force_T);    \
        if (!ok && P == 1) { /* long chain: three levels (tiles -> tiles over the separator chain -> top).  Many    \
               waves of tiles: throughput counts, so `long_tiles` tile CTAs of 512 / long_tiles threads share an SM  \
               (each with 1 / long_tiles of the shared memory) and overlap their latency-bound upper levels */  \
            const size_t smem_tile = long_tiles > 1 ? (smem + 5120) / long_tiles - 5120 - 1024 : smem;         \
            ok = cr_make_plan<D_>(p->plan, S, p->ctx->sm_count, smem_tile, 0, true) &&                         \
                 cr_make_plan<D_>(p->plan_mid, p->plan.K + 1, p->ctx->sm_count, smem);                         \
            p->three_level = ok;                                                                               \
            if (ok && long_tiles > 1) p->tile_threads = CR_THREADS / long_tiles;                               \
        }                                                                                                      \
        if (ok && P > 1)                                                                                       \
            ok = cr_make_plan<D_>(p->plan_mid, p->plan.K + 1, 1, smem, std::max(p->plan.K, 2)) &&              \
                 cr_make_plan<D_>(p->plan_top, P + 1, 1, smem, -1);                                            \
    } break;
        switch (d) {
            PLAN_CASE(1) PLAN_CASE(2) PLAN_CASE(3) PLAN_CASE(4) PLAN_CASE(6)
        }
#undef PLAN_CASE
        if (P > 1 && S < 2) return fail(GVIB200_EINVAL, "finalize: a multi-GPU segment needs at least two states");
        if (!ok) return fail(GVIB200_EINVAL, "finalize: the chain is too long for the two-level block-tridiagonal plan");
    }
    for (int i = 0; i < 2; ++i) {
        TRY(dev_alloc(&p->ws[i], p->plan.ws_doubles + 16));
        CUDA_TRY(cudaMemsetAsync(p->ws[i], 0, (p->plan.ws_doubles + 16) * sizeof(double), p->stream));
        if (p->three_level) {
            const DistLayout L = dist_layout(d, p->plan.K, 1);
            TRY(dev_alloc(&p->ws_mid[i], p->plan_mid.ws_doubles + 16));
            TRY(dev_alloc(&p->dist_buf[i], L.total + 16));
            CUDA_TRY(cudaMemsetAsync(p->ws_mid[i], 0, (p->plan_mid.ws_doubles + 16) * sizeof(double), p->stream));
            CUDA_TRY(cudaMemsetAsync(p->dist_buf[i], 0, (L.total + 16) * sizeof(double), p->stream));
        }
        if (p->ctx->world > 1) {
            const DistLayout L = dist_layout(d, p->plan.K, p->ctx->world);
            TRY(dev_alloc(&p->ws_mid[i], p->plan_mid.ws_doubles + 16));
            TRY(dev_alloc(&p->ws_top[i], p->plan_top.ws_doubles + 16));
            TRY(dev_alloc(&p->dist_buf[i], L.total + 16));
            CUDA_TRY(cudaMemsetAsync(p->ws_mid[i], 0, (p->plan_mid.ws_doubles + 16) * sizeof(double), p->stream));
            CUDA_TRY(cudaMemsetAsync(p->ws_top[i], 0, (p->plan_top.ws_doubles + 16) * sizeof(double), p->stream));
            CUDA_TRY(cudaMemsetAsync(p->dist_buf[i], 0, (L.total + 16) * sizeof(double), p->stream));
        }
    }
    if (p->ctx->world > 1) TRY(dev_alloc(&p->red_buf, (size_t)4 * (p->ctx->world + 1)));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->finalized = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C-ABI: state
// ------------------------------------------------------------------------------------------------
static int recompute_from_precision(gvib200_problem* p, int which, bool async = false) {
    TRY(clear_flag(p));
    TRY(do_selinv(p, p->LD[which], p->LO[which], p->CD[which], p->CO[which], p->scal + which));
    TRY(run_prologue_only(p, which));
    if (async) {  // the host does not wait: the flag is read by the next call that synchronises (resolve_pending)
        if (p->ctx->world > 1 && !p->flags_synced) TRY(dist_reduce(p, p->scal + 7));  // scal[7]: scratch
        LAUNCH(p, KC_OTHER, k_flags_to_host, 1, 1, 0, p->d_flag, p->zc_dev + 4);
        CUDA_TRY(cudaEventRecord(p->ev_pending, p->stream));
        p->pending_check = true;
        return check_launch("set_state_async");
    }
    int flag = 0;
    TRY(read_flag(p, &flag));
    if (flag) return fail(GVIB200_ENOTSPD, "precision matrix is not positive definite");
    return 0;
}

// gvib200_set_state_async left its validity check open: wait for the upload + selected inverse and read the flag
static int resolve_pending(gvib200_problem* p) {
    if (!p->pending_check) return 0;
    p->pending_check = false;
    CUDA_TRY(cudaEventSynchronize(p->ev_pending));
    const int flag = p->zc[4] != 0.0;
    if (flag) {
        p->has_state = false;
        return fail(GVIB200_ENOTSPD, "precision matrix is not positive definite (gvib200_set_state_async)");
    }
    return 0;
}

static int set_state_impl(gvib200_problem* p, const double* mu, const double* pd, const double* po, bool async) {
    if (!p || !p->finalized) return fail(GVIB200_ESTATE, "set_state: problem not finalized");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    const int S = p->S, d = p->d, c = p->cur;
    const size_t dd = (size_t)d * d;
    if (mu) CUDA_TRY(cudaMemcpyAsync(p->mu[c], mu, (size_t)S * d * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    if (pd) {
        CUDA_TRY(cudaMemcpyAsync(p->LD[c], pd, (size_t)S * dd * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        if (S > 1) {
            if (po) CUDA_TRY(cudaMemcpyAsync(p->LO[c], po, (size_t)(S - 1) * dd * sizeof(double), cudaMemcpyHostToDevice, p->stream));
            else CUDA_TRY(cudaMemsetAsync(p->LO[c], 0, (size_t)(S - 1) * dd * sizeof(double), p->stream));
        }
    }
    p->sweep_valid = false;
    p->asm_valid = false;
    p->grads_valid = false;
    p->zc_ok[0] = p->zc_ok[1] = false;
    p->batch.ldn_valid[0] = p->batch.ldn_valid[1] = false;
    p->batch.cost_valid = false;
    if (pd || !p->has_state) {
        if (!pd) return fail(GVIB200_ESTATE, "set_state: the first call must provide a precision");
        TRY(recompute_from_precision(p, c, async));
    }
    p->has_state = true;
    return 0;
}

extern "C" int gvib200_set_state(gvib200_problem* p, const double* mu, const double* pd, const double* po) {
    if (p && p->pending_check) TRY(resolve_pending(p));
    return set_state_impl(p, mu, pd, po, false);
}

// Asynchronous variant for pipelines of independent problems: the uploads (from PINNED host memory), the selected inverse
// and the factor marginals are enqueued on the problem's stream and the call returns; a precision that is not positive
// definite is reported by the next gvib200_ngd_iterate / gvib200_prox_iterate / gvib200_get_* / gvib200_sync on this handle.
extern "C" int gvib200_set_state_async(gvib200_problem* p, const double* mu, const double* pd, const double* po) {
    if (p && p->pending_check) TRY(resolve_pending(p));
    return set_state_impl(p, mu, pd, po, true);
}

extern "C" int gvib200_sync(gvib200_problem* p) {
    if (!p) return fail(GVIB200_EINVAL, "sync: null handle");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    TRY(resolve_pending(p));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (p->stream2) CUDA_TRY(cudaStreamSynchronize(p->stream2));
    return 0;
}

static int download(gvib200_problem* p, double* dst, const double* src, size_t n) {
    if (!dst || n == 0) return 0;
    CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    return 0;
}

static int get_mean_impl(gvib200_problem* p, double* mu, bool async) {
    if (!p || !p->has_state) return fail(GVIB200_ESTATE, "get_mean: no state");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    if (!async) TRY(resolve_pending(p));
    TRY(download(p, mu, p->mu[p->cur], (size_t)p->S * p->d));
    if (!async) CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}
static int get_blocks_impl(gvib200_problem* p, double* diag, double* off, bool cov, bool async) {
    if (!p || !p->has_state) return fail(GVIB200_ESTATE, "get blocks: no state");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    if (!async) TRY(resolve_pending(p));
    const size_t dd = (size_t)p->d * p->d;
    TRY(download(p, diag, cov ? p->CD[p->cur] : p->LD[p->cur], p->S * dd));
    TRY(download(p, off, cov ? p->CO[p->cur] : p->LO[p->cur], (p->S - 1) * dd));
    if (!async) CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}
extern "C" int gvib200_get_mean(gvib200_problem* p, double* mu) { return get_mean_impl(p, mu, false); }
extern "C" int gvib200_get_prec_blocks(gvib200_problem* p, double* diag, double* off) {
    return get_blocks_impl(p, diag, off, false, false);
}
extern "C" int gvib200_get_cov_blocks(gvib200_problem* p, double* diag, double* off) {
    return get_blocks_impl(p, diag, off, true, false);
}
// Asynchronous downloads into PINNED host memory: enqueued on the problem's stream behind everything already issued on the
// handle; the buffers are valid after gvib200_sync (or any synchronous call on the handle).
extern "C" int gvib200_get_mean_async(gvib200_problem* p, double* mu) { return get_mean_impl(p, mu, true); }
extern "C" int gvib200_get_prec_blocks_async(gvib200_problem* p, double* diag, double* off) {
    return get_blocks_impl(p, diag, off, false, true);
}
extern "C" int gvib200_get_cov_blocks_async(gvib200_problem* p, double* diag, double* off) {
    return get_blocks_impl(p, diag, off, true, true);
}

// ------------------------------------------------------------------------------------------------
// C-ABI: moments / cost / gradients
// ------------------------------------------------------------------------------------------------
static int ensure_sweep(gvib200_problem* p, bool want_raw) {
    TRY(resolve_pending(p));
    if (p->sweep_valid && !want_raw) return 0;
    TRY(run_sweep(p, p->cur, false, true, want_raw));
    run_total(p, p->cur);
    p->sweep_valid = true;
    return check_launch("sweep");
}

extern "C" int gvib200_moments(gvib200_problem* p, double* E0, double* E1, double* E2) {
    if (!p || !p->has_state) return fail(GVIB200_ESTATE, "moments: no state");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    TRY(ensure_sweep(p, true));
    size_t o0 = 0, o1 = 0, o2 = 0;
    for (auto& g : p->gh) {
        double *d0 = nullptr, *d1 = nullptr, *d2 = nullptr;
        TRY(dev_alloc(&d0, (size_t)g.n));
        TRY(dev_alloc(&d1, (size_t)g.n * g.dim));
        TRY(dev_alloc(&d2, (size_t)g.n * g.dim * g.dim));
        const double* SR = g.d_SR[p->cur];
        switch (g.dim) {
            case 1: launch_raw_to_x<1>(p, g, SR, d0, d1, d2); break;
            case 2: launch_raw_to_x<2>(p, g, SR, d0, d1, d2); break;
            case 3: launch_raw_to_x<3>(p, g, SR, d0, d1, d2); break;
            case 4: launch_raw_to_x<4>(p, g, SR, d0, d1, d2); break;
            case 6: launch_raw_to_x<6>(p, g, SR, d0, d1, d2); break;
            case 8: launch_raw_to_x<8>(p, g, SR, d0, d1, d2); break;
            case 12: launch_raw_to_x<12>(p, g, SR, d0, d1, d2); break;
            default: return fail(GVIB200_EINVAL, "moments: unsupported factor dim");
        }
        if (E0) TRY(download(p, E0 + o0, d0, (size_t)g.n));
        if (E1) TRY(download(p, E1 + o1, d1, (size_t)g.n * g.dim));
        if (E2) TRY(download(p, E2 + o2, d2, (size_t)g.n * g.dim * g.dim));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        cudaFree(d0);
        cudaFree(d1);
        cudaFree(d2);
        o0 += g.n;
        o1 += (size_t)g.n * g.dim;
        o2 += (size_t)g.n * g.dim * g.dim;
    }
    return check_launch("moments");
}

extern "C" int gvib200_cost(gvib200_problem* p, const double* mu, const double* pd, const double* po, double* cost,
                            double* fac_costs) {
    if (!p || !p->finalized) return fail(GVIB200_ESTATE, "cost: problem not finalized");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    const int S = p->S, d = p->d;
    const size_t dd = (size_t)d * d;
    int which;
    if (!mu && !pd) {
        if (!p->has_state) return fail(GVIB200_ESTATE, "cost: no state");
        TRY(ensure_sweep(p, false));
        which = p->cur;
    } else {
        if (!p->has_state && (!mu || !pd)) return fail(GVIB200_ESTATE, "cost: no state to take defaults from");
        which = 1 - p->cur;
        const int c = p->cur;
        if (mu) CUDA_TRY(cudaMemcpyAsync(p->mu[which], mu, (size_t)S * d * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        else CUDA_TRY(cudaMemcpyAsync(p->mu[which], p->mu[c], (size_t)S * d * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
        if (pd) {
            CUDA_TRY(cudaMemcpyAsync(p->LD[which], pd, S * dd * sizeof(double), cudaMemcpyHostToDevice, p->stream));
            if (S > 1) {
                if (po) CUDA_TRY(cudaMemcpyAsync(p->LO[which], po, (S - 1) * dd * sizeof(double), cudaMemcpyHostToDevice, p->stream));
                else CUDA_TRY(cudaMemsetAsync(p->LO[which], 0, (S - 1) * dd * sizeof(double), p->stream));
            }
        } else {
            CUDA_TRY(cudaMemcpyAsync(p->LD[which], p->LD[c], S * dd * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
            CUDA_TRY(cudaMemcpyAsync(p->LO[which], p->LO[c], S * dd * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
        }
        TRY(clear_flag(p));
        TRY(do_selinv(p, p->LD[which], p->LO[which], p->CD[which], p->CO[which], p->scal + which));
        TRY(run_sweep(p, which, true, false, false));
        run_total(p, which);
        int flag = 0;
        TRY(read_flag(p, &flag));
        if (flag) return fail(GVIB200_ENOTSPD, "cost: precision matrix is not positive definite");
    }
    CUDA_TRY(cudaMemcpyAsync(p->h_scal, p->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (fac_costs) TRY(download(p, fac_costs, p->fcost[which], (size_t)p->n_factors));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (cost) *cost = p->h_scal[2 + which];
    return check_launch("cost");
}

static int dispatch_assemble(gvib200_problem* p, int which, bool alt) {
    switch (p->d) {
        case 1: launch_assemble<1>(p, which, alt); break;
        case 2: launch_assemble<2>(p, which, alt); break;
        case 3: launch_assemble<3>(p, which, alt); break;
        case 4: launch_assemble<4>(p, which, alt); break;
        case 6: launch_assemble<6>(p, which, alt); break;
        default: return fail(GVIB200_EINVAL, "unsupported state dim");
    }
    if (!alt) p->asm_valid = false;  // callers that assemble the current sweep set it themselves
    return 0;
}

// gradients at the current state: assemble + dmu solve; leaves Vdmu/VD/VO/dmu on the device
static int compute_gradients(gvib200_problem* p) {
    TRY(ensure_sweep(p, false));
    TRY(dispatch_assemble(p, p->cur));
    TRY(do_solve(p, p->VD, p->VO, p->rhs, p->dmu, nullptr));
    p->grads_valid = true;
    return check_launch("gradients");
}

extern "C" int gvib200_gradients(gvib200_problem* p, double* dmu, double* dD, double* dO) {
    if (!p || !p->has_state) return fail(GVIB200_ESTATE, "gradients: no state");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    const int S = p->S, d = p->d;
    const size_t dd = (size_t)d * d;
    TRY(clear_flag(p));
    TRY(compute_gradients(p));
    int flag = 0;
    TRY(read_flag(p, &flag));
    TRY(download(p, dmu, p->dmu, (size_t)S * d));
    if (dD || dO) {
        // dprecision = Vddmu - precision, staged through the candidate precision buffers
        const int w = 1 - p->cur;
        LAUNCH(p, KC_OTHER, k_sub, cdiv(S * dd, 256), 256, 0, S * dd, p->VD, p->LD[p->cur], p->LD[w]);
        LAUNCH(p, KC_OTHER, k_sub, cdiv(S * dd, 256), 256, 0, S * dd, p->VO, p->LO[p->cur], p->LO[w]);
        TRY(download(p, dD, p->LD[w], S * dd));
        TRY(download(p, dO, p->LO[w], (S - 1) * dd));
    }
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (flag) return fail(GVIB200_ENOTSPD, "gradients: Vddmu is not positive definite (the reference would hand it to CG)");
    return check_launch("gradients");
}

extern "C" int gvib200_get_V(gvib200_problem* p, double* Vdmu, double* VD, double* VO) {
    if (!p || !p->grads_valid) return fail(GVIB200_ESTATE, "get_V: call gradients first");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    const size_t dd = (size_t)p->d * p->d;
    TRY(download(p, Vdmu, p->Vdmu, (size_t)p->S * p->d));
    TRY(download(p, VD, p->VD, p->S * dd));
    TRY(download(p, VO, p->VO, (p->S - 1) * dd));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C-ABI: optimizer
// ------------------------------------------------------------------------------------------------
extern "C" void gvib200_default_opts(gvib200_opts* o) {
    if (!o) return;
    o->step_size_base = 0.55;
    o->backtrack_ratio = 0.75;
    o->max_backtrack = 10;
    o->niters_lowtemp = 10;
    o->reuse_accepted_sweep = 0;
    o->ema_alpha = 1.0;
}

// GVIGH::switch_to_high_temperature (gvibase/GVI-GH-GBP-impl.h:18-27)
static int switch_to_high_temperature(gvib200_problem* p) {
    for (auto& g : p->gh) {
        g.T = g.Thigh;
        CUDA_TRY(cudaMemcpyAsync(g.d_T, g.T.data(), g.T.size() * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    }
    for (auto& g : p->lin) {
        g.T = g.Thigh;
        CUDA_TRY(cudaMemcpyAsync(g.d_T, g.T.data(), g.T.size() * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    }
    TRY(upload_klin(p));
    p->sweep_valid = false;
    p->asm_valid = false;
    p->grads_valid = false;
    return 0;
}

// which: 1 = mean, 2 = precision, 3 = both
static int launch_candidate(gvib200_problem* p, double alpha, int which = 3) {
    const int S = p->S, d = p->d, c = p->cur, w = 1 - p->cur;
    const size_t dd = (size_t)d * d;
    const size_t nmu = (which & 1) ? (size_t)S * d : 0, nD = (which & 2) ? S * dd : 0, nO = (which & 2) ? (S - 1) * dd : 0;
    const size_t nmax = nD > nmu ? nD : nmu;
    LAUNCH(p, KC_CANDIDATE, k_candidate, cdiv(nmax, 256), 256, 0, nmu, nD, nO, alpha, p->mu[c], p->dmu, p->LD[c], p->LO[c],
           p->VD, p->VO, p->mu[w], p->LD[w], p->LO[w]);
    return 0;
}

extern "C" int gvib200_ngd_iterate(gvib200_problem* p, const gvib200_opts* opts_in, gvib200_iter_stats* st) {
    if (!p || !p->has_state) return fail(GVIB200_ESTATE, "ngd_iterate: no state");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    gvib200_opts o;
    if (opts_in) o = *opts_in;
    else gvib200_default_opts(&o);
    gvib200_iter_stats s;
    std::memset(&s, 0, sizeof(s));
    TRY(resolve_pending(p));  // gvib200_set_state_async: its flag is read before this iteration clears it
    if (p->converged) {
        s.converged = 1;
        if (st) *st = s;
        return 0;
    }
    // temperature switch at iteration niters_lowtemp (GVI-GH-GBP-impl.h:49-58)
    if (p->iter == o.niters_lowtemp && p->is_lowtemp) {
        TRY(switch_to_high_temperature(p));
        p->is_lowtemp = false;
        s.switched_high_T = 1;
    }
    TRY(clear_flag(p));
    // cost_iter + factor costs + gradients from ONE full-moment sweep at the current state
    if (!p->sweep_valid) s.n_moment_sweeps++;
    TRY(ensure_sweep(p, false));
    if (!p->asm_valid) {
        TRY(dispatch_assemble(p, p->cur));
        p->asm_valid = true;
    }
    // back-tracking (GVI-GH-GBP-impl.h:82-124).  The first trial's candidate precision Lambda + a (Vddmu - Lambda) does
    // not depend on dmu, so its selected inverse runs on the side stream while the main stream solves for dmu.
    int cnt = 0;
    int flag_solve = 0, flag_inv = 0;
    double step = o.step_size_base;
    double cost_iter = 0.0;
    while (true) {
        step *= o.backtrack_ratio;
        const int w = 1 - p->cur;
        bool linear_forked = false;
        static const bool no_fork_env = getenv("GVIB200_NO_FORK") != nullptr;  // development switch
        // multi-GPU: the side stream needs its own communicator (collectives of one communicator stay on one stream)
        const bool fork = (p->ctx->world == 1 || p->ctx->nccl_comm2 != nullptr) && !no_fork_env;
        if (cnt == 0 && !fork) {
            ChainFuse fi;
            fi.Dg2 = p->VD;
            fi.Og2 = p->VO;
            fi.alpha = step;
            fi.Dout = p->LD[w];
            fi.Oout = p->LO[w];
            TRY(do_selinv(p, p->LD[p->cur], p->LO[p->cur], p->CD[w], p->CO[w], p->scal + w, 1, &fi));
            TRY(run_prologue_only(p, w));
            ChainFuse fs;
            fs.xbase = p->mu[p->cur];
            fs.xalpha = step;
            fs.xout = p->mu[w];
            TRY(do_solve(p, p->VD, p->VO, p->rhs, p->dmu, nullptr, 0, &fs));
            p->grads_valid = true;
        } else if (cnt == 0) {
            // Critical path on the main stream: selected inverse of the candidate precision -> factor marginals ->
            // quadrature sweep.  The candidate precision Lambda + a (Vddmu - Lambda) is formed while the selected
            // inverse loads its tiles, the candidate mean mu + a dmu while the solve stores dmu: no separate candidate
            // kernel.  The side stream (higher priority, so that its CTAs are placed as soon as SMs free up) carries what
            // is off the critical path: the dmu solve and the closed-form linear factors of the candidate (HBM bound),
            // which run underneath the FP64-bound quadrature sweep.
            CUDA_TRY(cudaEventRecord(p->ev_fork, p->stream));
            ChainFuse fi;
            fi.Dg2 = p->VD;
            fi.Og2 = p->VO;
            fi.alpha = step;
            fi.Dout = p->LD[w];
            fi.Oout = p->LO[w];
            TRY(do_selinv(p, p->LD[p->cur], p->LO[p->cur], p->CD[w], p->CO[w], p->scal + w, 1, &fi));
            CUDA_TRY(cudaEventRecord(p->ev_pro, p->stream));  // candidate covariance blocks are complete
            CUDA_TRY(cudaStreamWaitEvent(p->stream2, p->ev_fork, 0));
            p->ls = p->stream2;
            ChainFuse fs;
            fs.xbase = p->mu[p->cur];
            fs.xalpha = step;
            fs.xout = p->mu[w];
            int rc = do_solve(p, p->VD, p->VO, p->rhs, p->dmu, nullptr, 0, &fs);
            p->ls = p->stream;
            if (rc != 0) return rc;
            p->grads_valid = true;
            CUDA_TRY(cudaEventRecord(p->ev_mu, p->stream2));  // candidate mean is complete
            // covariance part of the closed-form linear factors (HBM bound), behind the solve pass on its stream.  (Measured and
            // rejected, twice: starting it next to the backward half of the solve pass -- on a third stream, or on the main
            // stream right behind the selected inverse -- makes the quadrature kernel a little faster and the iteration
            // 0.402 -> 0.43 ms.)
            // (Also measured and rejected here: both halves in ONE pass per group right behind the solve pass -- the linear
            // factors' launches take half the time, 0.124 -> 0.064 ms per iteration, and the iteration 0.399 -> 0.406 ms.
            // GVIB200_LINEAR_ONEPASS switches it on; every other caller of run_linear uses the one-pass form.)
            static const bool lin_split_env = getenv("GVIB200_LINEAR_ONEPASS") == nullptr;
            CUDA_TRY(cudaStreamWaitEvent(p->stream2, p->ev_pro, 0));
            p->ls = p->stream2;
            {
                SweepTarget t{p->mu[w], p->CD[w], p->CO[w], w};
                rc = run_linear(p, t, o.reuse_accepted_sweep != 0, lin_split_env ? 1 : 3);
            }
            p->ls = p->stream;
            if (rc != 0) return rc;
            TRY(run_prologue_only(p, w));
            CUDA_TRY(cudaStreamWaitEvent(p->stream, p->ev_mu, 0));
            linear_forked = true;
        } else {
            TRY(launch_candidate(p, step, 3));
            TRY(do_selinv(p, p->LD[w], p->LO[w], p->CD[w], p->CO[w], p->scal + w, 1));
        }
        TRY(run_sweep(p, w, cnt != 0, o.reuse_accepted_sweep != 0, false, !linear_forked));
        if (linear_forked) {
            // the linear factors start on the side stream once the candidate is complete and the (short, latency bound)
            // culling pass of the quadrature sweep is through; they then run underneath the moment kernel
            // mean part of the linear factors: as soon as the dmu solve (same stream) has delivered the candidate mean;
            // it ends up underneath the moment kernel.  (Measured: neither holding it back behind the culling pass nor a
            // lowest-priority stream that only fills the moment kernel's last partial wave is any faster.)
            static const bool lin_split_env2 = getenv("GVIB200_LINEAR_ONEPASS") == nullptr;
            if (lin_split_env2) {
                p->ls = p->stream2;
                SweepTarget t{p->mu[w], p->CD[w], p->CO[w], w};
                const int rc = run_linear(p, t, o.reuse_accepted_sweep != 0, 2);
                p->ls = p->stream;
                if (rc != 0) return rc;
            }
            CUDA_TRY(cudaEventRecord(p->ev_join, p->stream2));
            CUDA_TRY(cudaStreamWaitEvent(p->stream, p->ev_join, 0));
        }
        if (o.reuse_accepted_sweep) s.n_moment_sweeps++;
        else s.n_cost_sweeps++;
        // speculation: if this trial is accepted its sweep is the next iteration's gradient sweep -- assemble it now,
        // into the second set of buffers, so that the host round trip below costs no device time
        const bool speculate = (o.reuse_accepted_sweep != 0);
        const bool side_total = linear_forked && speculate && (p->ctx->world == 1 || p->ctx->mbox_on) && p->n_factors > 8192 &&
                                p->zc_ok[p->cur];
        if (side_total) {
            // the total cost (a short latency-bound kernel whose result only the host needs) moves to the side stream,
            // next to the speculative assembly on the main stream; k_total hands cost and flags over in mapped memory
            CUDA_TRY(cudaEventRecord(p->ev_k1, p->stream));
            CUDA_TRY(cudaStreamWaitEvent(p->stream2, p->ev_k1, 0));
            p->ls = p->stream2;
            run_total(p, w);
            p->ls = p->stream;
            CUDA_TRY(cudaEventRecord(p->ev_host, p->stream2));
            TRY(dispatch_assemble(p, w, true));
            CUDA_TRY(cudaEventSynchronize(p->ev_host));
            flag_solve = p->zc[2] != 0.0;
            flag_inv = p->zc[3] != 0.0;
            p->h_scal[2] = p->zc[0];
            p->h_scal[3] = p->zc[1];
        } else {
            run_total(p, w);
            // k_total hands the cost and the flags to the host through mapped memory; copies only where that is not valid
            const bool zc = p->zc_ok[w] && p->zc_ok[p->cur];
            if (!zc) CUDA_TRY(cudaMemcpyAsync(p->h_scal, p->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
            TRY(read_flags(p, &flag_solve, &flag_inv, speculate ? w : -1, zc));
            if (zc) {
                p->h_scal[2] = p->zc[0];
                p->h_scal[3] = p->zc[1];
            }
        }
        if (cnt == 0) {
            cost_iter = p->h_scal[2 + p->cur];
            s.cost = cost_iter;
            if (flag_solve) {
                // An indefinite Vddmu has no Cholesky factor: the step is not taken.  Everything the trial touched
                // lives in the candidate buffers (mu / precision / covariance [w], the second V set), so the
                // current state is exactly what it was; the iteration is not counted and the handle stays usable.
                // (The reference hands the matrix to CG, ngd/NGD-GH-impl.h:59-60, and continues with whatever comes
                // back; the oracle's direct solve reports it the same way as this library.)
                s.status = GVIB200_ENOTSPD;
                if (st) *st = s;
                p->grads_valid = false;
                TRY(clear_flag(p));
                return fail(GVIB200_ENOTSPD, "ngd_iterate: Vddmu is not positive definite (state unchanged)");
            }
        }
        const double new_cost = p->h_scal[2 + w];
        s.new_cost = new_cost;
        // a candidate precision that is not SPD has no finite cost: treat as a rejected trial
        const bool ok = (flag_inv == 0) && (new_cost < cost_iter);
        if (flag_inv) TRY(clear_flag(p));
        if (ok && o.ema_alpha != 1.0) {
            // EMA of the reference's GPU path (gvibase/GVI-GH-Cuda-impl.h:112-114): the proposal that is taken is
            //   alpha * new + (1 - alpha) * current  =  mu + (alpha step) dmu,  Lambda + (alpha step)(Vddmu - Lambda),
            // i.e. the candidate at the step alpha * step (whose cost is NOT evaluated: the trial at `step` decided).
            // Rebuilt in the candidate buffers with its own selected inverse and factor marginals, then flipped.
            TRY(launch_candidate(p, o.ema_alpha * step, 3));
            TRY(clear_flag(p));
            TRY(do_selinv(p, p->LD[w], p->LO[w], p->CD[w], p->CO[w], p->scal + w, 1));
            TRY(run_prologue_only(p, w));
            int f0 = 0, f1 = 0;
            TRY(read_flags(p, &f0, &f1));
            if (f0 | f1) {
                TRY(clear_flag(p));
                return fail(GVIB200_ENOTSPD, "ngd_iterate: the EMA proposal is not positive definite");
            }
            p->cur = w;
            p->sweep_valid = false;
            p->grads_valid = false;
            p->asm_valid = false;
            p->zc_ok[0] = p->zc_ok[1] = false;
            s.accepted = 1;
            s.step = step;
            s.n_backtrack = cnt;
            break;
        }
        if (ok) {
            // update_proposal (ngd/NGD-GH-impl.h:151-156): the candidate's mu, precision, covariance and factor
            // marginals become current -- a buffer flip, everything is already on the device
            p->cur = w;
            p->sweep_valid = (o.reuse_accepted_sweep != 0);
            p->grads_valid = false;
            p->asm_valid = false;
            if (speculate) {
                std::swap(p->Vdmu, p->Vdmu2);
                std::swap(p->VD, p->VD2);
                std::swap(p->VO, p->VO2);
                std::swap(p->rhs, p->rhs2);
                p->asm_valid = true;
            }
            s.accepted = 1;
            s.step = step;
            s.n_backtrack = cnt;
            break;
        }
        cnt++;
        if (cnt > o.max_backtrack) {
            if (p->is_lowtemp) {
                TRY(switch_to_high_temperature(p));
                p->is_lowtemp = false;
                s.switched_high_T = 1;
            } else {
                p->converged = true;
                s.converged = 1;
            }
            s.n_backtrack = cnt;
            break;
        }
    }
    p->iter++;
    if (st) *st = s;
    return 0;
}

extern "C" int gvib200_optimize(gvib200_problem* p, const gvib200_opts* opts, int n_iters, gvib200_iter_stats* stats,
                                int* n_done, double* fac_costs_trace, double* mean_trace) {
    if (!p || !p->has_state) return fail(GVIB200_ESTATE, "optimize: no state");
    int done = 0;
    for (int it = 0; it < n_iters; ++it) {
        if (p->converged) break;
        gvib200_iter_stats s;
        if (mean_trace) TRY(gvib200_get_mean(p, mean_trace + (size_t)it * p->S * p->d));
        int rc = gvib200_ngd_iterate(p, opts, &s);
        if (stats) stats[it] = s;
        if (rc != 0) {
            if (n_done) *n_done = done;
            return rc;
        }
        // factor costs recorded at the start of the iteration: still in the pre-flip buffer
        if (fac_costs_trace) {
            const int src = s.accepted ? 1 - p->cur : p->cur;
            TRY(download(p, fac_costs_trace + (size_t)it * p->n_factors, p->fcost[src], (size_t)p->n_factors));
            CUDA_TRY(cudaStreamSynchronize(p->stream));
        }
        done++;
    }
    if (n_done) *n_done = done;
    return 0;
}


// ------------------------------------------------------------------------------------------------
// C-ABI: batches of independent problems, line search PER PROBLEM.  The problems are concatenated block-diagonally into
// one chain (nothing couples consecutive problems), so sweeps, assembly and chain passes run over the whole batch in single
// launches; what is per problem is what GVIGH::optimize keeps per optimizer object (gvibase/GVI-GH-GBP-impl.h:33-130): the
// cost (factor costs + log det / 2 of ITS precision), the back-tracking count and step size, the temperature phase and the
// converged flag.  A trial is run for the whole batch with one step size per problem; problems that accepted keep their
// step (their candidate is recomputed bit for bit), problems that rejected shrink theirs, problems that are through take
// step 0 (candidate == current state) -- until every problem has accepted or exhausted its back-tracking.
// ------------------------------------------------------------------------------------------------
static void free_batch(gvib200_problem* p) {
    auto& B = p->batch;
    auto F = [](void* q) {
        if (q) cudaFree(q);
    };
    F(B.d_sprob); F(B.d_soff); F(B.d_seg); F(B.d_gfirst); F(B.d_step); F(B.d_alpha_node); F(B.d_ldn[0]); F(B.d_ldn[1]);
    F(B.d_ldn_solve); F(B.d_pcost); F(B.d_pcost2);
    if (B.h_pcost) cudaFreeHost(B.h_pcost);
    if (B.h_pcost2) cudaFreeHost(B.h_pcost2);
    if (B.h_step) cudaFreeHost(B.h_step);
    B = gvib200_problem::Batch();
}

extern "C" int gvib200_set_batch(gvib200_problem* p, int n_problems, const int32_t* state_offsets) {
    if (!p || !p->finalized) return fail(GVIB200_ESTATE, "set_batch: problem not finalized");
    if (n_problems < 1 || !state_offsets) return fail(GVIB200_EINVAL, "set_batch: bad arguments");
    if (p->prox) return fail(GVIB200_EINVAL, "set_batch: not available for Prox-GVI problems");
    if (p->ctx->world > 1) return fail(GVIB200_EINVAL, "set_batch: a batch lives on one GPU (shard the problems over the ranks)");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    free_batch(p);
    auto& B = p->batch;
    const int P = n_problems, S = p->S, d = p->d;
    if (state_offsets[0] != 0 || state_offsets[P] != S) return fail(GVIB200_EINVAL, "set_batch: offsets must run from 0 to num_states");
    B.soff.assign(state_offsets, state_offsets + P + 1);
    B.sprob_h.assign((size_t)S, 0);
    for (int q = 0; q < P; ++q) {
        if (B.soff[q + 1] <= B.soff[q]) return fail(GVIB200_EINVAL, "set_batch: offsets must increase");
        for (int st = B.soff[q]; st < B.soff[q + 1]; ++st) B.sprob_h[(size_t)st] = q;
    }
    // every group's factors must be ordered by problem and must not span two problems
    std::vector<int> seg, gfirst;
    auto add_group = [&](const std::vector<int>& start, int span, int first_id) -> bool {
        const int n = (int)start.size();
        for (int f = 0; f < n; ++f) {
            if (f && B.sprob_h[(size_t)start[f]] < B.sprob_h[(size_t)start[f - 1]]) return false;
            if (B.sprob_h[(size_t)start[f]] != B.sprob_h[(size_t)(start[f] + span - 1)]) return false;
        }
        int f = 0;
        for (int q = 0; q < P; ++q) {
            seg.push_back(f);
            while (f < n && B.sprob_h[(size_t)start[f]] == q) ++f;
        }
        seg.push_back(n);
        gfirst.push_back(first_id);
        return true;
    };
    for (auto& g : p->gh)
        if (!add_group(g.start, std::max(1, g.dim / d), g.first_id))
            return fail(GVIB200_EINVAL, "set_batch: factors must be ordered by problem and stay inside one problem");
    for (auto& g : p->lin)
        if (!add_group(g.start, std::max(1, g.dim / d), g.first_id))
            return fail(GVIB200_EINVAL, "set_batch: factors must be ordered by problem and stay inside one problem");
    B.P = P;
    B.G = (int)gfirst.size();
    TRY(dev_upload(&B.d_sprob, B.sprob_h, p->stream));
    TRY(dev_upload(&B.d_soff, B.soff, p->stream));
    TRY(dev_upload(&B.d_seg, seg, p->stream));
    TRY(dev_upload(&B.d_gfirst, gfirst, p->stream));
    TRY(dev_alloc(&B.d_step, (size_t)P));
    TRY(dev_alloc(&B.d_alpha_node, (size_t)S));
    TRY(dev_alloc(&B.d_ldn[0], (size_t)S));
    TRY(dev_alloc(&B.d_ldn[1], (size_t)S));
    TRY(dev_alloc(&B.d_ldn_solve, (size_t)S));
    TRY(dev_alloc(&B.d_pcost, (size_t)P));
    TRY(dev_alloc(&B.d_pcost2, (size_t)P));
    CUDA_TRY(cudaMallocHost((void**)&B.h_pcost, (size_t)P * sizeof(double)));
    CUDA_TRY(cudaMallocHost((void**)&B.h_pcost2, (size_t)P * sizeof(double)));
    CUDA_TRY(cudaMallocHost((void**)&B.h_step, (size_t)P * sizeof(double)));
    B.cost_cur.assign((size_t)P, 0.0);
    B.lowtemp.assign((size_t)P, p->is_lowtemp ? 1 : 0);
    B.converged.assign((size_t)P, 0);
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

// per-problem sums of buffer `which`'s factor costs + log det / 2 (or, fcost_on = false, of the log pivots in `ldn` alone)
static int batch_costs(gvib200_problem* p, int which, const double* ldn, bool fcost_on, double* host_out) {
    auto& B = p->batch;
    LAUNCH(p, KC_SUM, k_problem_costs, cdiv((long long)B.P * 32, 128), 128, 0, B.P, B.G, B.d_seg, B.d_gfirst,
           fcost_on ? p->fcost[which] : nullptr, B.d_soff, ldn, fcost_on ? 0.5 : 1.0, B.d_pcost);
    CUDA_TRY(cudaMemcpyAsync(host_out, B.d_pcost, (size_t)B.P * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return check_launch("k_problem_costs");
}

// temperature switch of the problems marked in `mask` (GVIGH::switch_to_high_temperature, one optimizer object each)
static int batch_switch_high_T(gvib200_problem* p, const std::vector<char>& mask) {
    auto& B = p->batch;
    for (auto& g : p->gh) {
        for (int f = 0; f < g.n; ++f)
            if (mask[(size_t)B.sprob_h[(size_t)g.start[f]]]) g.T[f] = g.Thigh[f];
        CUDA_TRY(cudaMemcpyAsync(g.d_T, g.T.data(), g.T.size() * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    }
    for (auto& g : p->lin) {
        for (int f = 0; f < g.n; ++f)
            if (mask[(size_t)B.sprob_h[(size_t)g.start[f]]]) g.T[f] = g.Thigh[f];
        CUDA_TRY(cudaMemcpyAsync(g.d_T, g.T.data(), g.T.size() * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    }
    TRY(upload_klin(p));
    p->sweep_valid = false;
    p->asm_valid = false;
    p->grads_valid = false;
    B.cost_valid = false;
    return 0;
}

extern "C" int gvib200_batch_iterate(gvib200_problem* p, const gvib200_opts* opts_in, gvib200_iter_stats* st, int* n_trials) {
    if (!p || !p->has_state) return fail(GVIB200_ESTATE, "batch_iterate: no state");
    auto& B = p->batch;
    if (B.P < 1) return fail(GVIB200_ESTATE, "batch_iterate: call gvib200_set_batch first");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    gvib200_opts o;
    if (opts_in) o = *opts_in;
    else gvib200_default_opts(&o);
    if (o.ema_alpha != 1.0) return fail(GVIB200_EINVAL, "batch_iterate: ema_alpha != 1 is not available for batches");
    TRY(resolve_pending(p));
    const int P = B.P, S = p->S, d = p->d;
    std::vector<gvib200_iter_stats> stats((size_t)P);
    std::memset(stats.data(), 0, stats.size() * sizeof(gvib200_iter_stats));
    p->ls = p->stream;
    // temperature switch at iteration niters_lowtemp (GVI-GH-GBP-impl.h:49-58), for every problem still at low temperature
    if (p->iter == o.niters_lowtemp) {
        std::vector<char> mask((size_t)P, 0);
        bool any = false;
        for (int q = 0; q < P; ++q)
            if (B.lowtemp[q] && !B.converged[q]) {
                mask[q] = 1;
                B.lowtemp[q] = 0;
                stats[q].switched_high_T = 1;
                any = true;
            }
        if (any) TRY(batch_switch_high_T(p, mask));
    }
    TRY(clear_flag(p));
    // sweep at the current state (kept from the accepted trial when reuse_accepted_sweep), per-problem cost_iter
    TRY(ensure_sweep(p, false));
    if (!B.ldn_valid[p->cur]) {
        ChainFuse f;
        f.ldnode = B.d_ldn[p->cur];
        TRY(do_selinv(p, p->LD[p->cur], p->LO[p->cur], p->CD[p->cur], p->CO[p->cur], p->scal + p->cur, 1, &f));
        B.ldn_valid[p->cur] = true;
        B.cost_valid = false;
    }
    if (!B.cost_valid) {
        TRY(batch_costs(p, p->cur, B.d_ldn[p->cur], true, B.h_pcost));
        for (int q = 0; q < P; ++q) B.cost_cur[q] = B.h_pcost[q];
        B.cost_valid = true;
    }
    if (!p->asm_valid) {
        TRY(dispatch_assemble(p, p->cur));
        p->asm_valid = true;
    }
    // dmu = -Vddmu^-1 Vdmu for the whole batch on the side stream, next to the first trial's selected inverse (whose candidate
    // precision does not depend on dmu); the per-problem sums of its log pivots (not finite = that problem's Vddmu is not
    // SPD) are read together with the first trial's costs: one host round trip per trial
    TRY(clear_flag(p));  // the batch-wide flags say nothing about a single problem: the log pivots do
    CUDA_TRY(cudaEventRecord(p->ev_fork, p->stream));
    CUDA_TRY(cudaStreamWaitEvent(p->stream2, p->ev_fork, 0));
    p->ls = p->stream2;
    {
        ChainFuse f;
        f.ldnode = B.d_ldn_solve;
        const int rc = do_solve(p, p->VD, p->VO, p->rhs, p->dmu, nullptr, 0, &f);
        if (rc == 0)
            LAUNCH(p, KC_SUM, k_problem_costs, cdiv((long long)P * 32, 128), 128, 0, P, B.G, B.d_seg, B.d_gfirst, (const double*)nullptr,
                   B.d_soff, B.d_ldn_solve, 1.0, B.d_pcost2);
        p->ls = p->stream;
        if (rc != 0) return rc;
        CUDA_TRY(cudaMemcpyAsync(B.h_pcost2, B.d_pcost2, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, p->stream2));
        CUDA_TRY(cudaEventRecord(p->ev_mu, p->stream2));
        p->grads_valid = true;
    }
    enum { RUN = 0, ACCEPTED, PARK, DONE };  // PARK: through, but one more trial with step 0 makes its candidate the current state
    std::vector<int> phase((size_t)P, RUN), cnt((size_t)P, 0);
    std::vector<double> step((size_t)P, o.step_size_base);
    for (int q = 0; q < P; ++q) {
        stats[q].cost = B.cost_cur[q];
        stats[q].new_cost = B.cost_cur[q];
        if (B.converged[q]) {
            stats[q].converged = 1;
            phase[q] = PARK;
        }
    }
    const int w = 1 - p->cur, c = p->cur;
    int trials = 0;
    while (true) {
        for (int q = 0; q < P; ++q) {
            if (phase[q] == RUN) step[q] *= o.backtrack_ratio;
            B.h_step[q] = (phase[q] == RUN || phase[q] == ACCEPTED) ? step[q] : 0.0;
        }
        CUDA_TRY(cudaMemcpyAsync(B.d_step, B.h_step, (size_t)P * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        LAUNCH(p, KC_CANDIDATE, k_batch_alpha, cdiv(S, 256), 256, 0, S, B.d_sprob, B.d_step, B.d_alpha_node);
        ChainFuse fi;
        fi.Dg2 = p->VD;
        fi.Og2 = p->VO;
        fi.alpha_node = B.d_alpha_node;
        fi.Dout = p->LD[w];
        fi.Oout = p->LO[w];
        fi.ldnode = B.d_ldn[w];
        TRY(do_selinv(p, p->LD[c], p->LO[c], p->CD[w], p->CO[w], p->scal + w, 1, &fi));
        TRY(run_prologue_only(p, w));
        if (trials == 0) CUDA_TRY(cudaStreamWaitEvent(p->stream, p->ev_mu, 0));  // dmu (and the solve's log pivots on the host)
        LAUNCH(p, KC_CANDIDATE, k_batch_candidate_mu, cdiv((long long)S * d, 256), 256, 0, (size_t)S * d, d, B.d_alpha_node,
               p->mu[c], p->dmu, p->mu[w]);
        TRY(run_sweep(p, w, false, true, false));
        TRY(batch_costs(p, w, B.d_ldn[w], true, B.h_pcost));
        TRY(clear_flag(p));
        ++trials;
        bool again = false;
        for (int q = 0; q < P; ++q) {
            if (trials == 1 && phase[q] == RUN && !std::isfinite(B.h_pcost2[q])) {
                // this problem's Vddmu has no Cholesky factor: its step is not taken (whatever the trial produced from the
                // garbage of its solve is discarded); one more trial with step 0 parks it at its current state
                stats[q].status = GVIB200_ENOTSPD;
                phase[q] = PARK;
                again = true;
                continue;
            }
            if (phase[q] == PARK) {
                phase[q] = DONE;  // this trial ran it with step 0
            } else if (phase[q] == RUN) {
                const double nc = B.h_pcost[q];
                stats[q].new_cost = nc;
                if (nc < B.cost_cur[q]) {  // NaN (candidate precision not SPD) compares false: a rejected trial
                    phase[q] = ACCEPTED;
                    stats[q].accepted = 1;
                    stats[q].step = step[q];
                    stats[q].n_backtrack = cnt[q];
                } else {
                    cnt[q]++;
                    if (cnt[q] > o.max_backtrack) {  // GVI-GH-GBP-impl.h:104-119
                        stats[q].n_backtrack = cnt[q];
                        if (B.lowtemp[q]) stats[q].switched_high_T = 1;
                        else stats[q].converged = 1;
                        phase[q] = PARK;
                        again = true;
                    } else {
                        again = true;
                    }
                }
            }
        }
        if (!again) break;
    }
    // every problem's candidate is now what it takes into the next iteration: the batch flips as one
    std::vector<char> mask((size_t)P, 0);
    bool any_switch = false, all_conv = true;
    for (int q = 0; q < P; ++q) {
        if (stats[q].accepted) B.cost_cur[q] = stats[q].new_cost;
        if (stats[q].switched_high_T && B.lowtemp[q] && !stats[q].accepted) {
            mask[q] = 1;
            B.lowtemp[q] = 0;
            any_switch = true;
        }
        if (stats[q].converged) B.converged[q] = 1;
        all_conv = all_conv && B.converged[q];
    }
    p->cur = w;
    B.ldn_valid[w] = true;
    p->sweep_valid = true;
    p->asm_valid = false;
    p->grads_valid = false;
    p->zc_ok[0] = p->zc_ok[1] = false;
    run_total(p, w);  // the batch total (gvib200_cost etc. read it)
    if (any_switch) TRY(batch_switch_high_T(p, mask));
    p->converged = all_conv;
    p->iter++;
    if (st) std::memcpy(st, stats.data(), stats.size() * sizeof(gvib200_iter_stats));
    if (n_trials) *n_trials = trials;
    return check_launch("batch_iterate");
}

// GVIGH::optimize of every problem of the batch: up to n_iters iterations (all problems iterate in lock step; a problem that
// has converged stays where it is), stats[it * n_problems + q] = problem q's record of iteration it
extern "C" int gvib200_batch_optimize(gvib200_problem* p, const gvib200_opts* opts, int n_iters, gvib200_iter_stats* stats,
                                      int* n_done) {
    if (!p || !p->has_state || p->batch.P < 1) return fail(GVIB200_ESTATE, "batch_optimize: no batch / no state");
    int done = 0;
    for (int it = 0; it < n_iters; ++it) {
        if (p->converged) break;
        const int rc = gvib200_batch_iterate(p, opts, stats ? stats + (size_t)it * p->batch.P : nullptr, nullptr);
        if (rc != 0) {
            if (n_done) *n_done = done;
            return rc;
        }
        done++;
    }
    if (n_done) *n_done = done;
    return 0;
}

extern "C" int gvib200_batch_costs(gvib200_problem* p, double* cost_per_problem) {
    if (!p || !p->has_state || p->batch.P < 1 || !cost_per_problem) return fail(GVIB200_ESTATE, "batch_costs: no batch / no state");
    auto& B = p->batch;
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    TRY(resolve_pending(p));
    p->ls = p->stream;
    TRY(ensure_sweep(p, false));
    if (!B.ldn_valid[p->cur]) {
        ChainFuse f;
        f.ldnode = B.d_ldn[p->cur];
        TRY(clear_flag(p));
        TRY(do_selinv(p, p->LD[p->cur], p->LO[p->cur], p->CD[p->cur], p->CO[p->cur], p->scal + p->cur, 1, &f));
        B.ldn_valid[p->cur] = true;
        B.cost_valid = false;
    }
    if (!B.cost_valid) {
        TRY(batch_costs(p, p->cur, B.d_ldn[p->cur], true, B.h_pcost));
        for (int q = 0; q < B.P; ++q) B.cost_cur[q] = B.h_pcost[q];
        B.cost_valid = true;
    }
    for (int q = 0; q < B.P; ++q) cost_per_problem[q] = B.cost_cur[q];
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C-ABI: result recorder (helpers/DataRecorder.h, banded)
// ------------------------------------------------------------------------------------------------
extern "C" int gvib200_optimize_traced(gvib200_problem* p, const gvib200_opts* opts, int n_iters, int prox,
                                       gvib200_iter_stats* stats, int* n_done, gvib200_trace* tr) {
    if (!p || !p->has_state) return fail(GVIB200_ESTATE, "optimize_traced: no state");
    if (!tr) return fail(GVIB200_EINVAL, "optimize_traced: null trace");
    const size_t nmu = (size_t)p->S * p->d, nblk = (size_t)p->S * p->d * p->d, noff = (size_t)(p->S - 1) * p->d * p->d;
    tr->n_recorded = 0;
    int done = 0;
    for (int it = 0; it < n_iters; ++it) {
        if (p->converged) break;
        const bool rec = it < tr->capacity;
        // the state before the step (the covariance / precision blocks are resident on the device)
        if (rec && tr->mean) TRY(gvib200_get_mean(p, tr->mean + (size_t)it * nmu));
        if (rec && tr->cov_diag) {
            TRY(download(p, tr->cov_diag + (size_t)it * nblk, p->CD[p->cur], nblk));
            CUDA_TRY(cudaStreamSynchronize(p->stream));
        }
        if (rec && tr->prec_diag) {
            TRY(download(p, tr->prec_diag + (size_t)it * nblk, p->LD[p->cur], nblk));
            CUDA_TRY(cudaStreamSynchronize(p->stream));
        }
        if (rec && tr->cov_off && noff) {
            TRY(download(p, tr->cov_off + (size_t)it * noff, p->CO[p->cur], noff));
            CUDA_TRY(cudaStreamSynchronize(p->stream));
        }
        if (rec && tr->prec_off && noff) {
            TRY(download(p, tr->prec_off + (size_t)it * noff, p->LO[p->cur], noff));
            CUDA_TRY(cudaStreamSynchronize(p->stream));
        }
        gvib200_iter_stats s;
        const int rc = prox ? gvib200_prox_iterate(p, opts, &s) : gvib200_ngd_iterate(p, opts, &s);
        if (stats) stats[it] = s;
        if (rc != 0) {
            if (n_done) *n_done = done;
            return rc;
        }
        if (rec && tr->cost) tr->cost[it] = s.cost;
        if (rec && tr->fac_costs) {  // factor costs of the iteration's start state: still in the pre-flip buffer
            const int src = s.accepted ? 1 - p->cur : p->cur;
            TRY(download(p, tr->fac_costs + (size_t)it * p->n_factors, p->fcost[src], (size_t)p->n_factors));
            CUDA_TRY(cudaStreamSynchronize(p->stream));
        }
        if (rec) tr->n_recorded = it + 1;
        done++;
    }
    if (n_done) *n_done = done;
    return 0;
}

extern "C" int gvib200_csv_write(const char* path, int rows, int cols, const double* a) {
    if (!path || rows < 0 || cols < 0 || (!a && rows * cols > 0)) return fail(GVIB200_EINVAL, "csv_write: bad arguments");
    FILE* f = std::fopen(path, "w");
    if (!f) return fail(GVIB200_EINVAL, std::string("csv_write: cannot open ") + path);
    for (int i = 0; i < rows; ++i) {
        for (int j = 0; j < cols; ++j) std::fprintf(f, j ? ", %.15g" : "%.15g", a[(size_t)i + (size_t)j * rows]);
        if (i + 1 < rows) std::fputc('\n', f);  // Eigen's IOFormat puts the row separator BETWEEN rows
    }
    const bool ok = std::fclose(f) == 0;
    return ok ? 0 : fail(GVIB200_EINVAL, std::string("csv_write: write failed: ") + path);
}

extern "C" int gvib200_trace_save(const gvib200_trace* tr, int S, int d, int n_factors, const char* prefix, const char* afterfix,
                                  int dense_limit) {
    if (!tr || S < 1 || d < 1) return fail(GVIB200_EINVAL, "trace_save: bad arguments");
    const int n = tr->n_recorded;  // "early ended": only the recorded iterations are written (DataRecorder.h:179-181)
    if (n < 1) return fail(GVIB200_ESTATE, "trace_save: nothing recorded");
    const std::string pre = prefix ? prefix : "", post = (afterfix && afterfix[0]) ? std::string("_") + afterfix : "";
    auto name = [&](const char* base) { return pre + base + post + ".csv"; };
    const int nmu = S * d, nblk = S * d * d;
    if (tr->mean) TRY(gvib200_csv_write(name("mean").c_str(), nmu, n, tr->mean));
    if (tr->cov_diag) TRY(gvib200_csv_write(name("cov").c_str(), nblk, n, tr->cov_diag));
    if (tr->prec_diag) TRY(gvib200_csv_write(name("precision").c_str(), nblk, n, tr->prec_diag));
    if (tr->cost) TRY(gvib200_csv_write(name("cost").c_str(), n, 1, tr->cost));
    if (tr->fac_costs) TRY(gvib200_csv_write(name("factor_costs").c_str(), n_factors, n, tr->fac_costs));
    // last iteration in the planner's layout: zk_sdf is d x S, Sk_sdf is d*d x S (DataRecorder.h:205-217)
    if (tr->mean) TRY(gvib200_csv_write(name("zk_sdf").c_str(), d, S, tr->mean + (size_t)(n - 1) * nmu));
    if (tr->cov_diag) TRY(gvib200_csv_write(name("Sk_sdf").c_str(), d * d, S, tr->cov_diag + (size_t)(n - 1) * nblk));
    if (nmu <= dense_limit) {
        // dense joint files from the recorded blocks (block tridiagonal)
        const int noff = (S - 1) * d * d;
        for (int which = 0; which < 2; ++which) {
            const double* src = which ? tr->prec_diag : tr->cov_diag;
            const double* off = which ? tr->prec_off : tr->cov_off;
            if (!src || (S > 1 && !off)) continue;
            std::vector<double> J((size_t)nmu * nmu * n, 0.0);
            for (int it = 0; it < n; ++it) {
                double* Jt = J.data() + (size_t)it * nmu * nmu;
                for (int s = 0; s < S; ++s)
                    for (int j = 0; j < d; ++j)
                        for (int i = 0; i < d; ++i) {
                            Jt[(size_t)(s * d + i) + (size_t)(s * d + j) * nmu] = src[(size_t)it * nblk + (size_t)s * d * d + i + j * d];
                            if (s + 1 < S) {
                                const double o = off[(size_t)it * noff + (size_t)s * d * d + i + j * d];  // block (s, s+1)
                                Jt[(size_t)(s * d + i) + (size_t)((s + 1) * d + j) * nmu] = o;
                                Jt[(size_t)((s + 1) * d + j) + (size_t)(s * d + i) * nmu] = o;
                            }
                        }
            }
            TRY(gvib200_csv_write(name(which ? "joint_precision" : "joint_cov").c_str(), nmu * nmu, n, J.data()));
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C-ABI: Prox-GVI (proxgd/ProxGVI-GH-impl.h:45-86,124-205)
// ------------------------------------------------------------------------------------------------
template <int DIM>
static void launch_prox_gh(gvib200_problem* p, const GhGroup& g, double eta, int which) {
    if constexpr (DIM > 4) {  // one warp per factor, matrices in shared memory
        static const bool thread_env = getenv("GVIB200_PROX_THREAD") != nullptr;  // development switch
        if (!thread_env) {
            constexpr int WD = 5 * DIM * DIM + 4 * ((DIM + 1) / 2) + DIM + 1;
            constexpr int GPB = 256 / JG;
            const size_t smem = (size_t)GPB * WD * sizeof(double);
            if (p->ctx->need_config(reinterpret_cast<const void*>(k_prox_gh_warp<DIM>)))
                cudaFuncSetAttribute(k_prox_gh_warp<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            LAUNCH(p, KC_OTHER, (k_prox_gh_warp<DIM>), cdiv(g.n, GPB), 256, smem, g.n, eta, g.d_raw, g.d_SR[which],
                   p->fcost[which] + g.first_id, p->fVdmu[which] + g.voff, p->fVdd[which] + g.moff);
            return;
        }
    }
    LAUNCH(p, KC_OTHER, (k_prox_gh<DIM>), cdiv(g.n, 64), 64, 0, g.n, eta, g.d_raw, g.d_SR[which],
           p->fcost[which] + g.first_id, p->fVdmu[which] + g.voff, p->fVdd[which] + g.moff);
}

// ProxGVIGH::compute_gradients(step): per-factor BW gradients + JKO step, scattered into dmu / dprecision (Vdmu, VD, VO)
static int prox_gradients(gvib200_problem* p, double eta) {
    const int c = p->cur;
    SweepTarget t{p->mu[c], p->CD[c], p->CO[c], c};
    TRY(run_sweep(p, c, false, true, true));  // full moments with the xi-space sums kept in `raw`
    for (auto& g : p->gh) {
        switch (g.dim) {
            case 1: launch_prox_gh<1>(p, g, eta, c); break;
            case 2: launch_prox_gh<2>(p, g, eta, c); break;
            case 3: launch_prox_gh<3>(p, g, eta, c); break;
            case 4: launch_prox_gh<4>(p, g, eta, c); break;
            case 6: launch_prox_gh<6>(p, g, eta, c); break;
            case 8: launch_prox_gh<8>(p, g, eta, c); break;
            case 12: launch_prox_gh<12>(p, g, eta, c); break;
            default: return fail(GVIB200_EINVAL, "prox: unsupported GH factor dim");
        }
    }
    for (auto& g : p->lin) {
        LinearArgs a = linear_args(p, g, t, true);
        double* fm = p->fVdd[c] + g.moff;
        const int sd = p->d;
#define PROX_LIN(DIM_, M_, SD_)                                                                            \
    if (g.dim == DIM_ && g.m == M_ && sd == SD_) {                                                         \
        LAUNCH(p, KC_LINEAR, (k_prox_linear<DIM_, M_, SD_>), cdiv(g.n, 64), 64, 0, a, eta, fm);            \
        continue;                                                                                          \
    }
        PROX_LIN(1, 1, 1) PROX_LIN(2, 2, 2) PROX_LIN(2, 1, 1) PROX_LIN(4, 4, 4) PROX_LIN(8, 4, 4) PROX_LIN(6, 6, 6) PROX_LIN(12, 6, 6)
#undef PROX_LIN
        return fail(GVIB200_EINVAL, "prox: unsupported linear factor shape");
    }
    run_total(p, c);
    p->sweep_valid = false;
    p->asm_valid = false;  // the NGD outputs of the sweep were overwritten
    TRY(dispatch_assemble(p, c));
    return check_launch("prox_gradients");
}

extern "C" int gvib200_prox_iterate(gvib200_problem* p, const gvib200_opts* opts_in, gvib200_iter_stats* st) {
    if (!p || !p->has_state) return fail(GVIB200_ESTATE, "prox_iterate: no state");
    if (!p->prox) return fail(GVIB200_ESTATE, "prox_iterate: the problem was not created with option \"prox\"");
    if (p->ctx->world > 1) return fail(GVIB200_EINVAL, "prox_iterate: single-GPU only");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    gvib200_opts o;
    if (opts_in) o = *opts_in;
    else gvib200_default_opts(&o);
    gvib200_iter_stats s;
    std::memset(&s, 0, sizeof(s));
    TRY(resolve_pending(p));
    if (p->iter == o.niters_lowtemp && p->is_lowtemp) {  // ProxGVI-GH-impl.h:130-133
        TRY(switch_to_high_temperature(p));
        p->is_lowtemp = false;
        s.switched_high_T = 1;
    }
    TRY(clear_flag(p));
    const int S = p->S, d = p->d;
    const size_t dd = (size_t)d * d;
    const size_t nmu = (size_t)S * d, nD = S * dd, nO = (S - 1) * dd;
    // cost_iter and the gradients at eta = base (:136-154)
    TRY(prox_gradients(p, o.step_size_base));
    s.n_moment_sweeps++;
    CUDA_TRY(cudaMemcpyAsync(p->h_scal, p->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    int flag = 0;
    TRY(read_flag(p, &flag));
    const double cost_iter = p->h_scal[2 + p->cur];
    s.cost = cost_iter;
    int cnt = 0, B = 1;
    while (true) {
        const double step = std::pow(o.step_size_base, B);  // :163
        const int c = p->cur, w = 1 - p->cur;
        LAUNCH(p, KC_CANDIDATE, k_candidate_add, cdiv(nD, 256), 256, 0, nmu, nD, nO, step, p->mu[c], p->Vdmu, p->LD[c], p->LO[c],
               p->VD, p->VO, p->mu[w], p->LD[w], p->LO[w]);
        TRY(do_selinv(p, p->LD[w], p->LO[w], p->CD[w], p->CO[w], p->scal + w));
        TRY(run_sweep(p, w, true, false, false));
        s.n_cost_sweeps++;
        run_total(p, w);
        CUDA_TRY(cudaMemcpyAsync(p->h_scal, p->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        TRY(read_flag(p, &flag));
        const double new_cost = p->h_scal[2 + w];
        s.new_cost = new_cost;
        s.step = step;
        const bool ok = (flag == 0) && (new_cost < cost_iter);
        if (ok) {
            p->cur = w;
            s.accepted = 1;
            s.n_backtrack = cnt;
            break;
        }
        B += 1;
        cnt += 1;
        if (cnt > o.max_backtrack) {  // :192-199: the last candidate is taken anyway -- unless it is not even SPD
            s.n_backtrack = cnt;
            if (flag) {
                s.status = GVIB200_ENOTSPD;
                TRY(clear_flag(p));
                if (st) *st = s;
                p->iter++;
                return fail(GVIB200_ENOTSPD, "prox_iterate: back-tracking exhausted on a candidate precision that is not SPD");
            }
            p->cur = w;
            break;
        }
        if (flag) TRY(clear_flag(p));
    }
    p->sweep_valid = false;
    p->asm_valid = false;
    p->grads_valid = false;
    p->iter++;
    if (st) *st = s;
    return 0;
}

extern "C" int gvib200_prox_optimize(gvib200_problem* p, const gvib200_opts* opts, int n_iters, gvib200_iter_stats* stats,
                                     int* n_done) {
    if (!p || !p->has_state) return fail(GVIB200_ESTATE, "prox_optimize: no state");
    int done = 0;
    for (int it = 0; it < n_iters; ++it) {
        gvib200_iter_stats s;
        const int rc = gvib200_prox_iterate(p, opts, &s);
        if (stats) stats[it] = s;
        if (rc != 0) {
            if (n_done) *n_done = done;
            return rc;
        }
        done++;
    }
    if (n_done) *n_done = done;
    return 0;
}

extern "C" int gvib200_switch_to_high_temperature(gvib200_problem* p) {
    if (!p || !p->finalized) return fail(GVIB200_ESTATE, "switch_to_high_temperature: problem not finalized");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    TRY(switch_to_high_temperature(p));
    p->is_lowtemp = false;
    p->zc_ok[0] = p->zc_ok[1] = false;
    return 0;
}

extern "C" int gvib200_reset_schedule(gvib200_problem* p) {
    if (!p) return fail(GVIB200_EINVAL, "reset_schedule: null");
    p->iter = 0;
    p->converged = false;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C-ABI: stand-alone block-tridiagonal engine
// ------------------------------------------------------------------------------------------------
static int standalone(gvib200_ctx* ctx, int S, int d, const double* diag, const double* off, const double* rhs,
                      double* x, double* cD, double* cO, double* logdet) {
    if (!ctx || S < 1 || !diag || (S > 1 && !off)) return fail(GVIB200_EINVAL, "blocktri: bad arguments");
    gvib200_problem* p = nullptr;
    TRY(gvib200_problem_create(ctx, S, d, &p));
    int rc = gvib200_problem_finalize(p);
    const size_t dd = (size_t)d * d;
    auto body = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(p->LD[0], diag, S * dd * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        if (S > 1) CUDA_TRY(cudaMemcpyAsync(p->LO[0], off, (S - 1) * dd * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        TRY(clear_flag(p));
        if (rhs) {
            CUDA_TRY(cudaMemcpyAsync(p->rhs, rhs, (size_t)S * d * sizeof(double), cudaMemcpyHostToDevice, p->stream));
            TRY(do_solve(p, p->LD[0], p->LO[0], p->rhs, p->dmu, p->scal));
            TRY(download(p, x, p->dmu, (size_t)S * d));
        } else {
            TRY(do_selinv(p, p->LD[0], p->LO[0], p->CD[0], p->CO[0], p->scal));
            TRY(download(p, cD, p->CD[0], S * dd));
            TRY(download(p, cO, p->CO[0], (S - 1) * dd));
        }
        CUDA_TRY(cudaMemcpyAsync(p->h_scal, p->scal, sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        int flag = 0;
        TRY(read_flag(p, &flag));
        if (logdet) *logdet = p->h_scal[0];
        if (flag) return fail(GVIB200_ENOTSPD, "blocktri: matrix is not positive definite");
        return 0;
    };
    if (rc == 0) rc = body();
    gvib200_problem_destroy(p);
    return rc;
}

extern "C" int gvib200_selected_inverse(gvib200_ctx* ctx, int S, int d, const double* diag, const double* off,
                                        double* cov_diag, double* cov_off, double* logdet) {
    return standalone(ctx, S, d, diag, off, nullptr, nullptr, cov_diag, cov_off, logdet);
}
extern "C" int gvib200_blocktri_solve(gvib200_ctx* ctx, int S, int d, const double* diag, const double* off,
                                      const double* rhs, double* x, double* logdet) {
    if (!rhs || !x) return fail(GVIB200_EINVAL, "blocktri_solve: null rhs/x");
    return standalone(ctx, S, d, diag, off, rhs, x, nullptr, nullptr, logdet);
}


// ------------------------------------------------------------------------------------------------
// C-ABI: device-side snapshot of the optimizer state (bench.py rewinds the trajectory between blocks of
// timed iterations so that every step does the same work)
// ------------------------------------------------------------------------------------------------
static size_t snapshot_size(const gvib200_problem* p) {
    const size_t dd = (size_t)p->d * p->d, S = (size_t)p->S;
    size_t n = S * p->d + 4 * S * dd + (size_t)p->n_factors + p->nV + p->nM + 8;
    for (auto& g : p->gh) n += (size_t)g.n * 2 * g.dim * g.dim;
    return n;
}

static int snapshot_copy(gvib200_problem* p, bool save) {
    const size_t dd = (size_t)p->d * p->d, S = (size_t)p->S;
    const int c = p->cur;
    double* q = p->snap;
    auto cp = [&](double* live, size_t n) -> int {
        CUDA_TRY(cudaMemcpyAsync(save ? q : live, save ? live : q, n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
        q += n;
        return 0;
    };
    TRY(cp(p->mu[c], S * p->d));
    TRY(cp(p->LD[c], S * dd));
    TRY(cp(p->LO[c], S * dd));
    TRY(cp(p->CD[c], S * dd));
    TRY(cp(p->CO[c], S * dd));
    for (auto& g : p->gh) TRY(cp(g.d_SR[c], (size_t)g.n * 2 * g.dim * g.dim));
    // the factor sweep at the current state (valid when sweep_valid) and the device scalars
    TRY(cp(p->fcost[c], (size_t)p->n_factors));
    TRY(cp(p->fVdmu[c], p->nV));
    TRY(cp(p->fVdd[c], p->nM));
    TRY(cp(p->scal, 8));
    return 0;
}

extern "C" int gvib200_snapshot_save(gvib200_problem* p) {
    if (!p || !p->has_state) return fail(GVIB200_ESTATE, "snapshot_save: no state");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    const size_t n = snapshot_size(p);
    if (p->snap == nullptr || p->snap_doubles != n) {
        if (p->snap) cudaFree(p->snap);
        TRY(dev_alloc(&p->snap, n));
        p->snap_doubles = n;
    }
    TRY(snapshot_copy(p, true));
    p->snap_iter = p->iter;
    p->snap_lowtemp = p->is_lowtemp;
    p->snap_cur = p->cur;
    p->snap_sweep_valid = p->sweep_valid;
    p->snap_sr_full.clear();
    for (auto& g : p->gh) p->snap_sr_full.push_back(g.sr_full[p->cur] ? 1 : 0);
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" int gvib200_snapshot_restore(gvib200_problem* p) {
    if (!p || !p->snap) return fail(GVIB200_ESTATE, "snapshot_restore: nothing saved");
    if (p->snap_lowtemp != p->is_lowtemp)
        return fail(GVIB200_ESTATE, "snapshot_restore: the temperature phase changed since the snapshot");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    if (p->snap_cur != p->cur) {
        // the scalars are slot-indexed (cur / candidate): restore into the slot layout of the snapshot
        p->cur = p->snap_cur;
    }
    TRY(snapshot_copy(p, false));
    p->iter = p->snap_iter;
    p->converged = false;
    p->batch.ldn_valid[0] = p->batch.ldn_valid[1] = false;
    p->batch.cost_valid = false;
    p->sweep_valid = p->snap_sweep_valid;
    for (size_t i = 0; i < p->gh.size() && i < p->snap_sr_full.size(); ++i) p->gh[i].sr_full[p->cur] = p->snap_sr_full[i] != 0;
    p->asm_valid = false;
    p->zc_ok[0] = p->zc_ok[1] = false;
    p->grads_valid = false;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C-ABI: timing on the problem's stream
// ------------------------------------------------------------------------------------------------
extern "C" int gvib200_timer_start(gvib200_problem* p) {
    if (!p) return fail(GVIB200_EINVAL, "timer_start: null");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    if (!p->t0) {
        CUDA_TRY(cudaEventCreate(&p->t0));
        CUDA_TRY(cudaEventCreate(&p->t1));
    }
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    CUDA_TRY(cudaEventRecord(p->t0, p->stream));
    return 0;
}

extern "C" int gvib200_timer_stop(gvib200_problem* p, float* ms) {
    if (!p || !p->t0 || !ms) return fail(GVIB200_ESTATE, "timer_stop: timer not started");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    CUDA_TRY(cudaEventRecord(p->t1, p->stream));
    CUDA_TRY(cudaEventSynchronize(p->t1));
    CUDA_TRY(cudaEventElapsedTime(ms, p->t0, p->t1));
    return 0;
}

// development aid: GVIB200_CHAIN_CLOCKS=1 -> clock64() stamps of CTA 0 of the chain kernels (kernels.cuh, cr_stamp)
static long long* g_clk_buf = nullptr;
static void chain_clocks_enable() {
    if (g_clk_buf || !getenv("GVIB200_CHAIN_CLOCKS")) return;
    cudaMalloc((void**)&g_clk_buf, 128 * sizeof(long long));
    cudaMemset(g_clk_buf, 0, 128 * sizeof(long long));
    cudaMemcpyToSymbol(g_cr_clk, &g_clk_buf, sizeof(g_clk_buf));
}
static void chain_clocks_dump() {
    if (!g_clk_buf) return;
    long long h[128];
    cudaMemcpy(h, g_clk_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[4] = {"forward<solve>", "forward<selinv>", "backward<solve>", "backward<selinv>"};
    for (int k = 0; k < 4; ++k) {
        const long long* c = h + 32 * k;
        if (c[0] == 0) continue;
        fprintf(stderr, "[chain clocks] %s: in %lld", names[k], c[1] - c[0]);
        long long prev = c[1];
        for (int l = 2; l < 20 && c[l] != 0; ++l) {
            fprintf(stderr, " L%d %lld", l - 2, c[l] - prev);
            prev = c[l];
        }
        fprintf(stderr, " out %lld total %lld\n", c[20] - prev, c[20] - c[0]);
    }
}

extern "C" int gvib200_profile_begin(gvib200_problem* p) {
    chain_clocks_enable();
    if (!p) return fail(GVIB200_EINVAL, "profile_begin: null");
    for (auto& r : p->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    p->prof.clear();
    p->profile = true;
    return 0;
}

extern "C" int gvib200_profile_end(gvib200_problem* p, gvib200_profile* out) {
    if (!p || !out) return fail(GVIB200_EINVAL, "profile_end: null");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    p->profile = false;
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    std::memset(out, 0, sizeof(*out));
    out->n_classes = KC_COUNT;
    // development aid: GVIB200_TIMELINE=<file> writes one line per launch (class, stream, start and end in ms since
    // the first launch of the profiled region)
    FILE* tl = nullptr;
    static bool tl_written = false;  // the first profiled region of the process (bench.py: the K timed steps)
    if (const char* tlp = getenv("GVIB200_TIMELINE"))
        if (!tl_written) {
            tl = fopen(tlp, "w");
            tl_written = true;
        }
    for (auto& r : p->prof) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
        if (tl) {
            float t0 = 0.f;
            cudaEventElapsedTime(&t0, p->prof.front().a, r.a);
            fprintf(tl, "%s %d %.4f %.4f\n", KC_NAMES[r.kc], r.side ? 1 : 0, t0, t0 + ms);
        }
        out->count[r.kc]++;
        out->ms[r.kc] += ms;
    }
    for (auto& r : p->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    if (tl) fclose(tl);
    chain_clocks_dump();
    p->prof.clear();
    return 0;
}

extern "C" int gvib200_problem_set_option(gvib200_problem* p, const char* name, int value) {
    if (!p || !name) return fail(GVIB200_EINVAL, "set_option: null");
    if (std::strcmp(name, "prox") == 0) {
        if (p->finalized) return fail(GVIB200_ESTATE, "set_option: \"prox\" must be set before finalize");
        p->prox = (value != 0);
        return 0;
    }
    if (std::strcmp(name, "cull") == 0) {  // free-space culling of the sign-group kernel (default on; results are identical)
        p->cull = (value != 0);
        p->sweep_valid = false;
        p->asm_valid = false;
        return 0;
    }
    if (std::strcmp(name, "generic_k1") == 0) {
        p->force_generic_k1 = (value != 0);
        p->sweep_valid = false;
        p->asm_valid = false;
        return 0;
    }
    return fail(GVIB200_EINVAL, std::string("set_option: unknown option ") + name);
}

extern "C" int gvib200_evaluated_factors(gvib200_problem* p, long long* count, int reset) {
    if (!p || !p->finalized || !count) return fail(GVIB200_EINVAL, "evaluated_factors: bad arguments");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    unsigned long long v = 0;
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    CUDA_TRY(cudaMemcpy(&v, p->d_evaluated, sizeof(v), cudaMemcpyDeviceToHost));
    if (reset) CUDA_TRY(cudaMemset(p->d_evaluated, 0, sizeof(v)));
    *count = (long long)v;
    return 0;
}

extern "C" const char* gvib200_kernel_class_name(int kc) { return (kc >= 0 && kc < KC_COUNT) ? KC_NAMES[kc] : ""; }

// algorithmic size of the problem, for the roofline arithmetic of bench.py
extern "C" int gvib200_problem_info(gvib200_problem* p, gvib200_info* out) {
    if (!p || !out) return fail(GVIB200_EINVAL, "problem_info: null");
    std::memset(out, 0, sizeof(*out));
    out->num_states = p->S;
    out->dim_state = p->d;
    out->n_factors = p->n_factors;
    for (auto& g : p->gh) {
        out->n_gh_factors += g.n;
        out->sigma_points_per_sweep += (long long)g.n * g.table->n;
    }
    for (auto& g : p->lin) out->n_linear_factors += g.n;
    out->chain_levels = p->three_level ? 3 : (p->plan.K > 0 ? 2 : 1);
    out->chain_tiles = p->plan.K;
    out->chain_tile_links = p->plan.T;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C-ABI: measurement hooks
// ------------------------------------------------------------------------------------------------
extern "C" int gvib200_time_stage(gvib200_problem* p, int stage, int reps, const gvib200_opts* opts, float* ms_per_rep,
                                  long long* kernel_launches) {
    if (!p || !p->has_state || reps < 1) return fail(GVIB200_ESTATE, "time_stage: no state");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    gvib200_opts o;
    if (opts) o = *opts;
    else gvib200_default_opts(&o);
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int c = p->cur, w = 1 - p->cur;
    // make sure gradients exist so that stages 2/3 have inputs
    if (stage == 2 || stage == 3) TRY(compute_gradients(p));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    const long long l0 = p->ctx->launches;
    CUDA_TRY(cudaEventRecord(e0, p->stream));
    for (int r = 0; r < reps; ++r) {
        switch (stage) {
            case 0: TRY(run_sweep(p, c, true, true, false)); break;
            case 1: TRY(run_sweep(p, c, true, false, false)); break;
            case 5:  // K1 alone (full moments), prologue outputs reused
            case 6:  // K1 alone (cost only)
            {
                SweepTarget t{p->mu[c], p->CD[c], p->CO[c], c};
                for (auto& g : p->gh) TRY(gh_group_dispatch(p, g, t, false, true, stage == 5, nullptr));
                break;
            }
            case 2: {
                TRY(dispatch_assemble(p, c));
                TRY(do_solve(p, p->VD, p->VO, p->rhs, p->dmu, nullptr));
                break;
            }
            case 3: {
                TRY(launch_candidate(p, o.step_size_base * o.backtrack_ratio));
                TRY(do_selinv(p, p->LD[w], p->LO[w], p->CD[w], p->CO[w], p->scal + w));
                break;
            }
            default: cudaEventDestroy(e0); cudaEventDestroy(e1); return fail(GVIB200_EINVAL, "time_stage: bad stage");
        }
    }
    CUDA_TRY(cudaEventRecord(e1, p->stream));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_per_rep) *ms_per_rep = ms / reps;
    if (kernel_launches) *kernel_launches = (p->ctx->launches - l0) / reps;
    if (stage <= 1 || stage >= 5) {
        // the sweep overwrote fcost[cur] consistently (same state), totals need refreshing
        run_total(p, c);
        p->sweep_valid = (stage == 0);
        p->asm_valid = false;
    }
    return check_launch("time_stage");
}

extern "C" int gvib200_fp64_peak(gvib200_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return fail(GVIB200_EINVAL, "fp64_peak: bad arguments");
    CUDA_TRY(cudaSetDevice(ctx->device));
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_out, sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 1 << 14, block = 256, grid = ctx->sm_count * 8;
    k_fp64_peak<<<grid, block>>>(1024, d_out);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0));
        k_fp64_peak<<<grid, block>>>(iters, d_out);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * block * (double)grid;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    ctx->launches += 6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops = best;
    return check_launch("k_fp64_peak");
}

extern "C" long long gvib200_launch_count(gvib200_ctx* ctx) { return ctx ? ctx->launches : 0; }
