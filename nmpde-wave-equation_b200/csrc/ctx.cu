// ctx.cu -- the context behind the C ABI (include/wavegpu.h): owns device memory, the stream, the
// compiled expressions and (for nranks > 1) the NCCL communicator, and sequences the kernels of
// kernels.cu into the reference's setup / init / step operations.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "../../include/wavegpu.h"
#include "expr.hpp"
#include "cg_fused.cuh"
#include "kernels.cuh"

namespace wv {
Quadrature make_quadrature(int n1d);
}

using namespace wv;

namespace {

constexpr int kSMsFallback = 148;  // B200

thread_local std::string g_create_error;
thread_local void *g_child_comm = nullptr;  // mg_setup -> wave_create: the parent's communicator for a multigrid level

// ---- NCCL, loaded on demand (single-GPU contexts never touch it) ----------------------------------
struct Nccl {
    void *lib = nullptr;
    typedef struct ncclComm *comm_t;
    struct Uid { char internal[128]; };
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(comm_t *, int, Uid, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    static constexpr int kFloat64 = 8, kSum = 0;  // ncclFloat64, ncclSum
    bool load(std::string &err) {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define WV_SYM(field, name)                                              \
    *(void **)(&field) = dlsym(lib, name);                               \
    if (!field) { err = std::string("NCCL symbol missing: ") + name; return false; }
        WV_SYM(GetUniqueId, "ncclGetUniqueId");
        WV_SYM(CommInitRank, "ncclCommInitRank");
        WV_SYM(CommDestroy, "ncclCommDestroy");
        WV_SYM(AllReduce, "ncclAllReduce");
        WV_SYM(Broadcast, "ncclBroadcast");
        WV_SYM(Send, "ncclSend");
        WV_SYM(Recv, "ncclRecv");
        WV_SYM(GroupStart, "ncclGroupStart");
        WV_SYM(GroupEnd, "ncclGroupEnd");
        WV_SYM(GetErrorString, "ncclGetErrorString");
#undef WV_SYM
        return true;
    }
};
Nccl g_nccl;

constexpr uint32_t kFlagChild = 1u << 31;  // internal: a multigrid level created by mg_setup

enum Phase { PH_RHS = 0, PH_BC, PH_CG, PH_UPDATE, PH_ENERGY, PH_OTHER, PH_COUNT };

}  // namespace

// one level of the multigrid hierarchy (WAVE_PRECOND_MG); level 0 is the solver context itself
struct MgLevel {
    wave_ctx *c = nullptr;
    const double *S = nullptr, *dinv = nullptr;
    double *x = nullptr, *x2 = nullptr, *b = nullptr, *r = nullptr;
    double omega = 0.8;
};
struct Mg {
    int nlev = 0, nu = 2, nu_coarse = 8;
    MgLevel lev[16];
    double s = 0.0;
};

struct wave_ctx {
    wave_config cfg{};
    Layout L{};
    std::string err;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    long long launches = 0;
    cudaError_t launch_error = cudaSuccess;
    Launcher launcher{};
    bool is_setup = false, is_init = false;

    // expressions
    Program hprog[WAVE_EXPR_COUNT]{};
    bool has[WAVE_EXPR_COUNT]{};
    Program *dprog = nullptr;  // WAVE_EXPR_COUNT programs
    bool forcing_active = false;
    // F(x,y,t) = T(t) S(x,y): the load vector of S lives in fvec (assembled once), T is evaluated on the host
    bool forcing_separable = false;
    Program f_time{}, f_space{};
    std::string f_text[3];  // expression, variables, constants as given to wave_set_expr
    Quadrature q_asm{}, q_err{};

    // SELL-32-sigma pattern shared by all matrices (owned rows, local column indices) + CSR row pointer
    uint32_t *rowptr = nullptr;
    uint32_t *slice_ptr = nullptr;
    int32_t *col = nullptr, *row_of = nullptr, *slot_of = nullptr;
    int32_t *c2i = nullptr, *i2c = nullptr;  // canonical <-> storage numbering of the local range (P2)
    bool permuted = false;
    double *tmp = nullptr;                   // nloc staging for the permutation at the ABI
    int nslices = 0;
    int64_t nnz = 0;      // entries of the pattern (CSR)
    int64_t nnz_pad = 0;  // stored entries including padding
    Sell A{};
    double *M = nullptr, *K = nullptr, *S1 = nullptr, *S2 = nullptr;
    // stencil operator (kernels.cuh): tables of M, K, SYS1, SYS2 (4 x kStencilKinds x kStencilMax), metadata,
    // per-slice marks; stencil_rows = owned rows served by the tables, sell_nnz = entries of the other rows
    double *st_tab = nullptr;
    int32_t *st_meta = nullptr;
    int2 *slice_info = nullptr;
    int32_t *sell_list = nullptr;
    int2 *st_walk = nullptr;
    std::vector<int32_t> h_st_meta;  // host copies: the launches pass them as kernel parameters
    std::vector<double> h_tab;
    int64_t stencil_rows = 0, sell_nnz = 0;
    double *dinv1 = nullptr, *dinv2 = nullptr;
    double *d0 = nullptr;  // [2]

    // vectors: u, v, a, unew, d are local-layout (ghosts included); rhs, fvec, g, h are row-indexed
    double *u = nullptr, *v = nullptr, *a = nullptr, *unew = nullptr, *d = nullptr;
    double *rhs = nullptr, *fvec = nullptr, *g = nullptr, *h = nullptr;
    double *cellvec = nullptr;  // per-cell load vectors of the forcing kernel (non-separable forcing only)
    double *scratch = nullptr;  // global-size staging for gathers (allocated on demand)
    int64_t scratch_n = 0;

    // boundary
    int nb = 0;
    int64_t nb_global = 0;
    int32_t *brow = nullptr;
    double *bx = nullptr, *by = nullptr;
    std::vector<int32_t> h_bdof_global;  // all boundary DoFs (global ids, sorted)

    // reductions / CG state
    double *partials = nullptr, *partials2 = nullptr, *partials3 = nullptr;  // SpMV / update / other sums
    unsigned *counter = nullptr;
    CgScalars *S = nullptr;
    CgScalars *hS = nullptr;  // pinned
    double *res = nullptr;    // device scalars [8]
    double *hres = nullptr;   // pinned [8]
    int prev_its[2] = {0, 0};

    // multigrid preconditioner
    Mg *mg = nullptr;
    double *mg_buf[3] = {nullptr, nullptr, nullptr};  // level-0 work vectors

    // K6f, opt-in (WAVE_CG_FUSED=1): the whole Jacobi-PCG solve as one cooperative kernel (cg_fused.cuh)
    struct FusedPlan {
        bool ok = false;
        int grid = 0, wpb = 0, stage_cap = 0;
        size_t smem = 0;
        int32_t *blk_c0 = nullptr, *blk_cn = nullptr;
        double *partials = nullptr;
        unsigned long long *pub = nullptr;  // several ranks: block 0 publishes the all-rank totals here
    } fused;

    // NCCL + NVLink peer exchange
    Nccl::comm_t comm = nullptr;
    bool own_comm = true;                // multigrid levels borrow the solver context's communicator
    PeerComm pc{};                       // enabled only for 1 < nranks <= kMaxPeers
    PeerMailbox *mailbox = nullptr;
    void *ipc_opened[2 * kMaxPeers]{};   // mapped peer allocations (closed in wave_destroy)
    int n_ipc_opened = 0;
    unsigned long long ar_seq = 0, halo_seq = 0;

    // instrumentation
    bool timers_on = false;
    double phase_ms[PH_COUNT]{};
    cudaEvent_t ev[2 * PH_COUNT + 4]{};
    double cg_stats[4]{};
    // live kernel timing inside the CG solves: slot 0 = SpMV (A d), 1 = k_cg_update, 2 = k_cg_direction
    bool spmv_timing = false;
    std::vector<cudaEvent_t> spmv_ev;  // pairs
    std::vector<int> spmv_ev_tag;      // per pair: slot | iteration << 2
    size_t spmv_ev_used = 0;
    double kt_ms[3] = {0, 0, 0}, kt_count[3] = {0, 0, 0};
    double *flush_buf = nullptr;
    int64_t flush_n = 0;
};

namespace {

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            return WAVE_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)
#define NK(call)                                                                              \
    do {                                                                                      \
        int e_ = (call);                                                                      \
        if (e_ != 0) {                                                                        \
            ctx->err = std::string(#call) + ": " + g_nccl.GetErrorString(e_);                 \
            return WAVE_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)
#define RET(call)                                \
    do {                                         \
        int rc_ = (call);                        \
        if (rc_ != WAVE_OK) return rc_;          \
    } while (0)

int fail(wave_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return code;
}

void quad_row_split(int ny, int rank, int nranks, int &j0, int &j1) {
    j0 = (int)((int64_t)ny * rank / nranks);
    j1 = (int)((int64_t)ny * (rank + 1) / nranks);
}

Layout make_layout(const Mesh &m, int rank, int nranks) {
    Layout L{};
    L.mesh = m;
    quad_row_split(m.ny, rank, nranks, L.jq0, L.jq1);
    L.row0 = block_start(m, L.jq0);
    const int64_t row1 = L.jq1 >= m.ny ? n_dofs(m) : block_start(m, L.jq1);
    L.nown = (int)(row1 - L.row0);
    L.col0 = block_start(m, L.jq0 > 0 ? L.jq0 - 1 : 0);
    const int jhi = L.jq1 + 1;
    const int64_t col1 = jhi >= m.ny ? n_dofs(m) : block_start(m, jhi);
    L.nloc = (int)(col1 - L.col0);
    L.own_off = (int)(L.row0 - L.col0);
    return L;
}

int launch_check(wave_ctx *ctx) {
    if (ctx->launch_error != cudaSuccess) {
        ctx->err = std::string("kernel launch failed: ") + cudaGetErrorString(ctx->launch_error);
        ctx->launch_error = cudaSuccess;
        return WAVE_ERR_CUDA;
    }
    return WAVE_OK;
}
int sync_check(wave_ctx *ctx) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return launch_check(ctx);
}

// ---- multi-GPU plumbing: contiguous halo blocks and small all-reduces over NCCL -------------------
int halo_exchange(wave_ctx *ctx, double *vec) {
    if (ctx->cfg.nranks == 1) return WAVE_OK;
    const Layout &L = ctx->L;
    const int rank = ctx->cfg.rank, nr = ctx->cfg.nranks;
    const Mesh &m = L.mesh;
    NK(g_nccl.GroupStart());
    if (rank > 0) {
        // lower neighbour owns block jq0-1 (my lower ghost); it needs my first block jq0
        const size_t ghost = (size_t)L.own_off;
        const size_t mine = (size_t)(block_start(m, L.jq0 + 1) - block_start(m, L.jq0));
        NK(g_nccl.Recv(vec, ghost, Nccl::kFloat64, rank - 1, ctx->comm, ctx->stream));
        NK(g_nccl.Send(vec + L.own_off, mine, Nccl::kFloat64, rank - 1, ctx->comm, ctx->stream));
    }
    if (rank < nr - 1) {
        // upper neighbour owns block jq1 (my upper ghost); it needs my last block jq1-1
        const size_t ghost = (size_t)(L.nloc - L.own_off - L.nown);
        const size_t mine = (size_t)(block_start(m, L.jq1) - block_start(m, L.jq1 - 1));
        NK(g_nccl.Recv(vec + L.own_off + L.nown, ghost, Nccl::kFloat64, rank + 1, ctx->comm, ctx->stream));
        NK(g_nccl.Send(vec + L.own_off + L.nown - mine, mine, Nccl::kFloat64, rank + 1, ctx->comm, ctx->stream));
    }
    NK(g_nccl.GroupEnd());
    return WAVE_OK;
}
int allreduce(wave_ctx *ctx, double *dev, size_t count) {
    if (ctx->cfg.nranks == 1) return WAVE_OK;
    NK(g_nccl.AllReduce(dev, dev, count, Nccl::kFloat64, Nccl::kSum, ctx->comm, ctx->stream));
    return WAVE_OK;
}

// ---- phase timers -----------------------------------------------------------------------------------
struct PhaseTimer {
    wave_ctx *ctx;
    int ph;
    PhaseTimer(wave_ctx *c, int p) : ctx(c), ph(p) {
        if (ctx->timers_on) cudaEventRecord(ctx->ev[2 * ph], ctx->stream);
    }
    ~PhaseTimer() {
        if (ctx->timers_on) {
            cudaEventRecord(ctx->ev[2 * ph + 1], ctx->stream);
            cudaEventSynchronize(ctx->ev[2 * ph + 1]);
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->ev[2 * ph], ctx->ev[2 * ph + 1]);
            ctx->phase_ms[ph] += ms;
        }
    }
};

// resolve the recorded event pairs into per-slot (count, ms).  Only launches of iterations below
// `valid_iterations` are counted: launches enqueued after the solve converged return at once and must not
// dilute the averages.
void drain_spmv_events(wave_ctx *ctx, int valid_iterations = 1 << 29) {
    if (!ctx->spmv_ev_used) return;
    cudaEventSynchronize(ctx->spmv_ev[ctx->spmv_ev_used - 1]);
    for (size_t k = 0; k + 1 < ctx->spmv_ev_used; k += 2) {
        const int tag = ctx->spmv_ev_tag[k / 2];
        if ((tag >> 2) >= valid_iterations) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->spmv_ev[k], ctx->spmv_ev[k + 1]) == cudaSuccess) {
            ctx->kt_ms[tag & 3] += ms;
            ctx->kt_count[tag & 3] += 1;
        }
    }
    ctx->spmv_ev_used = 0;
}
struct SpmvBracket {
    wave_ctx *ctx;
    bool on;
    SpmvBracket(wave_ctx *c, int slot, int iteration) : ctx(c), on(c->spmv_timing) {
        if (!on) return;
        if (ctx->spmv_ev_used + 2 > ctx->spmv_ev.size()) { on = false; return; }  // full: skip, never drain mid-solve
        ctx->spmv_ev_tag[ctx->spmv_ev_used / 2] = slot | (iteration << 2);
        cudaEventRecord(ctx->spmv_ev[ctx->spmv_ev_used], ctx->stream);
    }
    ~SpmvBracket() {
        if (!on) return;
        cudaEventRecord(ctx->spmv_ev[ctx->spmv_ev_used + 1], ctx->stream);
        ctx->spmv_ev_used += 2;
    }
};

SpmvArgs spmv_base(wave_ctx *ctx) {
    SpmvArgs a{};
    a.A = ctx->A;
    a.partials = ctx->partials;
    a.counter = ctx->counter;
    a.pc = PeerComm{};  // peer exchange only inside the CG loop (cg_solve sets it)
    return a;
}

constexpr int kTabSize = kStencilKinds * kStencilMax;
// which stencil table belongs to a value array of `ctx` (-1: none), and its host copy
int tab_index(const wave_ctx *ctx, const double *val) {
    if (!ctx->st_tab || !val) return -1;
    if (val == ctx->M) return 0;
    if (val == ctx->K) return 1;
    if (val == ctx->S1) return 2;
    if (val == ctx->S2) return 3;
    return -1;
}
const double *tab_of(const wave_ctx *ctx, const double *val) {
    const int k = tab_index(ctx, val);
    return k < 0 || ctx->h_tab.empty() ? nullptr : ctx->h_tab.data() + (size_t)k * kTabSize;
}
// every SpMV of the library goes through here: `owner` is the context whose pattern and values are used
void spmv(const wave_ctx *owner, const Launcher &l, SpmvArgs &a) {
    for (auto &t : a.t) t.tab = tab_of(owner, t.val);
    launch_spmv(l, a);
}

// SolverCG::solve (src/WaveNewmark.cpp:256-261): Jacobi-PCG on the BC-modified matrix `Sval`,
// start vector x (local layout), right-hand side b (row-indexed).
// V-cycle on level l: lev.x <- approximate solution of S x = lev.b from x = 0 (all launches asynchronous)
// x, x2, r of a level are local-layout vectors (ghost blocks included), b is row-indexed.  With several
// ranks every SpMV is preceded by the halo exchange of its input (NCCL send/recv of the two ghost blocks).
int mg_smooth(wave_ctx *ctx, MgLevel &lv, int sweeps, bool x_is_zero, const int *skip) {
    wave_ctx *c = lv.c;
    const int off = c->L.own_off;
    for (int sw = 0; sw < sweeps; ++sw) {
        if (sw == 0 && x_is_zero) {
            launch_scale_rows(ctx->launcher, c->L.nown, lv.omega, lv.dinv, lv.b, lv.x + off, skip);
            continue;
        }
        RET(halo_exchange(c, lv.x));
        SpmvArgs a = spmv_base(c);
        a.partials = ctx->partials;
        a.counter = ctx->counter;
        a.t[0] = {lv.S, lv.x, nullptr, 1.0, 0.0, -1.0};
        a.add0 = lv.b; a.addc0 = 1.0;
        a.dinv = lv.dinv; a.jac_x = lv.x + off; a.jac_omega = lv.omega;
        a.y = lv.x2 + off;
        a.skip_flag = skip;
        spmv(c, ctx->launcher, a);
        std::swap(lv.x, lv.x2);
    }
    return WAVE_OK;
}
int mg_vcycle(wave_ctx *ctx, int l, const int *skip) {
    Mg &m = *ctx->mg;
    MgLevel &lv = m.lev[l];
    if (l == m.nlev - 1) return mg_smooth(ctx, lv, m.nu_coarse, true, skip);
    RET(mg_smooth(ctx, lv, m.nu, true, skip));
    {   // r = b - S x
        RET(halo_exchange(lv.c, lv.x));
        SpmvArgs a = spmv_base(lv.c);
        a.partials = ctx->partials;
        a.counter = ctx->counter;
        a.t[0] = {lv.S, lv.x, nullptr, 1.0, 0.0, -1.0};
        a.add0 = lv.b; a.addc0 = 1.0;
        a.y = lv.r + lv.c->L.own_off;
        a.skip_flag = skip;
        spmv(lv.c, ctx->launcher, a);
    }
    MgLevel &cv = m.lev[l + 1];
    const bool p_coarsening = lv.c->L.mesh.r == 2;
    RET(halo_exchange(lv.c, lv.r));  // restriction reads the fine residual of the upper ghost block
    if (p_coarsening) launch_restrict_p2p1(ctx->launcher, lv.c->L, cv.c->L, lv.r, cv.b, skip);
    else launch_restrict_p1(ctx->launcher, lv.c->L, cv.c->L, lv.r, cv.b, skip);
    RET(mg_vcycle(ctx, l + 1, skip));
    RET(halo_exchange(cv.c, cv.x));  // prolongation reads the coarse correction of the lower ghost block
    if (p_coarsening) launch_prolong_add_p2p1(ctx->launcher, lv.c->L, cv.c->L, cv.x, lv.x, skip);
    else launch_prolong_add_p1(ctx->launcher, lv.c->L, cv.c->L, cv.x, lv.x, skip);
    return mg_smooth(ctx, lv, m.nu, false, skip);
}
// z = V-cycle(g) on the fine level; *z = the (row-indexed) vector holding it
int mg_apply(wave_ctx *ctx, const double *Sval, const double *dinv, const double *g, const int *skip,
             const double **z) {
    MgLevel &f = ctx->mg->lev[0];
    f.S = Sval;
    f.dinv = dinv;
    f.b = const_cast<double *>(g);
    RET(mg_vcycle(ctx, 0, skip));
    *z = f.x + ctx->L.own_off;
    return WAVE_OK;
}

int cg_solve(wave_ctx *ctx, const double *Sval, const double *dinv, double *x, const double *b, int slot,
             int *iters, bool use_mg = false) {
    const Layout &L = ctx->L;
    const Launcher &l = ctx->launcher;
    use_mg = use_mg && ctx->mg != nullptr;
    cudaEvent_t e0 = ctx->ev[2 * PH_COUNT], e1 = ctx->ev[2 * PH_COUNT + 1];
    CK(cudaEventRecord(e0, ctx->stream));
    // the status of the previous solve must not gate the kernels of this one (V-cycle before k_cg_start)
    CK(cudaMemsetAsync(ctx->S->status, 0, 2 * sizeof(int), ctx->stream));
    RET(halo_exchange(ctx, x));
    {   // g = A x - b ; h = D^-1 g ; d = -h ; gg, gh
        SpmvArgs a = spmv_base(ctx);
        a.t[0] = {Sval, x, nullptr, 1.0, 0.0, 1.0};
        a.add0 = b; a.addc0 = -1.0;
        a.y = ctx->g;
        if (!use_mg) { a.dinv = dinv; a.h_out = ctx->h; a.d_out = ctx->d + L.own_off; }
        a.dot_mode = 2;
        a.result = &ctx->S->gg;
        spmv(ctx, l, a);
    }
    if (use_mg) {  // h = V-cycle(g) ; d = -h ; gh = g.h
        const double *z = nullptr;
        RET(mg_apply(ctx, Sval, dinv, ctx->g, &ctx->S->status[0], &z));
        launch_dot_gz(l, L.nown, ctx->g, z, ctx->d + L.own_off, ctx->partials, ctx->counter, &ctx->S->gh_new, nullptr);
    }
    RET(allreduce(ctx, &ctx->S->gg, 2));
    launch_cg_start(l, ctx->S);
    const bool fused = ctx->fused.ok && !use_mg;
    if (fused) {  // K6f: every iteration inside one cooperative kernel; the loop below only reads the outcome
        CgFusedArgs fa{};
        fa.A = ctx->A;
        fa.val = Sval;
        fa.dinv = dinv;
        fa.x_own = x + L.own_off;
        fa.d = ctx->d;
        fa.own_off = L.own_off;
        fa.g = ctx->g;
        fa.S = ctx->S;
        fa.partials = ctx->fused.partials;
        fa.nwin = ctx->nslices / (kWindow / kSlice);
        fa.wpb = ctx->fused.wpb;
        fa.blk_c0 = ctx->fused.blk_c0;
        fa.blk_cn = ctx->fused.blk_cn;
        fa.stage_cap = ctx->fused.stage_cap;
        if (ctx->pc.enabled) {  // the first iteration's halo of d comes over NCCL, later ones inside the kernel
            RET(halo_exchange(ctx, ctx->d));
            fa.pc = ctx->pc;
            fa.pub = ctx->fused.pub;
            fa.ar_seq0 = ctx->ar_seq;
            fa.halo_seq0 = ctx->halo_seq;
        }
        CK(launch_cg_fused(l, ctx->fused.grid, ctx->fused.smem, fa));
    }
    int enq = 0;
    // iteration counts of consecutive time steps are nearly equal (warm start): enqueue as many
    // iterations as the previous solve needed, then poll in pairs
    int chunk = ctx->prev_its[slot] > 0 ? ctx->prev_its[slot] : 4;
    const int maxit = ctx->hS->maxit;
    if (fused) chunk = 0;
    // where the sums of an iteration travel: per-block partials summed by the consumer kernel on one rank,
    // the NVLink mailboxes for 2..8 ranks, NCCL all-reduces of the scalars otherwise
    const bool p2p = ctx->pc.enabled != 0;
    const int sum_mode = p2p ? SUM_MAILBOX : (ctx->cfg.nranks > 1 ? SUM_SCALAR : SUM_PARTIALS);
    const int nvec_blocks = cg_vector_blocks(L.nown);
    double *p1 = ctx->partials, *p2 = ctx->partials2;
    for (;;) {
        for (int k = 0; k < chunk; ++k) {
            const int it = enq + k, parity = it & 1;
            const bool first_it = it == 0;
            // halo of d: NCCL for the first iteration (d comes from the residual kernel), afterwards the
            // neighbours' k_cg_direction wrote it into the ghost blocks and raised halo flag `halo_seq`
            if (!p2p || first_it) RET(halo_exchange(ctx, ctx->d));
            SpmvArgs a = spmv_base(ctx);
            a.t[0] = {Sval, ctx->d, nullptr, 1.0, 0.0, 1.0};
            a.y = ctx->h;
            a.dot_mode = 1;
            a.dotv = ctx->d + L.own_off;
            a.skip_flag = &ctx->S->status[parity];
            a.partials = p1;
            if (sum_mode == SUM_SCALAR) a.result = &ctx->S->dAd;
            else a.dot_publish = 1;
            if (p2p) {
                a.pc = ctx->pc;
                a.ar_seq = ++ctx->ar_seq;
                a.halo_wait_seq = first_it ? 0ull : ctx->halo_seq;
                a.ghost_lo_slices = ctx->cfg.rank > 0 ? ((ctx->pc.lo_count + kWindow - 1) / kWindow) * (kWindow / kSlice) : 0;
                a.ghost_hi_slice0 = ctx->cfg.rank < ctx->cfg.nranks - 1
                                        ? ((L.nown - ctx->pc.hi_count) / kWindow) * (kWindow / kSlice)
                                        : ctx->nslices;
            }
            for (auto &t : a.t) t.tab = tab_of(ctx, t.val);
            const CgSumIo in1{sum_mode, p1, spmv_launch_blocks(a), 2, a.ar_seq};
            {
                SpmvBracket br(ctx, 0, it);
                launch_spmv(l, a);
            }
            if (sum_mode == SUM_SCALAR) RET(allreduce(ctx, &ctx->S->dAd, 1));
            const unsigned long long seq_b = p2p ? ++ctx->ar_seq : 0ull;
            {
                SpmvBracket br(ctx, 1, it);
                launch_cg_update(l, L.nown, parity, ctx->S, ctx->g, ctx->h, use_mg ? nullptr : dinv, in1, p2,
                                 ctx->counter, p2p ? ctx->pc : PeerComm{}, seq_b, sum_mode);
            }
            const double *z = ctx->h;
            if (use_mg) {  // h' = V-cycle(g) ; gh' = g.h' (all-reduced over NCCL: a handful of iterations per solve)
                RET(mg_apply(ctx, Sval, dinv, ctx->g, &ctx->S->status[parity], &z));
                launch_dot_gz(l, L.nown, ctx->g, z, nullptr, ctx->partials3, ctx->counter, &ctx->S->gh_new,
                              &ctx->S->status[parity]);
                if (sum_mode != SUM_SCALAR) RET(allreduce(ctx, &ctx->S->gh_new, 1));
            }
            if (sum_mode == SUM_SCALAR) RET(allreduce(ctx, &ctx->S->gg, 2));
            const unsigned long long seq_h = p2p ? ++ctx->halo_seq : 0ull;
            const CgSumIo in2{sum_mode, p2, nvec_blocks, 2, seq_b};
            {
                SpmvBracket br(ctx, 2, it);
                launch_cg_direction(l, L.nown, it, ctx->S, x + L.own_off, ctx->d + L.own_off, z, in2, use_mg ? 1 : 0,
                                    ctx->counter, p2p ? ctx->pc : PeerComm{}, seq_h);
            }
        }
        enq += chunk;
        CK(cudaMemcpyAsync(ctx->hS, ctx->S, sizeof(CgScalars), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        RET(launch_check(ctx));
        if (ctx->hS->status[enq & 1] != 0 || fused) break;
        if (enq > maxit + 8) break;
        chunk = use_mg ? 1 : 2;  // a skipped multigrid iteration still costs ~50 no-op launches
    }
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const int status = ctx->hS->status[enq & 1];
    *iters = ctx->hS->it;
    if (fused && ctx->pc.enabled) {  // the sequence numbers the kernel consumed (identical on every rank)
        ctx->ar_seq += 2ull * (unsigned long long)ctx->hS->it;
        ctx->halo_seq += (unsigned long long)ctx->hS->it;
    }
    if (ctx->spmv_timing) drain_spmv_events(ctx, ctx->hS->it);
    ctx->prev_its[slot] = ctx->hS->it;
    ctx->cg_stats[0] += 1;
    ctx->cg_stats[1] += ctx->hS->it;
    ctx->cg_stats[2] += ctx->hS->it + 1;
    ctx->cg_stats[3] += ms;
    if (ctx->hS->peer_timeout) {
        CK(cudaMemsetAsync(&ctx->S->peer_timeout, 0, sizeof(int), ctx->stream));
        return fail(ctx, WAVE_ERR_CUDA, "peer exchange timed out: a neighbouring rank did not arrive (NVLink mailbox)");
    }
    if (status != 1)
        return fail(ctx, WAVE_ERR_NOCONV, "CG did not converge within the iteration limit (SolverControl::NoConvergence)");
    return WAVE_OK;
}

// scheme matrix: out = bc(M + s K); also d0 and the Jacobi diagonal
int build_system_matrix(wave_ctx *ctx, double s, double *out, double *dinv, double *d0) {
    const Launcher &l = ctx->launcher;
    launch_axpy_vals(l, ctx->nnz_pad, ctx->M, ctx->K, s, out);
    if (ctx->st_tab && tab_index(ctx, out) >= 0) {
        // the same expression on the representative rows (Dirichlet rows are never stencil rows), then the
        // host copy the launches pass as kernel parameters
        const int k = tab_index(ctx, out);
        launch_axpy_vals(l, kTabSize, ctx->st_tab, ctx->st_tab + kTabSize, s, ctx->st_tab + (size_t)k * kTabSize);
        CK(cudaMemcpyAsync(ctx->h_tab.data() + (size_t)k * kTabSize, ctx->st_tab + (size_t)k * kTabSize,
                           sizeof(double) * kTabSize, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    launch_find_d0(l, ctx->L, ctx->A, out, d0);
    launch_bc_rows(l, ctx->L, ctx->nb, ctx->brow, ctx->A, out, d0);
    launch_dinv(l, ctx->L, ctx->A, out, ctx->cfg.precond == WAVE_PRECOND_NONE, dinv);
    return WAVE_OK;
}

int upload_cg_control(wave_ctx *ctx) {
    CgScalars s{};
    s.tol = ctx->cfg.cg_tol;
    s.reduce = ctx->cfg.cg_reduce;
    s.maxit = ctx->cfg.cg_maxit;
    *ctx->hS = s;
    CK(cudaMemcpyAsync(ctx->S, ctx->hS, sizeof(CgScalars), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return WAVE_OK;
}

// Load vector of w_np1 F(t_np1) + w_n F(t_n) as `scale * fvec`.  Separable forcing: fvec holds the load
// vector of S and only the scalar T changes; otherwise the per-cell quadrature kernel runs.
int compute_forcing(wave_ctx *ctx, double t_np1, double t_n, double w_np1, double w_n, int two_levels,
                    double *scale) {
    *scale = 1.0;
    if (ctx->forcing_separable) {
        const double T1 = eval(&ctx->f_time, 0.0, 0.0, t_np1);
        *scale = two_levels ? w_np1 * T1 + w_n * eval(&ctx->f_time, 0.0, 0.0, t_n) : T1;
        return WAVE_OK;
    }
    launch_forcing(ctx->launcher, ctx->L, ctx->dprog + WAVE_EXPR_F, &ctx->q_asm, t_np1, t_n, w_np1, w_n,
                   two_levels, ctx->cellvec, ctx->fvec);
    return WAVE_OK;
}

int finish_norms(wave_ctx *ctx, double norms[2]) {
    RET(allreduce(ctx, ctx->res, 2));
    CK(cudaMemcpyAsync(ctx->hres, ctx->res, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (norms) { norms[0] = std::sqrt(ctx->hres[0]); norms[1] = std::sqrt(ctx->hres[1]); }
    return WAVE_OK;
}

// WaveNewmark: assemble_rhs + solve_a + update_u_v (src/WaveNewmark.cpp:116-278)
int newmark_step(wave_ctx *ctx, double t, int32_t iters[2], double norms[2]) {
    const Layout &L = ctx->L;
    const Launcher &l = ctx->launcher;
    const double dt = ctx->cfg.dt, beta = ctx->cfg.beta, gamma = ctx->cfg.gamma;
    int its = 0;
    {
        PhaseTimer pt(ctx, PH_RHS);
        // u <- z = u + dt v + dt^2(1/2-beta) a ; v <- v + dt(1-gamma) a
        launch_newmark_predict(l, L.nown, dt, dt * dt * (0.5 - beta), dt * (1.0 - gamma), ctx->u + L.own_off,
                               ctx->v + L.own_off, ctx->a + L.own_off);
        RET(halo_exchange(ctx, ctx->u));
        double fscale = 1.0;
        if (ctx->forcing_active) RET(compute_forcing(ctx, t, 0.0, 1.0, 0.0, 0, &fscale));
        SpmvArgs a = spmv_base(ctx);
        a.t[0] = {ctx->K, ctx->u, nullptr, 1.0, 0.0, -1.0};
        if (ctx->forcing_active) { a.add0 = ctx->fvec; a.addc0 = fscale; }
        a.y = ctx->rhs;
        spmv(ctx, l, a);
    }
    {
        PhaseTimer pt(ctx, PH_BC);
        const int mode = beta > 1e-12 ? BC_NEWMARK_IMPLICIT : BC_SECOND_DIFF;
        launch_bc_values(l, mode, ctx->nb, ctx->brow, ctx->bx, ctx->by, ctx->dprog + WAVE_EXPR_G, t, dt,
                         beta * dt * dt, ctx->u + L.own_off, ctx->a + L.own_off, ctx->rhs, ctx->d0);
    }
    {
        PhaseTimer pt(ctx, PH_CG);
        RET(cg_solve(ctx, ctx->S1, ctx->dinv1, ctx->a, ctx->rhs, 0, &its, true));
    }
    {
        PhaseTimer pt(ctx, PH_UPDATE);
        launch_newmark_correct(l, L.nown, dt * dt * beta, dt * gamma, ctx->u + L.own_off, ctx->v + L.own_off,
                               ctx->a + L.own_off, ctx->partials, ctx->counter, ctx->res);
        RET(finish_norms(ctx, norms));
    }
    if (iters) { iters[0] = its; iters[1] = 0; }
    return WAVE_OK;
}

// WaveTheta: assemble_rhs_u + solve_u + assemble_rhs_v + solve_v (src/WaveTheta.cpp:119-339)
int theta_step(wave_ctx *ctx, double t, int32_t iters[2], double norms[2]) {
    const Layout &L = ctx->L;
    const Launcher &l = ctx->launcher;
    const double dt = ctx->cfg.dt, th = ctx->cfg.theta;
    int its_u = 0, its_v = 0;
    double fscale = 1.0;
    {
        PhaseTimer pt(ctx, PH_RHS);
        RET(halo_exchange(ctx, ctx->u));
        RET(halo_exchange(ctx, ctx->v));
        if (ctx->forcing_active) RET(compute_forcing(ctx, t, t - dt, th, 1.0 - th, 1, &fscale));
        // rhs = M u + dt M v - dt^2 theta (1-theta) K u + theta dt^2 F_theta
        SpmvArgs a = spmv_base(ctx);
        a.t[0] = {ctx->M, ctx->u, ctx->v, 1.0, dt, 1.0};
        a.t[1] = {ctx->K, ctx->u, nullptr, 1.0, 0.0, -dt * dt * th * (1 - th)};
        if (ctx->forcing_active) { a.add0 = ctx->fvec; a.addc0 = th * dt * dt * fscale; }
        a.y = ctx->rhs;
        spmv(ctx, l, a);
        launch_copy(l, L.nloc, ctx->u, ctx->unew);
    }
    {
        PhaseTimer pt(ctx, PH_BC);
        launch_bc_values(l, BC_DIRECT, ctx->nb, ctx->brow, ctx->bx, ctx->by, ctx->dprog + WAVE_EXPR_G, t, dt, 0.0,
                         nullptr, ctx->unew + L.own_off, ctx->rhs, ctx->d0);
    }
    {
        PhaseTimer pt(ctx, PH_CG);
        RET(cg_solve(ctx, ctx->S1, ctx->dinv1, ctx->unew, ctx->rhs, 0, &its_u, true));
    }
    {
        PhaseTimer pt(ctx, PH_RHS);
        RET(halo_exchange(ctx, ctx->unew));
        // rhs = M v - dt (1-theta) K u^n - dt theta K u^{n+1} + dt F_theta
        SpmvArgs a = spmv_base(ctx);
        a.t[0] = {ctx->M, ctx->v, nullptr, 1.0, 0.0, 1.0};
        a.t[1] = {ctx->K, ctx->u, ctx->unew, 1.0 - th, th, -dt};
        if (ctx->forcing_active) { a.add0 = ctx->fvec; a.addc0 = dt * fscale; }
        a.y = ctx->rhs;
        spmv(ctx, l, a);
    }
    {
        PhaseTimer pt(ctx, PH_BC);
        launch_bc_values(l, BC_DIRECT, ctx->nb, ctx->brow, ctx->bx, ctx->by, ctx->dprog + WAVE_EXPR_DGDT, t, dt, 0.0,
                         nullptr, ctx->v + L.own_off, ctx->rhs, ctx->d0 + 1);
    }
    {
        PhaseTimer pt(ctx, PH_CG);
        RET(cg_solve(ctx, ctx->S2, ctx->dinv2, ctx->v, ctx->rhs, 1, &its_v));
    }
    {
        PhaseTimer pt(ctx, PH_UPDATE);
        std::swap(ctx->u, ctx->unew);
        launch_norms2(l, L.nown, ctx->u + L.own_off, ctx->v + L.own_off, ctx->partials, ctx->counter, ctx->res);
        RET(finish_norms(ctx, norms));
    }
    if (iters) { iters[0] = its_u; iters[1] = its_v; }
    return WAVE_OK;
}

double *vec_ptr(wave_ctx *ctx, int which) {
    switch (which) {
    case WAVE_VEC_U: return ctx->u;
    case WAVE_VEC_V: return ctx->v;
    case WAVE_VEC_A: return ctx->a;
    default: return nullptr;
    }
}
const double *mat_ptr(wave_ctx *ctx, int which) {
    switch (which) {
    case WAVE_MAT_M: return ctx->M;
    case WAVE_MAT_K: return ctx->K;
    case WAVE_MAT_SYS1: return ctx->S1;
    case WAVE_MAT_SYS2: return ctx->S2;
    default: return nullptr;
    }
}

// algorithmic bytes of one SpMV launch: 12 B per entry + row map, x and y (20 B per row) for the rows kept in
// SELL form; x and y only (16 B per row) for the rows served by the stencil tables
double spmv_bytes(const wave_ctx *ctx) {
    const double sell_rows = (double)(ctx->L.nown - ctx->stencil_rows);
    return 12.0 * (double)ctx->sell_nnz + 20.0 * sell_rows + 16.0 * (double)ctx->stencil_rows;
}

int ensure_scratch(wave_ctx *ctx, int64_t n) {
    if (ctx->scratch_n >= n) return WAVE_OK;
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_n = 0;
    CK(cudaMalloc(&ctx->scratch, sizeof(double) * n));
    ctx->scratch_n = n;
    return WAVE_OK;
}

// canonical local-range vector (device) -> storage order
int to_storage(wave_ctx *ctx, const double *canon_local, double *dst_local) {
    const Layout &L = ctx->L;
    if (ctx->permuted) launch_gather(ctx->launcher, L.nloc, ctx->i2c, canon_local, dst_local);
    else CK(cudaMemcpyAsync(dst_local, canon_local, sizeof(double) * L.nloc, cudaMemcpyDeviceToDevice, ctx->stream));
    return WAVE_OK;
}
// row-indexed (owned, storage order) vector -> canonical order of the owned rows
int own_to_canonical(wave_ctx *ctx, const double *src_own, double *dst_own) {
    const Layout &L = ctx->L;
    if (ctx->permuted) launch_gather(ctx->launcher, L.nown, ctx->c2i + L.own_off, src_own - L.own_off, dst_own);
    else CK(cudaMemcpyAsync(dst_own, src_own, sizeof(double) * L.nown, cudaMemcpyDeviceToDevice, ctx->stream));
    return WAVE_OK;
}
// owned rows -> host (canonical order), asynchronous on the context stream
int download_own_async(wave_ctx *ctx, const double *src_own, double *host_own) {
    const Layout &L = ctx->L;
    const double *from = src_own;
    if (ctx->permuted) {
        RET(own_to_canonical(ctx, src_own, ctx->tmp));
        from = ctx->tmp;
    }
    CK(cudaMemcpyAsync(host_own, from, sizeof(double) * L.nown, cudaMemcpyDeviceToHost, ctx->stream));
    return WAVE_OK;
}
int upload_local(wave_ctx *ctx, const double *host_global, double *dst_local) {
    const Layout &L = ctx->L;
    double *stage = ctx->permuted ? ctx->tmp : dst_local;
    CK(cudaMemcpyAsync(stage, host_global + L.col0, sizeof(double) * L.nloc, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->permuted) RET(to_storage(ctx, stage, dst_local));
    return WAVE_OK;
}

// owned part of a local-layout vector -> all ranks' canonical vector in device scratch
int gather_global(wave_ctx *ctx, const double *own_src) {
    const int64_t n = n_dofs(ctx->L.mesh);
    RET(ensure_scratch(ctx, n));
    const Layout &L = ctx->L;
    RET(own_to_canonical(ctx, own_src, ctx->scratch + L.row0));
    if (ctx->cfg.nranks > 1) {
        for (int r = 0; r < ctx->cfg.nranks; ++r) {
            const Layout Lr = make_layout(L.mesh, r, ctx->cfg.nranks);
            NK(g_nccl.Broadcast(ctx->scratch + Lr.row0, ctx->scratch + Lr.row0, (size_t)Lr.nown, Nccl::kFloat64, r,
                                ctx->comm, ctx->stream));
        }
    }
    return WAVE_OK;
}

// temporary device allocation released on every exit path
template <class T>
struct DevTmp {
    T *p = nullptr;
    ~DevTmp() { if (p) cudaFree(p); }
    DevTmp() = default;
    DevTmp(const DevTmp &) = delete;
    DevTmp &operator=(const DevTmp &) = delete;
};

template <class T>
int dev_alloc(wave_ctx *ctx, T **p, size_t count, bool zero = true) {
    CK(cudaMalloc((void **)p, sizeof(T) * (count ? count : 1)));
    if (zero) CK(cudaMemsetAsync(*p, 0, sizeof(T) * (count ? count : 1), ctx->stream));
    return WAVE_OK;
}

// K6f plan: block b of the cooperative kernel owns the windows [b wpb, (b+1) wpb) and stages the column
// range of their entries in shared memory.  Possible when every block's rows (x, g: 16 B per row) and
// its staged range fit in one SM's shared memory; otherwise the three-kernel iteration stays.
int fused_plan(wave_ctx *ctx) {
    auto &f = ctx->fused;
    f.ok = false;
    // default: on whenever the rows fit on chip -- one rank, or 2..8 ranks over the NVLink mailboxes (measured
    // on B200: c2, one GPU: 32 us per iteration against 48 us of the round-1 three-kernel path; 1 M-DoF strips
    // on 8 GPUs: 47 us against 55 us); WAVE_CG_FUSED=0 switches it off.  Multigrid levels never use it.
    const char *env = std::getenv("WAVE_CG_FUSED");
    const int want = env ? std::atoi(env) : -1;
    if (want == 0 || (ctx->cfg.flags & kFlagChild)) return WAVE_OK;
    if (ctx->cfg.precond != WAVE_PRECOND_JACOBI) return WAVE_OK;
    if (ctx->cfg.nranks != 1 && !ctx->pc.enabled) return WAVE_OK;  // several ranks: only over the NVLink mailboxes
    // every rank evaluates its own strip; with several ranks all of them must agree (the kernel's sums and
    // halo flags pair up across ranks), so the local verdicts are combined below
    const int nwin = ctx->nslices / (kWindow / kSlice);
    int sms = kSMsFallback;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int wpb = (nwin + sms - 1) / sms;
    const int grid = (nwin + wpb - 1) / wpb;
    bool can = wpb <= kFusedMaxWin;
    std::vector<int32_t> c0((size_t)grid), cn((size_t)grid);
    int stage = 1;
    size_t smem = 0;
    if (can) {
        DevTmp<int32_t> cmin, cmax;
        RET(dev_alloc(ctx, &cmin.p, (size_t)nwin, false));
        RET(dev_alloc(ctx, &cmax.p, (size_t)nwin, false));
        launch_window_col_range(ctx->launcher, ctx->A, nwin, cmin.p, cmax.p);
        std::vector<int32_t> lo(nwin), hi(nwin);
        CK(cudaMemcpyAsync(lo.data(), cmin.p, sizeof(int32_t) * nwin, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(hi.data(), cmax.p, sizeof(int32_t) * nwin, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int b = 0; b < grid; ++b) {
            int32_t l = INT32_MAX, h = -1;
            for (int w = b * wpb; w < std::min(nwin, (b + 1) * wpb); ++w) {
                l = std::min(l, lo[w]);
                h = std::max(h, hi[w]);
            }
            if (h < l) { l = ctx->L.own_off; h = ctx->L.own_off; }  // a block of padding only
            c0[b] = l;
            cn[b] = h - l + 1;
            stage = std::max(stage, (int)cn[b]);
        }
        smem = cg_fused_smem_bytes(wpb, stage);
        can = cg_fused_supported(grid, smem);
    }
    if (ctx->cfg.nranks > 1) {  // unanimous or not at all
        ctx->hres[0] = can ? 1.0 : 0.0;
        CK(cudaMemcpyAsync(ctx->res, ctx->hres, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        RET(allreduce(ctx, ctx->res, 1));
        CK(cudaMemcpyAsync(ctx->hres, ctx->res, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        can = ctx->hres[0] == (double)ctx->cfg.nranks;
    }
    if (!can) return WAVE_OK;
    RET(dev_alloc(ctx, &f.blk_c0, (size_t)grid, false));
    RET(dev_alloc(ctx, &f.blk_cn, (size_t)grid, false));
    RET(dev_alloc(ctx, &f.partials, (size_t)grid * 4));
    RET(dev_alloc(ctx, &f.pub, 8));
    CK(cudaMemcpyAsync(f.blk_c0, c0.data(), sizeof(int32_t) * grid, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.blk_cn, cn.data(), sizeof(int32_t) * grid, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    f.grid = grid;
    f.wpb = wpb;
    f.stage_cap = stage;
    f.smem = smem;
    f.ok = true;
    return WAVE_OK;
}

// The coarse levels of the V-cycle for the scheme matrix bc(M + s K) on cfg's mesh, as a pure function of the
// configuration (host only; exported as wave_mg_plan for the CPU tests): [P1 on the same mesh when R = 2,
// then] P1 on Nel/2, Nel/4, ... while the stiffness part still matters (s c0^2 / (dx dy) > 1/4), the mesh
// halves evenly and -- with several ranks -- every strip begins and ends at an even quad row and keeps at
// least two coarse quad rows (coarse quad row J covers the fine quad rows 2J, 2J+1, so fine and coarse
// strips then split at the same physical lines).  Every rank takes the same decisions.
struct MgPlanLevel { int nx, ny, r; };
int mg_plan(const wave_config &cfg, double s, double c0, MgPlanLevel *out, int max_levels) {
    int nx = cfg.nx, ny = cfg.ny, r = cfg.r, n = 0;
    while (n < max_levels) {
        int nnx = nx, nny = ny;
        if (r == 1) {
            const double dx = (cfg.x1 - cfg.x0) / nx, dy = (cfg.y1 - cfg.y0) / ny;
            const bool matters = s * c0 * c0 / (dx * dy) > 0.25;
            if (!(matters && nx % 2 == 0 && ny % 2 == 0 && std::min(nx, ny) / 2 >= 2)) break;
            bool strips_ok = true;
            for (int rk = 0; rk < cfg.nranks && cfg.nranks > 1; ++rk) {
                int j0, j1, c0r, c1r;
                quad_row_split(ny, rk, cfg.nranks, j0, j1);
                quad_row_split(ny / 2, rk, cfg.nranks, c0r, c1r);
                strips_ok = strips_ok && j0 % 2 == 0 && j1 % 2 == 0 && c0r == j0 / 2 && c1r == j1 / 2 && c1r - c0r >= 2;
            }
            if (!strips_ok) break;
            nnx = nx / 2;
            nny = ny / 2;
        }
        out[n++] = MgPlanLevel{nnx, nny, 1};
        nx = nnx; ny = nny; r = 1;
    }
    return n;
}

// Detect the translation-invariant rows of M and K (kernels.cuh) and switch the SpMV of those slices to
// the table-driven path.  Off with WAVE_FLAG_NO_STENCIL / WAVE_NO_STENCIL=1, on meshes too small to have a
// generic middle quad, and when fewer than half of the rows match (variable wave speed).
int stencil_setup(wave_ctx *ctx) {
    const Layout &L = ctx->L;
    const char *env = std::getenv("WAVE_NO_STENCIL");
    if ((ctx->cfg.flags & WAVE_FLAG_NO_STENCIL) || (env && std::atoi(env) != 0)) return WAVE_OK;
    if (L.mesh.nx < 8 || L.mesh.ny < 8) return WAVE_OK;
    RET(dev_alloc(ctx, &ctx->st_tab, (size_t)4 * kTabSize));
    RET(dev_alloc(ctx, &ctx->st_meta, (size_t)kTabSize + kStencilKinds));
    RET(dev_alloc(ctx, &ctx->slice_info, (size_t)ctx->nslices, false));
    DevTmp<int8_t> row_kind;
    DevTmp<unsigned long long> counts;
    RET(dev_alloc(ctx, &row_kind.p, (size_t)L.nown, false));
    RET(dev_alloc(ctx, &counts.p, 2));
    launch_stencil_tables(ctx->launcher, L, ctx->dprog + WAVE_EXPR_C, &ctx->q_asm, ctx->st_meta, ctx->st_tab,
                          ctx->st_tab + kTabSize);
    launch_stencil_classify(ctx->launcher, L, ctx->A, ctx->M, ctx->K, ctx->st_meta, ctx->st_tab,
                            ctx->st_tab + kTabSize, row_kind.p, ctx->slice_info, counts.p);
    unsigned long long h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, counts.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // host copies of the tables; the kernel's unrolled rows assume the row lengths of the element
    ctx->h_st_meta.assign((size_t)kTabSize + kStencilKinds, 0);
    ctx->h_tab.assign((size_t)4 * kTabSize, 0.0);
    CK(cudaMemcpyAsync(ctx->h_st_meta.data(), ctx->st_meta, sizeof(int32_t) * ctx->h_st_meta.size(),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_tab.data(), ctx->st_tab, sizeof(double) * 2 * kTabSize, cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int32_t *len = ctx->h_st_meta.data() + kTabSize;
    const bool lengths_ok = L.mesh.r == 1 ? len[0] == 7 : (len[0] == 19 && len[1] == 9 && len[2] == 9 && len[3] == 9);
    if (!lengths_ok || (int64_t)h[0] * 2 < (int64_t)L.nown) {
        ctx->h_st_meta.clear();
        ctx->h_tab.clear();
        cudaFree(ctx->st_tab); ctx->st_tab = nullptr;
        cudaFree(ctx->st_meta); ctx->st_meta = nullptr;
        cudaFree(ctx->slice_info); ctx->slice_info = nullptr;
        return WAVE_OK;
    }
    // the slices that stay in SELL form, in ascending order (built on the host: the order fixes which warp
    // takes which slice and with it the order of the partial sums)
    std::vector<int2> info((size_t)ctx->nslices);
    CK(cudaMemcpyAsync(info.data(), ctx->slice_info, sizeof(int2) * info.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<int32_t> list;
    for (int sl = 0; sl < ctx->nslices; ++sl)
        if (info[(size_t)sl].x < 0) list.push_back(sl);
    // the stencil slices in tile order: key = (group of DoF lines, column chunk, line in the group, kind), so
    // that the 8 warps of a block take slices of one column range in neighbouring lines / kinds
    {
        const Mesh &m = L.mesh;
        const int lines_per_tile = m.r == 1 ? 8 : 2;
        std::vector<std::pair<uint64_t, int32_t>> keyed;
        keyed.reserve((size_t)ctx->nslices - list.size());
        const int64_t b0 = block_start(m, 1), bsz = m.r == 1 ? (m.nx + 1) : (4LL * m.nx + 2);
        for (int sl = 0; sl < ctx->nslices; ++sl) {
            const int2 si = info[(size_t)sl];
            if (si.x < 0) continue;
            uint64_t key;
            if (si.y < 0) key = ~0ull - (uint64_t)(ctx->nslices - sl);  // rows not consecutive: keep at the end
            else {
                const int64_t g = (int64_t)si.y + L.row0;
                const int64_t j = g < b0 ? 0 : 1 + (g - b0) / bsz;
                const int64_t pos = g - block_start(m, (int)j);
                int64_t i = pos;
                if (m.r == 2 && j >= 1) {
                    const int64_t seg[4] = {0, m.nx + 1, 2LL * m.nx + 1, 3LL * m.nx + 2};
                    i = pos - seg[si.x];
                }
                key = ((uint64_t)(j / lines_per_tile) << 40) | ((uint64_t)(i / kSlice) << 16) |
                      ((uint64_t)(j % lines_per_tile) << 4) | (uint64_t)si.x;
            }
            keyed.emplace_back(key, sl);
        }
        const char *ord = std::getenv("WAVE_STENCIL_ORDER");  // experiment knob: "linear" = ascending slices
        if (ord && std::string(ord) == "linear")
            for (auto &kv : keyed) kv.first = (uint64_t)kv.second;
        std::sort(keyed.begin(), keyed.end());
        // several ranks: rotate the walk by half its length (ghost-reading slices mid-kernel) and mark them
        const int nr = ctx->cfg.nranks, rk = ctx->cfg.rank;
        int lo_count = 0, hi_count = 0;
        if (nr > 1 && rk > 0) lo_count = (int)(block_start(m, L.jq0 + 1) - block_start(m, L.jq0));
        if (nr > 1 && rk < nr - 1) hi_count = (int)(block_start(m, L.jq1) - block_start(m, L.jq1 - 1));
        const int ghost_lo = lo_count ? ((lo_count + kWindow - 1) / kWindow) * (kWindow / kSlice) : 0;
        const int ghost_hi0 = hi_count ? ((L.nown - hi_count) / kWindow) * (kWindow / kSlice) : ctx->nslices;
        const size_t n_st = keyed.size(), rot = nr > 1 ? n_st / 2 : 0;
        std::vector<int2> order(n_st + (size_t)kStencilPad, make_int2(-1, -1));
        for (size_t k = 0; k < n_st; ++k) {
            const int32_t sl = keyed[(k + rot) % n_st].second;
            const int ghost = (sl < ghost_lo || sl >= ghost_hi0) ? 1 : 0;
            order[k] = make_int2(sl | (ghost << 27) | (info[(size_t)sl].x << 28), info[(size_t)sl].y);
        }
        RET(dev_alloc(ctx, &ctx->st_walk, order.size(), false));
        CK(cudaMemcpyAsync(ctx->st_walk, order.data(), sizeof(int2) * order.size(), cudaMemcpyHostToDevice,
                           ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->A.st_walk = ctx->st_walk;
        ctx->A.n_st = (int)n_st;
    }
    RET(dev_alloc(ctx, &ctx->sell_list, list.size(), false));
    if (!list.empty())
        CK(cudaMemcpyAsync(ctx->sell_list, list.data(), sizeof(int32_t) * list.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stencil_rows = (int64_t)h[0];
    ctx->sell_nnz = (int64_t)h[1];
    ctx->A.slice_info = ctx->slice_info;
    ctx->A.st_meta = ctx->h_st_meta.data();
    ctx->A.sell_list = ctx->sell_list;
    ctx->A.n_sell = (int)list.size();
    return WAVE_OK;
}

// Build the multigrid hierarchy for the scheme matrix bc(M + s K): [P2 on the mesh ->] P1 on the mesh ->
// P1 on Nel/2, Nel/4, ... while the stiffness part still matters (s c^2 / (dx dy) > 1/4) and the mesh
// halves evenly.  Every coarse level is a child context (pattern, rediscretised M and K, Dirichlet rows,
// Jacobi diagonal) running on the parent's stream; its u / unew / rhs / g vectors serve as x / x2 / b / r.
int mg_setup(wave_ctx *ctx, double s) {
    if (ctx->mg && ctx->mg->s == s) return WAVE_OK;
    if (ctx->mg) {
        for (int l = 1; l < ctx->mg->nlev; ++l) wave_destroy(ctx->mg->lev[l].c);
        delete ctx->mg;
        ctx->mg = nullptr;
    }
    Mg *m = new Mg();
    m->s = s;
    const Layout &L = ctx->L;
    for (int k = 0; k < 3; ++k)
        if (!ctx->mg_buf[k]) RET(dev_alloc(ctx, &ctx->mg_buf[k], (size_t)L.nloc));
    m->lev[0].c = ctx;
    m->lev[0].x = ctx->mg_buf[0];   // local layout (nloc), as x2 and r
    m->lev[0].x2 = ctx->mg_buf[1];
    m->lev[0].r = ctx->mg_buf[2];
    m->lev[0].omega = L.mesh.r == 2 ? 0.5 : 0.8;
    m->nlev = 1;
    const double cx = 0.5 * (ctx->cfg.x0 + ctx->cfg.x1), cy = 0.5 * (ctx->cfg.y0 + ctx->cfg.y1);
    const double c0 = eval(&ctx->hprog[WAVE_EXPR_C], cx, cy, 0.0);
    MgPlanLevel plan[12];
    const int n_coarse = mg_plan(ctx->cfg, s, c0, plan, 11);
    for (int lev = 0; lev < n_coarse; ++lev) {
        const int nnx = plan[lev].nx, nny = plan[lev].ny;
        wave_config cfg = ctx->cfg;
        cfg.nx = nnx; cfg.ny = nny; cfg.r = 1;
        cfg.scheme = WAVE_SCHEME_NEWMARK;
        cfg.nccl_unique_id = nullptr; cfg.device = -1;  // rank / nranks of the solver context, its communicator
        cfg.precond = WAVE_PRECOND_JACOBI;
        cfg.flags = kFlagChild | (ctx->cfg.flags & WAVE_FLAG_NO_STENCIL);
        cfg.stream = ctx->stream;
        wave_ctx *c = nullptr;
        auto drop_levels = [&]() {
            for (int l = 1; l < m->nlev; ++l) wave_destroy(m->lev[l].c);
            delete m;
        };
        g_child_comm = ctx->comm;
        const int crc = wave_create(&cfg, &c);
        g_child_comm = nullptr;
        if (crc != WAVE_OK) { drop_levels(); return fail(ctx, WAVE_ERR_CUDA, wave_last_error(nullptr)); }
        const Program zero = compile_expression("0.0", "x, y, t", "");
        for (int k = 0; k < WAVE_EXPR_SOLUTION; ++k) {
            c->hprog[k] = k == WAVE_EXPR_C ? ctx->hprog[k] : zero;
            c->has[k] = true;
        }
        cudaMemcpyAsync(c->dprog, c->hprog, sizeof(Program) * WAVE_EXPR_COUNT, cudaMemcpyHostToDevice, c->stream);
        int rc = wave_setup(c);
        if (rc == WAVE_OK) rc = build_system_matrix(c, s, c->S1, c->dinv1, c->d0);
        if (rc != WAVE_OK) {
            const std::string msg = std::string("multigrid level setup: ") + wave_last_error(c);
            wave_destroy(c);
            drop_levels();
            return fail(ctx, rc, msg);
        }
        MgLevel &lv = m->lev[m->nlev++];
        lv.c = c;
        lv.S = c->S1;
        lv.dinv = c->dinv1;
        lv.x = c->u;
        lv.x2 = c->unew;
        lv.b = c->rhs;
        lv.r = c->d;  // local layout: the restriction reads its upper ghost block
        lv.omega = 0.8;
        ctx->launches += c->launches;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->mg = m;
    return WAVE_OK;
}

// Map every rank's mailbox and the neighbours' search-direction vectors through CUDA IPC; the 64-byte
// handles travel over NCCL broadcasts.  Disabled (NCCL collectives stay) for nranks > kMaxPeers or
// when WAVE_NO_P2P is set.
int setup_peer_exchange(wave_ctx *ctx) {
    const int R = ctx->cfg.nranks, rank = ctx->cfg.rank;
    if (R == 1 || R > kMaxPeers || std::getenv("WAVE_NO_P2P") || (ctx->cfg.flags & kFlagChild)) return WAVE_OK;
    RET(dev_alloc(ctx, &ctx->mailbox, 1));
    struct Handles { cudaIpcMemHandle_t box, d; };
    std::vector<Handles> all((size_t)R);
    CK(cudaIpcGetMemHandle(&all[rank].box, ctx->mailbox));
    CK(cudaIpcGetMemHandle(&all[rank].d, ctx->d));
    Handles *dev = nullptr;
    RET(dev_alloc(ctx, &dev, (size_t)R));
    CK(cudaMemcpyAsync(dev + rank, &all[rank], sizeof(Handles), cudaMemcpyHostToDevice, ctx->stream));
    for (int r = 0; r < R; ++r)
        NK(g_nccl.Broadcast(dev + r, dev + r, sizeof(Handles), /*ncclInt8*/ 0, r, ctx->comm, ctx->stream));
    CK(cudaMemcpyAsync(all.data(), dev, sizeof(Handles) * R, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(dev));
    PeerComm pc{};
    pc.rank = rank;
    pc.nranks = R;
    pc.status = &ctx->S->peer_timeout;
    for (int r = 0; r < R; ++r) {
        if (r == rank) { pc.box[r] = ctx->mailbox; continue; }
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, all[r].box, cudaIpcMemLazyEnablePeerAccess));
        ctx->ipc_opened[ctx->n_ipc_opened++] = p;
        pc.box[r] = (PeerMailbox *)p;
    }
    const Layout &L = ctx->L;
    const Mesh &m = L.mesh;
    if (rank > 0) {  // my first block is the upper ghost block of rank-1
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, all[rank - 1].d, cudaIpcMemLazyEnablePeerAccess));
        ctx->ipc_opened[ctx->n_ipc_opened++] = p;
        const Layout Lo = make_layout(m, rank - 1, R);
        pc.d_lo = (double *)p + Lo.own_off + Lo.nown;
        pc.lo_count = (int)(block_start(m, L.jq0 + 1) - block_start(m, L.jq0));
    }
    if (rank < R - 1) {  // my last block is the lower ghost block of rank+1 (its local index 0)
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, all[rank + 1].d, cudaIpcMemLazyEnablePeerAccess));
        ctx->ipc_opened[ctx->n_ipc_opened++] = p;
        pc.d_hi = (double *)p;
        pc.hi_count = (int)(block_start(m, L.jq1) - block_start(m, L.jq1 - 1));
    }
    pc.enabled = 1;
    // nobody may store into a peer before every rank has mapped everything
    RET(allreduce(ctx, ctx->res, 1));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->pc = pc;
    return WAVE_OK;
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

void wave_default_config(wave_config *c) {
    std::memset(c, 0, sizeof(*c));
    c->nx = c->ny = 40;                 // "Nel" default, src/ParameterReader.cpp:41-44
    c->x0 = 0.0; c->x1 = 1.0; c->y0 = 0.0; c->y1 = 1.0;
    c->r = 1; c->scheme = WAVE_SCHEME_NEWMARK;
    c->dt = 0.01; c->theta = 0.5; c->beta = 0.25; c->gamma = 0.5;
    c->cg_maxit = 10000; c->cg_tol = 1e-12; c->cg_reduce = 1e-6;  // src/WaveNewmark.cpp:256
    c->precond = WAVE_PRECOND_JACOBI;
    c->rank = 0; c->nranks = 1; c->device = -1;
}

const char *wave_last_error(const wave_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int wave_device_count(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
        cudaGetLastError();  // no driver / no device is an answer (0), not a sticky error
        return 0;
    }
    return ndev;
}

int wave_comm_unique_id(void *out128) {
    std::string err;
    if (!g_nccl.load(err)) return fail(nullptr, WAVE_ERR_CUDA, err);
    Nccl::Uid id;
    if (g_nccl.GetUniqueId(&id) != 0) return fail(nullptr, WAVE_ERR_CUDA, "ncclGetUniqueId failed");
    std::memcpy(out128, &id, sizeof id);
    return WAVE_OK;
}

int wave_create(const wave_config *cfg, wave_ctx **out) {
    if (!cfg || !out) return fail(nullptr, WAVE_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->nx < 1 || cfg->ny < 1) return fail(nullptr, WAVE_ERR_ARG, "Nel must be >= 1");
    if (cfg->r != 1 && cfg->r != 2)
        return fail(nullptr, WAVE_ERR_UNSUPPORTED, "only FE_SimplexP degree R = 1 or 2 is implemented");
    if (!(cfg->x1 > cfg->x0) || !(cfg->y1 > cfg->y0)) return fail(nullptr, WAVE_ERR_ARG, "empty geometry");
    if (!(cfg->dt > 0.0)) return fail(nullptr, WAVE_ERR_ARG, "Dt must be positive");
    if (cfg->scheme != WAVE_SCHEME_NEWMARK && cfg->scheme != WAVE_SCHEME_THETA)
        return fail(nullptr, WAVE_ERR_ARG, "unknown scheme");
    if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks)
        return fail(nullptr, WAVE_ERR_ARG, "bad rank / nranks");
    if (cfg->ny < cfg->nranks) return fail(nullptr, WAVE_ERR_ARG, "need at least one quad row per rank");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, WAVE_ERR_CUDA, "no CUDA device: libwavegpu has no CPU fallback");
    wave_ctx *ctx = new wave_ctx();
    ctx->cfg = *cfg;
    if (ctx->cfg.cg_maxit <= 0) ctx->cfg.cg_maxit = 10000;
    if (ctx->cfg.cg_tol <= 0.0) ctx->cfg.cg_tol = 1e-12;
    if (ctx->cfg.cg_reduce <= 0.0) ctx->cfg.cg_reduce = 1e-6;
    auto bail = [&](int code) { g_create_error = ctx->err; wave_destroy(ctx); return code; };
    if (cfg->device >= 0 && cudaSetDevice(cfg->device) != cudaSuccess) {
        ctx->err = "cudaSetDevice failed";
        return bail(WAVE_ERR_CUDA);
    }
    if (cfg->stream) {
        ctx->stream = (cudaStream_t)cfg->stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            ctx->err = "cudaStreamCreate failed";
            return bail(WAVE_ERR_CUDA);
        }
        ctx->own_stream = true;
    }
    ctx->launcher = Launcher{ctx->stream, &ctx->launches, &ctx->launch_error};
    for (auto &e : ctx->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { ctx->err = "cudaEventCreate failed"; return bail(WAVE_ERR_CUDA); }
    Mesh m{};
    m.nx = cfg->nx; m.ny = cfg->ny; m.r = cfg->r;
    m.x0 = cfg->x0; m.y0 = cfg->y0;
    m.dx = (cfg->x1 - cfg->x0) / cfg->nx;
    m.dy = (cfg->y1 - cfg->y0) / cfg->ny;
    ctx->L = make_layout(m, cfg->rank, cfg->nranks);
    if ((int64_t)ctx->L.nloc * (cfg->r == 1 ? 7 : 12) >= (1LL << 32)) {  // padded entries must fit uint32 offsets
        ctx->err = "local problem exceeds 32-bit CSR offsets: partition over more GPUs";
        return bail(WAVE_ERR_UNSUPPORTED);
    }
    if (cfg->nranks > 1 && (cfg->flags & kFlagChild)) {
        ctx->comm = (Nccl::comm_t)g_child_comm;
        ctx->own_comm = false;
        if (!ctx->comm) { ctx->err = "multigrid level without a communicator"; return bail(WAVE_ERR_ARG); }
    } else if (cfg->nranks > 1) {
        if (!cfg->nccl_unique_id) { ctx->err = "nccl_unique_id required for nranks > 1"; return bail(WAVE_ERR_ARG); }
        if (!g_nccl.load(ctx->err)) return bail(WAVE_ERR_CUDA);
        Nccl::Uid id;
        std::memcpy(&id, cfg->nccl_unique_id, sizeof id);
        const int rc = g_nccl.CommInitRank(&ctx->comm, cfg->nranks, id, cfg->rank);
        if (rc != 0) { ctx->err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc); return bail(WAVE_ERR_CUDA); }
    }
    if (cudaMalloc(&ctx->dprog, sizeof(Program) * WAVE_EXPR_COUNT) != cudaSuccess) {
        ctx->err = "cudaMalloc failed";
        return bail(WAVE_ERR_CUDA);
    }
    ctx->q_asm = make_quadrature(cfg->r + 1);  // src/WaveEquationBase.cpp:82
    ctx->q_err = make_quadrature(cfg->r + 2);  // src/WaveEquationBase.cpp:371
    *out = ctx;
    return WAVE_OK;
}

void wave_destroy(wave_ctx *ctx) {
    if (!ctx) return;
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->mg) {
        for (int l = 1; l < ctx->mg->nlev; ++l) wave_destroy(ctx->mg->lev[l].c);
        delete ctx->mg;
    }
    for (double *b : ctx->mg_buf)
        if (b) cudaFree(b);
    for (int k = 0; k < ctx->n_ipc_opened; ++k) cudaIpcCloseMemHandle(ctx->ipc_opened[k]);
    if (ctx->mailbox) cudaFree(ctx->mailbox);
    for (void *q : {(void *)ctx->fused.blk_c0, (void *)ctx->fused.blk_cn, (void *)ctx->fused.partials,
                    (void *)ctx->fused.pub})
        if (q) cudaFree(q);
    void *ptrs[] = {ctx->st_tab, ctx->st_meta, ctx->slice_info, ctx->sell_list, ctx->st_walk, ctx->dprog, ctx->rowptr, ctx->slice_ptr, ctx->row_of, ctx->slot_of, ctx->col, ctx->c2i, ctx->i2c, ctx->tmp, ctx->M, ctx->K, ctx->S1, ctx->S2, ctx->dinv1, ctx->dinv2,
                    ctx->d0, ctx->u, ctx->v, ctx->a, ctx->unew, ctx->d, ctx->rhs, ctx->fvec, ctx->cellvec, ctx->g, ctx->h,
                    ctx->scratch, ctx->brow, ctx->bx, ctx->by, ctx->partials, ctx->partials2, ctx->partials3, ctx->counter, ctx->S, ctx->res,
                    ctx->flush_buf};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (ctx->hS) cudaFreeHost(ctx->hS);
    if (ctx->hres) cudaFreeHost(ctx->hres);
    if (ctx->comm && ctx->own_comm) g_nccl.CommDestroy(ctx->comm);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : ctx->spmv_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int wave_set_expr(wave_ctx *ctx, int which, const char *expression, const char *variable_names,
                  const char *constants) {
    if (!ctx || which < 0 || which >= WAVE_EXPR_COUNT || !expression)
        return fail(ctx, WAVE_ERR_ARG, "bad expression slot");
    try {
        ctx->hprog[which] = compile_expression(expression, variable_names ? variable_names : "",
                                               constants ? constants : "");
    } catch (const std::exception &e) {
        return fail(ctx, WAVE_ERR_EXPR, e.what());
    }
    ctx->has[which] = true;
    if (which == WAVE_EXPR_F) {
        ctx->f_text[0] = expression;
        ctx->f_text[1] = variable_names ? variable_names : "";
        ctx->f_text[2] = constants ? constants : "";
    }
    CK(cudaMemcpyAsync(ctx->dprog + which, &ctx->hprog[which], sizeof(Program), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return WAVE_OK;
}

struct wave_expr {
    Program prog;
};
int wave_expr_create(const char *expression, const char *variable_names, const char *constants, wave_expr **out,
                     char *errbuf, size_t errbuf_len) {
    if (!expression || !out) return WAVE_ERR_ARG;
    *out = nullptr;
    try {
        wave_expr *e = new wave_expr();
        e->prog = compile_expression(expression, variable_names ? variable_names : "", constants ? constants : "");
        *out = e;
        return WAVE_OK;
    } catch (const std::exception &ex) {
        if (errbuf && errbuf_len) std::snprintf(errbuf, errbuf_len, "%s", ex.what());
        return WAVE_ERR_EXPR;
    }
}
double wave_expr_value(const wave_expr *e, double x, double y, double t) { return eval(&e->prog, x, y, t); }
int wave_expr_is_time_dependent(const wave_expr *e) { return e->prog.time_dependent; }
void wave_expr_destroy(wave_expr *e) { delete e; }

int wave_eval_expr(wave_ctx *ctx, int which, double x, double y, double t, double *out) {
    if (!ctx || which < 0 || which >= WAVE_EXPR_COUNT || !ctx->has[which] || !out)
        return fail(ctx, WAVE_ERR_ARG, "expression not set");
    *out = eval(&ctx->hprog[which], x, y, t);
    return WAVE_OK;
}

int wave_setup(wave_ctx *ctx) {
    if (!ctx) return WAVE_ERR_ARG;
    for (int k = WAVE_EXPR_C; k <= WAVE_EXPR_DGDT; ++k)
        if (!ctx->has[k]) {
            static const char *names[] = {"C", "F", "U0", "V0", "G", "DGDT"};
            return fail(ctx, WAVE_ERR_EXPR,
                        std::string("Function expression for '") + names[k] + "' must be specified in the parameter file.");
        }
    if (ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_setup called twice");
    const Layout &L = ctx->L;
    const Launcher &l = ctx->launcher;
    double fconst = 1.0;
    ctx->forcing_active = !(is_constant(ctx->hprog[WAVE_EXPR_F], &fconst) && fconst == 0.0) ||
                          (ctx->cfg.flags & WAVE_FLAG_FORCING_EVERY_STEP);

    // ---- storage numbering (kind-major inside each block for P2, see mesh.h) -------------------------
    ctx->permuted = L.mesh.r == 2;
    RET(dev_alloc(ctx, &ctx->tmp, (size_t)L.nloc));
    if (ctx->permuted) {
        RET(dev_alloc(ctx, &ctx->c2i, (size_t)L.nloc, false));
        RET(dev_alloc(ctx, &ctx->i2c, (size_t)L.nloc, false));
        launch_build_perm(l, L, ctx->c2i, ctx->i2c);
    }

    // ---- sparsity: row lengths -> CSR row pointer; window sort -> SELL slices -> columns ----------
    const int nslots = ((L.nown + kWindow - 1) / kWindow) * kWindow;
    ctx->nslices = nslots / kSlice;
    uint32_t *rowlen = nullptr, *slice_cnt = nullptr;
    RET(dev_alloc(ctx, &rowlen, (size_t)L.nown + 1));
    RET(dev_alloc(ctx, &slice_cnt, (size_t)ctx->nslices + 1));
    RET(dev_alloc(ctx, &ctx->rowptr, (size_t)L.nown + 1));
    RET(dev_alloc(ctx, &ctx->slice_ptr, (size_t)ctx->nslices + 1));
    RET(dev_alloc(ctx, &ctx->row_of, (size_t)nslots, false));
    RET(dev_alloc(ctx, &ctx->slot_of, (size_t)L.nown, false));
    launch_row_lengths(l, L, rowlen);
    launch_window_sort(l, L.nown, nslots, rowlen, ctx->row_of, ctx->slot_of);
    launch_slice_sizes(l, ctx->nslices, ctx->row_of, rowlen, slice_cnt);
    {
        void *tmp = nullptr;
        size_t bytes = 0, bytes2 = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, rowlen, ctx->rowptr, L.nown + 1, ctx->stream));
        CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes2, slice_cnt, ctx->slice_ptr, ctx->nslices + 1, ctx->stream));
        CK(cudaMalloc(&tmp, std::max(bytes, bytes2)));
        CK(cub::DeviceScan::ExclusiveSum(tmp, bytes, rowlen, ctx->rowptr, L.nown + 1, ctx->stream));
        CK(cub::DeviceScan::ExclusiveSum(tmp, bytes2, slice_cnt, ctx->slice_ptr, ctx->nslices + 1, ctx->stream));
        ctx->launches += 2;
        uint32_t total = 0, total_pad = 0;
        CK(cudaMemcpyAsync(&total, ctx->rowptr + L.nown, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(&total_pad, ctx->slice_ptr + ctx->nslices, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                           ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(tmp));
        ctx->nnz = total;
        ctx->nnz_pad = total_pad;
    }
    CK(cudaFree(rowlen));
    CK(cudaFree(slice_cnt));
    RET(dev_alloc(ctx, &ctx->col, (size_t)ctx->nnz_pad, false));
    ctx->A = Sell{ctx->slice_ptr, ctx->col, ctx->row_of, ctx->slot_of, ctx->rowptr, ctx->nslices, L.nown,
                  L.mesh.r == 1 ? 7 : 10, L.own_off, nullptr, nullptr, nullptr, 0, nullptr, 0};
    launch_fill_int(l, ctx->nnz_pad, L.own_off, ctx->col);  // padding entries: a valid column, value 0
    launch_fill_cols(l, L, ctx->A, ctx->col);

    // ---- M, K ---------------------------------------------------------------------------------------
    RET(dev_alloc(ctx, &ctx->M, (size_t)ctx->nnz_pad));
    RET(dev_alloc(ctx, &ctx->K, (size_t)ctx->nnz_pad));
    launch_assemble(l, L, ctx->dprog + WAVE_EXPR_C, &ctx->q_asm, ctx->A, ctx->M, ctx->K);
    ctx->sell_nnz = ctx->nnz;
    RET(stencil_setup(ctx));

    // ---- boundary list (closed form, host) --------------------------------------------------------
    {
        const Mesh &m = L.mesh;
        struct B { int64_t dof, idof; double x, y; };
        std::vector<B> mine;
        ctx->h_bdof_global.clear();
        const int nk = m.r == 1 ? 1 : 4;
        for (int j = 0; j <= m.ny; ++j) {
            const bool edge_row = (j == 0 || j == m.ny);
            for (int i = 0; i <= m.nx; ++i) {
                if (!edge_row && i != 0 && i != m.nx) continue;
                for (int kind = 0; kind < nk; ++kind) {
                    if (!entity_on_boundary(m, i, j, kind)) continue;
                    const int64_t dof = entity_dof(m, i, j, kind);
                    if (dof < 0) continue;
                    ctx->h_bdof_global.push_back((int32_t)dof);
                    if (dof >= L.row0 && dof < L.row0 + L.nown) {
                        double x, y;
                        entity_point(m, i, j, kind, x, y);
                        mine.push_back({dof, entity_dof_internal(m, i, j, kind), x, y});
                    }
                }
            }
        }
        std::sort(ctx->h_bdof_global.begin(), ctx->h_bdof_global.end());
        std::sort(mine.begin(), mine.end(), [](const B &a, const B &b) { return a.dof < b.dof; });
        ctx->nb = (int)mine.size();
        ctx->nb_global = (int64_t)ctx->h_bdof_global.size();
        std::vector<int32_t> hrow(mine.size());
        std::vector<double> hx(mine.size()), hy(mine.size());
        for (size_t k = 0; k < mine.size(); ++k) {
            hrow[k] = (int32_t)(mine[k].idof - L.row0);  // storage row
            hx[k] = mine[k].x;
            hy[k] = mine[k].y;
        }
        RET(dev_alloc(ctx, &ctx->brow, mine.size(), false));
        RET(dev_alloc(ctx, &ctx->bx, mine.size(), false));
        RET(dev_alloc(ctx, &ctx->by, mine.size(), false));
        if (!mine.empty()) {
            CK(cudaMemcpyAsync(ctx->brow, hrow.data(), sizeof(int32_t) * hrow.size(), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->bx, hx.data(), sizeof(double) * hx.size(), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->by, hy.data(), sizeof(double) * hy.size(), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }

    // ---- vectors, reduction scratch, CG state ----------------------------------------------------
    RET(dev_alloc(ctx, &ctx->u, (size_t)L.nloc));
    RET(dev_alloc(ctx, &ctx->v, (size_t)L.nloc));
    RET(dev_alloc(ctx, &ctx->a, (size_t)L.nloc));
    RET(dev_alloc(ctx, &ctx->unew, (size_t)L.nloc));
    RET(dev_alloc(ctx, &ctx->d, (size_t)L.nloc));
    RET(dev_alloc(ctx, &ctx->rhs, (size_t)L.nown));
    RET(dev_alloc(ctx, &ctx->fvec, (size_t)L.nown));
    RET(dev_alloc(ctx, &ctx->g, (size_t)L.nown));
    RET(dev_alloc(ctx, &ctx->h, (size_t)L.nown));
    {
        const int64_t cells = 2LL * (L.jq1 - L.jq0) * L.mesh.nx;
        const int64_t blocks = std::max<int64_t>({(cells + 127) / 128, (int64_t)(nslots / kSlice + 7) / 8,
                                                  (int64_t)reduction_blocks(L.nown)}) + 1;
        RET(dev_alloc(ctx, &ctx->partials, (size_t)blocks * 4));
        RET(dev_alloc(ctx, &ctx->partials2, (size_t)blocks * 4));
        RET(dev_alloc(ctx, &ctx->partials3, (size_t)blocks * 4));
    }
    RET(dev_alloc(ctx, &ctx->counter, 4));
    RET(dev_alloc(ctx, &ctx->S, 1));
    RET(dev_alloc(ctx, &ctx->res, 8));
    RET(dev_alloc(ctx, &ctx->d0, 2));
    CK(cudaMallocHost((void **)&ctx->hS, sizeof(CgScalars)));
    CK(cudaMallocHost((void **)&ctx->hres, 8 * sizeof(double)));
    RET(upload_cg_control(ctx));

    // ---- separable forcing F = T(t) S(x,y): assemble the load vector of S once ---------------------
    if (ctx->forcing_active && !(ctx->cfg.flags & WAVE_FLAG_FORCING_EVERY_STEP) && !ctx->f_text[0].empty()) {
        try {
            ctx->forcing_separable =
                compile_separable(ctx->f_text[0], ctx->f_text[1], ctx->f_text[2], &ctx->f_time, &ctx->f_space);
        } catch (const std::exception &) {
            ctx->forcing_separable = false;
        }
        if (ctx->forcing_separable) {
            DevTmp<Program> dS;
            DevTmp<double> cells;
            RET(dev_alloc(ctx, &dS.p, 1, false));
            RET(dev_alloc(ctx, &cells.p, (size_t)forcing_cells(L) * dofs_per_cell(L.mesh.r), false));
            CK(cudaMemcpyAsync(dS.p, &ctx->f_space, sizeof(Program), cudaMemcpyHostToDevice, ctx->stream));
            launch_forcing(l, L, dS.p, &ctx->q_asm, 0.0, 0.0, 1.0, 0.0, 0, cells.p, ctx->fvec);
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    if (ctx->forcing_active && !ctx->forcing_separable)
        RET(dev_alloc(ctx, &ctx->cellvec, (size_t)forcing_cells(L) * dofs_per_cell(L.mesh.r), false));

    // ---- scheme matrices (src/WaveNewmark.cpp:110-112, :372-374; src/WaveTheta.cpp:110-115) -----
    RET(dev_alloc(ctx, &ctx->S1, (size_t)ctx->nnz_pad, false));
    RET(dev_alloc(ctx, &ctx->dinv1, (size_t)L.nown, false));
    const double dt = ctx->cfg.dt;
    if (ctx->cfg.scheme == WAVE_SCHEME_NEWMARK) {
        // SYS1 starts as the BC-modified mass matrix of the a^0 solve; wave_init switches it to
        // bc(M + beta dt^2 K) afterwards
        RET(build_system_matrix(ctx, 0.0, ctx->S1, ctx->dinv1, ctx->d0));
    } else {
        const double th = ctx->cfg.theta;
        RET(build_system_matrix(ctx, (th * dt) * (th * dt), ctx->S1, ctx->dinv1, ctx->d0));
        RET(dev_alloc(ctx, &ctx->S2, (size_t)ctx->nnz_pad, false));
        RET(dev_alloc(ctx, &ctx->dinv2, (size_t)L.nown, false));
        RET(build_system_matrix(ctx, 0.0, ctx->S2, ctx->dinv2, ctx->d0 + 1));
    }
    RET(sync_check(ctx));
    RET(setup_peer_exchange(ctx));
    RET(fused_plan(ctx));
    ctx->is_setup = true;
    return WAVE_OK;
}

int wave_init(wave_ctx *ctx) {
    if (!ctx) return WAVE_ERR_ARG;
    if (!ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_init before wave_setup");
    const Layout &L = ctx->L;
    const Launcher &l = ctx->launcher;
    launch_interpolate(l, L, ctx->dprog + WAVE_EXPR_U0, 0.0, ctx->u, nullptr, nullptr);
    launch_interpolate(l, L, ctx->dprog + WAVE_EXPR_V0, 0.0, ctx->v, nullptr, nullptr);
    if (ctx->cfg.scheme == WAVE_SCHEME_NEWMARK) {
        const double dt = ctx->cfg.dt, beta = ctx->cfg.beta;
        if (ctx->is_init)  // SYS1 must again be the BC-modified mass matrix
            RET(build_system_matrix(ctx, 0.0, ctx->S1, ctx->dinv1, ctx->d0));
        // M a0 = F(0) - K u0 with a0 = second difference of g on the boundary (src/WaveNewmark.cpp:300-385)
        double fscale = 1.0;
        if (ctx->forcing_active) RET(compute_forcing(ctx, 0.0, 0.0, 1.0, 0.0, 0, &fscale));
        SpmvArgs a = spmv_base(ctx);
        a.t[0] = {ctx->K, ctx->u, nullptr, 1.0, 0.0, -1.0};
        if (ctx->forcing_active) { a.add0 = ctx->fvec; a.addc0 = fscale; }
        a.y = ctx->rhs;
        spmv(ctx, l, a);
        launch_fill(l, L.nloc, 0.0, ctx->a);
        launch_bc_values(l, BC_SECOND_DIFF, ctx->nb, ctx->brow, ctx->bx, ctx->by, ctx->dprog + WAVE_EXPR_G, dt, dt, 0.0,
                         nullptr, ctx->a + L.own_off, ctx->rhs, ctx->d0);
        int its = 0;
        ctx->prev_its[0] = 0;
        RET(cg_solve(ctx, ctx->S1, ctx->dinv1, ctx->a, ctx->rhs, 0, &its));
        RET(halo_exchange(ctx, ctx->a));
        RET(build_system_matrix(ctx, beta * dt * dt, ctx->S1, ctx->dinv1, ctx->d0));
        ctx->prev_its[0] = 0;
    }
    if (ctx->cfg.precond == WAVE_PRECOND_MG) {
        const double dt = ctx->cfg.dt;
        const double s = ctx->cfg.scheme == WAVE_SCHEME_NEWMARK ? ctx->cfg.beta * dt * dt
                                                                : (ctx->cfg.theta * dt) * (ctx->cfg.theta * dt);
        if (s > 0.0) RET(mg_setup(ctx, s));
    }
    RET(sync_check(ctx));
    ctx->is_init = true;
    return WAVE_OK;
}

int wave_step(wave_ctx *ctx, double t_np1, int32_t iters[2], double norms[2]) {
    if (!ctx) return WAVE_ERR_ARG;
    if (!ctx->is_init) return fail(ctx, WAVE_ERR_STATE, "wave_step before wave_init");
    return ctx->cfg.scheme == WAVE_SCHEME_NEWMARK ? newmark_step(ctx, t_np1, iters, norms)
                                                  : theta_step(ctx, t_np1, iters, norms);
}

int wave_run(wave_ctx *ctx, double t_start, int32_t n_steps, double *t_end, int32_t *steps_done, int32_t iters[2],
             double norms[2], int64_t *total_iters) {
    if (!ctx) return WAVE_ERR_ARG;
    if (!ctx->is_init) return fail(ctx, WAVE_ERR_STATE, "wave_run before wave_init");
    double t = t_start, nrm[2] = {0, 0};
    int32_t it[2] = {0, 0};
    int64_t tot = 0;
    int done = 0, rc = WAVE_OK;
    for (int s = 0; s < n_steps; ++s) {
        t += ctx->cfg.dt;  // src/WaveNewmark.cpp:409
        ++done;
        rc = wave_step(ctx, t, it, nrm);
        if (rc != WAVE_OK) break;
        tot += it[0] + it[1];
        // check_divergence, threshold 1e130 (src/WaveEquationBase.cpp:425-431, src/WaveNewmark.cpp:399)
        if (!std::isfinite(nrm[0]) || !std::isfinite(nrm[1]) || nrm[0] > 1e130 || nrm[1] > 1e130) {
            rc = fail(ctx, WAVE_ERR_DIVERGED, "divergence detected");
            break;
        }
    }
    if (t_end) *t_end = t;
    if (steps_done) *steps_done = done;
    if (iters) { iters[0] = it[0]; iters[1] = it[1]; }
    if (norms) { norms[0] = nrm[0]; norms[1] = nrm[1]; }
    if (total_iters) *total_iters = tot;
    return rc;
}

int wave_set_vector(wave_ctx *ctx, int which, const double *host, size_t n) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_set_vector before wave_setup");
    double *dst = vec_ptr(ctx, which);
    if (!dst || !host || (int64_t)n != n_dofs(ctx->L.mesh)) return fail(ctx, WAVE_ERR_ARG, "bad vector id or size");
    // every rank takes its local slice (ghosts included) straight from the canonical host array
    RET(upload_local(ctx, host, dst));
    CK(cudaStreamSynchronize(ctx->stream));
    return WAVE_OK;
}

int wave_get_vector(wave_ctx *ctx, int which, double *host, size_t n) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_get_vector before wave_setup");
    const Layout &L = ctx->L;
    const double *src = which == WAVE_VEC_RHS ? ctx->rhs : (vec_ptr(ctx, which) ? vec_ptr(ctx, which) + L.own_off : nullptr);
    if (!src || !host) return fail(ctx, WAVE_ERR_ARG, "bad vector id");
    if ((int64_t)n == (int64_t)L.nown && ctx->cfg.nranks > 1) {
        RET(own_to_canonical(ctx, src, ctx->tmp));
        CK(cudaMemcpyAsync(host, ctx->tmp, sizeof(double) * L.nown, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return WAVE_OK;
    }
    if ((int64_t)n != n_dofs(L.mesh)) return fail(ctx, WAVE_ERR_ARG, "bad vector size");
    if (ctx->cfg.nranks == 1) {
        RET(own_to_canonical(ctx, src, ctx->tmp));
        CK(cudaMemcpyAsync(host, ctx->tmp, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        RET(gather_global(ctx, src));
        CK(cudaMemcpyAsync(host, ctx->scratch, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return WAVE_OK;
}

int wave_step_host(wave_ctx *ctx, double t_np1, double *u, double *v, double *a, int32_t iters[2], double norms[2]) {
    if (!ctx || !ctx->is_init) return fail(ctx, WAVE_ERR_STATE, "wave_step_host before wave_init");
    const size_t n = (size_t)n_dofs(ctx->L.mesh);
    const Layout &L = ctx->L;
    RET(upload_local(ctx, u, ctx->u));
    RET(upload_local(ctx, v, ctx->v));
    if (ctx->cfg.scheme == WAVE_SCHEME_NEWMARK) {
        if (!a) return fail(ctx, WAVE_ERR_ARG, "Newmark needs the acceleration vector");
        RET(upload_local(ctx, a, ctx->a));
    }
    RET(wave_step(ctx, t_np1, iters, norms));
    // three downloads back to back, one synchronisation.  With several ranks every rank refreshes its
    // owned rows of the caller's arrays (the ghost parts are re-exchanged on the device each step).
    RET(download_own_async(ctx, ctx->u + L.own_off, u + L.row0));
    RET(download_own_async(ctx, ctx->v + L.own_off, v + L.row0));
    if (ctx->cfg.scheme == WAVE_SCHEME_NEWMARK) RET(download_own_async(ctx, ctx->a + L.own_off, a + L.row0));
    CK(cudaStreamSynchronize(ctx->stream));
    (void)n;
    return WAVE_OK;
}

int wave_norms(wave_ctx *ctx, double out[2]) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_norms before wave_setup");
    const Layout &L = ctx->L;
    launch_norms2(ctx->launcher, L.nown, ctx->u + L.own_off, ctx->v + L.own_off, ctx->partials, ctx->counter, ctx->res);
    return finish_norms(ctx, out);
}

int wave_energy(wave_ctx *ctx, double *out) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_energy before wave_setup");
    const Layout &L = ctx->L;
    PhaseTimer pt(ctx, PH_ENERGY);
    RET(halo_exchange(ctx, ctx->u));
    RET(halo_exchange(ctx, ctx->v));
    SpmvArgs a = spmv_base(ctx);
    a.t[0] = {ctx->K, ctx->u, nullptr, 1.0, 0.0, 1.0};
    a.dot_mode = 1; a.dotv = ctx->u + L.own_off; a.result = ctx->res + 2;
    spmv(ctx, ctx->launcher, a);
    SpmvArgs b = spmv_base(ctx);
    b.t[0] = {ctx->M, ctx->v, nullptr, 1.0, 0.0, 1.0};
    b.dot_mode = 1; b.dotv = ctx->v + L.own_off; b.result = ctx->res + 3;
    spmv(ctx, ctx->launcher, b);
    RET(allreduce(ctx, ctx->res + 2, 2));
    CK(cudaMemcpyAsync(ctx->hres + 2, ctx->res + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out = 0.5 * (ctx->hres[3] + ctx->hres[2]);  // 0.5 * (v.Mv + u.Ku), src/WaveEquationBase.cpp:154
    return WAVE_OK;
}

int wave_errors(wave_ctx *ctx, double t, double out[4]) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_errors before wave_setup");
    if (!ctx->has[WAVE_EXPR_SOLUTION]) return fail(ctx, WAVE_ERR_STATE, "no exact Solution expression");
    RET(halo_exchange(ctx, ctx->u));
    launch_errors(ctx->launcher, ctx->L, ctx->dprog + WAVE_EXPR_SOLUTION, &ctx->q_err, t, ctx->u, ctx->partials,
                  ctx->counter, ctx->res + 4);
    RET(allreduce(ctx, ctx->res + 4, 4));
    CK(cudaMemcpyAsync(ctx->hres + 4, ctx->res + 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const double l2 = std::sqrt(ctx->hres[4]), h1 = std::sqrt(ctx->hres[5]);
    const double nl2 = std::sqrt(ctx->hres[6]), nh1 = std::sqrt(ctx->hres[7]);
    out[0] = l2;
    out[1] = h1;
    out[2] = nl2 < 1e-14 ? l2 : l2 / nl2;  // src/WaveEquationBase.cpp:419-422
    out[3] = nh1 < 1e-14 ? h1 : h1 / nh1;
    return WAVE_OK;
}

int wave_probe(wave_ctx *ctx, double x, double y, double *out) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_probe before wave_setup");
    RET(halo_exchange(ctx, ctx->u));
    launch_probe(ctx->launcher, ctx->L, x, y, ctx->u, ctx->res);
    RET(allreduce(ctx, ctx->res, 1));
    CK(cudaMemcpyAsync(ctx->hres, ctx->res, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out = ctx->hres[0];
    return WAVE_OK;
}

int64_t wave_n_dofs(const wave_ctx *ctx) { return ctx ? n_dofs(ctx->L.mesh) : 0; }
int64_t wave_n_cells(const wave_ctx *ctx) { return ctx ? n_cells(ctx->L.mesh) : 0; }
int64_t wave_local_rows(const wave_ctx *ctx, int64_t *first_row) {
    if (!ctx) return 0;
    if (first_row) *first_row = ctx->L.row0;
    return ctx->L.nown;
}
int64_t wave_local_nnz(const wave_ctx *ctx) { return ctx ? ctx->nnz : 0; }
int64_t wave_nnz(const wave_ctx *cctx) {
    if (!cctx) return 0;
    if (cctx->cfg.nranks == 1) return cctx->nnz;
    // sum of the ranks' owned entries (collective; exact in a double up to 2^53)
    wave_ctx *ctx = const_cast<wave_ctx *>(cctx);
    ctx->hres[0] = (double)ctx->nnz;
    if (cudaMemcpyAsync(ctx->res, ctx->hres, sizeof(double), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return -1;
    if (allreduce(ctx, ctx->res, 1) != WAVE_OK) return -1;
    if (cudaMemcpyAsync(ctx->hres, ctx->res, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    return (int64_t)ctx->hres[0];
}

int wave_get_csr(wave_ctx *ctx, int which, int64_t *rowptr, int32_t *col, double *val) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_get_csr before wave_setup");
    const Layout &L = ctx->L;
    const double *src = mat_ptr(ctx, which);
    if (val && !src) return fail(ctx, WAVE_ERR_ARG, "matrix not available for this scheme");
    // canonical row pointer: lengths of the storage rows in canonical row order, then a scan
    DevTmp<uint32_t> len_c, rp_c;
    DevTmp<char> scan_tmp;
    RET(dev_alloc(ctx, &len_c.p, (size_t)L.nown + 1));
    RET(dev_alloc(ctx, &rp_c.p, (size_t)L.nown + 1));
    launch_canonical_lengths(ctx->launcher, L, ctx->A, ctx->c2i, len_c.p);
    {
        size_t bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, len_c.p, rp_c.p, L.nown + 1, ctx->stream));
        CK(cudaMalloc((void **)&scan_tmp.p, bytes));
        CK(cub::DeviceScan::ExclusiveSum(scan_tmp.p, bytes, len_c.p, rp_c.p, L.nown + 1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (rowptr) {
        std::vector<uint32_t> rp((size_t)L.nown + 1);
        CK(cudaMemcpy(rp.data(), rp_c.p, sizeof(uint32_t) * rp.size(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < rp.size(); ++i) rowptr[i] = rp[i];
    }
    if (col || val) {
        DevTmp<double> dval;
        DevTmp<int32_t> dcol;
        if (val) RET(dev_alloc(ctx, &dval.p, (size_t)ctx->nnz, false));
        if (col) RET(dev_alloc(ctx, &dcol.p, (size_t)ctx->nnz, false));
        launch_export_csr(ctx->launcher, L, ctx->A, ctx->c2i, ctx->i2c, rp_c.p, src ? src : ctx->M, dval.p, dcol.p);
        if (val) CK(cudaMemcpyAsync(val, dval.p, sizeof(double) * ctx->nnz, cudaMemcpyDeviceToHost, ctx->stream));
        if (col) CK(cudaMemcpyAsync(col, dcol.p, sizeof(int32_t) * ctx->nnz, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return launch_check(ctx);
}

int wave_get_support_points(wave_ctx *ctx, double *x, double *y, size_t n) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "before wave_setup");
    if (ctx->cfg.nranks != 1 || (int64_t)n != n_dofs(ctx->L.mesh))
        return fail(ctx, WAVE_ERR_ARG, "support points: single rank, n = n_dofs");
    DevTmp<double> dx, dy;
    RET(dev_alloc(ctx, &dx.p, n));
    RET(dev_alloc(ctx, &dy.p, n));
    launch_interpolate(ctx->launcher, ctx->L, ctx->dprog, 0.0, nullptr, dx.p, dy.p);
    CK(cudaMemcpyAsync(x, dx.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(y, dy.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return launch_check(ctx);
}

int64_t wave_n_boundary_dofs(const wave_ctx *ctx) { return ctx ? ctx->nb_global : 0; }
int wave_get_boundary_dofs(wave_ctx *ctx, int32_t *out, size_t n) {
    if (!ctx || !ctx->is_setup || n != ctx->h_bdof_global.size()) return fail(ctx, WAVE_ERR_ARG, "bad size");
    std::memcpy(out, ctx->h_bdof_global.data(), sizeof(int32_t) * n);
    return WAVE_OK;
}

int wave_spmv(wave_ctx *ctx, int which, const double *x, double *y, size_t n) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_spmv before wave_setup");
    const double *val = mat_ptr(ctx, which);
    if (!val || ctx->cfg.nranks != 1 || (int64_t)n != n_dofs(ctx->L.mesh))
        return fail(ctx, WAVE_ERR_ARG, "wave_spmv: single rank, n = n_dofs, valid matrix id");
    RET(upload_local(ctx, x, ctx->d));
    SpmvArgs a = spmv_base(ctx);
    a.t[0] = {val, ctx->d, nullptr, 1.0, 0.0, 1.0};
    a.y = ctx->h;
    spmv(ctx, ctx->launcher, a);
    RET(own_to_canonical(ctx, ctx->h, ctx->tmp));
    CK(cudaMemcpyAsync(y, ctx->tmp, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_check(ctx);
}

int wave_cg(wave_ctx *ctx, int which, double *x, const double *b, size_t n, int32_t *iters) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_cg before wave_setup");
    const double *val = mat_ptr(ctx, which);
    if (!val || ctx->cfg.nranks != 1 || (int64_t)n != n_dofs(ctx->L.mesh) ||
        (which != WAVE_MAT_SYS1 && which != WAVE_MAT_SYS2))
        return fail(ctx, WAVE_ERR_ARG, "wave_cg: single rank, n = n_dofs, SYS1 or SYS2");
    const double *dinv = which == WAVE_MAT_SYS1 ? ctx->dinv1 : ctx->dinv2;
    RET(upload_local(ctx, x, ctx->unew));
    RET(upload_local(ctx, b, ctx->g));  // single rank: nloc == nown; g is scratch until the solve starts
    CK(cudaMemcpyAsync(ctx->rhs, ctx->g, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    int its = 0;
    ctx->prev_its[0] = 0;
    const int rc = cg_solve(ctx, val, dinv, ctx->unew, ctx->rhs, 0, &its);
    ctx->prev_its[0] = 0;
    if (iters) *iters = its;
    RET(own_to_canonical(ctx, ctx->unew, ctx->tmp));
    CK(cudaMemcpyAsync(x, ctx->tmp, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return rc;
}

static int ensure_flush(wave_ctx *ctx) {
    if (ctx->flush_buf) return WAVE_OK;
    ctx->flush_n = (256LL << 20) / 8;  // 256 MiB > 126 MB L2
    CK(cudaMalloc(&ctx->flush_buf, sizeof(double) * ctx->flush_n));
    CK(cudaMemsetAsync(ctx->flush_buf, 0, sizeof(double) * ctx->flush_n, ctx->stream));
    return WAVE_OK;
}

int wave_bench_spmv(wave_ctx *ctx, int which, int reps, int flush_l2, double *ms_avg, double *bytes) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_bench_spmv before wave_setup");
    const double *val = mat_ptr(ctx, which);
    if (!val || reps < 1) return fail(ctx, WAVE_ERR_ARG, "bad matrix id / reps");
    if (flush_l2) RET(ensure_flush(ctx));
    // the launch of the CG iteration: y = A x with the fused x . y (per-block partials left for a consumer)
    SpmvArgs a = spmv_base(ctx);
    a.t[0] = {val, ctx->u, nullptr, 1.0, 0.0, 1.0};
    a.y = ctx->h;
    a.dot_mode = 1;
    a.dotv = ctx->u + ctx->L.own_off;
    a.dot_publish = 1;
    cudaEvent_t e0 = ctx->ev[2 * PH_COUNT + 2], e1 = ctx->ev[2 * PH_COUNT + 3];
    for (int w = 0; w < 3; ++w) spmv(ctx, ctx->launcher, a);
    CK(cudaStreamSynchronize(ctx->stream));
    double total = 0.0;
    if (flush_l2) {
        for (int r = 0; r < reps; ++r) {
            launch_flush_l2(ctx->launcher, ctx->flush_buf, ctx->flush_n);
            CK(cudaEventRecord(e0, ctx->stream));
            spmv(ctx, ctx->launcher, a);
            CK(cudaEventRecord(e1, ctx->stream));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            total += ms;
        }
    } else {
        CK(cudaEventRecord(e0, ctx->stream));
        for (int r = 0; r < reps; ++r) spmv(ctx, ctx->launcher, a);
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        total = ms;
    }
    if (ms_avg) *ms_avg = total / reps;
    if (bytes) *bytes = spmv_bytes(ctx);
    return WAVE_OK;
}

int wave_bench_cg_iter(wave_ctx *ctx, int which, int reps, double *ms_avg, double *bytes) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_bench_cg_iter before wave_setup");
    const double *val = mat_ptr(ctx, which);
    if (!val || (which != WAVE_MAT_SYS1 && which != WAVE_MAT_SYS2) || reps < 1)
        return fail(ctx, WAVE_ERR_ARG, "bad matrix id / reps");
    const double *dinv = which == WAVE_MAT_SYS1 ? ctx->dinv1 : ctx->dinv2;
    const Layout &L = ctx->L;
    // right-hand side b = A * 1 on interior-compatible data: start from x = 0 each repetition
    launch_fill(ctx->launcher, L.nloc, 1.0, ctx->d);
    SpmvArgs a = spmv_base(ctx);
    a.t[0] = {val, ctx->d, nullptr, 1.0, 0.0, 1.0};
    a.y = ctx->rhs;
    spmv(ctx, ctx->launcher, a);
    launch_zero_rows(ctx->launcher, ctx->nb, ctx->brow, ctx->rhs);  // x_B = 0 must satisfy the Dirichlet rows
    double ms_total = 0.0, its_total = 0.0;
    for (int r = 0; r < reps; ++r) {
        launch_fill(ctx->launcher, L.nloc, 0.0, ctx->unew);
        const double before_ms = ctx->cg_stats[3], before_it = ctx->cg_stats[1];
        int its = 0;
        ctx->prev_its[0] = 0;
        RET(cg_solve(ctx, val, dinv, ctx->unew, ctx->rhs, 0, &its));
        ms_total += ctx->cg_stats[3] - before_ms;
        its_total += ctx->cg_stats[1] - before_it;
    }
    ctx->prev_its[0] = 0;
    if (ms_avg) *ms_avg = its_total > 0 ? ms_total / its_total : 0.0;
    if (bytes) *bytes = spmv_bytes(ctx) + 80.0 * (double)L.nown;
    return WAVE_OK;
}

int64_t wave_launch_count(const wave_ctx *ctx) { return ctx ? ctx->launches : 0; }

int wave_timers_enable(wave_ctx *ctx, int on) {
    if (!ctx) return WAVE_ERR_ARG;
    ctx->timers_on = on != 0;
    return WAVE_OK;
}
int wave_timers(wave_ctx *ctx, double out_ms[6], int reset) {
    if (!ctx) return WAVE_ERR_ARG;
    for (int k = 0; k < PH_COUNT; ++k) {
        out_ms[k] = ctx->phase_ms[k];
        if (reset) ctx->phase_ms[k] = 0.0;
    }
    return WAVE_OK;
}
int wave_kernel_timing(wave_ctx *ctx, int on, double launches[3], double ms_total[3]) {
    if (!ctx) return WAVE_ERR_ARG;
    drain_spmv_events(ctx);
    for (int k = 0; k < 3; ++k) {
        if (launches) launches[k] = ctx->kt_count[k];
        if (ms_total) ms_total[k] = ctx->kt_ms[k];
        ctx->kt_count[k] = 0.0;
        ctx->kt_ms[k] = 0.0;
    }
    if (on && ctx->spmv_ev.empty()) {
        ctx->spmv_ev.resize(16384, nullptr);
        ctx->spmv_ev_tag.resize(ctx->spmv_ev.size() / 2, 0);
        for (auto &e : ctx->spmv_ev) CK(cudaEventCreate(&e));
    }
    ctx->spmv_timing = on != 0;
    return WAVE_OK;
}
int wave_spmv_timing(wave_ctx *ctx, int on, double *launches, double *ms_total) {
    double n[3] = {0, 0, 0}, ms[3] = {0, 0, 0};
    const int rc = wave_kernel_timing(ctx, on, n, ms);
    if (launches) *launches = n[0];
    if (ms_total) *ms_total = ms[0];
    return rc;
}
int wave_cg_stats(wave_ctx *ctx, double out[4], int reset) {
    if (!ctx) return WAVE_ERR_ARG;
    for (int k = 0; k < 4; ++k) {
        out[k] = ctx->cg_stats[k];
        if (reset) ctx->cg_stats[k] = 0.0;
    }
    return WAVE_OK;
}

int wave_operator_info(const wave_ctx *ctx, int64_t out[4]) {
    if (!ctx || !out) return WAVE_ERR_ARG;
    out[0] = ctx->stencil_rows;
    out[1] = ctx->L.nown - ctx->stencil_rows;
    out[2] = ctx->sell_nnz;
    out[3] = (int64_t)spmv_bytes(ctx);
    return WAVE_OK;
}

int wave_cg_fused_active(const wave_ctx *ctx) { return ctx && ctx->fused.ok ? 1 : 0; }

int wave_cell_dofs(int32_t nx, int32_t ny, int32_t r, int64_t cell, int32_t *out) {
    if (nx < 1 || ny < 1 || (r != 1 && r != 2) || cell < 0 || cell >= 2LL * nx * ny || !out) return WAVE_ERR_ARG;
    Mesh m{};
    m.nx = nx; m.ny = ny; m.r = r;
    int64_t d[6];
    cell_dofs(m, cell, d);
    for (int k = 0; k < dofs_per_cell(r); ++k) out[k] = (int32_t)d[k];
    return WAVE_OK;
}

int wave_cell_dofs_storage(int32_t nx, int32_t ny, int32_t r, int64_t cell, int32_t *out) {
    if (nx < 1 || ny < 1 || (r != 1 && r != 2) || cell < 0 || cell >= 2LL * nx * ny || !out) return WAVE_ERR_ARG;
    Mesh m{};
    m.nx = nx; m.ny = ny; m.r = r;
    int64_t d[6];
    cell_dofs_internal(m, cell, d);
    for (int k = 0; k < dofs_per_cell(r); ++k) out[k] = (int32_t)d[k];
    return WAVE_OK;
}

int wave_quadrature(int32_t n_points_1d, double *xi, double *eta, double *w) {
    if (!xi || !eta || !w) return WAVE_ERR_ARG;
    try {
        const Quadrature q = make_quadrature(n_points_1d);
        for (int k = 0; k < q.nq; ++k) { xi[k] = q.xi[k]; eta[k] = q.eta[k]; w[k] = q.w[k]; }
        return q.nq;
    } catch (const std::exception &) {
        return WAVE_ERR_ARG;
    }
}

int wave_mg_plan(const wave_config *cfg, double s, double c0, int32_t *nx_out, int32_t *ny_out, int32_t max_levels) {
    if (!cfg || !nx_out || !ny_out || max_levels < 1 || cfg->nx < 1 || cfg->ny < 1 || cfg->nranks < 1) return WAVE_ERR_ARG;
    MgPlanLevel plan[12];
    const int n = mg_plan(*cfg, s, c0, plan, std::min<int>(max_levels, 11));
    for (int k = 0; k < n; ++k) { nx_out[k] = plan[k].nx; ny_out[k] = plan[k].ny; }
    return n;
}

int wave_partition_plan(int32_t nx, int32_t ny, int32_t r, int32_t rank, int32_t nranks, wave_partition *out) {
    if (nx < 1 || ny < 1 || (r != 1 && r != 2) || nranks < 1 || rank < 0 || rank >= nranks || ny < nranks || !out)
        return WAVE_ERR_ARG;
    Mesh m{};
    m.nx = nx; m.ny = ny; m.r = r;
    const Layout L = make_layout(m, rank, nranks);
    out->quad_row_begin = L.jq0;
    out->quad_row_end = L.jq1;
    out->row_begin = L.row0;
    out->row_end = L.row0 + L.nown;
    out->ghost_lo_begin = L.col0;
    out->ghost_hi_end = L.col0 + L.nloc;
    return WAVE_OK;
}

}  // extern "C"
