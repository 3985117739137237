/*
 * wavegpu.h -- C ABI of libwavegpu.so: the B200-native (sm_100a CUDA + NCCL) replacement of the
 * time-stepping hot path of nmpde-wave-equation.
 *
 * The reference has no plugin/FFI seam; every numerical call goes from the WaveNewmark /
 * WaveTheta / WaveEquationBase classes straight into deal.II + Trilinos.  This header is the
 * seam a maintainer would cut: each entry point names the reference code it replaces
 * (paths relative to the reference root).  The C++ host classes in
 * nmpde-wave-equation_b200/host/ (same names and constructor arguments as the reference's)
 * call only these functions; INTEGRATION.md shows the binding.
 *
 * Conventions: plain C types only; opaque context; every function returns WAVE_OK (0) or a
 * negative wave_status, with a message available from wave_last_error().  The context owns all
 * device memory, streams and the NCCL communicator; host buffers are caller-owned.  One context
 * is driven by one host thread.  There is no CPU fallback: without a CUDA device wave_create
 * fails with WAVE_ERR_CUDA.
 */
#ifndef WAVEGPU_H
#define WAVEGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wave_ctx wave_ctx;

typedef enum {
    WAVE_OK = 0,
    WAVE_ERR_ARG = -1,       /* invalid argument / configuration                         */
    WAVE_ERR_EXPR = -2,      /* expression did not parse (std::invalid_argument upstream) */
    WAVE_ERR_CUDA = -3,      /* CUDA runtime / driver / NCCL failure, or no device        */
    WAVE_ERR_STATE = -4,     /* call order violated (e.g. step before init)               */
    WAVE_ERR_NOCONV = -5,    /* CG hit max iterations (deal.II SolverControl::NoConvergence) */
    WAVE_ERR_DIVERGED = -6,  /* norms non-finite or above threshold (check_divergence)    */
    WAVE_ERR_UNSUPPORTED = -7
} wave_status;

/* scheme ids: which class the context stands behind */
enum { WAVE_SCHEME_NEWMARK = 0, WAVE_SCHEME_THETA = 1 };

/* expression slots: the seven FunctionParser objects of src/main-newmark.cpp:64-73 */
enum { WAVE_EXPR_C = 0, WAVE_EXPR_F, WAVE_EXPR_U0, WAVE_EXPR_V0, WAVE_EXPR_G, WAVE_EXPR_DGDT,
       WAVE_EXPR_SOLUTION, WAVE_EXPR_COUNT };

/* vector ids for wave_get_vector / wave_set_vector */
enum { WAVE_VEC_U = 0, WAVE_VEC_V = 1, WAVE_VEC_A = 2, WAVE_VEC_RHS = 3 };

/* matrix ids for wave_get_csr / wave_spmv / wave_cg:
   M = mass_matrix, K = stiffness_matrix (include/WaveEquationBase.hpp),
   SYS1 = BC-modified matrix_a (Newmark) / matrix_u (theta), SYS2 = BC-modified matrix_v (theta),
   SYS0 = BC-modified mass matrix of the a^0 solve (src/WaveNewmark.cpp:372-374). */
enum { WAVE_MAT_M = 0, WAVE_MAT_K = 1, WAVE_MAT_SYS1 = 2, WAVE_MAT_SYS2 = 3 };

/* preconditioners (north star: Jacobi / SSOR-equivalent in place of Trilinos ML AMG,
   src/WaveNewmark.cpp:246-251) */
/* WAVE_PRECOND_MG: geometric multigrid V-cycle ([P2 ->] P1 -> P1 on Nel/2, Nel/4, ...; rediscretised
   coarse operators, damped-Jacobi smoothing), the structured-mesh counterpart of the reference's
   Trilinos ML AMG with Chebyshev smoothing; used for the solves with the stiffness term (matrix_a /
   matrix_u), Jacobi for the mass-only solves; one rank or strips over several ranks (wave_mg_plan). */
enum { WAVE_PRECOND_JACOBI = 0, WAVE_PRECOND_NONE = 1, WAVE_PRECOND_MG = 2 };

typedef struct {
    /* mesh + FE: WaveEquationBase ctor arguments N_el, geometry, r (include/WaveEquationBase.hpp:72-95) */
    int32_t nx, ny;
    double x0, x1, y0, y1;
    int32_t r; /* 1 or 2 */
    /* scheme: WaveNewmark(gamma, beta, delta_t) include/WaveNewmark.hpp:62-84,
               WaveTheta(theta, delta_t)         include/WaveTheta.hpp:69-89 */
    int32_t scheme;
    double dt, theta, beta, gamma;
    /* SolverCG control: ReductionControl(10000, 1e-12, 1e-6), src/WaveNewmark.cpp:256.
       Zero / negative values select those defaults. */
    int32_t cg_maxit;
    double cg_tol, cg_reduce;
    int32_t precond;
    /* strip partition over ranks (one process per GPU); rank 0 of 1 = single GPU */
    int32_t rank, nranks;
    int32_t device;             /* CUDA ordinal, -1 = current device */
    const void *nccl_unique_id; /* 128 bytes from wave_comm_unique_id (rank 0), NULL if nranks==1 */
    uint32_t flags;             /* WAVE_FLAG_* */
    void *stream;               /* cudaStream_t to run on (caller-owned), NULL = private stream */
} wave_config;

#define WAVE_FLAG_FORCING_EVERY_STEP 1u /* assemble F even when it folds to 0 (as the reference) */
#define WAVE_FLAG_NO_STENCIL 2u         /* keep every row in SELL form: no table-driven (matrix-free) rows */

/* ---- life cycle ---------------------------------------------------------------------- */
/* Fills a config with the reference's declared defaults (src/ParameterReader.cpp:39-105). */
void wave_default_config(wave_config *cfg);
/* Replaces the constructor of WaveNewmark / WaveTheta. */
int wave_create(const wave_config *cfg, wave_ctx **out);
void wave_destroy(wave_ctx *ctx);
/* Message of the last failure on this context (ctx may be NULL for wave_create failures). */
const char *wave_last_error(const wave_ctx *ctx);
/* CUDA devices visible to this process (0 without a driver or device); launchers size -np with it
   and the host classes pick device local_rank % count. */
int wave_device_count(void);
/* 128-byte NCCL unique id to broadcast to all ranks before wave_create (rank 0 calls it). */
int wave_comm_unique_id(void *out128);

/* Replaces ParameterReader::load_functions -> FunctionParser::initialize
   (src/ParameterReader.cpp:139-175): expression text, "x, y[, t]" variable list,
   "k=v, ..." constants with pi / n*pi values.  `pi` is always defined. */
int wave_set_expr(wave_ctx *ctx, int which, const char *expression, const char *variable_names,
                  const char *constants);
/* Evaluate one compiled expression on the host side of the ABI (parser self-test). */
int wave_eval_expr(wave_ctx *ctx, int which, double x, double y, double t, double *out);

/* Context-free expression handles for the host classes (FunctionParser::value on the host side:
   boundary tables, convergence.csv bookkeeping).  Host only, no device needed. */
typedef struct wave_expr wave_expr;
int wave_expr_create(const char *expression, const char *variable_names, const char *constants,
                     wave_expr **out, char *errbuf, size_t errbuf_len);
double wave_expr_value(const wave_expr *e, double x, double y, double t);
int wave_expr_is_time_dependent(const wave_expr *e);
void wave_expr_destroy(wave_expr *e);

/* Replaces setup() + assemble_matrices(): mesh, DoF numbering, sparsity pattern, M, K, the
   scheme matrices and their Dirichlet rows (src/WaveNewmark.cpp:12-114, src/WaveTheta.cpp:12-117,
   src/WaveEquationBase.cpp:37-94). */
int wave_setup(wave_ctx *ctx);
/* Replaces the initial-condition block of run(): interpolate u0, v0 and (Newmark) the
   consistent a^0 solve (src/WaveNewmark.cpp:290-390, src/WaveTheta.cpp:350-356). */
int wave_init(wave_ctx *ctx);

/* ---- time stepping ------------------------------------------------------------------- */
/* One pass of the loop body: assemble_rhs + solve_a + update_u_v (src/WaveNewmark.cpp:424-426)
   or assemble_rhs_u + solve_u + assemble_rhs_v + solve_v (src/WaveTheta.cpp:377-383), followed by
   the old_* = * rotation.  t_np1 is the caller's accumulated time (src/WaveNewmark.cpp:409).
   iters[0..1] receive solver_control.last_step() of the solve(s).  norms[0..1] (optional, may be
   NULL) receive ||u||_2, ||v||_2 (src/WaveNewmark.cpp:429-430). */
int wave_step(wave_ctx *ctx, double t_np1, int32_t iters[2], double norms[2]);
/* n_steps passes without returning to the host in between (time accumulated as the reference
   does: t += dt from t_start).  Stops early on divergence.  steps_done, last iters and norms
   are returned.  Used by the throughput benchmark; identical arithmetic to wave_step. */
int wave_run(wave_ctx *ctx, double t_start, int32_t n_steps, double *t_end, int32_t *steps_done,
             int32_t iters[2], double norms[2], int64_t *total_iters);
/* Stateless form for callers that keep their vectors in host memory (the reference keeps them in
   Trilinos vectors): uploads u, v, (a) in canonical numbering, performs one step, downloads the
   new u, v, (a).  a may be NULL for the theta scheme.  With several ranks the arrays have global
   length and every rank reads / refreshes its own rows (plus read-only ghost rows). */
int wave_step_host(wave_ctx *ctx, double t_np1, double *u, double *v, double *a, int32_t iters[2],
                   double norms[2]);

/* ||u||_2, ||v||_2 of the current state */
int wave_norms(wave_ctx *ctx, double out[2]);
/* compute_and_log_energy: E = 1/2 (v^T M v + u^T K u) (src/WaveEquationBase.cpp:148-154) */
int wave_energy(wave_ctx *ctx, double *out);
/* compute_error + compute_relative_error: {L2, H1, rel L2, rel H1} against the Solution
   expression at time t (src/WaveEquationBase.cpp:367-423) */
int wave_errors(wave_ctx *ctx, double t, double out[4]);
/* log_point_probe: u_h at (x, y) (src/WaveEquationBase.cpp:170-206) */
int wave_probe(wave_ctx *ctx, double x, double y, double *out);

/* ---- data access (canonical 1-rank DoF numbering; collective over ranks) -------------- */
int64_t wave_n_dofs(const wave_ctx *ctx);
int64_t wave_nnz(const wave_ctx *ctx);      /* global */
int64_t wave_n_cells(const wave_ctx *ctx);
int64_t wave_local_rows(const wave_ctx *ctx, int64_t *first_row);
int64_t wave_local_nnz(const wave_ctx *ctx);
int wave_get_vector(wave_ctx *ctx, int which, double *host, size_t n);
int wave_set_vector(wave_ctx *ctx, int which, const double *host, size_t n);
/* Locally owned rows of one matrix: rowptr has local_rows+1 entries (starting at 0), col holds
   global column indices.  val may be NULL (pattern only). */
int wave_get_csr(wave_ctx *ctx, int which, int64_t *rowptr, int32_t *col, double *val);
/* Support point of every DoF and the sorted boundary-DoF list (interpolate_boundary_values). */
int wave_get_support_points(wave_ctx *ctx, double *x, double *y, size_t n);
int64_t wave_n_boundary_dofs(const wave_ctx *ctx);
int wave_get_boundary_dofs(wave_ctx *ctx, int32_t *out, size_t n);
/* Closed-form cell -> DoF map (distribute_dofs numbering), dpc entries per cell; host only. */
int wave_cell_dofs(int32_t nx, int32_t ny, int32_t r, int64_t cell, int32_t *out);
/* The same cell's DoFs in the internal storage numbering (kind-major inside each block for R = 2,
   identical to the canonical numbering for R = 1); host only, for tests of the permutation. */
int wave_cell_dofs_storage(int32_t nx, int32_t ny, int32_t r, int64_t cell, int32_t *out);
/* [deal.II] QGaussSimplex<2>(n_points_1d) as the library integrates with it: n = r+1 for assembly and
   forcing (src/WaveEquationBase.cpp:82), n = r+2 for the error norms (:371,405).  Fills up to 16
   reference-triangle points and weights (sum 1/2) and returns their number (< 0: unsupported n);
   host only. */
int wave_quadrature(int32_t n_points_1d, double *xi, double *eta, double *w);

/* ---- kernel-level entry points (parity tests, roofline measurement) ------------------- */
/* y = A x on the device path (TrilinosWrappers::SparseMatrix::vmult); host vectors, canonical
   numbering, single rank only. */
int wave_spmv(wave_ctx *ctx, int which, const double *x, double *y, size_t n);
/* SolverCG::solve(A, x, b, P) with the context's control; x is the start vector in, solution out. */
int wave_cg(wave_ctx *ctx, int which, double *x, const double *b, size_t n, int32_t *iters);
/* Time `reps` back-to-back launches of the SpMV kernel on device-resident data; returns the
   average launch duration in ms and the algorithmic bytes per launch (12 nnz + 20 n). */
int wave_bench_spmv(wave_ctx *ctx, int which, int reps, int flush_l2, double *ms_avg, double *bytes);
/* Same for one Jacobi-PCG iteration (algorithmic bytes 12 nnz + 100 n). */
int wave_bench_cg_iter(wave_ctx *ctx, int which, int reps, double *ms_avg, double *bytes);

/* ---- instrumentation ------------------------------------------------------------------ */
/* Kernels launched by this context since creation (bench.py's gpu_launches claim). */
int64_t wave_launch_count(const wave_ctx *ctx);
/* Accumulated device milliseconds per phase since the last reset, measured with CUDA events on
   the context stream when enabled: {rhs, bc, cg, update, energy, other}. */
int wave_timers_enable(wave_ctx *ctx, int on);
int wave_timers(wave_ctx *ctx, double out_ms[6], int reset);
/* Bracket every SpMV launch of the CG solves with CUDA events on the context stream (on != 0) and
   read back the accumulated count and device milliseconds (roofline.achieved of bench.py is
   algorithmic bytes / (ms / launches) taken live inside the timed steps). */
int wave_spmv_timing(wave_ctx *ctx, int on, double *launches, double *ms_total);
/* The same for the three kernels of a CG iteration: slot 0 = SpMV A d (+ d.Ad), 1 = k_cg_update
   (g += alpha Ad, h = D^-1 g, g.g, g.h), 2 = k_cg_direction (x += alpha d, d = beta d - h). */
int wave_kernel_timing(wave_ctx *ctx, int on, double launches[3], double ms_total[3]);
/* The operator behind SpMV (replaces TrilinosWrappers::SparseMatrix::vmult, src/WaveNewmark.cpp:139):
   out = {rows served by the translation-invariant stencil tables (constant wave speed, structured mesh),
   rows kept in SELL form, pattern entries of those rows, algorithmic bytes of one SpMV launch}. */
int wave_operator_info(const wave_ctx *ctx, int64_t out[4]);
/* 1 when this context runs its Jacobi-PCG solves as one cooperative kernel (K6f, csrc/cg_fused.cu): the
   default whenever the rank's rows fit on chip (about 1.2 M rows per GPU; WAVE_CG_FUSED=0 in the
   environment of wave_setup keeps the three-kernel iteration), else 0. */
int wave_cg_fused_active(const wave_ctx *ctx);
/* Device-time and iteration statistics of the CG solves since the last reset:
   out = {solves, iterations, spmv_launches, ms_total}. */
int wave_cg_stats(wave_ctx *ctx, double out[4], int reset);

/* ---- partition plan (host only, no device needed; used by the world_size-2 CPU tests) -- */
typedef struct {
    int32_t quad_row_begin, quad_row_end; /* owned quad rows [begin, end)                  */
    int64_t row_begin, row_end;           /* owned canonical DoF range                     */
    int64_t ghost_lo_begin;               /* local vector covers [ghost_lo_begin, ghost_hi_end) */
    int64_t ghost_hi_end;
} wave_partition;
int wave_partition_plan(int32_t nx, int32_t ny, int32_t r, int32_t rank, int32_t nranks,
                        wave_partition *out);
/* Coarse levels of the multigrid V-cycle (WAVE_PRECOND_MG; stands in for PreconditionAMG,
   src/WaveNewmark.cpp:246-251) for the scheme matrix M + s K on cfg's mesh, wave speed c0 at the centre of
   the box: fills (nx, ny) of the levels below the fine one -- P1 on the same mesh first when R = 2, then
   the halved meshes -- and returns their number (< 0: error).  Host only; depends on cfg->nranks (every
   strip must keep whole coarse quad rows) but not on cfg->rank: all ranks plan the same hierarchy. */
int wave_mg_plan(const wave_config *cfg, double s, double c0, int32_t *nx_out, int32_t *ny_out,
                 int32_t max_levels);

#ifdef __cplusplus
}
#endif
#endif /* WAVEGPU_H */
