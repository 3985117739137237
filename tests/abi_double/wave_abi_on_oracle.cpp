// wave_abi_on_oracle.cpp -- TEST DOUBLE of the C ABI (include/wavegpu.h) on top of the CPU oracle.
//
// TEST INFRASTRUCTURE ONLY.  This file is compiled by tests/abi_double/build_double.py into executables
// under tests/_build/ and nowhere else: it is not part of libwavegpu.so, nmpde-wave-equation_b200/bin/ never
// links it, and nothing under nmpde-wave-equation_b200/ refers to it.  The product has no CPU path:
// libwavegpu's wave_create fails with WAVE_ERR_CUDA without a device.
//
// Purpose: the host side of the drop-in -- ParameterReader, WaveEquationBase, WaveNewmark, WaveTheta, the
// CSV / folder / exit-code conventions and the launcher -- is plain C++ above the ABI.  Linked against this
// double instead of libwavegpu it runs where there is no GPU, so that
//   * the host classes are tested end to end on CPU (tests/test_host_on_oracle_cpu.py), and
//   * the reference's own sweep drivers (scripts/convergence_sweep.py, dissipation_dispersion_sweep.py,
//     scalability_sweep.py) can be run UNMODIFIED against the launcher + executables in a scratch checkout
//     layout and their output tables compared with the reference's result tables
//     (tests/test_reference_scripts_cpu.py, tools/reference_scripts_report.py).
// Only the entry points the host classes call are implemented; numerics = oracle/wave_oracle.c (the
// restatement of src/WaveNewmark.cpp / src/WaveTheta.cpp), single rank.  The host-only entry points
// (expressions, quadrature, cell DoFs, partition plan) reuse the product's host code (csrc/expr.cpp,
// csrc/quadrature.cpp, csrc/mesh.h), which needs no device.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>

#include "expr.hpp"
#include "mesh.h"
#include "wavegpu.h"

namespace wv {
Quadrature make_quadrature(int n1d);  // csrc/quadrature.cpp
}

extern "C" {
// oracle/wave_oracle.c
struct oracle_problem;
oracle_problem *oracle_create(int Nx, int Ny, double x0, double x1, double y0, double y1, int r);
const char *oracle_last_error(oracle_problem *p);
int oracle_set_expr(oracle_problem *p, int which, const char *expr, const char *vars, const char *consts);
int oracle_setup(oracle_problem *p);
int oracle_assemble(oracle_problem *p);
void oracle_set_cg(oracle_problem *p, int maxit, double tol, double reduce, int precond);
int oracle_newmark_init(oracle_problem *p, double dt, double beta, double gamma);
int oracle_newmark_step(oracle_problem *p);
int oracle_theta_init(oracle_problem *p, double dt, double theta);
int oracle_theta_step(oracle_problem *p);
double oracle_energy(oracle_problem *p);
int oracle_errors(oracle_problem *p, double t, double *out);
double oracle_probe(oracle_problem *p);
int64_t oracle_n(oracle_problem *p);
int64_t oracle_nnz(oracle_problem *p);
double oracle_time(oracle_problem *p);
void oracle_last_iterations(oracle_problem *p, int *its);
int oracle_get_vector(oracle_problem *p, int which, double *out);
double oracle_norm(oracle_problem *p, int which);
void oracle_destroy(oracle_problem *p);
}

struct wave_ctx {
    oracle_problem *o = nullptr;
    wave_config cfg{};
    std::string err;
    bool is_setup = false, is_init = false;
    bool has[WAVE_EXPR_COUNT] = {};
    double solves = 0.0, iterations = 0.0;
};

struct wave_expr {
    wv::Program prog;
};

namespace {
std::string g_create_error;

int fail(wave_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

void count_solves(wave_ctx *ctx, int32_t iters[2]) {
    int its[2] = {0, 0};
    oracle_last_iterations(ctx->o, its);
    const int nsolves = ctx->cfg.scheme == WAVE_SCHEME_THETA && ctx->is_init ? 2 : 1;
    ctx->solves += nsolves;
    ctx->iterations += its[0] + its[1];
    if (iters) { iters[0] = its[0]; iters[1] = its[1]; }
}
}  // namespace

extern "C" {

void wave_default_config(wave_config *c) {  // the declared defaults, src/ParameterReader.cpp:39-105
    std::memset(c, 0, sizeof(*c));
    c->nx = c->ny = 40;
    c->x0 = 0.0; c->x1 = 1.0; c->y0 = 0.0; c->y1 = 1.0;
    c->r = 1; c->scheme = WAVE_SCHEME_NEWMARK;
    c->dt = 0.01; c->theta = 0.5; c->beta = 0.25; c->gamma = 0.5;
    c->cg_maxit = 10000; c->cg_tol = 1e-12; c->cg_reduce = 1e-6;
    c->precond = WAVE_PRECOND_JACOBI;
    c->rank = 0; c->nranks = 1; c->device = -1;
}

const char *wave_last_error(const wave_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int wave_device_count(void) { return 0; }  // the launcher then starts one process, as on a box without GPUs

int wave_comm_unique_id(void *) { return fail(nullptr, WAVE_ERR_UNSUPPORTED, "test double: single rank only"); }

int wave_create(const wave_config *cfg, wave_ctx **out) {
    if (!cfg || !out) return fail(nullptr, WAVE_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->nranks != 1) return fail(nullptr, WAVE_ERR_UNSUPPORTED, "test double: single rank only");
    if (cfg->nx < 1 || cfg->ny < 1 || !(cfg->x1 > cfg->x0) || !(cfg->y1 > cfg->y0))
        return fail(nullptr, WAVE_ERR_ARG, "invalid mesh");
    if (cfg->r != 1 && cfg->r != 2) return fail(nullptr, WAVE_ERR_UNSUPPORTED, "only R = 1, 2");
    if (!(cfg->dt > 0.0)) return fail(nullptr, WAVE_ERR_ARG, "Dt must be positive");
    wave_ctx *ctx = new (std::nothrow) wave_ctx();
    if (!ctx) return fail(nullptr, WAVE_ERR_ARG, "out of memory");
    ctx->cfg = *cfg;
    ctx->o = oracle_create(cfg->nx, cfg->ny, cfg->x0, cfg->x1, cfg->y0, cfg->y1, cfg->r);
    if (!ctx->o) {
        delete ctx;
        return fail(nullptr, WAVE_ERR_ARG, "oracle_create failed");
    }
    *out = ctx;
    return WAVE_OK;
}

void wave_destroy(wave_ctx *ctx) {
    if (!ctx) return;
    if (ctx->o) oracle_destroy(ctx->o);
    delete ctx;
}

int wave_set_expr(wave_ctx *ctx, int which, const char *expression, const char *variable_names, const char *constants) {
    if (!ctx || which < 0 || which >= WAVE_EXPR_COUNT || !expression) return fail(ctx, WAVE_ERR_ARG, "bad expression slot");
    if (oracle_set_expr(ctx->o, which, expression, variable_names ? variable_names : "", constants ? constants : ""))
        return fail(ctx, WAVE_ERR_EXPR, oracle_last_error(ctx->o));
    ctx->has[which] = true;
    return WAVE_OK;
}

int wave_expr_create(const char *expression, const char *variable_names, const char *constants, wave_expr **out,
                     char *errbuf, size_t errbuf_len) {
    if (!expression || !out) return WAVE_ERR_ARG;
    *out = nullptr;
    try {
        wave_expr *e = new wave_expr();
        e->prog = wv::compile_expression(expression, variable_names ? variable_names : "", constants ? constants : "");
        *out = e;
        return WAVE_OK;
    } catch (const std::exception &ex) {
        if (errbuf && errbuf_len) std::snprintf(errbuf, errbuf_len, "%s", ex.what());
        return WAVE_ERR_EXPR;
    }
}
double wave_expr_value(const wave_expr *e, double x, double y, double t) { return wv::eval(&e->prog, x, y, t); }
int wave_expr_is_time_dependent(const wave_expr *e) { return e->prog.time_dependent; }
void wave_expr_destroy(wave_expr *e) { delete e; }

int wave_setup(wave_ctx *ctx) {
    if (!ctx) return WAVE_ERR_ARG;
    for (int k = 0; k < WAVE_EXPR_SOLUTION; ++k)
        if (!ctx->has[k]) return fail(ctx, WAVE_ERR_STATE, "wave_setup before every expression is set");
    if (oracle_setup(ctx->o) || oracle_assemble(ctx->o)) return fail(ctx, WAVE_ERR_ARG, oracle_last_error(ctx->o));
    const wave_config &c = ctx->cfg;
    oracle_set_cg(ctx->o, c.cg_maxit > 0 ? c.cg_maxit : 10000, c.cg_tol > 0.0 ? c.cg_tol : 1e-12,
                  c.cg_reduce > 0.0 ? c.cg_reduce : 1e-6, c.precond);
    ctx->is_setup = true;
    return WAVE_OK;
}

int wave_init(wave_ctx *ctx) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_init before wave_setup");
    const wave_config &c = ctx->cfg;
    const int rc = c.scheme == WAVE_SCHEME_NEWMARK ? oracle_newmark_init(ctx->o, c.dt, c.beta, c.gamma)
                                                   : oracle_theta_init(ctx->o, c.dt, c.theta);
    if (c.scheme == WAVE_SCHEME_NEWMARK) count_solves(ctx, nullptr);
    ctx->is_init = true;
    return rc ? fail(ctx, WAVE_ERR_NOCONV, "CG did not converge in the a0 solve") : WAVE_OK;
}

int wave_step(wave_ctx *ctx, double t_np1, int32_t iters[2], double norms[2]) {
    if (!ctx) return WAVE_ERR_ARG;
    if (!ctx->is_init) return fail(ctx, WAVE_ERR_STATE, "wave_step before wave_init");
    const int rc = ctx->cfg.scheme == WAVE_SCHEME_NEWMARK ? oracle_newmark_step(ctx->o) : oracle_theta_step(ctx->o);
    // the oracle accumulates time exactly as the caller does (time += dt from 0, src/WaveNewmark.cpp:409)
    if (oracle_time(ctx->o) != t_np1) return fail(ctx, WAVE_ERR_ARG, "test double: caller's time is not the accumulated time");
    count_solves(ctx, iters);
    if (norms) { norms[0] = oracle_norm(ctx->o, WAVE_VEC_U); norms[1] = oracle_norm(ctx->o, WAVE_VEC_V); }
    return rc ? fail(ctx, WAVE_ERR_NOCONV, "CG did not converge") : WAVE_OK;
}

int wave_norms(wave_ctx *ctx, double out[2]) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_norms before wave_setup");
    out[0] = oracle_norm(ctx->o, WAVE_VEC_U);
    out[1] = oracle_norm(ctx->o, WAVE_VEC_V);
    return WAVE_OK;
}

int wave_energy(wave_ctx *ctx, double *out) {
    if (!ctx || !ctx->is_setup || !out) return fail(ctx, WAVE_ERR_STATE, "wave_energy before wave_setup");
    *out = oracle_energy(ctx->o);
    return WAVE_OK;
}

int wave_errors(wave_ctx *ctx, double t, double out[4]) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_errors before wave_setup");
    if (!ctx->has[WAVE_EXPR_SOLUTION] || oracle_errors(ctx->o, t, out)) return fail(ctx, WAVE_ERR_STATE, "no exact solution");
    return WAVE_OK;
}

int wave_probe(wave_ctx *ctx, double x, double y, double *out) {
    if (!ctx || !ctx->is_setup || !out) return fail(ctx, WAVE_ERR_STATE, "wave_probe before wave_setup");
    const wave_config &c = ctx->cfg;  // the oracle probes the box centre, the only point the host classes ask for
    if (x != 0.5 * (c.x0 + c.x1) || y != 0.5 * (c.y0 + c.y1)) return fail(ctx, WAVE_ERR_UNSUPPORTED, "test double: centre only");
    *out = oracle_probe(ctx->o);
    return WAVE_OK;
}

int64_t wave_n_dofs(const wave_ctx *ctx) {
    if (!ctx) return 0;
    wv::Mesh m{};
    m.nx = ctx->cfg.nx; m.ny = ctx->cfg.ny; m.r = ctx->cfg.r;
    return wv::n_dofs(m);
}
int64_t wave_local_nnz(const wave_ctx *ctx) { return ctx && ctx->is_setup ? oracle_nnz(ctx->o) : 0; }

int wave_get_vector(wave_ctx *ctx, int which, double *host, size_t n) {
    if (!ctx || !ctx->is_setup) return fail(ctx, WAVE_ERR_STATE, "wave_get_vector before wave_setup");
    if (!host || (int64_t)n != oracle_n(ctx->o) || which < 0 || which > WAVE_VEC_RHS) return fail(ctx, WAVE_ERR_ARG, "bad vector id or size");
    return oracle_get_vector(ctx->o, which, host) ? fail(ctx, WAVE_ERR_ARG, "bad vector id") : WAVE_OK;
}

int wave_cg_stats(wave_ctx *ctx, double out[4], int reset) {
    if (!ctx || !out) return WAVE_ERR_ARG;
    out[0] = ctx->solves; out[1] = ctx->iterations; out[2] = ctx->iterations + ctx->solves; out[3] = 0.0;
    if (reset) ctx->solves = ctx->iterations = 0.0;
    return WAVE_OK;
}

int wave_cell_dofs(int32_t nx, int32_t ny, int32_t r, int64_t cell, int32_t *out) {
    if (nx < 1 || ny < 1 || (r != 1 && r != 2) || cell < 0 || cell >= 2LL * nx * ny || !out) return WAVE_ERR_ARG;
    wv::Mesh m{};
    m.nx = nx; m.ny = ny; m.r = r;
    int64_t d[6];
    wv::cell_dofs(m, cell, d);
    for (int k = 0; k < wv::dofs_per_cell(r); ++k) out[k] = (int32_t)d[k];
    return WAVE_OK;
}

int wave_quadrature(int32_t n_points_1d, double *xi, double *eta, double *w) {
    if (!xi || !eta || !w) return WAVE_ERR_ARG;
    try {
        const wv::Quadrature q = wv::make_quadrature(n_points_1d);
        for (int k = 0; k < q.nq; ++k) { xi[k] = q.xi[k]; eta[k] = q.eta[k]; w[k] = q.w[k]; }
        return q.nq;
    } catch (const std::exception &) {
        return WAVE_ERR_ARG;
    }
}

int wave_partition_plan(int32_t nx, int32_t ny, int32_t r, int32_t rank, int32_t nranks, wave_partition *out) {
    if (nx < 1 || ny < 1 || (r != 1 && r != 2) || nranks != 1 || rank != 0 || !out) return WAVE_ERR_ARG;
    wv::Mesh m{};
    m.nx = nx; m.ny = ny; m.r = r;
    out->quad_row_begin = 0;
    out->quad_row_end = ny;
    out->row_begin = out->ghost_lo_begin = 0;
    out->row_end = out->ghost_hi_end = wv::n_dofs(m);
    return WAVE_OK;
}

}  // extern "C"
