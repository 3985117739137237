/*
 * wave_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the time-stepping hot path of
 * AlessandroGhiotto/nmpde-wave-equation.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (libwavegpu.so) never links, loads or calls anything in here.
 *
 * The reference cannot be compiled in this environment (needs deal.II >= 9.3.1,
 * Trilinos, MPI, muParser -- reference CMakeLists.txt:21,37-39), so this file
 * restates the algorithm from the reference sources plus the published deal.II
 * semantics at the reference's call sites.  Parity pinning: this oracle is
 * checked against the reference's own shipped result tables
 * (analysis/data/convergence-results.csv, analysis/data/dissdisp-results.csv;
 * see tests/test_oracle_golden.py and tests/golden/).  Items no reference
 * artefact pins (DoF numbering, sparsity pattern, solution vectors, inhomogeneous
 * BC / forcing paths) are "parity unpinned": there this oracle is the definition.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * the reference root).  "[deal.II]" marks behaviour of the un-vendored library
 * restated at the reference's call site.
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC wave_oracle.c -lm -o _build/libwaveoracle.so
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
/* loops shorter than this run on the calling thread: small test meshes would spend their time in
   the fork/join of thousands of tiny parallel regions */
#define OMP_MIN_WORK 50000

/* ------------------------------------------------------------------------- */
/* Expression evaluator: the muParser subset deal.II's FunctionParser accepts   */
/* and the shipped parameter files use (src/ParameterReader.cpp:139-175,        */
/* the shipped parameter files).  AST walker, double arithmetic, `^` = pow, comparisons   */
/* and && / || yield 1.0 / 0.0, if(c,a,b).                                      */
/* ------------------------------------------------------------------------- */
enum {
    N_NUM, N_VAR, N_NEG, N_ADD, N_SUB, N_MUL, N_DIV, N_POW, N_LT, N_LE, N_GT, N_GE,
    N_EQ, N_NE, N_AND, N_OR, N_IF, N_F1, N_F2, N_NOT
};
enum {
    F_SIN, F_COS, F_TAN, F_ASIN, F_ACOS, F_ATAN, F_SINH, F_COSH, F_TANH, F_ASINH,
    F_ACOSH, F_ATANH, F_EXP, F_LOG, F_LOG2, F_LOG10, F_SQRT, F_ABS, F_SIGN, F_RINT,
    F_FLOOR, F_CEIL, F_INT, F_ERFC, F_COT, F_CSC, F_SEC, F_MIN, F_MAX, F_POW2
};
typedef struct node {
    int kind, fn, var;
    double val;
    struct node *a, *b, *c;
} node;

typedef struct {
    node *root;
    int time_dependent;
    int defined;
} expr_t;

typedef struct {
    const char *s;
    int pos;
    char err[256];
    char names[64][32];
    double vals[64];
    int nconst;
    char vars[3][32];
    int nvars;
} parser;

static node *mk(int kind) {
    node *n = (node *)calloc(1, sizeof(node));
    n->kind = kind;
    return n;
}
static void free_node(node *n) {
    if (!n) return;
    free_node(n->a);
    free_node(n->b);
    free_node(n->c);
    free(n);
}
static void skipws(parser *p) {
    while (p->s[p->pos] && isspace((unsigned char)p->s[p->pos])) p->pos++;
}
static node *parse_expr(parser *p);
static node *parse_unary(parser *p);

static const struct { const char *name; int fn; int nargs; } FUNCS[] = {
    {"sin", F_SIN, 1},   {"cos", F_COS, 1},     {"tan", F_TAN, 1},     {"asin", F_ASIN, 1},
    {"acos", F_ACOS, 1}, {"atan", F_ATAN, 1},   {"sinh", F_SINH, 1},   {"cosh", F_COSH, 1},
    {"tanh", F_TANH, 1}, {"asinh", F_ASINH, 1}, {"acosh", F_ACOSH, 1}, {"atanh", F_ATANH, 1},
    {"exp", F_EXP, 1},   {"log", F_LOG, 1},     {"ln", F_LOG, 1},      {"log2", F_LOG2, 1},
    {"log10", F_LOG10, 1}, {"sqrt", F_SQRT, 1}, {"abs", F_ABS, 1},     {"sign", F_SIGN, 1},
    {"rint", F_RINT, 1}, {"floor", F_FLOOR, 1}, {"ceil", F_CEIL, 1},   {"int", F_INT, 1},
    {"erfc", F_ERFC, 1}, {"cot", F_COT, 1},     {"csc", F_CSC, 1},     {"sec", F_SEC, 1},
    {"min", F_MIN, 2},   {"max", F_MAX, 2},     {"pow", F_POW2, 2},    {NULL, 0, 0}};

static node *parse_primary(parser *p) {
    skipws(p);
    char ch = p->s[p->pos];
    if (ch == '(') {
        p->pos++;
        node *e = parse_expr(p);
        if (!e) return NULL;
        skipws(p);
        if (p->s[p->pos] != ')') {
            snprintf(p->err, sizeof p->err, "expected ')' at %d", p->pos);
            free_node(e);
            return NULL;
        }
        p->pos++;
        return e;
    }
    if (isdigit((unsigned char)ch) || ch == '.') {
        char *end;
        double v = strtod(p->s + p->pos, &end);
        if (end == p->s + p->pos) {
            snprintf(p->err, sizeof p->err, "bad number at %d", p->pos);
            return NULL;
        }
        p->pos = (int)(end - p->s);
        node *n = mk(N_NUM);
        n->val = v;
        return n;
    }
    if (isalpha((unsigned char)ch) || ch == '_') {
        char id[64];
        int k = 0;
        while ((isalnum((unsigned char)p->s[p->pos]) || p->s[p->pos] == '_') && k < 63)
            id[k++] = p->s[p->pos++];
        id[k] = 0;
        skipws(p);
        if (p->s[p->pos] == '(') {
            p->pos++;
            node *args[3] = {0, 0, 0};
            int na = 0;
            for (;;) {
                if (na >= 3) {
                    snprintf(p->err, sizeof p->err, "too many arguments to %s", id);
                    goto fail;
                }
                args[na] = parse_expr(p);
                if (!args[na]) goto fail;
                na++;
                skipws(p);
                if (p->s[p->pos] == ',') { p->pos++; continue; }
                if (p->s[p->pos] == ')') { p->pos++; break; }
                snprintf(p->err, sizeof p->err, "expected ',' or ')' at %d", p->pos);
                goto fail;
            }
            if (!strcmp(id, "if")) {
                if (na != 3) { snprintf(p->err, sizeof p->err, "if() needs 3 arguments"); goto fail; }
                node *n = mk(N_IF);
                n->a = args[0]; n->b = args[1]; n->c = args[2];
                return n;
            }
            for (int f = 0; FUNCS[f].name; ++f)
                if (!strcmp(id, FUNCS[f].name)) {
                    if (na != FUNCS[f].nargs) {
                        snprintf(p->err, sizeof p->err, "wrong argument count for %s", id);
                        goto fail;
                    }
                    node *n = mk(na == 1 ? N_F1 : N_F2);
                    n->fn = FUNCS[f].fn; n->a = args[0]; n->b = args[1];
                    return n;
                }
            snprintf(p->err, sizeof p->err, "unknown function '%s'", id);
        fail:
            for (int i = 0; i < 3; ++i) free_node(args[i]);
            return NULL;
        }
        for (int v = 0; v < p->nvars; ++v)
            if (!strcmp(id, p->vars[v])) {
                node *n = mk(N_VAR);
                n->var = v;
                return n;
            }
        for (int c = 0; c < p->nconst; ++c)
            if (!strcmp(id, p->names[c])) {
                node *n = mk(N_NUM);
                n->val = p->vals[c];
                return n;
            }
        snprintf(p->err, sizeof p->err, "unknown identifier '%s'", id);
        return NULL;
    }
    snprintf(p->err, sizeof p->err, "unexpected character '%c' at %d", ch ? ch : '$', p->pos);
    return NULL;
}
/* pow binds tighter than unary minus: -a^2 = -(a^2); a^-b allowed; right assoc */
static node *parse_pow(parser *p) {
    node *base = parse_primary(p);
    if (!base) return NULL;
    skipws(p);
    if (p->s[p->pos] == '^') {
        p->pos++;
        node *ex = parse_unary(p);
        if (!ex) { free_node(base); return NULL; }
        node *n = mk(N_POW);
        n->a = base; n->b = ex;
        return n;
    }
    return base;
}
static node *parse_unary(parser *p) {
    skipws(p);
    if (p->s[p->pos] == '-') {
        p->pos++;
        node *a = parse_unary(p);
        if (!a) return NULL;
        node *n = mk(N_NEG);
        n->a = a;
        return n;
    }
    if (p->s[p->pos] == '+') { p->pos++; return parse_unary(p); }
    if (p->s[p->pos] == '!' && p->s[p->pos + 1] != '=') {
        p->pos++;
        node *a = parse_unary(p);
        if (!a) return NULL;
        node *n = mk(N_NOT);
        n->a = a;
        return n;
    }
    return parse_pow(p);
}
static node *parse_mul(parser *p) {
    node *l = parse_unary(p);
    while (l) {
        skipws(p);
        char c = p->s[p->pos];
        if (c != '*' && c != '/') break;
        p->pos++;
        node *r = parse_unary(p);
        if (!r) { free_node(l); return NULL; }
        node *n = mk(c == '*' ? N_MUL : N_DIV);
        n->a = l; n->b = r; l = n;
    }
    return l;
}
static node *parse_add(parser *p) {
    node *l = parse_mul(p);
    while (l) {
        skipws(p);
        char c = p->s[p->pos];
        if (c != '+' && c != '-') break;
        p->pos++;
        node *r = parse_mul(p);
        if (!r) { free_node(l); return NULL; }
        node *n = mk(c == '+' ? N_ADD : N_SUB);
        n->a = l; n->b = r; l = n;
    }
    return l;
}
static node *parse_cmp(parser *p) {
    node *l = parse_add(p);
    while (l) {
        skipws(p);
        const char *s = p->s + p->pos;
        int kind = -1, len = 0;
        if (s[0] == '<' && s[1] == '=') { kind = N_LE; len = 2; }
        else if (s[0] == '>' && s[1] == '=') { kind = N_GE; len = 2; }
        else if (s[0] == '=' && s[1] == '=') { kind = N_EQ; len = 2; }
        else if (s[0] == '!' && s[1] == '=') { kind = N_NE; len = 2; }
        else if (s[0] == '<') { kind = N_LT; len = 1; }
        else if (s[0] == '>') { kind = N_GT; len = 1; }
        if (kind < 0) break;
        p->pos += len;
        node *r = parse_add(p);
        if (!r) { free_node(l); return NULL; }
        node *n = mk(kind);
        n->a = l; n->b = r; l = n;
    }
    return l;
}
static node *parse_and(parser *p) {
    node *l = parse_cmp(p);
    while (l) {
        skipws(p);
        if (!(p->s[p->pos] == '&' && p->s[p->pos + 1] == '&')) break;
        p->pos += 2;
        node *r = parse_cmp(p);
        if (!r) { free_node(l); return NULL; }
        node *n = mk(N_AND);
        n->a = l; n->b = r; l = n;
    }
    return l;
}
static node *parse_or(parser *p) {
    node *l = parse_and(p);
    while (l) {
        skipws(p);
        if (!(p->s[p->pos] == '|' && p->s[p->pos + 1] == '|')) break;
        p->pos += 2;
        node *r = parse_and(p);
        if (!r) { free_node(l); return NULL; }
        node *n = mk(N_OR);
        n->a = l; n->b = r; l = n;
    }
    return l;
}
static node *parse_expr(parser *p) {
    node *c = parse_or(p);
    if (!c) return NULL;
    skipws(p);
    if (p->s[p->pos] == '?') {
        p->pos++;
        node *a = parse_expr(p);
        if (!a) { free_node(c); return NULL; }
        skipws(p);
        if (p->s[p->pos] != ':') {
            snprintf(p->err, sizeof p->err, "expected ':' at %d", p->pos);
            free_node(c); free_node(a);
            return NULL;
        }
        p->pos++;
        node *b = parse_expr(p);
        if (!b) { free_node(c); free_node(a); return NULL; }
        node *n = mk(N_IF);
        n->a = c; n->b = a; n->c = b;
        return n;
    }
    return c;
}

static double eval_node(const node *n, const double *v) {
    switch (n->kind) {
    case N_NUM: return n->val;
    case N_VAR: return v[n->var];
    case N_NEG: return -eval_node(n->a, v);
    case N_NOT: return eval_node(n->a, v) == 0.0 ? 1.0 : 0.0;
    case N_ADD: return eval_node(n->a, v) + eval_node(n->b, v);
    case N_SUB: return eval_node(n->a, v) - eval_node(n->b, v);
    case N_MUL: return eval_node(n->a, v) * eval_node(n->b, v);
    case N_DIV: return eval_node(n->a, v) / eval_node(n->b, v);
    case N_POW: return pow(eval_node(n->a, v), eval_node(n->b, v));
    case N_LT: return eval_node(n->a, v) < eval_node(n->b, v) ? 1.0 : 0.0;
    case N_LE: return eval_node(n->a, v) <= eval_node(n->b, v) ? 1.0 : 0.0;
    case N_GT: return eval_node(n->a, v) > eval_node(n->b, v) ? 1.0 : 0.0;
    case N_GE: return eval_node(n->a, v) >= eval_node(n->b, v) ? 1.0 : 0.0;
    case N_EQ: return eval_node(n->a, v) == eval_node(n->b, v) ? 1.0 : 0.0;
    case N_NE: return eval_node(n->a, v) != eval_node(n->b, v) ? 1.0 : 0.0;
    case N_AND: return (eval_node(n->a, v) != 0.0 && eval_node(n->b, v) != 0.0) ? 1.0 : 0.0;
    case N_OR: return (eval_node(n->a, v) != 0.0 || eval_node(n->b, v) != 0.0) ? 1.0 : 0.0;
    case N_IF: return eval_node(n->a, v) != 0.0 ? eval_node(n->b, v) : eval_node(n->c, v);
    case N_F2: {
        double a = eval_node(n->a, v), b = eval_node(n->b, v);
        switch (n->fn) {
        case F_MIN: return a < b ? a : b;
        case F_MAX: return a > b ? a : b;
        default: return pow(a, b);
        }
    }
    case N_F1: {
        double a = eval_node(n->a, v);
        switch (n->fn) {
        case F_SIN: return sin(a);
        case F_COS: return cos(a);
        case F_TAN: return tan(a);
        case F_ASIN: return asin(a);
        case F_ACOS: return acos(a);
        case F_ATAN: return atan(a);
        case F_SINH: return sinh(a);
        case F_COSH: return cosh(a);
        case F_TANH: return tanh(a);
        case F_ASINH: return asinh(a);
        case F_ACOSH: return acosh(a);
        case F_ATANH: return atanh(a);
        case F_EXP: return exp(a);
        case F_LOG: return log(a);
        case F_LOG2: return log2(a);
        case F_LOG10: return log10(a);
        case F_SQRT: return sqrt(a);
        case F_ABS: return fabs(a);
        case F_SIGN: return a > 0 ? 1.0 : (a < 0 ? -1.0 : 0.0);
        case F_RINT: return rint(a);
        case F_FLOOR: return floor(a);
        case F_CEIL: return ceil(a);
        case F_INT: return rint(a);
        case F_ERFC: return erfc(a);
        case F_COT: return 1.0 / tan(a);
        case F_CSC: return 1.0 / sin(a);
        case F_SEC: return 1.0 / cos(a);
        }
    }
    }
    return NAN;
}

static void trim(char *s) {
    char *b = s;
    while (*b && isspace((unsigned char)*b)) b++;
    if (b != s) memmove(s, b, strlen(b) + 1);
    size_t n = strlen(s);
    while (n && isspace((unsigned char)s[n - 1])) s[--n] = 0;
}

/* src/ParameterReader.cpp:237-265 parse_value_with_pi: "pi", "<num>*pi", or a number */
static int parse_value_with_pi(const char *value, double *out) {
    char buf[128];
    strncpy(buf, value, 127);
    buf[127] = 0;
    trim(buf);
    char low[128];
    size_t i;
    for (i = 0; buf[i]; ++i) low[i] = (char)tolower((unsigned char)buf[i]);
    low[i] = 0;
    if (!strcmp(low, "pi")) { *out = M_PI; return 0; }
    char *star = strchr(low, '*');
    if (star) {
        char lhs[128], rhs[128];
        size_t l = (size_t)(star - low);
        memcpy(lhs, low, l); lhs[l] = 0;
        strcpy(rhs, star + 1);
        trim(lhs); trim(rhs);
        if (!strcmp(rhs, "pi")) {
            /* the reference's pattern for <num>: [0-9]*\.?[0-9]+ -- digits only, no sign, no exponent,
               no trailing dot; anything else falls through to the numeric-prefix parse below */
            size_t a = 0, n = strlen(lhs), ok = 0;
            while (a < n && isdigit((unsigned char)lhs[a])) ++a;
            if (a < n && lhs[a] == '.') {
                size_t b = a + 1;
                while (b < n && isdigit((unsigned char)lhs[b])) ++b;
                ok = (b > a + 1 && b == n);
            } else
                ok = (a > 0 && a == n);
            if (ok) { *out = strtod(lhs, NULL) * M_PI; return 0; }
        }
    }
    char *end;
    double c = strtod(buf, &end);
    if (end == buf) return -1;
    *out = c;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Problem object                                                             */
/* ------------------------------------------------------------------------- */
enum { EX_C = 0, EX_F, EX_U0, EX_V0, EX_G, EX_DGDT, EX_SOL, EX_COUNT };

typedef struct {
    int nq;
    double xi[16], eta[16], w[16];
} quadrule_t;

struct mg_t;
typedef struct oracle_problem {
    int Nx, Ny, r;
    double x0, x1, y0, y1;
    expr_t ex[EX_COUNT];
    char err[256];
    /* mesh (src/WaveEquationBase.cpp:37-72) */
    int64_t ncells, nverts;
    double *vx, *vy;
    int *cell_v; /* 3 per cell */
    /* dofs (src/WaveEquationBase.cpp:86-94) */
    int dpc;
    int64_t n;
    int *cell_dof; /* dpc per cell */
    double *sx, *sy; /* DoF support points */
    int64_t nb;
    int *bdof;       /* sorted boundary DoFs */
    unsigned char *is_b;
    /* CSR (src/WaveNewmark.cpp:30-41) */
    int64_t *rowptr;
    int *col;
    int64_t nnz;
    double *M, *K, *A, *A2, *S; /* mass, stiffness, matrix_a/matrix_u, matrix_v, system matrix scratch */
    quadrule_t q, qerr;
    /* vectors */
    double *u, *v, *a, *ou, *ov, *oa, *rhs, *t1, *t2, *t3, *cg_g, *cg_d, *cg_h, *dinv;
    /* scheme */
    int scheme; /* 0 newmark 1 theta */
    double dt, beta, gamma, theta, time;
    int step;
    /* CG control: ReductionControl(10000, 1e-12, 1e-6) src/WaveNewmark.cpp:256 */
    int cg_maxit;
    double cg_tol, cg_reduce;
    int last_its[2];
    int precond; /* 0 = Jacobi (north-star replacement of AMG), 1 = identity, 2 = multigrid V-cycle */
    struct mg_t *mg;   /* hierarchy for precond == 2 (built by the scheme init) */
    int borrowed_expr; /* coarse-level problems share the fine problem's expression trees */
    int forcing_every_step; /* 1 = as the reference (assemble F even if zero) */
} oracle_problem;

static double ev(const oracle_problem *p, int which, double x, double y, double t) {
    double v[3] = {x, y, t};
    return eval_node(p->ex[which].root, v);
}

oracle_problem *oracle_create(int Nx, int Ny, double x0, double x1, double y0, double y1, int r) {
    if (Nx < 1 || Ny < 1 || (r != 1 && r != 2)) return NULL;
    oracle_problem *p = (oracle_problem *)calloc(1, sizeof *p);
    p->Nx = Nx; p->Ny = Ny; p->r = r;
    p->x0 = x0; p->x1 = x1; p->y0 = y0; p->y1 = y1;
    p->cg_maxit = 10000; p->cg_tol = 1e-12; p->cg_reduce = 1e-6;
    p->forcing_every_step = 1;
    return p;
}
const char *oracle_last_error(oracle_problem *p) { return p->err; }

/* src/ParameterReader.cpp:139-175 load_functions + :267-294 constants */
int oracle_set_expr(oracle_problem *p, int which, const char *expr, const char *vars,
                    const char *consts) {
    if (which < 0 || which >= EX_COUNT) return -1;
    parser ps;
    memset(&ps, 0, sizeof ps);
    /* constants "k=v, k2=v2" */
    char cbuf[1024];
    strncpy(cbuf, consts ? consts : "", 1023);
    cbuf[1023] = 0;
    for (char *item = strtok(cbuf, ","); item; item = strtok(NULL, ",")) {
        char *eq = strchr(item, '=');
        if (!eq) continue;
        *eq = 0;
        char key[64];
        strncpy(key, item, 63); key[63] = 0;
        trim(key);
        double val;
        if (parse_value_with_pi(eq + 1, &val)) {
            snprintf(p->err, sizeof p->err, "bad constant value for %s", key);
            return -2;
        }
        strncpy(ps.names[ps.nconst], key, 31);
        ps.vals[ps.nconst++] = val;
    }
    strcpy(ps.names[ps.nconst], "pi"); /* :167 constants["pi"] */
    ps.vals[ps.nconst++] = M_PI;
    /* variables; time-dependent iff the string contains 't' (:168, substring test) */
    char vbuf[256];
    strncpy(vbuf, vars ? vars : "", 255);
    vbuf[255] = 0;
    int td = strchr(vbuf, 't') != NULL;
    for (char *item = strtok(vbuf, ","); item; item = strtok(NULL, ",")) {
        if (ps.nvars >= 3) { snprintf(p->err, sizeof p->err, "too many variables"); return -3; }
        strncpy(ps.vars[ps.nvars], item, 31);
        trim(ps.vars[ps.nvars]);
        ps.nvars++;
    }
    if (ps.nvars != 2 + td) { /* [deal.II] FunctionParser::initialize dimension check */
        snprintf(p->err, sizeof p->err, "variable list must name %d variables", 2 + td);
        return -3;
    }
    ps.s = expr;
    node *root = parse_expr(&ps);
    if (root) {
        skipws(&ps);
        if (ps.s[ps.pos]) {
            snprintf(ps.err, sizeof ps.err, "trailing input at %d", ps.pos);
            free_node(root);
            root = NULL;
        }
    }
    if (!root) {
        snprintf(p->err, sizeof p->err, "%s", ps.err);
        return -4;
    }
    free_node(p->ex[which].root);
    p->ex[which].root = root;
    p->ex[which].time_dependent = td;
    p->ex[which].defined = 1;
    return 0;
}
double oracle_eval(oracle_problem *p, int which, double x, double y, double t) {
    return ev(p, which, x, y, t);
}

/* ------------------------------------------------------------------------- */
/* Quadrature [deal.II QGaussSimplex<2>(n), deal.II >= 9.4 tables]            */
/* Call sites: src/WaveEquationBase.cpp:82 (n = r+1) and :371,405 (n = r+2).  */
/* n=2: 4 points, degree 3 (Hillion's scheme = 2x2 Gauss x Gauss-Jacobi(1,0)  */
/*      collapsed onto the triangle; not symmetric under vertex permutation); */
/* n=3: 7 points, degree 5 (Hammer-Marlowe-Stroud / Radon);                   */
/* n=4: 15 points, degree 7 (Witherden-Vincent).                              */
/* Which release the reference was built with is pinned by its own results:   */
/* the P2 rows of analysis/data/convergence-results.csv (error norms through  */
/* n=4) reproduce to the 7 printed digits with the Witherden-Vincent rule and */
/* only to 5e-5 with other degree-7 rules; deal.II 9.3 has no n=4 rule in 2-D.*/
/* Weights sum to 1/2 (reference triangle area).                              */
/* ------------------------------------------------------------------------- */
static void put_point(quadrule_t *q, double xi, double eta, double w) {
    q->xi[q->nq] = xi; q->eta[q->nq] = eta; q->w[q->nq] = w; q->nq++;
}
/* first two barycentric coordinates of every distinct permutation of (b0,b1,b2), in
   lexicographic order of the sorted triple */
static void put_orbit(quadrule_t *q, double b0, double b1, double b2, double w) {
    double b[3] = {b0, b1, b2};
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (b[j] < b[i]) { double t = b[i]; b[i] = b[j]; b[j] = t; }
    static const int perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    double seen[6][2];
    int ns = 0;
    for (int p = 0; p < 6; ++p) {
        const double x = b[perm[p][0]], y = b[perm[p][1]];
        int dup = 0;
        for (int k = 0; k < ns; ++k) dup |= (seen[k][0] == x && seen[k][1] == y);
        if (dup) continue;
        seen[ns][0] = x; seen[ns][1] = y; ns++;
        put_point(q, x, y, w);
    }
}
static void make_quadrature(int n1d, quadrule_t *q) {
    q->nq = 0;
    if (n1d == 2) {
        put_point(q, 0.17855872826361643, 0.1550510257216822, 0.5 * 0.31804138174397717);
        put_point(q, 0.07503111022260812, 0.6449489742783178, 0.5 * 0.18195861825602283);
        put_point(q, 0.6663902460147014, 0.1550510257216822, 0.5 * 0.31804138174397717);
        put_point(q, 0.28001991549907407, 0.6449489742783178, 0.5 * 0.18195861825602283);
    } else if (n1d == 3) {
        const double s15 = sqrt(15.0);
        const double p0 = 2.0 / 7.0 - s15 / 21.0, p1 = 2.0 / 7.0 + s15 / 21.0;
        const double p2 = 3.0 / 7.0 - 2.0 * s15 / 21.0, p3 = 3.0 / 7.0 + 2.0 * s15 / 21.0;
        const double w0 = 9.0 / 40.0, w1 = 31.0 / 240.0 - s15 / 1200.0, w2 = 31.0 / 240.0 + s15 / 1200.0;
        put_point(q, 1.0 / 3.0, 1.0 / 3.0, 0.5 * w0);
        put_point(q, p3, p0, 0.5 * w1);
        put_point(q, p0, p3, 0.5 * w1);
        put_point(q, p0, p0, 0.5 * w1);
        put_point(q, p2, p1, 0.5 * w2);
        put_point(q, p1, p2, 0.5 * w2);
        put_point(q, p1, p1, 0.5 * w2);
    } else {
        /* three orbits (a, a, 1-2a) and one orbit (b1, b2, 1-b1-b2); weights for unit total */
        static const double a[3] = {0.03373064855458785, 0.24157738259540357, 0.47430969250471822};
        static const double wa[3] = {0.016545050110792132, 0.12794417123015558, 0.07708664618598607};
        static const double b1 = 0.047036644652595234, b2 = 0.19868331479735159, wb = 0.05587873290319978;
        for (int o = 0; o < 3; ++o) put_orbit(q, a[o], a[o], 1.0 - 2.0 * a[o], 0.5 * wa[o]);
        put_orbit(q, b1, b2, 1.0 - b1 - b2, 0.5 * wb);
    }
}
int oracle_get_quadrature(int n1d, double *xi, double *eta, double *w) {
    quadrule_t q;
    make_quadrature(n1d, &q);
    for (int i = 0; i < q.nq; ++i) { xi[i] = q.xi[i]; eta[i] = q.eta[i]; w[i] = q.w[i]; }
    return q.nq;
}

/* Shape functions [deal.II FE_SimplexP(r)], reference triangle (0,0),(1,0),(0,1).
   r=1: 1-xi-eta, xi, eta.  r=2: vertex k: l_k(2 l_k - 1); then line0 (v0v1) 4 l0 l1,
   line1 (v1v2) 4 l1 l2, line2 (v2v0) 4 l2 l0. */
static void shape(int r, double xi, double eta, double *phi, double *dxi, double *deta) {
    double l0 = 1.0 - xi - eta, l1 = xi, l2 = eta;
    if (r == 1) {
        phi[0] = l0; phi[1] = l1; phi[2] = l2;
        dxi[0] = -1; dxi[1] = 1; dxi[2] = 0;
        deta[0] = -1; deta[1] = 0; deta[2] = 1;
    } else {
        phi[0] = l0 * (2 * l0 - 1); phi[1] = l1 * (2 * l1 - 1); phi[2] = l2 * (2 * l2 - 1);
        phi[3] = 4 * l0 * l1; phi[4] = 4 * l1 * l2; phi[5] = 4 * l2 * l0;
        dxi[0] = -(4 * l0 - 1); dxi[1] = 4 * l1 - 1; dxi[2] = 0;
        deta[0] = -(4 * l0 - 1); deta[1] = 0; deta[2] = 4 * l2 - 1;
        dxi[3] = 4 * (l0 - l1); deta[3] = -4 * l1;
        dxi[4] = 4 * l2; deta[4] = 4 * l1;
        dxi[5] = -4 * l2; deta[5] = 4 * (l0 - l2);
    }
}

/* ------------------------------------------------------------------------- */
/* Mesh, DoFs, sparsity                                                       */
/* ------------------------------------------------------------------------- */
typedef struct { int64_t key; int val; } hslot;
static int64_t hkey(int a, int b) {
    int lo = a < b ? a : b, hi = a < b ? b : a;
    return ((int64_t)lo << 32) | (uint32_t)hi;
}
static uint64_t hmix(int64_t k) {
    uint64_t x = (uint64_t)k;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

static int cmp_int(const void *a, const void *b) {
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

int oracle_setup(oracle_problem *p) {
    const int Nx = p->Nx, Ny = p->Ny, r = p->r;
    /* --- mesh: [deal.II] GridGenerator::subdivided_hyper_rectangle_with_simplices,
       src/WaveEquationBase.cpp:42-46.  Vertices j*(Nx+1)+i; per quad (j outer, i inner)
       T0={q0,q1,q2}, T1={q3,q2,q1}. */
    p->nverts = (int64_t)(Nx + 1) * (Ny + 1);
    p->ncells = 2LL * Nx * Ny;
    p->vx = (double *)malloc(sizeof(double) * p->nverts);
    p->vy = (double *)malloc(sizeof(double) * p->nverts);
    const double dx = (p->x1 - p->x0) / Nx, dy = (p->y1 - p->y0) / Ny;
    for (int j = 0; j <= Ny; ++j)
        for (int i = 0; i <= Nx; ++i) {
            p->vx[(int64_t)j * (Nx + 1) + i] = p->x0 + i * dx;
            p->vy[(int64_t)j * (Nx + 1) + i] = p->y0 + j * dy;
        }
    p->cell_v = (int *)malloc(sizeof(int) * 3 * p->ncells);
    {
        int64_t c = 0;
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) {
                int q0 = j * (Nx + 1) + i, q1 = q0 + 1, q2 = q0 + Nx + 1, q3 = q2 + 1;
                int *t0 = p->cell_v + 3 * c, *t1 = p->cell_v + 3 * (c + 1);
                t0[0] = q0; t0[1] = q1; t0[2] = q2;
                t1[0] = q3; t1[1] = q2; t1[2] = q1;
                c += 2;
            }
    }
    /* --- DoFs: [deal.II] DoFHandler::distribute_dofs(FE_SimplexP(r)), 1 rank,
       src/WaveEquationBase.cpp:90-91.  First touch in cell order: per cell the
       not-yet-numbered vertices (v0,v1,v2), then lines (v0v1, v1v2, v2v0). */
    p->dpc = (r == 1) ? 3 : 6;
    p->cell_dof = (int *)malloc(sizeof(int) * p->dpc * p->ncells);
    int *vdof = (int *)malloc(sizeof(int) * p->nverts);
    for (int64_t i = 0; i < p->nverts; ++i) vdof[i] = -1;
    /* line table: hash (lo,hi) -> line id; line_cells counts incidence */
    int64_t nlines_max = 3LL * Nx * Ny + Nx + Ny;
    int64_t hcap = 1;
    while (hcap < 2 * nlines_max + 16) hcap <<= 1;
    hslot *ht = (hslot *)malloc(sizeof(hslot) * hcap);
    for (int64_t i = 0; i < hcap; ++i) ht[i].key = -1;
    int *line_dof = (int *)malloc(sizeof(int) * nlines_max);
    int *line_cnt = (int *)calloc(nlines_max, sizeof(int));
    int *line_va = (int *)malloc(sizeof(int) * nlines_max);
    int *line_vb = (int *)malloc(sizeof(int) * nlines_max);
    int nlines = 0;
    int next = 0;
    for (int64_t c = 0; c < p->ncells; ++c) {
        const int *cv = p->cell_v + 3 * c;
        int *cd = p->cell_dof + p->dpc * c;
        for (int k = 0; k < 3; ++k) {
            if (vdof[cv[k]] < 0) vdof[cv[k]] = next++;
            cd[k] = vdof[cv[k]];
        }
        for (int k = 0; k < 3; ++k) {
            int a = cv[k], b = cv[(k + 1) % 3];
            int64_t key = hkey(a, b);
            uint64_t h = hmix(key) & (uint64_t)(hcap - 1);
            while (ht[h].key != -1 && ht[h].key != key) h = (h + 1) & (uint64_t)(hcap - 1);
            int lid;
            if (ht[h].key == -1) {
                ht[h].key = key;
                ht[h].val = lid = nlines++;
                line_dof[lid] = -1;
                line_va[lid] = a; line_vb[lid] = b;
            } else
                lid = ht[h].val;
            line_cnt[lid]++;
            if (r == 2) {
                if (line_dof[lid] < 0) line_dof[lid] = next++;
                cd[3 + k] = line_dof[lid];
            }
        }
    }
    p->n = next;
    /* support points and boundary DoFs ([deal.II] interpolate_boundary_values on
       all boundary ids, src/WaveNewmark.cpp:191-192; colorize=false => whole boundary) */
    p->sx = (double *)malloc(sizeof(double) * p->n);
    p->sy = (double *)malloc(sizeof(double) * p->n);
    p->is_b = (unsigned char *)calloc(p->n, 1);
    for (int64_t v = 0; v < p->nverts; ++v) { p->sx[vdof[v]] = p->vx[v]; p->sy[vdof[v]] = p->vy[v]; }
    for (int l = 0; l < nlines; ++l) {
        int a = line_va[l], b = line_vb[l];
        if (r == 2) {
            p->sx[line_dof[l]] = 0.5 * (p->vx[a] + p->vx[b]);
            p->sy[line_dof[l]] = 0.5 * (p->vy[a] + p->vy[b]);
        }
        if (line_cnt[l] == 1) {
            p->is_b[vdof[a]] = 1; p->is_b[vdof[b]] = 1;
            if (r == 2) p->is_b[line_dof[l]] = 1;
        }
    }
    p->nb = 0;
    for (int64_t i = 0; i < p->n; ++i) p->nb += p->is_b[i];
    p->bdof = (int *)malloc(sizeof(int) * (p->nb ? p->nb : 1));
    {
        int64_t k = 0;
        for (int64_t i = 0; i < p->n; ++i) if (p->is_b[i]) p->bdof[k++] = (int)i;
    }
    free(ht); free(line_dof); free(line_cnt); free(line_va); free(line_vb); free(vdof);
    /* --- sparsity: [deal.II] DoFTools::make_sparsity_pattern + compress,
       src/WaveNewmark.cpp:33-35: all (i,j) sharing a cell; rows sorted ascending. */
    const int dpc = p->dpc;
    int64_t *dcnt = (int64_t *)calloc(p->n + 1, sizeof(int64_t));
    for (int64_t c = 0; c < p->ncells; ++c)
        for (int k = 0; k < dpc; ++k) dcnt[p->cell_dof[dpc * c + k] + 1]++;
    for (int64_t i = 0; i < p->n; ++i) dcnt[i + 1] += dcnt[i];
    int *dcell = (int *)malloc(sizeof(int) * dcnt[p->n]);
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * p->n);
    memcpy(fill, dcnt, sizeof(int64_t) * p->n);
    for (int64_t c = 0; c < p->ncells; ++c)
        for (int k = 0; k < dpc; ++k) dcell[fill[p->cell_dof[dpc * c + k]]++] = (int)c;
    p->rowptr = (int64_t *)malloc(sizeof(int64_t) * (p->n + 1));
    p->rowptr[0] = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int64_t i = 0; i < p->n; ++i) {
            int tmp[64], m = 0;
            for (int64_t k = dcnt[i]; k < dcnt[i + 1]; ++k)
                for (int j = 0; j < dpc; ++j) tmp[m++] = p->cell_dof[dpc * (int64_t)dcell[k] + j];
            qsort(tmp, m, sizeof(int), cmp_int);
            int u = 0;
            for (int k = 0; k < m; ++k)
                if (k == 0 || tmp[k] != tmp[k - 1]) tmp[u++] = tmp[k];
            if (pass == 0)
                p->rowptr[i + 1] = p->rowptr[i] + u;
            else
                memcpy(p->col + p->rowptr[i], tmp, sizeof(int) * u);
        }
        if (pass == 0) {
            p->nnz = p->rowptr[p->n];
            p->col = (int *)malloc(sizeof(int) * p->nnz);
        }
    }
    free(dcnt); free(dcell); free(fill);
    /* src/WaveEquationBase.cpp:82 QGaussSimplex(r+1); :371 QGaussSimplex(r+2) */
    make_quadrature(r + 1, &p->q);
    make_quadrature(r + 2, &p->qerr);
    p->M = (double *)calloc(p->nnz, sizeof(double));
    p->K = (double *)calloc(p->nnz, sizeof(double));
    p->A = (double *)calloc(p->nnz, sizeof(double));
    p->A2 = (double *)calloc(p->nnz, sizeof(double));
    p->S = (double *)calloc(p->nnz, sizeof(double));
    double **vecs[] = {&p->u, &p->v, &p->a, &p->ou, &p->ov, &p->oa, &p->rhs, &p->t1,
                       &p->t2, &p->t3, &p->cg_g, &p->cg_d, &p->cg_h, &p->dinv};
    for (size_t i = 0; i < sizeof vecs / sizeof vecs[0]; ++i)
        *vecs[i] = (double *)calloc(p->n, sizeof(double));
    return 0;
}

static inline int64_t csr_find(const oracle_problem *p, int row, int c) {
    int64_t lo = p->rowptr[row], hi = p->rowptr[row + 1] - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) >> 1;
        if (p->col[mid] == c) return mid;
        if (p->col[mid] < c) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* affine cell geometry: J = [v1-v0, v2-v0] */
static inline void cell_geom(const oracle_problem *p, int64_t c, double *X0, double *Y0, double J[4],
                             double *det) {
    const int *cv = p->cell_v + 3 * c;
    *X0 = p->vx[cv[0]]; *Y0 = p->vy[cv[0]];
    J[0] = p->vx[cv[1]] - *X0; J[1] = p->vx[cv[2]] - *X0;
    J[2] = p->vy[cv[1]] - *Y0; J[3] = p->vy[cv[2]] - *Y0;
    *det = J[0] * J[3] - J[1] * J[2];
}

/* src/WaveNewmark.cpp:56-108 == src/WaveTheta.cpp:56-108 assemble_matrices:
   M_e(i,j) += phi_i phi_j JxW ; K_e(i,j) += c^2 grad phi_i . grad phi_j JxW,
   c evaluated at the quadrature point at parser time 0. */
int oracle_assemble(oracle_problem *p) {
    if (!p->ex[EX_C].defined) { snprintf(p->err, sizeof p->err, "C undefined"); return -1; }
    const int dpc = p->dpc, nq = p->q.nq;
    memset(p->M, 0, sizeof(double) * p->nnz);
    memset(p->K, 0, sizeof(double) * p->nnz);
    double phi[16][6], gx[16][6], gy[16][6];
    for (int64_t c = 0; c < p->ncells; ++c) {
        double X0, Y0, J[4], det;
        cell_geom(p, c, &X0, &Y0, J, &det);
        const double adet = fabs(det);
        /* J^{-T} */
        const double i00 = J[3] / det, i01 = -J[2] / det, i10 = -J[1] / det, i11 = J[0] / det;
        double Me[6][6] = {{0}}, Ke[6][6] = {{0}};
        for (int q = 0; q < nq; ++q) {
            double dxi[6], deta[6];
            shape(p->r, p->q.xi[q], p->q.eta[q], phi[q], dxi, deta);
            for (int i = 0; i < dpc; ++i) {
                gx[q][i] = i00 * dxi[i] + i01 * deta[i];
                gy[q][i] = i10 * dxi[i] + i11 * deta[i];
            }
            const double JxW = p->q.w[q] * adet;
            const double xq = X0 + J[0] * p->q.xi[q] + J[1] * p->q.eta[q];
            const double yq = Y0 + J[2] * p->q.xi[q] + J[3] * p->q.eta[q];
            const double cv = ev(p, EX_C, xq, yq, 0.0);
            const double c2 = cv * cv;
            for (int i = 0; i < dpc; ++i)
                for (int j = 0; j < dpc; ++j) {
                    Me[i][j] += phi[q][i] * phi[q][j] * JxW;
                    Ke[i][j] += c2 * (gx[q][i] * gx[q][j] + gy[q][i] * gy[q][j]) * JxW;
                }
        }
        const int *cd = p->cell_dof + dpc * c;
        for (int i = 0; i < dpc; ++i)
            for (int j = 0; j < dpc; ++j) {
                int64_t k = csr_find(p, cd[i], cd[j]);
                p->M[k] += Me[i][j];
                p->K[k] += Ke[i][j];
            }
    }
    return 0;
}

/* Trilinos vmult (Epetra_CrsMatrix::Multiply) */
static void spmv(const oracle_problem *p, const double *val, const double *x, double *y) {
#pragma omp parallel for schedule(static) if ((p->n) > OMP_MIN_WORK)
    for (int64_t i = 0; i < p->n; ++i) {
        double s = 0.0;
        for (int64_t k = p->rowptr[i]; k < p->rowptr[i + 1]; ++k) s += val[k] * x[p->col[k]];
        y[i] = s;
    }
}
static double dot(const oracle_problem *p, const double *x, const double *y) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static) if ((p->n) > OMP_MIN_WORK)
    for (int64_t i = 0; i < p->n; ++i) s += x[i] * y[i];
    return s;
}
static void axpy(const oracle_problem *p, double a, const double *x, double *y) {
#pragma omp parallel for schedule(static) if ((p->n) > OMP_MIN_WORK)
    for (int64_t i = 0; i < p->n; ++i) y[i] += a * x[i];
}
static void vcopy(const oracle_problem *p, const double *x, double *y) {
#pragma omp parallel for schedule(static) if ((p->n) > OMP_MIN_WORK)
    for (int64_t i = 0; i < p->n; ++i) y[i] = x[i];
}

/* forcing load vector: out_i += scale * sum_q f(x_q) phi_i JxW
   src/WaveNewmark.cpp:151-171 (one time level) / src/WaveTheta.cpp:151-180 (theta blend) */
static void add_forcing(oracle_problem *p, double *out, double scale, double t_np1, double t_n,
                        double w_np1, double w_n, int two_levels) {
    const int dpc = p->dpc, nq = p->q.nq;
#pragma omp parallel for schedule(static) if ((p->ncells) > OMP_MIN_WORK)
    for (int64_t c = 0; c < p->ncells; ++c) {
        double X0, Y0, J[4], det;
        cell_geom(p, c, &X0, &Y0, J, &det);
        const double adet = fabs(det);
        double cell_rhs[6] = {0};
        for (int q = 0; q < nq; ++q) {
            double phi[6], dxi[6], deta[6];
            shape(p->r, p->q.xi[q], p->q.eta[q], phi, dxi, deta);
            const double JxW = p->q.w[q] * adet;
            const double xq = X0 + J[0] * p->q.xi[q] + J[1] * p->q.eta[q];
            const double yq = Y0 + J[2] * p->q.xi[q] + J[3] * p->q.eta[q];
            double fv;
            if (two_levels) {
                const double f_n = ev(p, EX_F, xq, yq, t_n);
                const double f_np1 = ev(p, EX_F, xq, yq, t_np1);
                fv = w_np1 * f_np1 + w_n * f_n;
            } else
                fv = ev(p, EX_F, xq, yq, t_np1);
            for (int i = 0; i < dpc; ++i) cell_rhs[i] += scale * fv * phi[i] * JxW;
        }
        const int *cd = p->cell_dof + dpc * c;
        for (int i = 0; i < dpc; ++i) {
#pragma omp atomic
            out[cd[i]] += cell_rhs[i];
        }
    }
}

/* [deal.II] MatrixTools::apply_boundary_values, Trilinos overload, Release build
   (src/WaveNewmark.cpp:240-241): eliminate_columns ignored; d0 = |first non-zero
   diagonal|; boundary rows cleared to d0 on the diagonal; rhs_i = v d0; x_i = v. */
static void apply_bc(oracle_problem *p, double *Sval, const double *bval, double *x, double *rhs) {
    double d0 = 0.0;
    for (int64_t i = 0; i < p->n; ++i) {
        int64_t k = csr_find(p, (int)i, (int)i);
        if (Sval[k] != 0.0) { d0 = fabs(Sval[k]); break; }
    }
    for (int64_t b = 0; b < p->nb; ++b) {
        int i = p->bdof[b];
        for (int64_t k = p->rowptr[i]; k < p->rowptr[i + 1]; ++k) Sval[k] = (p->col[k] == i) ? d0 : 0.0;
        rhs[i] = bval[b] * d0;
        x[i] = bval[b];
    }
}

/* ------------------------------------------------------------------------- */
/* Geometric multigrid V-cycle preconditioner (precond == 2).                   */
/* Stands in for TrilinosWrappers::PreconditionAMG (ML smoothed aggregation,     */
/* Chebyshev smoother, src/WaveNewmark.cpp:246-251) -- "next" row (f).1 of       */
/* SURVEY section 8.  Hierarchy: [P2 on the mesh ->] P1 on the mesh -> P1 on the */
/* meshes with Nel/2, Nel/4, ...; coarse operators by rediscretisation           */
/* bc(M_l + s K_l); transfers = FE interpolation (P[i][j] = phi_j^coarse at the  */
/* fine support point i) and its transpose; smoother = damped Jacobi, nu pre and */
/* nu post sweeps (symmetric, so the V-cycle is an SPD preconditioner for CG).   */
/* ------------------------------------------------------------------------- */
typedef struct mg_level {
    oracle_problem *prob; /* level 0: the fine problem itself (not owned) */
    double *S, *dinv, *x, *b, *r, *t;
    int64_t *Pptr; /* prolongation from the next coarser level onto this one (rows = this level) */
    int *Pcol;
    double *Pval;
    double omega;
} mg_level;
typedef struct mg_t {
    int nlev, nu, nu_coarse;
    mg_level lev[16];
} mg_t;

int oracle_setup(oracle_problem *p);
int oracle_assemble(oracle_problem *p);
static void apply_bc(oracle_problem *p, double *Sval, const double *bval, double *x, double *rhs);

static void mg_build_prolongation(const oracle_problem *f, const oracle_problem *c, mg_level *L) {
    const double dxc = (c->x1 - c->x0) / c->Nx, dyc = (c->y1 - c->y0) / c->Ny;
    L->Pptr = (int64_t *)malloc(sizeof(int64_t) * (f->n + 1));
    L->Pcol = (int *)malloc(sizeof(int) * 6 * f->n);
    L->Pval = (double *)malloc(sizeof(double) * 6 * f->n);
    int64_t k = 0;
    for (int64_t i = 0; i < f->n; ++i) {
        L->Pptr[i] = k;
        const double x = f->sx[i], y = f->sy[i];
        int ic = (int)floor((x - c->x0) / dxc + 1e-9), jc = (int)floor((y - c->y0) / dyc + 1e-9);
        if (ic >= c->Nx) ic = c->Nx - 1;
        if (jc >= c->Ny) jc = c->Ny - 1;
        const double a = (x - (c->x0 + ic * dxc)) / dxc, b = (y - (c->y0 + jc * dyc)) / dyc;
        const int t1 = a + b > 1.0 + 1e-9;
        const int64_t cell = 2 * ((int64_t)jc * c->Nx + ic) + t1;
        const double xi = t1 ? 1.0 - a : a, eta = t1 ? 1.0 - b : b;
        double phi[6], d1[6], d2[6];
        shape(c->r, xi, eta, phi, d1, d2);
        for (int q = 0; q < c->dpc; ++q) {
            const double w = rint(phi[q] * 1024.0) / 1024.0; /* weights are exact dyadic numbers */
            if (w != 0.0) { L->Pcol[k] = c->cell_dof[c->dpc * cell + q]; L->Pval[k] = w; ++k; }
        }
    }
    L->Pptr[f->n] = k;
}

static void mg_free(mg_t *m);
static void oracle_destroy_internal(oracle_problem *p);

/* s: the scheme matrix is M + s K; fineS: the fine level's BC-modified system matrix (borrowed) */
static int mg_setup(oracle_problem *p, double s, double *fineS) {
    if (p->mg) { mg_free(p->mg); p->mg = NULL; }
    mg_t *m = (mg_t *)calloc(1, sizeof(mg_t));
    m->nu = 2;
    m->nu_coarse = 8;
    m->lev[0].prob = p;
    m->lev[0].S = fineS;
    m->nlev = 1;
    int Nx = p->Nx, Ny = p->Ny, r = p->r;
    /* h-coarsening continues while the stiffness part still matters on the current level:
       s c^2 / (dx dy) > 1/4 with c taken at the centre of the domain */
    const double c0 = ev(p, EX_C, 0.5 * (p->x0 + p->x1), 0.5 * (p->y0 + p->y1), 0.0);
    for (;;) {
        int nNx = Nx, nNy = Ny, nr = 1;
        if (r == 2) nr = 1; /* p-coarsening first: same mesh, P1 */
        else {
            const double dx = (p->x1 - p->x0) / Nx, dy = (p->y1 - p->y0) / Ny;
            const int matters = s * c0 * c0 / (dx * dy) > 0.25;
            if (matters && Nx % 2 == 0 && Ny % 2 == 0 && (Nx / 2 < Ny / 2 ? Nx / 2 : Ny / 2) >= 2) { nNx = Nx / 2; nNy = Ny / 2; }
            else break;
        }
        if (m->nlev >= 12) break;
        oracle_problem *q = oracle_create(nNx, nNy, p->x0, p->x1, p->y0, p->y1, nr);
        q->ex[EX_C] = p->ex[EX_C];
        q->borrowed_expr = 1;
        oracle_setup(q);
        oracle_assemble(q);
        mg_level *L = &m->lev[m->nlev];
        L->prob = q;
        L->S = (double *)malloc(sizeof(double) * q->nnz);
        for (int64_t k = 0; k < q->nnz; ++k) L->S[k] = q->M[k] + s * q->K[k];
        double *zero = (double *)calloc(q->nb ? q->nb : 1, sizeof(double));
        apply_bc(q, L->S, zero, q->t1, q->t2);
        free(zero);
        mg_build_prolongation(m->lev[m->nlev - 1].prob, q, &m->lev[m->nlev - 1]);
        m->nlev++;
        Nx = nNx; Ny = nNy; r = nr;
    }
    for (int l = 0; l < m->nlev; ++l) {
        mg_level *L = &m->lev[l];
        const oracle_problem *q = L->prob;
        L->omega = q->r == 2 ? 0.5 : 0.8;
        L->dinv = (double *)malloc(sizeof(double) * q->n);
        L->x = (double *)calloc(q->n, sizeof(double));
        L->b = (double *)calloc(q->n, sizeof(double));
        L->r = (double *)calloc(q->n, sizeof(double));
        L->t = (double *)calloc(q->n, sizeof(double));
        if (l > 0)
            for (int64_t i = 0; i < q->n; ++i) L->dinv[i] = 1.0 / L->S[csr_find(q, (int)i, (int)i)];
    }
    {   /* fine level: p->S is rebuilt every step (copy + apply_boundary_values); its diagonal is that
           of bc(M + s K), formed here once */
        mg_level *L = &m->lev[0];
        double *tmpS = (double *)malloc(sizeof(double) * p->nnz);
        for (int64_t k = 0; k < p->nnz; ++k) tmpS[k] = p->M[k] + s * p->K[k];
        double *zero = (double *)calloc(p->nb ? p->nb : 1, sizeof(double));
        apply_bc(p, tmpS, zero, L->t, L->r);
        for (int64_t i = 0; i < p->n; ++i) L->dinv[i] = 1.0 / tmpS[csr_find(p, (int)i, (int)i)];
        memset(L->t, 0, sizeof(double) * p->n);
        memset(L->r, 0, sizeof(double) * p->n);
        free(zero);
        free(tmpS);
    }
    p->mg = m;
    return m->nlev;
}

/* x <- x + omega D^-1 (b - S x), `sweeps` times; the first sweep from x = 0 is x = omega D^-1 b */
static void mg_smooth(mg_level *L, int sweeps, int x_is_zero) {
    const oracle_problem *q = L->prob;
    for (int sw = 0; sw < sweeps; ++sw) {
        if (sw == 0 && x_is_zero) {
            for (int64_t i = 0; i < q->n; ++i) L->x[i] = L->omega * (L->dinv[i] * L->b[i]);
            continue;
        }
        spmv(q, L->S, L->x, L->t);
        for (int64_t i = 0; i < q->n; ++i) L->x[i] = L->x[i] + L->omega * (L->dinv[i] * (L->b[i] - L->t[i]));
    }
}
static void mg_vcycle(mg_t *m, int l) {
    mg_level *L = &m->lev[l];
    const oracle_problem *q = L->prob;
    if (l == m->nlev - 1) { mg_smooth(L, m->nu_coarse, 1); return; }
    mg_smooth(L, m->nu, 1);
    spmv(q, L->S, L->x, L->t);
    for (int64_t i = 0; i < q->n; ++i) L->r[i] = L->b[i] - L->t[i];
    /* restriction = transpose of the prolongation; Dirichlet rows of the coarse level get 0 */
    mg_level *C = &m->lev[l + 1];
    const oracle_problem *qc = C->prob;
    memset(C->b, 0, sizeof(double) * qc->n);
    for (int64_t i = 0; i < q->n; ++i)
        for (int64_t k = L->Pptr[i]; k < L->Pptr[i + 1]; ++k) C->b[L->Pcol[k]] += L->Pval[k] * L->r[i];
    for (int64_t b = 0; b < qc->nb; ++b) C->b[qc->bdof[b]] = 0.0;
    mg_vcycle(m, l + 1);
    for (int64_t i = 0; i < q->n; ++i) {
        double e = 0.0;
        for (int64_t k = L->Pptr[i]; k < L->Pptr[i + 1]; ++k) e += L->Pval[k] * C->x[L->Pcol[k]];
        L->x[i] += e;
    }
    mg_smooth(L, m->nu, 0);
}
/* h = V-cycle(g) on the fine level */
static void mg_apply(oracle_problem *p, const double *g, double *h) {
    mg_t *m = p->mg;
    memcpy(m->lev[0].b, g, sizeof(double) * p->n);
    mg_vcycle(m, 0);
    memcpy(h, m->lev[0].x, sizeof(double) * p->n);
}
static void mg_free(mg_t *m) {
    for (int l = 0; l < m->nlev; ++l) {
        mg_level *L = &m->lev[l];
        free(L->dinv); free(L->x); free(L->b); free(L->r); free(L->t);
        free(L->Pptr); free(L->Pcol); free(L->Pval);
        if (l > 0) { free(L->S); oracle_destroy_internal(L->prob); }
    }
    free(m);
}
int oracle_mg_levels(oracle_problem *p) { return p->mg ? p->mg->nlev : 0; }

/* [deal.II] SolverCG::solve + ReductionControl (src/WaveNewmark.cpp:256-261).
   Preconditioner: Jacobi (north-star replacement of PreconditionAMG/SSOR). */
static int cg_solve_ex(oracle_problem *p, const double *Aval, double *x, const double *b, int use_mg);
static int cg_solve(oracle_problem *p, const double *Aval, double *x, const double *b) {
    return cg_solve_ex(p, Aval, x, b, 0);
}
static int cg_solve_ex(oracle_problem *p, const double *Aval, double *x, const double *b, int use_mg) {
    double *g = p->cg_g, *d = p->cg_d, *h = p->cg_h;
    const int64_t n = p->n;
    use_mg = use_mg && p->precond == 2 && p->mg != NULL;
#pragma omp parallel for schedule(static) if ((n) > OMP_MIN_WORK)
    for (int64_t i = 0; i < n; ++i) {
        double dg = 1.0;
        for (int64_t k = p->rowptr[i]; k < p->rowptr[i + 1]; ++k)
            if (p->col[k] == i) dg = Aval[k];
        p->dinv[i] = p->precond == 1 ? 1.0 : 1.0 / dg;
    }
    spmv(p, Aval, x, g);
    axpy(p, -1.0, b, g);
    double res = sqrt(dot(p, g, g));
    const double reduced_tol = res * p->cg_reduce;
    int it = 0;
    if (res <= reduced_tol || res <= p->cg_tol) return 0;
    if (use_mg) {
        mg_apply(p, g, h);
        for (int64_t i = 0; i < n; ++i) d[i] = -h[i];
    } else {
#pragma omp parallel for schedule(static) if ((n) > OMP_MIN_WORK)
        for (int64_t i = 0; i < n; ++i) { h[i] = p->dinv[i] * g[i]; d[i] = -h[i]; }
    }
    double gh = dot(p, g, h);
    for (;;) {
        it++;
        spmv(p, Aval, d, h);
        double alpha = gh / dot(p, d, h);
        axpy(p, alpha, d, x);
        axpy(p, alpha, h, g);
        res = sqrt(fabs(dot(p, g, g)));
        if (res <= reduced_tol || res <= p->cg_tol) return it;
        if (it >= p->cg_maxit || isnan(res)) return -it;
        if (use_mg)
            mg_apply(p, g, h);
        else {
#pragma omp parallel for schedule(static) if ((n) > OMP_MIN_WORK)
            for (int64_t i = 0; i < n; ++i) h[i] = p->dinv[i] * g[i];
        }
        double beta = gh;
        gh = dot(p, g, h);
        beta = gh / beta;
#pragma omp parallel for schedule(static) if ((n) > OMP_MIN_WORK)
        for (int64_t i = 0; i < n; ++i) d[i] = beta * d[i] - h[i];
    }
}

void oracle_set_cg(oracle_problem *p, int maxit, double tol, double reduce, int precond) {
    p->cg_maxit = maxit; p->cg_tol = tol; p->cg_reduce = reduce; p->precond = precond;
}
void oracle_set_forcing_every_step(oracle_problem *p, int flag) { p->forcing_every_step = flag; }

static void interp(oracle_problem *p, int which, double t, double *out) {
    for (int64_t i = 0; i < p->n; ++i) out[i] = ev(p, which, p->sx[i], p->sy[i], t);
}

/* ------------------------------------------------------------------------- */
/* Newmark: src/WaveNewmark.cpp:280-456                                        */
/* ------------------------------------------------------------------------- */
int oracle_newmark_init(oracle_problem *p, double dt, double beta, double gamma) {
    p->scheme = 0; p->dt = dt; p->beta = beta; p->gamma = gamma; p->time = 0.0; p->step = 0;
    /* :110-112 matrix_a = M + beta dt^2 K */
    for (int64_t k = 0; k < p->nnz; ++k) p->A[k] = p->M[k] + beta * dt * dt * p->K[k];
    /* :292-296 */
    interp(p, EX_U0, 0.0, p->ou);
    interp(p, EX_V0, 0.0, p->ov);
    vcopy(p, p->ou, p->u);
    vcopy(p, p->ov, p->v);
    /* :300-390 consistent a0: M a0 = F(0) - K u0 */
    spmv(p, p->K, p->ou, p->rhs);
    for (int64_t i = 0; i < p->n; ++i) p->rhs[i] *= -1.0;
    memset(p->t1, 0, sizeof(double) * p->n);
    add_forcing(p, p->t1, 1.0, 0.0, 0.0, 1.0, 0.0, 0);
    axpy(p, 1.0, p->t1, p->rhs);
    double *bv = (double *)malloc(sizeof(double) * (p->nb ? p->nb : 1));
    const double inv_dt2 = 1.0 / (dt * dt);
    for (int64_t b = 0; b < p->nb; ++b) { /* :348-367 */
        int i = p->bdof[b];
        double gp = ev(p, EX_G, p->sx[i], p->sy[i], dt);
        double g0 = ev(p, EX_G, p->sx[i], p->sy[i], 0.0);
        double gm = ev(p, EX_G, p->sx[i], p->sy[i], -dt);
        bv[b] = (gp - 2.0 * g0 + gm) * inv_dt2;
    }
    memcpy(p->S, p->M, sizeof(double) * p->nnz); /* :372 */
    memset(p->oa, 0, sizeof(double) * p->n);
    apply_bc(p, p->S, bv, p->oa, p->rhs);
    int its = cg_solve(p, p->S, p->oa, p->rhs); /* :378-385 (SSOR there; Jacobi here) */
    vcopy(p, p->oa, p->a);
    p->last_its[0] = its; p->last_its[1] = 0;
    free(bv);
    if (p->precond == 2 && beta * dt * dt > 0.0) mg_setup(p, beta * dt * dt, p->S);
    return its < 0 ? -1 : 0;
}

int oracle_newmark_step(oracle_problem *p) {
    const double dt = p->dt, beta = p->beta, gamma = p->gamma;
    p->time += dt; /* :409-410 */
    p->step++;
    /* assemble_rhs :116-175 */
    memset(p->rhs, 0, sizeof(double) * p->n);
    double *z = p->t1, *w = p->t2, *rf = p->t3;
    vcopy(p, p->ou, z);
    axpy(p, dt, p->ov, z);
    axpy(p, dt * dt * (0.5 - beta), p->oa, z);
    spmv(p, p->K, z, w);
    axpy(p, -1.0, w, p->rhs);
    memset(rf, 0, sizeof(double) * p->n);
    if (p->forcing_every_step) add_forcing(p, rf, 1.0, p->time, 0.0, 1.0, 0.0, 0);
    axpy(p, 1.0, rf, p->rhs);
    /* solve_a :177-262 */
    memcpy(p->S, p->A, sizeof(double) * p->nnz);
    double *bv = (double *)malloc(sizeof(double) * (p->nb ? p->nb : 1));
    if (beta > 1e-12) {
        const double beta_dt2 = beta * dt * dt;
        for (int64_t b = 0; b < p->nb; ++b) {
            int i = p->bdof[b];
            double gv = ev(p, EX_G, p->sx[i], p->sy[i], p->time);
            double u_pred = p->ou[i] + dt * p->ov[i] + dt * dt * (0.5 - beta) * p->oa[i];
            bv[b] = (gv - u_pred) / beta_dt2;
        }
    } else {
        const double inv_dt2 = 1.0 / (dt * dt);
        for (int64_t b = 0; b < p->nb; ++b) {
            int i = p->bdof[b];
            double g1 = ev(p, EX_G, p->sx[i], p->sy[i], p->time);
            double g0 = ev(p, EX_G, p->sx[i], p->sy[i], p->time - dt);
            double gm = ev(p, EX_G, p->sx[i], p->sy[i], p->time - 2.0 * dt);
            bv[b] = (g1 - 2.0 * g0 + gm) * inv_dt2;
        }
    }
    apply_bc(p, p->S, bv, p->a, p->rhs);
    free(bv);
    int its = cg_solve_ex(p, p->S, p->a, p->rhs, 1);
    p->last_its[0] = its; p->last_its[1] = 0;
    /* update_u_v :264-278 */
    vcopy(p, p->ou, p->u);
    axpy(p, dt, p->ov, p->u);
    axpy(p, dt * dt * (0.5 - beta), p->oa, p->u);
    axpy(p, dt * dt * beta, p->a, p->u);
    vcopy(p, p->ov, p->v);
    axpy(p, dt * (1.0 - gamma), p->oa, p->v);
    axpy(p, dt * gamma, p->a, p->v);
    /* :438-440 */
    vcopy(p, p->u, p->ou);
    vcopy(p, p->v, p->ov);
    vcopy(p, p->a, p->oa);
    return its < 0 ? -1 : 0;
}

/* ------------------------------------------------------------------------- */
/* Theta: src/WaveTheta.cpp:341-411                                            */
/* ------------------------------------------------------------------------- */
int oracle_theta_init(oracle_problem *p, double dt, double theta) {
    p->scheme = 1; p->dt = dt; p->theta = theta; p->time = 0.0; p->step = 0;
    for (int64_t k = 0; k < p->nnz; ++k) { /* :110-115 */
        p->A[k] = p->M[k] + (theta * dt) * (theta * dt) * p->K[k];
        p->A2[k] = p->M[k];
    }
    interp(p, EX_U0, 0.0, p->ou);
    interp(p, EX_V0, 0.0, p->ov);
    vcopy(p, p->ou, p->u);
    vcopy(p, p->ov, p->v);
    if (p->precond == 2 && theta * dt > 0.0) mg_setup(p, (theta * dt) * (theta * dt), p->S);
    return 0;
}

int oracle_theta_step(oracle_problem *p) {
    const double dt = p->dt, theta = p->theta;
    p->time += dt;
    p->step++;
    double *bv = (double *)malloc(sizeof(double) * (p->nb ? p->nb : 1));
    double *tmp = p->t1, *rf = p->t3;
    /* assemble_rhs_u :119-186 */
    spmv(p, p->M, p->ou, p->rhs);
    spmv(p, p->K, p->ou, tmp);
    axpy(p, -dt * dt * theta * (1 - theta), tmp, p->rhs);
    spmv(p, p->M, p->ov, tmp);
    axpy(p, dt, tmp, p->rhs);
    memset(rf, 0, sizeof(double) * p->n);
    if (p->forcing_every_step)
        add_forcing(p, rf, theta * dt * dt, p->time, p->time - dt, theta, 1.0 - theta, 1);
    axpy(p, 1.0, rf, p->rhs);
    /* solve_u :251-294 */
    memcpy(p->S, p->A, sizeof(double) * p->nnz);
    for (int64_t b = 0; b < p->nb; ++b) {
        int i = p->bdof[b];
        bv[b] = ev(p, EX_G, p->sx[i], p->sy[i], p->time);
    }
    apply_bc(p, p->S, bv, p->u, p->rhs);
    int its_u = cg_solve_ex(p, p->S, p->u, p->rhs, 1);
    /* assemble_rhs_v :188-249 */
    spmv(p, p->M, p->ov, p->rhs);
    spmv(p, p->K, p->ou, tmp);
    axpy(p, -dt * (1.0 - theta), tmp, p->rhs);
    spmv(p, p->K, p->u, tmp);
    axpy(p, -dt * theta, tmp, p->rhs);
    memset(rf, 0, sizeof(double) * p->n);
    if (p->forcing_every_step)
        add_forcing(p, rf, dt, p->time, p->time - dt, theta, 1.0 - theta, 1);
    axpy(p, 1.0, rf, p->rhs);
    /* solve_v :296-339 */
    memcpy(p->S, p->A2, sizeof(double) * p->nnz);
    for (int64_t b = 0; b < p->nb; ++b) {
        int i = p->bdof[b];
        bv[b] = ev(p, EX_DGDT, p->sx[i], p->sy[i], p->time);
    }
    apply_bc(p, p->S, bv, p->v, p->rhs);
    int its_v = cg_solve(p, p->S, p->v, p->rhs);
    free(bv);
    p->last_its[0] = its_u; p->last_its[1] = its_v;
    vcopy(p, p->u, p->ou); /* :394-395 */
    vcopy(p, p->v, p->ov);
    return (its_u < 0 || its_v < 0) ? -1 : 0;
}

/* src/WaveEquationBase.cpp:148-154 E = 1/2 (v^T M v + u^T K u) */
double oracle_energy(oracle_problem *p) {
    spmv(p, p->K, p->u, p->t1);
    spmv(p, p->M, p->v, p->t2);
    return 0.5 * (dot(p, p->t2, p->v) + dot(p, p->t1, p->u));
}

/* src/WaveEquationBase.cpp:367-423 compute_error / compute_relative_error.
   [deal.II] VectorTools::integrate_difference with QGaussSimplex(r+2), affine mapping;
   H1_norm per cell = sqrt(L2^2 + seminorm^2); exact gradient: FunctionParser derives
   from AutoDerivativeFunction (h = 1e-8, centred difference "Euler").  Numerator cells
   are stored as float (:384), denominator cells as double (:412); global = sqrt(sum sq).
   out = {L2, H1, relL2, relH1}. */
int oracle_errors(oracle_problem *p, double t, double *out) {
    if (!p->ex[EX_SOL].defined) return -1;
    const int dpc = p->dpc, nq = p->qerr.nq;
    const double hfd = 1e-8;
    double e_l2 = 0, e_h1 = 0, n_l2 = 0, n_h1 = 0;
#pragma omp parallel for reduction(+ : e_l2, e_h1, n_l2, n_h1) schedule(static) if ((p->ncells) > OMP_MIN_WORK)
    for (int64_t c = 0; c < p->ncells; ++c) {
        double X0, Y0, J[4], det;
        cell_geom(p, c, &X0, &Y0, J, &det);
        const double adet = fabs(det);
        const double i00 = J[3] / det, i01 = -J[2] / det, i10 = -J[1] / det, i11 = J[0] / det;
        const int *cd = p->cell_dof + dpc * c;
        double dl2 = 0, dsemi = 0, xl2 = 0, xsemi = 0;
        for (int q = 0; q < nq; ++q) {
            double phi[6], dxi[6], deta[6];
            shape(p->r, p->qerr.xi[q], p->qerr.eta[q], phi, dxi, deta);
            const double JxW = p->qerr.w[q] * adet;
            const double xq = X0 + J[0] * p->qerr.xi[q] + J[1] * p->qerr.eta[q];
            const double yq = Y0 + J[2] * p->qerr.xi[q] + J[3] * p->qerr.eta[q];
            double uh = 0, ux = 0, uy = 0;
            for (int i = 0; i < dpc; ++i) {
                const double ui = p->u[cd[i]];
                uh += ui * phi[i];
                ux += ui * (i00 * dxi[i] + i01 * deta[i]);
                uy += ui * (i10 * dxi[i] + i11 * deta[i]);
            }
            const double ue = ev(p, EX_SOL, xq, yq, t);
            const double uex = (ev(p, EX_SOL, xq + hfd, yq, t) - ev(p, EX_SOL, xq - hfd, yq, t)) / (2 * hfd);
            const double uey = (ev(p, EX_SOL, xq, yq + hfd, t) - ev(p, EX_SOL, xq, yq - hfd, t)) / (2 * hfd);
            dl2 += (uh - ue) * (uh - ue) * JxW;
            dsemi += ((ux - uex) * (ux - uex) + (uy - uey) * (uy - uey)) * JxW;
            xl2 += ue * ue * JxW;
            xsemi += (uex * uex + uey * uey) * JxW;
        }
        const float fl2 = (float)sqrt(dl2), fh1 = (float)sqrt(dl2 + dsemi);
        e_l2 += (double)fl2 * (double)fl2;
        e_h1 += (double)fh1 * (double)fh1;
        n_l2 += xl2;
        n_h1 += xl2 + xsemi;
    }
    out[0] = sqrt(e_l2); out[1] = sqrt(e_h1);
    const double nl2 = sqrt(n_l2), nh1 = sqrt(n_h1);
    out[2] = nl2 < 1e-14 ? out[0] : out[0] / nl2;
    out[3] = nh1 < 1e-14 ? out[1] : out[1] / nh1;
    return 0;
}

/* src/WaveEquationBase.cpp:170-193 probe: u_h at the box centre ([deal.II] point_value) */
double oracle_probe(oracle_problem *p) {
    const double px = 0.5 * (p->x0 + p->x1), py = 0.5 * (p->y0 + p->y1);
    for (int64_t c = 0; c < p->ncells; ++c) {
        double X0, Y0, J[4], det;
        cell_geom(p, c, &X0, &Y0, J, &det);
        const double rx = px - X0, ry = py - Y0;
        const double xi = (J[3] * rx - J[1] * ry) / det, eta = (-J[2] * rx + J[0] * ry) / det;
        const double eps = 1e-12;
        if (xi >= -eps && eta >= -eps && xi + eta <= 1.0 + eps) {
            double phi[6], dxi[6], deta[6];
            shape(p->r, xi, eta, phi, dxi, deta);
            double s = 0;
            for (int i = 0; i < p->dpc; ++i) s += p->u[p->cell_dof[p->dpc * c + i]] * phi[i];
            return s;
        }
    }
    return 0.0;
}

/* ------------------------------------------------------------------------- */
/* accessors                                                                  */
/* ------------------------------------------------------------------------- */
void oracle_destroy(oracle_problem *p);
int64_t oracle_n(oracle_problem *p) { return p->n; }
int64_t oracle_nnz(oracle_problem *p) { return p->nnz; }
int64_t oracle_nb(oracle_problem *p) { return p->nb; }
int64_t oracle_ncells(oracle_problem *p) { return p->ncells; }
double oracle_time(oracle_problem *p) { return p->time; }
void oracle_last_iterations(oracle_problem *p, int *its) { its[0] = p->last_its[0]; its[1] = p->last_its[1]; }
void oracle_get_csr(oracle_problem *p, int64_t *rowptr, int *col) {
    memcpy(rowptr, p->rowptr, sizeof(int64_t) * (p->n + 1));
    memcpy(col, p->col, sizeof(int) * p->nnz);
}
/* which: 0 M, 1 K, 2 A (matrix_a / matrix_u), 3 A2 (matrix_v), 4 S (last BC-modified system matrix) */
int oracle_get_values(oracle_problem *p, int which, double *val) {
    const double *src[] = {p->M, p->K, p->A, p->A2, p->S};
    if (which < 0 || which > 4) return -1;
    memcpy(val, src[which], sizeof(double) * p->nnz);
    return 0;
}
/* which: 0 u, 1 v, 2 a, 3 rhs */
int oracle_get_vector(oracle_problem *p, int which, double *out) {
    const double *src[] = {p->u, p->v, p->a, p->rhs};
    if (which < 0 || which > 3) return -1;
    memcpy(out, src[which], sizeof(double) * p->n);
    return 0;
}
int oracle_set_vector(oracle_problem *p, int which, const double *in) {
    double *dst[] = {p->u, p->v, p->a};
    double *odst[] = {p->ou, p->ov, p->oa};
    if (which < 0 || which > 2) return -1;
    memcpy(dst[which], in, sizeof(double) * p->n);
    memcpy(odst[which], in, sizeof(double) * p->n);
    return 0;
}
void oracle_get_cell_dofs(oracle_problem *p, int *out) { memcpy(out, p->cell_dof, sizeof(int) * p->dpc * p->ncells); }
void oracle_get_support_points(oracle_problem *p, double *sx, double *sy) {
    memcpy(sx, p->sx, sizeof(double) * p->n);
    memcpy(sy, p->sy, sizeof(double) * p->n);
}
void oracle_get_boundary_dofs(oracle_problem *p, int *out) { memcpy(out, p->bdof, sizeof(int) * p->nb); }
/* y = val[which] * x */
int oracle_spmv(oracle_problem *p, int which, const double *x, double *y) {
    const double *src[] = {p->M, p->K, p->A, p->A2, p->S};
    if (which < 0 || which > 4) return -1;
    spmv(p, src[which], x, y);
    return 0;
}
/* standalone CG on val[which] (used by kernel-level parity tests) */
int oracle_cg(oracle_problem *p, int which, double *x, const double *b) {
    const double *src[] = {p->M, p->K, p->A, p->A2, p->S};
    return cg_solve(p, src[which], x, b);
}
double oracle_norm(oracle_problem *p, int which) {
    const double *src[] = {p->u, p->v, p->a, p->rhs};
    return sqrt(dot(p, src[which], src[which]));
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
static void oracle_destroy_internal(oracle_problem *p) { oracle_destroy(p); }
void oracle_destroy(oracle_problem *p) {
    if (!p) return;
    if (p->mg) mg_free(p->mg);
    if (!p->borrowed_expr)
        for (int i = 0; i < EX_COUNT; ++i) free_node(p->ex[i].root);
    free(p->vx); free(p->vy); free(p->cell_v); free(p->cell_dof); free(p->sx); free(p->sy);
    free(p->bdof); free(p->is_b); free(p->rowptr); free(p->col);
    free(p->M); free(p->K); free(p->A); free(p->A2); free(p->S);
    free(p->u); free(p->v); free(p->a); free(p->ou); free(p->ov); free(p->oa); free(p->rhs);
    free(p->t1); free(p->t2); free(p->t3); free(p->cg_g); free(p->cg_d); free(p->cg_h); free(p->dinv);
    free(p);
}
