// expr_vm.h -- stack-machine bytecode for the parameter-file expressions, evaluated on the
// device (and on the host for constant folding / wave_eval_expr).
//
// Replaces deal.II FunctionParser::value (muParser) as used by the reference for C, F, U0, V0,
// G, DGDT and Solution (src/main-newmark.cpp:64-79, src/ParameterReader.cpp:139-175).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define WV_HD __host__ __device__ __forceinline__
#else
#define WV_HD inline
#endif

namespace wv {

enum Op : int32_t {
    OP_CONST = 0, OP_VAR, OP_NEG, OP_NOT, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_POWI,
    OP_LT, OP_LE, OP_GT, OP_GE, OP_EQ, OP_NE, OP_AND, OP_OR, OP_SELECT, OP_F1, OP_F2
};
enum Fn : int32_t {
    FN_SIN = 0, FN_COS, FN_TAN, FN_ASIN, FN_ACOS, FN_ATAN, FN_SINH, FN_COSH, FN_TANH, FN_ASINH,
    FN_ACOSH, FN_ATANH, FN_EXP, FN_LOG, FN_LOG2, FN_LOG10, FN_SQRT, FN_ABS, FN_SIGN, FN_RINT,
    FN_FLOOR, FN_CEIL, FN_ERFC, FN_COT, FN_CSC, FN_SEC, FN_MIN, FN_MAX, FN_POW
};

struct Instr {
    int32_t op;
    int32_t arg;
    double val;
};

constexpr int kMaxProgram = 192;
constexpr int kMaxStack = 24;

// A compiled expression as the kernels see it (plain data, passed by pointer in global memory).
struct Program {
    int32_t len;
    int32_t time_dependent;
    Instr code[kMaxProgram];
};

WV_HD double apply_f1(int fn, double a) {
    switch (fn) {
    case FN_SIN: return sin(a);
    case FN_COS: return cos(a);
    case FN_TAN: return tan(a);
    case FN_ASIN: return asin(a);
    case FN_ACOS: return acos(a);
    case FN_ATAN: return atan(a);
    case FN_SINH: return sinh(a);
    case FN_COSH: return cosh(a);
    case FN_TANH: return tanh(a);
    case FN_ASINH: return asinh(a);
    case FN_ACOSH: return acosh(a);
    case FN_ATANH: return atanh(a);
    case FN_EXP: return exp(a);
    case FN_LOG: return log(a);
    case FN_LOG2: return log2(a);
    case FN_LOG10: return log10(a);
    case FN_SQRT: return sqrt(a);
    case FN_ABS: return fabs(a);
    case FN_SIGN: return a > 0.0 ? 1.0 : (a < 0.0 ? -1.0 : 0.0);
    case FN_RINT: return rint(a);
    case FN_FLOOR: return floor(a);
    case FN_CEIL: return ceil(a);
    case FN_ERFC: return erfc(a);
    case FN_COT: return 1.0 / tan(a);
    case FN_CSC: return 1.0 / sin(a);
    case FN_SEC: return 1.0 / cos(a);
    default: return NAN;
    }
}

WV_HD double powi(double a, int n) {
    // exact repeated squaring for the small integer exponents the parameter files use (x^2)
    bool inv = n < 0;
    unsigned m = inv ? (unsigned)(-n) : (unsigned)n;
    double r = 1.0, b = a;
    while (m) {
        if (m & 1u) r *= b;
        b *= b;
        m >>= 1;
    }
    return inv ? 1.0 / r : r;
}

WV_HD double eval(const Program *p, double x, double y, double t) {
    double st[kMaxStack];
    int sp = 0;
    const int n = p->len;
    for (int pc = 0; pc < n; ++pc) {
        const Instr in = p->code[pc];
        switch (in.op) {
        case OP_CONST: st[sp++] = in.val; break;
        case OP_VAR: st[sp++] = in.arg == 0 ? x : (in.arg == 1 ? y : t); break;
        case OP_NEG: st[sp - 1] = -st[sp - 1]; break;
        case OP_NOT: st[sp - 1] = st[sp - 1] == 0.0 ? 1.0 : 0.0; break;
        case OP_ADD: --sp; st[sp - 1] = st[sp - 1] + st[sp]; break;
        case OP_SUB: --sp; st[sp - 1] = st[sp - 1] - st[sp]; break;
        case OP_MUL: --sp; st[sp - 1] = st[sp - 1] * st[sp]; break;
        case OP_DIV: --sp; st[sp - 1] = st[sp - 1] / st[sp]; break;
        case OP_POW: --sp; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
        case OP_POWI: st[sp - 1] = powi(st[sp - 1], in.arg); break;
        case OP_LT: --sp; st[sp - 1] = st[sp - 1] < st[sp] ? 1.0 : 0.0; break;
        case OP_LE: --sp; st[sp - 1] = st[sp - 1] <= st[sp] ? 1.0 : 0.0; break;
        case OP_GT: --sp; st[sp - 1] = st[sp - 1] > st[sp] ? 1.0 : 0.0; break;
        case OP_GE: --sp; st[sp - 1] = st[sp - 1] >= st[sp] ? 1.0 : 0.0; break;
        case OP_EQ: --sp; st[sp - 1] = st[sp - 1] == st[sp] ? 1.0 : 0.0; break;
        case OP_NE: --sp; st[sp - 1] = st[sp - 1] != st[sp] ? 1.0 : 0.0; break;
        case OP_AND: --sp; st[sp - 1] = (st[sp - 1] != 0.0 && st[sp] != 0.0) ? 1.0 : 0.0; break;
        case OP_OR: --sp; st[sp - 1] = (st[sp - 1] != 0.0 || st[sp] != 0.0) ? 1.0 : 0.0; break;
        case OP_SELECT: sp -= 2; st[sp - 1] = st[sp - 1] != 0.0 ? st[sp] : st[sp + 1]; break;
        case OP_F1: st[sp - 1] = apply_f1(in.arg, st[sp - 1]); break;
        case OP_F2: {
            --sp;
            const double a = st[sp - 1], b = st[sp];
            st[sp - 1] = in.arg == FN_MIN ? (a < b ? a : b) : (in.arg == FN_MAX ? (a > b ? a : b) : pow(a, b));
            break;
        }
        default: break;
        }
    }
    return st[0];
}

}  // namespace wv
