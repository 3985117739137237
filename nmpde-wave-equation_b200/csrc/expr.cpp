// expr.cpp -- recursive-descent compiler for the muParser subset the reference's parameter files
// and deal.II's FunctionParser accept.  AST -> constant folding -> RPN bytecode.
#include "expr.hpp"

#include <cctype>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace wv {
namespace {

struct Ast {
    int op = OP_CONST;
    int arg = 0;
    double val = 0.0;
    std::vector<std::unique_ptr<Ast>> kids;
};
using AstP = std::unique_ptr<Ast>;

AstP leaf(double v) {
    auto n = std::make_unique<Ast>();
    n->op = OP_CONST;
    n->val = v;
    return n;
}
AstP node(int op, int arg, AstP a, AstP b = nullptr, AstP c = nullptr) {
    auto n = std::make_unique<Ast>();
    n->op = op;
    n->arg = arg;
    n->kids.push_back(std::move(a));
    if (b) n->kids.push_back(std::move(b));
    if (c) n->kids.push_back(std::move(c));
    return n;
}

struct FnEntry { const char *name; int fn; int nargs; };
const FnEntry kFns[] = {
    {"sin", FN_SIN, 1},     {"cos", FN_COS, 1},     {"tan", FN_TAN, 1},     {"asin", FN_ASIN, 1},
    {"acos", FN_ACOS, 1},   {"atan", FN_ATAN, 1},   {"sinh", FN_SINH, 1},   {"cosh", FN_COSH, 1},
    {"tanh", FN_TANH, 1},   {"asinh", FN_ASINH, 1}, {"acosh", FN_ACOSH, 1}, {"atanh", FN_ATANH, 1},
    {"exp", FN_EXP, 1},     {"log", FN_LOG, 1},     {"ln", FN_LOG, 1},      {"log2", FN_LOG2, 1},
    {"log10", FN_LOG10, 1}, {"sqrt", FN_SQRT, 1},   {"abs", FN_ABS, 1},     {"sign", FN_SIGN, 1},
    {"rint", FN_RINT, 1},   {"int", FN_RINT, 1},    {"floor", FN_FLOOR, 1}, {"ceil", FN_CEIL, 1},
    {"erfc", FN_ERFC, 1},   {"cot", FN_COT, 1},     {"csc", FN_CSC, 1},     {"sec", FN_SEC, 1},
    {"min", FN_MIN, 2},     {"max", FN_MAX, 2},     {"pow", FN_POW, 2}};

class Parser {
  public:
    Parser(const std::string &s, const std::vector<std::string> &vars,
           const std::map<std::string, double> &consts)
        : s_(s), vars_(vars), consts_(consts) {}

    AstP parse() {
        AstP e = expr();
        ws();
        if (pos_ < s_.size()) fail("unexpected trailing input");
        return e;
    }

  private:
    const std::string &s_;
    const std::vector<std::string> &vars_;
    const std::map<std::string, double> &consts_;
    size_t pos_ = 0;

    [[noreturn]] void fail(const std::string &msg) const {
        std::ostringstream o;
        o << "expression '" << s_ << "': " << msg << " at position " << pos_;
        throw std::invalid_argument(o.str());
    }
    void ws() {
        while (pos_ < s_.size() && std::isspace((unsigned char)s_[pos_])) ++pos_;
    }
    char peek(size_t k = 0) const { return pos_ + k < s_.size() ? s_[pos_ + k] : '\0'; }
    bool eat(char c) {
        ws();
        if (peek() == c) { ++pos_; return true; }
        return false;
    }

    AstP expr() {  // ternary
        AstP c = lor();
        ws();
        if (peek() == '?') {
            ++pos_;
            AstP a = expr();
            if (!eat(':')) fail("expected ':'");
            AstP b = expr();
            return node(OP_SELECT, 0, std::move(c), std::move(a), std::move(b));
        }
        return c;
    }
    AstP lor() {
        AstP l = land();
        for (;;) {
            ws();
            if (peek() == '|' && peek(1) == '|') { pos_ += 2; l = node(OP_OR, 0, std::move(l), land()); }
            else return l;
        }
    }
    AstP land() {
        AstP l = cmp();
        for (;;) {
            ws();
            if (peek() == '&' && peek(1) == '&') { pos_ += 2; l = node(OP_AND, 0, std::move(l), cmp()); }
            else return l;
        }
    }
    AstP cmp() {
        AstP l = add();
        for (;;) {
            ws();
            int op = -1, len = 0;
            if (peek() == '<' && peek(1) == '=') { op = OP_LE; len = 2; }
            else if (peek() == '>' && peek(1) == '=') { op = OP_GE; len = 2; }
            else if (peek() == '=' && peek(1) == '=') { op = OP_EQ; len = 2; }
            else if (peek() == '!' && peek(1) == '=') { op = OP_NE; len = 2; }
            else if (peek() == '<') { op = OP_LT; len = 1; }
            else if (peek() == '>') { op = OP_GT; len = 1; }
            if (op < 0) return l;
            pos_ += len;
            l = node(op, 0, std::move(l), add());
        }
    }
    AstP add() {
        AstP l = mul();
        for (;;) {
            ws();
            if (peek() == '+') { ++pos_; l = node(OP_ADD, 0, std::move(l), mul()); }
            else if (peek() == '-') { ++pos_; l = node(OP_SUB, 0, std::move(l), mul()); }
            else return l;
        }
    }
    AstP mul() {
        AstP l = unary();
        for (;;) {
            ws();
            if (peek() == '*') { ++pos_; l = node(OP_MUL, 0, std::move(l), unary()); }
            else if (peek() == '/') { ++pos_; l = node(OP_DIV, 0, std::move(l), unary()); }
            else return l;
        }
    }
    AstP unary() {  // sign binds looser than '^': -a^2 == -(a^2)
        ws();
        if (peek() == '-') { ++pos_; return node(OP_NEG, 0, unary()); }
        if (peek() == '+') { ++pos_; return unary(); }
        if (peek() == '!' && peek(1) != '=') { ++pos_; return node(OP_NOT, 0, unary()); }
        return power();
    }
    AstP power() {  // right associative, signed exponent allowed
        AstP base = primary();
        ws();
        if (peek() == '^') {
            ++pos_;
            return node(OP_POW, 0, std::move(base), unary());
        }
        return base;
    }
    AstP primary() {
        ws();
        const char c = peek();
        if (c == '(') {
            ++pos_;
            AstP e = expr();
            if (!eat(')')) fail("expected ')'");
            return e;
        }
        if (std::isdigit((unsigned char)c) || c == '.') {
            const char *begin = s_.c_str() + pos_;
            char *end = nullptr;
            const double v = std::strtod(begin, &end);
            if (end == begin) fail("malformed number");
            pos_ += (size_t)(end - begin);
            return leaf(v);
        }
        if (std::isalpha((unsigned char)c) || c == '_') {
            std::string id;
            while (std::isalnum((unsigned char)peek()) || peek() == '_') id.push_back(s_[pos_++]);
            ws();
            if (peek() == '(') {
                ++pos_;
                std::vector<AstP> args;
                for (;;) {
                    args.push_back(expr());
                    if (eat(',')) continue;
                    if (eat(')')) break;
                    fail("expected ',' or ')'");
                }
                if (id == "if") {
                    if (args.size() != 3) fail("if() takes three arguments");
                    return node(OP_SELECT, 0, std::move(args[0]), std::move(args[1]), std::move(args[2]));
                }
                for (const auto &f : kFns)
                    if (id == f.name) {
                        if ((int)args.size() != f.nargs) fail("wrong number of arguments to " + id);
                        if (f.nargs == 1) return node(OP_F1, f.fn, std::move(args[0]));
                        if (f.fn == FN_POW) return node(OP_POW, 0, std::move(args[0]), std::move(args[1]));
                        return node(OP_F2, f.fn, std::move(args[0]), std::move(args[1]));
                    }
                fail("unknown function '" + id + "'");
            }
            for (size_t v = 0; v < vars_.size(); ++v)
                if (id == vars_[v]) {
                    auto n = std::make_unique<Ast>();
                    n->op = OP_VAR;
                    n->arg = (int)v;
                    return n;
                }
            auto it = consts_.find(id);
            if (it != consts_.end()) return leaf(it->second);
            fail("unknown identifier '" + id + "'");
        }
        fail(std::string("unexpected character '") + (c ? c : '$') + "'");
    }
};

void emit(const Ast &a, Program &p, int &depth, int &maxdepth);

bool all_const(const Ast &a) {
    for (const auto &k : a.kids)
        if (k->op != OP_CONST) return false;
    return true;
}

// bottom-up constant folding through the same VM the device runs
void fold(AstP &a) {
    for (auto &k : a->kids) fold(k);
    if (a->op == OP_CONST || a->op == OP_VAR) return;
    if (a->op == OP_POW && a->kids[1]->op == OP_CONST) {
        const double e = a->kids[1]->val;
        if (e == std::floor(e) && std::fabs(e) <= 16.0 && a->kids[0]->op != OP_CONST) {
            a->op = OP_POWI;
            a->arg = (int)e;
            a->kids.pop_back();
            return;
        }
    }
    if (all_const(*a)) {
        Program tmp{};
        int d = 0, md = 0;
        emit(*a, tmp, d, md);
        const double v = eval(&tmp, 0.0, 0.0, 0.0);
        a = leaf(v);
    }
}

void emit(const Ast &a, Program &p, int &depth, int &maxdepth) {
    for (const auto &k : a.kids) emit(*k, p, depth, maxdepth);
    if (p.len >= kMaxProgram) throw std::invalid_argument("expression too long for the device evaluator");
    Instr in{};
    in.op = a.op;
    in.arg = a.arg;
    in.val = a.val;
    p.code[p.len++] = in;
    const int pops = (int)a.kids.size();
    depth += 1 - pops;
    if (a.op == OP_CONST || a.op == OP_VAR) { /* depth already +1 */ }
    if (depth > maxdepth) maxdepth = depth;
    if (maxdepth > kMaxStack) throw std::invalid_argument("expression nests too deeply for the device evaluator");
}

std::string trimmed(std::string x) {
    const auto b = x.find_first_not_of(" \t\r\n");
    if (b == std::string::npos) return "";
    const auto e = x.find_last_not_of(" \t\r\n");
    return x.substr(b, e - b + 1);
}

}  // namespace

// src/ParameterReader.cpp:244-265: "pi" (any case), or <number> * pi where <number> matches the reference's
// pattern [0-9]*\.?[0-9]+ (digits, no sign, no exponent, no trailing dot); everything else goes through
// std::stod, which parses the longest numeric prefix ("-2*pi" -> -2, "1e3*pi" -> 1000, "2.*pi" -> 2) and
// throws std::invalid_argument when there is none -- the same values the host ParameterReader produces.
static bool plain_decimal(const std::string &t) {
    size_t i = 0, n = t.size();
    while (i < n && std::isdigit((unsigned char)t[i])) ++i;
    const size_t int_digits = i;
    if (i < n && t[i] == '.') {
        size_t j = i + 1;
        while (j < n && std::isdigit((unsigned char)t[j])) ++j;
        if (j > i + 1) return j == n;      // digits after the dot: [0-9]*\.[0-9]+
        return false;                      // "2." or "." : the pattern needs digits after the dot
    }
    return int_digits > 0 && i == n;       // [0-9]+
}
double parse_value_with_pi(std::string value) {
    const std::string t = trimmed(value);
    std::string low;
    for (char c : t) low.push_back((char)std::tolower((unsigned char)c));
    if (low == "pi") return M_PI;
    const auto star = low.find('*');
    if (star != std::string::npos && trimmed(low.substr(star + 1)) == "pi") {
        const std::string lhs = trimmed(low.substr(0, star));
        if (plain_decimal(lhs)) return std::stod(lhs) * M_PI;
    }
    return std::stod(value);  // throws std::invalid_argument like the reference (:264)
}

std::map<std::string, double> parse_constants(const std::string &s) {
    std::map<std::string, double> m;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, ',')) {
        const auto pos = item.find('=');
        if (pos == std::string::npos) continue;
        m[trimmed(item.substr(0, pos))] = parse_value_with_pi(item.substr(pos + 1));
    }
    return m;
}

Program compile_expression(const std::string &expression, const std::string &variables,
                           const std::string &constants) {
    if (trimmed(expression).empty()) throw std::invalid_argument("empty function expression");
    auto consts = parse_constants(constants);
    consts["pi"] = M_PI;  // src/ParameterReader.cpp:167
    std::vector<std::string> vars;
    {
        std::stringstream ss(variables);
        std::string item;
        while (std::getline(ss, item, ',')) {
            item = trimmed(item);
            if (!item.empty()) vars.push_back(item);
        }
    }
    const bool time_dependent = variables.find('t') != std::string::npos;  // :168 substring test
    if (vars.size() != (size_t)(2 + (time_dependent ? 1 : 0)))
        throw std::invalid_argument("variable list '" + variables + "' must name x, y" +
                                    (time_dependent ? ", t" : ""));
    Parser ps(expression, vars, consts);
    AstP root = ps.parse();
    fold(root);
    Program p{};
    p.time_dependent = time_dependent ? 1 : 0;
    int depth = 0, maxdepth = 0;
    emit(*root, p, depth, maxdepth);
    return p;
}

namespace {

// which variables a subtree reads: bit 0 = space (x or y), bit 1 = time
int dependence(const Ast &a) {
    int d = 0;
    if (a.op == OP_VAR) d = a.arg == 2 ? 2 : 1;
    for (const auto &k : a.kids) d |= dependence(*k);
    return d;
}
AstP clone(const Ast &a) {
    auto n = std::make_unique<Ast>();
    n->op = a.op;
    n->arg = a.arg;
    n->val = a.val;
    for (const auto &k : a.kids) n->kids.push_back(clone(*k));
    return n;
}
// flatten a chain of * and / (and unary minus) into (factor, is_divisor) pairs
void factors(const Ast &a, bool inverse, std::vector<std::pair<const Ast *, bool>> &out, double &sign) {
    if (a.op == OP_MUL) {
        factors(*a.kids[0], inverse, out, sign);
        factors(*a.kids[1], inverse, out, sign);
    } else if (a.op == OP_DIV) {
        factors(*a.kids[0], inverse, out, sign);
        factors(*a.kids[1], !inverse, out, sign);
    } else if (a.op == OP_NEG) {
        sign = -sign;
        factors(*a.kids[0], inverse, out, sign);
    } else
        out.emplace_back(&a, inverse);
}
AstP product(AstP acc, AstP f, bool inverse) {
    return node(inverse ? OP_DIV : OP_MUL, 0, std::move(acc), std::move(f));
}

}  // namespace

bool compile_separable(const std::string &expression, const std::string &variables, const std::string &constants,
                       Program *time_part, Program *space_part) {
    auto consts = parse_constants(constants);
    consts["pi"] = M_PI;
    std::vector<std::string> vars;
    {
        std::stringstream ss(variables);
        std::string item;
        while (std::getline(ss, item, ',')) {
            item = trimmed(item);
            if (!item.empty()) vars.push_back(item);
        }
    }
    if (vars.size() != 3) return false;  // no time variable: nothing to separate
    Parser ps(expression, vars, consts);
    AstP root = ps.parse();
    std::vector<std::pair<const Ast *, bool>> fs;
    double sign = 1.0;
    factors(*root, false, fs, sign);
    AstP t = leaf(sign), s = leaf(1.0);
    for (const auto &f : fs) {
        const int d = dependence(*f.first);
        if (d == 3) return false;  // a factor mixes space and time
        if (d == 2) t = product(std::move(t), clone(*f.first), f.second);
        else s = product(std::move(s), clone(*f.first), f.second);
    }
    fold(t);
    fold(s);
    auto finish = [](AstP &r, Program *p, int td) {
        *p = Program{};
        p->time_dependent = td;
        int depth = 0, maxdepth = 0;
        emit(*r, *p, depth, maxdepth);
    };
    finish(t, time_part, 1);
    finish(s, space_part, 1);
    return true;
}

bool is_constant(const Program &p, double *value) {
    if (p.len == 1 && p.code[0].op == OP_CONST) {
        if (value) *value = p.code[0].val;
        return true;
    }
    return false;
}

}  // namespace wv
