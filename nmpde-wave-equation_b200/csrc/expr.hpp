// expr.hpp -- host-side compiler: parameter-file expression text -> wv::Program bytecode.
#pragma once
#include <map>
#include <string>

#include "expr_vm.h"

namespace wv {

// Mirrors ParameterReader::load_functions + FunctionParser::initialize
// (src/ParameterReader.cpp:139-175): `constants` is the "k=v, ..." string (values: number, pi,
// n*pi; :237-294), `variables` the "x, y[, t]" list; time-dependent iff it contains 't' (:168).
// Throws std::invalid_argument on any error (main maps it to exit code 1,
// src/main-newmark.cpp:92-97).
Program compile_expression(const std::string &expression, const std::string &variables,
                           const std::string &constants);

std::map<std::string, double> parse_constants(const std::string &s);
double parse_value_with_pi(std::string value);

// Space-time separation of a forcing term: if the expression is a product / quotient chain whose
// factors each depend on t only or on (x, y) only, returns true and compiles f = T(t) * S(x, y)
// into `time_part` and `space_part`.  The load vector of S is then assembled once and scaled by T(t)
// every step instead of re-running the quadrature loop (src/WaveNewmark.cpp:151-171) per step.
bool compile_separable(const std::string &expression, const std::string &variables, const std::string &constants,
                       Program *time_part, Program *space_part);

// true if the program is the literal constant `value` after folding
bool is_constant(const Program &p, double *value);

}  // namespace wv
