// wave_types.hpp -- the small local stand-ins for the deal.II types that appear in the reference's
// public interface (include/WaveEquationBase.hpp:72-95, include/ParameterReader.hpp:52-98):
// Point<dim>, Function<dim>, FunctionParser<dim>, ParameterHandler, ConditionalOStream.
// They carry the same member names the reference's mains and classes use, and nothing else.
#ifndef WAVE_TYPES_HPP
#define WAVE_TYPES_HPP

#include <array>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "wavegpu.h"

template <int dim>
class Point
{
  public:
    Point() { c.fill(0.0); }
    Point(double x, double y)
    {
        static_assert(dim == 2, "2-D only");
        c[0] = x;
        c[1] = y;
    }
    double operator[](unsigned int i) const { return c[i]; }
    double& operator[](unsigned int i) { return c[i]; }

  private:
    std::array<double, dim> c;
};

/// Scalar function of space and time with the set_time / value protocol of dealii::Function.
template <int dim>
class Function
{
  public:
    virtual ~Function() = default;
    virtual double value(const Point<dim>& p, unsigned int component = 0) const = 0;
    virtual void set_time(const double t) { time = t; }
    double get_time() const { return time; }

  protected:
    double time = 0.0;
};

/// Expression-backed function: the role dealii::FunctionParser (muParser) plays in the reference
/// (src/main-newmark.cpp:64-79).  The text is compiled by libwavegpu; the same text is what the
/// solver hands to wave_set_expr for evaluation on the device.
template <int dim>
class FunctionParser : public Function<dim>
{
  public:
    FunctionParser() = default;
    ~FunctionParser() override
    {
        if (handle)
            wave_expr_destroy(handle);
    }
    FunctionParser(const FunctionParser&) = delete;
    FunctionParser& operator=(const FunctionParser&) = delete;

    /// Same argument meaning as FunctionParser::initialize(vars, expression, constants, time_dependent).
    /// Throws std::invalid_argument when the expression does not parse.
    void initialize(const std::string& vars,
                    const std::string& expr,
                    const std::map<std::string, double>& consts,
                    const bool time_dep = false)
    {
        variables = vars;
        expression = expr;
        constants.clear();
        for (const auto& kv : consts)
        {
            if (kv.first == "pi")
                continue; // always defined by the library
            char buf[64];
            std::snprintf(buf, sizeof buf, "%.17g", kv.second);
            if (!constants.empty())
                constants += ", ";
            constants += kv.first + "=" + buf;
        }
        time_dependent = time_dep;
        if (handle)
        {
            wave_expr_destroy(handle);
            handle = nullptr;
        }
        char err[512] = {0};
        if (wave_expr_create(expression.c_str(), variables.c_str(), constants.c_str(), &handle, err, sizeof err) !=
            WAVE_OK)
            throw std::invalid_argument(err);
        initialized = true;
    }

    double value(const Point<dim>& p, unsigned int = 0) const override
    {
        if (!initialized)
            throw std::logic_error("FunctionParser used before initialize()");
        return wave_expr_value(handle, p[0], p[1], this->time);
    }

    bool is_initialized() const { return initialized; }
    const std::string& get_expression() const { return expression; }
    const std::string& get_variables() const { return variables; }
    const std::string& get_constants() const { return constants; }

  private:
    std::string expression, variables, constants;
    bool time_dependent = false;
    bool initialized = false;
    wave_expr* handle = nullptr;
};

/// Output stream that prints only when a condition holds (dealii::ConditionalOStream).
class ConditionalOStream
{
  public:
    ConditionalOStream(std::ostream& s, const bool active_) : out(s), active(active_) {}
    template <typename T>
    const ConditionalOStream& operator<<(const T& t) const
    {
        if (active)
            out << t;
        return *this;
    }
    const ConditionalOStream& operator<<(std::ostream& (*p)(std::ostream&)) const
    {
        if (active)
            out << p;
        return *this;
    }

  private:
    std::ostream& out;
    bool active;
};

/// Declared-entry parameter store with JSON and PRM input: the subset of dealii::ParameterHandler the
/// reference uses (declare_entry / enter_subsection / get* / parse_input).
class ParameterHandler
{
  public:
    enum class Kind
    {
        Anything,
        Integer,
        Double,
        Bool,
        IntegerList
    };
    void declare_entry(const std::string& name,
                       const std::string& default_value,
                       Kind kind = Kind::Anything,
                       double lower = -1e300,
                       double upper = 1e300,
                       const std::string& doc = "");
    void enter_subsection(const std::string& name);
    void leave_subsection();
    std::string get(const std::string& name) const;
    long get_integer(const std::string& name) const;
    double get_double(const std::string& name) const;
    bool get_bool(const std::string& name) const;
    /// Reads a .json parameter file (the format every shipped file and script of the reference uses) or a
    /// deal.II .prm file (set / subsection / end), chosen by the extension like dealii::ParameterHandler.
    void parse_input(const std::string& filename);

  private:
    struct Entry
    {
        std::string value;
        Kind kind;
        double lower, upper;
    };
    std::string key(const std::string& name) const;
    void set_checked(const std::string& full_key, const std::string& value);
    std::map<std::string, Entry> entries;
    std::vector<std::string> path;
};

#endif
