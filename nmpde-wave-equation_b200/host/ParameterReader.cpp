// ParameterReader.cpp -- schema, JSON input and function loading of the parameter files.
// Mirrors src/ParameterReader.cpp of the reference; deal.II's ParameterHandler is replaced by the
// local one in wave_types.hpp, whose JSON reader lives at the end of this file.
#include "ParameterReader.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <fstream>
#include <functional>
#include <regex>
#include <sstream>

namespace
{
std::string trim_copy(std::string x)
{
    const auto b = x.find_first_not_of(" \t\r\n");
    if (b == std::string::npos)
        return "";
    const auto e = x.find_last_not_of(" \t\r\n");
    return x.substr(b, e - b + 1);
}
} // namespace

ParameterReader::ParameterReader(ParameterHandler& paramhandler)
    : prm(paramhandler)
{
}

void ParameterReader::declare_scalar_parameters()
{
    using K = ParameterHandler::Kind;
    // names, defaults and ranges of src/ParameterReader.cpp:41-104
    prm.declare_entry("Nel", "40", K::IntegerList, 1);
    prm.declare_entry("Geometry", "[0.0, 1.0] x [0.0, 1.0]", K::Anything);
    prm.declare_entry("Mesh File Name", "../mesh/mesh-square-40.msh", K::Anything); // declared, unused
    prm.declare_entry("R", "1", K::Integer, 1);
    prm.declare_entry("T", "1.0", K::Double, 0.0);
    prm.declare_entry("Theta", "0.5", K::Double, 0.0, 1.0);
    prm.declare_entry("Beta", "0.25", K::Double, 0.0, 1.0);
    prm.declare_entry("Gamma", "0.5", K::Double, 0.0, 1.0);
    prm.declare_entry("Dt", "0.01", K::Double, 0.0);
    prm.declare_entry("Save Solution", "true", K::Bool);
    prm.declare_entry("Enable Logging", "true", K::Bool);
    prm.declare_entry("Log Every", "10", K::Integer, 0);
    prm.declare_entry("Print Every", "10", K::Integer, 1);
}

void ParameterReader::declare_function_subsections(const std::vector<std::string>& names)
{
    for (const auto& n : names)
    {
        prm.enter_subsection(n);
        prm.declare_entry("Function constants", "");
        prm.declare_entry("Function expression", "");
        prm.declare_entry("Variable names", "");
        prm.leave_subsection();
    }
}

void ParameterReader::declare(const std::vector<std::string>& function_names)
{
    declare_scalar_parameters();
    declare_function_subsections(function_names);
}

void ParameterReader::parse(const std::string& filename)
{
    prm.parse_input(filename);
}

void ParameterReader::load_functions(const std::vector<std::string>& names,
                                     const std::vector<FunctionParser<dim>*>& funcs)
{
    if (names.size() != funcs.size())
    {
        std::cerr << "Mismatch names/functions size\n";
        return;
    }
    for (unsigned int i = 0; i < names.size(); ++i)
    {
        prm.enter_subsection(names[i]);
        const std::string expr = prm.get("Function expression");
        const std::string var_names = prm.get("Variable names");
        const std::string constants_str = prm.get("Function constants");
        prm.leave_subsection();

        if (names[i] == "Solution" && expr.empty())
            continue; // optional
        if (expr.empty())
            throw std::invalid_argument("Function expression for '" + names[i] +
                                        "' must be specified in the parameter file.");

        auto constants = parse_constants_with_pi_and_multiplication(constants_str);
        constants["pi"] = M_PI;
        const bool time_dependent = (var_names.find("t") != std::string::npos); // substring test, as upstream
        funcs[i]->initialize(var_names, expr, constants, time_dependent);
    }
}

std::pair<Point<ParameterReader::dim>, Point<ParameterReader::dim>> ParameterReader::get_geometry() const
{
    const auto geom_str = prm.get("Geometry");
    std::regex pattern(R"(\[\s*([-\d\.]+)\s*,\s*([-\d\.]+)\s*\]\s*x\s*\[\s*([-\d\.]+)\s*,\s*([-\d\.]+)\s*\])");
    std::smatch match;
    if (!std::regex_match(geom_str, match, pattern))
        throw std::invalid_argument("Invalid Geometry format in parameters.");
    const double x_min = std::stod(match[1].str()), x_max = std::stod(match[2].str());
    const double y_min = std::stod(match[3].str()), y_max = std::stod(match[4].str());
    return { Point<dim>(x_min, y_min), Point<dim>(x_max, y_max) };
}

std::pair<unsigned int, unsigned int> ParameterReader::get_nel() const
{
    std::vector<std::string> tokens;
    std::stringstream ss(trim_copy(prm.get("Nel")));
    std::string item;
    while (std::getline(ss, item, ','))
    {
        item = trim_copy(item);
        if (!item.empty())
            tokens.push_back(item);
    }
    if (tokens.size() == 1)
    {
        const auto nel = static_cast<unsigned int>(std::stoul(tokens[0]));
        return { nel, nel };
    }
    if (tokens.size() == 2)
        return { static_cast<unsigned int>(std::stoul(tokens[0])), static_cast<unsigned int>(std::stoul(tokens[1])) };
    throw std::invalid_argument("Invalid Nel format. Expected single value or two space-separated values.");
}

double parse_value_with_pi(std::string value)
{
    value = trim_copy(value);
    std::string lower = value;
    std::transform(lower.begin(), lower.end(), lower.begin(), ::tolower);
    if (lower == "pi")
        return M_PI;
    static const std::regex mul_pattern(R"(^\s*([0-9]*\.?[0-9]+)\s*\*\s*(pi)\s*$)", std::regex::icase);
    std::smatch match;
    if (std::regex_match(value, match, mul_pattern))
        return std::stod(match[1].str()) * M_PI;
    return std::stod(value);
}

std::map<std::string, double> parse_constants_with_pi_and_multiplication(const std::string& s)
{
    std::map<std::string, double> m;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, ','))
    {
        const auto pos = item.find('=');
        if (pos == std::string::npos)
            continue;
        m[trim_copy(item.substr(0, pos))] = parse_value_with_pi(item.substr(pos + 1));
    }
    return m;
}

// =====================================================================================================
// ParameterHandler: declared entries + a small recursive JSON reader (objects of strings, numbers,
// booleans; the sweep scripts write `false` / `0` literals, scripts/convergence_sweep.py:175-177)
// =====================================================================================================
std::string ParameterHandler::key(const std::string& name) const
{
    std::string k;
    for (const auto& p : path)
        k += p + "/";
    return k + name;
}

void ParameterHandler::declare_entry(const std::string& name, const std::string& default_value, Kind kind,
                                     double lower, double upper, const std::string&)
{
    entries[key(name)] = Entry{ default_value, kind, lower, upper };
}

void ParameterHandler::enter_subsection(const std::string& name) { path.push_back(name); }
void ParameterHandler::leave_subsection()
{
    if (!path.empty())
        path.pop_back();
}

std::string ParameterHandler::get(const std::string& name) const
{
    const auto it = entries.find(key(name));
    if (it == entries.end())
        throw std::runtime_error("ParameterHandler: entry '" + key(name) + "' was not declared");
    return it->second.value;
}

long ParameterHandler::get_integer(const std::string& name) const
{
    const std::string v = get(name);
    size_t used = 0;
    const long r = std::stol(v, &used);
    if (used != trim_copy(v).size())
        throw std::invalid_argument("Cannot convert '" + v + "' to an integer (entry " + name + ")");
    return r;
}

double ParameterHandler::get_double(const std::string& name) const
{
    const std::string v = trim_copy(get(name));
    size_t used = 0;
    const double r = std::stod(v, &used);
    if (used != v.size())
        throw std::invalid_argument("Cannot convert '" + v + "' to a double (entry " + name + ")");
    return r;
}

bool ParameterHandler::get_bool(const std::string& name) const
{
    std::string v = trim_copy(get(name));
    std::transform(v.begin(), v.end(), v.begin(), ::tolower);
    if (v == "true" || v == "yes" || v == "on" || v == "1")
        return true;
    if (v == "false" || v == "no" || v == "off" || v == "0")
        return false;
    throw std::invalid_argument("Cannot convert '" + v + "' to a boolean (entry " + name + ")");
}

void ParameterHandler::set_checked(const std::string& full_key, const std::string& value)
{
    const auto it = entries.find(full_key);
    if (it == entries.end())
        throw std::runtime_error("ParameterHandler: no entry with name '" + full_key + "' was declared");
    Entry& e = it->second;
    auto bad = [&](const char* what) {
        throw std::invalid_argument("The value '" + value + "' of entry '" + full_key + "' does not match its pattern (" +
                                    what + ")");
    };
    const std::string v = trim_copy(value);
    try
    {
        switch (e.kind)
        {
            case Kind::Integer:
            {
                size_t used = 0;
                const long x = std::stol(v, &used);
                if (used != v.size() || x < e.lower || x > e.upper)
                    bad("integer out of range");
                break;
            }
            case Kind::Double:
            {
                size_t used = 0;
                const double x = std::stod(v, &used);
                if (used != v.size() || x < e.lower || x > e.upper)
                    bad("floating point number out of range");
                break;
            }
            case Kind::Bool:
            {
                std::string l = v;
                std::transform(l.begin(), l.end(), l.begin(), ::tolower);
                if (l != "true" && l != "false" && l != "yes" && l != "no" && l != "on" && l != "off" && l != "1" &&
                    l != "0")
                    bad("boolean");
                break;
            }
            case Kind::IntegerList:
            {
                std::stringstream ss(v);
                std::string item;
                int count = 0;
                while (std::getline(ss, item, ','))
                {
                    item = trim_copy(item);
                    size_t used = 0;
                    const long x = std::stol(item, &used);
                    if (used != item.size() || x < e.lower)
                        bad("list of integers >= 1");
                    ++count;
                }
                if (count < 1)
                    bad("non-empty list of integers");
                break;
            }
            case Kind::Anything:
                break;
        }
    }
    catch (const std::out_of_range&)
    {
        bad("number out of range");
    }
    catch (const std::invalid_argument& ex)
    {
        if (std::string(ex.what()).find("does not match") != std::string::npos)
            throw;
        bad("not a number");
    }
    e.value = value;
}

namespace
{
class JsonReader
{
  public:
    JsonReader(const std::string& text_, ParameterHandler& prm_,
               const std::function<void(const std::string&, const std::string&)>& set_)
        : text(text_), set(set_)
    {
        (void)prm_;
    }
    void run()
    {
        ws();
        object("");
        ws();
        if (pos != text.size())
            fail("trailing characters");
    }

  private:
    const std::string& text;
    std::function<void(const std::string&, const std::string&)> set;
    size_t pos = 0;

    [[noreturn]] void fail(const std::string& msg) const
    {
        throw std::runtime_error("JSON parameter file: " + msg + " at offset " + std::to_string(pos));
    }
    void ws()
    {
        while (pos < text.size() && std::isspace(static_cast<unsigned char>(text[pos])))
            ++pos;
    }
    std::string string_literal()
    {
        if (text[pos] != '"')
            fail("expected string");
        ++pos;
        std::string out;
        while (pos < text.size() && text[pos] != '"')
        {
            char c = text[pos++];
            if (c == '\\' && pos < text.size())
            {
                const char e = text[pos++];
                switch (e)
                {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u':
                    {
                        if (pos + 4 > text.size())
                            fail("bad unicode escape");
                        const unsigned code = std::stoul(text.substr(pos, 4), nullptr, 16);
                        pos += 4;
                        if (code < 0x80)
                            out += static_cast<char>(code);
                        else
                            out += '?';
                        break;
                    }
                    default: out += e;
                }
            }
            else
                out += c;
        }
        if (pos >= text.size())
            fail("unterminated string");
        ++pos;
        return out;
    }
    void object(const std::string& prefix)
    {
        if (text[pos] != '{')
            fail("expected '{'");
        ++pos;
        ws();
        if (text[pos] == '}')
        {
            ++pos;
            return;
        }
        for (;;)
        {
            ws();
            const std::string name = string_literal();
            ws();
            if (text[pos] != ':')
                fail("expected ':'");
            ++pos;
            ws();
            const std::string full = prefix.empty() ? name : prefix + "/" + name;
            if (text[pos] == '{')
                object(full);
            else if (text[pos] == '"')
                set(full, string_literal());
            else
            {
                // number / true / false / null literal: keep its text
                const size_t b = pos;
                while (pos < text.size() && text[pos] != ',' && text[pos] != '}' &&
                       !std::isspace(static_cast<unsigned char>(text[pos])))
                    ++pos;
                if (pos == b)
                    fail("expected a value");
                set(full, text.substr(b, pos - b));
            }
            ws();
            if (text[pos] == ',')
            {
                ++pos;
                continue;
            }
            if (text[pos] == '}')
            {
                ++pos;
                return;
            }
            fail("expected ',' or '}'");
        }
    }
};
} // namespace

void ParameterHandler::parse_input(const std::string& filename)
{
    std::ifstream in(filename);
    if (!in)
        throw std::runtime_error("Cannot open parameter file " + filename);
    const auto dot = filename.find_last_of('.');
    const std::string ext = dot == std::string::npos ? "" : filename.substr(dot);
    if (ext == ".prm")
    {
        // [deal.II] ParameterHandler's native format: "set <name> = <value>", "subsection <name>" ... "end",
        // '#' comments, a trailing backslash continues the line
        std::vector<std::string> sections;
        std::string line, logical;
        int line_no = 0;
        auto strip = [](std::string t) {
            const auto b = t.find_first_not_of(" \t\r");
            if (b == std::string::npos)
                return std::string();
            const auto e = t.find_last_not_of(" \t\r");
            return t.substr(b, e - b + 1);
        };
        while (std::getline(in, line))
        {
            ++line_no;
            const auto hash = line.find('#');
            if (hash != std::string::npos)
                line.erase(hash);
            line = strip(line);
            if (!line.empty() && line.back() == '\\')
            {
                line.pop_back();
                logical += strip(line) + " ";
                continue;
            }
            logical += line;
            const std::string stmt = strip(logical);
            logical.clear();
            if (stmt.empty())
                continue;
            if (stmt.rfind("subsection", 0) == 0 && (stmt.size() == 10 || std::isspace(static_cast<unsigned char>(stmt[10]))))
                sections.push_back(strip(stmt.substr(10)));
            else if (stmt == "end" || stmt == "END")
            {
                if (sections.empty())
                    throw std::runtime_error(filename + ":" + std::to_string(line_no) + ": 'end' without a subsection");
                sections.pop_back();
            }
            else if (stmt.rfind("set", 0) == 0 && stmt.size() > 3 && std::isspace(static_cast<unsigned char>(stmt[3])))
            {
                const auto eq = stmt.find('=');
                if (eq == std::string::npos)
                    throw std::runtime_error(filename + ":" + std::to_string(line_no) + ": 'set' without '='");
                std::string full;
                for (const auto& sct : sections)
                    full += sct + "/";
                set_checked(full + strip(stmt.substr(3, eq - 3)), strip(stmt.substr(eq + 1)));
            }
            else
                throw std::runtime_error(filename + ":" + std::to_string(line_no) + ": cannot parse '" + stmt + "'");
        }
        if (!sections.empty())
            throw std::runtime_error(filename + ": subsection '" + sections.back() + "' is not closed");
        return;
    }
    if (ext != ".json")
        throw std::runtime_error("Unknown input file name extension '" + ext + "': .json and .prm parameter files are supported");
    std::stringstream buf;
    buf << in.rdbuf();
    const std::string text = buf.str();
    JsonReader reader(text, *this, [this](const std::string& k, const std::string& v) { set_checked(k, v); });
    reader.run();
}
