// ParameterReader.hpp -- same interface as the reference's include/ParameterReader.hpp:38-110:
// declares the schema on a ParameterHandler, parses the JSON file, loads the FunctionParser
// objects and exposes Nel / Geometry in typed form.
#ifndef PARAMETER_READER_HPP
#define PARAMETER_READER_HPP

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "wave_types.hpp"

class ParameterReader
{
  public:
    static constexpr unsigned int dim = 2;

    explicit ParameterReader(ParameterHandler& paramhandler);

    /// Declare scalar entries and one subsection per function name (src/ParameterReader.cpp:39-126).
    void declare(const std::vector<std::string>& function_names);
    /// Parse the parameter file (src/ParameterReader.cpp:134-137).
    void parse(const std::string& filename);
    /// Initialise every FunctionParser from its subsection (src/ParameterReader.cpp:139-175).
    /// Throws std::invalid_argument for a missing expression (except "Solution") or a parse error.
    void load_functions(const std::vector<std::string>& names,
                        const std::vector<FunctionParser<dim>*>& funcs);
    /// "[x_min, x_max] x [y_min, y_max]" -> corner points (src/ParameterReader.cpp:177-196).
    std::pair<Point<dim>, Point<dim>> get_geometry() const;
    /// "N" or "Nx, Ny" (src/ParameterReader.cpp:198-230).
    std::pair<unsigned int, unsigned int> get_nel() const;

  private:
    void declare_scalar_parameters();
    void declare_function_subsections(const std::vector<std::string>& names);
    ParameterHandler& prm;
};

/// "pi", "<number>*pi" or a plain number (src/ParameterReader.cpp:237-265).
double parse_value_with_pi(std::string value);
/// "k=v, k2=v2" -> map (src/ParameterReader.cpp:267-294).
std::map<std::string, double> parse_constants_with_pi_and_multiplication(const std::string& s);

#endif
