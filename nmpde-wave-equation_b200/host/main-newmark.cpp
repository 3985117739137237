// main-newmark.cpp -- command-line entry point with the reference's behaviour (src/main-newmark.cpp): one optional
// positional argument (the JSON parameter file, default ../parameters/sine-membrane.json), the
// NMPDE_* environment side channel, parse errors -> message + exit code 1.  One process drives one
// B200 through libwavegpu.
#include <cstdlib>
#include <filesystem>

#include "ParameterReader.hpp"
#include "WaveNewmark.hpp"

int main(int argc, char* argv[])
{
    ConditionalOStream pcout(std::cout, true);
    pcout << "Backend: libwavegpu (CUDA sm_100a), 1 GPU" << std::endl;
    pcout << "===============================================" << std::endl;
    std::string parameters_file =
        (argc > 1) ? std::string(argv[1]) : std::string("../parameters/sine-membrane.json");
    if (argc <= 1)
    {
        pcout << "Usage:./main <path-to-arguments-file> \nRemember you are inside /build" << std::endl;
        pcout << "Using default parameter file: " << parameters_file << std::endl;
    }
    else
        pcout << "Using parameter file from argument: " << parameters_file << std::endl;
    pcout << "===============================================" << std::endl;

    // Make the parameter file path available to downstream code (copied into the run folder).
    ::setenv("NMPDE_PARAM_FILE", parameters_file.c_str(), 1);

    constexpr unsigned int dim = WaveNewmark::dim;

    std::string problem_name = "newmark-" + std::filesystem::path(parameters_file).stem().string();
    ParameterHandler prm;
    ParameterReader param(prm);

    FunctionParser<dim> c, f, u0, v0, g, dgdt, exact_solution;

    std::vector<std::string> function_names{ "C", "F", "U0", "V0", "G", "DGDT", "Solution" };
    param.declare(function_names);

    try
    {
        param.parse(parameters_file);
        param.load_functions(function_names, { &c, &f, &u0, &v0, &g, &dgdt, &exact_solution });

        pcout << "Parsed parameters:" << std::endl;
        pcout << "  Problem name: " << problem_name << std::endl;
        pcout << "  Geometry: " << prm.get("Geometry") << std::endl;
        pcout << "  Nel: " << prm.get("Nel") << std::endl;
        pcout << "  R (degree): " << prm.get_integer("R") << std::endl;
        pcout << "  T: " << prm.get_double("T") << std::endl;
        pcout << "  Beta: " << prm.get_double("Beta") << std::endl;
        pcout << "  Gamma: " << prm.get_double("Gamma") << std::endl;
        pcout << "  Dt: " << prm.get_double("Dt") << std::endl;
    }
    catch (const std::invalid_argument& e)
    {
        pcout << "Error while parsing parameters/functions: " << e.what() << std::endl;
        pcout << "Hint: check JSON fields (Geometry, Nel, R, T, Beta, Gamma, Dt) and function strings; ensure "
                 "numeric fields are valid numbers and not empty."
              << std::endl;
        return 1;
    }
    catch (const std::exception& e)
    {
        pcout << "Unexpected error while parsing parameters: " << e.what() << std::endl;
        return 1;
    }

    // Export runtime flags (used by WaveEquationBase without changing class APIs).
    const bool save_solution = prm.get_bool("Save Solution");
    const bool enable_logging = prm.get_bool("Enable Logging");
    ::setenv("NMPDE_SAVE_SOLUTION", save_solution ? "1" : "0", 1);

    int log_every = static_cast<int>(prm.get_integer("Log Every"));
    if (!enable_logging)
        log_every = 0;
    ::setenv("NMPDE_LOG_EVERY", std::to_string(log_every).c_str(), 1);

    Function<dim>* exact_solution_ptr = exact_solution.is_initialized() ? &exact_solution : nullptr;

    try
    {
        WaveNewmark problem(problem_name, param.get_nel(), param.get_geometry(),
                            static_cast<unsigned int>(prm.get_integer("R")), prm.get_double("T"),
                            /* gamma */ prm.get_double("Gamma"), /* beta */ prm.get_double("Beta"),
                            /* delta_t */ prm.get_double("Dt"), c, f, u0, v0, g, dgdt,
                            static_cast<unsigned int>(log_every),
                            static_cast<unsigned int>(prm.get_integer("Print Every")), exact_solution_ptr);
        problem.run();
    }
    catch (const std::invalid_argument& e)
    {
        pcout << "Error while initializing or running WaveNewmark: " << e.what() << std::endl;
        pcout << "Likely cause: a non-numeric or malformed value in the parameter file (stod failure)." << std::endl;
        pcout << "Please verify fields like 'R', 'T', 'Beta', 'Gamma', 'Dt' and function definitions C/F/U0/V0/G/DGDT in "
              << parameters_file << std::endl;
        return 1;
    }
    catch (const std::exception& e)
    {
        pcout << "Unexpected error: " << e.what() << std::endl;
        return 1;
    }

    return 0;
}
