// cli.cpp -- see cli.hpp.
#include "cli.hpp"

#include <cstdlib>
#include <filesystem>
#include <memory>

#include "ParameterReader.hpp"
#include "launch_env.hpp"
#include "WaveNewmark.hpp"
#include "WaveTheta.hpp"

namespace
{
constexpr unsigned int dim = WaveEquationBase::dim;
const char* const kDefaultParameters = "../parameters/sine-membrane.json";
const char* const kRule = "===============================================";

struct SchemeTraits
{
    const char* prefix;     // results folder prefix and problem-name prefix
    const char* class_name; // for messages
    std::vector<const char*> scalars; // scheme parameters echoed after T
    bool exports_param_file;          // only main-newmark sets NMPDE_PARAM_FILE
};

SchemeTraits traits_of(Scheme s)
{
    if (s == Scheme::Newmark)
        return { "newmark", "WaveNewmark", { "Beta", "Gamma" }, true };
    return { "theta", "WaveTheta", { "Theta" }, false };
}

std::string joined(const std::vector<const char*>& names, bool quoted)
{
    std::string out;
    for (size_t k = 0; k < names.size(); ++k)
    {
        if (k)
            out += ", ";
        out += quoted ? std::string("'") + names[k] + "'" : std::string(names[k]);
    }
    return out;
}

// the seven expression-backed functions of a run, in the order of the parameter file sections
struct ProblemFunctions
{
    FunctionParser<dim> c, f, u0, v0, g, dgdt, exact;
    std::vector<std::string> names() const { return { "C", "F", "U0", "V0", "G", "DGDT", "Solution" }; }
    std::vector<FunctionParser<dim>*> slots() { return { &c, &f, &u0, &v0, &g, &dgdt, &exact }; }
};

void echo_parameters(const ConditionalOStream& out, ParameterHandler& prm, const std::string& problem_name,
                     const SchemeTraits& tr)
{
    out << "Parsed parameters:" << std::endl;
    out << "  Problem name: " << problem_name << std::endl;
    out << "  Geometry: " << prm.get("Geometry") << std::endl;
    out << "  Nel: " << prm.get("Nel") << std::endl;
    out << "  R (degree): " << prm.get_integer("R") << std::endl;
    out << "  T: " << prm.get_double("T") << std::endl;
    for (const char* name : tr.scalars)
        out << "  " << name << ": " << prm.get_double(name) << std::endl;
    out << "  Dt: " << prm.get_double("Dt") << std::endl;
}

// Side channel main -> WaveEquationBase (kept from the reference): returns the effective log interval.
unsigned int export_runtime_flags(ParameterHandler& prm)
{
    ::setenv("NMPDE_SAVE_SOLUTION", prm.get_bool("Save Solution") ? "1" : "0", 1);
    long log_every = prm.get_integer("Log Every");
    if (!prm.get_bool("Enable Logging"))
        log_every = 0; // "Enable Logging": false is Log Every = 0
    ::setenv("NMPDE_LOG_EVERY", std::to_string(log_every).c_str(), 1);
    return static_cast<unsigned int>(log_every);
}

std::unique_ptr<WaveEquationBase> make_solver(Scheme scheme, const std::string& problem_name, ParameterReader& reader,
                                              ParameterHandler& prm, ProblemFunctions& fn, unsigned int log_every)
{
    const auto nel = reader.get_nel();
    const auto box = reader.get_geometry();
    const auto degree = static_cast<unsigned int>(prm.get_integer("R"));
    const auto print_every = static_cast<unsigned int>(prm.get_integer("Print Every"));
    Function<dim>* exact = fn.exact.is_initialized() ? &fn.exact : nullptr;
    if (scheme == Scheme::Newmark)
        return std::make_unique<WaveNewmark>(problem_name, nel, box, degree, prm.get_double("T"),
                                             prm.get_double("Gamma"), prm.get_double("Beta"), prm.get_double("Dt"),
                                             fn.c, fn.f, fn.u0, fn.v0, fn.g, fn.dgdt, log_every, print_every, exact);
    return std::make_unique<WaveTheta>(problem_name, nel, box, degree, prm.get_double("T"), prm.get_double("Theta"),
                                       prm.get_double("Dt"), fn.c, fn.f, fn.u0, fn.v0, fn.g, fn.dgdt, log_every,
                                       print_every, exact);
}
} // namespace

int wave_cli_main(int argc, char* argv[], Scheme scheme)
{
    const SchemeTraits tr = traits_of(scheme);
    // rank 0 speaks for the run, as with the reference's pcout (include/WaveEquationBase.hpp:119)
    LaunchEnvironment launch;
    try
    {
        launch = detect_launch_environment();
    }
    catch (const std::exception& e)
    {
        std::cout << "Error in the launcher environment: " << e.what() << std::endl;
        return 1;
    }
    const ConditionalOStream pcout(std::cout, launch.rank == 0);
    const bool from_argument = argc > 1;
    const std::string parameters_file = from_argument ? argv[1] : kDefaultParameters;

    pcout << "Backend: libwavegpu (CUDA sm_100a), " << launch.size << (launch.size == 1 ? " GPU" : " GPUs") << std::endl
          << kRule << std::endl;
    if (from_argument)
        pcout << "Using parameter file from argument: " << parameters_file << std::endl;
    else
    {
        pcout << "Usage:./main <path-to-arguments-file> \nRemember you are inside /build" << std::endl;
        pcout << "Using default parameter file: " << parameters_file << std::endl;
    }
    pcout << kRule << std::endl;

    if (tr.exports_param_file)
        ::setenv("NMPDE_PARAM_FILE", parameters_file.c_str(), 1);

    const std::string problem_name =
        std::string(tr.prefix) + "-" + std::filesystem::path(parameters_file).stem().string();

    ParameterHandler prm;
    ParameterReader reader(prm);
    ProblemFunctions fn;
    reader.declare(fn.names());

    // stage 1: parameter file and expressions
    try
    {
        reader.parse(parameters_file);
        reader.load_functions(fn.names(), fn.slots());
        echo_parameters(pcout, prm, problem_name, tr);
    }
    catch (const std::invalid_argument& e)
    {
        pcout << "Error while parsing parameters/functions: " << e.what() << std::endl;
        pcout << "Hint: check JSON fields (Geometry, Nel, R, T, " << joined(tr.scalars, false)
              << ", Dt) and function strings; ensure numeric fields are valid numbers and not empty." << std::endl;
        return 1;
    }
    catch (const std::exception& e)
    {
        pcout << "Unexpected error while parsing parameters: " << e.what() << std::endl;
        return 1;
    }

    // stage 2: the run
    try
    {
        const unsigned int log_every = export_runtime_flags(prm);
        make_solver(scheme, problem_name, reader, prm, fn, log_every)->run();
    }
    catch (const std::invalid_argument& e)
    {
        pcout << "Error while initializing or running " << tr.class_name << ": " << e.what() << std::endl;
        if (launch.rank != 0)
            std::cerr << "rank " << launch.rank << ": error while initializing or running " << tr.class_name << ": "
                      << e.what() << std::endl;
        pcout << "Likely cause: a non-numeric or malformed value in the parameter file (stod failure)." << std::endl;
        pcout << "Please verify fields like 'R', 'T', " << joined(tr.scalars, true)
              << ", 'Dt' and function definitions C/F/U0/V0/G/DGDT in " << parameters_file << std::endl;
        return 1;
    }
    catch (const std::exception& e)
    {
        pcout << "Unexpected error: " << e.what() << std::endl;
        if (launch.rank != 0) // pcout is rank 0's: a failure of any other rank must not be silent
            std::cerr << "rank " << launch.rank << ": unexpected error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
