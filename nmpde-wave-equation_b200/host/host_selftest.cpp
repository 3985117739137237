// host_selftest.cpp -- CPU-only checks of the host mirror (no device needed): clean_double, the
// ParameterHandler JSON reader and pattern checks, ParameterReader::get_nel / get_geometry / constants,
// FunctionParser over the library's expression handles.  Run by tests/test_host_cpu.py; prints one
// "ok <name>" line per check and exits non-zero on the first failure.
#include <cmath>
#include <cstdio>
#include <fstream>
#include <functional>
#include <iostream>

#include "ParameterReader.hpp"
#include "launch_env.hpp"
#include "vtu_writer.hpp"
#include "WaveEquationBase.hpp"

namespace
{
int failures = 0;
void check(bool cond, const char* name)
{
    std::cout << (cond ? "ok " : "FAIL ") << name << std::endl;
    if (!cond)
        ++failures;
}
template <class E>
bool throws(const std::function<void()>& f)
{
    try
    {
        f();
    }
    catch (const E&)
    {
        return true;
    }
    catch (...)
    {
        return false;
    }
    return false;
}
} // namespace

// `host_selftest --rendezvous`: one rank of a launched group (wave-mpirun or rank variables set by
// hand).  Rank 0 publishes a recognisable 128-byte record, every rank prints what it holds.
int rendezvous_mode()
{
    try
    {
        const LaunchEnvironment env = detect_launch_environment();
        unsigned char id[128];
        for (int k = 0; k < 128; ++k)
            id[k] = env.rank == 0 ? static_cast<unsigned char>(7 * k + 3) : 0;
        share_communicator_id(env, id, 20.0);
        unsigned long sum = 0;
        for (int k = 0; k < 128; ++k)
            sum += id[k] * static_cast<unsigned long>(k + 1);
        std::printf("rank %u of %u local %u via %s id %lu\n", env.rank, env.size, env.local_rank,
                    env.source.empty() ? "none" : env.source.c_str(), sum);
        return 0;
    }
    catch (const std::exception& e)
    {
        std::printf("error: %s\n", e.what());
        return 1;
    }
}

// `host_selftest --vtu <dir>`: two cells (one quad) with recognisable field values through the VTU
// writer; tests/test_host_cpu.py parses the files back.
int vtu_mode(const std::string& dir)
{
    try
    {
        const std::vector<float> xyz = { 0, 0, 0, 1, 0, 0, 0, 2, 0, /* T1 */ 1, 2, 0, 0, 2, 0, 1, 0, 0 };
        std::vector<VtuField> fields = { { "u", { 0.5, 1.5, 2.5, 3.5, 2.5, 1.5 } },
                                         { "partitioning", { 0, 0, 0, 1, 1, 1 } } };
        const std::string piece = vtu_piece_name("solution", 7, 0);
        write_vtu_piece(dir + "/" + piece, xyz, fields);
        write_pvtu_record(dir + "/" + pvtu_record_name("solution", 7), { piece }, { "u", "partitioning" });
        std::printf("%s %s\n", piece.c_str(), pvtu_record_name("solution", 12345).c_str());
        fields[0].values.pop_back();
        try
        {
            write_vtu_piece(dir + "/bad.vtu", xyz, fields);
            return 1;
        }
        catch (const std::invalid_argument&)
        {
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        std::printf("error: %s\n", e.what());
        return 1;
    }
}

int main(int argc, char** argv)
{
    if (argc > 1 && std::string(argv[1]) == "--rendezvous")
        return rendezvous_mode();
    if (argc > 2 && std::string(argv[1]) == "--vtu")
        return vtu_mode(argv[2]);
    const std::string dir = argc > 1 ? argv[1] : ".";

    // clean_double: src/WaveEquationBase.cpp:433-452 and scripts/dissipation_dispersion_sweep.py:333-357
    check(clean_double(0.05) == "0_05", "clean_double 0.05");
    check(clean_double(1.0) == "1", "clean_double 1.0");
    check(clean_double(10.0) == "10", "clean_double 10 keeps the integer zero");
    check(clean_double(0.015625) == "0_015625", "clean_double 1/64");
    check(clean_double(8e-5) == "0_00008", "clean_double 8e-5");
    check(clean_double(1e-7) == "0", "clean_double below 6 decimals");
    check(clean_double(0.25) == "0_25" && clean_double(0.5) == "0_5", "clean_double beta gamma");
    check(clean_double(-1.0, 2) == "-1", "clean_double negative, precision 2");

    // constants with pi: src/ParameterReader.cpp:237-294
    check(std::fabs(parse_value_with_pi(" pi ") - M_PI) < 1e-15, "constant pi");
    check(std::fabs(parse_value_with_pi("4.0*pi") - 4 * M_PI) < 1e-15, "constant 4.0*pi");
    check(std::fabs(parse_value_with_pi(" 2 * PI") - 2 * M_PI) < 1e-15, "constant 2 * PI");
    check(parse_value_with_pi("1e-3") == 1e-3, "plain number");
    check(throws<std::invalid_argument>([] { parse_value_with_pi("abc"); }), "bad constant throws invalid_argument");
    {
        auto m = parse_constants_with_pi_and_multiplication("TT=0.5, XX=0.5, ya=0.333, k=4.0*pi");
        check(m.size() == 4 && m["TT"] == 0.5 && std::fabs(m["k"] - 4 * M_PI) < 1e-15, "constant list");
    }

    // JSON parameter file with string, numeric and boolean literals (scripts/convergence_sweep.py:175-177)
    const std::string file = dir + "/selftest.json";
    {
        std::ofstream f(file);
        f << "{\n \"Geometry\": \"[-1.0, 1.0] x [0.0, 3.5]\", \"Nel\": \"180, 60\", \"R\": 2, \"T\": \"1.5\",\n"
             " \"Dt\": 0.005, \"Save Solution\": false, \"Enable Logging\": \"true\", \"Log Every\": 0,\n"
             " \"U0\": {\"Function constants\": \"A=2.0\", \"Function expression\": \"A*sin(pi*x)*y\", \"Variable "
             "names\": \"x, y\"},\n"
             " \"G\": {\"Function constants\": \"\", \"Function expression\": \"if(t<=0.5 && x<0.1, sin(t), 0.0)\", "
             "\"Variable names\": \"x, y, t\"},\n"
             " \"C\": {\"Function expression\": \"1.0\", \"Variable names\": \"x, y, t\"},\n"
             " \"F\": {\"Function expression\": \"0.0\", \"Variable names\": \"x, y, t\"},\n"
             " \"V0\": {\"Function expression\": \"0.0\", \"Variable names\": \"x, y\"},\n"
             " \"DGDT\": {\"Function expression\": \"0.0\", \"Variable names\": \"x, y, t\"}\n}\n";
    }
    ParameterHandler prm;
    ParameterReader reader(prm);
    const std::vector<std::string> names{ "C", "F", "U0", "V0", "G", "DGDT", "Solution" };
    reader.declare(names);
    reader.parse(file);
    check(prm.get_integer("R") == 2 && prm.get_double("Dt") == 0.005 && prm.get_double("T") == 1.5, "scalars");
    check(prm.get_double("Theta") == 0.5 && prm.get_double("Beta") == 0.25 && prm.get_integer("Print Every") == 10,
          "defaults of undeclared-in-file entries");
    check(!prm.get_bool("Save Solution") && prm.get_bool("Enable Logging") && prm.get_integer("Log Every") == 0,
          "bool / int literals");
    const auto nel = reader.get_nel();
    check(nel.first == 180 && nel.second == 60, "get_nel two values");
    const auto geo = reader.get_geometry();
    check(geo.first[0] == -1.0 && geo.second[0] == 1.0 && geo.first[1] == 0.0 && geo.second[1] == 3.5, "get_geometry");
    FunctionParser<2> c, f, u0, v0, g, dgdt, sol;
    reader.load_functions(names, { &c, &f, &u0, &v0, &g, &dgdt, &sol });
    check(!sol.is_initialized() && u0.is_initialized(), "Solution block optional");
    check(std::fabs(u0.value(Point<2>(0.5, 3.0)) - 6.0) < 1e-14, "U0 value with constant");
    g.set_time(0.25);
    check(std::fabs(g.value(Point<2>(0.05, 0.0)) - std::sin(0.25)) < 1e-15, "G inside the window");
    g.set_time(0.75);
    check(g.value(Point<2>(0.05, 0.0)) == 0.0, "G after the window");

    // the same entries from a deal.II .prm file (set / subsection / end, comments, continued lines):
    // dealii::ParameterHandler::parse_input picks the format by the extension (src/ParameterReader.cpp:134-137)
    {
        std::ofstream f(dir + "/selftest.prm");
        f << "# wave equation parameters\n"
             "set Geometry = [-1.0, 1.0] x [0.0, 3.5]\n"
             "set Nel      = 180, 60   # two values\n"
             "set R = 2\nset T = 1.5\nset Dt = 0.005\nset Save Solution = false\nset Log Every = 0\n"
             "subsection U0\n  set Function constants  = A=2.0\n  set Function expression = A*sin(pi*x) \\\n"
             "                            *y\n  set Variable names = x, y\nend\n"
             "subsection C\n  set Function expression = 1.0\n  set Variable names = x, y, t\nend\n"
             "subsection F\n  set Function expression = 0.0\n  set Variable names = x, y, t\nend\n"
             "subsection V0\n  set Function expression = 0.0\n  set Variable names = x, y\nend\n"
             "subsection G\n  set Function expression = 0.0\n  set Variable names = x, y, t\nend\n"
             "subsection DGDT\n  set Function expression = 0.0\n  set Variable names = x, y, t\nend\n";
    }
    {
        ParameterHandler p3;
        ParameterReader r3(p3);
        r3.declare(names);
        r3.parse(dir + "/selftest.prm");
        check(p3.get_integer("R") == 2 && p3.get_double("Dt") == 0.005 && !p3.get_bool("Save Solution") &&
                  p3.get_integer("Log Every") == 0 && p3.get_double("Beta") == 0.25,
              "prm scalars and defaults");
        const auto nel3 = r3.get_nel();
        check(nel3.first == 180 && nel3.second == 60, "prm Nel with a trailing comment");
        FunctionParser<2> c3, f3, u3, v3, g3, d3, s3;
        r3.load_functions(names, { &c3, &f3, &u3, &v3, &g3, &d3, &s3 });
        check(std::fabs(u3.value(Point<2>(0.5, 3.0)) - 6.0) < 1e-14, "prm subsection with a continued line");
    }
    check(throws<std::runtime_error>([&] {
              ParameterHandler p2;
              ParameterReader r2(p2);
              r2.declare(names);
              std::ofstream(dir + "/open.prm") << "subsection U0\n set Function expression = 1.0\n";
              r2.parse(dir + "/open.prm");
          }),
          "unclosed subsection in a prm file");
    check(throws<std::runtime_error>([&] {
              ParameterHandler p2;
              ParameterReader r2(p2);
              r2.declare(names);
              std::ofstream(dir + "/x.yaml") << "R: 2\n";
              r2.parse(dir + "/x.yaml");
          }),
          "unknown file name extension");

    // error behaviour
    check(throws<std::invalid_argument>([&] {
              ParameterHandler p2;
              ParameterReader r2(p2);
              r2.declare(names);
              std::ofstream(dir + "/bad.json") << "{\"R\": \"0\"}";
              r2.parse(dir + "/bad.json");
          }),
          "R = 0 violates Patterns::Integer(1)");
    check(throws<std::runtime_error>([&] {
              ParameterHandler p2;
              ParameterReader r2(p2);
              r2.declare(names);
              std::ofstream(dir + "/bad2.json") << "{\"Unknown Key\": \"1\"}";
              r2.parse(dir + "/bad2.json");
          }),
          "undeclared entry is an error");
    check(throws<std::invalid_argument>([&] {
              ParameterHandler p2;
              ParameterReader r2(p2);
              r2.declare(names);
              std::ofstream(dir + "/bad3.json") << "{\"Geometry\": \"0,1,0,1\"}";
              r2.parse(dir + "/bad3.json");
              r2.get_geometry();
          }),
          "malformed Geometry");
    check(throws<std::invalid_argument>([&] {
              FunctionParser<2> bad;
              bad.initialize("x, y", "sin(pi*x", {}, false);
          }),
          "unparsable expression throws invalid_argument");
    check(throws<std::invalid_argument>([&] {
              ParameterHandler p2;
              ParameterReader r2(p2);
              r2.declare(names);
              std::ofstream(dir + "/bad4.json") << "{\"C\": {\"Function expression\": \"\"}}";
              r2.parse(dir + "/bad4.json");
              FunctionParser<2> a1, a2, a3, a4, a5, a6, a7;
              r2.load_functions(names, { &a1, &a2, &a3, &a4, &a5, &a6, &a7 });
          }),
          "missing expression for C");
    std::cout << (failures ? "FAILED" : "ALL OK") << std::endl;
    return failures ? 1 : 0;
}
