// WaveEquationBase.hpp -- shared base of the two time integrators, with the reference's surface
// (include/WaveEquationBase.hpp:55-370): constructor arguments, run(), the logging helpers and the
// CSV formats.  Mesh, DoFs, matrices and vectors live on the GPU behind a wave_ctx; this class
// only sequences C-ABI calls and writes the files.
#ifndef WAVE_EQUATION_BASE_HPP
#define WAVE_EQUATION_BASE_HPP

#include <chrono>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

#include "launch_env.hpp"
#include "wave_types.hpp"

/// A CSV time series that is created (with its header line) on the first row: runs with
/// log_every = 0 leave no files behind, as in the reference.
class LazyCsv
{
  public:
    /// Stream positioned after the header; opens `path` on first use.  Formatting flags set on the
    /// stream persist between rows (the reference's sticky std::scientific).
    std::ofstream& rows(const std::string& path, const char* header)
    {
        if (!file.is_open())
        {
            file.open(path);
            if (file.is_open())
                file << header << std::endl;
        }
        return file;
    }
    bool is_open() const { return file.is_open(); }
    void close()
    {
        if (file.is_open())
            file.close();
    }

  private:
    std::ofstream file;
};

class WaveEquationBase
{
  public:
    static constexpr unsigned int dim = 2;

    virtual ~WaveEquationBase();
    /// Run the full simulation (setup, assembly, time loop, output).
    virtual void run() = 0;

  protected:
    WaveEquationBase(const std::string& problem_name_,
                     const std::pair<unsigned int, unsigned int>& N_el_,
                     const std::pair<Point<dim>, Point<dim>>& geometry_,
                     const unsigned int& r_,
                     const double& T_,
                     const double& delta_t_,
                     const Function<dim>& c_,
                     Function<dim>& f_,
                     const Function<dim>& u0_,
                     const Function<dim>& v0_,
                     Function<dim>& g_,
                     Function<dim>& dgdt_,
                     const unsigned int log_every_,
                     const unsigned int print_every_,
                     Function<dim>* exact_solution_);

    // ---- Mesh & FE initialisation (device side: wave_create / wave_set_expr / wave_setup) -------
    void create_context(int scheme, double theta, double beta, double gamma);
    void setup_mesh();
    void setup_fe();
    void setup_dof_handler();

    // ---- Output & logging ---------------------------------------------------------------------
    void prepare_output_filename(const std::string& method_params);
    void compute_and_log_energy();
    void log_point_probe();
    void log_iterations(const unsigned int n_iterations_1, const unsigned int n_iterations_2);
    void compute_and_log_error();
    void compute_final_errors();
    void compute_final_errors(const std::string& theta_str, const std::string& beta_str,
                              const std::string& gamma_str);
    void print_step_info();
    void output() const;
    bool check_divergence(const double norm_u, const double norm_v, const double threshold) const;
    void close_logs();
    void check(int status, const char* what) const;

    // ---- Problem description --------------------------------------------------------------------
    const std::string problem_name;
    std::string output_folder;
    LazyCsv energy_log, error_log, iterations_log, probe_log; // per-run series in output_folder
    std::ofstream convergence_file;                           // one row per run, shared across runs

    std::pair<unsigned int, unsigned int> N_el;
    const std::pair<Point<dim>, Point<dim>> geometry;
    const unsigned int r;

    const double T;
    const double delta_t;
    double time = 0.0;
    unsigned int timestep_number = 0;

    double current_energy = 0.0;
    double simulation_time = 0.0;
    double norm_u = 0.0, norm_v = 0.0; // ||u||_2, ||v||_2 of the last step

    const Function<dim>& c;
    Function<dim>& f;
    const Function<dim>& u0;
    const Function<dim>& v0;
    Function<dim>& g;
    Function<dim>& dgdt;

    const unsigned int log_every;
    const unsigned int print_every;
    Function<dim>* exact_solution;

    // One process drives one GPU.  Rank and size come from the launcher's environment (launch_env.hpp)
    // where the reference asks MPI (include/WaveEquationBase.hpp:111-114); rank p owns the p-th strip
    // of quad rows and rank 0 writes every file, as in the reference.
    const LaunchEnvironment launch;
    const unsigned int mpi_size;
    const unsigned int mpi_rank;

    wave_ctx* ctx = nullptr;
    ConditionalOStream pcout;

  private:
    // what output() needs of the mesh, built on its first call: per output point (three per cell)
    // the DoF and the grid vertex it sits on, the coordinates and the owning rank of its cell
    struct OutputMesh
    {
        bool built = false;
        std::vector<int32_t> dof, vertex;
        std::vector<float> xyz;
        std::vector<double> partitioning;
    };
    mutable OutputMesh output_mesh;
    mutable std::vector<double> host_u, host_v;
    void build_output_mesh() const;
};

/// Format a double for folder names: fixed notation, trailing zeros stripped, '.' -> '_'
/// (src/WaveEquationBase.cpp:433-452).
std::string clean_double(double x, int precision = 6);

#endif
