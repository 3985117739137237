// WaveEquationBase.cpp -- logging, file naming and C-ABI plumbing shared by WaveNewmark / WaveTheta.
// Follows src/WaveEquationBase.cpp of the reference for every user-visible artefact (folder names,
// CSV headers, stream formatting, step line); the numerics behind each call are libwavegpu's.
#include "WaveEquationBase.hpp"

#include "vtu_writer.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <sstream>
#include <stdexcept>
#include <utility>

namespace
{
bool env_flag_enabled(const char* name, const bool default_value)
{
    const char* v = std::getenv(name);
    if (!v)
        return default_value;
    const std::string s(v);
    if (s == "0" || s == "false" || s == "FALSE" || s == "False")
        return false;
    if (s == "1" || s == "true" || s == "TRUE" || s == "True")
        return true;
    return default_value;
}

const FunctionParser<2>& as_parser(const Function<2>& fn, const char* name)
{
    const auto* p = dynamic_cast<const FunctionParser<2>*>(&fn);
    if (!p || !p->is_initialized())
        throw std::invalid_argument(std::string("function '") + name +
                                    "' must be an initialised FunctionParser to be evaluated on the device");
    return *p;
}
} // namespace

WaveEquationBase::WaveEquationBase(const std::string& problem_name_,
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
                                   Function<dim>* exact_solution_)
    : problem_name(problem_name_), N_el(N_el_), geometry(geometry_), r(r_), T(T_), delta_t(delta_t_), c(c_), f(f_),
      u0(u0_), v0(v0_), g(g_), dgdt(dgdt_), log_every(log_every_), print_every(print_every_),
      exact_solution(exact_solution_), launch(detect_launch_environment()), mpi_size(launch.size),
      mpi_rank(launch.rank), pcout(std::cout, launch.rank == 0)
{
}

WaveEquationBase::~WaveEquationBase()
{
    if (ctx)
        wave_destroy(ctx);
}

void WaveEquationBase::check(int status, const char* what) const
{
    if (status == WAVE_OK)
        return;
    const std::string msg = std::string(what) + ": " + wave_last_error(ctx);
    if (status == WAVE_ERR_EXPR || status == WAVE_ERR_ARG)
        throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

void WaveEquationBase::create_context(int scheme, double theta, double beta, double gamma)
{
    wave_config cfg;
    wave_default_config(&cfg);
    cfg.nx = static_cast<int32_t>(N_el.first);
    cfg.ny = static_cast<int32_t>(N_el.second);
    cfg.x0 = geometry.first[0];
    cfg.x1 = geometry.second[0];
    cfg.y0 = geometry.first[1];
    cfg.y1 = geometry.second[1];
    cfg.r = static_cast<int32_t>(r);
    cfg.scheme = scheme;
    cfg.dt = delta_t;
    cfg.theta = theta;
    cfg.beta = beta;
    cfg.gamma = gamma;
    // solver options without a field in the JSON schema come through the environment, like NMPDE_*:
    // WAVE_PRECOND=mg selects the multigrid V-cycle instead of Jacobi
    if (const char* pc = std::getenv("WAVE_PRECOND"))
    {
        const std::string v(pc);
        if (v == "mg" || v == "MG" || v == "2")
            cfg.precond = WAVE_PRECOND_MG;
        else if (v == "none" || v == "1")
            cfg.precond = WAVE_PRECOND_NONE;
    }
    // WAVE_CG_REDUCE / WAVE_CG_TOL override ReductionControl's reduction factor and absolute tolerance
    // (src/WaveNewmark.cpp:256: 1e-6, 1e-12) -- e.g. to reproduce the printed digits of the reference's
    // convergence tables, which its AMG-preconditioned solves reach far below the stopping bar
    for (const auto& opt : { std::make_pair("WAVE_CG_REDUCE", &cfg.cg_reduce), std::make_pair("WAVE_CG_TOL", &cfg.cg_tol) })
        if (const char* v = std::getenv(opt.first))
        {
            char* end = nullptr;
            const double x = std::strtod(v, &end);
            if (end == v || !(x > 0.0))
                throw std::invalid_argument(std::string(opt.first) + "='" + v + "' is not a positive number");
            *opt.second = x;
        }
    // several ranks: one GPU each, strips of quad rows, the communicator id broadcast through the
    // launcher's rendezvous file (the reference's MPI_COMM_WORLD needs no such step)
    unsigned char comm_id[128] = {};
    if (mpi_size > 1)
    {
        // One GPU per rank.  Every rank evaluates the same condition before the id is published, so either
        // all of them stop here or none does (a single rank that threw would leave the others waiting in the
        // communicator set-up).  Ranks per node: the launcher's count when it exports one, else all ranks.
        const int n_devices = wave_device_count();
        unsigned int per_node = mpi_size;
        for (const char* name : { "WAVE_LOCAL_NRANKS", "OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS", "SLURM_NTASKS_PER_NODE",
                                  "LOCAL_WORLD_SIZE" })
            if (const char* v = std::getenv(name))
            {
                const long x = std::strtol(v, nullptr, 10);
                if (x > 0)
                {
                    per_node = static_cast<unsigned int>(x);
                    break;
                }
            }
        if (n_devices > 0 && per_node > static_cast<unsigned int>(n_devices))
            throw std::runtime_error("rank " + std::to_string(mpi_rank) + " of " + std::to_string(mpi_size) + ": " +
                                     std::to_string(per_node) + " ranks on this node but " + std::to_string(n_devices) +
                                     " device(s) visible; launch at most one rank per GPU");
        if (mpi_rank == 0 && wave_comm_unique_id(comm_id) != WAVE_OK)
            throw std::runtime_error(std::string("wave_comm_unique_id: ") + wave_last_error(nullptr));
        share_communicator_id(launch, comm_id);
        cfg.rank = static_cast<int32_t>(mpi_rank);
        cfg.nranks = static_cast<int32_t>(mpi_size);
        cfg.device = n_devices > 0 ? static_cast<int32_t>(launch.local_rank % n_devices) : -1;
        cfg.nccl_unique_id = comm_id;
    }
    const int created = wave_create(&cfg, &ctx);
    if (mpi_size > 1 && mpi_rank == 0)
        std::remove(launch.rendezvous.c_str()); // every rank has joined (or the run is over)
    if (created != WAVE_OK)
        throw std::runtime_error(std::string("wave_create: ") + wave_last_error(nullptr));

    struct Slot { int id; const Function<dim>* fn; const char* name; };
    const Slot slots[] = { { WAVE_EXPR_C, &c, "C" }, { WAVE_EXPR_F, &f, "F" }, { WAVE_EXPR_U0, &u0, "U0" },
                           { WAVE_EXPR_V0, &v0, "V0" }, { WAVE_EXPR_G, &g, "G" }, { WAVE_EXPR_DGDT, &dgdt, "DGDT" } };
    for (const auto& s : slots)
    {
        const auto& p = as_parser(*s.fn, s.name);
        check(wave_set_expr(ctx, s.id, p.get_expression().c_str(), p.get_variables().c_str(),
                            p.get_constants().c_str()),
              s.name);
    }
    if (exact_solution != nullptr)
    {
        const auto& p = as_parser(*exact_solution, "Solution");
        check(wave_set_expr(ctx, WAVE_EXPR_SOLUTION, p.get_expression().c_str(), p.get_variables().c_str(),
                            p.get_constants().c_str()),
              "Solution");
    }
}

void WaveEquationBase::setup_mesh()
{
    pcout << "Initializing the mesh" << std::endl;
    // the structured simplex mesh is generated analytically on the device; no mesh file is written
    pcout << "  Number of elements = " << 2ull * N_el.first * N_el.second << std::endl;
    if (mpi_size > 1)
        pcout << "  Partition          = " << mpi_size << " strips of quad rows, one GPU each (" << launch.source
              << ")" << std::endl;
}

void WaveEquationBase::setup_fe()
{
    pcout << "Initializing the finite element space" << std::endl;
    pcout << "  Degree                     = " << r << std::endl;
    pcout << "  DoFs per cell              = " << (r == 1 ? 3 : 6) << std::endl;
    double xi[16], eta[16], w[16];
    pcout << "  Quadrature points per cell = " << wave_quadrature(static_cast<int32_t>(r) + 1, xi, eta, w) << std::endl;
}

void WaveEquationBase::setup_dof_handler()
{
    pcout << "Initializing the DoF handler" << std::endl;
    pcout << "  Number of DoFs = " << wave_n_dofs(ctx) << std::endl;
}

void WaveEquationBase::prepare_output_filename(const std::string& method_params)
{
    output_folder = "../results/" + problem_name + "/run-R" + std::to_string(r) + "-N" +
                    std::to_string(N_el.first) + "x" + std::to_string(N_el.second) + "-dt" + clean_double(delta_t) +
                    "-T" + clean_double(T) + method_params + "/";

    pcout << "Output folder: " << output_folder << std::endl;

    if (mpi_rank == 0)
    {
        if (!std::filesystem::exists(output_folder))
            std::filesystem::create_directories(output_folder);

        if (const char* param_env = std::getenv("NMPDE_PARAM_FILE"))
        {
            try
            {
                const std::filesystem::path src(param_env);
                const std::filesystem::path dst = std::filesystem::path(output_folder) / "parameters.json";
                if (std::filesystem::exists(src))
                {
                    std::filesystem::copy_file(src, dst, std::filesystem::copy_options::overwrite_existing);
                    pcout << "  Parameters copied to " << dst << std::endl;
                }
                else
                    pcout << "  Parameter file not found: " << src << std::endl;
            }
            catch (const std::exception& e)
            {
                pcout << "  Warning: could not copy parameter file (" << e.what() << ")" << std::endl;
            }
        }

        // energy/error CSV files are opened lazily so that log_every = 0 produces no files
        if (exact_solution != nullptr)
        {
            const std::string convergence_file_path = "../results/" + problem_name + "/convergence.csv";
            const bool file_exists = std::filesystem::exists(convergence_file_path);
            convergence_file.open(convergence_file_path, std::ios_base::app);
            if (convergence_file.is_open() && !file_exists)
                convergence_file << "h,N_el_x,N_el_y,r,dt,T,method,theta,beta,gamma,rel_L2_error_final,"
                                    "rel_H1_error_final,elapsed_time_s"
                                 << std::endl;
        }
    }
}

// energy.csv: E = 1/2 (v^T M v + u^T K u) from the device, default stream formatting (6 significant
// digits) for time and energy (src/WaveEquationBase.cpp:148-168)
void WaveEquationBase::compute_and_log_energy()
{
    check(wave_energy(ctx, &current_energy), "wave_energy");
    if (mpi_rank != 0)
        return;
    auto& out = energy_log.rows(output_folder + "energy.csv", "timestep,time,energy");
    if (out)
        out << timestep_number << ',' << time << ',' << current_energy << std::endl;
}

// probe.csv: u_h at the centre of the box, scientific with 10 digits (src/WaveEquationBase.cpp:170-222)
void WaveEquationBase::log_point_probe()
{
    double u_probe = 0.0;
    check(wave_probe(ctx, 0.5 * (geometry.first[0] + geometry.second[0]),
                     0.5 * (geometry.first[1] + geometry.second[1]), &u_probe),
          "wave_probe");
    if (mpi_rank != 0)
        return;
    auto& out = probe_log.rows(output_folder + "probe.csv", "timestep,time,u_probe");
    if (out)
        out << timestep_number << ',' << time << ',' << std::scientific << std::setprecision(10) << u_probe
            << std::endl;
}

// iterations.csv: CG iterations of the one (Newmark) or two (theta) solves of the step (:224-239)
void WaveEquationBase::log_iterations(const unsigned int n_iterations_1, const unsigned int n_iterations_2)
{
    if (mpi_rank != 0)
        return;
    auto& out = iterations_log.rows(output_folder + "iterations.csv", "timestep,time,iterations_1,iterations_2");
    if (out)
        out << timestep_number << ',' << time << ',' << n_iterations_1 << ',' << n_iterations_2 << std::endl;
}

// error.csv: L2 / H1 errors against the Solution expression and their relative versions, scientific
// with 6 digits (:241-272); nothing is written for problems without an exact solution
void WaveEquationBase::compute_and_log_error()
{
    if (exact_solution == nullptr)
        return;
    double e[4];
    check(wave_errors(ctx, time, e), "wave_errors");
    if (mpi_rank != 0)
        return;
    auto& out = error_log.rows(output_folder + "error.csv",
                               "timestep,time,L2_error,H1_error,rel_L2_error,rel_H1_error");
    if (out)
        out << timestep_number << ',' << time << ',' << std::scientific << std::setprecision(6) << e[0] << ',' << e[1]
            << ',' << e[2] << ',' << e[3] << std::endl;
}

void WaveEquationBase::compute_final_errors() { compute_final_errors("", "", ""); }

void WaveEquationBase::compute_final_errors(const std::string& theta_str, const std::string& beta_str,
                                            const std::string& gamma_str)
{
    if (exact_solution == nullptr)
        return;
    double e[4];
    check(wave_errors(ctx, time, e), "wave_errors");
    const double rel_error_L2 = e[2], rel_error_H1 = e[3];
    if (mpi_rank == 0 && convergence_file.is_open())
    {
        const double h = 1.0 / std::sqrt(N_el.first * N_el.second);
        convergence_file << h << "," << N_el.first << "," << N_el.second << "," << r << "," << delta_t << "," << T
                         << "," << problem_name << ",";
        convergence_file << (theta_str.empty() ? "N/A" : theta_str) << "," << (beta_str.empty() ? "N/A" : beta_str)
                         << "," << (gamma_str.empty() ? "N/A" : gamma_str) << ",";
        convergence_file << std::scientific << std::setprecision(6) << rel_error_L2 << "," << rel_error_H1 << ",";
        convergence_file << std::fixed << std::setprecision(3) << simulation_time << std::endl;

        pcout << "Final (last-iteration) errors:" << std::endl;
        pcout << "  Relative L2 error  = " << std::scientific << std::setprecision(6) << rel_error_L2 << std::endl;
        pcout << "  Relative H1 error  = " << std::scientific << std::setprecision(6) << rel_error_H1 << std::endl;
    }
}

void WaveEquationBase::print_step_info()
{
    std::ostringstream oss;
    oss << "Step " << std::setw(6) << timestep_number << ",  t=" << std::scientific << std::setprecision(3)
        << std::setw(9) << time << ",  ||u||=" << std::scientific << std::setprecision(3) << std::setw(9) << norm_u
        << ",  ||v||=" << std::scientific << std::setprecision(3) << std::setw(9) << norm_v;
    const char* env_p = std::getenv("NMPDE_LOG_EVERY");
    if (!env_p || std::atoi(env_p) != 0)
        oss << ",  E=" << std::scientific << std::setprecision(3) << std::setw(9) << current_energy;
    pcout << oss.str() << std::endl;
}

// Corner points of every cell in the reference's cell order (two triangles per quad, quads row by
// row: T0 = {q0, q1, q2}, T1 = {q3, q2, q1}; src/WaveEquationBase.cpp:42-46) with the DoF that lives
// on each corner (the first three entries of the cell's DoF list, for P1 and P2 alike).
void WaveEquationBase::build_output_mesh() const
{
    OutputMesh& m = output_mesh;
    const int32_t nx = static_cast<int32_t>(N_el.first), ny = static_cast<int32_t>(N_el.second);
    const size_t n_cells = 2ull * nx * ny;
    const double dx = (geometry.second[0] - geometry.first[0]) / nx, dy = (geometry.second[1] - geometry.first[1]) / ny;
    std::vector<int32_t> owner_of_quad_row(static_cast<size_t>(ny), 0);
    for (unsigned int p = 0; p < mpi_size; ++p)
    {
        wave_partition plan;
        if (wave_partition_plan(nx, ny, static_cast<int32_t>(r), static_cast<int32_t>(p), static_cast<int32_t>(mpi_size),
                                &plan) != WAVE_OK)
            throw std::runtime_error("wave_partition_plan failed");
        for (int64_t j = plan.quad_row_begin; j < plan.quad_row_end; ++j)
            owner_of_quad_row[static_cast<size_t>(j)] = static_cast<int32_t>(p);
    }
    m.dof.resize(3 * n_cells);
    m.vertex.resize(3 * n_cells);
    m.xyz.resize(9 * n_cells);
    m.partitioning.resize(3 * n_cells);
    for (size_t cell = 0; cell < n_cells; ++cell)
    {
        int32_t dofs[6];
        if (wave_cell_dofs(nx, ny, static_cast<int32_t>(r), static_cast<int64_t>(cell), dofs) != WAVE_OK)
            throw std::runtime_error("wave_cell_dofs failed");
        const int32_t quad = static_cast<int32_t>(cell / 2), i = quad % nx, j = quad / nx;
        const bool upper = cell % 2 == 1;
        const int32_t ci[3] = { upper ? i + 1 : i, upper ? i : i + 1, upper ? i + 1 : i };
        const int32_t cj[3] = { upper ? j + 1 : j, upper ? j + 1 : j, upper ? j : j + 1 };
        for (int k = 0; k < 3; ++k)
        {
            const size_t p = 3 * cell + k;
            m.dof[p] = dofs[k];
            m.vertex[p] = cj[k] * (nx + 1) + ci[k];
            m.xyz[3 * p + 0] = static_cast<float>(geometry.first[0] + ci[k] * dx);
            m.xyz[3 * p + 1] = static_cast<float>(geometry.first[1] + cj[k] * dy);
            m.xyz[3 * p + 2] = 0.0f;
            m.partitioning[p] = owner_of_quad_row[static_cast<size_t>(j)];
        }
    }
    m.built = true;
}

// solution_NNNN.0.vtu + solution_NNNN.pvtu with the fields u, v, [u_exact], partitioning
// (src/WaveEquationBase.cpp:330-365); written at step 0 and after every step unless
// "Save Solution" is false.  The vectors come back from the device in canonical numbering
// (collective over the ranks), rank 0 writes the single piece.
void WaveEquationBase::output() const
{
    if (!env_flag_enabled("NMPDE_SAVE_SOLUTION", true))
        return;
    const size_t n = static_cast<size_t>(wave_n_dofs(ctx));
    host_u.resize(n);
    host_v.resize(n);
    check(wave_get_vector(ctx, WAVE_VEC_U, host_u.data(), n), "wave_get_vector(u)");
    check(wave_get_vector(ctx, WAVE_VEC_V, host_v.data(), n), "wave_get_vector(v)");
    if (mpi_rank != 0)
        return;
    if (!output_mesh.built)
        build_output_mesh();
    const OutputMesh& m = output_mesh;
    const size_t n_points = m.dof.size();

    std::vector<VtuField> fields;
    fields.push_back({ "u", std::vector<double>(n_points) });
    fields.push_back({ "v", std::vector<double>(n_points) });
    for (size_t p = 0; p < n_points; ++p)
    {
        fields[0].values[p] = host_u[static_cast<size_t>(m.dof[p])];
        fields[1].values[p] = host_v[static_cast<size_t>(m.dof[p])];
    }
    if (exact_solution != nullptr)
    {
        // VectorTools::interpolate of the exact solution at the current time; the corner value is
        // evaluated once per grid vertex at its double-precision coordinates
        exact_solution->set_time(time);
        const int32_t nx = static_cast<int32_t>(N_el.first);
        const double dx = (geometry.second[0] - geometry.first[0]) / N_el.first,
                     dy = (geometry.second[1] - geometry.first[1]) / N_el.second;
        const size_t n_vertices = (static_cast<size_t>(N_el.first) + 1) * (N_el.second + 1);
        std::vector<double> at_vertex(n_vertices);
        std::vector<char> known(n_vertices, 0);
        fields.push_back({ "u_exact", std::vector<double>(n_points) });
        for (size_t p = 0; p < n_points; ++p)
        {
            const size_t vtx = static_cast<size_t>(m.vertex[p]);
            if (!known[vtx])
            {
                const int32_t i = m.vertex[p] % (nx + 1), j = m.vertex[p] / (nx + 1);
                at_vertex[vtx] =
                    exact_solution->value(Point<dim>(geometry.first[0] + i * dx, geometry.first[1] + j * dy));
                known[vtx] = 1;
            }
            fields.back().values[p] = at_vertex[vtx];
        }
    }
    fields.push_back({ "partitioning", m.partitioning });

    const std::string piece = vtu_piece_name("solution", timestep_number, 0);
    write_vtu_piece(output_folder + piece, m.xyz, fields);
    std::vector<std::string> names;
    for (const VtuField& f : fields)
        names.push_back(f.name);
    write_pvtu_record(output_folder + pvtu_record_name("solution", timestep_number), { piece }, names);
}

bool WaveEquationBase::check_divergence(const double nu, const double nv, const double threshold) const
{
    return (!std::isfinite(nu) || !std::isfinite(nv) || nu > threshold || nv > threshold);
}

void WaveEquationBase::close_logs()
{
    if (mpi_rank != 0)
        return;
    for (LazyCsv* log : { &energy_log, &error_log, &iterations_log, &probe_log })
        log->close();
    if (convergence_file.is_open())
        convergence_file.close();
}

std::string clean_double(double x, int precision)
{
    std::ostringstream out;
    out << std::fixed << std::setprecision(precision) << x;
    std::string s = out.str();
    if (s.find('.') != std::string::npos)
    {
        while (!s.empty() && s.back() == '0')
            s.pop_back();
        if (!s.empty() && s.back() == '.')
            s.pop_back();
    }
    std::replace(s.begin(), s.end(), '.', '_');
    return s.empty() ? "0" : s;
}
