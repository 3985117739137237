// WaveNewmark.cpp -- the run() sequence of src/WaveNewmark.cpp:280-491 over the C ABI:
// setup + assembly (wave_setup), initial conditions and consistent a^0 (wave_init), then
// while (time < T) { time += dt; wave_step; divergence check; logging }.
#include "WaveNewmark.hpp"

void WaveNewmark::setup()
{
    pcout << "===============================================" << std::endl;
    create_context(WAVE_SCHEME_NEWMARK, 0.5, beta, gamma);
    setup_mesh();
    pcout << "-----------------------------------------------" << std::endl;
    setup_fe();
    pcout << "-----------------------------------------------" << std::endl;
    setup_dof_handler();
    pcout << "-----------------------------------------------" << std::endl;
    pcout << "Initializing the linear system" << std::endl;
    pcout << "  Initializing the sparsity pattern" << std::endl;
    pcout << "  Initializing matrices" << std::endl;
    pcout << "  Initializing vectors" << std::endl;
}

void WaveNewmark::assemble_matrices()
{
    pcout << "Assembling mass and stiffness matrices" << std::endl;
    check(wave_setup(ctx), "wave_setup");
    pcout << "  Setup complete!  (" << wave_local_nnz(ctx) << " matrix entries on the device)" << std::endl;
}

void WaveNewmark::run()
{
    setup();
    assemble_matrices();

    const std::string method_params = "-gamma" + clean_double(gamma) + "-beta" + clean_double(beta);
    prepare_output_filename(method_params);

    pcout << "Setting initial conditions..." << std::endl;
    pcout << "Computing consistent initial acceleration a^0..." << std::endl;
    check(wave_init(ctx), "wave_init");
    {
        double st[4];
        wave_cg_stats(ctx, st, 1);
        pcout << "  a0 solved (" << static_cast<long>(st[1]) << " CG iterations)" << std::endl;
        double nrm[2];
        check(wave_norms(ctx, nrm), "wave_norms");
        pcout << "||u0|| = " << nrm[0] << std::endl;
        pcout << "||v0|| = " << nrm[1] << std::endl;
    }
    pcout << "-----------------------------------------------" << std::endl;

    output();
    timestep_number = 0;
    time = 0.0;
    const double divergence_threshold = 1e130;
    unsigned long total_iterations = 0;

    const auto start_time = std::chrono::high_resolution_clock::now();

    while (time < T)
    {
        time += delta_t;
        ++timestep_number;

        int32_t its[2] = { 0, 0 };
        double nrm[2] = { 0.0, 0.0 };
        check(wave_step(ctx, time, its, nrm), "wave_step"); // assemble_rhs + solve_a + update_u_v
        current_iterations = static_cast<unsigned int>(its[0]);
        total_iterations += current_iterations;
        norm_u = nrm[0];
        norm_v = nrm[1];

        if (check_divergence(norm_u, norm_v, divergence_threshold))
        {
            pcout << "Divergence detected at step " << timestep_number << ", t = " << time
                  << "; stopping simulation." << std::endl;
            break;
        }

        if (log_every > 0 && (timestep_number % log_every == 0))
        {
            compute_and_log_energy();
            compute_and_log_error();
            log_point_probe();
            log_iterations(current_iterations, 0);
        }

        if (timestep_number % print_every == 0)
            print_step_info();

        output();
    }

    const auto end_time = std::chrono::high_resolution_clock::now();
    simulation_time = std::chrono::duration<double>(end_time - start_time).count();

    pcout << "\nSimulation completed: " << timestep_number << " steps, final time t = " << time << std::endl;
    pcout << "Elapsed time: " << std::fixed << std::setprecision(3) << simulation_time << " seconds" << std::endl;
    pcout << "Total CG iterations: " << total_iterations << ", avg per step: " << std::fixed << std::setprecision(1)
          << (timestep_number > 0 ? static_cast<double>(total_iterations) / timestep_number : 0.0) << std::endl;

    compute_final_errors("", std::to_string(beta), std::to_string(gamma));
    close_logs();
}
