// WaveTheta.cpp -- the run() sequence of src/WaveTheta.cpp:341-447 over the C ABI.
#include "WaveTheta.hpp"

void WaveTheta::setup()
{
    pcout << "===============================================" << std::endl;
    create_context(WAVE_SCHEME_THETA, theta, 0.25, 0.5);
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

void WaveTheta::assemble_matrices()
{
    pcout << "Assembling mass and stiffness matrices" << std::endl;
    check(wave_setup(ctx), "wave_setup");
    pcout << "Setup complete!  (" << wave_local_nnz(ctx) << " matrix entries on the device)" << std::endl;
}

void WaveTheta::run()
{
    setup();
    assemble_matrices();

    const std::string method_params = "-theta" + clean_double(theta);
    prepare_output_filename(method_params);

    pcout << "Setting initial conditions..." << std::endl;
    check(wave_init(ctx), "wave_init");
    {
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
    unsigned long total_iterations_u = 0, total_iterations_v = 0;

    const auto start_time = std::chrono::high_resolution_clock::now();

    while (time < T)
    {
        time += delta_t;
        ++timestep_number;

        int32_t its[2] = { 0, 0 };
        double nrm[2] = { 0.0, 0.0 };
        // assemble_rhs_u + solve_u + assemble_rhs_v + solve_v
        check(wave_step(ctx, time, its, nrm), "wave_step");
        current_iterations_u = static_cast<unsigned int>(its[0]);
        current_iterations_v = static_cast<unsigned int>(its[1]);
        total_iterations_u += current_iterations_u;
        total_iterations_v += current_iterations_v;
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
            log_iterations(current_iterations_u, current_iterations_v);
        }

        if (timestep_number % print_every == 0)
            print_step_info();

        output();
    }

    const auto end_time = std::chrono::high_resolution_clock::now();
    simulation_time = std::chrono::duration<double>(end_time - start_time).count();

    pcout << "\nSimulation completed: " << timestep_number << " steps, final time t = " << time << std::endl;
    pcout << "Elapsed time: " << std::fixed << std::setprecision(3) << simulation_time << " seconds" << std::endl;
    pcout << "Total CG iterations (u): " << total_iterations_u << ", avg per step: " << std::fixed
          << std::setprecision(1)
          << (timestep_number > 0 ? static_cast<double>(total_iterations_u) / timestep_number : 0.0) << std::endl;
    pcout << "Total CG iterations (v): " << total_iterations_v << ", avg per step: " << std::fixed
          << std::setprecision(1)
          << (timestep_number > 0 ? static_cast<double>(total_iterations_v) / timestep_number : 0.0) << std::endl;

    compute_final_errors(std::to_string(theta), "", "");
    close_logs();
}
