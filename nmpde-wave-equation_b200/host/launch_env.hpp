// launch_env.hpp -- what a process learns from its launcher: rank, number of ranks, rank on this
// node, and how the ranks share the 128-byte communicator id before wave_create.
//
// The reference reads these from MPI (Utilities::MPI::this_mpi_process / n_mpi_processes,
// include/WaveEquationBase.hpp:111-114) after `mpirun -np P main-...`.  libwavegpu runs one process
// per GPU and needs no MPI library: the rank variables a launcher exports are enough
// (bin/wave-mpirun, Open MPI / MPICH / Slurm launchers, torchrun), and the one broadcast a run
// needs -- the NCCL unique id -- goes through a rendezvous file on the node.
#ifndef WAVE_LAUNCH_ENV_HPP
#define WAVE_LAUNCH_ENV_HPP

#include <string>

struct LaunchEnvironment
{
    unsigned int rank = 0;
    unsigned int size = 1;
    unsigned int local_rank = 0;
    std::string rendezvous; // file rank 0 publishes the communicator id in
    std::string source;     // which family of variables was found ("WAVE_*", "OMPI_*", ...)
};

/// Read WAVE_RANK/WAVE_NRANKS/WAVE_LOCAL_RANK, else OMPI_COMM_WORLD_*, PMI_*; SLURM_PROCID/SLURM_NTASKS
/// only with WAVE_LAUNCHER=slurm and RANK/WORLD_SIZE/LOCAL_RANK only with WAVE_LAUNCHER=torchrun (both
/// are also set where nothing was launched in parallel); no variables = a single rank.  Throws std::invalid_argument on
/// inconsistent values (rank >= size, non-numeric text).
LaunchEnvironment detect_launch_environment();

/// Rank 0 publishes `id` (128 bytes), every other rank waits for it and fills `id`.
/// Throws std::runtime_error when the file does not appear within `timeout_s` seconds.
void share_communicator_id(const LaunchEnvironment& env, unsigned char id[128], double timeout_s = 120.0);

#endif
