// launch_env.cpp -- see launch_env.hpp.
#include "launch_env.hpp"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <stdexcept>
#include <thread>

namespace
{
constexpr size_t kIdBytes = 128;
constexpr char kMagic[8] = { 'W', 'A', 'V', 'E', 'I', 'D', '0', '2' };
// the record rank 0 publishes: magic, number of ranks, communicator id.  A reader accepts it only when the
// rank count matches and the file is fresh (a leftover of a crashed earlier run is ignored).
struct Record
{
    char magic[8];
    unsigned int nranks;
    unsigned int reserved;
    unsigned char id[kIdBytes];
};

bool read_unsigned(const char* name, unsigned int& out)
{
    const char* v = std::getenv(name);
    if (!v || !*v)
        return false;
    char* end = nullptr;
    const long x = std::strtol(v, &end, 10);
    if (end == v || *end != '\0' || x < 0)
        throw std::invalid_argument(std::string(name) + "='" + v + "' is not a rank number");
    out = static_cast<unsigned int>(x);
    return true;
}

// start time of a process in clock ticks since boot (field 22 of /proc/<pid>/stat): together with
// the pid it names one launcher instance even after the pid has been recycled
std::string process_start_ticks(long pid)
{
    std::ifstream f("/proc/" + std::to_string(pid) + "/stat");
    std::string line;
    if (!std::getline(f, line))
        return "0";
    const size_t close = line.rfind(')'); // the command name may contain spaces and parentheses
    if (close == std::string::npos)
        return "0";
    size_t pos = close + 1;
    std::string field;
    for (int k = 3; k <= 22; ++k) // fields after the name start with number 3
    {
        while (pos < line.size() && line[pos] == ' ')
            ++pos;
        const size_t next = line.find(' ', pos);
        field = line.substr(pos, next == std::string::npos ? std::string::npos : next - pos);
        if (next == std::string::npos)
            break;
        pos = next;
    }
    return field.empty() ? "0" : field;
}

std::string default_rendezvous_path()
{
    const char* tmp = std::getenv("TMPDIR");
    const long parent = static_cast<long>(::getppid());
    return std::string(tmp && *tmp ? tmp : "/tmp") + "/wavegpu-id-" + std::to_string(parent) + "-" +
           process_start_ticks(parent);
}
} // namespace

LaunchEnvironment detect_launch_environment()
{
    // Variables only a process launcher sets are trusted as they are.  SLURM_* also exist in a plain
    // batch shell (a directly started binary is then one rank, as it is for the reference without srun)
    // and RANK / WORLD_SIZE are common pod-level settings, so those two families must be asked for:
    // WAVE_LAUNCHER=slurm or WAVE_LAUNCHER=torchrun.
    struct Family { const char* label; const char* rank; const char* size; const char* local; const char* opt_in; };
    static const Family families[] = {
        { "WAVE_*", "WAVE_RANK", "WAVE_NRANKS", "WAVE_LOCAL_RANK", nullptr },
        { "OMPI_COMM_WORLD_*", "OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_RANK", nullptr },
        { "PMI_*", "PMI_RANK", "PMI_SIZE", "MPI_LOCALRANKID", nullptr },
        { "SLURM_*", "SLURM_PROCID", "SLURM_NTASKS", "SLURM_LOCALID", "slurm" },
        { "RANK/WORLD_SIZE", "RANK", "WORLD_SIZE", "LOCAL_RANK", "torchrun" },
    };
    const char* chosen = std::getenv("WAVE_LAUNCHER");
    LaunchEnvironment env;
    for (const Family& f : families)
    {
        unsigned int rank = 0, size = 1, local = 0;
        if (f.opt_in && !(chosen && std::string(chosen) == f.opt_in))
            continue;
        if (!read_unsigned(f.size, size) || !read_unsigned(f.rank, rank))
            continue;
        if (size == 0 || rank >= size)
            throw std::invalid_argument(std::string(f.rank) + "=" + std::to_string(rank) + " is outside " + f.size +
                                        "=" + std::to_string(size));
        env.rank = rank;
        env.size = size;
        env.local_rank = read_unsigned(f.local, local) ? local : rank;
        env.source = f.label;
        break;
    }
    if (const char* path = std::getenv("WAVE_RENDEZVOUS"))
        env.rendezvous = path;
    if (env.rendezvous.empty())
        env.rendezvous = default_rendezvous_path();
    return env;
}

void share_communicator_id(const LaunchEnvironment& env, unsigned char id[128], const double timeout_s)
{
    if (env.size <= 1)
        return;
    if (env.rank == 0)
    {
        // A record left behind by an earlier run must never be read as this run's: remove it first.  The
        // new record is written beside the final name (exclusive create, no symlink following, owner-only)
        // and renamed, so readers see nothing or the complete record.
        ::unlink(env.rendezvous.c_str());
        const std::string staging = env.rendezvous + ".part";
        ::unlink(staging.c_str());
        const int fd = ::open(staging.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW | O_CLOEXEC, 0600);
        if (fd < 0)
            throw std::runtime_error("cannot create the rendezvous file " + staging);
        Record rec{};
        std::memcpy(rec.magic, kMagic, sizeof kMagic);
        rec.nranks = env.size;
        std::memcpy(rec.id, id, kIdBytes);
        const bool ok = ::write(fd, &rec, sizeof rec) == static_cast<ssize_t>(sizeof rec);
        if (::close(fd) != 0 || !ok || std::rename(staging.c_str(), env.rendezvous.c_str()) != 0)
        {
            std::remove(staging.c_str());
            throw std::runtime_error("cannot publish the communicator id in " + env.rendezvous);
        }
        return;
    }
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout_s);
    for (;;)
    {
        const int fd = ::open(env.rendezvous.c_str(), O_RDONLY | O_NOFOLLOW | O_CLOEXEC);
        if (fd >= 0)
        {
            Record rec{};
            struct stat st{};
            const bool read_ok = ::read(fd, &rec, sizeof rec) == static_cast<ssize_t>(sizeof rec) && ::fstat(fd, &st) == 0;
            ::close(fd);
            // fresh = written within this wait (plus slack for ranks that started late)
            const bool fresh = read_ok && std::difftime(std::time(nullptr), st.st_mtime) <= timeout_s + 60.0;
            if (read_ok && fresh && std::memcmp(rec.magic, kMagic, sizeof kMagic) == 0 && rec.nranks == env.size)
            {
                std::memcpy(id, rec.id, kIdBytes);
                return;
            }
        }
        if (std::chrono::steady_clock::now() > deadline)
            throw std::runtime_error("rank " + std::to_string(env.rank) + ": no communicator id from rank 0 in " +
                                     env.rendezvous + " after " + std::to_string(static_cast<int>(timeout_s)) +
                                     " s (were all ranks started by the same launcher?)");
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
}
