// wave-mpirun -- process launcher with mpirun's command line for the GPU build.
//
// The reference is started as `mpirun -np P [binding options] ./main-newmark file.json`
// (README.md:113-114, scripts/scalability_sweep.py:40-44, scripts/*.pbs).  libwavegpu runs one
// process per GPU, so this launcher starts min(P, visible GPUs) copies of the program, tells each
// its rank through WAVE_RANK / WAVE_NRANKS / WAVE_LOCAL_RANK, names the rendezvous file the ranks
// share the communicator id through (launch_env.hpp), waits for all of them and returns the first
// non-zero exit status.  Placement options of mpirun (hostfile, binding, mapping, MCA parameters)
// have no meaning on one node of GPUs and are accepted and ignored; `-x VAR=value` is applied.
#include <signal.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "wavegpu.h"

namespace
{
std::vector<pid_t> children;

void forward_signal(int sig)
{
    for (const pid_t pid : children)
        if (pid > 0)
            ::kill(pid, sig);
}

struct CommandLine
{
    long requested_ranks = 1;
    std::vector<std::string> exported; // -x VAR=value
    std::vector<char*> program;        // argv of the program, null-terminated
};

// number of option values to skip for the mpirun options that take any
int option_values(const std::string& opt)
{
    static const char* const one[] = { "--hostfile", "-hostfile", "--machinefile", "-machinefile", "--host", "-host",
                                       "-H", "--bind-to", "-bind-to", "--map-by", "-map-by", "--rank-by", "-rank-by",
                                       "--prefix", "--wdir", "-wdir", "--output-filename", "--timeout", "--npernode",
                                       "-npernode", "--cpus-per-proc", "--cpus-per-rank" };
    static const char* const two[] = { "--mca", "-mca", "--gmca", "-gmca" };
    for (const char* o : two)
        if (opt == o)
            return 2;
    for (const char* o : one)
        if (opt == o)
            return 1;
    return 0;
}

bool parse(int argc, char** argv, CommandLine& cl)
{
    int k = 1;
    for (; k < argc; ++k)
    {
        const std::string a = argv[k];
        if (a == "-np" || a == "-n" || a == "--np" || a == "--n" || a == "-c")
        {
            if (k + 1 >= argc)
                return false;
            char* end = nullptr;
            cl.requested_ranks = std::strtol(argv[++k], &end, 10);
            if (*end != '\0' || cl.requested_ranks < 1)
                return false;
        }
        else if (a == "-x")
        {
            if (k + 1 >= argc)
                return false;
            cl.exported.emplace_back(argv[++k]);
        }
        else if (const int skip = option_values(a))
            k += skip;
        else if (a.size() > 1 && a[0] == '-')
            continue; // flags without a value: --oversubscribe, --allow-run-as-root, --report-bindings, ...
        else
            break;
    }
    if (k >= argc)
        return false;
    for (; k < argc; ++k)
        cl.program.push_back(argv[k]);
    cl.program.push_back(nullptr);
    return true;
}

// at most one rank per visible GPU; WAVE_LAUNCH_MAX_RANKS overrides the device count (tests,
// or several ranks on hosts whose GPUs the launcher cannot see)
long usable_ranks(long requested)
{
    long cap = wave_device_count();
    if (const char* v = std::getenv("WAVE_LAUNCH_MAX_RANKS"))
        cap = std::atol(v);
    if (cap < 1)
        cap = 1;
    return requested < cap ? requested : cap;
}

int exit_status_of(int status)
{
    if (WIFEXITED(status))
        return WEXITSTATUS(status);
    if (WIFSIGNALED(status))
        return 128 + WTERMSIG(status);
    return 1;
}
} // namespace

int main(int argc, char** argv)
{
    CommandLine cl;
    if (!parse(argc, argv, cl))
    {
        std::fprintf(stderr, "usage: wave-mpirun [-np P] [mpirun placement options] <program> [arguments]\n");
        return 2;
    }
    for (const std::string& e : cl.exported)
    {
        const size_t eq = e.find('=');
        if (eq != std::string::npos) // plain `-x VAR` exports the caller's value, which children inherit anyway
            ::setenv(e.substr(0, eq).c_str(), e.substr(eq + 1).c_str(), 1);
    }

    const long ranks = usable_ranks(cl.requested_ranks);
    if (ranks < cl.requested_ranks)
        std::fprintf(stderr, "wave-mpirun: -np %ld requested, starting %ld (one rank per visible GPU)\n",
                     cl.requested_ranks, ranks);

    if (ranks == 1)
    {
        ::unsetenv("WAVE_RANK");
        ::unsetenv("WAVE_NRANKS");
        ::unsetenv("WAVE_LOCAL_RANK");
        ::execvp(cl.program[0], cl.program.data());
        std::fprintf(stderr, "wave-mpirun: cannot execute %s: %s\n", cl.program[0], std::strerror(errno));
        return 127;
    }

    const char* tmp = std::getenv("TMPDIR");
    std::string rendezvous = std::string(tmp && *tmp ? tmp : "/tmp") + "/wavegpu-id-XXXXXX";
    {
        const int fd = ::mkstemp(&rendezvous[0]); // reserves a unique name; rank 0 renames its record over it
        if (fd < 0)
        {
            std::fprintf(stderr, "wave-mpirun: cannot create %s: %s\n", rendezvous.c_str(), std::strerror(errno));
            return 1;
        }
        ::close(fd);
    }
    ::setenv("WAVE_RENDEZVOUS", rendezvous.c_str(), 1);
    ::setenv("WAVE_NRANKS", std::to_string(ranks).c_str(), 1);

    struct sigaction sa;
    std::memset(&sa, 0, sizeof sa);
    sa.sa_handler = forward_signal;
    ::sigaction(SIGINT, &sa, nullptr);
    ::sigaction(SIGTERM, &sa, nullptr);

    // SIGINT / SIGTERM stay blocked from before the fork until the child has put the default actions back:
    // a stop request that arrives between fork and exec (a sibling rank failed at once) then stays pending and
    // terminates the child, instead of being swallowed by the handler it inherited from the launcher
    sigset_t stop_signals, previous_mask;
    ::sigemptyset(&stop_signals);
    ::sigaddset(&stop_signals, SIGINT);
    ::sigaddset(&stop_signals, SIGTERM);
    ::sigprocmask(SIG_BLOCK, &stop_signals, &previous_mask);
    children.assign(static_cast<size_t>(ranks), -1);
    for (long r = 0; r < ranks; ++r)
    {
        const pid_t pid = ::fork();
        if (pid < 0)
        {
            std::fprintf(stderr, "wave-mpirun: fork failed: %s\n", std::strerror(errno));
            forward_signal(SIGTERM);
            break;
        }
        if (pid == 0)
        {
            ::signal(SIGINT, SIG_DFL);
            ::signal(SIGTERM, SIG_DFL);
            ::sigprocmask(SIG_SETMASK, &previous_mask, nullptr);
            ::setenv("WAVE_RANK", std::to_string(r).c_str(), 1);
            ::setenv("WAVE_LOCAL_RANK", std::to_string(r).c_str(), 1);
            ::execvp(cl.program[0], cl.program.data());
            std::fprintf(stderr, "wave-mpirun: cannot execute %s: %s\n", cl.program[0], std::strerror(errno));
            ::_exit(127);
        }
        children[static_cast<size_t>(r)] = pid;
    }
    ::sigprocmask(SIG_SETMASK, &previous_mask, nullptr);

    // the first failure decides the exit status; the surviving ranks would wait for the lost one in
    // their next collective, so they are told to stop
    int first_failure = 0;
    size_t alive = 0;
    for (const pid_t pid : children)
        alive += pid > 0;
    if (alive < children.size())
        first_failure = 1;
    while (alive > 0)
    {
        int status = 0;
        const pid_t pid = ::waitpid(-1, &status, 0);
        if (pid < 0)
        {
            if (errno == EINTR)
                continue;
            break;
        }
        for (pid_t& c : children)
            if (c == pid)
            {
                c = -1;
                --alive;
            }
        const int code = exit_status_of(status);
        if (code != 0 && first_failure == 0)
        {
            first_failure = code;
            forward_signal(SIGTERM);
        }
    }
    ::unlink(rendezvous.c_str());
    ::unlink((rendezvous + ".part").c_str());
    return first_failure;
}
