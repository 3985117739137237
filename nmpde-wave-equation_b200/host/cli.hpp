// cli.hpp -- shared command-line driver of main-newmark / main-theta.
//
// Behaviour kept from the reference's two mains (src/main-newmark.cpp:24-169, src/main-theta.cpp:23-152):
// one optional positional argument (default ../parameters/sine-membrane.json), the parsed-parameter
// echo, NMPDE_SAVE_SOLUTION / NMPDE_LOG_EVERY (and NMPDE_PARAM_FILE for Newmark) exported to the
// solver classes, std::invalid_argument and std::exception mapped to a message and exit code 1.
#ifndef WAVE_CLI_HPP
#define WAVE_CLI_HPP

enum class Scheme
{
    Newmark,
    Theta
};

int wave_cli_main(int argc, char* argv[], Scheme scheme);

#endif
