// main-newmark -- Newmark-beta executable (same name and command line as the reference's target).
#include "cli.hpp"

int main(int argc, char* argv[]) { return wave_cli_main(argc, argv, Scheme::Newmark); }
