// `main <entry-point> args...` -- the reference's dispatcher (main.cpp:27-44).  Only the simulation entry
// point is on the accelerated hot path; the code-search entry points are the reference's own CPU programs
// and are not rebuilt here (out of scope).
#include <csignal>
#include <cstdio>
#include <cstring>

int main_simulation(int argc, char* argv[]);
void bp_simulation_request_interrupt();

// SIGUSR1 stands in for the reference's 'x' console key: the current (code, SNR) point ends with (-1, -1), the sweep goes on
static void on_sigusr1(int) { bp_simulation_request_interrupt(); }

int main(int argc, char* argv[])
{
    if (argc <= 1) {
        fprintf(stderr, "Error: no entry point specified\nUsage: %s <entry-point> args...\n    where <entry-point> is one of the following:\n        simulation\n", argv[0]);
        return 1;
    }
    if (!strcmp(argv[1], "simulation")) { signal(SIGUSR1, on_sigusr1); return main_simulation(argc - 1, argv + 1); }
    if (!strcmp(argv[1], "search") || !strcmp(argv[1], "tests") || !strcmp(argv[1], "ggp")) {
        fprintf(stderr, "Error: entry point '%s' belongs to ldpc-lib's CPU code-search side and is not part of the B200 engine; use the reference binary for it\n", argv[1]);
        return 2;
    }
    fprintf(stderr, "Error: unknown entry point '%s'\n", argv[1]);
    return 1;
}
