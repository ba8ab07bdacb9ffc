// `main <entry-point> args...` -- the reference's dispatcher (main.cpp:27-44).  Only the simulation entry
// point is on the accelerated hot path; the code-search entry points are the reference's own CPU programs
// and are not rebuilt here (out of scope).
#include <cstdio>
#include <cstring>

int main_simulation(int argc, char* argv[]);

int main(int argc, char* argv[])
{
    if (argc <= 1) {
        fprintf(stderr, "Error: no entry point specified\nUsage: %s <entry-point> args...\n    where <entry-point> is one of the following:\n        simulation\n", argv[0]);
        return 1;
    }
    if (!strcmp(argv[1], "simulation")) return main_simulation(argc - 1, argv + 1);
    if (!strcmp(argv[1], "search") || !strcmp(argv[1], "tests") || !strcmp(argv[1], "ggp")) {
        fprintf(stderr, "Error: entry point '%s' belongs to ldpc-lib's CPU code-search side and is not part of the B200 engine; use the reference binary for it\n", argv[1]);
        return 2;
    }
    fprintf(stderr, "Error: unknown entry point '%s'\n", argv[1]);
    return 1;
}
