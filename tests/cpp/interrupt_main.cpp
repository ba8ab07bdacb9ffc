// bp_simulation() of the drop-in host layer: an interrupt request makes the call return (-1, -1) like the reference's 'x'
// console hook (bp_simulation.cpp:590, :825-829), and the next call runs normally.
//   interrupt_main <b> <c> <Z> <hd.bin int16>     prints "ber fer" of the interrupted and of the following call
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "bp_simulation.h"
#include "decoders.h"

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    const int b = atoi(argv[1]), c = atoi(argv[2]), Z = atoi(argv[3]);
    std::vector<short> hd((size_t)b * c);
    FILE* f = fopen(argv[4], "rb");
    if (!f || fread(hd.data(), 2, hd.size(), f) != hd.size()) return 3;
    fclose(f);
    matrix<int> H(b, c), coef(1, 1);
    for (int i = 0; i < b; i++)
        for (int j = 0; j < c; j++) H(i, j) = hd[(size_t)i * c + j];
    bp_simulation_request_interrupt();
    std::pair<double, double> r1 = bp_simulation(2, H, coef, 0, Z, 10, 20, 100000, 2.0, 1.0, LMS_DEC, 0, 0, 1, 1, 0, 0);
    std::pair<double, double> r2 = bp_simulation(2, H, coef, 0, Z, 10, 20, 100000, 2.0, 1.0, LMS_DEC, 0, 0, 1, 1, 0, 0);
    printf("%g %g %g %g\n", r1.first, r1.second, r2.first, r2.second);
    return 0;
}
