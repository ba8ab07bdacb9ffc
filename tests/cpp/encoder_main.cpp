// random_codeword() of the host layer, called the way the reference's drivers call it
// (scenario_based_code_generation.cpp:505, bp_simulation.cpp:512): argv = <b> <c> <M> <seed> <hd.bin int32> <out.bin>;
// out = int32 exit code followed by c*M bytes (only when the code is 0).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "bp_simulation.h"

int main(int argc, char** argv)
{
    if (argc != 7) return 2;
    int b = atoi(argv[1]), c = atoi(argv[2]), M = atoi(argv[3]);
    initial_random_seed = atoi(argv[4]);
    reset_random();
    std::vector<int> hd((size_t)b * c);
    FILE* f = fopen(argv[5], "rb");
    if (!f || fread(hd.data(), 4, hd.size(), f) != hd.size()) return 3;
    fclose(f);
    matrix<int> H(b, c);
    for (int i = 0; i < b; i++)
        for (int j = 0; j < c; j++) H(i, j) = hd[(size_t)i * c + j];
    std::vector<bit> cw;
    int rc = random_codeword(H, M, cw);
    FILE* out = fopen(argv[6], "wb");
    fwrite(&rc, 4, 1, out);
    if (rc == 0)
        for (size_t i = 0; i < cw.size(); i++) { unsigned char v = (bool)cw[i]; fwrite(&v, 1, 1, out); }
    fclose(out);
    return 0;
}
