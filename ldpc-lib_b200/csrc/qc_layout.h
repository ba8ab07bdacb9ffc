// Device-side quasi-cyclic Tanner-graph layout.
//
// The reference keeps only the b x c matrix of shifts hd[j][i] (decoders.h:146) and re-scans it,
// -1 entries included, on every pass, rotating whole circulant columns with two memcpy
// (decoders.cpp:327-346).  Here the matrix is "lifted" once per handle into packed per-row edge
// lists: block row j owns edges rp[j]..rp[j+1]-1 in ascending block-column order; edge e joins
// check row (j, n) with bit col[e]*Z + (n + sh[e]) mod Z for every lane n in [0, Z).  Column lists
// (ascending block row) serve the flooding decoders, whose sums run over the rows of a column.
// The fast kernels read the packed word pk[e] = col | shift << 8 | local index << 20 (one 32-bit
// uniform load per edge); rows are padded to a common stride in pkpad for unrolled loops.
#pragma once
#include <cstdint>
#include <vector>

namespace ldpcb200 {

struct QcHost {
    int b = 0, c = 0, Z = 0, E = 0, N = 0, R = 0, maxdeg = 0, mindeg = 0, maxcdeg = 0;
    int all_cw_2 = 0;                 // every block column has weight 2 (decod_init, decoders.cpp:1026-1041)
    std::vector<int> rp, col, sh, row, cp, cedge;
    std::vector<uint32_t> pk;         // E packed edge words
    std::vector<int16_t> hd;          // the caller's matrix
    bool build(const int16_t* hd, int b, int c, int Z);
};

// Plain-old-data view handed to kernels by value (all pointers are device pointers).
struct QcDev {
    int b, c, Z, E, N, R, maxdeg, maxcdeg, all_cw_2, nwords;
    const int* rp;      // b+1
    const int* col;     // E
    const int* sh;      // E
    const int* row;     // E
    const int* cp;      // c+1
    const int* cedge;   // E
    const uint32_t* pk; // E
};

} // namespace ldpcb200
