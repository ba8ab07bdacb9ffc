// bp_simulation() of the reference (bp_simulation.h:9-27, bp_simulation.cpp:305-841) on the B200 engine.
#pragma once
#include <utility>
#ifdef LDPCB200_WITH_REFERENCE_HEADERS
// Level-2 integration: this file is compiled inside ldpc-lib's own tree, against ITS matrix<T>, die()
// and seed globals (data_structures.h:10-61, commons_portable.h)
#include "data_structures.h"
#include "commons_portable.h"
#else
#include "commons.h"
#endif

enum MODULATION_TYPE { MODULATION_SKIP = 0, MODULATION_QAM4, MODULATION_QAM16, MODULATION_QAM64, MODULATION_QAM256 };   // modulation.h:4-11

// Same signature and return value as the reference: (BER over the information bits, FER), or
// (-1, -1) if the run was interrupted.  See bp_simulation.cpp for what differs inside.
std::pair<double, double> bp_simulation(
    int q_mod,
    matrix<int> const& code_generating_matrix,
    matrix<int>& coef_matrix,
    int ncols2convert,
    int tailbite_length,
    int max_iterations,
    int n_frame_errors,
    int n_experiments,
    double snr,
    double reference_frame_error,
    int decoder_type,
    int modulation_type,
    int permutation_type,
    int permutation_block,
    int permutation_inter,
    int punctured_blocks,
    int show_process);

// The reference's QC encoder (bp_simulation.h:29-33; bp_simulation.cpp:22-192): a random codeword of the code with base
// matrix mx (parity block columns first, uni-/bi-diagonal parity structure).  Returns 0, < 0 ("bad matrix": no positive
// shift in the last bidiagonal column) or > 0 ("bad encoding": the structure is not one the encoder understands).
// bp_simulation() itself transmits the all-zero codeword, as the reference does (bp_simulation.cpp:568); the search
// drivers call this function directly to reject matrices.
#include <vector>
int random_codeword(matrix<int> const& mx, int tailbite_length, std::vector<bit>& codeword);
int qc_encode(matrix<int> const& mx, int tailbite_length, std::vector<bit>& cword);

// Bookkeeping of the last bp_simulation() call (an addition: the reference prints nothing comparable).
struct bp_simulation_stats {
    long long frames_counted;       // `experiment`: frames that entered the result
    long long frames_decoded;       // frames the GPUs decoded (rounds are decoded whole)
    long long frame_errors, info_bit_errors, undetected;
    double seconds;                 // wall time of the frame loop
    double gpu_ms;                  // sum over rounds of the slowest GPU's kernel time
    int gpus;
};
bp_simulation_stats const& bp_simulation_last_stats();

// The reference lets the operator interrupt the current (code, SNR) point with the 'x' console key (console_exception_hook,
// bp_simulation.cpp:590, :825-829) and then returns (-1, -1).  The drop-in has no console hook; a caller (or a signal handler: the
// function only sets a flag) asks for the same through this call: the running -- or, if none is running, the next --
// bp_simulation() returns (-1, -1) at its next round boundary and clears the request.
void bp_simulation_request_interrupt();
