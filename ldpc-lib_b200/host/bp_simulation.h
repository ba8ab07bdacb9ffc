// bp_simulation() of the reference (bp_simulation.h:9-27, bp_simulation.cpp:305-841) on the B200 engine.
#pragma once
#include <utility>
#include "commons.h"

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
