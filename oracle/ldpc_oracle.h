/* TEST INFRASTRUCTURE ONLY -- the CPU oracle for ldpc-lib's Monte-Carlo BP hot path.
 *
 * A plain-C restatement, in this repository's own words, of the reference algorithms
 * (eovs/ldpc-lib; file:line citations are into the reference tree).  It exists so the CUDA
 * kernels can be checked against something that runs anywhere; it is validated against the
 * compiled reference itself (oracle/_ref, built by oracle/Makefile) and against the golden
 * vectors in tests/golden/ that were produced by that reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * The product (ldpc-lib_b200/) never includes, links or calls anything in oracle/.
 *
 * Formulation: the reference walks the dense b x c base matrix and rotates whole circulant
 * columns with memcpy (decoders.cpp:327-346).  Here every decoder is written over per-row
 * edge lists (col, shift) with one independent "lane" per check row n of a block row:
 * lane n of block row j touches bit col*Z + (n+shift) mod Z.  Lanes of one block row touch
 * disjoint bits, so the reference's loop order over n is immaterial; the order over block rows
 * (layers) and over edges inside a row IS kept, because floating-point sums and the integer
 * saturating sums depend on it.
 *
 * All functions decode `nf` frames laid out back to back (frame f at y + f*N) and return 0.
 * iters[f] holds the reference decoder's return value (sign conventions in SURVEY.md §8a).
 */
#ifndef LDPC_ORACLE_H
#define LDPC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* decoder ids = enum DEC_ID, decoders.h:16-28 */
enum { ORC_BP = 0, ORC_SP = 1, ORC_ASP = 2, ORC_MS = 3, ORC_IMS = 4, ORC_IASP = 5,
       ORC_TASP = 7, ORC_LMS = 8, ORC_LCHE = 9 };

/* layered offset min-sum, lmin_sum_decod_qc_lm, decoders.cpp:5064-5425 (beta = 0.4 hard-coded) */
int orc_lms_f64(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
                uint8_t* hard, int32_t* iters, double* post);
int orc_lms_f32(const int16_t* hd, int b, int c, int Z, const float* y, int nf, int maxiter,
                uint8_t* hard, int32_t* iters, float* post);

/* flooding normalised min-sum, min_sum_decod_qc_lm (MS_MUL_CORRECTION variant), decoders.cpp:4554-4767 */
int orc_ms_f64(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
               double alpha, uint8_t* hard, int32_t* iters, double* post);
int orc_ms_f32(const int16_t* hd, int b, int c, int Z, const float* y, int nf, int maxiter,
               float alpha, uint8_t* hard, int32_t* iters, float* post);

/* fixed-point min-sum, imin_sum_decod_qc_lm, decoders.cpp:5430-5690.
 * post = ims_soft, iy = quantised channel values ims_y (both int16, optional). */
int orc_ims(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
            double alpha, double thr, int qbits, int dbits,
            uint8_t* hard, int32_t* iters, int16_t* post, int16_t* iy);

/* layered sum-product, probability domain, tdmp_sum_prod_gf2_decod_qc_lm, decoders.cpp:2584-2744.
 * post = tasp_soft_out = P(bit = 1). */
int orc_tasp(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
             uint8_t* hard, int32_t* iters, double* post);

/* flooding sum-product, probability domain, sum_prod_gf2_decod_qc_lm, decoders.cpp:2324-2581 */
int orc_asp(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
            uint8_t* hard, int32_t* iters, double* post);

/* Gallager log-domain flooding BP, bp_decod_qc_lm, decoders.cpp:1708-1920.
 * chain != 0 reproduces the reference's stale st->syndr carried from frame to frame
 * (decoders.cpp:1742-1759); chain == 0 starts every frame from a zeroed syndrome. */
int orc_bp(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
           int chain, uint8_t* hard, int32_t* iters, double* post);

/* flooding sum-product, likelihood-ratio domain, sum_prod_decod_qc_lm, decoders.cpp:1923-2185 */
int orc_sp(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
           uint8_t* hard, int32_t* iters, double* post);

/* layered table-lookup decoder, lche_decod, decoders.cpp:2893-3010 */
int orc_lche(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
             uint8_t* hard, int32_t* iters, double* post);

/* 12-bit fixed-point flooding sum-product, isum_prod_gf2_decod_qc_lm, decoders.cpp:3822-4121.
 * post = iasp_soft_out (ui16). */
int orc_iasp(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
             uint8_t* hard, int32_t* iters, uint16_t* post);

/* Demodulate(), QAM_demodulator.cpp:99-566, called with m = log2(Q).  x: 2*ns (I,Q interleaved). */
int orc_demodulate(int m, int ns, double sigma, double T, int out_type, const double* x, double* res);

/* QAM_modulator(), QAM_modulator.cpp:142-194: ns*m bits -> 2*ns lattice coordinates. */
int orc_modulate(int m, int ns, const uint8_t* bits, double* out);

/* BPSK/AWGN LLR, bp_simulation.cpp:600-605, and sigma, bp_simulation.cpp:444-449 */
double orc_sigma_bpsk(double snr_db, int b, int c, int punctured_blocks);
double orc_sigma_qam(double snr_db, int b, int c, int punctured_blocks, int Q);
void orc_bpsk_llr(const double* noise, const uint8_t* cw, int n, double sigma, double* llr);

/* error counting of one frame, bp_simulation.cpp:731-743: all-bit and info-bit (index >= R) errors
 * against codeword cw (NULL = all-zero). */
void orc_count_errors(const uint8_t* hard, const uint8_t* cw, int N, int R, int* nse, int* nse_info);

#ifdef __cplusplus
}
#endif
#endif
