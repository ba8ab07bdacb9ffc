// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// extern "C" harness around the UNMODIFIED reference (eovs/ldpc-lib) functions on the
// Monte-Carlo BP hot path.  It is compiled by oracle/Makefile together with the
// reference's own translation units, straight from /root/reference (nothing is copied
// into this repository), into oracle/_ref/libldpcref.so.  Tests, the golden-vector
// generator and bench.py's CPU baseline call it through ctypes.
//
// Reference interfaces exercised here:
//   decod_open / decod_init / decod_close        decoders.h:293-295
//   the nine binary *_decod* functions            decoders.h:296-305
//   QAM_demodulator_open / Demodulate             modulation.h:113-115
//   QAM_modulator_open / QAM_modulator            modulation.h:109-111
//   bp_simulation                                 bp_simulation.h:9-27
//   reset_random / next_random_gaussian           commons_portable.h:34-39
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <utility>

#include "decoders.h"
#include "modulation.h"
#include "bp_simulation.h"
#include "commons_portable.h"
#include "data_structures.h"
#include "trace_pm.h"

namespace {

DEC_STATE* open_state(int decoder_id, const short* hd, int b, int c, int M)
{
    DEC_STATE* st = decod_open(decoder_id, 1, b, c, M);
    if (!st) return NULL;
    // the caller fills hd[][] directly, exactly as bp_simulation.cpp:357-362 does
    for (int i = 0; i < b; i++)
        for (int j = 0; j < c; j++)
            st->hd[i][j] = hd[i * c + j];
    if (!decod_init(st)) { decod_close(st); return NULL; }
    return st;
}

// same dispatch and same compile-time decoder parameters as bp_simulation.cpp:716-729
int run_decoder(DEC_STATE* st, int id, int maxiter, int decision = DEC_DECISION)
{
#undef DEC_DECISION
#define DEC_DECISION decision
    switch (id) {
    case BP_DEC:   return bp_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION);
    case SP_DEC:   return sum_prod_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION);
    case ASP_DEC:  return sum_prod_gf2_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION);
    case MS_DEC:   return min_sum_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION, MS_ALPHA);
    case IMS_DEC:  return imin_sum_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION, MS_ALPHA, MS_THR, MS_QBITS, MS_DBITS);
    case IASP_DEC: return isum_prod_gf2_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION);
    case TASP_DEC: return tdmp_sum_prod_gf2_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION);
    case LMS_DEC:  return lmin_sum_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION, MS_ALPHA, MS_BETA);
    case LCHE_DEC: return lche_decod(st, st->y, st->decword, maxiter, DEC_DECISION);
    default:       return -100000;
    }
}

void copy_posterior(DEC_STATE* st, int id, double* out)
{
    int N = st->n;
    switch (id) {
    case BP_DEC:
    case SP_DEC:   for (int i = 0; i < N; i++) out[i] = st->y[i]; break;          // overwritten input
    case ASP_DEC:  for (int i = 0; i < N; i++) out[i] = st->asp_soft_out[i]; break;
    case MS_DEC:   for (int i = 0; i < N; i++) out[i] = st->ms_soft[i]; break;
    case IMS_DEC:  for (int i = 0; i < N; i++) out[i] = st->ims_soft[i]; break;
    case IASP_DEC: for (int i = 0; i < N; i++) out[i] = st->iasp_soft_out[i]; break;
    case TASP_DEC: for (int i = 0; i < N; i++) out[i] = st->tasp_soft_out[i]; break;
    case LMS_DEC:  for (int i = 0; i < N; i++) out[i] = st->lms_soft[i]; break;
    case LCHE_DEC: for (int i = 0; i < N; i++) out[i] = st->lche_soft_out[i]; break;
    }
}

} // namespace

extern "C" {

// Decode n_frames frames of N=c*M channel LLRs with reference decoder `decoder_id`.
//   fresh_state != 0 : a new DEC_STATE per frame (no carry-over between frames)
//   fresh_state == 0 : one DEC_STATE for the whole batch, frame after frame, as bp_simulation does
// Outputs: hard[f*N+i] in {0,1}, iters[f] = the decoder's return value,
//          posterior (optional) = the decoder's final soft state (see copy_posterior),
//          aux (optional, IMS only) = the quantised channel values ims_y.
int ref_decode(int decoder_id, const short* hd, int b, int c, int M,
               const double* llr, int n_frames, int maxiter, int fresh_state,
               unsigned char* hard, int* iters, double* posterior, double* aux)
{
    int N = c * M;
    DEC_STATE* st = NULL;
    for (int f = 0; f < n_frames; f++) {
        if (!st) st = open_state(decoder_id, hd, b, c, M);
        if (!st) return -1;
        memcpy(st->y, llr + (size_t)f * N, N * sizeof(double));
        int it = run_decoder(st, decoder_id, maxiter);
        iters[f] = it;
        if (hard)
            for (int i = 0; i < N; i++) hard[(size_t)f * N + i] = st->decword[i] != 0.0;
        if (posterior) copy_posterior(st, decoder_id, posterior + (size_t)f * N);
        if (aux && decoder_id == IMS_DEC)
            for (int i = 0; i < N; i++) aux[(size_t)f * N + i] = st->ims_y[i];
        if (fresh_state) { decod_close(st); st = NULL; }
    }
    if (st) decod_close(st);
    return 0;
}

// The caller-visible arrays after ONE call with the given `decision` (fresh state): decword[N] as the decoder left it
// (0 / 1 or soft values) and soft_after[N] = the input array st->y after the call (some decoders overwrite it).
int ref_decode_arrays(int decoder_id, const short* hd, int b, int c, int M, const double* llr, int maxiter, int decision,
                      double* decword, double* soft_after, int* iter)
{
    int N = c * M;
    DEC_STATE* st = open_state(decoder_id, hd, b, c, M);
    if (!st) return -1;
    memcpy(st->y, llr, N * sizeof(double));
    *iter = run_decoder(st, decoder_id, maxiter, decision);
    memcpy(decword, st->decword, N * sizeof(double));
    memcpy(soft_after, st->y, N * sizeof(double));
    decod_close(st);
    return 0;
}

// The reference's interleaver as bp_simulation.cpp:417-425, 573, 684 uses it, applied to index ramps: direct[j] = index of the
// input element that Permutation(direction 0) puts at j, inverse[i] likewise for direction 1; -1 where nothing is written.
int ref_permutation(const short* hd, int b, int c, int M, int QAM, int halfmlog, int mode, int block, int step, int* direct, int* inverse)
{
    PERMSTATE* st = Permutations_Open(b, c, M, QAM, halfmlog, mode, block, step);
    if (!st) return -1;
    std::vector<short*> rows(b);
    std::vector<short> flat(hd, hd + (size_t)b * c);
    for (int i = 0; i < b; i++) rows[i] = flat.data() + (size_t)i * c;
    Permutation_Init(st, rows.data());
    int N = c * M;
    std::vector<double> in(N + N), out(N + N, -1.0);
    for (int i = 0; i < N; i++) in[i] = i;
    Permutation(st, 0, in.data(), out.data());
    for (int i = 0; i < N; i++) direct[i] = (int)out[i];
    std::fill(out.begin(), out.end(), -1.0);
    Permutation(st, 1, in.data(), out.data());
    for (int i = 0; i < N; i++) inverse[i] = (int)out[i];
    Permutations_Close(st);
    return 0;
}

// trace_bound_pol_mon_pm() as main_simulation.cpp:148-205 (trace_matrix) calls it: S[20] cycle counts, SA[20] ACE.
int ref_girth(const int* H, int b, int c, int M, int gtarget, int* S, int* SA)
{
    int minACE[GMAX], maxACEspec[GMAX];
    for (int i = 0; i < GMAX; i++) { S[i] = 0; SA[i] = 0; minACE[i] = 0; maxACEspec[i] = 100000000; }
    ARRAY matr;
    matr.ndim = 2;
    put_nrow(&matr, b);
    put_ncol(&matr, c);
    put_addr(&matr, Alloc2d_int(b, c));
    for (int i = 0; i < b; i++)
        for (int j = 0; j < c; j++) matr.addr[i][j] = H[i * c + j];
    trace_bound_pol_mon_pm(matr, M, GMAX, gtarget, S, SA, minACE, maxACEspec);
    free(matr.addr);
    return 0;
}

// Demodulate() at the function boundary, with the CORRECT m = log2(Q)
// (bp_simulation.cpp:471 passes tailbite_length instead).
//   pMod: 2*ns doubles (I,Q interleaved), pRes: n = ns*m doubles.
int ref_demodulate(int Q, int m, int n, int ns, double sigma, double T, int out_type,
                   const double* pMod, double* pRes)
{
    QAM_DEMODULATOR_STATE* st = QAM_demodulator_open(T, sigma, (short)Q, n, m, ns, out_type);
    if (!st) return -1;
    std::vector<double> in(pMod, pMod + 2 * (size_t)ns);
    Demodulate(st, in.data(), pRes);
    QAM_demodulator_close(st);
    return 0;
}

// QAM_modulator(): L bits (as doubles 0/1) -> 2*ns lattice coordinates.
int ref_modulate(int Q, int L, int m, const double* bits, double* out)
{
    QAM_MODULATOR_STATE* st = QAM_modulator_open(Q, L, m);
    if (!st) return -1;
    std::vector<double> in(bits, bits + L);
    in.resize(st->Lfact, 0.0);
    QAM_modulator(st, in.data(), out);
    int ns = st->ns;
    QAM_modulator_close(st);
    return ns;
}

// The reference's whole simulation kernel for one (code, SNR) point.
int ref_bp_simulation(const int* H, int b, int c, int M, int max_iterations,
                      int n_frame_errors, int n_experiments, double snr,
                      double reference_frame_error, int decoder_type, int modulation_type,
                      int punctured_blocks, int seed, double* ber, double* fer)
{
    matrix<int> HM(b, c), HC(b, c);
    for (int i = 0; i < b; i++)
        for (int j = 0; j < c; j++) HM(i, j) = H[i * c + j];
    initial_random_seed = seed;
    reset_random();
    std::pair<double, double> r = bp_simulation(2, HM, HC, 0, M, max_iterations, n_frame_errors,
                                                n_experiments, snr, reference_frame_error,
                                                decoder_type, modulation_type, 0, 1, 1,
                                                punctured_blocks, 0);
    *ber = r.first;
    *fer = r.second;
    return 0;
}

// random_codeword() (bp_simulation.cpp:142-192) right after reset_random() with the given seed.
// out: c*M bytes; returns the reference's exit code (0 ok, < 0 bad matrix, > 0 bad encoding).
int ref_random_codeword(const int* H, int b, int c, int M, int seed, unsigned char* out)
{
    matrix<int> HM(b, c);
    for (int i = 0; i < b; i++)
        for (int j = 0; j < c; j++) HM(i, j) = H[i * c + j];
    initial_random_seed = seed;
    reset_random();
    std::vector<bit> cw;
    int rc = random_codeword(HM, M, cw);
    if (rc == 0)
        for (size_t i = 0; i < cw.size(); i++) out[i] = (bool)cw[i];
    return rc;
}

// n samples of the reference's N(0,1) stream (commons_portable.cpp:174-178).
void ref_gaussian(int seed, int n, double* out)
{
    initial_random_seed = seed;
    reset_random();
    for (int i = 0; i < n; i++) out[i] = next_random_gaussian();
}

} // extern "C"
