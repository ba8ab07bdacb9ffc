/* ldpcb200 -- C ABI of the B200-native engine for ldpc-lib's Monte-Carlo BP simulation hot path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, caller-owned buffers.  The
 * reference (eovs/ldpc-lib) is a single C++ binary with no FFI of its own; the entry points below
 * are what its kept host entry points bind to:
 *
 *   reference interface (file:line)                          replaced by
 *   -------------------------------------------------------  -----------------------------------
 *   decod_open + caller fills st->hd + decod_init             ldpcb200_create
 *       decoders.h:293-294, bp_simulation.cpp:353-382
 *   decod_close                        decoders.h:295         ldpcb200_destroy
 *   the nine binary *_decod* functions decoders.h:296-305     ldpcb200_decode_batch (1..n frames)
 *   Demodulate                         modulation.h:115       ldpcb200_demodulate
 *   QAM_modulator                      modulation.h:111       ldpcb200_modulate
 *   Permutations_Open / _Init / Permutation                   ldpcb200_interleaver_tables,
 *       direct_inverse_perm.cpp:139, :312, :785               ldpcb200_set_interleaver
 *   bp_simulation's frame loop         bp_simulation.cpp:591-824
 *       (noise -> LLR -> puncture -> decode -> count)         ldpcb200_simulate
 *   the search caller's loop over candidate codes             ldpcb200_simulate_codes
 *       main_good_code_search.cpp:267-411
 *
 * Conventions kept from the reference: LLR = log P(0)/P(1), positive => bit 0
 * (bp_simulation.cpp:603); a frame is N = c*Z values, block column i at [i*Z, (i+1)*Z), parity
 * block columns first (bits >= R = b*Z are information bits, bp_simulation.cpp:738); the iteration
 * count returned per frame follows each reference decoder's own return convention (>= 0 success,
 * < 0 = -iterations on failure).
 *
 * Threading: one handle per (GPU, code, decoder); a handle must not be used from two host threads
 * at once.  All functions return 0 on success or a negative LDPCB200_E* code;
 * ldpcb200_last_error() gives the message of the calling thread's last failure.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * LDPCB200_ENODEV.
 */
#ifndef LDPCB200_H
#define LDPCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPCB200_VERSION 100
#define LDPCB200_MAX_ROW_WEIGHT 32   /* largest base-matrix row weight a handle accepts */

/* enum DEC_ID, decoders.h:16-28 */
enum ldpcb200_decoder {
    LDPCB200_BP_DEC = 0, LDPCB200_SP_DEC = 1, LDPCB200_ASP_DEC = 2, LDPCB200_MS_DEC = 3,
    LDPCB200_IMS_DEC = 4, LDPCB200_IASP_DEC = 5, /* 6 = FHT_DEC (GF(q)): out of scope */
    LDPCB200_TASP_DEC = 7, LDPCB200_LMS_DEC = 8, LDPCB200_LCHE_DEC = 9
};

/* enum MODULATION_TYPE, modulation.h:4-11 */
enum ldpcb200_modulation {
    LDPCB200_MOD_BPSK = 0, LDPCB200_MOD_QAM4 = 1, LDPCB200_MOD_QAM16 = 2,
    LDPCB200_MOD_QAM64 = 3, LDPCB200_MOD_QAM256 = 4
};

enum ldpcb200_dtype { LDPCB200_F64 = 0, LDPCB200_F32 = 1, LDPCB200_I16 = 2, LDPCB200_U16 = 3 };

enum ldpcb200_error {
    LDPCB200_OK = 0, LDPCB200_EINVAL = -1, LDPCB200_ENODEV = -2, LDPCB200_ECUDA = -3,
    LDPCB200_ENOMEM = -4, LDPCB200_EUNSUPPORTED = -6
};

enum ldpcb200_flags {
    LDPCB200_LLR_ON_DEVICE   = 1 << 0,  /* llr points to device memory                           */
    LDPCB200_OUT_ON_DEVICE   = 1 << 1,  /* hard / iters / posterior point to device memory       */
    LDPCB200_HARD_PACKED     = 1 << 2,  /* hard = ceil(N/32) uint32 words per frame, bit i of the
                                           frame at word i/32, bit i%32; default is one byte/bit  */
    LDPCB200_NO_EARLY_EXIT   = 1 << 3,  /* run exactly maxiter iterations, with the syndrome check in
                                           every one (worst-case timing).  iters reports the count the
                                           reference would return (its first zero syndrome); hard
                                           decisions, posteriors and error counters are those AFTER
                                           the extra iterations                                        */
    LDPCB200_BP_CHAIN_SYNDROME = 1 << 4 /* BP_DEC only: carry the previous frame's syndrome into the
                                           pre-iteration check, as decoders.cpp:1742-1759 does     */
};

typedef struct ldpcb200_handle_s* ldpcb200_handle;

/* Decoder parameters.  Zero-initialise, then call ldpcb200_default_params(). */
typedef struct {
    double alpha;      /* MS_DEC / IMS_DEC normalisation, MS_ALPHA = 0.8        decoders.h:43   */
    double beta;       /* kept for signature parity; LMS_DEC ignores it (hard-coded 0.4,
                          decoders.cpp:5163)                                     decoders.h:44   */
    double thr;        /* IMS_DEC quantiser threshold, MS_THR = 1.4              decoders.h:46   */
    int    qbits;      /* IMS_DEC input bits, MS_QBITS = 6                       decoders.h:47   */
    int    dbits;      /* IMS_DEC message bits, MS_DBITS = 8                     decoders.h:48   */
    int    precision;  /* arithmetic of the float decoders: 64 (default: the reference's double,
                          bit-exact for LMS/MS) or 32 (LMS/MS only: same decisions and iteration
                          counts on >= 99.99 % of frames)                         */
    int    device;     /* CUDA device ordinal, -1 = current device                               */
    int    use_fast;   /* 0: the table-driven parity kernels only (one frame per CTA, state in an
                          L2-resident workspace); 1 (default): the shared-memory throughput kernels
                          where one exists (LMS_DEC) and the code fits, code-specialised when the
                          matrix is one of the built-in benchmark matrices; 2: additionally compile a
                          code-specialised kernel for THIS matrix at create time (NVRTC, a few
                          seconds once per matrix and process; falls back to 1 if NVRTC is absent)  */
    int    reserved[8];
} ldpcb200_params;

/* One (code, SNR) simulation round: frames [first_frame, first_frame + n_frames) of the stream
 * keyed by (seed, stream).  Noise sample i of frame f is Philox4x32-10(key = seed,stream;
 * counter = f, i/4) -> Box-Muller, so results do not depend on batch size or on how frames are
 * sharded over GPUs. */
typedef struct {
    double   snr_db;            /* Eb/N0 in dB, bp_simulation.cpp:445                            */
    int      modulation;        /* enum ldpcb200_modulation                                       */
    int      punctured_blocks;  /* last block columns, LLR := 0.5 or 0, bp_simulation.cpp:697-710 */
    int      max_iterations;
    uint64_t seed;
    uint32_t stream;            /* e.g. SNR index                                                 */
    uint64_t first_frame;
    uint32_t n_frames;
    uint32_t flags;             /* LDPCB200_NO_EARLY_EXIT | LDPCB200_OUT_ON_DEVICE (per_frame)    */
    double   qam_T;             /* Demodulate clip T, 26.0 in bp_simulation.cpp:339; 0 => 26.0    */
} ldpcb200_sim_params;

/* Sums over the round's frames (bp_simulation.cpp:731-743, 805-810). */
typedef struct {
    uint64_t frames;
    uint64_t frame_errors;      /* nde: frames with >= 1 wrong bit among all N                    */
    uint64_t info_bit_errors;   /* nse: wrong bits with index >= R, summed over erroneous frames  */
    uint64_t undetected;        /* nue: erroneous frames whose decoder returned >= 0              */
    uint64_t iter_sum;          /* sum of |iterations| run                                        */
    uint64_t bit_errors;        /* wrong bits among all N                                         */
} ldpcb200_counters;

void        ldpcb200_default_params(ldpcb200_params* p);
const char* ldpcb200_last_error(void);
int         ldpcb200_version(void);

/* hd: b*c shorts, row-major, -1 = no circulant, otherwise the shift (reduced mod Z). */
int ldpcb200_create(const int16_t* hd, int b, int c, int Z, int decoder_id,
                    const ldpcb200_params* params, ldpcb200_handle* out);
int ldpcb200_destroy(ldpcb200_handle h);

/* Geometry of a handle: N, R, number of edges E, and the CUDA device it lives on. */
int ldpcb200_info(ldpcb200_handle h, int* N, int* R, int* E, int* device);

/* Which kernel a handle runs: *fast = 0 table-driven parity kernel, >= 1 a shared-memory throughput
 * kernel (bits 0-3: variant number + 1; bit 4: its check-to-variable messages live in tensor memory); launch
 * geometry of that kernel. */
int ldpcb200_kernel_info(ldpcb200_handle h, int* fast, int* threads, int* frames_per_cta, int* ctas_per_sm,
                         int* smem_bytes);

/* Decode n_frames frames of N channel LLRs (frame f at llr + f*N elements of llr_dtype F64|F32).
 *   hard      : n_frames*N bytes (0/1), or packed words with LDPCB200_HARD_PACKED; may be NULL
 *   iters     : n_frames int32; may be NULL
 *   posterior : optional n_frames*N values of post_dtype -- the decoder's final soft state
 *               (LMS/MS/BP/LCHE: LLR; ASP/TASP: P(bit=1); SP: likelihood ratio; IMS: int16
 *               ims_soft; IASP: uint16), F64 or F32 for float decoders, I16 / U16 for IMS / IASP
 *   aux       : optional, IMS_DEC only: n_frames*N int16 quantised channel values (ims_y)        */
int ldpcb200_decode_batch(ldpcb200_handle h, const void* llr, int llr_dtype, int n_frames,
                          int maxiter, uint32_t flags, void* hard, int32_t* iters,
                          void* posterior, int post_dtype, void* aux);

/* The same decode on LLRs generated on the device from the handle's channel model.
 * per_frame (optional, host, n_frames uint32): bit 31 = frame in error, bit 30 = decoder reported
 * success, bits 0..23 = info-bit errors; lets the caller apply bp_simulation's stopping rules
 * (bp_simulation.cpp:591, 820) in exact frame order.
 * The in-kernel channel hands the decoder fp32 LLRs.  For QAM-16/64/256 it evaluates Demodulate's likelihood sums in a
 * factored form (one exponential per PAM component; doubles within a few 1e-15 of QAM_demodulator.cpp:181-561, i.e. the same
 * fp32 value but for a last bit about once in 1e7); ldpcb200_generate_llr / ldpcb200_demodulate keep the reference's
 * evaluation order, and LDPCB200_QAM_EXACT=1 in the environment puts it into the kernels as well. */
int ldpcb200_simulate(ldpcb200_handle h, const ldpcb200_sim_params* sp,
                      ldpcb200_counters* out, uint32_t* per_frame);

/* The same round for MANY candidate codes in one launch -- what the code-search caller does with every candidate it
 * generates (main_good_code_search.cpp:267-411: generate_code, then bp_simulation at one SNR, :320-338).  hds: n_codes
 * matrices of the handle's shape (b x c, same lifting, the same number of circulants E; the masks may differ), row-major one
 * after the other.  The handle (TASP_DEC, the decoder files/input32_16.jsonx names) is only the launch plan: its own matrix
 * does not take part, and it is kept across calls.  out: n_codes counter blocks; per_frame (optional, host): n_codes x
 * n_frames records as in ldpcb200_simulate.  Every code sees the same noise (frames first_frame .. of stream sp->stream):
 * common random numbers, as the reference's reset_random() before each candidate (:316) gives it.
 * LDPCB200_EUNSUPPORTED if the handle is not a TASP_DEC handle on the tensor-memory kernel or a matrix does not fit it. */
int ldpcb200_simulate_codes(ldpcb200_handle h, int n_codes, const int16_t* hds, const ldpcb200_sim_params* sp,
                            ldpcb200_counters* out, uint32_t* per_frame);

/* The channel LLRs ldpcb200_simulate would feed the decoder, written out (host, F32 or F64) so the
 * identical buffer can be given to the reference. */
int ldpcb200_generate_llr(ldpcb200_handle h, const ldpcb200_sim_params* sp, void* llr, int llr_dtype);

/* Demodulate() at the function boundary, QAM_demodulator.cpp:99-566, with m = log2(Q).
 * x: 2*ns doubles (I,Q interleaved), res: ns*log2(Q) doubles; out_type 0 = log(P1/P0), 1 = P1.
 * Host pointers. */
int ldpcb200_demodulate(int Q, int ns, double sigma, double T, int out_type,
                        const double* x, double* res, int device);

/* QAM_modulator(), QAM_modulator.cpp:142-194: ns*log2(Q) bits (bytes 0/1) -> 2*ns coordinates. */
int ldpcb200_modulate(int Q, int ns, const uint8_t* bits, double* out, int device);

/* sigma / sigmaQAM of bp_simulation.cpp:444-449 (host arithmetic, no device needed). */
double ldpcb200_sigma(int b, int c, int punctured_blocks, double snr_db, int modulation);

/* Girth, ACE spectrum and cycle spectrum of a base matrix, as the reference's driver computes them before every
 * simulation (main_simulation.cpp:148-205 trace_matrix -> trace_pm.cpp:58 trace_bound_pol_mon_pm with GMAX = 20):
 * *girth = shortest cycle length, ace[k] / spectrum[k] = ACE value and number of protograph cycles of the k-th
 * shortest cycle length present, k < gtarget (the driver uses GTARGET = 4).  Host arithmetic, no device needed. */
int ldpcb200_girth_spectrum(const int16_t* hd, int b, int c, int Z, int gtarget, int* girth, int* ace, int* spectrum);

/* Bit interleaver tables of the simulation path: Permutations_Open / Permutation_Init / Permutation,
 * direct_inverse_perm.cpp:139, :312-782, :785-896 with the LCG myrand :131-135, as bp_simulation.cpp:417-425 opens them.
 * modulation = enum ldpcb200_modulation (decides the bits per PAM component mode 2 deals the block columns to), mode =
 * the scenario's permutation_type 0..4, block / inter = permutation_block / permutation_inter.  direct[j] = codeword bit
 * sent at transmitted position j (Permutation direction 0, :573), inverse[i] = transmitted position that feeds decoder
 * input i (direction 1, :684); N = c*Z entries each.  LDPCB200_EUNSUPPORTED where the reference's parameters do not
 * define a permutation of the N positions (it would read stale memory).  Host arithmetic, no device needed. */
int ldpcb200_interleaver_tables(const int16_t* hd, int b, int c, int Z, int modulation, int mode, int block, int inter,
                                int32_t* direct, int32_t* inverse);

/* Attach an interleaver to a handle: from then on ldpcb200_simulate / ldpcb200_generate_llr feed decoder input i with
 * the LLR received at transmitted position inverse[i] (a gather / scatter inside the decoder's first load; puncturing
 * is applied after it, bp_simulation.cpp:697-710).  direct / inverse: host arrays of N entries, mutually inverse
 * permutations (checked); both NULL removes the interleaver. */
int ldpcb200_set_interleaver(ldpcb200_handle h, const int32_t* direct, const int32_t* inverse);

/* The transmitted codeword of ldpcb200_simulate / ldpcb200_generate_llr: N bytes 0 / 1 in codeword order (host), NULL or
 * all-zero = the all-zero codeword the reference's loop sends (bp_simulation.cpp:567).  With a real codeword (host/encoder.cpp
 * qc_encode / random_codeword make them, bp_simulation.cpp:22-192) the bits go through the direct permutation (:573), the Gray
 * map of QAM_modulator (QAM_modulator.cpp:127-194) or the BPSK map (:600-612), the channel, Demodulate, the inverse
 * permutation and the puncturing; errors are counted against the codeword (:731-743).  The round then runs as three launches
 * (LLRs into a device buffer, decode, compare) instead of the fused one. */
int ldpcb200_set_codeword(ldpcb200_handle h, const uint8_t* bits);

/* The generator's unit-variance noise samples 0 .. n_samples-1 of frames first_frame .. first_frame + n_frames - 1 (host, fp32,
 * frame-major): sample t is the one added to transmitted position t (BPSK / QAM-4) or to PAM component t = 2 symbol +
 * component (QAM-16/64/256).  For tests that rebuild the channel with the reference's own modulator and demodulator. */
int ldpcb200_generate_noise(ldpcb200_handle h, const ldpcb200_sim_params* sp, int n_samples, float* out);

/* Diagnostic: generate and compile (NVRTC, no device needed) the code-specialised LMS_DEC kernel for a matrix
 * and target architecture sm_<major><minor>; *cubin_bytes = size of the result. */
int ldpcb200_jit_check(const int16_t* hd, int b, int c, int Z, int sm_major, int sm_minor, int* cubin_bytes);

/* Kernel timing of the last decode_batch / simulate call on this handle, measured with CUDA events
 * on the handle's stream around the decode kernel launches only. */
int ldpcb200_last_kernel_ms(ldpcb200_handle h, float* ms, int* launches);

/* The handle's CUDA stream (cudaStream_t) so callers can order their own work with it. */
void* ldpcb200_stream(ldpcb200_handle h);

#ifdef __cplusplus
}
#endif
#endif /* LDPCB200_H */
