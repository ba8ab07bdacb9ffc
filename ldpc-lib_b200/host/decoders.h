// The reference's decoder interface (decoders.h:16-28, 132-308) over the B200 engine.
//
// Same entry points, argument meaning and return conventions as the reference, so code written
// against ldpc-lib's decoders.h keeps compiling and behaving:
//   DEC_STATE* st = decod_open(id, q_bits, b, c, M);  st->hd[j][i] = shift or -1;  decod_init(st);
//   iter = lmin_sum_decod_qc_lm(st, st->y, st->decword, maxiter, 0, alpha, beta);   decod_close(st);
// The DEC_STATE members a caller may touch (bp_simulation.cpp:357-362, 702-736) are kept: hd, y,
// decword, syndr, m, rh, nh, n, codec_id, q_bits; the per-decoder scratch arrays of the reference are
// gone (that state lives on the GPU).  Each *_decod* call decodes ONE frame through
// ldpcb200_decode_batch; callers with many frames should use decod_batch() below or the C ABI itself.
// There is no CPU implementation behind these functions: without a CUDA device decod_init fails.
#pragma once
#include "../../include/ldpcb200.h"

enum DEC_ID { BP_DEC, SP_DEC, ASP_DEC, MS_DEC, IMS_DEC, IASP_DEC, FHT_DEC, TASP_DEC, LMS_DEC, LCHE_DEC };
extern char const* const DEC_FULL_NAME[];

// compile-time decoder parameters of the reference (decoders.h:5-48)
#define DEC_DECISION 0
#define MS_ALPHA 0.8
#define MS_BETA  0.4
#define MS_THR   1.4
#define MS_QBITS 6
#define MS_DBITS 8

typedef struct {
    int q_bits;
    int q;
    int nh;             // block columns
    int rh;             // block rows
    int m;              // lifting size
    int n;              // codeword length nh * m
    int maxiter;
    int codec_id;
    short** hd;         // rh x nh shifts, -1 = no circulant; filled by the caller before decod_init
    double* y;          // n channel LLRs (caller's I/O buffer)
    double* decword;    // n decisions (caller's I/O buffer)
    short* syndr;       // rh * m
    // engine side
    ldpcb200_handle engine;     // created by decod_init (or lazily by the first decode)
    int engine_precision;       // 64 (default: the reference's arithmetic) or 32, see decod_set_precision
    double engine_alpha, engine_thr;
    int engine_qbits, engine_dbits;
    int bp_chain;               // BP_DEC: carry the syndrome from frame to frame (decoders.cpp:1742-1759); default 1
} DEC_STATE;

DEC_STATE* decod_open(int decoder_id, int q_bits, int mh, int nh, int M);
int decod_init(void* st);                           // 1 = ok, 0 = failure (decoders.cpp:1009-1014: a null state is "ok")
void decod_close(DEC_STATE* st);

int bp_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision);
int sum_prod_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision);
int sum_prod_gf2_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision);
int min_sum_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision, double alpha);
int imin_sum_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision, double alpha, double thr, int qbits, int dbits);
int isum_prod_gf2_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision);
int tdmp_sum_prod_gf2_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision);
int lmin_sum_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision, double alpha, double beta);
int lche_decod(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision);

// ---- additions (not in the reference)
// 32 selects the fp32 throughput kernels for LMS_DEC / MS_DEC; call before decod_init.
void decod_set_precision(DEC_STATE* st, int precision);
// n_frames frames back to back in soft[] -> decword[] (0/1 doubles) and iters[]; same return value
// conventions per frame.  Returns 0, or a negative ldpcb200 error code.
int decod_batch(DEC_STATE* st, const double* soft, int n_frames, int maxiter, double* decword, int* iters);
