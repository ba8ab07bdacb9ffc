#include "decoders.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

char const* const DEC_FULL_NAME[] = {      // decoders.cpp:18-30
    "Belief Propagation", "Sum-Product", "Advanced Sum-Product", "Min-Sum", "Integer Min-Sum",
    "Integer Advanced Sum-Product", "FHT Sum-Product", "TDMP Advanced Sum-Product", "Layered Min-Sum",
    "Low complexity-high efficiency"
};

DEC_STATE* decod_open(int decoder_id, int q_bits, int mh, int nh, int M)
{
    switch (decoder_id) {
    case BP_DEC: case SP_DEC: case ASP_DEC: case MS_DEC: case IMS_DEC: case IASP_DEC: case TASP_DEC: case LMS_DEC: case LCHE_DEC: break;
    default: return NULL;                   // FHT_DEC (GF(q)) is out of scope; unknown ids are NULL in the reference too
    }
    if (mh <= 0 || nh <= 0 || M <= 0) return NULL;
    DEC_STATE* st = (DEC_STATE*)calloc(1, sizeof(DEC_STATE));
    if (!st) return NULL;
    st->q_bits = q_bits; st->q = 1 << q_bits; st->nh = nh; st->rh = mh; st->m = M; st->n = nh * M;
    st->codec_id = decoder_id;
    st->hd = (short**)calloc(mh, sizeof(short*));
    short* flat = (short*)calloc((size_t)mh * nh, sizeof(short));
    st->y = (double*)calloc(st->n, sizeof(double));
    st->decword = (double*)calloc(st->n, sizeof(double));
    st->syndr = (short*)calloc((size_t)mh * M, sizeof(short));
    if (!st->hd || !flat || !st->y || !st->decword || !st->syndr) {
        free(flat); free(st->hd); free(st->y); free(st->decword); free(st->syndr); free(st);
        return NULL;
    }
    for (int j = 0; j < mh; j++) st->hd[j] = flat + (size_t)j * nh;
    st->engine = NULL;
    st->engine_precision = 64;
    st->engine_alpha = MS_ALPHA; st->engine_thr = MS_THR; st->engine_qbits = MS_QBITS; st->engine_dbits = MS_DBITS;
    st->bp_chain = 1;
    return st;
}

void decod_set_precision(DEC_STATE* st, int precision)
{
    if (st) st->engine_precision = precision;
}

static int make_engine(DEC_STATE* st)
{
    if (st->engine) { ldpcb200_destroy(st->engine); st->engine = NULL; }
    ldpcb200_params p;
    ldpcb200_default_params(&p);
    p.alpha = st->engine_alpha; p.thr = st->engine_thr; p.qbits = st->engine_qbits; p.dbits = st->engine_dbits;
    bool has32 = st->codec_id == LMS_DEC || st->codec_id == MS_DEC;
    p.precision = has32 ? st->engine_precision : 64;
    int rc = ldpcb200_create(st->hd[0], st->rh, st->nh, st->m, st->codec_id, &p, &st->engine);
    if (rc) { fprintf(stderr, "decod_init: %s\n", ldpcb200_last_error()); st->engine = NULL; return 0; }
    return 1;
}

int decod_init(void* state)
{
    DEC_STATE* st = (DEC_STATE*)state;
    if (!st) return 1;                      // decoders.cpp:1014
    return make_engine(st);
}

void decod_close(DEC_STATE* st)
{
    if (!st) return;
    if (st->engine) ldpcb200_destroy(st->engine);
    if (st->hd) { free(st->hd[0]); free(st->hd); }
    free(st->y); free(st->decword); free(st->syndr);
    free(st);
}

// the decoder parameters arrive per call in the reference; the engine fixes them per handle
static int ensure_engine(DEC_STATE* st, double alpha, double thr, int qbits, int dbits)
{
    bool changed = st->engine_alpha != alpha || st->engine_thr != thr || st->engine_qbits != qbits || st->engine_dbits != dbits;
    st->engine_alpha = alpha; st->engine_thr = thr; st->engine_qbits = qbits; st->engine_dbits = dbits;
    if (st->engine && !changed) return 1;
    return make_engine(st);
}

// What the reference leaves in its caller-visible arrays, decoder by decoder:
//   decword, decision == 0: 0 / 1; decision != 0: the soft output -- BP_DEC / SP_DEC soft[] itself (:1767, :1907, :1991, :2173),
//     ASP_DEC the probabilities soft_out (make_output :2308, :2396, :2570), IASP_DEC soft_out / 2^16 (imake_output :3806),
//     MS_DEC / LMS_DEC the posterior LLRs (:4670, :5413), IMS_DEC the integer posteriors (:5579-5589); TASP_DEC and LCHE_DEC
//     ignore `decision` (:2736, lche_decod);
//   soft[] (the input): BP_DEC / SP_DEC overwrite it with their posteriors; ASP_DEC / TASP_DEC replace it by the channel
//     probabilities e^-y / (e^y + e^-y), y = clip(soft / 2, +-20) (:2351-2358, :2611-2618); IASP_DEC by 1 / (1 + e^clip(soft, +-20))
//     (:3849-3854); the min-sum family and LCHE_DEC leave it alone.
static int decode_one(DEC_STATE* st, int id, double soft[], double decword[], int maxiter, int decision,
                      double alpha = MS_ALPHA, double thr = MS_THR, int qbits = MS_QBITS, int dbits = MS_DBITS)
{
    if (!st || st->codec_id != id) { fprintf(stderr, "decoder called on a state opened for decoder %d\n", st ? st->codec_id : -1); return -100000; }
    if (!ensure_engine(st, alpha, thr, qbits, dbits)) return -100000;
    const int n = st->n;
    std::vector<unsigned char> hard(n);
    const bool soft_out = decision != 0 && id != TASP_DEC && id != LCHE_DEC;
    const bool want_post = soft_out || id == BP_DEC || id == SP_DEC;
    // the posterior comes in the engine's own type for this decoder (ldpcb200.h): int16 / uint16 for the fixed-point pair,
    // float when the handle runs the fp32 kernels, else double
    const bool f32 = (id == LMS_DEC || id == MS_DEC) && st->engine_precision == 32;
    const int ptype = id == IMS_DEC ? LDPCB200_I16 : id == IASP_DEC ? LDPCB200_U16 : f32 ? LDPCB200_F32 : LDPCB200_F64;
    std::vector<double> post;                                   // 8 bytes per value: large enough (and aligned) for every type
    if (want_post) post.resize(n);
    int32_t iters = 0;
    uint32_t flags = (id == BP_DEC && st->bp_chain) ? LDPCB200_BP_CHAIN_SYNDROME : 0;
    int rc = ldpcb200_decode_batch(st->engine, soft, LDPCB200_F64, 1, maxiter, flags, hard.data(), &iters,
                                   want_post ? (void*)post.data() : NULL, ptype, NULL);
    if (rc) { fprintf(stderr, "decode: %s\n", ldpcb200_last_error()); return -100000; }
    if (want_post && ptype != LDPCB200_F64) {                   // widen in place, back to front
        const void* raw = post.data();
        for (int i = n - 1; i >= 0; i--)
            post[i] = ptype == LDPCB200_I16 ? (double)((const int16_t*)raw)[i]
                    : ptype == LDPCB200_U16 ? (double)((const uint16_t*)raw)[i] / 65536.0
                    : (double)((const float*)raw)[i];
    }
    if (soft_out) for (int i = 0; i < n; i++) decword[i] = post[i];
    else for (int i = 0; i < n; i++) decword[i] = hard[i];
    if (id == BP_DEC || id == SP_DEC) memcpy(soft, post.data(), sizeof(double) * n);
    else if (id == ASP_DEC || id == TASP_DEC) {
        for (int i = 0; i < n; i++) {
            const double x = soft[i] * 0.5, y = x > 20.0 ? 20.0 : x < -20.0 ? -20.0 : x;
            const double e0 = exp(y), e1 = exp(-y);
            soft[i] = e1 / (e0 + e1);
        }
    } else if (id == IASP_DEC) {
        for (int i = 0; i < n; i++) {
            const double x = soft[i], y = x > 20.0 ? 20.0 : x < -20.0 ? -20.0 : x;
            soft[i] = 1.0 / (1.0 + exp(y));
        }
    }
    st->maxiter = maxiter;
    return iters;
}

int bp_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision) { return decode_one(st, BP_DEC, soft, decword, maxiter, decision); }
int sum_prod_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision) { return decode_one(st, SP_DEC, soft, decword, maxiter, decision); }
int sum_prod_gf2_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision) { return decode_one(st, ASP_DEC, soft, decword, maxiter, decision); }
int min_sum_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision, double alpha) { return decode_one(st, MS_DEC, soft, decword, maxiter, decision, alpha); }
int imin_sum_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision, double alpha, double thr, int qbits, int dbits)
{
    return decode_one(st, IMS_DEC, soft, decword, maxiter, decision, alpha, thr, qbits, dbits);
}
int isum_prod_gf2_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision) { return decode_one(st, IASP_DEC, soft, decword, maxiter, decision); }
int tdmp_sum_prod_gf2_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision) { return decode_one(st, TASP_DEC, soft, decword, maxiter, decision); }
// alpha and beta are ignored, exactly as in the reference (decoders.cpp:5163: beta = 0.4 is hard-coded)
int lmin_sum_decod_qc_lm(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision, double, double) { return decode_one(st, LMS_DEC, soft, decword, maxiter, decision); }
int lche_decod(DEC_STATE* st, double soft[], double decword[], int maxiter, int decision) { return decode_one(st, LCHE_DEC, soft, decword, maxiter, decision); }

int decod_batch(DEC_STATE* st, const double* soft, int n_frames, int maxiter, double* decword, int* iters)
{
    if (!st) return LDPCB200_EINVAL;
    if (!st->engine && !make_engine(st)) return LDPCB200_ECUDA;
    std::vector<unsigned char> hard((size_t)n_frames * st->n);
    uint32_t flags = (st->codec_id == BP_DEC && st->bp_chain) ? LDPCB200_BP_CHAIN_SYNDROME : 0;
    int rc = ldpcb200_decode_batch(st->engine, soft, LDPCB200_F64, n_frames, maxiter, flags, hard.data(), iters, NULL, LDPCB200_F64, NULL);
    if (rc) return rc;
    if (decword) for (size_t i = 0; i < hard.size(); i++) decword[i] = hard[i];
    return 0;
}
