/* TEST INFRASTRUCTURE ONLY -- included twice by ldpc_oracle.c with REAL = double / float.
 * Min-sum family of the oracle (add / subtract / compare / one multiply only), so the same text is
 * an exact specification at either precision.  See ldpc_oracle.h for the conventions. */

#ifndef REAL
#error "define REAL and SFX before including"
#endif

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define FN(name) CAT(name, SFX)

/* per-check-row compressed state, MS_DEC_STATE, decoders.h:115-121 */
typedef struct { REAL min1, min2; int pos, sign; } FN(msrow);

/* syndrome of the hard decisions (x < 0) of `soft`; check_syndrome, decoders.cpp:793-814 */
static int FN(syndrome_neg)(const qc_graph* g, const REAL* soft)
{
    int Z = g->Z, parity = 0;
    for (int j = 0; j < g->b; j++)
        for (int n = 0; n < Z; n++) {
            int s = 0;
            for (int e = g->rp[j]; e < g->rp[j + 1]; e++)
                s ^= soft[g->col[e] * Z + (n + g->sh[e]) % Z] < 0;
            parity |= s;
        }
    return parity;
}

/* lmin_sum_decod_qc_lm, decoders.cpp:5064-5425 (MY_VERSION branch :5140-5207) */
static int FN(lms_frame)(const qc_graph* g, const REAL* y, int maxiter, REAL* soft,
                         FN(msrow)* prev, uint8_t* esign, REAL* v2c)
{
    const int Z = g->Z, N = g->c * Z, R = g->b * Z;
    const REAL beta = (REAL)0.4;                 /* decoders.cpp:5163 -- the beta argument is ignored */
    const REAL MAXV = (REAL)32767;               /* MAX_VAL, decoders.cpp:4301 */
    int iter, parity;

    for (int i = 0; i < N; i++) soft[i] = y[i];                             /* :5088 */
    for (int i = 0; i < R; i++) { prev[i].min1 = 0; prev[i].min2 = 0; prev[i].pos = 0; prev[i].sign = 0; }
    memset(esign, 0, (size_t)g->E * Z);                                     /* :5108 */

    parity = FN(syndrome_neg)(g, soft);                                     /* :5111-5115 */

    for (iter = 0; iter < maxiter; iter++) {
        if (!parity) break;                                                 /* :5119 */
        for (int j = 0; j < g->b; j++) {
            for (int n = 0; n < Z; n++) {
                FN(msrow) cur = { MAXV, MAXV, 0, 0 };                       /* :5131-5137 */
                FN(msrow)* pr = &prev[j * Z + n];
                for (int e = g->rp[j]; e < g->rp[j + 1]; e++) {
                    int k = g->col[e];
                    int idx = k * Z + (n + g->sh[e]) % Z;
                    REAL pabs = pr->pos == k ? pr->min2 : pr->min1;          /* :5152 */
                    int psgn = esign[(size_t)e * Z + n] ^ pr->sign;          /* :5156 */
                    REAL pval = psgn ? -pabs : pabs;
                    REAL v = soft[idx] - pval;                               /* :5158 */
                    int s = v < 0;                                           /* :5164 */
                    REAL a = v < (REAL)0.0 ? -v : v;
                    a -= beta;                                               /* :5166 */
                    a = a < 0 ? 0 : a;                                       /* :5168 */
                    v2c[e - g->rp[j]] = v;                                   /* :5170 */
                    esign[(size_t)e * Z + n] = (uint8_t)s;                   /* :5171 */
                    cur.sign ^= s;                                           /* process_check_node :5012-5027 */
                    if (a < cur.min1) { cur.pos = k; cur.min2 = cur.min1; cur.min1 = a; }
                    else if (a < cur.min2) cur.min2 = a;
                }
                *pr = cur;                                                  /* :5179 */
                for (int e = g->rp[j]; e < g->rp[j + 1]; e++) {
                    int k = g->col[e];
                    int idx = k * Z + (n + g->sh[e]) % Z;
                    REAL cabs = cur.pos == k ? cur.min2 : cur.min1;          /* :5193 */
                    REAL cval = (esign[(size_t)e * Z + n] ^ cur.sign) ? -cabs : cabs;
                    soft[idx] = v2c[e - g->rp[j]] + cval;                    /* :5199-5204 */
                }
            }
        }
        parity = FN(syndrome_neg)(g, soft);                                 /* :5281-5284 */
        if (!parity) break;
    }
    return parity ? -iter : iter + 1;                                       /* :5424 */
}

int FN(orc_lms)(const int16_t* hd, int b, int c, int Z, const REAL* y, int nf, int maxiter,
                uint8_t* hard, int32_t* iters, REAL* post)
{
    qc_graph g;
    if (qc_build(&g, hd, b, c, Z)) return -1;
    int N = c * Z, R = b * Z;
    REAL* soft = (REAL*)malloc(sizeof(REAL) * N);
    FN(msrow)* prev = (FN(msrow)*)malloc(sizeof(FN(msrow)) * R);
    uint8_t* esign = (uint8_t*)malloc((size_t)g.E * Z);
    REAL* v2c = (REAL*)malloc(sizeof(REAL) * (g.maxdeg + 1));
    for (int f = 0; f < nf; f++) {
        iters[f] = FN(lms_frame)(&g, y + (size_t)f * N, maxiter, soft, prev, esign, v2c);
        if (hard) for (int i = 0; i < N; i++) hard[(size_t)f * N + i] = soft[i] < 0;   /* :5421 */
        if (post) memcpy(post + (size_t)f * N, soft, sizeof(REAL) * N);
    }
    free(soft); free(prev); free(esign); free(v2c);
    qc_free(&g);
    return 0;
}

/* min_sum_decod_qc_lm, decoders.cpp:4554-4767 */
static int FN(ms_frame)(const qc_graph* g, const REAL* y, int maxiter, REAL alpha, REAL* soft,
                        FN(msrow)* dcs, uint8_t* esign, uint8_t* hard)
{
    const int Z = g->Z, N = g->c * Z, R = g->b * Z;
    const REAL MAXV = (REAL)32767;
    int iter, parity = 0;

    for (int i = 0; i < R; i++) { dcs[i].min1 = 0; dcs[i].min2 = 0; dcs[i].pos = 0; dcs[i].sign = 0; }   /* :4579-4585 */
    memset(esign, 0, (size_t)g->E * Z);                                     /* :4596 */
    /* the pre-loop syndrome (:4601-4622) is never read: synd is cleared at :4632 before any use.
       With maxiter <= 0 the reference would return on that stale value; not restated. */

    for (iter = 0; iter < maxiter; iter++) {
        for (int i = 0; i < N; i++) soft[i] = 0;                            /* :4633 */
        /* STATE 1 (:4636-4667): sum of check-to-variable messages, block rows ascending */
        for (int j = 0; j < g->b; j++)
            for (int e = g->rp[j]; e < g->rp[j + 1]; e++) {
                int k = g->col[e];
                for (int n = 0; n < Z; n++) {
                    FN(msrow)* d = &dcs[j * Z + n];
                    REAL tmp = d->pos == k ? d->min2 : d->min1;              /* :4649 */
                    REAL val = (esign[(size_t)e * Z + n] ^ d->sign) ? -tmp : tmp;
                    int idx = k * Z + (n + g->sh[e]) % Z;
                    soft[idx] = soft[idx] + val;                             /* :4658 */
                }
            }
        /* STATE 2 (:4678-4685) */
        for (int i = 0; i < N; i++) {
            soft[i] = y[i] + soft[i] * alpha;
            hard[i] = soft[i] < 0;
        }
        /* STATE 3 (:4688-4755) */
        parity = 0;
        for (int j = 0; j < g->b; j++)
            for (int n = 0; n < Z; n++) {
                FN(msrow) t = { MAXV, MAXV, 0, 0 };
                FN(msrow)* d = &dcs[j * Z + n];
                int synd = 0;
                for (int e = g->rp[j]; e < g->rp[j + 1]; e++) {
                    int k = g->col[e];
                    REAL rs = soft[k * Z + (n + g->sh[e]) % Z];
                    synd ^= rs < 0;                                          /* :4711 */
                    REAL old = d->pos == k ? d->min2 : d->min1;              /* :4714 */
                    REAL val = old * alpha;                                  /* :4719 */
                    REAL tt = (esign[(size_t)e * Z + n] ^ d->sign) ? -val : val;
                    tt = rs - tt;                                            /* :4722 */
                    int s = tt < 0;
                    esign[(size_t)e * Z + n] = (uint8_t)s;
                    t.sign ^= s;
                    val = tt < (REAL)0.0 ? -tt : tt;                         /* :4729 */
                    val = (val > MAXV) ? MAXV : val;                         /* :4730 */
                    if (val < t.min1) { t.pos = k; t.min2 = t.min1; t.min1 = val; }
                    else if (val < t.min2) t.min2 = val;
                }
                *d = t;                                                     /* :4753 */
                parity |= synd;
            }
        if (!parity) break;                                                 /* :4761 */
    }
    return parity ? -iter : iter + 1;                                       /* :4766 */
}

int FN(orc_ms)(const int16_t* hd, int b, int c, int Z, const REAL* y, int nf, int maxiter,
               REAL alpha, uint8_t* hard, int32_t* iters, REAL* post)
{
    qc_graph g;
    if (qc_build(&g, hd, b, c, Z)) return -1;
    int N = c * Z, R = b * Z;
    REAL* soft = (REAL*)calloc(N, sizeof(REAL));
    FN(msrow)* dcs = (FN(msrow)*)malloc(sizeof(FN(msrow)) * R);
    uint8_t* esign = (uint8_t*)malloc((size_t)g.E * Z);
    uint8_t* hd1 = (uint8_t*)calloc(N, 1);
    for (int f = 0; f < nf; f++) {
        iters[f] = FN(ms_frame)(&g, y + (size_t)f * N, maxiter, alpha, soft, dcs, esign, hd1);
        if (hard) memcpy(hard + (size_t)f * N, hd1, N);
        if (post) memcpy(post + (size_t)f * N, soft, sizeof(REAL) * N);
    }
    free(soft); free(dcs); free(esign); free(hd1);
    qc_free(&g);
    return 0;
}

#undef FN
#undef CAT
#undef CAT2
