/* TEST INFRASTRUCTURE ONLY -- see ldpc_oracle.h.  Plain C99, libm only.
 * Build: make -C oracle oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, like the
 * reference's generic x86-64 -O3 build). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ldpc_oracle.h"

/* ------------------------------------------------------------------------------------------
 * QC Tanner graph as edge lists.  The reference never builds one: it scans hd[j][i]
 * (decoders.h:146) and skips -1.  Edge order inside a row = ascending block column, inside a
 * column = ascending block row -- the orders in which the reference accumulates.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int b, c, Z, E, maxdeg, maxcdeg;
    int *rp;      /* b+1 : row j owns edges rp[j] .. rp[j+1]-1                    */
    int *col;     /* E   : block column of edge e                                 */
    int *sh;      /* E   : circulant shift of edge e, reduced into [0, Z)         */
    int *row;     /* E   : block row of edge e                                    */
    int *cp;      /* c+1 : column i owns cedge[cp[i] .. cp[i+1]-1], rows ascending */
    int *cedge;   /* E                                                            */
} qc_graph;

static void qc_free(qc_graph* g)
{
    free(g->rp); free(g->col); free(g->sh); free(g->row); free(g->cp); free(g->cedge);
}

static int qc_build(qc_graph* g, const int16_t* hd, int b, int c, int Z)
{
    memset(g, 0, sizeof(*g));
    g->b = b; g->c = c; g->Z = Z;
    int E = 0;
    for (int i = 0; i < b * c; i++) E += hd[i] != -1;
    g->E = E;
    g->rp = (int*)calloc(b + 1, sizeof(int));
    g->cp = (int*)calloc(c + 1, sizeof(int));
    g->col = (int*)calloc(E + 1, sizeof(int));
    g->sh = (int*)calloc(E + 1, sizeof(int));
    g->row = (int*)calloc(E + 1, sizeof(int));
    g->cedge = (int*)calloc(E + 1, sizeof(int));
    int e = 0;
    for (int j = 0; j < b; j++) {
        g->rp[j] = e;
        for (int i = 0; i < c; i++) {
            int v = hd[j * c + i];
            if (v == -1) continue;
            g->col[e] = i;
            g->sh[e] = ((v % Z) + Z) % Z;     /* rotate() reduces any shift mod M, decoders.cpp:335-339 */
            g->row[e] = j;
            e++;
        }
        if (e - g->rp[j] > g->maxdeg) g->maxdeg = e - g->rp[j];
    }
    g->rp[b] = e;
    int k = 0;
    for (int i = 0; i < c; i++) {
        g->cp[i] = k;
        for (int j = 0; j < b; j++)
            for (int q = g->rp[j]; q < g->rp[j + 1]; q++)
                if (g->col[q] == i) g->cedge[k++] = q;
        if (k - g->cp[i] > g->maxcdeg) g->maxcdeg = k - g->cp[i];
    }
    g->cp[c] = k;
    return 0;
}

/* mind / maxd / absd exactly as decoders.cpp:104-109 (matters for NaN operands) */
static double mind(double a, double b) { if (a < b) return a; else return b; }
static double maxd(double a, double b) { if (a < b) return b; else return a; }
static double absd(double x) { if (x < 0) return -x; else return x; }

#define REAL double
#define SFX _f64
#include "ldpc_oracle_impl.h"
#undef REAL
#undef SFX

#define REAL float
#define SFX _f32
#include "ldpc_oracle_impl.h"
#undef REAL
#undef SFX

/* ------------------------------------------------------------------------------------------
 * IMS_DEC -- imin_sum_decod_qc_lm, decoders.cpp:5430-5690
 * ---------------------------------------------------------------------------------------- */
typedef struct { int16_t min1, min2; int pos, sign; } imsrow;    /* IMS_DEC_STATE, decoders.h:123-129 */

static inline int16_t limit_s16(int16_t x, int16_t mx)            /* limit_val, decoders.cpp:4308 */
{
    return x > mx ? mx : (x < -mx ? (int16_t)-mx : x);
}

static int ims_frame(const qc_graph* g, const double* y, int maxiter, double alpha, double thr,
                     int qbits, int dbits, int16_t* soft, int16_t* iy, imsrow* dcs,
                     uint8_t* esign, uint8_t* hard)
{
    const int Z = g->Z, N = g->c * Z, R = g->b * Z;
    const int16_t max_data = (int16_t)((1L << (dbits - 1)) - 1);     /* :5445 */
    const int16_t max_quant = (int16_t)((1L << (qbits - 1)) - 1);    /* :5446 */
    const int ialpha = (int)(alpha * (1L << 4));                     /* MS_ALPHA_FPP = 4, :5458 */
    int iter, parity = 0;

    for (int i = 0; i < R; i++) { dcs[i].min1 = 0; dcs[i].min2 = 0; dcs[i].pos = 0; dcs[i].sign = 0; }

    {   /* per-frame energy normalisation + quantiser, :5472-5500 */
        double en = 0;
        for (int i = 0; i < N; i++) en += y[i] * y[i];                /* sequential, double */
        double coef = sqrt(N / en);
        for (int i = 0; i < N; i++) {
            double val = y[i];
            int sign = 0;
            if (val < 0) { val = -val; sign = 1; }
            val *= coef;
            if (val > thr) val = thr;
            int ival = (int16_t)floor(val * max_quant / thr + 0.5);
            iy[i] = (int16_t)(sign ? -ival : ival);
        }
    }
    memset(esign, 0, (size_t)g->E * Z);

    for (iter = 0; iter < maxiter; iter++) {
        for (int i = 0; i < N; i++) soft[i] = 0;                      /* :5536 */
        /* STATE 1 (:5540-5576): saturate after EVERY add, block rows ascending */
        for (int j = 0; j < g->b; j++)
            for (int e = g->rp[j]; e < g->rp[j + 1]; e++) {
                int k = g->col[e];
                for (int n = 0; n < Z; n++) {
                    imsrow* d = &dcs[j * Z + n];
                    int16_t tmp = d->pos == k ? d->min2 : d->min1;
                    tmp = (int16_t)((tmp * ialpha) >> 4);             /* :5554 */
                    int16_t val = (int16_t)((esign[(size_t)e * Z + n] ^ d->sign) ? -tmp : tmp);
                    int idx = k * Z + (n + g->sh[e]) % Z;
                    int16_t t = (int16_t)(soft[idx] + val);           /* :5567 */
                    soft[idx] = limit_s16(t, max_data);               /* :5568 */
                }
            }
        /* STATE 2 (:5594-5603) */
        for (int i = 0; i < N; i++) {
            soft[i] = (int16_t)(iy[i] + soft[i]);
            soft[i] = limit_s16(soft[i], max_data);
            hard[i] = soft[i] < 0;
        }
        /* STATE 3 (:5608-5678) */
        parity = 0;
        for (int j = 0; j < g->b; j++)
            for (int n = 0; n < Z; n++) {
                imsrow t = { max_data, max_data, 0, 0 };
                imsrow* d = &dcs[j * Z + n];
                int synd = 0;
                for (int e = g->rp[j]; e < g->rp[j + 1]; e++) {
                    int k = g->col[e];
                    int16_t rs = soft[k * Z + (n + g->sh[e]) % Z];
                    synd ^= rs < 0;                                   /* :5631 */
                    int16_t old = d->pos == k ? d->min2 : d->min1;
                    int sign = esign[(size_t)e * Z + n] ^ d->sign;
                    int16_t val = (int16_t)((old * ialpha) >> 4);     /* :5640 */
                    int16_t tt = (int16_t)(sign ? -val : val);
                    int16_t v2c = (int16_t)(rs - tt);                 /* :5646 */
                    int s = v2c < 0;
                    esign[(size_t)e * Z + n] = (uint8_t)s;
                    t.sign ^= s;
                    val = (int16_t)(v2c < 0 ? -v2c : v2c);
                    val = (val > max_data) ? max_data : val;          /* :5653 */
                    if (val < t.min1) { t.pos = k; t.min2 = t.min1; t.min1 = val; }
                    else if (val < t.min2) t.min2 = val;
                }
                *d = t;
                parity |= synd;
            }
        if (!parity) break;
    }
    return parity ? -iter : iter + 1;
}

int orc_ims(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
            double alpha, double thr, int qbits, int dbits,
            uint8_t* hard, int32_t* iters, int16_t* post, int16_t* iyout)
{
    qc_graph g;
    if (qc_build(&g, hd, b, c, Z)) return -1;
    int N = c * Z, R = b * Z;
    int16_t* soft = (int16_t*)calloc(N, sizeof(int16_t));
    int16_t* iy = (int16_t*)calloc(N, sizeof(int16_t));
    imsrow* dcs = (imsrow*)malloc(sizeof(imsrow) * R);
    uint8_t* esign = (uint8_t*)malloc((size_t)g.E * Z);
    uint8_t* h1 = (uint8_t*)calloc(N, 1);
    for (int f = 0; f < nf; f++) {
        iters[f] = ims_frame(&g, y + (size_t)f * N, maxiter, alpha, thr, qbits, dbits, soft, iy, dcs, esign, h1);
        if (hard) memcpy(hard + (size_t)f * N, h1, N);
        if (post) memcpy(post + (size_t)f * N, soft, sizeof(int16_t) * N);
        if (iyout) memcpy(iyout + (size_t)f * N, iy, sizeof(int16_t) * N);
    }
    free(soft); free(iy); free(dcs); free(esign); free(h1);
    qc_free(&g);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * probability-domain helpers shared by TASP_DEC and ASP_DEC
 * ---------------------------------------------------------------------------------------- */

/* LLR -> P(bit = 1), decoders.cpp:2611-2618 (= :2351-2358) */
static double llr_to_p1(double llr)
{
    double x = llr * 0.5;
    double yv = maxd(mind(x, 20.0), -20.0);       /* INPUT_LIMIT, decoders.cpp:94 */
    double e0 = exp(yv);
    double e1 = exp(-yv);
    return e1 / (e0 + e1);
}

/* map_bin, decoders.cpp:2191-2228, contiguous (step = 1) on rw >= 2 values */
static void map_bin(double* s, int rw, double* SF, double* SB, double* P)
{
    for (int i = 0; i < rw; i++) P[i] = 1 - 2 * s[i];
    SF[0] = P[0];
    for (int i = 1; i < rw - 1; i++) SF[i] = P[i] * SF[i - 1];
    SB[rw - 1] = P[rw - 1];
    for (int i = rw - 2; i > 0; i--) SB[i] = P[i] * SB[i + 1];
    s[0] = (1 - SB[1]) / 2;
    for (int i = 1; i < rw - 1; i++) {
        double Zv = SF[i - 1] * SB[i + 1];
        s[i] = (1 - Zv) / 2;
    }
    s[rw - 1] = (1 - SF[rw - 2]) / 2;
}

/* check_syndrome_thr, decoders.cpp:2274-2306 */
static int syndrome_gt(const qc_graph* g, const double* soft, double thr)
{
    int Z = g->Z, parity = 0;
    for (int j = 0; j < g->b; j++)
        for (int n = 0; n < Z; n++) {
            int s = 0;
            for (int e = g->rp[j]; e < g->rp[j + 1]; e++)
                s ^= soft[g->col[e] * Z + (n + g->sh[e]) % Z] > thr;
            parity |= s;
        }
    return parity;
}

static int min_row_weight(const qc_graph* g)
{
    int m = 1 << 30;
    for (int j = 0; j < g->b; j++)
        if (g->rp[j + 1] - g->rp[j] < m) m = g->rp[j + 1] - g->rp[j];
    return m;
}

/* ------------------------------------------------------------------------------------------
 * TASP_DEC -- tdmp_sum_prod_gf2_decod_qc_lm, decoders.cpp:2584-2744
 * lam[e*Z + n] replaces the reference's dense Z[row][cnt]
 * ---------------------------------------------------------------------------------------- */
static int tasp_frame(const qc_graph* g, const double* y, int maxiter, double* gam, double* lam,
                      double* rho, double* a, double* w)
{
    const int Z = g->Z, N = g->c * Z;
    const double T = 0.0001, TT = 0;                                  /* :2597-2598 */
    int synd, steps;

    for (int i = 0; i < N; i++) gam[i] = llr_to_p1(y[i]);
    for (size_t i = 0; i < (size_t)g->E * Z; i++) lam[i] = 0.5;       /* :2620-2641 */

    synd = syndrome_gt(g, gam, 0.5);                                  /* :2653 */
    if (synd == 0) return 0;                                          /* :2654-2660 */

    steps = 0;
    while (steps < maxiter) {
        for (int j = 0; j < g->b; j++) {
            int e0 = g->rp[j], cnt = g->rp[j + 1] - e0;
            for (int n = 0; n < Z; n++) {
                for (int q = 0; q < cnt; q++) {
                    int e = e0 + q;
                    double x = gam[g->col[e] * Z + (n + g->sh[e]) % Z];
                    double av = lam[(size_t)e * Z + n];
                    rho[q] = x * (1.0 - av) / (av + x - 2.0 * av * x);   /* :2686 */
                }
                for (int q = 0; q < cnt; q++) {                       /* :2692-2697 */
                    if (rho[q] < TT) rho[q] = TT;
                    if (rho[q] > 1 - TT) rho[q] = 1 - TT;
                    a[q] = rho[q];
                }
                map_bin(a, cnt, w, w + cnt + 1, w + 2 * (cnt + 1));   /* :2699 */
                for (int q = 0; q < cnt; q++) {                       /* :2701-2705 */
                    if (a[q] < T) a[q] = T;
                    if (a[q] > 1.0 - T) a[q] = 1.0 - T;
                }
                for (int q = 0; q < cnt; q++) {
                    int e = e0 + q;
                    gam[g->col[e] * Z + (n + g->sh[e]) % Z] =
                        rho[q] * a[q] / (1.0 - rho[q] - a[q] + 2 * rho[q] * a[q]);   /* :2716 */
                    lam[(size_t)e * Z + n] = a[q];
                }
            }
        }
        /* the reference recomputes the syndrome after every layer (:2723) but only the value
           after the last layer survives to :2733 */
        synd = syndrome_gt(g, gam, 0.5);
        steps = steps + 1;
        if (synd == 0) break;
    }
    return synd ? -steps : steps;                                     /* :2740-2743 */
}

int orc_tasp(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
             uint8_t* hard, int32_t* iters, double* post)
{
    qc_graph g;
    if (qc_build(&g, hd, b, c, Z)) return -1;
    if (min_row_weight(&g) < 2) { qc_free(&g); return -2; }            /* map_bin reads SB[1] uninitialised, :2219 */
    int N = c * Z;
    double* gam = (double*)malloc(sizeof(double) * N);
    double* lam = (double*)malloc(sizeof(double) * (size_t)g.E * Z);
    double* rho = (double*)malloc(sizeof(double) * (g.maxdeg + 1));
    double* a = (double*)malloc(sizeof(double) * (g.maxdeg + 1));
    double* w = (double*)malloc(sizeof(double) * 3 * (g.maxdeg + 2));
    for (int f = 0; f < nf; f++) {
        iters[f] = tasp_frame(&g, y + (size_t)f * N, maxiter, gam, lam, rho, a, w);
        if (hard) for (int i = 0; i < N; i++) hard[(size_t)f * N + i] = gam[i] > 0.5;   /* :2738 */
        if (post) memcpy(post + (size_t)f * N, gam, sizeof(double) * N);
    }
    free(gam); free(lam); free(rho); free(a); free(w);
    qc_free(&g);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * ASP_DEC -- sum_prod_gf2_decod_qc_lm, decoders.cpp:2324-2581
 * msg[e*Z + n] replaces state[cnt][row] (messages live at the check-row lane n)
 * ---------------------------------------------------------------------------------------- */
static int asp_frame(const qc_graph* g, const double* y, int maxiter, int all_cw_2,
                     double* p, double* so, double* msg, double* a, double* w)
{
    const int Z = g->Z, N = g->c * Z;
    int synd, steps;

    for (int i = 0; i < N; i++) p[i] = llr_to_p1(y[i]);               /* :2351-2358 */
    for (int e = 0; e < g->E; e++)                                    /* :2361-2378 */
        for (int n = 0; n < Z; n++)
            msg[(size_t)e * Z + n] = p[g->col[e] * Z + (n + g->sh[e]) % Z];
    for (int i = 0; i < N; i++) so[i] = p[i];

    synd = syndrome_gt(g, so, 0.5);                                   /* :2392 */
    if (synd == 0) return 0;

    steps = 0;
    while (steps < maxiter) {
        /* check nodes, :2406-2428 */
        for (int j = 0; j < g->b; j++) {
            int e0 = g->rp[j], cnt = g->rp[j + 1] - e0;
            for (int n = 0; n < Z; n++) {
                for (int q = 0; q < cnt; q++) a[q] = msg[(size_t)(e0 + q) * Z + n];
                map_bin(a, cnt, w, w + cnt + 1, w + 2 * (cnt + 1));
                for (int q = 0; q < cnt; q++) msg[(size_t)(e0 + q) * Z + n] = a[q];
            }
        }
        if (all_cw_2) {
            /* every block column has weight 2, :2432-2482 */
            for (int i = 0; i < g->c; i++) {
                int ea = g->cedge[g->cp[i]], eb = g->cedge[g->cp[i] + 1];
                for (int k = 0; k < Z; k++) {
                    int na = (k - g->sh[ea] + Z) % Z, nb = (k - g->sh[eb] + Z) % Z;
                    double d0 = msg[(size_t)ea * Z + na], d1 = msg[(size_t)eb * Z + nb];
                    double p1 = p[i * Z + k];
                    double q10 = p1, q11 = p1;
                    double q00 = 1.0 - p1, q01 = 1.0 - p1, p0;
                    q10 = q10 * d1;
                    q00 = q00 * (1 - d1);
                    q11 = q11 * d0;
                    q01 = q01 * (1 - d0);
                    p1 = q10 * d0;
                    p0 = q00 * (1 - d0);
                    so[i * Z + k] = p1 / (p0 + p1);
                    msg[(size_t)ea * Z + na] = q10 / (q10 + q00);
                    msg[(size_t)eb * Z + nb] = q11 / (q11 + q01);
                }
            }
        } else {
            /* overall products, :2489-2522: block rows ascending */
            for (int i = 0; i < g->c; i++)
                for (int k = 0; k < Z; k++) {
                    double P1 = p[i * Z + k];
                    double P0 = 1 - p[i * Z + k];
                    for (int q = g->cp[i]; q < g->cp[i + 1]; q++) {
                        int e = g->cedge[q];
                        double d = msg[(size_t)e * Z + (k - g->sh[e] + Z) % Z];
                        P1 *= d;
                        P0 *= 1 - d;
                    }
                    so[i * Z + k] = P1 / (P0 + P1);
                }
            /* local data updating, :2525-2558 */
            for (int e = 0; e < g->E; e++) {
                int i = g->col[e];
                for (int k = 0; k < Z; k++) {
                    size_t m = (size_t)e * Z + (k - g->sh[e] + Z) % Z;
                    double s1 = so[i * Z + k];
                    double sos = msg[m];
                    double p1 = s1 / sos;
                    double p0 = (1 - s1) / (1 - sos);
                    double d = p1 / (p1 + p0);
                    msg[m] = maxd(mind(d, 1.0 - 0.000001), 0.000001);   /* SP_DEC_MIN/MAX_VAL, :96-97 */
                }
            }
        }
        synd = syndrome_gt(g, so, 0.5);                               /* :2566 */
        if (synd == 0) return steps + 1;
        steps = steps + 1;
    }
    return -steps;
}

int orc_asp(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
            uint8_t* hard, int32_t* iters, double* post)
{
    qc_graph g;
    if (qc_build(&g, hd, b, c, Z)) return -1;
    if (min_row_weight(&g) < 2) { qc_free(&g); return -2; }
    int N = c * Z;
    int all_cw_2 = 1;                                                 /* decod_init, decoders.cpp:1026-1041 */
    for (int i = 0; i < c; i++) if (g.cp[i + 1] - g.cp[i] != 2) { all_cw_2 = 0; break; }
    double* p = (double*)malloc(sizeof(double) * N);
    double* so = (double*)malloc(sizeof(double) * N);
    double* msg = (double*)malloc(sizeof(double) * (size_t)g.E * Z);
    double* a = (double*)malloc(sizeof(double) * (g.maxdeg + 1));
    double* w = (double*)malloc(sizeof(double) * 3 * (g.maxdeg + 2));
    for (int f = 0; f < nf; f++) {
        iters[f] = asp_frame(&g, y + (size_t)f * N, maxiter, all_cw_2, p, so, msg, a, w);
        if (hard) for (int i = 0; i < N; i++) hard[(size_t)f * N + i] = so[i] > 0.5;   /* make_output :2308 */
        if (post) memcpy(post + (size_t)f * N, so, sizeof(double) * N);
    }
    free(p); free(so); free(msg); free(a); free(w);
    qc_free(&g);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * BP_DEC -- bp_decod_qc_lm, decoders.cpp:1708-1920
 * ZZ/BB are stored per edge at the VARIABLE position k of the block column (as the reference
 * does); the check row lane of (edge, k) is n = (k - shift) mod Z.
 * ---------------------------------------------------------------------------------------- */
static int bp_frame(const qc_graph* g, const double* y, int maxiter, double* soft, double* yd,
                    double* ZZ, uint8_t* BB, double* s, uint8_t* bs, uint8_t* syndr)
{
    const int Z = g->Z, N = g->c * Z, R = g->b * Z;
    int synd, iter = 0;

    for (size_t i = 0; i < (size_t)g->E * Z; i++) ZZ[i] = 0.0;        /* :1732-1734 */
    for (int i = 0; i < N; i++) yd[i] = soft[i] = maxd(mind(y[i], 20.0), -20.0);   /* :1738 */

    /* pre-iteration check XORs into the syndrome left by the previous frame, :1742-1759 */
    for (int j = 0; j < g->b; j++)
        for (int e = g->rp[j]; e < g->rp[j + 1]; e++)
            for (int n = 0; n < Z; n++)
                syndr[j * Z + n] ^= soft[g->col[e] * Z + (n + g->sh[e]) % Z] < 0;
    synd = 0;
    for (int i = 0; i < R; i++) synd |= syndr[i];
    if (!synd) return 0;                                              /* :1765-1779 */

    while (iter < maxiter) {
        memset(syndr, 0, R);
        memset(bs, 0, R);
        for (int i = 0; i < R; i++) s[i] = 0;
        /* variable-node step + check sums, block columns outer, :1790-1827 */
        for (int i = 0; i < g->c; i++)
            for (int q = g->cp[i]; q < g->cp[i + 1]; q++) {
                int e = g->cedge[q], j = g->row[e];
                for (int k = 0; k < Z; k++) {
                    double A = exp(soft[i * Z + k] - ZZ[(size_t)e * Z + k]);
                    double x = log(absd((A - 1) / (A + 1)));
                    BB[(size_t)e * Z + k] = A < 1;
                    ZZ[(size_t)e * Z + k] = x;
                }
                for (int n = 0; n < Z; n++) {
                    int k = (n + g->sh[e]) % Z;
                    s[j * Z + n] += ZZ[(size_t)e * Z + k];
                    bs[j * Z + n] ^= BB[(size_t)e * Z + k];
                }
            }
        memcpy(soft, yd, sizeof(double) * N);                         /* :1832 */
        /* check-node step + posterior, :1834-1862 */
        for (int i = 0; i < g->c; i++)
            for (int q = g->cp[i]; q < g->cp[i + 1]; q++) {
                int e = g->cedge[q], j = g->row[e];
                for (int k = 0; k < Z; k++) {
                    int n = (k - g->sh[e] + Z) % Z;
                    double A = exp(s[j * Z + n] - ZZ[(size_t)e * Z + k]);
                    int bb = bs[j * Z + n] ^ BB[(size_t)e * Z + k];
                    A = (1 - 2 * bb) * log((1 + A) / (1 - A));
                    ZZ[(size_t)e * Z + k] = maxd(mind(A, 19.07), -19.07);
                }
                for (int k = 0; k < Z; k++) soft[i * Z + k] += ZZ[(size_t)e * Z + k];
            }
        for (int j = 0; j < g->b; j++)                                /* :1865-1882 */
            for (int e = g->rp[j]; e < g->rp[j + 1]; e++)
                for (int n = 0; n < Z; n++)
                    syndr[j * Z + n] ^= soft[g->col[e] * Z + (n + g->sh[e]) % Z] < 0;
        synd = 0;
        for (int i = 0; i < R; i++) synd |= syndr[i];
        if (!synd) { iter++; return iter; }
        iter++;
    }
    return -iter;
}

int orc_bp(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
           int chain, uint8_t* hard, int32_t* iters, double* post)
{
    qc_graph g;
    if (qc_build(&g, hd, b, c, Z)) return -1;
    int N = c * Z, R = b * Z;
    double* soft = (double*)malloc(sizeof(double) * N);
    double* yd = (double*)malloc(sizeof(double) * N);
    double* ZZ = (double*)malloc(sizeof(double) * (size_t)g.E * Z);
    uint8_t* BB = (uint8_t*)malloc((size_t)g.E * Z);
    double* s = (double*)malloc(sizeof(double) * R);
    uint8_t* bs = (uint8_t*)malloc(R);
    uint8_t* syndr = (uint8_t*)calloc(R, 1);                          /* calloc'ed in decod_open :429 */
    for (int f = 0; f < nf; f++) {
        if (!chain) memset(syndr, 0, R);
        iters[f] = bp_frame(&g, y + (size_t)f * N, maxiter, soft, yd, ZZ, BB, s, bs, syndr);
        if (hard) for (int i = 0; i < N; i++) hard[(size_t)f * N + i] = soft[i] < 0;
        if (post) memcpy(post + (size_t)f * N, soft, sizeof(double) * N);
    }
    free(soft); free(yd); free(ZZ); free(BB); free(s); free(bs); free(syndr);
    qc_free(&g);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * SP_DEC -- sum_prod_decod_qc_lm, decoders.cpp:1923-2185 (likelihood-ratio domain)
 * ---------------------------------------------------------------------------------------- */
static int syndrome_lt(const qc_graph* g, const double* soft, double thr)
{
    int Z = g->Z, parity = 0;
    for (int j = 0; j < g->b; j++)
        for (int n = 0; n < Z; n++) {
            int s = 0;
            for (int e = g->rp[j]; e < g->rp[j + 1]; e++)
                s ^= soft[g->col[e] * Z + (n + g->sh[e]) % Z] < thr;
            parity |= s;
        }
    return parity;
}

static int sp_frame(const qc_graph* g, const double* y, int maxiter, double* soft, double* yd,
                    double* ZZ, double* ZZ0, double* s)
{
    const int Z = g->Z, N = g->c * Z, R = g->b * Z;
    const double SP_THR = 1.0;                                        /* :1922 */
    int synd, iter = 0;

    for (int i = 0; i < N; i++) {
        double v = maxd(mind(y[i], 20.0), -20.0);
        yd[i] = soft[i] = exp(v);                                     /* :1947-1951 */
    }
    for (size_t i = 0; i < (size_t)g->E * Z; i++) ZZ[i] = 1.0;        /* :1957-1959 */

    synd = syndrome_lt(g, soft, SP_THR);                              /* :1964-1987 */
    if (!synd) return 0;

    while (iter < maxiter) {
        for (int i = 0; i < R; i++) s[i] = 1.0;
        for (int i = 0; i < N; i++) soft[i] = yd[i];
        for (int i = 0; i < g->c; i++) {                              /* :2013-2061 */
            for (int q = g->cp[i]; q < g->cp[i + 1]; q++) {
                int e = g->cedge[q], j = g->row[e];
                for (int k = 0; k < Z; k++) {
                    double AA = yd[i * Z + k];
                    for (int q2 = g->cp[i]; q2 < g->cp[i + 1]; q2++) {
                        if (q2 == q) continue;
                        AA *= ZZ[(size_t)g->cedge[q2] * Z + k];       /* all of the column except row j */
                    }
                    ZZ0[(size_t)e * Z + k] = (AA - 1) / (AA + 1);
                }
                for (int n = 0; n < Z; n++)
                    s[j * Z + n] *= ZZ0[(size_t)e * Z + (n + g->sh[e]) % Z];
            }
            for (int q = g->cp[i]; q < g->cp[i + 1]; q++) {
                int e = g->cedge[q];
                memcpy(&ZZ[(size_t)e * Z], &ZZ0[(size_t)e * Z], sizeof(double) * Z);
            }
        }
        for (int i = 0; i < g->c; i++)                                /* :2103-2127 */
            for (int q = g->cp[i]; q < g->cp[i + 1]; q++) {
                int e = g->cedge[q], j = g->row[e];
                for (int k = 0; k < Z; k++) {
                    int n = (k - g->sh[e] + Z) % Z;
                    double A = s[j * Z + n] / ZZ[(size_t)e * Z + k];
                    A = (1 + A) / (1 - A);
                    A = maxd(mind(A, 1.9e+8), -5.2e-9);
                    ZZ[(size_t)e * Z + k] = A;
                    soft[i * Z + k] *= A;
                }
            }
        synd = syndrome_lt(g, soft, SP_THR);                          /* :2129-2149 */
        if (!synd) { iter++; return iter; }
        iter++;
    }
    return -iter;
}

int orc_sp(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
           uint8_t* hard, int32_t* iters, double* post)
{
    qc_graph g;
    if (qc_build(&g, hd, b, c, Z)) return -1;
    int N = c * Z, R = b * Z;
    double* soft = (double*)malloc(sizeof(double) * N);
    double* yd = (double*)malloc(sizeof(double) * N);
    double* ZZ = (double*)malloc(sizeof(double) * (size_t)g.E * Z);
    double* ZZ0 = (double*)malloc(sizeof(double) * (size_t)g.E * Z);
    double* s = (double*)malloc(sizeof(double) * R);
    for (int f = 0; f < nf; f++) {
        iters[f] = sp_frame(&g, y + (size_t)f * N, maxiter, soft, yd, ZZ, ZZ0, s);
        if (hard) for (int i = 0; i < N; i++) hard[(size_t)f * N + i] = soft[i] < 1.0;
        if (post) memcpy(post + (size_t)f * N, soft, sizeof(double) * N);
    }
    free(soft); free(yd); free(ZZ); free(ZZ0); free(s);
    qc_free(&g);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * LCHE_DEC -- lche_decod, decoders.cpp:2893-3010; logexp_int :2777-2836; map_bin_llr :2837-2890
 * ---------------------------------------------------------------------------------------- */
static const double LCHE_A[32] = {
    1.41e+00, 7.72e-01, 4.54e-01, 2.72e-01, 1.65e-01, 9.97e-02, 6.04e-02, 3.66e-02,
    2.22e-02, 1.35e-02, 8.17e-03, 4.96e-03, 3.01e-03, 1.82e-03, 1.11e-03, 6.71e-04,
    4.07e-04, 2.47e-04, 1.50e-04, 9.08e-05, 5.51e-05, 3.34e-05, 2.03e-05, 1.23e-05,
    7.45e-06, 4.52e-06, 2.74e-06, 1.66e-06, 1.01e-06, 6.12e-07, 3.71e-07, 2.25e-07 };
static const double LCHE_B[32] = {
    3.47, 2.77, 2.37, 2.08, 1.86, 1.69, 1.54, 1.41, 1.29, 1.19, 1.11, 1.03, 0.95, 0.89, 0.83, 0.77,
    0.72, 0.67, 0.63, 0.59, 0.55, 0.52, 0.48, 0.45, 0.43, 0.40, 0.37, 0.35, 0.33, 0.31, 0.29, 0.27 };
static const double LCHE_C[32] = {
    6.93, 6.24, 5.83, 5.55, 5.32, 5.14, 4.99, 4.85, 4.73, 4.63, 4.53, 4.45, 4.37, 4.29, 4.22, 4.16,
    4.10, 4.04, 3.99, 3.94, 3.89, 3.84, 3.80, 3.75, 3.71, 3.67, 3.64, 3.60, 3.56, 3.53, 3.50, 3.47 };

static double logexp_int(double x)
{
    if (x <= 0) x = 1.0 / 4096.0;
    if (x > 16.0) x = 16.0;
    if (x >= 2.0) return -LCHE_A[(int)(2 * x + 0.5) - 1];
    else if (x > 1.0 / 16.0) return -LCHE_B[(int)(16 * x + 0.5) - 1];
    else if (x > 1.0 / 512.0) return -LCHE_C[(int)(512 * x + 0.5) - 1];
    else {
        double s = 0;
        while (x < 1.0 / 512.0) { x *= 32; s -= 3.46; }
        return s - LCHE_C[(int)(512 * x + 0.5) - 1];
    }
}

static void map_bin_llr(double* yv, int n, double* alog)
{
    int synd = 0;
    double sum = 0;
    for (int i = 0; i < n; i++) synd ^= yv[i] < 0;
    for (int i = 0; i < n; i++) {
        double ay = yv[i] < 0.0 ? -yv[i] : yv[i];
        alog[i] = logexp_int(ay);
    }
    for (int i = 0; i < n; i++) sum += alog[i];
    for (int i = 0; i < n; i++) {
        int hardb = (yv[i] < 0) ^ synd;
        double A = alog[i] - sum;
        double av = logexp_int(A);
        yv[i] = (2 * hardb - 1) * av;
    }
}

static int lche_frame(const qc_graph* g, const double* y, int maxiter, double* so, double* st,
                      double* u, double* yv, double* w)
{
    const int Z = g->Z, N = g->c * Z;
    int synd, steps;
    for (size_t i = 0; i < (size_t)g->E * Z; i++) st[i] = 0.0;        /* :2913-2915 */
    for (int i = 0; i < N; i++) so[i] = y[i];
    synd = syndrome_neg_f64(g, so);                                   /* :2927-2931 */
    if (synd == 0) return 0;
    steps = 0;
    while (steps < maxiter) {
        for (int j = 0; j < g->b; j++) {
            int e0 = g->rp[j], cnt = g->rp[j + 1] - e0;
            for (int n = 0; n < Z; n++) {
                for (int q = 0; q < cnt; q++) {
                    int e = e0 + q;
                    u[q] = yv[q] = so[g->col[e] * Z + (n + g->sh[e]) % Z] - st[(size_t)e * Z + n];   /* :2962 */
                }
                map_bin_llr(u, cnt, w);
                for (int q = 0; q < cnt; q++) {
                    int e = e0 + q;
                    so[g->col[e] * Z + (n + g->sh[e]) % Z] = u[q] + yv[q];   /* :2979 */
                    st[(size_t)e * Z + n] = u[q];
                }
            }
        }
        steps = steps + 1;
        synd = syndrome_neg_f64(g, so);
        if (synd == 0) break;
    }
    return synd ? -steps : steps;                                     /* :3006-3009 */
}

int orc_lche(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
             uint8_t* hard, int32_t* iters, double* post)
{
    qc_graph g;
    if (qc_build(&g, hd, b, c, Z)) return -1;
    int N = c * Z;
    double* so = (double*)malloc(sizeof(double) * N);
    double* st = (double*)malloc(sizeof(double) * (size_t)g.E * Z);
    double* u = (double*)malloc(sizeof(double) * (g.maxdeg + 1));
    double* yv = (double*)malloc(sizeof(double) * (g.maxdeg + 1));
    double* w = (double*)malloc(sizeof(double) * (g.maxdeg + 1));
    for (int f = 0; f < nf; f++) {
        iters[f] = lche_frame(&g, y + (size_t)f * N, maxiter, so, st, u, yv, w);
        if (hard) for (int i = 0; i < N; i++) hard[(size_t)f * N + i] = so[i] < 0.0;
        if (post) memcpy(post + (size_t)f * N, so, sizeof(double) * N);
    }
    free(so); free(st); free(u); free(yv); free(w);
    qc_free(&g);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * IASP_DEC -- isum_prod_gf2_decod_qc_lm (IASP_FIXED_POINT build), decoders.cpp:3822-4121,
 * imap_bin :2235-2271, icheck_syndrome :3772-3804.  SOFT_FPP = 12, ONE_SOFT = 4096.
 * ---------------------------------------------------------------------------------------- */
#define ONE_SOFT 4096
#define MAX_SOFT 4095
#define DIVR(x, n) (((x) + (1 << ((n) - 1))) >> (n))                 /* div_power2r, decoders.cpp:80 */

static void imap_bin(uint16_t* s, int rw, int16_t* SF, int16_t* SB, int16_t* P)
{
    for (int i = 0; i < rw; i++) P[i] = (int16_t)(ONE_SOFT - 2 * s[i]);
    SF[0] = P[0];
    for (int i = 1; i < rw - 1; i++) SF[i] = (int16_t)DIVR((int)P[i] * SF[i - 1], 12);
    SB[rw - 1] = P[rw - 1];
    for (int i = rw - 2; i > 0; i--) SB[i] = (int16_t)DIVR((int)P[i] * SB[i + 1], 12);
    s[0] = (uint16_t)DIVR(ONE_SOFT - SB[1], 1);
    s[0] = s[0] < 1 ? 1 : s[0];
    for (int i = 1; i < rw - 1; i++) {
        int Zv = DIVR((int)SF[i - 1] * SB[i + 1], 12);
        s[i] = (uint16_t)DIVR(ONE_SOFT - Zv, 1);
        s[i] = s[i] < 1 ? 1 : s[i];
    }
    s[rw - 1] = (uint16_t)DIVR(ONE_SOFT - SF[rw - 2], 1);
    s[rw - 1] = s[rw - 1] < 1 ? 1 : s[rw - 1];
}

static int isyndrome(const qc_graph* g, const uint16_t* soft)
{
    int Z = g->Z, parity = 0;
    for (int j = 0; j < g->b; j++)
        for (int n = 0; n < Z; n++) {
            int s = 0;
            for (int e = g->rp[j]; e < g->rp[j + 1]; e++)
                s ^= soft[g->col[e] * Z + (n + g->sh[e]) % Z] >> 15;
            parity |= s;
        }
    return parity;
}

static int iasp_frame(const qc_graph* g, const double* y, int maxiter, int all_cw_2,
                      uint16_t* yq, uint16_t* so, uint16_t* msg, uint16_t* a, int16_t* w)
{
    const int Z = g->Z, N = g->c * Z;
    int synd, steps;

    for (int i = 0; i < N; i++) {                                     /* :3849-3861 */
        double v = maxd(mind(y[i], 20.0), -20.0);
        double pr = 1.0 / (1.0 + exp(v));
        int x = (int)(pr * ONE_SOFT + 0.5);
        x = MAX_SOFT < x ? MAX_SOFT : x;
        yq[i] = (uint16_t)(x < 1 ? 1 : x);
    }
    for (int i = 0; i < N; i++) so[i] = yq[i];
    for (int i = 0; i < N; i++) yq[i] = (uint16_t)(yq[i] << 4);       /* :3867 */
    for (int e = 0; e < g->E; e++)                                    /* :3869-3886 (12-bit values) */
        for (int n = 0; n < Z; n++)
            msg[(size_t)e * Z + n] = so[g->col[e] * Z + (n + g->sh[e]) % Z];
    for (int i = 0; i < N; i++) so[i] = (uint16_t)(so[i] << 4);       /* :3889 */

    synd = isyndrome(g, so);
    if (synd == 0) return 0;

    steps = 0;
    while (steps < maxiter) {
        for (int j = 0; j < g->b; j++) {                              /* :3906-3911 */
            int e0 = g->rp[j], cnt = g->rp[j + 1] - e0;
            for (int n = 0; n < Z; n++) {
                for (int q = 0; q < cnt; q++) a[q] = msg[(size_t)(e0 + q) * Z + n];
                imap_bin(a, cnt, w, w + cnt + 1, w + 2 * (cnt + 1));
                for (int q = 0; q < cnt; q++) msg[(size_t)(e0 + q) * Z + n] = a[q];
            }
        }
        if (all_cw_2) {                                               /* :3915-3977 */
            for (int i = 0; i < g->c; i++) {
                int ea = g->cedge[g->cp[i]], eb = g->cedge[g->cp[i] + 1];
                for (int k = 0; k < Z; k++) {
                    size_t ma = (size_t)ea * Z + (k - g->sh[ea] + Z) % Z;
                    size_t mb = (size_t)eb * Z + (k - g->sh[eb] + Z) % Z;
                    uint16_t ip1 = yq[i * Z + k];
                    uint16_t ip0 = (uint16_t)((ONE_SOFT << 4) - ip1);
                    uint16_t d1 = (uint16_t)(msg[mb] << 4);
                    uint16_t d0 = (uint16_t)(msg[ma] << 4);
                    uint16_t t1 = (uint16_t)((ONE_SOFT << 4) - d1);
                    uint16_t t0 = (uint16_t)((ONE_SOFT << 4) - d0);
                    uint16_t q10 = (uint16_t)DIVR((uint32_t)ip1 * d1, 16);
                    uint16_t q11 = (uint16_t)DIVR((uint32_t)ip1 * d0, 16);
                    uint16_t q00 = (uint16_t)DIVR((uint32_t)ip0 * t1, 16);
                    uint16_t q01 = (uint16_t)DIVR((uint32_t)ip0 * t0, 16);
                    uint16_t p1 = (uint16_t)DIVR((uint32_t)q10 * d0, 16);
                    uint16_t p0 = (uint16_t)DIVR((uint32_t)q00 * t0, 16);
                    p0 = (uint16_t)(p1 + p0);
                    p0 = p0 < 1 ? 1 : p0;
                    so[i * Z + k] = (uint16_t)(((uint32_t)p1 << 16) / p0);
                    so[i * Z + k] = so[i * Z + k] < (1 << 4) ? (1 << 4) : so[i * Z + k];
                    q00 = (uint16_t)(q00 + q10);
                    q01 = (uint16_t)(q01 + q11);
                    q00 = q00 < 1 ? 1 : q00;
                    q01 = q01 < 1 ? 1 : q01;
                    uint16_t n0 = (uint16_t)(((uint32_t)q10 << 12) / q00);
                    uint16_t n1 = (uint16_t)(((uint32_t)q11 << 12) / q01);
                    msg[ma] = n0 < 1 ? 1 : n0;
                    msg[mb] = n1 < 1 ? 1 : n1;
                }
            }
        } else {
            for (int i = 0; i < g->c; i++)                            /* :3984-4052 */
                for (int k = 0; k < Z; k++) {
                    uint32_t P1 = (uint32_t)yq[i * Z + k] << 16;
                    uint32_t P0 = (uint32_t)((ONE_SOFT << 4) - yq[i * Z + k]) << 16;
                    for (int q = g->cp[i]; q < g->cp[i + 1]; q++) {
                        int e = g->cedge[q];
                        uint16_t d = msg[(size_t)e * Z + (k - g->sh[e] + Z) % Z];
                        uint16_t d1 = (uint16_t)(d << 4);
                        uint16_t d0 = (uint16_t)((MAX_SOFT - d) << 4);        /* MAX_SOFT, not ONE_SOFT (:4010) */
                        uint64_t pp1 = (uint64_t)P1 * d1;
                        uint64_t pp0 = (uint64_t)P0 * d0;
                        P1 = (uint32_t)(pp1 >> 16);
                        P0 = (uint32_t)(pp0 >> 16);
                    }
                    int s;
                    uint32_t x = P1 >> 1;
                    uint32_t yv = (P0 >> 1) + x;
                    int flg = yv > (ONE_SOFT << 4);
                    if (flg) yv = yv >> 12; else x = x << 12;
                    yv = yv < 1 ? 1 : yv;
                    s = (int)(x / yv);
                    s = MAX_SOFT < s ? MAX_SOFT : s;
                    so[i * Z + k] = (uint16_t)(s < 1 ? 1 : s);
                    so[i * Z + k] = (uint16_t)(so[i * Z + k] << 4);
                }
            for (int e = 0; e < g->E; e++) {                          /* :4055-4102 */
                int i = g->col[e];
                for (int k = 0; k < Z; k++) {
                    size_t m = (size_t)e * Z + (k - g->sh[e] + Z) % Z;
                    int sv = so[i * Z + k] << (12 - 4);
                    int sos = msg[m] < 1 ? 1 : msg[m];
                    int p1 = sv / sos;
                    int t = (ONE_SOFT - sos) < 1 ? 1 : (ONE_SOFT - sos);
                    int p0 = (ONE_SOFT * ONE_SOFT - sv) / t;
                    int yy = DIVR(p1 + p0, 6);
                    int y1 = yy < 1 ? 1 : yy;
                    int d = (p1 << 6) / y1;
                    d = d < 1 ? 1 : d;
                    msg[m] = (uint16_t)(MAX_SOFT < d ? MAX_SOFT : d);
                }
            }
        }
        synd = isyndrome(g, so);
        if (synd == 0) return steps + 1;
        steps = steps + 1;
    }
    return -steps;
}

int orc_iasp(const int16_t* hd, int b, int c, int Z, const double* y, int nf, int maxiter,
             uint8_t* hard, int32_t* iters, uint16_t* post)
{
    qc_graph g;
    if (qc_build(&g, hd, b, c, Z)) return -1;
    if (min_row_weight(&g) < 2) { qc_free(&g); return -2; }
    int N = c * Z;
    int all_cw_2 = 1;
    for (int i = 0; i < c; i++) if (g.cp[i + 1] - g.cp[i] != 2) { all_cw_2 = 0; break; }
    uint16_t* yq = (uint16_t*)malloc(sizeof(uint16_t) * N);
    uint16_t* so = (uint16_t*)malloc(sizeof(uint16_t) * N);
    uint16_t* msg = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)g.E * Z);
    uint16_t* a = (uint16_t*)malloc(sizeof(uint16_t) * (g.maxdeg + 1));
    int16_t* w = (int16_t*)malloc(sizeof(int16_t) * 3 * (g.maxdeg + 2));
    for (int f = 0; f < nf; f++) {
        iters[f] = iasp_frame(&g, y + (size_t)f * N, maxiter, all_cw_2, yq, so, msg, a, w);
        if (hard) for (int i = 0; i < N; i++) hard[(size_t)f * N + i] = (uint8_t)(so[i] >> 15);
        if (post) memcpy(post + (size_t)f * N, so, sizeof(uint16_t) * N);
    }
    free(yq); free(so); free(msg); free(a); free(w);
    qc_free(&g);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * channel: QAM demodulator / modulator, BPSK LLR, sigma, error counting
 * ---------------------------------------------------------------------------------------- */

/* one output of Demodulate's repeated if-ladder, e.g. QAM_demodulator.cpp:215-239 */
static double demod_out(double p0, double p1, double T, int out_type)
{
    if (p0 == 0.0) return out_type == 0 ? T : 1.0;
    if (p1 == 0.0) return out_type == 0 ? -T : 0.0;
    return out_type == 0 ? log(p1 / p0) : p1;
}

int orc_demodulate(int m, int ns, double sigma, double T, int out_type, const double* x, double* res)
{
    int n = ns * m;
    if (m == 2) {                                                     /* QAM-4, :114-140 */
        double sigma2 = sigma * sigma;
        for (int j = 0; j < ns; j++) {
            res[2 * j] = 2.0 * x[2 * j] / sigma2;
            res[2 * j + 1] = 2.0 * x[2 * j + 1] / sigma2;
        }
        if (out_type) {
            double P = 0.0;
            for (int i = 0; i < n; i++) { res[i] = exp(res[i]); P += res[i]; }
            for (int i = 0; i < n; i++) res[i] /= P;
        }
        return 0;
    }
    if (m != 4 && m != 6 && m != 8) return -1;
    const double N0 = 2.0 * sigma * sigma;                            /* :144 */
    const int SQ = 1 << (m / 2);
    double P[16];
    for (int ix = 0; ix < ns; ix++) {
        int h = 0;
        for (int j = 0; j < 2; j++) {                                 /* I then Q, :173 */
            double sum = 0;
            for (int i1 = 0; i1 < SQ; i1++) {
                double tmp = x[2 * ix + j] - (2 * i1 - (SQ - 1));     /* lattice, :150-155 */
                tmp *= tmp;
                tmp /= N0;
                P[i1] = tmp < T ? exp(-tmp) : 0.0;                    /* :185-188 */
                sum += P[i1];
            }
            for (int i1 = 0; i1 < SQ; i1++) P[i1] /= sum;             /* :196-199 */
            double* o = res + (size_t)ix * m + h;
            if (m == 4) {                                             /* :203-275 */
                o[0] = demod_out(P[0] + P[1], P[2] + P[3], T, out_type);
                o[1] = demod_out(P[0] + P[3], P[1] + P[2], T, out_type);
                h += 2;
            } else if (m == 6) {                                      /* :276-395 */
                double p12 = P[0] + P[1], p34 = P[2] + P[3], p56 = P[4] + P[5], p78 = P[6] + P[7];
                double p1234 = p12 + p34, p5678 = p56 + p78, p1278 = p12 + p78, p3456 = p34 + p56;
                o[0] = demod_out(p1234, p5678, T, out_type);
                o[1] = demod_out(p1278, p3456, T, out_type);
                o[2] = demod_out(P[0] + P[3] + P[4] + P[7], P[1] + P[2] + P[5] + P[6], T, out_type);
                h += 3;
            } else {                                                  /* :396-561 */
                double p12 = P[0] + P[1], p34 = P[2] + P[3], p56 = P[4] + P[5], p78 = P[6] + P[7];
                double p9A = P[8] + P[9], pBC = P[10] + P[11], pDE = P[12] + P[13], pFG = P[14] + P[15];
                double p1234 = p12 + p34, p5678 = p56 + p78, p9ABC = p9A + pBC, pDEFG = pDE + pFG;
                double p1to8 = p1234 + p5678, p9toG = p9ABC + pDEFG;
                o[0] = demod_out(p1to8, p9toG, T, out_type);
                o[1] = demod_out(p1234 + pDEFG, p5678 + p9ABC, T, out_type);
                o[2] = demod_out(p12 + p78 + p9A + pFG, p34 + p56 + pBC + pDE, T, out_type);
                o[3] = demod_out(P[0] + P[3] + P[4] + P[7] + P[8] + P[11] + P[12] + P[15],
                                 P[1] + P[2] + P[5] + P[6] + P[9] + P[10] + P[13] + P[14], T, out_type);
                h += 4;
            }
        }
    }
    return 0;
}

int orc_modulate(int m, int ns, const uint8_t* bits, double* out)
{
    static const int gray[16] = { 0, 1, 3, 2, 7, 6, 4, 5, 15, 14, 12, 13, 8, 9, 11, 10 };   /* QAM_modulator.cpp:127 */
    int half = m / 2, off = (1 << half) - 1;
    for (int j = 0; j < ns; j++) {
        int z1 = 0, z2 = 0;
        for (int i = 0; i < half; i++) {                              /* MSB first, :93-94, :151-171 */
            z1 = (z1 << 1) | (bits[j * m + i] & 1);
            z2 = (z2 << 1) | (bits[j * m + half + i] & 1);
        }
        out[2 * j] = 2 * gray[z1] - off;                              /* GrayPAM, :129-140 */
        out[2 * j + 1] = 2 * gray[z2] - off;
    }
    return 0;
}

double orc_sigma_bpsk(double snr_db, int b, int c, int punctured_blocks)
{
    double bitrate = (double)(c - b) / (c - punctured_blocks);        /* bp_simulation.cpp:444 */
    return sqrt(pow(10, -snr_db / 10) / 2 / bitrate);                 /* :445 */
}

double orc_sigma_qam(double snr_db, int b, int c, int punctured_blocks, int Q)
{
    double bitrate = (double)(c - b) / (c - punctured_blocks);
    int halfmlog = Q == 4 ? 1 : Q == 16 ? 2 : Q == 64 ? 3 : Q == 256 ? 4 : 1;   /* :403-411 */
    double norm_factor = 2.0 * (Q - 1.0) / 3.0;                       /* :447 */
    return sqrt(pow(10., -snr_db / 10.) / (2 * bitrate * halfmlog * 2) * norm_factor);   /* :449 */
}

void orc_bpsk_llr(const double* noise, const uint8_t* cw, int n, double sigma, double* llr)
{
    for (int i = 0; i < n; i++) {
        double cwv = cw ? (double)cw[i] : 0.0;
        llr[i] = -2.0 * (sigma * noise[i] + 2.0 * cwv - 1.0) / (sigma * sigma);   /* :603 */
    }
}

void orc_count_errors(const uint8_t* hard, const uint8_t* cw, int N, int R, int* nse, int* nse_info)
{
    int a = 0, bi = 0;
    for (int i = 0; i < N; i++) {
        int cwv = cw ? cw[i] : 0;
        if (hard[i] != cwv) { ++a; if (i >= R) ++bi; }                /* bp_simulation.cpp:735-742 */
    }
    *nse = a;
    *nse_info = bi;
}
