// Uses the decoder interface exactly the way the reference's callers do (bp_simulation.cpp:353-382, 716-729):
// decod_open -> fill st->hd -> decod_init -> copy LLRs into st->y -> call the decoder -> read st->decword.
// argv: <decoder id> <b> <c> <M> <maxiter> <hd.bin int16> <llr.bin f64> <out.bin> [decision]; out = per frame: int32 iter + N bytes,
// or with decision given: int32 iter + N doubles decword + N doubles = the input array after the call
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "decoders.h"

int main(int argc, char** argv)
{
    if (argc != 9 && argc != 10) return 2;
    const int decision = argc == 10 ? atoi(argv[9]) : DEC_DECISION;
#undef DEC_DECISION
#define DEC_DECISION decision
    int id = atoi(argv[1]), b = atoi(argv[2]), c = atoi(argv[3]), M = atoi(argv[4]), maxiter = atoi(argv[5]);
    std::vector<short> hd((size_t)b * c);
    FILE* f = fopen(argv[6], "rb");
    if (!f || fread(hd.data(), 2, hd.size(), f) != hd.size()) return 3;
    fclose(f);
    DEC_STATE* st = decod_open(id, 1, b, c, M);
    if (!st) return 4;
    for (int i = 0; i < b; i++)
        for (int j = 0; j < c; j++) st->hd[i][j] = hd[(size_t)i * c + j];
    if (!decod_init(st)) return 5;
    st->bp_chain = 1;
    FILE* in = fopen(argv[7], "rb");
    FILE* out = fopen(argv[8], "wb");
    if (!in || !out) return 6;
    const int n = st->n;
    std::vector<unsigned char> bits(n);
    while (fread(st->y, sizeof(double), n, in) == (size_t)n) {
        int iter;
        switch (id) {
        case BP_DEC:   iter = bp_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION); break;
        case SP_DEC:   iter = sum_prod_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION); break;
        case ASP_DEC:  iter = sum_prod_gf2_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION); break;
        case MS_DEC:   iter = min_sum_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION, MS_ALPHA); break;
        case IMS_DEC:  iter = imin_sum_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION, MS_ALPHA, MS_THR, MS_QBITS, MS_DBITS); break;
        case IASP_DEC: iter = isum_prod_gf2_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION); break;
        case TASP_DEC: iter = tdmp_sum_prod_gf2_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION); break;
        case LMS_DEC:  iter = lmin_sum_decod_qc_lm(st, st->y, st->decword, maxiter, DEC_DECISION, MS_ALPHA, MS_BETA); break;
        case LCHE_DEC: iter = lche_decod(st, st->y, st->decword, maxiter, DEC_DECISION); break;
        default: return 7;
        }
        fwrite(&iter, 4, 1, out);
        if (argc == 10) {
            fwrite(st->decword, sizeof(double), n, out);
            fwrite(st->y, sizeof(double), n, out);
        } else {
            for (int i = 0; i < n; i++) bits[i] = st->decword[i] != 0.0;
            fwrite(bits.data(), 1, n, out);
        }
    }
    fclose(in); fclose(out);
    decod_close(st);
    return 0;
}
