// Girth, cycle spectrum and ACE spectrum of a QC base matrix -- what the reference's driver prints and stores with every
// result (main_simulation.cpp:148-205 trace_matrix -> trace_pm.cpp:58 trace_bound_pol_mon_pm).  Host code.
//
// The reference's method, restated: the protograph's E edges become 2E directed edges (variable -> check carries the
// circulant shift s, check -> variable carries -s mod Z).  A[I][J] describes the two-step non-backtracking move "forward
// edge I, then back along another edge of I's check to a variable v, then out of v along forward edge J (not the edge just
// used)"; its entry is the monomial x^(shift sum of the two appended edges) with an "ACE" weight 2 * (column weight of v)
// (tanner_mon :301, hp2a_mon_pm :352).  B = A^k holds, per pair of forward edges, a polynomial mod x^Z - 1 whose coefficient
// at x^t counts such walks of 2k edges with shift sum t, and per term the smallest accumulated weight over those walks.
// The constant term of a diagonal entry B[I][I] counts closed walks of 2k edges that close in the lifted graph too: cycles
// of length 2k, each counted 2k times over all I.  After counting, that term is removed from B[I][I] so that longer walks
// through a counted cycle are not counted again (:205-222) -- with the reference's quirk that the weights of the remaining
// terms of that entry move up by one term ("know how", :213-218), which is reproduced.  The sweep stops after `gtarget`
// distinct cycle lengths (or at length gmax).
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../../include/ldpcb200.h"

namespace {

struct Poly {                       // polynomial mod x^Z - 1 with a weight per term; empty() <=> zero polynomial
    std::vector<int> coef, ace;
    bool zero() const { return coef.empty(); }
};

} // namespace

extern "C" int ldpcb200_girth_spectrum(const int16_t* hd, int b, int c, int Z, int gtarget, int* girth, int* ace, int* spectrum)
{
    if (!hd || b <= 0 || c <= 0 || Z <= 0 || gtarget <= 0 || gtarget > 16) return LDPCB200_EINVAL;
    const int GMAX = 20;                                        // trace_pm.h:8
    // protograph edges in the reference's order: block column by block column, rows ascending (tanner_mon :322-334)
    std::vector<int> ecol, erow, esh, cw(c, 0);
    for (int i = 0; i < c; i++)
        for (int j = 0; j < b; j++)
            if (hd[(size_t)j * c + i] >= 0) { ecol.push_back(i); erow.push_back(j); esh.push_back(hd[(size_t)j * c + i] % Z); cw[i]++; }
    const int N = (int)ecol.size();
    // two-step transitions between forward edges: I (v -> check r), back edge jb (r -> v', jb != I), forward edge J out of v' (J != jb)
    struct Step { int to, shift, weight; };
    std::vector<std::vector<Step>> A(N);                        // A[I] = the J reachable from I (hp2a_mon_pm :440-460: Am * Ap)
    for (int I = 0; I < N; I++)
        for (int jb = 0; jb < N; jb++) {
            if (jb == I || erow[jb] != erow[I]) continue;
            for (int J = 0; J < N; J++) {
                if (J == jb || ecol[J] != ecol[jb]) continue;
                A[I].push_back({ J, ((Z - esh[jb]) % Z + esh[J]) % Z, 2 * cw[ecol[jb]] });
            }
        }
    // incoming lists: for column J of the product, the (j, shift, weight) with A[j][J] set, j ascending (mul_mat_mat_mon :690)
    std::vector<std::vector<Step>> in(N);
    for (int j = 0; j < N; j++)
        for (const Step& s : A[j]) in[s.to].push_back({ j, s.shift, s.weight });
    for (auto& v : in) std::sort(v.begin(), v.end(), [](const Step& x, const Step& y) { return x.to < y.to; });

    std::vector<Poly> B((size_t)N * N), Y((size_t)N * N);
    for (int I = 0; I < N; I++)
        for (const Step& s : A[I]) {                             // B = A (:113-129)
            Poly& p = B[(size_t)I * N + s.to];
            if (p.zero()) { p.coef.assign(Z, 0); p.ace.assign(Z, 0); }
            p.coef[s.shift] = 1;
            p.ace[s.shift] = s.weight;
        }
    int S[GMAX] = { 0 }, SA[GMAX] = { 0 }, found = 0;
    std::vector<int> keep;
    for (int d = 3; d < GMAX; d += 2) {
        for (auto& y : Y) { y.coef.clear(); y.ace.clear(); }
        for (int I = 0; I < N; I++)
            for (int J = 0; J < N; J++) {
                Poly& y = Y[(size_t)I * N + J];
                for (const Step& s : in[J]) {                    // Y[I][J] += B[I][j] * x^shift, weights + weight (:690-727)
                    const Poly& p = B[(size_t)I * N + s.to];
                    if (p.zero()) continue;
                    if (y.zero()) { y.coef.assign(Z, 0); y.ace.assign(Z, 0); }
                    for (int t = 0; t < Z; t++) {
                        const int cf = p.coef[t];
                        if (!cf) continue;
                        const int u = t + s.shift >= Z ? t + s.shift - Z : t + s.shift, w = p.ace[t] + s.weight;
                        if (y.coef[u] == 0) { y.coef[u] = cf; y.ace[u] = w; }            // add_pol :632-656
                        else { y.coef[u] += cf; y.ace[u] = std::min(y.ace[u], w); }
                    }
                }
            }
        B.swap(Y);
        for (int I = 0; I < N; I++) {                            // closed walks that close in the lifted graph (:187-222)
            Poly& p = B[(size_t)I * N + I];
            if (p.zero() || p.coef[0] == 0) continue;
            S[d] += p.coef[0];
            SA[d] = SA[d] > 0 ? std::min(SA[d], p.ace[0]) : p.ace[0];
            // remove the constant term; the k-th remaining term takes the weight of the k-th term of the old list
            keep.clear();
            for (int t = 0; t < Z; t++) if (p.coef[t]) keep.push_back(p.ace[t]);
            p.coef[0] = 0;
            int k = 0;
            for (int t = 0; t < Z; t++) p.ace[t] = p.coef[t] ? keep[k++] : 0;
        }
        if (S[d] > 0 && ++found == gtarget) break;               // (:224-246 with the driver's limits: nothing is ever rejected)
    }
    for (int i = 3; i < GMAX; i += 2) { S[i] /= i + 1; SA[i] /= 2; }     // :250-254
    int g = 1;
    while (g < GMAX + 1 && !S[g - 1]) g++;                       // main_simulation.cpp:175-179
    if (girth) *girth = g;
    for (int k = 0; k < gtarget; k++) { if (ace) ace[k] = 0; if (spectrum) spectrum[k] = 0; }
    for (int j = 0, i = g - 1; i < GMAX && j < gtarget; i++) if (SA[i] && ace) ace[j++] = SA[i]; else if (SA[i]) j++;
    for (int j = 0, i = g - 1; i < GMAX && j < gtarget; i++) if (S[i] && spectrum) spectrum[j++] = S[i]; else if (S[i]) j++;
    return 0;
}
