// Bit interleavers of the simulation path: the index tables of the reference's five permutation modes
// (direct_inverse_perm.cpp: Permutations_Open :139, Permutation_Init :312-782, Permutation :785-896, LCG myrand :131-135),
// restated as "which codeword bit travels at transmitted position j" (direct) and "which transmitted position feeds
// decoder input i" (inverse).  Host code; the tables are uploaded once per handle (ldpcb200_set_interleaver) and applied
// as a gather / scatter inside the decoder's first load (channel.cuh).
//
//   mode 0  identity
//   mode 1  one random permutation of all N positions
//   mode 2  deterministic: the block columns are dealt to the bit positions of a PAM symbol so that the columns of maximum
//           weight (the reference assumes they are the first `mrp` block columns) take one fixed position of every symbol
//   mode 3  the same random permutation inside every block of `block` positions (+ one for the short last block)
//   mode 4  the same random permutation along every stride-`inter` comb
#include <cstdint>
#include <vector>

#include "../../include/ldpcb200.h"

namespace {

struct Lcg {                                    // myrand(), :131-135; Permutations_Open resets the state to 1 (:172)
    unsigned state = 1;
    int next() { state = state * 1103515245u + 12345u; return (int)(state & 0x3FFFFFFF); }
};

// random_perm_gen (:284-302): draw until `size` distinct values below `size` have appeared, in order of first appearance
std::vector<int> random_perm(Lcg& g, int size)
{
    std::vector<int> p;
    std::vector<char> seen(size, 0);
    p.reserve(size);
    while ((int)p.size() < size) {
        const int r = g.next() % size;
        if (seen[r]) continue;
        seen[r] = 1;
        p.push_back(r);
    }
    return p;
}

// invperm of a table (:728-771): position of value i
std::vector<int> inverse_of(const std::vector<int>& p)
{
    std::vector<int> inv(p.size(), -1);
    for (size_t j = 0; j < p.size(); j++)
        if (p[j] >= 0 && p[j] < (int)p.size()) inv[p[j]] = (int)j;
    return inv;
}

// mode 2 (:470-676).  h = bits per PAM component, cw = block column weights
bool deterministic_table(const std::vector<int>& cw, int c, int M, int h, std::vector<int>& perm)
{
    const int N = c * M, c0 = c / h, c0_mod = c % h;
    int mw = 0;
    for (int j = 0; j < c; j++) mw = cw[j] > mw ? cw[j] : mw;
    // the block columns dealt round-robin to the h bit positions (:363-380): 0, h, 2h, ... | 1, h + 1, ... | ...; the
    // columns beyond c0 * h keep their place
    std::vector<int> V(c + h + 1, 0);
    for (int j = 0; j < h; j++)
        for (int i = 0; i * h < c; i++) V[j * c0 + i] = j + i * h;
    if (c0_mod)
        for (int k = c0 * h; k < c; k++) V[k] = k;
    int mrp = 0;                                // number of maximum-weight columns (:382-386)
    for (int i = 0; i < c; i++) mrp += cw[V[i]] == mw ? 1 : 0;
    // the first mrp block columns go to the end of the list (:389-403)
    std::vector<int> order;
    for (int i = 0; i < c; i++)
        if (V[i] >= mrp) order.push_back(V[i]);
    for (int i = 0; i < mrp; i++) order.push_back(i);
    if ((int)order.size() != c) return false;
    perm.assign(N, -1);
    if (c0_mod == 0) {                          // (:472-488) bit position i of every symbol component takes the i-th c0 columns of the list
        for (int i = 0; i < h; i++)
            for (int j = 0; j < c0; j++)
                for (int k = 0; k < M; k++) perm[i + h * (j * M + k)] = order[i * c0 + j] * M + k;
        return true;
    }
    // (:490-676) `limit` columns travel as they are; then, per maximum-weight column, h columns are interleaved bit by bit
    // so that the maximum-weight one keeps the same bit position of the component whatever the offset
    const int ibad = c - mrp, limit = ibad - (h - 1) * mrp;
    if (limit < 0) return false;
    for (int i = 0; i < limit; i++)
        for (int k = 0; k < M; k++) perm[i * M + k] = order[i] * M + k;
    const int rem = (limit * M) % h;
    // slot -> which of the h column groups (0 .. h-2 = "mix" groups counted from the one next to the bad group, h-1 = bad)
    static const int ORDER[5][4][4] = {
        {}, {},
        { { 0, 1 }, { 1, 0 } },                                                         // h = 2: (mix1, bad) | (bad, mix1)
        { { 0, 1, 2 }, { 0, 2, 1 }, { 2, 0, 1 } },                                      // h = 3
        { { 0, 1, 2, 3 }, { 0, 1, 3, 2 }, { 2, 3, 0, 1 }, { 3, 2, 0, 1 } } };           // h = 4
    int l = limit * M;
    for (int i = 0; i < mrp; i++)
        for (int k = 0; k < M; k++)
            for (int slot = 0; slot < h; slot++) {
                const int grp = ORDER[h][rem][slot];
                const int col = grp == h - 1 ? order[ibad + i] : order[ibad - (grp + 1) * mrp + i];
                perm[l++] = col * M + k;
            }
    return l == N;
}

} // namespace

// direct[j] = codeword bit at transmitted position j; inverse[i] = transmitted position that feeds decoder input i.
// Returns LDPCB200_EUNSUPPORTED when the parameters do not define a permutation of the N positions (the reference would
// then read stale buffer contents).
extern "C" int ldpcb200_interleaver_tables(const int16_t* hd, int b, int c, int Z, int modulation, int mode, int block, int inter,
                                           int32_t* direct, int32_t* inverse)
{
    if (!hd || b <= 0 || c <= 0 || Z <= 0 || !direct || !inverse) return LDPCB200_EINVAL;
    if (mode < 0 || mode > 4 || modulation < 0 || modulation > 4) return LDPCB200_EINVAL;
    static const int HALF[5] = { 1, 1, 2, 3, 4 };              // bits per PAM component: no modulation, QAM-4, 16, 64, 256 (bp_simulation.cpp:403-411)
    const int h = HALF[modulation], N = c * Z;
    std::vector<int> dir(N, -1), inv(N, -1);
    Lcg g;
    if (mode == 0) {
        for (int i = 0; i < N; i++) dir[i] = inv[i] = i;
    } else if (mode == 1) {
        dir = random_perm(g, N);
        inv = inverse_of(dir);
    } else if (mode == 2) {
        std::vector<int> cw(c, 0);
        for (int j = 0; j < c; j++)
            for (int i = 0; i < b; i++) cw[j] += hd[(size_t)i * c + j] >= 0 ? 1 : 0;
        if (!deterministic_table(cw, c, Z, h, dir)) return LDPCB200_EUNSUPPORTED;
        inv = inverse_of(dir);
    } else {
        // modes 3 / 4 (:158-171): a table for the full blocks and one for the short last block
        if ((mode == 3 && block <= 0) || (mode == 4 && (inter <= 0 || inter > N))) return LDPCB200_EINVAL;
        const int bs = mode == 3 ? block : N / inter;
        if (bs <= 0) return LDPCB200_EINVAL;
        const int nblocks = N / bs, shortb = N % bs;
        const std::vector<int> p = random_perm(g, bs), ps = shortb ? random_perm(g, shortb) : std::vector<int>();
        const std::vector<int> ip = inverse_of(p), ips = inverse_of(ps);
        if (mode == 3) {                         // Permutation :806-851
            for (int k = 0; k < nblocks; k++)
                for (int i = 0; i < bs; i++) { dir[k * bs + i] = k * bs + p[i]; inv[k * bs + i] = k * bs + ip[i]; }
            for (int i = 0; i < shortb; i++) { dir[nblocks * bs + i] = nblocks * bs + ps[i]; inv[nblocks * bs + i] = nblocks * bs + ips[i]; }
        } else {                                 // Permutation :852-893: comb k takes positions k, k + inter, ...
            for (int k = 0; k < nblocks; k++)
                for (int i = 0; i < bs; i++) {
                    const long long j = k + (long long)i * inter;
                    if (j >= N || k + (long long)p[i] * inter >= N) return LDPCB200_EUNSUPPORTED;
                    dir[j] = k + p[i] * inter; inv[j] = k + ip[i] * inter;
                }
            for (int i = 0; i < shortb; i++) {
                const long long j = nblocks + (long long)i * inter;
                if (j >= N || nblocks + (long long)ps[i] * inter >= N) return LDPCB200_EUNSUPPORTED;
                dir[j] = nblocks + ps[i] * inter; inv[j] = nblocks + ips[i] * inter;
            }
        }
    }
    // both must be permutations and inverse to each other
    std::vector<char> seen(N, 0);
    for (int j = 0; j < N; j++) {
        if (dir[j] < 0 || dir[j] >= N || seen[dir[j]]) return LDPCB200_EUNSUPPORTED;
        seen[dir[j]] = 1;
    }
    for (int i = 0; i < N; i++)
        if (inv[i] < 0 || inv[i] >= N || dir[inv[i]] != i) return LDPCB200_EUNSUPPORTED;
    for (int i = 0; i < N; i++) { direct[i] = dir[i]; inverse[i] = inv[i]; }
    return 0;
}
