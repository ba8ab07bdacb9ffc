// The QC-LDPC encoder behind random_codeword() (reference bp_simulation.cpp:22-192), host code.
//
// Block columns 0 .. b-1 of the base matrix are parity columns arranged in diagonal blocks -- "unidiagonal" (identity:
// parity = partial syndrome) or "bidiagonal" (dual diagonal closed by the weight-3 column b-1) -- and the encoder
// peels them from the last block to the first.  Off the per-frame path: the simulation sends the all-zero codeword
// (bp_simulation.cpp:568); the search drivers use the return code to reject matrices.
#include <vector>

#include "bp_simulation.h"

namespace {

// syndrome of cword over the block columns [first_col, c): synd[i * M + h] = XOR of cword[j * M + (h + shift) % M]
void accumulate_syndrome(matrix<int> const& mx, int M, std::vector<bit> const& cword, int first_col, std::vector<bit>& synd)
{
    const int b = mx.n_rows(), c = mx.n_cols();
    for (int i = 0; i < b; ++i)
        for (int j = first_col; j < c; ++j) {
            const int shift = mx(i, j);
            if (shift < 0) continue;
            for (int h = 0; h < M; ++h) synd[(size_t)i * M + h] ^= cword[(size_t)j * M + (h + shift) % M];
        }
}

} // namespace

int qc_encode(matrix<int> const& mx, int M, std::vector<bit>& cword)
{
    const int b = mx.n_rows(), c = mx.n_cols();
    const int r = b * M, n = c * M;
    const bool single_diagonal = mx(1, 0) < 0;                      // bp_simulation.cpp:32

    int p = 0;                                                      // a positive shift in column b-1 (:35-41)
    while (p < b && mx(p, b - 1) <= 0) ++p;
    if (p >= b && !single_diagonal) return -1;
    if ((int)cword.size() != n) die("Input word has the length %d but should have been %d", (int)cword.size(), n);

    // partial syndromes of the information part and their sum over the block rows (:47-62)
    std::vector<bit> synd(r), sumsynd(M);
    accumulate_syndrome(mx, M, cword, b, synd);
    for (int i = 0; i < b; ++i)
        for (int h = 0; h < M; ++h) sumsynd[h] ^= synd[(size_t)i * M + h];

    if (single_diagonal) {
        for (int i = 0; i < r; ++i) cword[i] = synd[i];             // :64-68
    } else {
        for (int h = 0; h < M; ++h) {                               // :69-84
            const bool xh = sumsynd[(h + M - mx(p, b - 1)) % M];
            cword[(size_t)(b - 1) * M + h] = xh;
            cword[h] = synd[h];
            if (mx(0, b - 1) == 0) cword[h] ^= bit(xh);
            if (mx(0, b - 1) > 0) cword[h] ^= sumsynd[h];
            for (int i = 1; i < b - 1; ++i) {
                const size_t idx = (size_t)i * M + h;
                cword[idx] = (bool)synd[idx] != (bool)cword[idx - M];
                if (mx(i, b - 1) == 0) cword[idx] ^= cword[(size_t)(b - 1) * M + h];
                if (mx(i, b - 1) > 0) cword[idx] ^= sumsynd[h];
            }
        }
    }

    // is it a codeword? (:87-117; the reference also dumps cw.txt on failure -- not reproduced)
    std::vector<bit> check(r);
    accumulate_syndrome(mx, M, cword, 0, check);
    for (int i = 0; i < r; ++i)
        if (check[i]) return 1;
    return 0;
}

static void independent_validation(matrix<int> const& mx, int M, std::vector<bit> const& codeword)
{
    std::vector<bit> syndrome((size_t)mx.n_rows() * M);
    accumulate_syndrome(mx, M, codeword, 0, syndrome);
    for (size_t k = 0; k < syndrome.size(); ++k)
        if (syndrome[k]) die("Codeword is bad [independent validation]");
}

int random_codeword(matrix<int> const& mx, int M, std::vector<bit>& codeword)
{
    // where the diagonal blocks of the parity part begin (:143-155)
    std::vector<int> breaks(1, 0);
    for (int i = 1, b = mx.n_rows(); i + 1 < b; ++i) {
        if (mx(i, i) >= 0 && mx(i + 1, i) >= 0 && mx(i, i - 1) < 0) breaks.push_back(i);                      // a bidiagonal block begins
        if (i > 1 && mx(i, i) >= 0 && mx(i + 1, i) < 0 && mx(i, i - 1) < 0 && mx(i - 1, i - 1) >= 0 && mx(i - 1, i - 2) >= 0)
            breaks.push_back(i);                                                                              // a unidiagonal block begins
    }
    breaks.push_back(mx.n_rows());

    const int n = mx.n_cols() * M;
    std::vector<bit> cword(n);
    for (int i = breaks.back() * M; i < n; ++i) cword[i] = next_random_int(0, 2) == 1;                        // random information bits

    for (int k = (int)breaks.size() - 2; k >= 0; --k) {             // encode block by block, last block first (:164-185)
        const int offset = breaks[k], rows = breaks[k + 1] - offset;
        matrix<int> sub(rows, mx.n_cols() - offset);
        for (int rr = 0; rr < sub.n_rows(); ++rr)
            for (int cc = 0; cc < sub.n_cols(); ++cc) sub(rr, cc) = mx(rr + offset, cc + offset);
        std::vector<bit> local(cword.begin() + (size_t)offset * M, cword.end());
        const int rc = qc_encode(sub, M, local);
        if (rc != 0) return rc;
        for (size_t j = 0; j < local.size(); ++j) cword[(size_t)offset * M + j] = local[j];
    }
    codeword = cword;
    independent_validation(mx, M, codeword);
    return 0;
}
