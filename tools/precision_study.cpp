// Study tool (CPU, not product code): can the float class of LMS_DEC (lmin_sum_decod_qc_lm, decoders.cpp:5064-5425)
// run in reduced precision and still meet the parity bar of BASELINE.json -- identical hard decisions and iteration
// counts on >= 99.99 % of frames, posteriors within 1e-4 relative -- against the reference's double arithmetic?
//
// Variants (same schedule, same edge order, row-lane formulation of oracle/ldpc_oracle_impl.h):
//   f64      the reference's arithmetic
//   f32      everything in float (what the GPU kernels do)
//   m16f     float arithmetic, check-to-variable messages rounded to IEEE half when stored
//   mbf16    float arithmetic, messages rounded to bfloat16 when stored
//   p16f     float arithmetic, posteriors rounded to IEEE half when stored (and messages)
//   q7 / q6 / q5   int16 fixed point with 7 / 6 / 5 fractional bits throughout (saturating), channel LLR rounded to the grid
//
//   g++ -O2 -fopenmp tools/precision_study.cpp -o /tmp/precision_study
//   /tmp/precision_study configs/ref32x16_b.jsonx 256 10 2.0 200000 > profiles/r02_precision_study_c2_2.0dB.json
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

struct Graph { int b, c, Z, E; std::vector<int> rp, col, sh; };

static Graph load(const char* path, int Z)
{
    FILE* f = fopen(path, "r");
    if (!f) { perror(path); exit(1); }
    std::string t; char buf[4096]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) t.append(buf, n);
    fclose(f);
    size_t p = t.find("matrix");
    int b, c;
    sscanf(t.c_str() + t.find('(', p), "(%d %d)", &b, &c);
    const char* q = t.c_str() + t.find('{', p) + 1;
    Graph g; g.b = b; g.c = c; g.Z = Z; g.rp.push_back(0);
    for (int j = 0; j < b; j++) {
        for (int i = 0; i < c; i++) {
            char* e; long v = strtol(q, &e, 10); q = e;
            if (v >= 0) { g.col.push_back(i); g.sh.push_back((int)(v % Z)); }
        }
        g.rp.push_back((int)g.col.size());
    }
    g.E = (int)g.col.size();
    return g;
}

static float to_half(float x)      // round to nearest even onto the IEEE binary16 grid (no overflow handling needed: |x| <= 32767)
{
    if (x == 0.0f || !std::isfinite(x)) return x;
    int e; float m = std::frexp(std::fabs(x), &e);            // |x| = m * 2^e, m in [0.5, 1)
    if (e < -13) {                                           // subnormal half: grid 2^-24
        return std::copysign(std::nearbyint(std::fabs(x) * 16777216.0f) / 16777216.0f, x);
    }
    float q = std::nearbyint(m * 2048.0f) / 2048.0f;          // 11 significant bits
    float r = std::ldexp(q, e);
    if (r > 65504.0f) r = 65504.0f;
    return std::copysign(r, x);
}
static float to_bf16(float x)
{
    uint32_t u; memcpy(&u, &x, 4);
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb; u &= 0xffff0000u;
    float r; memcpy(&r, &u, 4); return r;
}

enum Mode { F64, F32, M16F, MBF16, P16F, Q7, Q6, Q5, NMODES };
static const char* NAMES[NMODES] = { "f64", "f32", "m16f", "mbf16", "p16f", "q7", "q6", "q5" };

template <class R>
struct Row { R min1, min2; int pos, sign; };

// float / double flavours; round_msg / round_post applied at the stores
template <class R, int MODE>
static int lms(const Graph& g, const double* y, int maxiter, std::vector<double>& post)
{
    const int Z = g.Z, N = g.c * Z, Rr = g.b * Z;
    std::vector<R> soft(N), v2c(32);
    std::vector<Row<R>> prev(Rr, Row<R>{0, 0, 0, 0});
    std::vector<uint8_t> esign((size_t)g.E * Z, 0);
    auto rm = [](R x) -> R { return MODE == M16F || MODE == P16F ? (R)to_half((float)x) : MODE == MBF16 ? (R)to_bf16((float)x) : x; };
    auto rp = [](R x) -> R { return MODE == P16F ? (R)to_half((float)x) : x; };
    for (int i = 0; i < N; i++) soft[i] = rp((R)y[i]);
    auto syndrome = [&]() {
        int parity = 0;
        for (int j = 0; j < g.b; j++)
            for (int n = 0; n < Z; n++) {
                int s = 0;
                for (int e = g.rp[j]; e < g.rp[j + 1]; e++) s ^= soft[g.col[e] * Z + (n + g.sh[e]) % Z] < 0;
                parity |= s;
            }
        return parity;
    };
    int parity = syndrome(), iter;
    for (iter = 0; iter < maxiter; iter++) {
        if (!parity) break;
        for (int j = 0; j < g.b; j++)
            for (int n = 0; n < Z; n++) {
                Row<R> cur{ (R)32767, (R)32767, 0, 0 };
                Row<R>& pr = prev[j * Z + n];
                for (int e = g.rp[j]; e < g.rp[j + 1]; e++) {
                    const int k = g.col[e], idx = k * Z + (n + g.sh[e]) % Z;
                    const R pabs = pr.pos == k ? pr.min2 : pr.min1;
                    const int psgn = esign[(size_t)e * Z + n] ^ pr.sign;
                    const R v = soft[idx] - (psgn ? -pabs : pabs);
                    const int s = v < 0;
                    R a = v < 0 ? -v : v;
                    a -= (R)0.4;
                    if (a < 0) a = 0;
                    v2c[e - g.rp[j]] = v;
                    esign[(size_t)e * Z + n] = (uint8_t)s;
                    cur.sign ^= s;
                    if (a < cur.min1) { cur.min2 = cur.min1; cur.min1 = a; cur.pos = k; }
                    else if (a < cur.min2) cur.min2 = a;
                }
                cur.min1 = rm(cur.min1); cur.min2 = rm(cur.min2);
                for (int e = g.rp[j]; e < g.rp[j + 1]; e++) {
                    const int k = g.col[e], idx = k * Z + (n + g.sh[e]) % Z;
                    const R mabs = cur.pos == k ? cur.min2 : cur.min1;
                    const int sg = esign[(size_t)e * Z + n] ^ cur.sign;
                    soft[idx] = rp(v2c[e - g.rp[j]] + (sg ? -mabs : mabs));
                }
                pr = cur;
            }
        parity = syndrome();
        if (!parity) { iter++; break; }
    }
    for (int i = 0; i < N; i++) post[i] = (double)soft[i];
    return parity ? -iter : (iter == 0 ? 1 : iter);            // iteration count convention is the same for every variant
}

// int16 fixed point with FB fractional bits, saturating adds; offset 0.4 rounded to the grid
template <int FB>
static int lms_q(const Graph& g, const double* y, int maxiter, std::vector<double>& post)
{
    const int Z = g.Z, N = g.c * Z, Rr = g.b * Z;
    const int MAXQ = 32767, BETA = (int)std::lround(0.4 * (1 << FB));
    auto sat = [&](int x) { return x > MAXQ ? MAXQ : x < -MAXQ ? -MAXQ : x; };
    std::vector<int> soft(N), v2c(32);
    std::vector<Row<int>> prev(Rr, Row<int>{0, 0, 0, 0});
    std::vector<uint8_t> esign((size_t)g.E * Z, 0);
    for (int i = 0; i < N; i++) soft[i] = sat((int)std::lround(y[i] * (1 << FB)));
    auto syndrome = [&]() {
        int parity = 0;
        for (int j = 0; j < g.b; j++)
            for (int n = 0; n < Z; n++) {
                int s = 0;
                for (int e = g.rp[j]; e < g.rp[j + 1]; e++) s ^= soft[g.col[e] * Z + (n + g.sh[e]) % Z] < 0;
                parity |= s;
            }
        return parity;
    };
    int parity = syndrome(), iter;
    for (iter = 0; iter < maxiter; iter++) {
        if (!parity) break;
        for (int j = 0; j < g.b; j++)
            for (int n = 0; n < Z; n++) {
                Row<int> cur{ MAXQ, MAXQ, 0, 0 };
                Row<int>& pr = prev[j * Z + n];
                for (int e = g.rp[j]; e < g.rp[j + 1]; e++) {
                    const int k = g.col[e], idx = k * Z + (n + g.sh[e]) % Z;
                    const int pabs = pr.pos == k ? pr.min2 : pr.min1;
                    const int psgn = esign[(size_t)e * Z + n] ^ pr.sign;
                    const int v = sat(soft[idx] - (psgn ? -pabs : pabs));
                    const int s = v < 0;
                    int a = (v < 0 ? -v : v) - BETA;
                    if (a < 0) a = 0;
                    v2c[e - g.rp[j]] = v;
                    esign[(size_t)e * Z + n] = (uint8_t)s;
                    cur.sign ^= s;
                    if (a < cur.min1) { cur.min2 = cur.min1; cur.min1 = a; cur.pos = k; }
                    else if (a < cur.min2) cur.min2 = a;
                }
                for (int e = g.rp[j]; e < g.rp[j + 1]; e++) {
                    const int k = g.col[e], idx = k * Z + (n + g.sh[e]) % Z;
                    const int mabs = cur.pos == k ? cur.min2 : cur.min1;
                    const int sg = esign[(size_t)e * Z + n] ^ cur.sign;
                    soft[idx] = sat(v2c[e - g.rp[j]] + (sg ? -mabs : mabs));
                }
                pr = cur;
            }
        parity = syndrome();
        if (!parity) { iter++; break; }
    }
    for (int i = 0; i < N; i++) post[i] = (double)soft[i] / (1 << FB);
    return parity ? -iter : (iter == 0 ? 1 : iter);
}

int main(int argc, char** argv)
{
    if (argc < 6) { fprintf(stderr, "usage: %s code.jsonx Z maxiter snr_db frames\n", argv[0]); return 2; }
    const Graph g = load(argv[1], atoi(argv[2]));
    const int maxiter = atoi(argv[3]);
    const double snr = atof(argv[4]);
    const long frames = atol(argv[5]);
    const int N = g.c * g.Z;
    const double rate = (double)(g.c - g.b) / g.c, sigma = std::sqrt(std::pow(10.0, -snr / 10.0) / (2.0 * rate));
    long diff_iter[NMODES] = {0}, diff_hard[NMODES] = {0}, diff_frame[NMODES] = {0}, post_frames[NMODES] = {0};
    double post_vals[NMODES] = {0}, worst[NMODES] = {0};
    long fails = 0; double itsum = 0;
#pragma omp parallel
    {
        std::vector<double> y(N), ref(N), got(N);
        long l_iter[NMODES] = {0}, l_hard[NMODES] = {0}, l_frame[NMODES] = {0}, l_pf[NMODES] = {0}, l_fail = 0;
        double l_pv[NMODES] = {0}, l_worst[NMODES] = {0}, l_it = 0;
#pragma omp for schedule(dynamic, 64)
        for (long f = 0; f < frames; f++) {
            std::mt19937_64 rng(0x9E3779B97F4A7C15ull * (unsigned long long)(f + 1));
            std::normal_distribution<double> nd(0.0, 1.0);
            for (int i = 0; i < N; i++) y[i] = -2.0 * (sigma * nd(rng) - 1.0) / (sigma * sigma);      // all-zero codeword, bp_simulation.cpp:603
            const int it0 = lms<double, F64>(g, y.data(), maxiter, ref);
            l_fail += it0 < 0; l_it += std::abs(it0);
            for (int m = 1; m < NMODES; m++) {
                int it;
                switch (m) {
                case F32: it = lms<float, F32>(g, y.data(), maxiter, got); break;
                case M16F: it = lms<float, M16F>(g, y.data(), maxiter, got); break;
                case MBF16: it = lms<float, MBF16>(g, y.data(), maxiter, got); break;
                case P16F: it = lms<float, P16F>(g, y.data(), maxiter, got); break;
                case Q7: it = lms_q<7>(g, y.data(), maxiter, got); break;
                case Q6: it = lms_q<6>(g, y.data(), maxiter, got); break;
                default: it = lms_q<5>(g, y.data(), maxiter, got); break;
                }
                bool hd = false; long nbad = 0; double w = 0;
                for (int i = 0; i < N; i++) {
                    hd |= (ref[i] < 0) != (got[i] < 0);
                    const double rel = std::fabs(got[i] - ref[i]) / std::max(std::fabs(ref[i]), 1.0);
                    if (rel > 1e-4) nbad++;
                    if (rel > w) w = rel;
                }
                l_iter[m] += it != it0; l_hard[m] += hd; l_frame[m] += (it != it0) || hd;
                if (it == it0 && !hd) { l_pf[m] += nbad > 0; l_pv[m] += (double)nbad; if (w > l_worst[m]) l_worst[m] = w; }
            }
        }
#pragma omp critical
        {
            fails += l_fail; itsum += l_it;
            for (int m = 0; m < NMODES; m++) {
                diff_iter[m] += l_iter[m]; diff_hard[m] += l_hard[m]; diff_frame[m] += l_frame[m]; post_frames[m] += l_pf[m];
                post_vals[m] += l_pv[m]; if (l_worst[m] > worst[m]) worst[m] = l_worst[m];
            }
        }
    }
    printf("{\"code\": \"%s\", \"Z\": %d, \"maxiter\": %d, \"snr_db\": %g, \"frames\": %ld, \"reference\": \"f64\", \"fer_f64\": %g, \"avg_iters_f64\": %g,\n"
           " \"bar\": \"<= 1e-4 of frames with a different iteration count or hard decision; posterior |d| / max(|LLR|, 1) <= 1e-4\",\n \"variants\": {\n",
           argv[1], g.Z, maxiter, snr, frames, (double)fails / frames, itsum / frames);
    for (int m = 1; m < NMODES; m++)
        printf("  \"%s\": {\"frames_iter_differs\": %ld, \"frames_hard_differs\": %ld, \"frames_differ\": %ld, \"frac_frames_differ\": %.3g, "
               "\"agreeing_frames_with_posterior_above_1e-4\": %ld, \"posterior_values_above_1e-4\": %.0f, \"worst_posterior_rel\": %.3g, \"meets_bar\": %s}%s\n",
               NAMES[m], diff_iter[m], diff_hard[m], diff_frame[m], (double)diff_frame[m] / frames, post_frames[m], post_vals[m], worst[m],
               ((double)diff_frame[m] / frames <= 1e-4 && post_frames[m] == 0) ? "true" : "false", m + 1 < NMODES ? "," : "");
    printf(" }\n}\n");
    return 0;
}
