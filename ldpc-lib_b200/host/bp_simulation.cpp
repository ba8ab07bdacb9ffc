// The Monte-Carlo frame loop of the reference's bp_simulation() (bp_simulation.cpp:591-824) restated
// over the engine's C ABI.
//
// What is the same: the signature; sigma / bitrate (bp_simulation.cpp:444-449, computed by
// ldpcb200_sigma); the all-zero codeword (:568); puncturing of the last `punctured_blocks` block columns
// (:697-710); the decoder's own return convention; per-frame error accounting (:731-743, 805-810:
// `nse` adds the information-bit errors of erroneous frames, `nde` counts erroneous frames, `nue` those
// the decoder reported as converged); both stop rules, applied IN FRAME ORDER (:591 and :820); the
// result (nse / experiment / (n - r), nde / experiment) (:840).
//
// What differs, by design: frames are decoded in rounds of thousands on the GPU(s) (ldpcb200_simulate)
// and the per-frame records of a round are then scanned in order, so the stop rules cut at exactly
// the frame where the reference's sequential loop would stop -- results do not depend on the round size
// or on the number of GPUs (frames after the stop point were decoded for nothing and are discarded).
// Noise comes from the engine's counter-based generator keyed by (random_seed, frame index), not from
// the reference's global std::mt19937 stream, so FER / BER agree statistically, not sample by sample.
// QAM-16/64/256 use the intended channel r = s + sigmaQAM n with Demodulate(m = log2 Q); the reference's
// own wiring of that path is broken.
//
// Bit interleavers (permutation_type 1-4, direct_inverse_perm.cpp) are applied on the device: the index tables are
// built once per call (ldpcb200_interleaver_tables) and the decoder's first load gathers through them.
// Out of scope here and refused loudly: GF(q) codes (q_mod > 2) -- SURVEY.md §8f "next".
#include "bp_simulation.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "decoders.h"

namespace {

bp_simulation_stats g_stats;

std::vector<int> devices_from_env()
{
    // LDPCB200_DEVICES = "all" | "0,2,3" ; default: the current device only
    std::vector<int> devs;
    const char* e = getenv("LDPCB200_DEVICES");
    if (!e || !*e) { devs.push_back(-1); return devs; }
    std::string s(e);
    if (s == "all") {
        // probe by creating handles until one fails is wasteful; ask for up to 64 ordinals and let create() reject
        for (int d = 0; d < 64; d++) devs.push_back(d);
        return devs;
    }
    size_t i = 0;
    while (i < s.size()) {
        size_t j = s.find(',', i);
        if (j == std::string::npos) j = s.size();
        devs.push_back(atoi(s.substr(i, j - i).c_str()));
        i = j + 1;
    }
    return devs;
}

unsigned long long g_next_frame = 0, g_epoch = ~0ull;
volatile int g_interrupt = 0;                      // bp_simulation_request_interrupt()

} // namespace

bp_simulation_stats const& bp_simulation_last_stats() { return g_stats; }
void bp_simulation_request_interrupt() { g_interrupt = 1; }

std::pair<double, double> bp_simulation(
    int q_mod, matrix<int> const& H, matrix<int>& /*coef_matrix*/, int /*ncols2convert*/, int tailbite_length,
    int max_iterations, int n_frame_errors, int n_experiments, double snr, double reference_frame_error,
    int decoder_type, int modulation_type, int permutation_type, int permutation_block, int permutation_inter,
    int punctured_blocks, int show_process)
{
    const int b = H.n_rows(), c = H.n_cols(), M = tailbite_length;
    const int r = b * M, n = c * M;
    if (q_mod != 2) die("bp_simulation: q_mod = %d: GF(q) codes are outside the B200 engine (binary decoders only)", q_mod);
    if (modulation_type < MODULATION_SKIP || modulation_type > MODULATION_QAM256) die("Unknown modulation type: %d", modulation_type);
    if (permutation_type < 0 || permutation_type > 4) die("Unknown permutation type: %d", permutation_type);
    switch (decoder_type) {
    case BP_DEC: case SP_DEC: case ASP_DEC: case MS_DEC: case IMS_DEC: case IASP_DEC: case TASP_DEC: case LMS_DEC: case LCHE_DEC: break;
    default: die("Unknown decoder type: %d", decoder_type);
    }

    std::vector<int16_t> hd((size_t)b * c);
    for (int i = 0; i < b; i++)
        for (int j = 0; j < c; j++) hd[(size_t)i * c + j] = (int16_t)H(i, j);            // bp_simulation.cpp:357-362

    // one engine handle per GPU
    ldpcb200_params p;
    ldpcb200_default_params(&p);
    const char* pe = getenv("LDPCB200_PRECISION");
    // LMS_DEC / MS_DEC default to the fp32 throughput kernels in the simulation loop (same decisions and
    // iteration counts as the double arithmetic on >= 99.99 % of frames); LDPCB200_PRECISION=64 forces double
    p.precision = (decoder_type == LMS_DEC || decoder_type == MS_DEC) ? 32 : 64;
    if (pe && atoi(pe) == 64) p.precision = 64;
    // compile a code-specialised kernel for this matrix unless told not to (LDPCB200_JIT=0); cached per process
    const char* je = getenv("LDPCB200_JIT");
    p.use_fast = (je && atoi(je) == 0) ? 1 : 2;
    // The engine handles (one per GPU) are kept between calls: an SNR sweep calls bp_simulation() once per point with the
    // same matrix and decoder, and opening a handle (streams, tables, launch plan, possibly a run-time compilation)
    // costs more than a short point.  A different matrix / decoder / lifting replaces the cached set.
    static std::vector<ldpcb200_handle> eng;
    static std::vector<int16_t> eng_hd;
    static int eng_key[6] = { -1, -1, -1, -1, -1, -1 };
    static std::string eng_devs;
    const char* de = getenv("LDPCB200_DEVICES");
    const int key[6] = { b, c, M, decoder_type, p.precision, p.use_fast };
    if (eng.empty() || eng_hd != hd || memcmp(eng_key, key, sizeof key) != 0 || eng_devs != (de ? de : "")) {
        for (auto h : eng) ldpcb200_destroy(h);
        eng.clear();
        std::vector<int> devs = devices_from_env();
        for (size_t k = 0; k < devs.size(); k++) {
            ldpcb200_handle h = NULL;
            p.device = devs[k];
            int rc = ldpcb200_create(hd.data(), b, c, M, decoder_type, &p, &h);
            if (rc) {
                // "all" asks for ordinals until the engine says there is no such device; any other failure is an error
                const bool past_the_last = rc == LDPCB200_EINVAL && strstr(ldpcb200_last_error(), "out of range") != NULL;
                if (de && std::string(de) == "all" && !eng.empty() && past_the_last) break;
                die("bp_simulation: cannot open the decoder on device %d: %s", devs[k], ldpcb200_last_error());
            }
            eng.push_back(h);
        }
        eng_hd = hd; memcpy(eng_key, key, sizeof key); eng_devs = de ? de : "";
    }
    const int G = (int)eng.size();

    // Bit interleaver (bp_simulation.cpp:417-425): the reference interleaves the (all-zero, :567) codeword, modulates, adds
    // noise, demodulates and de-interleaves the LLRs (:573, :684) before puncturing (:697-710).  The engine does the
    // de-interleaving inside the decoder's first load from the same index tables (csrc/interleaver.cpp).
    if (permutation_type != 0) {
        std::vector<int32_t> dir((size_t)n), inv((size_t)n);
        int rc = ldpcb200_interleaver_tables(hd.data(), b, c, M, modulation_type, permutation_type, permutation_block, permutation_inter, dir.data(), inv.data());
        if (rc) die("bp_simulation: permutation_type = %d (block %d, inter %d) does not define a permutation of the %d code bits%s",
                    permutation_type, permutation_block, permutation_inter, n, rc == LDPCB200_EUNSUPPORTED ? " (the reference would read stale memory)" : "");
        for (auto h : eng)
            if (ldpcb200_set_interleaver(h, dir.data(), inv.data())) die("bp_simulation: cannot attach the interleaver: %s", ldpcb200_last_error());
    } else {
        for (auto h : eng) ldpcb200_set_interleaver(h, NULL, NULL);
    }

    ensure_random_is_initialized();
#ifdef LDPCB200_WITH_REFERENCE_HEADERS
    g_next_frame = 0;       // the reference keeps no visible record of reset_random(): every call starts the stream over
    (void)g_epoch;
#else
    if (g_epoch != current_noise_epoch()) { g_epoch = current_noise_epoch(); g_next_frame = 0; }   // reset_random() restarts the stream
#endif

    long long nse = 0, nue = 0, nde = 0, experiment = 0, decoded = 0;
    double gpu_ms = 0;
    bool stop = false;
    const long long limit = (long long)n_experiments + 1;       // `experiment <= n_experiments` is tested before the increment (:591-594)
    long long per_gpu = 4096;
    auto t0 = std::chrono::steady_clock::now();

    std::vector<std::vector<uint32_t>> rec(G);
    std::vector<int> rcs(G);
    std::vector<float> ms(G);
    std::vector<std::string> errs(G);                           // ldpcb200_last_error() is per thread: taken inside the worker
    while (!stop && nde < n_frame_errors && experiment < limit) {
        if (g_interrupt) { g_interrupt = 0; return std::make_pair(-1.0, -1.0); }            // :825-829
        const long long want = std::min<long long>(limit - experiment, per_gpu * G);
        // GPU g decodes frames [base + off[g], base + off[g] + cnt[g]) of the stream
        std::vector<long long> cnt(G), off(G);
        long long acc = 0;
        for (int g = 0; g < G; g++) { cnt[g] = want / G + (g < want % G ? 1 : 0); off[g] = acc; acc += cnt[g]; }
        auto run = [&](int g) {
            rcs[g] = 0; ms[g] = 0;
            if (cnt[g] == 0) return;
            rec[g].resize((size_t)cnt[g]);
            ldpcb200_sim_params sp;
            memset(&sp, 0, sizeof sp);
            sp.snr_db = snr; sp.modulation = modulation_type; sp.punctured_blocks = punctured_blocks;
            sp.max_iterations = max_iterations; sp.seed = (uint64_t)(uint32_t)initial_random_seed; sp.stream = 0;
            sp.first_frame = g_next_frame + (uint64_t)off[g]; sp.n_frames = (uint32_t)cnt[g];
            ldpcb200_counters co;
            rcs[g] = ldpcb200_simulate(eng[g], &sp, &co, rec[g].data());
            if (!rcs[g]) ldpcb200_last_kernel_ms(eng[g], &ms[g], NULL);
            else errs[g] = ldpcb200_last_error();
        };
        if (G == 1) run(0);
        else {
            std::vector<std::thread> th;
            for (int g = 0; g < G; g++) th.emplace_back(run, g);
            for (auto& t : th) t.join();
        }
        for (int g = 0; g < G; g++)
            if (rcs[g]) die("bp_simulation: simulate failed on GPU %d (error %d): %s", g, rcs[g], errs[g].c_str());
        gpu_ms += *std::max_element(ms.begin(), ms.end());
        decoded += want;
        // the reference's loop body after the decoder call, frame by frame (:731-743, 805-823)
        for (int g = 0; g < G && !stop; g++)
            for (long long k = 0; k < cnt[g]; k++) {
                if (!(nde < n_frame_errors && experiment < limit)) { stop = true; break; }
                ++experiment;
                const uint32_t w = rec[g][(size_t)k];
                if (w & 0x80000000u) {
                    nse += w & 0xFFFFFFu;
                    ++nde;
                    if (w & 0x40000000u) ++nue;
                    if (show_process)
                        printf("SNR=%5.3lf,step=%4lld,s_ers=%lld,f_ers=%lld,u_ers=%lld,BER=%5.3le,FER=%5.3le\n", snr, experiment, nse, nde, nue,
                               (double)nse / experiment / (n - r), (double)nde / experiment);
                    if (nde >= 10 && (double)nde / experiment > 2.5 * reference_frame_error) { stop = true; break; }
                }
            }
        g_next_frame += (unsigned long long)want;
        per_gpu = std::min<long long>(per_gpu * 4, 1 << 18);
    }
    g_stats.frames_counted = experiment; g_stats.frames_decoded = decoded; g_stats.frame_errors = nde;
    g_stats.info_bit_errors = nse; g_stats.undetected = nue; g_stats.gpu_ms = gpu_ms; g_stats.gpus = G;
    g_stats.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (experiment == 0) return std::make_pair(0.0, 0.0);
    return std::make_pair((double)nse / experiment / (n - r), (double)nde / experiment);          // :840
}
