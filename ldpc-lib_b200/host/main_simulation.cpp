// `main simulation <scenario.jsonx> <result.jsonx> [marking file]` -- the reference's SNR-sweep driver
// (main_simulation.cpp:222-679) on the B200 engine.
//
// Kept: the input record layout (settings{...} + results = array of code records, :257-368), the shift
// reduction mod the lifting size with the special last bidiagonal column (:400-414), the SNR loop with
// reset_random() before every point (:477-518), the "good code" rule on the first SNR index whose FER is
// below TARGET_ERR (:543-571), the result descriptor and its keys (:588-616, appended to the result
// file), the progress table (:621-633), and the optional marking loop that re-labels the non-(-1)
// entries column by column from `{ data = array {...} min_modulo = n }` records (:66-119, 635-664).
// The girth / ACE / cycle spectrum of every simulated matrix (trace_matrix, :496) comes from ldpcb200_girth_spectrum
// (csrc/girth.cpp), once per matrix instead of once per SNR point; girth_, girth_ACE and girth_spectrum are written.  GF(q)
// codes (_q_mod > 2) are refused.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

#include "bp_simulation.h"
#include "decoders.h"
#include "settings.h"

#define DEFAULT_MARKING "skip"
#define TARGET_ERR 0.01
#define GTARGET 4

namespace {

// One `{ data = array { ... } min_modulo = n }` record of a marking file; false at end of file.
bool next_marking(FILE* fp, std::vector<int>& data, int& min_modulo)
{
    int ch;
    while ((ch = fgetc(fp)) != EOF && ch != '{') {}
    if (ch == EOF) return false;
    char word[128];
    for (int k = 0; k < 4; k++)
        if (fscanf(fp, "%120s", word) != 1) return false;      // data = array {
    data.clear();
    int v;
    while (fscanf(fp, "%d", &v) == 1) data.push_back(v);
    if (fscanf(fp, "%120s", word) != 1) return false;          // }
    if (fscanf(fp, "%120s", word) != 1) return false;          // min_modulo
    if (fscanf(fp, "%120s", word) != 1) return false;          // =
    if (fscanf(fp, "%d", &min_modulo) != 1) return false;
    if (fscanf(fp, "%120s", word) != 1) return false;          // }
    return true;
}

} // namespace

int main_simulation(int argc, char* argv[])
{
    if (argc != 3 && argc != 4) die("Expected arguments: <scenario file name> <result file name> [marking file]");

    settings f_scenario = settings::from_file(argv[1]);
    int num_experiments, num_frame_errors, modulation_type, permutation_type, permutation_block, permutation_inter;
    std::string error_name;
    double error_rate_threshold, good_code_multiple;
    f_scenario.select("settings/random_seed").cast_to(initial_random_seed);
    f_scenario.select("settings/num_codewords").cast_to(num_experiments);
    f_scenario.select("settings/error_blocks").cast_to(num_frame_errors);
    f_scenario.select("settings/error_minimization/name").cast_to(error_name);
    f_scenario.select("settings/error_minimization/threshold").cast_to(error_rate_threshold);
    f_scenario.select("settings/error_minimization/good_code_multiple").cast_to(good_code_multiple);
    if (error_name != "BER" && error_name != "FER") die("Unknown error name: '%s'", error_name.c_str());
    f_scenario.select("settings/modulation_type").cast_to(modulation_type);
    f_scenario.select("settings/permutation_type").cast_to(permutation_type);
    f_scenario.select("settings/permutation_block").cast_to(permutation_block);
    f_scenario.select("settings/permutation_inter").cast_to(permutation_inter);
    ensure_random_is_initialized();
    printf("scenario OK\n");

    std::vector<settings> input_codes;
    f_scenario.select("results").cast_to(input_codes);

    int rows = -1, columns = -1;
    for (int code_idx = 0; code_idx < (int)input_codes.size(); code_idx++) {
        settings const& rec = input_codes[code_idx];
        std::vector<double> snrs;
        std::vector<settings> logs;
        std::vector<int> column_weights;
        int q_mod, decoder_type, tailbite_length, punctured_blocks, num_iterations, config_index, code_index, matrix_index, girth;
        std::string mark_file;
        matrix<int> org_HM, coef;

        rec.select("_SNRs").cast_to(snrs);
        // The reference's `search` writes its result records WITHOUT `_q_mod` (main_good_code_search.cpp:383-399), so its
        // own `simulation` rejects them until the field is added by hand; the engine is binary only, so a
        // missing field means 2 and search -> simulation round-trips.
        if (rec.can_select("_q_mod")) rec.select("_q_mod").cast_to(q_mod);
        else q_mod = 2;
        rec.select("_decoder_type").cast_to(decoder_type);
        if (q_mod > 2) die("code #%d: _q_mod = %d: GF(q) codes are outside the B200 engine", code_idx, q_mod);
        rec.select("_lifting").cast_to(tailbite_length);
        rec.select("_punctured_blocks").cast_to(punctured_blocks);
        rec.select("_iterations").cast_to(num_iterations);
        rec.select("_marking").cast_to(mark_file);
        rec.select("code").cast_to(org_HM);
        rec.select("column_weights").cast_to(column_weights);
        rec.select("config_index").cast_to(config_index);
        rec.select("simulation_logs").cast_to(logs);
        rec.select("code_index").cast_to(code_index);
        rec.select("matrix_index").cast_to(matrix_index);
        rec.select("girth").cast_to(girth);
        if (argc == 4) mark_file = argv[3];

        if (rows == -1) { rows = org_HM.n_rows(); columns = org_HM.n_cols(); }
        else if (rows != org_HM.n_rows() || columns != org_HM.n_cols()) {
            std::cout << "Warning: unequal matrices in the input!" << std::endl;
            continue;
        }
        const int s_max = (int)snrs.size();
        const double bitrate = (double)(columns - rows) / (columns - punctured_blocks);
        std::vector<double> best_errors(s_max, error_rate_threshold), EsN0(s_max), BER(s_max), FER(s_max),
            best_BER(s_max), best_FER(s_max), org_BER(s_max), org_FER(s_max);
        std::vector<int> ACE(GTARGET, 0), girth_spectrum(GTARGET, 0);

        // shifts are reduced mod the lifting size; a zero in the last bidiagonal column becomes 1 (:400-414)
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < columns; j++)
                if (org_HM(i, j) > 0) {
                    int t = org_HM(i, j) % tailbite_length;
                    if (j == rows - 1) t = t == 0 ? 1 : t;
                    org_HM(i, j) = t;
                }
        matrix<int> current_HM = org_HM;

        std::vector<int> row_weights;
        for (int i = 0; i < rows; ++i) {
            unsigned count = 0;
            for (int j = 0; j < columns; ++j) count += org_HM(i, j) != -1;
            while (row_weights.size() <= count) row_weights.push_back(0);
            ++row_weights[count];
        }
        std::vector<std::vector<int>> row_weights_write;
        for (unsigned i = 0; i < row_weights.size(); ++i)
            if (row_weights[i] != 0) row_weights_write.push_back(std::vector<int>{(int)i, row_weights[i]});

        FILE* fp = NULL;
        bool mark_flag = mark_file != DEFAULT_MARKING;
        if (mark_flag && (fp = fopen(mark_file.c_str(), "rt")) == NULL) mark_flag = false;

        int mark_num = -1, best_mark = -1, best_snr_idx = s_max, curr_girth = 0;
        matrix<int> traced_HM;
        double best_err = 1.0, curr_BER = 1.0, curr_FER = 1.0;
        do {
            int s;
            for (s = 0; s < s_max; s++) {
                if (mark_num == -1) {
                    if (s == 0) printf("====================================================\n");
                    printf("code #%d, original matrix is being processed, SNR = %6.3f\n", code_idx, snrs[s]);
                } else
                    printf("code #%d, marked matrix #%d is being processed, SNR = %6.3f\n", code_idx, mark_num, snrs[s]);
                EsN0[s] = snrs[s] + 10.0 * log10(2.0 * bitrate);
                reset_random();                                   // all codes are tested with the same noise
                if (q_mod <= 2) {
                    // girth / ACE / cycle spectrum of the matrix about to be simulated (trace_matrix + show_matrix_property,
                    // :496-497).  The reference recomputes it for every SNR point (1.8 s each at E = 128); it only depends on
                    // the matrix, so it is computed when the matrix changes.
                    if (!(traced_HM == current_HM)) {
                        std::vector<int16_t> hd16((size_t)rows * columns);
                        for (int i = 0; i < rows; i++)
                            for (int j = 0; j < columns; j++) hd16[(size_t)i * columns + j] = (int16_t)current_HM(i, j);
                        if (ldpcb200_girth_spectrum(hd16.data(), rows, columns, tailbite_length, GTARGET, &curr_girth, ACE.data(), girth_spectrum.data()))
                            die("girth spectrum: bad matrix");
                        traced_HM = current_HM;
                    }
                    printf("girth: %3d, ACE: ", curr_girth);
                    for (int i = 0; i < GTARGET; i++) printf("%3d ", ACE[i]);
                    printf(",  SPEC: ");
                    for (int i = 0; i < GTARGET; i++) printf("%5d ", girth_spectrum[i]);
                }
                std::pair<double, double> result = bp_simulation(q_mod, current_HM, coef, 0, tailbite_length, num_iterations,
                                                                 num_frame_errors, num_experiments, snrs[s], best_errors[s], decoder_type,
                                                                 modulation_type, permutation_type, permutation_block, permutation_inter,
                                                                 punctured_blocks, 0);
                bp_simulation_stats const& st = bp_simulation_last_stats();
                printf("  frames %lld (decoded %lld on %d GPU%s), frame errors %lld, FER %.6g, BER %.6g, %.3f s\n", st.frames_counted,
                       st.frames_decoded, st.gpus, st.gpus == 1 ? "" : "s", st.frame_errors, result.second, result.first, st.seconds);
                if (result.first < 0 || result.second < 0) { curr_BER = 1.0; curr_FER = 1.0; }
                else { curr_BER = result.first; curr_FER = result.second; }
                BER[s] = curr_BER;
                FER[s] = curr_FER;
            }
            for (s = 0; s < s_max; s++)
                if (FER[s] < TARGET_ERR) break;
            if (s == s_max) s--;

            bool good_code = (s < best_snr_idx) || (s == best_snr_idx && FER[s] < best_err * good_code_multiple);
            if ((s < best_snr_idx) || (s == best_snr_idx && FER[s] < best_err)) {
                best_mark = mark_num; best_snr_idx = s; best_err = FER[s]; best_FER = FER; best_BER = BER;
            }
            if (mark_num == -1) { org_FER = FER; org_BER = BER; }

            if (good_code) {
                settings snr_curr;
                snr_curr.open("SNR_per_bit___").set(snrs);
                snr_curr.open("SNR_per_symbol").set(EsN0);
                snr_curr.open("BER").set(BER);
                snr_curr.open("FER").set(FER);
                std::vector<settings> snr_results(1, snr_curr);
                settings descriptor;
                descriptor.open("_decoder_name").set(DEC_FULL_NAME[decoder_type]);
                descriptor.open("_decoder_type").set(decoder_type);
                descriptor.open("_lifting").set(tailbite_length);
                descriptor.open("_SNRs").set(snrs);
                descriptor.open("_punctured_blocks").set(punctured_blocks);
                descriptor.open("_iterations").set(num_iterations);
                descriptor.open("_marking").set(DEFAULT_MARKING);
                descriptor.open("code_bitrate").set(bitrate);
                descriptor.open("code").set(current_HM);
                descriptor.open("column_weights").set(column_weights);
                descriptor.open("row_weights").set(row_weights_write);
                descriptor.open("girth").set(girth);
                descriptor.open("girth_").set(curr_girth);
                descriptor.open("girth_ACE").set(ACE);
                descriptor.open("girth_spectrum").set(girth_spectrum);
                descriptor.open("config_index").set(config_index);
                descriptor.open("matrix_index").set(matrix_index);
                descriptor.open("code_index").set(code_index);
                descriptor.open("simulation_logs").set(snr_results);
                descriptor.to_file(argv[2], true);
            }

            printf("\nbest mark: %d\n", best_mark);
            printf("        |             FER               |            BER \n");
            printf("  SNR   |    orig     best     curr     |   orig     best   curr\n");
            for (int k = 0; k < s_max; k++) {
                printf("%7.3f | ", snrs[k]);
                printf("%8.6f %8.6f %8.6f    | ", org_FER[k], best_FER[k], curr_FER);
                printf("%8.6f %8.6f %8.6f ", org_BER[k], best_BER[k], curr_BER);
                printf("\n");
            }
            printf("\n--------\n");

            mark_num++;
            if (mark_flag) {
                std::vector<int> data;
                int min_modulo = 0;
                if (!next_marking(fp, data, min_modulo)) mark_flag = false;
                else {
                    size_t k = 0;
                    for (int i = 0; i < columns && k < data.size(); i++)
                        for (int j = 0; j < rows && k < data.size(); j++)
                            if (current_HM(j, i) != -1) current_HM(j, i) = data[k++];
                }
            }
        } while (mark_flag);
        if (fp) fclose(fp);
    }
    printf("\n");
    return 0;
}
