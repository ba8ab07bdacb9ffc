// CPU check of ldpc-lib_b200/csrc/fastmath64.cuh (the branch-free exp / log of the BP_DEC throughput kernel) against libm.
//   g++ -O2 -o fastmath_check fastmath_check.cpp && ./fastmath_check     exit code 0 = within 2.3e-16 relative, special values right
#include <cmath>
#include <cstdio>
#include <random>
#include "../../ldpc-lib_b200/csrc/fastmath64.cuh"
using namespace ldpcb200;

int main()
{
    std::mt19937_64 g(1);
    double maxe = 0, maxl = 0, maxt = 0;
    int bad = 0;
    std::uniform_real_distribution<double> ue(-250, 250), ul(-40, 40);
    for (int i = 0; i < 4000000; i++) {
        const double x = ue(g), a = fx_exp(x), b = std::exp(x);
        maxe = std::fmax(maxe, std::fabs(a - b) / b);
        maxt = std::fmax(maxt, std::fabs(fx_exp_tab(x, FX_T32) - b) / b);
        const double v = std::exp(ul(g)), c = fx_log(v), d = std::log(v);
        if (std::fabs(d) > 1e-3) maxl = std::fmax(maxl, std::fabs(c - d) / std::fabs(d));
        else if (std::fabs(c - d) > 4e-19 + 3e-16 * std::fabs(d)) bad++;
    }
    for (int i = 0; i < 2000000; i++) {                                  // arguments next to 1: |tanh| of a saturated message
        const double v = 1.0 - std::ldexp((double)(g() >> 11), -53 - (int)(g() % 40)), c = fx_log(v), d = std::log(v);
        if (std::fabs(c - d) > 2.3e-16 * std::fabs(d) + 1e-300) bad++;
    }
    // fx_log_ratio(n, q) against log in long double of the exact quotient: the way bpsp4_kernel calls it (n = T + S, q = T - S)
    double maxr = 0;
    std::uniform_real_distribution<double> u01(0, 1), ux(-38, 0);
    for (int i = 0; i < 4000000; i++) {
        const double T = (i & 1) ? u01(g) : std::exp(ux(g)), S = T * ((i & 2) ? u01(g) : 1 - std::exp(ux(g)));
        const double n = T + S, q = T - S;
        if (!(q > 0) || q < n * 1e-9) continue;
        const double c = fx_log_ratio(n, q);
        const long double d = logl((long double)n / (long double)q);
        const double err = (double)fabsl(c - d);
        if (d > 1e-3L) maxr = std::fmax(maxr, err / (double)d);
        else if (err > 4e-19 + 3e-16 * (double)d) bad++;
    }
    bad += !(fx_log_ratio(3.0, 3.0) == 0.0) + !(fx_log_ratio(1e-200, 1e-200) == 0.0);
    bad += !(fx_log(0.0) == -INFINITY) + !(fx_log(INFINITY) == INFINITY) + !std::isnan(fx_log(-1.0)) + !std::isnan(fx_log(NAN)) + !(fx_log(1.0) == 0.0);
    bad += !(fx_exp(-INFINITY) == 0.0) + !std::isnan(fx_exp(NAN)) + !(fx_exp(-800.0) == 0.0) + !(fx_exp(0.0) == 1.0);
    bad += !(fx_exp_tab(0.0, FX_T32) == 1.0);
    printf("max relative error: exp %.3g, exp with table %.3g, log %.3g, log of a ratio %.3g; failures %d\n", maxe, maxt, maxl, maxr, bad);
    return (maxe <= 2.3e-16 && maxt <= 3.4e-16 && maxl <= 2.3e-16 && maxr <= 3.4e-16 && bad == 0) ? 0 : 1;
}
