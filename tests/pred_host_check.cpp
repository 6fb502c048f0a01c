// Host check of pyradiotracking_b200/csrc/predicate.h: the cheap predicate (Pred) must take the decision of the exact one
// (analyze.py:370-379) for every float, in particular for powers within a few ulp of thr and of snr * avg.
//   stdout: "<cases> <mismatches>"
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>

#include "../pyradiotracking_b200/csrc/predicate.h"

static float nudge(float x, int ulps) {
    int32_t u;
    std::memcpy(&u, &x, 4);
    u += ulps;
    std::memcpy(&x, &u, 4);
    return x;
}

int main() {
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> e(-14.0, -6.0), se(-1.0, 2.0), u01(0.0, 1.0);
    long cases = 0, bad = 0;
    for (int it = 0; it < 200000; ++it) {
        const float avg = (float)std::pow(10.0, e(rng));
        const float snr = it % 97 == 0 ? 0.f : (float)std::pow(10.0, se(rng));
        const float thr = (float)std::pow(10.0, e(rng));
        const rt::Pred pred(thr, avg, snr);
        const float edge = snr * avg;
        for (int k = -24; k <= 24; ++k) {
            const float ps[3] = {nudge(edge, k), nudge(thr, k), (float)(edge * (1.0 + (u01(rng) - 0.5) * 4e-6))};
            for (float p : ps) {
                ++cases;
                if (pred(p) != rt::above_exact(p, thr, avg, snr)) ++bad;
            }
        }
        const float far[4] = {0.f, avg * 1e-3f, edge * 10.f, INFINITY};
        for (float p : far) {
            ++cases;
            if (pred(p) != rt::above_exact(p, thr, avg, snr)) ++bad;
        }
    }
    // degenerate rows: zero / infinite / NaN row mean
    const float weird[4] = {0.f, INFINITY, NAN, 1e-38f};
    for (float avg : weird)
        for (float p : {0.f, 1e-12f, 1.f, INFINITY, NAN}) {
            const rt::Pred pred(1e-9f, avg, 3.1622777f);
            ++cases;
            if (pred(p) != rt::above_exact(p, 1e-9f, avg, 3.1622777f)) ++bad;
        }
    std::printf("%ld %ld\n", cases, bad);
    return bad ? 1 : 0;
}
