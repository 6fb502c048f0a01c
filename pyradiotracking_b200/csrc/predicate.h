// predicate.h -- the "above" test of extract_signals (/root/reference/radiotracking/analyze.py:370-379, 391-396, 403-410):
// a cell belongs to a run unless it undershoots the absolute threshold or the SNR-vs-row-mean threshold,
//     above(p) = !(p < thr) && !(p / avg < snr)            (float32, correctly rounded division)
// Compiles for the device (rt_engine.cu) and for the host (tests/pred_host_check.cpp).
#pragma once

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

namespace rt {

RT_HD bool above_exact(float p, float thr, float avg, float snr) {
#if defined(__CUDA_ARCH__)
    return !(p < thr) && !(__fdiv_rn(p, avg) < snr);
#else
    return !(p < thr) && !(p / avg < snr);
#endif
}

// The same decision, bit for bit, at 3 instead of ~12 instructions for almost every cell: p * (1 / avg) is within 2 ulp of the
// correctly rounded quotient, so it decides unless it falls within 2^-20 of the threshold -- only then the exact division runs.
struct Pred {
    float thr, avg, inv_avg, snr, snr_lo, snr_hi;
    RT_HD Pred(float thr_, float avg_, float snr_)
        : thr(thr_), avg(avg_), inv_avg(1.0f / avg_), snr(snr_), snr_lo(snr_ * 0.99999905f), snr_hi(snr_ * 1.00000095f) {}
    RT_HD bool operator()(float p) const {
        if (p < thr) return false;
        const float r = p * inv_avg;
        if (r < snr_lo) return false;
        if (r > snr_hi) return true;
        return above_exact(p, thr, avg, snr);
    }
};

}  // namespace rt
