// CPU emulation of the data flow of spectro_reg256_v8 (pyradiotracking_b200/csrc/spectro256.cuh) for
// ONE 256-sample segment, using the very same packed-complex FFT header (host flavour).  The build
// container has no GPU; this lets `pytest -m "not gpu"` check the index algebra (16x16 Cooley-Tukey,
// twiddles, byte->float trick, stored bin permutation) against numpy.  The front end is v8's: bytes -> fp16 pair
// (0x6400 | b), segment byte sums from the packed words, -(1024 + mean) by one FMA, widening add; every value is
// compared bit for bit with v7's 0x4700bb00 / FADD2 front end (exit code 2 on a difference).
//   stdin : 512 raw bytes, then 256 float32 window values
//   stdout: 256 float32 power values in FFT bin order
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../pyradiotracking_b200/csrc/fft_cpk.cuh"

static float magic_byte(unsigned b) {          // 0x4700bb00 == 32768 + b
    uint32_t u = 0x47000000u | (b << 8);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

static float half_bits_to_float(unsigned h) {   // IEEE binary16 -> binary32 (normal numbers only: 0x6400 | b is 1024 + b)
    const unsigned e = (h >> 10) & 31u, m = h & 1023u;
    return std::ldexp((float)(1024u + m), (int)e - 25);
}
static float bits_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

int main() {
    std::vector<uint8_t> raw(512);
    std::vector<float> win(256);
    if (fread(raw.data(), 1, 512, stdin) != 512) return 1;
    if (fread(win.data(), 4, 256, stdin) != 256) return 1;
    unsigned sI = 0, sQ = 0;
    for (int i = 0; i < 256; ++i) { sI += raw[2 * i]; sQ += raw[2 * i + 1]; }
    const rt::cpk c = rt::c_make(32768.f + (float)sI * 0.00390625f, 32768.f + (float)sQ * 0.00390625f);
    // v8: per "thread" j the 16 packed words hv = (0x6400 | I) | (0x6400 | Q) << 16; their integer sum minus 16 x 0x64006400 (mod 2^32)
    // is sum I + 65536 sum Q of the thread, the sum over the 16 threads (REDUX on the device) the segment's totals
    uint32_t tot = 0;
    for (int j = 0; j < 16; ++j) {
        uint32_t t = 0;
        for (int n1 = 0; n1 < 16; ++n1) {
            const int smp = 16 * n1 + j;
            t += (0x6400u | raw[2 * smp]) | ((0x6400u | raw[2 * smp + 1]) << 16);
        }
        tot += t - 0x40064000u;
    }
    if ((tot & 0xffffu) != sI || (tot >> 16) != sQ) return 2;
    const float ncI = std::fmaf(bits_float(0x4B000000u | (tot & 0xffffu)), -0.00390625f, 31744.f);   // -(1024 + mean_I), exact
    const float ncQ = std::fmaf(bits_float(0x4B000000u | (tot >> 16)), -0.00390625f, 31744.f);
    if ((double)ncI != -(1024.0 + sI / 256.0) || (double)ncQ != -(1024.0 + sQ / 256.0)) return 2;
    static rt::cpk xch[16][16];
    float out_pos[256];
    for (int j = 0; j < 16; ++j) {              // "thread" j: column n2 = j
        rt::cpk v[16];
        for (int n1 = 0; n1 < 16; ++n1) {
            const int smp = 16 * n1 + j;
            const rt::cpk old = rt::c_sub(rt::c_make(magic_byte(raw[2 * smp]), magic_byte(raw[2 * smp + 1])), c);      // v7's front end
            const float re = half_bits_to_float(0x6400u | raw[2 * smp]) + ncI, im = half_bits_to_float(0x6400u | raw[2 * smp + 1]) + ncQ;   // FHADD
            if (re != rt::c_re(old) || im != rt::c_im(old)) return 2;
            v[n1] = rt::c_scale(rt::c_make(re, im), win[smp]);
        }
        rt::cdft16(v);
        for (int k1 = 1; k1 < 16; ++k1) {
            const double ang = -2.0 * M_PI * (double)((j * k1) & 255) / 256.0;
            v[k1] = rt::c_mul(v[k1], (float)std::cos(ang), (float)std::sin(ang));
        }
        for (int k1 = 0; k1 < 16; ++k1) xch[k1][j] = v[k1];
    }
    for (int j = 0; j < 16; ++j) {              // "thread" j: row k1 = j
        rt::cpk v[16];
        for (int c2 = 0; c2 < 16; ++c2) v[c2] = xch[j][c2];
        rt::cdft16(v);
        for (int k2 = 0; k2 < 16; ++k2) out_pos[16 * j + k2] = rt::c_re(v[k2]) * rt::c_re(v[k2]) + rt::c_im(v[k2]) * rt::c_im(v[k2]);
    }
    float out[256];
    for (int fi = 0; fi < 256; ++fi) out[fi] = out_pos[((fi & 15) << 4) | (fi >> 4)];
    fwrite(out, 4, 256, stdout);
    return 0;
}
