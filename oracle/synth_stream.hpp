// oracle/synth_stream.hpp -- TEST / BENCH INFRASTRUCTURE ONLY (part of the CPU oracle).
//
// (1) The synthetic design of the benchmarks (README recipe, /root/reference/README.md:195-201:
//     X_ij ~ N(mean, sd^2), beta* = nsig leading U(0,1) values, y = X beta* + N(0, noise^2)),
//     produced from the same counter-based stream as the CUDA generator of the library
//     (Philox4x32-10; key = seed; counter = (column, global row group) for X) and with the same
//     float arithmetic, operation by operation: the logarithm and sin/cos of the Box-Muller
//     transform are fixed polynomial evaluations built from IEEE +, *, /, sqrt and fma only, so
//     CPU and GPU produce BIT-IDENTICAL matrices (tests/test_gpu_kernels.py checks this) and
//     the CPU arm of the bench fits the very design the GPU arm fits.
//     Arithmetic spec (all float, every operation rounded to nearest):
//       u01(k)   = (float(k) + 0.5f) * 2^-32
//       ln(u)    : u = m 2^e, m in [1,2) from the bit pattern; if m > 1.41421354f: m *= 0.5, e += 1;
//                  f = m - 1; s = f / (2 + f); z = s s;
//                  P = fma(z, c11, c9); P = fma(z, P, c7); P = fma(z, P, c5); P = fma(z, P, c3)   (c_k = 2/k)
//                  ln = fma(float(e), LN2, fma(s z, P, 2 s))
//       r        = sqrt(-2 ln(max(u1, 1e-12f)))
//       t = 2 u2; q = rint(2 t); f = fma(q, -0.5f, t); w = f f
//       sinpi(f) = f * fma(w, fma(w, fma(w, fma(w, S9, S7), S5), S3), S1)
//       cospi(f) = fma(w, fma(w, fma(w, fma(w, fma(w, C10, C8), C6), C4), C2), 1)
//       (sin, cos)(pi t) by the quadrant q & 3;  n0 = r cos, n1 = r sin;  x = mean + sd * n
// (2) A streamed full-size tall lasso fit: DataStd -> X'y -> Gram -> path, with X produced in row
//     chunks (it never has to exist as a whole: 40 GB at n = 1e6, p = 1e4), following
//     /root/reference/src/DataStd.h:128-150 (flag 3), ADMMLassoTall.h:164-216 and Lasso.cpp:78-124.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace oracle {
namespace synth {

struct U4 { uint32_t x, y, z, w; };
constexpr uint32_t STREAM_X = 0u, STREAM_NOISE = 0x5EEDu, STREAM_BETA = 0xBE7Au;   // stream ids in the counter's 4th word

inline U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return U4{c0, c1, c2, c3};
}

inline float u01(uint32_t k) { return ((float)k + 0.5f) * 2.3283064365386963e-10f; }

inline float ln_spec(float u)
{
    uint32_t b;
    std::memcpy(&b, &u, 4);
    int e = (int)(b >> 23) - 127;
    b = (b & 0x007fffffu) | 0x3f800000u;
    float m;
    std::memcpy(&m, &b, 4);
    if (m > 1.41421354f) { m = m * 0.5f; e += 1; }
    const float f = m - 1.0f;
    const float s = f / (2.0f + f);
    const float z = s * s;
    float P = std::fmaf(z, 0.181818187f, 0.222222224f);
    P = std::fmaf(z, P, 0.285714298f);
    P = std::fmaf(z, P, 0.400000006f);
    P = std::fmaf(z, P, 0.666666687f);
    const float lm = std::fmaf(s * z, P, 2.0f * s);
    return std::fmaf((float)e, 0.693147182f, lm);
}

inline void sincospi_spec(float t, float& sn, float& cs)     // sin(pi t), cos(pi t), t in [0, 2]
{
    const float q = std::nearbyintf(2.0f * t);               // round half to even (default mode), as cvt.rni / rintf
    const float f = std::fmaf(q, -0.5f, t);
    const float w = f * f;
    float S = std::fmaf(w, 0.0821458846f, -0.599264503f);
    S = std::fmaf(w, S, 2.55016398f);
    S = std::fmaf(w, S, -5.16771269f);
    S = std::fmaf(w, S, 3.14159274f);
    const float sp = f * S;
    float Cc = std::fmaf(w, -0.0258068908f, 0.235330626f);
    Cc = std::fmaf(w, Cc, -1.33526278f);
    Cc = std::fmaf(w, Cc, 4.05871201f);
    Cc = std::fmaf(w, Cc, -4.93480206f);
    const float cp = std::fmaf(w, Cc, 1.0f);
    switch (((int)q) & 3) {
    case 0: sn = sp; cs = cp; break;
    case 1: sn = cp; cs = -sp; break;
    case 2: sn = -sp; cs = -cp; break;
    default: sn = -cp; cs = sp; break;
    }
}

inline void box_muller(uint32_t a, uint32_t b, float& n0, float& n1)
{
    const float u1 = std::fmax(u01(a), 1e-12f), u2 = u01(b);
    const float r = std::sqrt(-2.0f * ln_spec(u1));
    float sn, cs;
    sincospi_spec(2.0f * u2, sn, cs);
    n0 = r * cs; n1 = r * sn;
}

// Branch-free lane forms of the same arithmetic (identical results, written so the compiler can vectorise
// them): the Philox rounds and both Box-Muller transforms of a block of row groups of one column.
constexpr int BLK = 64;

inline void philox_block(const uint32_t col_lo, const uint32_t col_hi, const int64_t g0, const int nb, const uint32_t k0, const uint32_t k1,
                         uint32_t* __restrict__ o0, uint32_t* __restrict__ o1, uint32_t* __restrict__ o2, uint32_t* __restrict__ o3)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma omp simd
    for (int i = 0; i < nb; i++) {
        const int64_t g = g0 + i;
        uint32_t c0 = col_lo, c1 = col_hi, c2 = (uint32_t)g, c3 = STREAM_X ^ (uint32_t)(g >> 32);
        uint32_t a0 = k0, a1 = k1;
        for (int r = 0; r < 10; r++) {
            const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a0, n1 = (uint32_t)p1;
            const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ a1, n3 = (uint32_t)p0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a0 += W0; a1 += W1;
        }
        o0[i] = c0; o1[i] = c1; o2[i] = c2; o3[i] = c3;
    }
}

inline void box_muller_lanes(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, const int nb,
                             float* __restrict__ n0, float* __restrict__ n1)
{
#pragma omp simd
    for (int i = 0; i < nb; i++) {
        // ln(u1)
        // float(k) for an unsigned 32-bit k, rounded to nearest: both halves convert exactly and their sum is
        // rounded once (AVX2 has no unsigned conversion; the signed one vectorises)
        const float ka = (float)(int)(a[i] >> 16) * 65536.0f + (float)(int)(a[i] & 0xffffu);
        const float kb = (float)(int)(b[i] >> 16) * 65536.0f + (float)(int)(b[i] & 0xffffu);
        float u1 = (ka + 0.5f) * 2.3283064365386963e-10f;
        u1 = u1 < 1e-12f ? 1e-12f : u1;
        const float u2 = (kb + 0.5f) * 2.3283064365386963e-10f;
        uint32_t bits = __builtin_bit_cast(uint32_t, u1);
        int e = (int)(bits >> 23) - 127;
        bits = (bits & 0x007fffffu) | 0x3f800000u;
        float m = __builtin_bit_cast(float, bits);
        const bool big = m > 1.41421354f;
        m = big ? m * 0.5f : m;
        e = big ? e + 1 : e;
        const float f = m - 1.0f;
        const float s = f / (2.0f + f);
        const float z = s * s;
        float P = std::fmaf(z, 0.181818187f, 0.222222224f);
        P = std::fmaf(z, P, 0.285714298f);
        P = std::fmaf(z, P, 0.400000006f);
        P = std::fmaf(z, P, 0.666666687f);
        const float lm = std::fmaf(s * z, P, 2.0f * s);
        const float ln = std::fmaf((float)e, 0.693147182f, lm);
        const float r = std::sqrt(-2.0f * ln);
        // sin / cos(2 pi u2)
        const float t = 2.0f * u2;
        const float q = std::nearbyintf(2.0f * t);
        const float fr = std::fmaf(q, -0.5f, t);
        const float w = fr * fr;
        float S = std::fmaf(w, 0.0821458846f, -0.599264503f);
        S = std::fmaf(w, S, 2.55016398f);
        S = std::fmaf(w, S, -5.16771269f);
        S = std::fmaf(w, S, 3.14159274f);
        const float sp = fr * S;
        float Cc = std::fmaf(w, -0.0258068908f, 0.235330626f);
        Cc = std::fmaf(w, Cc, -1.33526278f);
        Cc = std::fmaf(w, Cc, 4.05871201f);
        Cc = std::fmaf(w, Cc, -4.93480206f);
        const float cp = std::fmaf(w, Cc, 1.0f);
        const int qi = (int)q;
        const bool sw = (qi & 1) != 0;
        float sn = sw ? cp : sp, cs = sw ? sp : cp;
        sn = (qi & 2) ? -sn : sn;
        cs = ((qi + 1) & 2) ? -cs : cs;
        n0[i] = r * cs; n1[i] = r * sn;
    }
}

// X: rows [row0, row0 + nrows) of the global design, column-major with leading dimension ld >= nrows
inline void fill_x(float* X, int64_t nrows, int64_t ld, int64_t p, int64_t row0, uint64_t seed, float mean_x, float sd_x)
{
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int64_t g_first = row0 / 4, g_last = (row0 + nrows + 3) / 4;
#pragma omp parallel for schedule(static)
    for (int64_t col = 0; col < p; col++) {
        float* dst = X + col * ld;
        alignas(64) uint32_t o0[BLK], o1[BLK], o2[BLK], o3[BLK];
        alignas(64) float v0[BLK], v1[BLK], v2[BLK], v3[BLK];
        for (int64_t g0 = g_first; g0 < g_last; g0 += BLK) {
            const int nb = (int)std::min<int64_t>(BLK, g_last - g0);
            philox_block((uint32_t)col, (uint32_t)(col >> 32), g0, nb, k0, k1, o0, o1, o2, o3);
            box_muller_lanes(o0, o1, nb, v0, v1);
            box_muller_lanes(o2, o3, nb, v2, v3);
            const int64_t lr0 = g0 * 4 - row0;
            if (lr0 >= 0 && lr0 + 4 * (int64_t)nb <= nrows) {
                float* d = dst + lr0;
                for (int i = 0; i < nb; i++) {
                    d[4 * i] = mean_x + sd_x * v0[i]; d[4 * i + 1] = mean_x + sd_x * v1[i];
                    d[4 * i + 2] = mean_x + sd_x * v2[i]; d[4 * i + 3] = mean_x + sd_x * v3[i];
                }
            } else {
                for (int i = 0; i < nb; i++) {
                    const float v[4] = {v0[i], v1[i], v2[i], v3[i]};
                    for (int e = 0; e < 4; e++) {
                        const int64_t lr = lr0 + 4 * i + e;
                        if (lr >= 0 && lr < nrows) dst[lr] = mean_x + sd_x * v[e];
                    }
                }
            }
        }
    }
}

inline void beta_star(float* beta, int nsig, uint64_t seed)
{
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int j = 0; j < nsig; j++) beta[j] = u01(philox4x32_10((uint32_t)j, 0u, 0u, STREAM_BETA, k0, k1).x);
}

// y for the same rows, from X already in memory (first min(nsig, p) columns)
inline void fill_y(const float* X, int64_t nrows, int64_t ld, int64_t p, int64_t row0, uint64_t seed, int nsig, float noise, float* y)
{
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    std::vector<float> beta(std::max(nsig, 1));
    beta_star(beta.data(), nsig, seed);
    const int ns = (int)std::min<int64_t>(nsig, p);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nrows; i++) {
        const int64_t gr = row0 + i;
        const U4 r = philox4x32_10((uint32_t)(gr >> 1), (uint32_t)(gr >> 33), 0u, STREAM_NOISE, k0, k1);
        float e0, e1;
        box_muller(r.x, r.y, e0, e1);
        float acc = 0.f;
        for (int j = 0; j < ns; j++) acc = std::fmaf(X[i + (int64_t)j * ld], beta[j], acc);
        y[i] = acc + noise * ((gr & 1) ? e1 : e0);
    }
}

}  // namespace synth
}  // namespace oracle
