// pmb_wnrng.h -- the RANLUX (ranlxd1) generator, the seed tables and the per-column mode arithmetic of
// the white-noise generator, compiled for host and device: pmb_whitenoise.cu runs one column per GPU
// thread with these routines, tests/harness/host_harness.cpp walks the columns serially with the
// same routines so that the CPU-only test-suite can compare them bit for bit with the oracle.
//
// Reference: pmesh/gsl/ranlxd.c:36-245 (generator), pmesh/_whitenoise_imp.c:21-52 (SAMPLE, SETSEED),
// pmesh/_whitenoise_generics.h:73-94 (seed spiral), :106-232 (column fill).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "pmb_window.h"   /* PMB_HD */

#define WN_ONE_BIT (1.0 / 281474976710656.0)   /* 2^-48 */

// ---- generator, shared by host (master stream) and device (column streams) -----------------------
struct WnRng {
    double x[12];   // x[0] is the next word to hand out / the oldest word of the ring
    double c;       // borrow: 0 or 2^-48
    int left;       // words of the current batch not handed out yet
};

// seeds with 0 < s < 2^31 (s == 0 -> 1): pmesh/gsl/ranlxd.c:176-222
PMB_HD void wn_seed(WnRng &g, unsigned int s)
{
    if (s == 0) s = 1;
    unsigned int bits = s & 0x7fffffffu;    // 31-bit shift register, bit k = xbit[k]
    int a = 0, b = 18;
#pragma unroll
    for (int k = 0; k < 12; k++) {        // unrolled: x[] must keep compile-time indices (registers)
        unsigned long long w = 0;
#pragma unroll 1
        for (int l = 0; l < 48; l++) {
            const unsigned int ba = (bits >> a) & 1u, bb = (bits >> b) & 1u;
            w = (w << 1) | (ba ^ 1u);
            bits = (bits & ~(1u << a)) | ((ba ^ bb) << a);
            a = a == 30 ? 0 : a + 1;
            b = b == 30 ? 0 : b + 1;
        }
        g.x[k] = (double) w * WN_ONE_BIT;
    }
    g.c = 0;
    g.left = 0;
}

#define WN_STEP(i, j)                                  \
    do {                                               \
        double y_ = (g.x[j] - g.x[i]) - g.c;           \
        const bool neg_ = y_ < 0;                      \
        g.c = neg_ ? WN_ONE_BIT : 0.0;                 \
        g.x[i] = neg_ ? y_ + 1.0 : y_;                 \
    } while (0)

PMB_HD void wn_refill(WnRng &g)
{
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int blk = 0; blk < 16; blk++) {
        WN_STEP(0, 7); WN_STEP(1, 8); WN_STEP(2, 9); WN_STEP(3, 10); WN_STEP(4, 11); WN_STEP(5, 0);
        WN_STEP(6, 1); WN_STEP(7, 2); WN_STEP(8, 3); WN_STEP(9, 4); WN_STEP(10, 5); WN_STEP(11, 6);
    }
    WN_STEP(0, 7); WN_STEP(1, 8); WN_STEP(2, 9); WN_STEP(3, 10); WN_STEP(4, 11);
    WN_STEP(5, 0); WN_STEP(6, 1); WN_STEP(7, 2); WN_STEP(8, 3); WN_STEP(9, 4);
    // the oldest word is now register 10: rotate it to the front
    const double t0 = g.x[0], t1 = g.x[1], t2 = g.x[2], t3 = g.x[3], t4 = g.x[4], t5 = g.x[5], t6 = g.x[6], t7 = g.x[7],
                 t8 = g.x[8], t9 = g.x[9];
    g.x[0] = g.x[10]; g.x[1] = g.x[11];
    g.x[2] = t0; g.x[3] = t1; g.x[4] = t2; g.x[5] = t3; g.x[6] = t4; g.x[7] = t5; g.x[8] = t6; g.x[9] = t7;
    g.x[10] = t8; g.x[11] = t9;
    g.left = 12;
}

PMB_HD double wn_uniform(WnRng &g)
{
    if (g.left == 0) wn_refill(g);
    const double v = g.x[0];
    g.x[0] = g.x[1]; g.x[1] = g.x[2]; g.x[2] = g.x[3]; g.x[3] = g.x[4]; g.x[4] = g.x[5]; g.x[5] = g.x[6];
    g.x[6] = g.x[7]; g.x[7] = g.x[8]; g.x[8] = g.x[9]; g.x[9] = g.x[10]; g.x[10] = g.x[11]; g.x[11] = v;
    g.left--;
    return v;
}

// SAMPLE, pmesh/_whitenoise_imp.c:21-27
PMB_HD void wn_sample(WnRng &g, double &ampl, double &phase)
{
    phase = wn_uniform(g) * 2 * M_PI;
    do ampl = wn_uniform(g); while (ampl == 0);
}

// ---- host: the seed tables -------------------------------------------------------------------------
// The master stream takes the user's seed through `int` -> `unsigned long` (sign extension) and
// ranlxd's `int i = seed & 0xFFFFFFFF; xbit = i % 2; i /= 2` -- for seeds with bit 31 set the "bits"
// are 0 / -1.  Kept literally, on an int array (pmesh/_whitenoise_generics.h:75-76, ranlxd.c:187-199).
static void wn_seed_master(WnRng &g, unsigned int user_seed)
{
    int bit[31];
    long seed = (long) (unsigned long) (int) user_seed;
    if (seed == 0) seed = 1;
    int i = (int) (seed & 0xFFFFFFFFUL);
    for (int k = 0; k < 31; k++) { bit[k] = i % 2; i /= 2; }
    int a = 0, b = 18;
    for (int k = 0; k < 12; k++) {
        double x = 0;
        for (int l = 1; l <= 48; l++) {
            const double y = (double) ((bit[a] + 1) % 2);
            x += x + y;
            bit[a] = (bit[a] + bit[b]) % 2;
            a = (a + 1) % 31;
            b = (b + 1) % 31;
        }
        g.x[k] = WN_ONE_BIT * x;
    }
    g.c = 0;
    g.left = 0;
}

struct WnTables {
    int64_t N0, N1, s0, s1, m0, m1;   // mesh, local block
    unsigned int *t00, *t11;          // local block of seedtable[0][0] and [1][1], (m0, m1)
};

static inline void wn_setseed(WnTables &T, int64_t i, int64_t j, WnRng &master)
{
    // pmesh/_whitenoise_imp.c:29-52; only the tables the compressed fill reads are kept
    const unsigned int seed = 0x7fffffff * wn_uniform(master);
    const int64_t ci = (T.N0 - i) % T.N0, cj = (T.N1 - j) % T.N1;
    int64_t a = i - T.s0, b = j - T.s1;
    if (a >= 0 && a < T.m0 && b >= 0 && b < T.m1) T.t00[a * T.m1 + b] = seed;
    a = ci - T.s0; b = cj - T.s1;
    if (a >= 0 && a < T.m0 && b >= 0 && b < T.m1) T.t11[a * T.m1 + b] = seed;
}

static void wn_build_tables(WnTables &T, unsigned int user_seed)
{
    WnRng master;
    wn_seed_master(master, user_seed);
    const int64_t N0 = T.N0, N1 = T.N1;
    // the square spiral of pmesh/_whitenoise_generics.h:77-94, index expressions kept as written
    // (they mix Nmesh[0] and Nmesh[1]; the meshes of this engine are what they are given)
    for (int64_t i = 0; i < N0 / 2; i++) {
        for (int64_t j = 0; j < i; j++) wn_setseed(T, i, j, master);
        for (int64_t j = 0; j < i + 1; j++) wn_setseed(T, j, i, master);
        for (int64_t j = 0; j < i; j++) wn_setseed(T, N0 - 1 - i, j, master);
        for (int64_t j = 0; j < i + 1; j++) wn_setseed(T, N1 - 1 - j, i, master);
        for (int64_t j = 0; j < i; j++) wn_setseed(T, i, N1 - 1 - j, master);
        for (int64_t j = 0; j < i + 1; j++) wn_setseed(T, j, N0 - 1 - i, master);
        for (int64_t j = 0; j < i; j++) wn_setseed(T, N0 - 1 - i, N1 - 1 - j, master);
        for (int64_t j = 0; j < i + 1; j++) wn_setseed(T, N1 - 1 - j, N0 - 1 - i, master);
    }
}

// ---- one Fourier column (i, j, 0..N/2) -------------------------------------------------------------
struct WnArgs {
    int64_t N[3], start[3], size[3], strides[3];   // strides in bytes
    const unsigned int *t00, *t11;                 // local (size[0], size[1]) blocks of the seed tables
    int unitary;
    int fast_axis;       // 0: consecutive threads walk i, 1: they walk j (the column axis of smaller stride)
};

struct WnColumn {
    WnRng mine, lower;
    int64_t i, j;
    bool d;              // the k = 0 / Nyquist modes come from the Hermitian partner's stream, conjugated
    bool selfconj_ij;
};

PMB_HD void wn_column_init(WnColumn &c, const WnArgs &a, int64_t li, int64_t lj)
{
    c.i = li + a.start[0];
    c.j = lj + a.start[1];
    int64_t ci = a.N[0] - c.i, cj = a.N[1] - c.j;
    if (ci >= a.N[0]) ci -= a.N[0];
    if (cj >= a.N[1]) cj -= a.N[1];
    // pmesh/_whitenoise_generics.h:125-131
    c.d = (ci == c.i && cj < c.j) || (ci < c.i && cj != c.j) || (ci < c.i && cj == c.j);
    c.selfconj_ij = ci == c.i && cj == c.j;
    wn_seed(c.mine, a.t00[li * a.size[1] + lj]);
    if (c.d) wn_seed(c.lower, a.t11[li * a.size[1] + lj]);
}

// draws the samples of mode k (the streams advance whether or not the mode is stored) and returns
// true with (re, im), already rounded to the canvas precision T, when k lies inside the block
template <typename T>
PMB_HD bool wn_column_mode(WnColumn &c, const WnArgs &a, int64_t k, double &re, double &im)
{
    const int64_t kmax = a.N[2] / 2;
    double ampl, phase;
    // the two streams are independent: which one the reference samples first does not change what
    // either returns (pmesh/_whitenoise_generics.h:150-158)
    wn_sample(c.mine, ampl, phase);
    const bool use_conj = c.d && (k == 0 || k == kmax);
    if (c.d) {
        double al, pl;
        wn_sample(c.lower, al, pl);
        if (use_conj) { ampl = al; phase = pl; }
    }
    if (k < a.start[2] || k >= a.start[2] + a.size[2]) return false;
    ampl = a.unitary ? 1.0 : sqrt(-log(ampl));
    re = (double) (T) (ampl * cos(phase));
    im = (double) (T) (ampl * sin(phase));
    if (use_conj) im = -im;
    if (c.selfconj_ij && (a.N[2] - k) % a.N[2] == k) {
        // self-conjugate mode: purely real
        im = 0;
        if (a.unitary) re = 1;
    }
    if (c.i == 0 && c.j == 0 && k == 0) re = im = 0;   // the mean
    return true;
}
