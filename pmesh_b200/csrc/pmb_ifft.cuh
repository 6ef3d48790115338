// pmb_ifft.cuh -- the axis-0 pass of the three backward transforms of a force evaluation, FUSED with the gradient
// transfers (and, on one rank, with the last pass of r2c):   out_d = IDFT_axis0( i m_d(k_d) / k^2 * rho_k ),  d = 0, 1, 2
//
// Replaces, for `[rhok.apply(T_d).c2r() for d in 0..2]` (pmesh/pm.py:617-648 Field.apply, 987-1019 c2r;
// examples/nbody.py:162-170, 211-216), the transfer pass over the modes (pmb_k_transfer_grad3: 16 B read +
// 48 B written per complex cell) AND the axis-0 pass of cuFFT in each of the three transforms (3 x 32 B per
// cell): the modes are read once, multiplied on load, transformed along axis 0 in registers + shared memory and
// written once per direction: 16 + 48 bytes per complex cell instead of 64 + 96.  The remaining two axes of every
// transform stay with cuFFT (batched 2-D c2r over the planes, library time).
//
// Fewer transforms than directions.  The multipliers of directions 1 and 2 are constants of an axis-0 line, so both are
// ONE transform phi = IDFT(pre / k^2 * rho_k) scaled at the store (i m_1 phi, i m_2 phi).  Direction 0 needs m_0(k_0)
// inside the transform -- a second transform -- unless the gradient is the finite-difference one (PMB_TF_GRAVITY_FD4):
// i kfinite(k_0) is the symbol of the 4th-order central difference, i.e. a 5-point stencil along the line of the same
// phi, applied out of the exchange buffer.  With `forward` the line is first transformed forward (conj . IDFT . conj):
// the kernel then starts from the planes' 2-D r2c and the density modes are never stored.
//
// One line of N = 16 * R2 * R3 points (R3 = 1: two passes) is transformed by N / 16 threads, each holding 16
// points in registers: a Stockham autosort transform with radices (16, R2, R3) -- pass 1 reads global memory
// (stride N / 16: rows of the bundle coalesce), passes exchange through shared memory (index padded by one
// element per 16: conflict-free for the stride-16 stores of pass 1), the last pass stores to global memory in
// natural order.  A CTA owns a bundle of B lines:
//   STRIDED (one rank, complex layout (n0, n1, nc): the elements of a line are n1*nc apart): B ADJACENT lines,
//           thread = (point, line) with the line fastest, so every access of a row is B * 16 contiguous bytes
//           (B = 8 for 1024 double-precision points: whole 128-byte lines; 64-byte rows measured 2 x slower);
//   CONTIG  (P ranks, "transposed" layout (m1, mc, n0): lines contiguous): B consecutive lines.
// Twiddles W_N^k = exp(+2 pi i k / N) come from a table built on the host in double precision.
//
// The per-thread phases are PMB_HD so that tests/harness/host_harness.cpp can run the exact index arithmetic
// and butterflies on the CPU (one "thread" after the other, phase by phase) against numpy.fft.
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector_types.h>

#include "pmb_window.h"   // PMB_HD

template <typename C> struct pmb_cx_real;
template <> struct pmb_cx_real<double2> { typedef double type; };
template <> struct pmb_cx_real<float2> { typedef float type; };

template <typename C> PMB_HD C pmb_cx(typename pmb_cx_real<C>::type x, typename pmb_cx_real<C>::type y) { C r; r.x = x; r.y = y; return r; }
template <typename C> PMB_HD C pmb_cadd(C a, C b) { return pmb_cx<C>(a.x + b.x, a.y + b.y); }
template <typename C> PMB_HD C pmb_csub(C a, C b) { return pmb_cx<C>(a.x - b.x, a.y - b.y); }
// explicit fused multiply-adds: the library is compiled -fmad=false (paint / readout parity), this file wants them
template <typename C> PMB_HD C pmb_cmul(C a, C b)
{
    return pmb_cx<C>(fma(-a.y, b.y, a.x * b.x), fma(a.x, b.y, a.y * b.x));
}

// v * exp(+2 pi i K / R), K and R compile-time
template <int R, int K, typename C> PMB_HD C pmb_ctw(C v)
{
    typedef typename pmb_cx_real<C>::type T;
    if constexpr (K == 0) return v;
    else if constexpr (4 * K == R) return pmb_cx<C>(-v.y, v.x);                                   // * i
    else if constexpr (8 * K == R) { const T s = (T) 0.70710678118654752440; return pmb_cx<C>((v.x - v.y) * s, (v.x + v.y) * s); }
    else if constexpr (8 * K == 3 * R) { const T s = (T) 0.70710678118654752440; return pmb_cx<C>((-v.x - v.y) * s, (v.x - v.y) * s); }
    else {
        // R = 16, K in {1, 3, 5, 7}
        const T c1 = (T) 0.92387953251128675613, s1 = (T) 0.38268343236508977173;   // cos, sin (pi / 8)
        const T wr = K == 1 ? c1 : (K == 3 ? s1 : (K == 5 ? -s1 : -c1));
        const T wi = (K == 1 || K == 7) ? s1 : c1;
        return pmb_cmul<C>(v, pmb_cx<C>(wr, wi));
    }
}

template <int R, int K, typename C> struct pmb_idft_comb {
    static PMB_HD void run(C *v, const C *e, const C *o)
    {
        const C t = pmb_ctw<R, K, C>(o[K]);
        v[K] = pmb_cadd<C>(e[K], t);
        v[K + R / 2] = pmb_csub<C>(e[K], t);
        if constexpr (K + 1 < R / 2) pmb_idft_comb<R, K + 1, C>::run(v, e, o);
    }
};

// unnormalised inverse DFT of R points in place, natural order in and out (decimation in time, radix 2)
template <int R, typename C> PMB_HD void pmb_idft(C *v)
{
    if constexpr (R == 1) {
    } else if constexpr (R == 2) {
        const C a = v[0], b = v[1];
        v[0] = pmb_cadd<C>(a, b);
        v[1] = pmb_csub<C>(a, b);
    } else if constexpr (R == 4) {
        const C t0 = pmb_cadd<C>(v[0], v[2]), t1 = pmb_csub<C>(v[0], v[2]);
        const C t2 = pmb_cadd<C>(v[1], v[3]), d = pmb_csub<C>(v[1], v[3]);
        const C t3 = pmb_cx<C>(-d.y, d.x);
        v[0] = pmb_cadd<C>(t0, t2); v[1] = pmb_cadd<C>(t1, t3);
        v[2] = pmb_csub<C>(t0, t2); v[3] = pmb_csub<C>(t1, t3);
    } else {
        C e[R / 2], o[R / 2];
#pragma unroll
        for (int k = 0; k < R / 2; k++) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
        pmb_idft<R / 2, C>(e);
        pmb_idft<R / 2, C>(o);
        pmb_idft_comb<R, 0, C>::run(v, e, o);
    }
}

// radices of the passes after the first (which is always 16): N = 16 * R2 * R3
template <int N> struct pmb_ifft_radices {
    static constexpr int R2 = N == 64 ? 4 : (N == 128 ? 8 : (N == 512 ? 8 : 16));
    static constexpr int R3 = N / 16 / R2;
};

template <typename C, int N, bool CONTIG, int B> struct pmb_ifft_line {
    static constexpr int R2 = pmb_ifft_radices<N>::R2, R3 = pmb_ifft_radices<N>::R3;
    static constexpr int TPL = N / 16;                      // threads per line
    static constexpr int NP = N + N / 16;                   // padded points of a line
    static constexpr int LS = NP + ((10 - NP % 8) % 8);     // CONTIG: line stride = 2 mod 8 elements
    static constexpr int SMEM_ELEMS = CONTIG ? B * LS : NP * B;
    static_assert(16 * R2 * R3 == N && (R3 == 1 || R3 == 2 || R3 == 4 || R3 == 8 || R3 == 16), "unsupported line length");

    static PMB_HD int sidx(int b, int idx)
    {
        const int p = idx + (idx >> 4);
        return CONTIG ? b * LS + p : p * B + b;
    }
    // pass 1, radix 16, Ns = 1: thread t holds the points t + r * TPL (already multiplied by the transfer)
    static PMB_HD void p1(C *sm, int b, int t, C *v)
    {
        pmb_idft<16, C>(v);
#pragma unroll
        for (int q = 0; q < 16; q++) sm[sidx(b, t * 16 + q)] = v[q];
    }
    // pass 2, radix R2, Ns = 16: 16 / R2 butterflies j = t + m * TPL per thread
    static PMB_HD void p2_load(const C *sm, int b, int t, C *v)
    {
#pragma unroll
        for (int m = 0; m < 16 / R2; m++)
#pragma unroll
            for (int r = 0; r < R2; r++) v[m * R2 + r] = sm[sidx(b, t + m * TPL + r * (N / R2))];
    }
    // twiddles: tw2(r, k) = W_N^(k r R3) = W_(16 R2)^(k r), k < 16, r < R2;  tw3(r, j) = W_N^(j r), j < N / R3, r < R3
    template <typename TW> static PMB_HD void p2_compute(int t, C *v, TW tw2)
    {
#pragma unroll
        for (int m = 0; m < 16 / R2; m++) {
            const int k = (t + m * TPL) & 15;
#pragma unroll
            for (int r = 1; r < R2; r++) v[m * R2 + r] = pmb_cmul<C>(v[m * R2 + r], tw2(r, k));
            pmb_idft<R2, C>(v + m * R2);
        }
    }
    // where output q of butterfly m of thread t goes after pass 2 (in shared memory, or in the line when R3 == 1)
    static PMB_HD int p2_out(int t, int m, int q)
    {
        const int j = t + m * TPL;
        return (j >> 4) * (16 * R2) + (j & 15) + q * 16;
    }
    static PMB_HD void p2_store(C *sm, int b, int t, const C *v)
    {
#pragma unroll
        for (int m = 0; m < 16 / R2; m++)
#pragma unroll
            for (int q = 0; q < R2; q++) sm[sidx(b, p2_out(t, m, q))] = v[m * R2 + q];
    }
    // pass 3, radix R3, Ns = 16 * R2 = N / R3
    static PMB_HD void p3_load(const C *sm, int b, int t, C *v)
    {
#pragma unroll
        for (int m = 0; m < 16 / R3; m++)
#pragma unroll
            for (int r = 0; r < R3; r++) v[m * R3 + r] = sm[sidx(b, t + m * TPL + r * (N / R3))];
    }
    template <typename TW> static PMB_HD void p3_compute(int t, C *v, TW tw3)
    {
#pragma unroll
        for (int m = 0; m < 16 / R3; m++) {
            const int j = t + m * TPL;
#pragma unroll
            for (int r = 1; r < R3; r++) v[m * R3 + r] = pmb_cmul<C>(v[m * R3 + r], tw3(r, j));
            pmb_idft<R3, C>(v + m * R3);
        }
    }
    static PMB_HD int p3_out(int t, int m, int q) { return t + m * TPL + q * (N / R3); }
};

// lines per CTA.  CONTIG: 256 threads where the line is short, at least 4 lines (2 for 4096 double-precision points:
// shared memory).  STRIDED: the B adjacent lines of a bundle make rows of B * sizeof(C) contiguous bytes -- whole
// 128-byte lines for double precision up to 1024 points (B = 8; half lines measured 2 x slower), what shared memory
// allows beyond that.  WIDE = false selects the narrower bundle (A/B measurements).
template <typename C, int N, bool CONTIG = true, bool WIDE = true> struct pmb_ifft_bundle {
    static constexpr int BASE = (N == 4096 && sizeof(C) == 16) ? 2 : (4096 / N > 4 ? 4096 / N : 4);
    static constexpr int B = (!CONTIG && WIDE && N == 1024 && BASE < 8) ? 8 : BASE;
};

#if defined(__CUDACC__)

// One TRANSFORM of a launch: phi = IDFT_axis0( pre / k^2 * [m_0(i0)] * in ); output o = i * c_o * phi with c_o a constant
// of the LINE (1, m_1(i1) or m_2(i2)).  The multipliers of directions 1 and 2 do not depend on the
// position along the line, so ONE transform serves both: a force evaluation costs two transforms per line, not three.
struct PmbIfftTransform {
    int axis0mul;            // 1: the input is multiplied by mtab[0][i0] (direction 0)
    int nout;                // 0, 1 or 2
    void *out[2];
    int linemul[2];          // 0: none, 1: mtab[1][i1], 2: mtab[2][i2]
    // direction 0 of the FINITE-DIFFERENCE gradient (PMB_TF_GRAVITY_FD4): i kfinite(k_0) = i (8 sin w - sin 2w) / (6 C)
    // is the symbol of the 4th-order central difference, so direction 0 is that stencil applied ALONG THE LINE to the
    // transform phi of pre / k^2 * in that directions 1 and 2 store anyway:
    //   sten_out(x) = sten_c * (8 (phi(x+1) - phi(x-1)) - (phi(x+2) - phi(x-2))),  sten_c = 1 / (12 C), periodic in x.
    // One transform of the line then serves all three directions.
    void *sten_out;          // NULL: no stencil output
    double sten_c;
};

struct PmbIfftArgs {
    const void *in;
    int ntr;                 // transforms of this launch (1 or 2), all of the same input
    PmbIfftTransform tr[2];
    int P;                   // 1: (n0, n1, nc) strided lines; > 1: (m1, mc, n0) contiguous lines
    int64_t nlines;          // n1 * nc   |  m1 * mc
    int64_t nc, s1, mc, s2;
    const double *ktab[3];   // wavenumbers per global index of each axis
    const double *mtab[3];   // gradient multiplier per global index of each axis
    const void *tw;          // W_N^k, k = 0 .. N-1, in the field's precision
    double pre;
    // 1: `in` holds the field transformed along axes 1 and 2 only; the line is first transformed FORWARD along axis 0
    // (in registers + shared memory, conj . IDFT . conj), then multiplied and transformed back: the last pass of r2c,
    // the transfers and the first pass of the three c2r in one kernel -- the density modes never exist in HBM
    int forward;
};

// 1 / x for a normal, positive double: hardware seed (MUFU.RCP64H) + three Newton steps, no special cases -- the host
// takes this kernel only when every k^2 of the mesh lies in [1e-280, 1e280]
__device__ __forceinline__ double pmb_fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#pragma unroll
    for (int it = 0; it < 3; it++) r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// Tables of a CTA in shared memory (TSM; lines up to 2048 points): the wavenumbers and multipliers of axis 0 and the
// twiddles, laid out so that the lanes of a warp read consecutive entries.  Streaming 64 KB bundles through the SM
// leaves nothing of these tables in L1: read through the cache they cost an L2 round trip per entry.
template <typename C, int N> struct pmb_ifft_tables {
    static constexpr int R2 = pmb_ifft_radices<N>::R2, R3 = pmb_ifft_radices<N>::R3;
    static constexpr bool TSM = N <= 2048;
    static constexpr int NTW2 = 16 * R2, NTW3 = R3 > 1 ? N : 0;      // tw2[r * 16 + k], tw3[r * (N / R3) + j]
    static constexpr size_t BYTES = TSM ? (size_t) (NTW2 + NTW3) * sizeof(C) + 2 * (size_t) N * sizeof(double) : 0;
};

template <typename C, int N, bool CONTIG, int B, int MINB>
__global__ void __launch_bounds__(B * (N / 16), MINB)
pmb_k_ifft_grad(PmbIfftArgs a)
{
    typedef pmb_ifft_line<C, N, CONTIG, B> F;
    typedef pmb_ifft_tables<C, N> TB;
    typedef typename pmb_cx_real<C>::type T;
    constexpr int TPL = F::TPL;
    constexpr int NT = B * TPL;
    extern __shared__ __align__(16) unsigned char pmb_ifft_smem[];
    C *sm = (C *) pmb_ifft_smem;
    C *s_tw2 = sm + F::SMEM_ELEMS;
    C *s_tw3 = s_tw2 + TB::NTW2;
    double *s_k0 = (double *) (s_tw3 + TB::NTW3);
    double *s_m0 = s_k0 + N;
    const int tid = threadIdx.x;
    const int b = CONTIG ? tid / TPL : tid % B;
    const int t = CONTIG ? tid % TPL : tid / B;
    const C *__restrict__ in = (const C *) a.in;
    const C *__restrict__ twp = (const C *) a.tw;
    const double *__restrict__ k0tab = a.ktab[0];
    const double *__restrict__ m0tab = a.mtab[0];
    if constexpr (TB::TSM) {
        for (int i = tid; i < TB::NTW2; i += NT) s_tw2[i] = twp[(i & 15) * (i >> 4) * F::R3];
        for (int i = tid; i < TB::NTW3; i += NT) s_tw3[i] = twp[(i % (N / F::R3)) * (i / (N / F::R3))];
        for (int i = tid; i < N; i += NT) { s_k0[i] = k0tab[i]; s_m0[i] = m0tab[i]; }
        __syncthreads();
    }
    auto tw2 = [&](int r, int k) -> C { if constexpr (TB::TSM) return s_tw2[r * 16 + k]; else return __ldg(twp + k * r * F::R3); };
    auto tw3 = [&](int r, int j) -> C { if constexpr (TB::TSM) return s_tw3[r * (N / F::R3) + j]; else return __ldg(twp + j * r); };
    const int64_t nbundles = (a.nlines + B - 1) / B;
    for (int64_t bundle = blockIdx.x; bundle < nbundles; bundle += gridDim.x) {
        int64_t L = bundle * B + b;
        const bool live = L < a.nlines;
        if (!live) L = a.nlines - 1;
        int64_t i1, i2;
        if (CONTIG) { i2 = a.s2 + L % a.mc; i1 = a.s1 + L / a.mc; }
        else { i1 = L / a.nc; i2 = L % a.nc; }
        const double k1 = __ldg(a.ktab[1] + i1), k2v = __ldg(a.ktab[2] + i2);
        const double c12 = k1 * k1 + k2v * k2v;
        const T mline[3] = {(T) 1, (T) __ldg(a.mtab[1] + i1), (T) __ldg(a.mtab[2] + i2)};
        // element i0 of line L
        const int64_t base = CONTIG ? L * N : L;
        const int64_t estr = CONTIG ? 1 : a.nlines;
        for (int tr = 0; tr < a.ntr; tr++) {
            const bool ax0 = a.tr[tr].axis0mul != 0;
            C v[16];
            {
                const C *__restrict__ p = in + base + (int64_t) t * estr;
                const int64_t step = (int64_t) TPL * estr;
#pragma unroll
                for (int r = 0; r < 16; r++) { v[r] = __ldcg(p); p += step; }      // L2 only: the second transform re-reads it there
            }
            if (a.forward) {
                constexpr int RLf = F::R3 > 1 ? F::R3 : F::R2;
#pragma unroll
                for (int r = 0; r < 16; r++) v[r].y = -v[r].y;
                F::p1(sm, b, t, v);
                __syncthreads();
                F::p2_load(sm, b, t, v);
                __syncthreads();
                F::p2_compute(t, v, tw2);
                if constexpr (F::R3 > 1) {
                    F::p2_store(sm, b, t, v);
                    __syncthreads();
                    F::p3_load(sm, b, t, v);
                    __syncthreads();
                    F::p3_compute(t, v, tw3);
                }
                // output q of butterfly m is point t + TPL * (m + (16 / RL) q) of the line: exactly the points this thread
                // feeds to pass 1 of the next transform -- the spectrum changes registers, not threads
                C u[16];
#pragma unroll
                for (int m = 0; m < 16 / RLf; m++)
#pragma unroll
                    for (int q = 0; q < RLf; q++) u[m + (16 / RLf) * q] = pmb_cx<C>(v[m * RLf + q].x, -v[m * RLf + q].y);
#pragma unroll
                for (int r = 0; r < 16; r++) v[r] = u[r];
            }
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const int i0 = t + r * TPL;
                double k0, m0;
                if constexpr (TB::TSM) { k0 = s_k0[i0]; m0 = s_m0[i0]; }
                else { k0 = __ldg(k0tab + i0); m0 = __ldg(m0tab + i0); }
                double kk = fma(k0, k0, c12);
                kk = kk == 0 ? 1.0 : kk;
                double g = pmb_fast_rcp(kk) * a.pre;
                g = ax0 ? g * m0 : g;
                const T gr = (T) g;
                v[r] = pmb_cx<C>(v[r].x * gr, v[r].y * gr);
            }
            F::p1(sm, b, t, v);
            __syncthreads();
            F::p2_load(sm, b, t, v);
            __syncthreads();
            F::p2_compute(t, v, tw2);
            if constexpr (F::R3 > 1) {
                F::p2_store(sm, b, t, v);
                __syncthreads();
                F::p3_load(sm, b, t, v);
                __syncthreads();
                F::p3_compute(t, v, tw3);
            }
            constexpr int RL = F::R3 > 1 ? F::R3 : F::R2;
            // output q of butterfly m sits at idx0(m) + q * QS points of the line
            constexpr int QS = F::R3 > 1 ? N / F::R3 : 16;
            C *__restrict__ sten = (C *) a.tr[tr].sten_out;
            if (sten) {
                // park phi in the exchange buffer (every thread has read its last pass out of it: the barrier above)
#pragma unroll
                for (int m = 0; m < 16 / RL; m++) {
                    const int idx0 = F::R3 > 1 ? F::p3_out(t, m, 0) : F::p2_out(t, m, 0);
#pragma unroll
                    for (int q = 0; q < RL; q++) sm[F::sidx(b, idx0 + q * QS)] = v[m * RL + q];
                }
            }
            if (live) {
                const int nout = a.tr[tr].nout;
                const int64_t qstep = (int64_t) QS * estr;
#pragma unroll 1
                for (int o = 0; o < nout; o++) {
                    C *__restrict__ out = (C *) a.tr[tr].out[o] + base;
                    const int lm = a.tr[tr].linemul[o];
                    const T ml = lm == 0 ? mline[0] : (lm == 1 ? mline[1] : mline[2]);
#pragma unroll
                    for (int m = 0; m < 16 / RL; m++) {
                        const int idx0 = F::R3 > 1 ? F::p3_out(t, m, 0) : F::p2_out(t, m, 0);
                        C *__restrict__ p = out + (int64_t) idx0 * estr;
#pragma unroll
                        for (int q = 0; q < RL; q++) {
                            const C w = v[m * RL + q];
                            *p = pmb_cx<C>(-(w.y * ml), w.x * ml);       // i * ml * phi
                            p += qstep;
                        }
                    }
                }
            }
            if (sten) {
                __syncthreads();
                // thread t of the line now owns the 16 CONSECUTIVE points 16 t .. 16 t + 15
                const T c = (T) a.tr[tr].sten_c;
                const int x0 = 16 * t;
                auto at = [&](int x) -> C { return sm[F::sidx(b, (x + N) & (N - 1))]; };
                C pm2 = at(x0 - 2), pm1 = at(x0 - 1), p0 = at(x0), pp1 = at(x0 + 1);
                C *__restrict__ p = sten + base + (int64_t) x0 * estr;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const C pp2 = at(x0 + j + 2);
                    const T dx = (T) 8 * (pp1.x - pm1.x) - (pp2.x - pm2.x);
                    const T dy = (T) 8 * (pp1.y - pm1.y) - (pp2.y - pm2.y);
                    if (live) *p = pmb_cx<C>(dx * c, dy * c);
                    p += estr;
                    pm2 = pm1; pm1 = p0; p0 = pp1; pp1 = pp2;
                }
                __syncthreads();       // before the next transform's pass 1 overwrites the buffer
            }
        }
    }
}

#endif   // __CUDACC__
