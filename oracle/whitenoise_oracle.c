/* whitenoise_oracle.c -- CPU ORACLE of the reference's white-noise generator.  TEST INFRASTRUCTURE:
 * only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it; the product never does.
 *
 * Restates, in plain C and in INTEGER arithmetic, what the reference computes with doubles:
 *   - the RANLUX double-precision generator at luxury level 1 ("ranlxd1"): pmesh/gsl/ranlxd.c:36-245
 *     (seeding :176-222, the subtract-with-borrow recurrence :68-159, output order :167-174);
 *   - the per-(i, j) seed table drawn from one master stream:      pmesh/_whitenoise_generics.h:73-94,
 *                                                                  pmesh/_whitenoise_imp.c:29-52
 *   - the column fill (N-GenIC / Gadget scheme):                   pmesh/_whitenoise_generics.h:106-232,
 *     SAMPLE                                                       pmesh/_whitenoise_imp.c:21-27
 *
 * Parity status: PINNED -- tests/test_oracle.py compares it bit for bit with the compiled reference
 * (oracle/_ref/pmesh_ref/_whitenoise) and with tests/golden/whitenoise_golden.npz.
 *
 * Formulation.  Every state word of ranlxd is k * 2^-48 with an integer 0 <= k < 2^48, and the
 * reference's double arithmetic on them is exact; so the state is kept as uint64 integers here:
 *     z[n] = z[n-5] - z[n-12] - borrow  (mod 2^48),
 * a ring of the last 12 words.  One "refill" advances the ring by 202 steps; the 12 words then in
 * the ring are handed out oldest first (that is what ranlxd_get_double's ir / ir_old bookkeeping
 * amounts to), then the ring is advanced again.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define WN_MOD ((int64_t) 1 << 48)

typedef struct {
    int64_t z[12];   /* ring; z[head] is the oldest word */
    int head;
    int64_t borrow;
    int left;        /* words of the current batch not handed out yet */
    int cursor;
} wn_rng;

static void wn_seed(wn_rng *g, unsigned long s)
{
    /* ranlxd.c:176-222.  The 31 seed bits are taken from an `int` with C's truncating % and /, so a
     * seed with bit 31 set yields "bits" in {0, -1}; kept literally. */
    int bit[31];
    long seed = s == 0 ? 1 : (long) s;
    int i = (int) (seed & 0xFFFFFFFFUL);
    for (int k = 0; k < 31; k++) { bit[k] = i % 2; i /= 2; }
    int a = 0, b = 18;
    for (int k = 0; k < 12; k++) {
        /* x += x + y in doubles, 48 times: exact while x < 2^53.  y can only be 0 or 1 (or, for the
         * negative "bits" of the quirk above, 0 and -1 -> (bit + 1) % 2 in {0, 1, 0}) */
        double x = 0;
        for (int l = 1; l <= 48; l++) {
            double y = (double) ((bit[a] + 1) % 2);
            x += x + y;
            bit[a] = (bit[a] + bit[b]) % 2;
            a = (a + 1) % 31;
            b = (b + 1) % 31;
        }
        g->z[k] = (int64_t) x;
    }
    g->head = 0;
    g->borrow = 0;
    g->left = 0;
    g->cursor = 0;
}

static void wn_step(wn_rng *g)
{
    int64_t v = g->z[(g->head + 7) % 12] - g->z[g->head] - g->borrow;
    if (v < 0) { v += WN_MOD; g->borrow = 1; } else g->borrow = 0;
    g->z[g->head] = v;
    g->head = (g->head + 1) % 12;
}

static double wn_uniform(wn_rng *g)
{
    if (g->left == 0) {
        for (int k = 0; k < 202; k++) wn_step(g);
        g->left = 12;
        g->cursor = g->head;
    }
    const int64_t v = g->z[g->cursor];
    g->cursor = (g->cursor + 1) % 12;
    g->left--;
    return (double) v * (1.0 / 281474976710656.0);
}

/* exported for the tests: the first n uniforms of a stream */
void wn_oracle_stream(unsigned long seed, double *out, int n)
{
    wn_rng g;
    wn_seed(&g, seed);
    for (int i = 0; i < n; i++) out[i] = wn_uniform(&g);
}

static void wn_sample(wn_rng *g, double *ampl, double *phase)
{
    /* _whitenoise_imp.c:21-27 */
    *phase = wn_uniform(g) * 2 * M_PI;
    do *ampl = wn_uniform(g); while (*ampl == 0);
}

typedef struct {
    ptrdiff_t Nmesh[3], start[3], size[3], strides[3];   /* strides in BYTES */
    unsigned int *table[2][2];
} wn_geom;

static void wn_setseed(wn_geom *s, int i, int j, wn_rng *g)
{
    /* _whitenoise_imp.c:29-52 */
    unsigned int seed = 0x7fffffff * wn_uniform(g);
    int ii[2] = {i, (int) ((s->Nmesh[0] - i) % s->Nmesh[0])};
    int jj[2] = {j, (int) ((s->Nmesh[1] - j) % s->Nmesh[1])};
    for (int d = 0; d < 2; d++) { ii[d] -= s->start[0]; jj[d] -= s->start[1]; }
    for (int d1 = 0; d1 < 2; d1++)
        for (int d2 = 0; d2 < 2; d2++)
            if (ii[d1] >= 0 && ii[d1] < s->size[0] && jj[d2] >= 0 && jj[d2] < s->size[1])
                s->table[d1][d2][ii[d1] * s->size[1] + jj[d2]] = seed;
}

/* canvas: complex64 (elsize 8) or complex128 (elsize 16), local block [start, start + size) of the
 * Nmesh^3 Fourier mesh with byte strides.  Returns 0, or -1 for bad arguments. */
int wn_oracle_fill(void *canvas, int elsize, const ptrdiff_t *Nmesh, const ptrdiff_t *start, const ptrdiff_t *size,
                   const ptrdiff_t *strides, unsigned int seed, int unitary)
{
    if (elsize != 8 && elsize != 16) return -1;
    wn_geom s;
    for (int d = 0; d < 3; d++) { s.Nmesh[d] = Nmesh[d]; s.start[d] = start[d]; s.size[d] = size[d]; s.strides[d] = strides[d]; }
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) s.table[a][b] = calloc((size_t) (size[0] * size[1] > 0 ? size[0] * size[1] : 1), sizeof(unsigned int));

    /* which signs of k_z are asked for (_whitenoise_generics.h:42-71) */
    int signs[3] = {1, 0, 0};
    for (ptrdiff_t k = Nmesh[2] / 2 + 1; k < Nmesh[2]; k++)
        if (k - start[2] >= 0 && k - start[2] < size[2]) { signs[0] = -1; signs[1] = 1; break; }

    /* the seed table: one master stream walked in a square spiral (_whitenoise_generics.h:73-94) */
    wn_rng master;
    wn_seed(&master, (unsigned long) (int) seed);
    const int N0 = (int) Nmesh[0], N1 = (int) Nmesh[1];
    for (int i = 0; i < N0 / 2; i++) {
        for (int j = 0; j < i; j++) wn_setseed(&s, i, j, &master);
        for (int j = 0; j < i + 1; j++) wn_setseed(&s, j, i, &master);
        for (int j = 0; j < i; j++) wn_setseed(&s, N0 - 1 - i, j, &master);
        for (int j = 0; j < i + 1; j++) wn_setseed(&s, N1 - 1 - j, i, &master);
        for (int j = 0; j < i; j++) wn_setseed(&s, i, N1 - 1 - j, &master);
        for (int j = 0; j < i + 1; j++) wn_setseed(&s, j, N0 - 1 - i, &master);
        for (int j = 0; j < i; j++) wn_setseed(&s, N0 - 1 - i, N1 - 1 - j, &master);
        for (int j = 0; j < i + 1; j++) wn_setseed(&s, N1 - 1 - j, N0 - 1 - i, &master);
    }

    /* columns (_whitenoise_generics.h:106-232) */
    for (ptrdiff_t i = start[0]; i < start[0] + size[0]; i++) {
        ptrdiff_t ci = Nmesh[0] - i;
        if (ci >= Nmesh[0]) ci -= Nmesh[0];
        for (ptrdiff_t j = start[1]; j < start[1] + size[1]; j++) {
            ptrdiff_t cj = Nmesh[1] - j;
            if (cj >= Nmesh[1]) cj -= Nmesh[1];
            /* modes of the k_z = 0 and Nyquist planes whose conjugate partner lies in the "lower"
             * half are generated from the partner's stream and conjugated */
            int d = 0;
            if ((ci == i && cj < j) || (ci < i && cj != j) || (ci < i && cj == j)) d = 1;
            const ptrdiff_t li = i - start[0], lj = j - start[1];
            for (int is = 0; signs[is] != 0; is++) {
                const int sign = signs[is];
                wn_rng lower, mine;
                wn_seed(&lower, s.table[d][d][li * size[1] + lj]);
                wn_seed(&mine, sign == 1 ? s.table[0][0][li * size[1] + lj] : s.table[1][1][li * size[1] + lj]);
                for (ptrdiff_t k = 0; k <= Nmesh[2] / 2; k++) {
                    const int use_conj = d && (k == 0 || k == Nmesh[2] / 2);
                    double ampl, phase;
                    if (use_conj) { wn_sample(&mine, &ampl, &phase); wn_sample(&lower, &ampl, &phase); }
                    else { wn_sample(&lower, &ampl, &phase); wn_sample(&mine, &ampl, &phase); }
                    ptrdiff_t kk = k;
                    /* the reference tests the presence of the mode BEFORE mirroring k for the negative
                     * sign (_whitenoise_generics.h:160-167) and again, mirrored, when storing (:14-27) */
                    if (!(k - start[2] >= 0 && k - start[2] < size[2])) continue;
                    ampl = unitary ? 1.0 : sqrt(-log(ampl));
                    double re = ampl * cos(phase), im = ampl * sin(phase);
                    if (elsize == 8) { re = (float) re; im = (float) im; }
                    if (sign == -1) { kk = Nmesh[2] - k; im = -im; }
                    if (use_conj) im *= -1;
                    if ((Nmesh[0] - i) % Nmesh[0] == i && (Nmesh[1] - j) % Nmesh[1] == j && (Nmesh[2] - kk) % Nmesh[2] == kk) {
                        im = 0;
                        if (unitary) re = 1;
                    }
                    if (i == 0 && j == 0 && kk == 0) re = im = 0;
                    const ptrdiff_t lk = kk - start[2];
                    if (lk < 0 || lk >= size[2]) continue;
                    char *p = (char *) canvas + li * strides[0] + lj * strides[1] + lk * strides[2];
                    if (elsize == 16) { ((double *) p)[0] = re; ((double *) p)[1] = im; }
                    else { ((float *) p)[0] = (float) re; ((float *) p)[1] = (float) im; }
                }
            }
        }
    }
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) free(s.table[a][b]);
    return 0;
}
