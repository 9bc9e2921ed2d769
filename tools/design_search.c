// design_search.c - offline search for seed covering designs (tools/, not part of the product path).
//
// A seed design for (L, k) is a family of position masks M_c (|M_c| = s nt, at most R contiguous
// runs each) such that every set T of k mismatch positions is avoided by at least one mask
// (M_c & T == 0).  "b blocks, choose b-k" (the pigeonhole scheme) is one such family; smaller
// families with longer keys exist (covering designs).  Simulated annealing over C masks, cost =
// number of uncovered k-subsets.  Output: one line per design found: L k s C masks...
//
//   gcc -O2 -o /tmp/design_search tools/design_search.c && /tmp/design_search L k s C maxruns seed iters
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

static uint64_t rng_s = 88172645463325252ull;
static inline uint64_t rnd(void) { rng_s ^= rng_s << 13; rng_s ^= rng_s >> 7; rng_s ^= rng_s << 17; return rng_s; }

static int L, K, S, C, MAXRUNS;
static uint32_t* subsets; static int n_sub;
static int* cover;  // how many masks avoid subset i

static int runs_of(uint32_t m) { return __builtin_popcount(m & ~(m << 1)); }

static void gen_subsets(void) {
    n_sub = 0;
    uint32_t lim = L >= 32 ? 0xffffffffu : ((1u << L) - 1u);
    // enumerate k-subsets via Gosper
    int cap = 1; for (int i = 0; i < K; i++) cap = cap * (L - i) / (i + 1);
    subsets = malloc(sizeof(uint32_t) * (cap + 1));
    if (K == 0) { subsets[n_sub++] = 0; return; }
    uint64_t v = (1ull << K) - 1;
    while (v <= lim) {
        subsets[n_sub++] = (uint32_t)v;
        uint64_t c = v & -v, r = v + c;
        v = (((r ^ v) >> 2) / c) | r;
    }
}

static uint32_t random_mask(void) {
    for (;;) {
        uint32_t m = 0; int n = 0;
        // random runs: pick MAXRUNS run starts/lengths
        int nr = 1 + rnd() % MAXRUNS;
        int left = S;
        for (int r = 0; r < nr && left > 0; r++) {
            int len = (r == nr - 1) ? left : 1 + rnd() % left;
            int st = rnd() % (L - len + 1);
            for (int j = 0; j < len; j++) if (!((m >> (st + j)) & 1)) { m |= 1u << (st + j); n++; }
            left = S - n;
        }
        while (n < S) { int p = rnd() % L; if (!((m >> p) & 1)) { m |= 1u << p; n++; } }
        if (runs_of(m) <= MAXRUNS) return m;
    }
}

int main(int argc, char** argv) {
    if (argc < 8) { fprintf(stderr, "usage: L k s C maxruns seed iters\n"); return 2; }
    L = atoi(argv[1]); K = atoi(argv[2]); S = atoi(argv[3]); C = atoi(argv[4]); MAXRUNS = atoi(argv[5]);
    rng_s ^= (uint64_t)atoll(argv[6]) * 0x9E3779B97F4A7C15ull; for (int i = 0; i < 10; i++) rnd();
    long iters = atol(argv[7]);
    gen_subsets();
    cover = calloc(n_sub, sizeof(int));
    uint32_t* M = malloc(sizeof(uint32_t) * C);
    for (int c = 0; c < C; c++) M[c] = random_mask();
    int unc = 0;
    for (int i = 0; i < n_sub; i++) { for (int c = 0; c < C; c++) if (!(M[c] & subsets[i])) cover[i]++; if (!cover[i]) unc++; }
    double T = 2.0;
    int best = unc;
    for (long it = 0; it < iters && unc > 0; it++) {
        T = 2.0 * (1.0 - (double)it / iters) + 0.05;
        int c = rnd() % C;
        uint32_t old = M[c], nw;
        if (rnd() % 8 == 0) nw = random_mask();
        else {  // move one position: remove a set bit, add an unset bit
            for (;;) {
                int a = rnd() % L, b = rnd() % L;
                if (!((old >> a) & 1) || ((old >> b) & 1)) continue;
                nw = (old & ~(1u << a)) | (1u << b);
                break;
            }
            if (runs_of(nw) > MAXRUNS) continue;
        }
        int delta = 0;
        for (int i = 0; i < n_sub; i++) {
            int o = !(old & subsets[i]), n = !(nw & subsets[i]);
            if (o == n) continue;
            if (o && cover[i] == 1) delta++;
            if (n && cover[i] == 0) delta--;
        }
        if (delta <= 0 || (double)(rnd() % 1000000) / 1e6 < exp(-delta / T)) {
            for (int i = 0; i < n_sub; i++) {
                int o = !(old & subsets[i]), n = !(nw & subsets[i]);
                cover[i] += n - o;
            }
            M[c] = nw; unc += delta;
            if (unc < best) best = unc;
        }
    }
    if (unc == 0) {
        printf("%d %d %d %d", L, K, S, C);
        for (int c = 0; c < C; c++) printf(" 0x%x", M[c]);
        printf("\n");
        return 0;
    }
    fprintf(stderr, "no design: L=%d k=%d s=%d C=%d best uncovered=%d\n", L, K, S, C, best);
    return 1;
}
