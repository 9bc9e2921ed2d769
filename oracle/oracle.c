/*
 * oracle.c - CPU restatement of barcoder's spacer->genome mismatch search.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under barcoder_b200/ may import, link or
 * execute this file; it is the checker for tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY PIN: the reference delegates the search to the external bowtie 1.3.1
 * binary (environment.yml:26), which is neither vendored under /root/reference
 * nor installed here, and the reference has no tests.  The
 * oracle is therefore pinned against (a) the reference's only data fixture,
 * Example_Libraries/CN-32-zmo.tsv (772 plasmid rows, tests/golden/), and
 * (b) the SURVEY.md section 8c known answer (869 hits, sha256 5dbbc9c4...).
 * Against bowtie itself parity is UNPINNED.
 *
 * What it restates (file:line under /root/reference):
 *   BowtieRunner.py:111-125   bowtie -a -v k --best --tryhard: every end-to-end
 *                             ungapped alignment of every read to both strands
 *                             with <= k mismatches, qualities ignored.
 *   PySamParser.py:26-48      Start = 0-based leftmost '+'-strand position,
 *                             End = Start + L, Strand '+'/'-', Barcode always in
 *                             library orientation, Mismatches = NM.
 *   targets.py:184-190        mismatch positions are counted in spacer
 *                             orientation (get_diff over spacer vs target).
 *   PAMProcessor.py:65-94     PAM slice: '+' -> seq[End:End+P];
 *                             '-' -> revcomp(seq[Start-P:Start]).
 *   targets.py:227-307        direction-aware variant (upstream = 5' side).
 *
 * Search rules (bowtie 1.3.1 manual, -v mode):
 *   1. window length == read length L, no gaps;
 *   2. both strands: a '-' hit at leftmost '+' position p means
 *      revcomp(spacer) aligns to ref[p:p+L];
 *   3. valid iff Hamming distance <= k;
 *   4. a window may not cross a contig boundary;
 *   5. a window touching any non-ACGT reference character is invalid;
 *   6. a non-ACGT character in the spacer mismatches everything;
 *   7. reads with L <= k are skipped (bowtie warns and does not align them);
 *   8. a palindromic spacer is reported on both strands.
 *
 * Two independent search strategies are provided so they can check each other:
 *   orc_search_brute   exhaustive char-by-char comparison (ground truth);
 *   orc_search_seeded  generalised pigeonhole (b blocks, every (b-k)-subset a seed,
 *                      b chosen by a cost model) over a key-sorted copy of the 2-bit
 *                      packed library; rolling genome window; one XOR + POPCNT per
 *                      candidate, positions only for survivors - a reasonable
 *                      multi-threaded CPU implementation, used as cpu_baseline.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint32_t spacer_id; /* index into the caller's spacer array */
    uint32_t gpos;      /* 0-based leftmost '+' position in the concatenated genome */
    uint32_t mm_mask;   /* bit i = mismatch at spacer position i (spacer orientation) */
    uint32_t meta;      /* see include/barcoder_b200.h (BC_META_*) */
} orc_hit;

#define META_STRAND(m) ((m) & 1u)
#define META_NMM_SHIFT 1
#define META_PAM_OK (1u << 3)
#define META_PAM_FULL (1u << 4)
#define META_PAM_AMB (1u << 5)
#define META_PAM_LEN_SHIFT 8
#define META_PAM_CODES_SHIFT 16

#define PAM_FLAG_IUPAC 1u /* expand IUPAC letters in the pattern (extension) */
#define PAM_FLAG_GATE 2u  /* drop hits whose PAM does not match */

static inline int code_of(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

typedef struct {
    orc_hit *v;
    int64_t n, cap;
} hitvec;

static void hv_push(hitvec *h, orc_hit x) {
    if (h->n == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 1024;
        h->v = (orc_hit *)realloc(h->v, (size_t)h->cap * sizeof(orc_hit));
    }
    h->v[h->n++] = x;
}

/* ---- seed scheme of the seeded strategy: generalised pigeonhole.  The L positions are cut into
 * b blocks; an alignment with <= k mismatches leaves >= b-k blocks untouched, so indexing the
 * library under every (b-k)-subset of blocks ("combination") finds all of them.  b = k+1 is the
 * textbook k+1-seed filter; larger b trades more look-ups per window for far fewer candidates.
 * Keys are capped at ORC_KEY_CAP bases (a prefix of the chosen blocks); verification is always
 * over the full length. */
#define ORC_MAX_BLOCKS 8
#define ORC_MAX_COMBOS 70 /* C(8,4) */
#define ORC_KEY_CAP 12

typedef struct {
    uint32_t n_pieces, start[ORC_MAX_BLOCKS], len[ORC_MAX_BLOCKS];
    uint32_t key_nt, key_mask; /* key_mask: query positions covered by the key */
    uint64_t dir_off, ent_off;
} orc_combo;

typedef struct {
    const char *genome;
    const uint64_t *coff;
    uint32_t n_contigs;
    const char *spacers;
    uint32_t n, L;
    int k;
    /* derived */
    const signed char *fw; /* n*L codes, -1 for non-ACGT */
    const signed char *rc; /* n*L codes of the reverse complement */
    int lib_has_n;
    /* seeded */
    const uint32_t *dir;    /* one direct-address table per combination */
    const uint64_t *ent_w;  /* per combination: the 2-bit packed queries in key order */
    const uint32_t *ent_id; /* per combination: entry = spacer*2 + strand, same order */
    const orc_combo *combo;
    int n_combos;
    /* work split */
    uint64_t g_lo, g_hi; /* global position range handled by this thread */
    int c_lo, c_hi;      /* index build: combinations handled by this thread */
    const uint64_t *packed;
    uint32_t *dir_w;
    uint64_t *ent_w_w;
    uint32_t *ent_id_w;
    hitvec out;
} job;

static inline void emit(job *J, uint32_t sid, uint64_t gpos, int strand, uint32_t mask_query) {
    /* mask_query: bit j = mismatch at query (window) position j.  Convert to spacer
     * orientation: for '-' the query is the reverse complement, position j <-> L-1-j. */
    uint32_t m = mask_query;
    if (strand) {
        uint32_t r = 0;
        for (uint32_t j = 0; j < J->L; j++)
            if (m >> j & 1u) r |= 1u << (J->L - 1 - j);
        m = r;
    }
    orc_hit h;
    h.spacer_id = sid;
    h.gpos = (uint32_t)gpos;
    h.mm_mask = m;
    h.meta = (uint32_t)strand | ((uint32_t)__builtin_popcount(m) << META_NMM_SHIFT);
    hv_push(&J->out, h);
}

static inline int is_bad(char c) { return code_of((unsigned char)c) < 0; }

static void *brute_worker(void *arg) {
    job *J = (job *)arg;
    const uint32_t L = J->L;
    for (uint32_t c = 0; c < J->n_contigs; c++) {
        uint64_t a = J->coff[c], b = J->coff[c + 1];
        if (b - a < L) continue;
        uint64_t lo = a > J->g_lo ? a : J->g_lo;
        uint64_t hi = (b - L + 1) < J->g_hi ? (b - L + 1) : J->g_hi;
        uint32_t bad = 0; /* non-ACGT characters inside the current window (rule 5), kept rolling */
        for (uint64_t p = lo; p < hi; p++) {
            if (p == lo) {
                for (uint32_t j = 0; j < L; j++) bad += (uint32_t)is_bad(J->genome[p + j]);
            } else {
                bad += (uint32_t)is_bad(J->genome[p + L - 1]);
                bad -= (uint32_t)is_bad(J->genome[p - 1]);
            }
            if (bad) continue;
            signed char w[32];
            for (uint32_t j = 0; j < L; j++) w[j] = (signed char)code_of((unsigned char)J->genome[p + j]);
            for (uint32_t s = 0; s < J->n; s++) {
                for (int strand = 0; strand < 2; strand++) {
                    const signed char *q = (strand ? J->rc : J->fw) + (size_t)s * L;
                    int mm = 0;
                    uint32_t mask = 0;
                    for (uint32_t j = 0; j < L; j++) {
                        if (q[j] != w[j]) {
                            mask |= 1u << j;
                            if (++mm > J->k) break;
                        }
                    }
                    if (mm <= J->k) emit(J, s, p, strand, mask);
                }
            }
        }
    }
    return NULL;
}

static inline uint32_t combo_key(const orc_combo *cd, uint64_t w) {
    uint32_t key = 0;
    for (uint32_t i = 0; i < cd->n_pieces; i++)
        key = (key << (2 * cd->len[i])) | (uint32_t)((w >> (2 * cd->start[i])) & ((1ull << (2 * cd->len[i])) - 1));
    return key;
}

/* even bits of x (one per base) -> a dense L-bit mask */
static inline uint32_t squeeze_even(uint64_t x) {
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
    x = (x | (x >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)x;
}

static void *seeded_worker(void *arg) {
    job *J = (job *)arg;
    const uint32_t L = J->L;
    const uint64_t lowmask = 0x5555555555555555ull;
    const int k = J->k;
    for (uint32_t c = 0; c < J->n_contigs; c++) {
        uint64_t a = J->coff[c], b = J->coff[c + 1];
        if (b - a < L) continue;
        uint64_t lo = a > J->g_lo ? a : J->g_lo;
        uint64_t hi = (b - L + 1) < J->g_hi ? (b - L + 1) : J->g_hi;
        uint64_t w = 0;   /* rolling 2-bit window, base j at bits 2j (non-ACGT packed as 0) */
        uint32_t bad = 0; /* rolling count of non-ACGT characters in the window */
        for (uint64_t p = lo; p < hi; p++) {
            if (p == lo) {
                for (uint32_t j = 0; j < L; j++) {
                    int cd = code_of((unsigned char)J->genome[p + j]);
                    bad += cd < 0;
                    w |= (uint64_t)(cd < 0 ? 0 : cd) << (2 * j);
                }
            } else {
                int cd = code_of((unsigned char)J->genome[p + L - 1]);
                bad += cd < 0;
                bad -= (uint32_t)is_bad(J->genome[p - 1]);
                w = (w >> 2) | ((uint64_t)(cd < 0 ? 0 : cd) << (2 * (L - 1)));
            }
            if (bad) continue; /* rule 5 */
            for (int j = 0; j < J->n_combos; j++) {
                const orc_combo *cd = &J->combo[j];
                const uint32_t key = combo_key(cd, w);
                const uint32_t *dir = J->dir + cd->dir_off;
                const uint64_t *ew = J->ent_w + cd->ent_off;
                for (uint32_t e = dir[key]; e < dir[key + 1]; e++) {
                    /* popcount first: one XOR, one fold, one POPCNT per candidate */
                    const uint64_t x = w ^ ew[e];
                    const uint64_t m2 = (x | (x >> 1)) & lowmask;
                    if (__builtin_popcountll(m2) > k) continue;
                    /* survivor: positions, forced mismatches of non-ACGT spacer characters, owner */
                    const uint32_t id = J->ent_id[cd->ent_off + e]; /* entry = spacer*2 + strand */
                    uint32_t mask = squeeze_even(m2);
                    if (J->lib_has_n) {
                        const signed char *q = ((id & 1) ? J->rc : J->fw) + (size_t)(id >> 1) * L;
                        for (uint32_t t = 0; t < L; t++)
                            if (q[t] < 0) mask |= 1u << t;
                        if (__builtin_popcount(mask) > k) continue;
                    }
                    /* report from the first combination whose key is untouched, so each hit appears once */
                    int first = -1;
                    for (int jj = 0; jj < J->n_combos && first < 0; jj++)
                        if (!(mask & J->combo[jj].key_mask)) first = jj;
                    if (first == j) emit(J, id >> 1, p, (int)(id & 1), mask);
                }
            }
        }
    }
    return NULL;
}

/* key-order counting sort of the 2n entries under the combinations [c_lo, c_hi) */
static void *index_worker(void *arg) {
    job *J = (job *)arg;
    const uint32_t L = J->L, n = J->n;
    for (int j = J->c_lo; j < J->c_hi; j++) {
        const orc_combo *cd = &J->combo[j];
        uint32_t *d = J->dir_w + cd->dir_off;
        const uint32_t nk = 1u << (2 * cd->key_nt);
        for (int pass = 0; pass < 2; pass++) {
            for (uint32_t id = 0; id < 2 * n; id++) {
                if (J->lib_has_n) {
                    const signed char *q = ((id & 1) ? J->rc : J->fw) + (size_t)(id >> 1) * L;
                    int has_n = 0;
                    for (uint32_t t = 0; t < L; t++) has_n |= (q[t] < 0) && (cd->key_mask >> t & 1u);
                    if (has_n) continue; /* a seed containing a non-ACGT character can never be exact */
                }
                const uint32_t key = combo_key(cd, J->packed[id]);
                if (pass == 0) d[key + 1]++;
                else {
                    const uint32_t at = d[key]++;
                    J->ent_w_w[cd->ent_off + at] = J->packed[id];
                    J->ent_id_w[cd->ent_off + at] = id;
                }
            }
            if (pass == 0)
                for (uint32_t x = 0; x < nk; x++) d[x + 1] += d[x];
            else {
                for (uint32_t x = nk; x > 0; x--) d[x] = d[x - 1];
                d[0] = 0;
            }
        }
    }
    return NULL;
}

static int make_scheme(uint32_t L, int k, int b, orc_combo *combo, double *cand_per_pair) {
    if (b < k + 1 || b > ORC_MAX_BLOCKS || (uint32_t)b > L) return 0;
    uint32_t bstart[ORC_MAX_BLOCKS + 1];
    for (int j = 0; j <= b; j++) bstart[j] = (uint32_t)((uint64_t)j * L / (uint64_t)b);
    int nc = 0;
    double cand = 0;
    for (uint32_t mask = 0; mask < (1u << b); mask++) {
        if (__builtin_popcount(mask) != b - k) continue;
        if (nc == ORC_MAX_COMBOS) return 0;
        orc_combo *cd = &combo[nc++];
        memset(cd, 0, sizeof *cd);
        uint32_t budget = ORC_KEY_CAP;
        for (int j = 0; j < b && budget; j++) {
            if (!(mask >> j & 1u)) continue;
            uint32_t len = bstart[j + 1] - bstart[j];
            if (len > budget) len = budget;
            cd->start[cd->n_pieces] = bstart[j];
            cd->len[cd->n_pieces] = len;
            cd->n_pieces++;
            cd->key_mask |= ((len >= 32 ? 0xffffffffu : ((1u << len) - 1u))) << bstart[j];
            cd->key_nt += len;
            budget -= len;
        }
        cand += 1.0 / (double)(1ull << (2 * cd->key_nt));
    }
    *cand_per_pair = cand;
    return nc;
}

static int64_t run(const char *genome, const uint64_t *coff, uint32_t n_contigs, const char *spacers,
                   uint32_t n, uint32_t L, int k, orc_hit *out, int64_t cap, int nthreads, int seeded, int blocks) {
    if (L == 0 || L > 32 || k < 0 || n_contigs == 0) return -1;
    if ((int)L <= k) return 0; /* rule 7 */
    uint64_t G = coff[n_contigs];
    if (nthreads < 1) nthreads = 1;
    signed char *fw = (signed char *)malloc((size_t)n * L + 1);
    signed char *rc = (signed char *)malloc((size_t)n * L + 1);
    int lib_has_n = 0;
    for (uint32_t s = 0; s < n; s++)
        for (uint32_t j = 0; j < L; j++) {
            int c = code_of((unsigned char)spacers[(size_t)s * L + j]);
            lib_has_n |= c < 0;
            fw[(size_t)s * L + j] = (signed char)c;
            rc[(size_t)s * L + (L - 1 - j)] = (signed char)(c < 0 ? -1 : 3 - c);
        }

    uint32_t *dir = NULL, *ent_id = NULL;
    uint64_t *packed = NULL, *ent_w = NULL;
    orc_combo *combo = (orc_combo *)calloc(ORC_MAX_COMBOS, sizeof(orc_combo));
    int n_combos = 0;
    job *jobs = (job *)calloc((size_t)nthreads, sizeof(job));
    pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    if (seeded) {
        /* pick b: look-ups cost a cache miss each, candidates a few cycles, index entries a scatter */
        int best_b = 0;
        double best_cost = 0;
        for (int b = k + 1; b <= k + 4; b++) {
            if (blocks && b != blocks) continue;
            double cand;
            int nc = make_scheme(L, k, b, combo, &cand);
            if (!nc) continue;
            double cost = (double)G * nc * 20.0 + (double)G * 2.0 * n * cand * 1.0 + 2.0 * n * nc * 40.0 * nthreads;
            if (!best_b || cost < best_cost) { best_b = b; best_cost = cost; }
        }
        if (!best_b) { free(fw); free(rc); free(combo); free(jobs); free(th); return -1; }
        double cand;
        n_combos = make_scheme(L, k, best_b, combo, &cand);
        uint64_t dtot = 0;
        for (int j = 0; j < n_combos; j++) {
            combo[j].dir_off = dtot;
            dtot += (1ull << (2 * combo[j].key_nt)) + 1;
            combo[j].ent_off = (uint64_t)j * 2 * n;
        }
        dir = (uint32_t *)calloc(dtot, sizeof(uint32_t));
        ent_w = (uint64_t *)malloc(((size_t)n_combos * 2 * n + 1) * sizeof(uint64_t));
        ent_id = (uint32_t *)malloc(((size_t)n_combos * 2 * n + 1) * sizeof(uint32_t));
        packed = (uint64_t *)malloc((size_t)2 * n * sizeof(uint64_t) + 8);
        for (uint32_t id = 0; id < 2 * n; id++) {
            const signed char *q = ((id & 1) ? rc : fw) + (size_t)(id >> 1) * L;
            uint64_t w = 0;
            for (uint32_t j = 0; j < L; j++) w |= (uint64_t)(q[j] < 0 ? 0 : q[j]) << (2 * j);
            packed[id] = w;
        }
        int nt_idx = nthreads < n_combos ? nthreads : n_combos;
        for (int t = 0; t < nt_idx; t++) {
            job *J = &jobs[t];
            J->n = n; J->L = L; J->fw = fw; J->rc = rc; J->lib_has_n = lib_has_n;
            J->combo = combo; J->packed = packed; J->dir_w = dir; J->ent_w_w = ent_w; J->ent_id_w = ent_id;
            J->c_lo = (int)((int64_t)n_combos * t / nt_idx);
            J->c_hi = (int)((int64_t)n_combos * (t + 1) / nt_idx);
            pthread_create(&th[t], NULL, index_worker, J);
        }
        for (int t = 0; t < nt_idx; t++) pthread_join(th[t], NULL);
    }

    for (int t = 0; t < nthreads; t++) {
        job *J = &jobs[t];
        memset(J, 0, sizeof *J);
        J->genome = genome; J->coff = coff; J->n_contigs = n_contigs;
        J->spacers = spacers; J->n = n; J->L = L; J->k = k;
        J->fw = fw; J->rc = rc; J->lib_has_n = lib_has_n;
        J->dir = dir; J->ent_w = ent_w; J->ent_id = ent_id; J->combo = combo; J->n_combos = n_combos;
        J->g_lo = G * (uint64_t)t / (uint64_t)nthreads;
        J->g_hi = G * (uint64_t)(t + 1) / (uint64_t)nthreads;
        pthread_create(&th[t], NULL, seeded ? seeded_worker : brute_worker, J);
    }
    int64_t total = 0;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        for (int64_t i = 0; i < jobs[t].out.n; i++, total++)
            if (total < cap) out[total] = jobs[t].out.v[i];
        free(jobs[t].out.v);
    }
    free(jobs); free(th); free(fw); free(rc); free(dir); free(ent_w); free(ent_id); free(packed); free(combo);
    return total;
}

int64_t orc_search_brute(const char *genome, const uint64_t *coff, uint32_t n_contigs, const char *spacers,
                         uint32_t n, uint32_t L, int k, orc_hit *out, int64_t cap, int nthreads) {
    return run(genome, coff, n_contigs, spacers, n, L, k, out, cap, nthreads, 0, 0);
}

int64_t orc_search_seeded(const char *genome, const uint64_t *coff, uint32_t n_contigs, const char *spacers,
                          uint32_t n, uint32_t L, int k, orc_hit *out, int64_t cap, int nthreads) {
    return run(genome, coff, n_contigs, spacers, n, L, k, out, cap, nthreads, 1, 0);
}

/* same, with the number of pigeonhole blocks forced (k+1 <= blocks <= k+4); 0 = choose */
int64_t orc_search_seeded_b(const char *genome, const uint64_t *coff, uint32_t n_contigs, const char *spacers,
                            uint32_t n, uint32_t L, int k, orc_hit *out, int64_t cap, int nthreads, int blocks) {
    return run(genome, coff, n_contigs, spacers, n, L, k, out, cap, nthreads, 1, blocks);
}

/* IUPAC letter -> 4-bit set over {A=1,C=2,G=4,T=8}; 0 = matches nothing. */
static unsigned iupac_set(char c, unsigned flags) {
    switch (c) {
        case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8;
        case 'N': return 15;
        default: break;
    }
    if (!(flags & PAM_FLAG_IUPAC)) return 0; /* reference: only N is expanded (PAMProcessor.py:7, targets.py:224) */
    switch (c) {
        case 'R': return 1 | 4; case 'Y': return 2 | 8; case 'S': return 2 | 4; case 'W': return 1 | 8;
        case 'K': return 4 | 8; case 'M': return 1 | 2; case 'B': return 14; case 'D': return 13;
        case 'H': return 11; case 'V': return 7;
        default: return 0;
    }
}

/*
 * Fill the PAM bits of meta for each hit, in place; returns the number of hits
 * kept (== nhits unless PAM_FLAG_GATE).  direction: 0 = downstream (3' of the
 * protospacer in spacer orientation; PAMProcessor.py:65-94, targets.py:227-263),
 * 1 = upstream (5' side; targets.py:266-307).
 *
 * C-ABI-level semantics (shared with the CUDA path): pam_full = all P bases lie
 * inside the hit's contig; pam_amb = at least one of them is non-ACGT; pam_ok is
 * only ever set for full, unambiguous PAMs.  Truncated / ambiguous PAMs are
 * resolved by the Python host with the reference's string rules.  With GATE,
 * hits that are full, unambiguous and do not match - or are not full - are dropped.
 */
int64_t orc_annotate_pam(orc_hit *hits, int64_t nhits, const char *genome, const uint64_t *coff,
                         uint32_t n_contigs, uint32_t L, const char *pam, int direction, unsigned flags) {
    uint32_t P = (uint32_t)strlen(pam);
    if (P > 8) return -1;
    unsigned sets[8];
    for (uint32_t i = 0; i < P; i++) sets[i] = iupac_set(pam[i], flags);
    int64_t kept = 0;
    for (int64_t h = 0; h < nhits; h++) {
        orc_hit x = hits[h];
        uint32_t meta = x.meta & 7u; /* strand + nmm */
        int strand = (int)META_STRAND(meta);
        meta |= P << META_PAM_LEN_SHIFT;
        if (P == 0) {
            meta |= META_PAM_OK | META_PAM_FULL;
        } else {
            /* contig of the hit */
            uint32_t lo = 0, hi = n_contigs;
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) / 2;
                if (coff[mid] <= x.gpos) lo = mid; else hi = mid;
            }
            int64_t cs = (int64_t)coff[lo], ce = (int64_t)coff[lo + 1];
            /* side of the '+'-strand window the PAM sits on: right if (down,+) or (up,-) */
            int right = (direction == 0) == (strand == 0);
            int64_t a = right ? (int64_t)x.gpos + L : (int64_t)x.gpos - P;
            int full = a >= cs && a + P <= ce;
            int amb = 0, ok = 1;
            uint32_t codes = 0;
            if (full) {
                for (uint32_t i = 0; i < P; i++) {
                    /* PAM position i in spacer orientation */
                    int c = strand == 0 ? code_of((unsigned char)genome[a + i])
                                        : code_of((unsigned char)genome[a + P - 1 - i]);
                    if (c < 0) { amb = 1; continue; }
                    if (strand) c = 3 - c;
                    codes |= (uint32_t)c << (2 * i);
                    if (!(sets[i] >> c & 1u)) ok = 0;
                }
                meta |= META_PAM_FULL;
                if (amb) meta |= META_PAM_AMB;
                else if (ok) meta |= META_PAM_OK;
                meta |= codes << META_PAM_CODES_SHIFT;
            }
            if ((flags & PAM_FLAG_GATE) && !(meta & META_PAM_OK) && !(meta & META_PAM_AMB)) continue;
        }
        x.meta = meta;
        hits[kept++] = x;
    }
    return kept;
}
