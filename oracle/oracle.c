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
 *   orc_search_seeded  pigeonhole (k+1 seeds) hash index over the library with
 *                      2-bit packed XOR/popcount verification - a reasonable
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
    const uint32_t *badpref; /* prefix count of non-ACGT genome chars */
    /* seeded */
    const uint32_t *dir;   /* (k+1) tables */
    const uint32_t *ent;
    const uint64_t *packed; /* 2n entries, 2 bits per base, base j at bits 2j */
    const uint32_t *seed_start, *seed_len, *dir_off, *ent_off;
    int mode;
    /* work split */
    uint64_t g_lo, g_hi; /* global position range handled by this thread */
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

static void *brute_worker(void *arg) {
    job *J = (job *)arg;
    const uint32_t L = J->L;
    for (uint32_t c = 0; c < J->n_contigs; c++) {
        uint64_t a = J->coff[c], b = J->coff[c + 1];
        if (b - a < L) continue;
        uint64_t lo = a > J->g_lo ? a : J->g_lo;
        uint64_t hi = (b - L + 1) < J->g_hi ? (b - L + 1) : J->g_hi;
        for (uint64_t p = lo; p < hi; p++) {
            if (J->badpref[p + L] - J->badpref[p]) continue; /* rule 5 */
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

static inline uint64_t pack_window(const char *g, uint32_t L) {
    uint64_t w = 0;
    for (uint32_t j = 0; j < L; j++) w |= (uint64_t)code_of((unsigned char)g[j]) << (2 * j);
    return w;
}

static void *seeded_worker(void *arg) {
    job *J = (job *)arg;
    const uint32_t L = J->L;
    const int S = J->k + 1;
    const uint64_t lowmask = 0x5555555555555555ull;
    for (uint32_t c = 0; c < J->n_contigs; c++) {
        uint64_t a = J->coff[c], b = J->coff[c + 1];
        if (b - a < L) continue;
        uint64_t lo = a > J->g_lo ? a : J->g_lo;
        uint64_t hi = (b - L + 1) < J->g_hi ? (b - L + 1) : J->g_hi;
        for (uint64_t p = lo; p < hi; p++) {
            if (J->badpref[p + L] - J->badpref[p]) continue;
            uint64_t w = pack_window(J->genome + p, L);
            for (int j = 0; j < S; j++) {
                uint32_t key = (uint32_t)((w >> (2 * J->seed_start[j])) & ((1ull << (2 * J->seed_len[j])) - 1));
                const uint32_t *dir = J->dir + J->dir_off[j];
                const uint32_t *ent = J->ent + J->ent_off[j];
                for (uint32_t e = dir[key]; e < dir[key + 1]; e++) {
                    uint32_t id = ent[e]; /* entry = spacer*2 + strand */
                    uint64_t x = w ^ J->packed[id];
                    uint64_t m2 = (x | (x >> 1)) & lowmask;
                    const signed char *q = ((id & 1) ? J->rc : J->fw) + (size_t)(id >> 1) * L;
                    /* non-ACGT spacer characters were packed as 0; force their mismatch bit */
                    uint32_t mask = 0;
                    for (uint32_t t = 0; t < L; t++)
                        if ((m2 >> (2 * t) & 1ull) || q[t] < 0) mask |= 1u << t;
                    if (__builtin_popcount(mask) > J->k) continue;
                    /* report from the first exact seed only, so each hit appears once */
                    int first = -1;
                    for (int jj = 0; jj < S && first < 0; jj++) {
                        uint32_t sm = ((J->seed_len[jj] >= 32 ? 0xffffffffu : ((1u << J->seed_len[jj]) - 1u)) << J->seed_start[jj]);
                        if (!(mask & sm)) first = jj;
                    }
                    if (first == j) emit(J, id >> 1, p, (int)(id & 1), mask);
                }
            }
        }
    }
    return NULL;
}

static int64_t run(const char *genome, const uint64_t *coff, uint32_t n_contigs, const char *spacers,
                   uint32_t n, uint32_t L, int k, orc_hit *out, int64_t cap, int nthreads, int seeded) {
    if (L == 0 || L > 32 || k < 0 || n_contigs == 0) return -1;
    if ((int)L <= k) return 0; /* rule 7 */
    uint64_t G = coff[n_contigs];
    if (nthreads < 1) nthreads = 1;
    signed char *fw = (signed char *)malloc((size_t)n * L + 1);
    signed char *rc = (signed char *)malloc((size_t)n * L + 1);
    for (uint32_t s = 0; s < n; s++)
        for (uint32_t j = 0; j < L; j++) {
            int c = code_of((unsigned char)spacers[(size_t)s * L + j]);
            fw[(size_t)s * L + j] = (signed char)c;
            rc[(size_t)s * L + (L - 1 - j)] = (signed char)(c < 0 ? -1 : 3 - c);
        }
    uint32_t *badpref = (uint32_t *)malloc((G + 1) * sizeof(uint32_t));
    badpref[0] = 0;
    for (uint64_t i = 0; i < G; i++) badpref[i + 1] = badpref[i] + (code_of((unsigned char)genome[i]) < 0);

    uint32_t *dir = NULL, *ent = NULL;
    uint64_t *packed = NULL;
    uint32_t seed_start[8], seed_len[8], dir_off[9], ent_off[9];
    const int S = k + 1;
    if (seeded) {
        if (S > 8) return -1;
        uint64_t dtot = 0;
        for (int j = 0; j < S; j++) {
            seed_start[j] = (uint32_t)((uint64_t)j * L / S);
            seed_len[j] = (uint32_t)((uint64_t)(j + 1) * L / S) - seed_start[j];
            if (seed_len[j] > 12) seed_len[j] = 12; /* key prefix; verification is full-length */
            dir_off[j] = (uint32_t)dtot;
            dtot += (1ull << (2 * seed_len[j])) + 1;
            ent_off[j] = (uint32_t)((uint64_t)j * 2 * n);
        }
        dir = (uint32_t *)calloc(dtot, sizeof(uint32_t));
        ent = (uint32_t *)malloc((size_t)S * 2 * n * sizeof(uint32_t) + 4);
        packed = (uint64_t *)malloc((size_t)2 * n * sizeof(uint64_t) + 8);
        for (uint32_t id = 0; id < 2 * n; id++) {
            const signed char *q = ((id & 1) ? rc : fw) + (size_t)(id >> 1) * L;
            uint64_t w = 0;
            for (uint32_t j = 0; j < L; j++) w |= (uint64_t)(q[j] < 0 ? 0 : q[j]) << (2 * j);
            packed[id] = w;
        }
        for (int j = 0; j < S; j++) {
            uint32_t *d = dir + dir_off[j];
            uint64_t kmask = (1ull << (2 * seed_len[j])) - 1;
            uint32_t nk = (uint32_t)kmask + 1;
            for (int pass = 0; pass < 2; pass++) {
                for (uint32_t id = 0; id < 2 * n; id++) {
                    const signed char *q = ((id & 1) ? rc : fw) + (size_t)(id >> 1) * L;
                    int has_n = 0;
                    for (uint32_t t = 0; t < seed_len[j]; t++) has_n |= q[seed_start[j] + t] < 0;
                    if (has_n) continue; /* a seed containing N can never be exact */
                    uint32_t key = (uint32_t)((packed[id] >> (2 * seed_start[j])) & kmask);
                    if (pass == 0) d[key + 1]++;
                    else ent[ent_off[j] + d[key]++] = id;
                }
                if (pass == 0)
                    for (uint32_t x = 0; x < nk; x++) d[x + 1] += d[x];
                else {
                    for (uint32_t x = nk; x > 0; x--) d[x] = d[x - 1];
                    d[0] = 0;
                }
            }
        }
    }

    job *jobs = (job *)calloc((size_t)nthreads, sizeof(job));
    pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        job *J = &jobs[t];
        J->genome = genome; J->coff = coff; J->n_contigs = n_contigs;
        J->spacers = spacers; J->n = n; J->L = L; J->k = k;
        J->fw = fw; J->rc = rc; J->badpref = badpref;
        J->dir = dir; J->ent = ent; J->packed = packed;
        J->seed_start = seed_start; J->seed_len = seed_len; J->dir_off = dir_off; J->ent_off = ent_off;
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
    free(jobs); free(th); free(fw); free(rc); free(badpref); free(dir); free(ent); free(packed);
    return total;
}

int64_t orc_search_brute(const char *genome, const uint64_t *coff, uint32_t n_contigs, const char *spacers,
                         uint32_t n, uint32_t L, int k, orc_hit *out, int64_t cap, int nthreads) {
    return run(genome, coff, n_contigs, spacers, n, L, k, out, cap, nthreads, 0);
}

int64_t orc_search_seeded(const char *genome, const uint64_t *coff, uint32_t n_contigs, const char *spacers,
                          uint32_t n, uint32_t L, int k, orc_hit *out, int64_t cap, int nthreads) {
    return run(genome, coff, n_contigs, spacers, n, L, k, out, cap, nthreads, 1);
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
