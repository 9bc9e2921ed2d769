// bc_device.cuh - device-side data layout and the shared hit-emission path.
//
// HBM layout:
//   genome   three 1-bit planes H (code>>1), Lo (code&1), B (1 = not ACGT), one bit per base,
//            LSB-first in uint32 words; contigs are laid out back to back with ONE separator
//            base (B=1) after each contig, so a window or PAM that would cross a contig
//            boundary touches a B bit.  "dev position" = position in this layout;
//            gpos = dev position - contig index.
//   library  entry e = 2*spacer + strand; qh/ql hold the QUERY planes (the spacer itself for
//            '+', its reverse complement for '-'), bit j = base j of the query.
//   index    for every seed combination c a direct-address directory dir[c][key] into a
//            key-sorted copy of the entries (ent_hl = {qh,ql}, ent_id = e).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/barcoder_b200.h"

#define BC_MAX_COMBOS 40
#define BC_MAX_PIECES 6
#define BC_MAX_BLOCKS 8
#define BC_KEY_MAX_NT 12

// One seed combination = one set of query positions (the key) under which the library is indexed
// and the genome windows are probed / sorted.  The key positions are at most BC_MAX_PIECES runs;
// the remaining ("rem") positions are at most BC_MAX_PIECES + 1 runs.
struct ComboDesc {
    uint32_t blocks_mask;            // block schemes: which of the b blocks form this combination (0 for designs)
    uint32_t key_mask;               // query-position bits covered by the key pieces
    uint32_t dir_off;                // first directory slot of this combination
    uint8_t n_pieces;
    uint8_t key_nt;                  // total bases in the key
    uint8_t start[BC_MAX_PIECES];    // first base of each key piece
    uint8_t len[BC_MAX_PIECES];      // bases in each key piece
    uint8_t n_rem;                   // runs of non-key positions (compact join path)
    uint8_t rem_nt;                  // L - key_nt
    uint8_t rstart[BC_MAX_PIECES + 1];
    uint8_t rlen[BC_MAX_PIECES + 1];
    uint8_t top_bits;                // compact join path: key bits that select the pass-A bin
    uint8_t pad;
    uint32_t bin_off;                // compact join path: first pass-A bin of this combination
};

struct SearchParams {
    // genome planes
    const uint32_t* H;
    const uint32_t* Lo;
    const uint32_t* B;
    const uint32_t* start_dev;  // [n_contigs+1] dev position of each contig start; last = n_pos
    uint32_t n_pos;             // dev positions (bases + separators)
    uint32_t n_words;           // words per plane (whole tiles + padding, bc_api.cu)
    uint32_t pos_begin, pos_end;  // window start positions this context scans (genome-range sharding)
    uint32_t slot_lo, slot_hi;    // directory slots this context owns (slot-range sharding; everything = [0, all slots))
    uint32_t bin_aligned;         // compact join: slot_lo / slot_hi lie on pass-A bin boundaries (bc_build_index)
    uint32_t own_hash;            // ownership order: 0 = combination order, 1 = cyclic from a hash of (position, entry) (bc_owns)
    uint32_t n_contigs;
    // library
    const uint32_t* sn;         // [n] spacer-orientation mask of non-ACGT spacer characters
    uint32_t lib_has_n;
    uint32_t L, k, b, n_combos;
    uint32_t block_mask[BC_MAX_BLOCKS];
    ComboDesc combo[BC_MAX_COMBOS];
    // index
    const uint32_t* dir;
    const uint2* ent_hl;
    const uint32_t* ent_id;
    unsigned long long dir_entries;  // index entries over all combinations
    const uint32_t* pdir;            // probe path: packed directory (start | count << 26), one load per probe; or null
    const uint32_t* ent_h;           // probe path: H planes of the entries (ent_hl[e].x) as a dense array, padded to 4; or null
    // PAM
    uint32_t P, pam_dir, pam_flags;
    uint32_t pam_sets[8];       // per PAM position: allowed set over {A=1,C=2,G=4,T=8}
    uint32_t gate_first;        // 1: evaluate the PAM gate before any index work
    uint32_t window_sort;       // bucket-join path: 0 auto, 1 direct scatter, 2 radix scatter (BC_PARAM_WINDOW_SORT)
    uint32_t join_chunk;        // bucket-join path: max window positions per pass over the genome, 0 = as many as fit (BC_PARAM_JOIN_CHUNK)
    // output
    bc_hit* hits;
    unsigned long long* count;  // [0] hits, [1] candidates, [2] probes, [3] next tile of the probe kernel, [4] queued verify items
    uint4* items;               // compact join: global queue of {window, entry group} items for k_cfinish
    unsigned long long item_cap;
    unsigned long long cap;
    uint32_t count_candidates;
    uint32_t spacer_id_base;
};

// slot-range sharding: does a combination own any slot of [lo, hi)?
__device__ __forceinline__ bool bc_combo_in_range(const ComboDesc& cd, uint32_t lo, uint32_t hi) {
    return cd.dir_off < hi && cd.dir_off + (1u << (2u * cd.key_nt)) > lo;
}

__device__ __forceinline__ uint32_t bc_lmask(uint32_t n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }

// L-bit window of a plane starting at dev position pos (needs words pos>>5 and (pos>>5)+1).
__device__ __forceinline__ uint32_t bc_window(const uint32_t* __restrict__ plane, uint32_t pos) {
    uint32_t w = pos >> 5;
    return __funnelshift_r(plane[w], plane[w + 1], pos & 31u);
}

// The key bits of a plane word, packed to the low key_nt bits (ascending position order).
__device__ __forceinline__ uint32_t bc_combo_gather_key(const ComboDesc& cd, uint32_t v) {
    uint32_t out = 0, acc = 0;
    for (uint32_t i = 0; i < cd.n_pieces; i++) {
        const uint32_t len = cd.len[i];
        out |= ((v >> cd.start[i]) & ((1u << len) - 1u)) << acc;
        acc += len;
    }
    return out;
}

// Seed key of a window / query under a combination: the key positions of the hi plane above those
// of the lo plane.  Any bijection works as long as the library index and the genome side agree.
__device__ __forceinline__ uint32_t bc_combo_key(const ComboDesc& cd, uint32_t h, uint32_t l) {
    if (cd.n_pieces == 1) {  // one contiguous run (every k+1-seed block scheme): no loop
        const uint32_t st = cd.start[0], m = (1u << cd.key_nt) - 1u;
        return (((h >> st) & m) << cd.key_nt) | ((l >> st) & m);
    }
    return (bc_combo_gather_key(cd, h) << cd.key_nt) | bc_combo_gather_key(cd, l);
}

// The non-key bits of a plane word, packed to the low rem_nt bits (ascending position order).
__device__ __forceinline__ uint32_t bc_combo_rem(const ComboDesc& cd, uint32_t v) {
    uint32_t out = 0, acc = 0;
    for (uint32_t i = 0; i < cd.n_rem; i++) {
        const uint32_t len = cd.rlen[i];
        out |= ((v >> cd.rstart[i]) & ((1u << len) - 1u)) << acc;
        acc += len;
    }
    return out;
}

// Inverse of bc_combo_rem for a mismatch mask: packed rem bits back to their query positions.
__device__ __forceinline__ uint32_t bc_combo_rem_expand(const ComboDesc& cd, uint32_t r) {
    uint32_t out = 0, acc = 0;
    for (uint32_t i = 0; i < cd.n_rem; i++) {
        const uint32_t len = cd.rlen[i];
        out |= ((r >> acc) & ((1u << len) - 1u)) << cd.rstart[i];
        acc += len;
    }
    return out;
}

__device__ __forceinline__ uint32_t bc_rev_bits(uint32_t m, uint32_t L) { return __brev(m) >> (32 - L); }

// PAM-first gate (BC_PAM_GATE): can a window at dev position `pos` still produce a reportable hit
// on at least one strand?  Evaluated before any index work, so at NGG about 7 windows in 8 are
// dropped up front.  A PAM site that touches a non-ACGT base or a contig end counts as "maybe":
// bc_make_hit decides those exactly (ambiguous PAMs are kept for the host, truncated ones dropped).
struct PamGate {
    uint32_t P, L, right_for_plus;  // right_for_plus: '+' hits have their PAM right of the window
    uint32_t fwd[4];                // bit j set: base code c is allowed at window bit j ('+' strand reading)
    uint32_t rc[4];                 // same for the reverse-complement reading ('-' strand hits)
};

// Build the bit masks from the per-position sets (pam_sets[i] over {A=1,C=2,G=4,T=8}).
__device__ __forceinline__ void bc_gate_init(PamGate& g, uint32_t P, uint32_t L, uint32_t pam_dir,
                                             const uint32_t* pam_sets) {
    g.P = P; g.L = L; g.right_for_plus = pam_dir == 0;
    for (int c = 0; c < 4; c++) g.fwd[c] = g.rc[c] = 0;
    for (uint32_t i = 0; i < P; i++) {
        for (uint32_t c = 0; c < 4; c++) {
            if ((pam_sets[i] >> c) & 1u) {
                g.fwd[c] |= 1u << i;                // PAM position i is window bit i
                g.rc[3u - c] |= 1u << (P - 1 - i);  // read backwards, complemented
            }
        }
    }
}

// All P positions at once: a position is satisfied when its (h, l) code is in its allowed set.
__device__ __forceinline__ bool bc_gate_side(const PamGate& g, const uint32_t* __restrict__ H,
                                             const uint32_t* __restrict__ Lo, const uint32_t* __restrict__ B,
                                             uint32_t a, bool rc) {
    const uint32_t pm = (1u << g.P) - 1u;
    if (bc_window(B, a) & pm) return true;  // ambiguous or contig end: decided later
    const uint32_t h = bc_window(H, a), l = bc_window(Lo, a);
    const uint32_t* m = rc ? g.rc : g.fwd;
    const uint32_t sat = (~h & ~l & m[0]) | (~h & l & m[1]) | (h & ~l & m[2]) | (h & l & m[3]);
    return (sat & pm) == pm;
}

__device__ __forceinline__ bool bc_gate_window(const PamGate& g, const uint32_t* __restrict__ H,
                                               const uint32_t* __restrict__ Lo, const uint32_t* __restrict__ B,
                                               uint32_t pos) {
    // '+' strand hits read the PAM forward on one side, '-' strand hits read it reverse-complemented
    // on the other side
    const uint32_t right = pos + g.L;
    const bool left_ok = pos >= g.P;
    const bool plus = g.right_for_plus ? bc_gate_side(g, H, Lo, B, right, false)
                                       : (left_ok && bc_gate_side(g, H, Lo, B, pos - g.P, false));
    if (plus) return true;
    return g.right_for_plus ? (left_ok && bc_gate_side(g, H, Lo, B, pos - g.P, true))
                            : bc_gate_side(g, H, Lo, B, right, true);
}

// Ownership: several combinations may find the same alignment (every combination whose key
// positions are mismatch-free does).  It is reported by exactly ONE of them, so no dedup pass is
// needed.  The owner is a function of the mismatch mask alone (the same for every combination that
// finds the pair, and known before the position or the entry id are fetched):
//   own_hash = 0   the first combination in index order whose key is mismatch-free;
//   own_hash = 1   the (hash(m) mod n_free)-th of the n_free mismatch-free combinations.
// The plain order gives the low combinations most of the multi-key hits: with the seed directory
// sharded over 8 GPUs, rank 0 then produced 1.66x the average number of records and its merge / D2H
// became the tail, so slot-range sharding uses the hashed pick.  (A hash of (position, entry) would
// balance just as well but needs two random DRAM sectors per candidate before it can decide:
// measured +30 % on the finish kernel.)  m = mismatch mask in QUERY orientation; the caller found the
// pair through combination c, so c's own key is mismatch-free by construction.
__device__ __forceinline__ bool bc_owns(const SearchParams& p, uint32_t c, uint32_t m) {
    if (!p.own_hash) {
        for (uint32_t j = 0; j < c; j++)
            if (!(m & p.combo[j].key_mask)) return false;
        return true;
    }
    uint32_t n_free = 0, mine = 0;
    for (uint32_t j = 0; j < p.n_combos; j++) {
        if (j == c) mine = n_free;
        n_free += (m & p.combo[j].key_mask) ? 0u : 1u;
    }
    uint32_t h = m * 2654435761u;
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    return (uint32_t)(((unsigned long long)h * n_free) >> 32) == mine;
}

// Rare path: a (window, entry) pair passed the popcount filter in seed combination `c`.
// Decides whether this combination owns the hit and annotates the PAM.  Returns false when the
// pair is not to be reported from here.
//   pos   dev position of the window
//   e     library entry (2*spacer + strand)
//   m     mismatch mask in QUERY orientation (bit j = query/window base j differs)
// (bc_make_hit_inl is the body; bc_make_hit, below, is the out-of-line copy the scan kernels call from their rare paths.
// Out of line the SearchParams reference is a generic pointer and every field is a load; k_cfinish, whose whole job
// is this function, inlines the body so that the fields come from the constant bank.)
static __device__ __forceinline__ bool bc_make_hit_inl(const SearchParams& p, uint32_t c, uint32_t pos, uint32_t e,
                                                       uint32_t m, uint4* out) {
    const uint32_t L = p.L;
    const uint32_t strand = e & 1u, sid = e >> 1;
    if (p.lib_has_n) {  // non-ACGT spacer characters mismatch everything (oracle.c rule 6)
        uint32_t nm = p.sn[sid];
        if (strand) nm = bc_rev_bits(nm, L);
        m |= nm;
        if (__popc(m) > (int)p.k) return false;
    }
    if (!bc_owns(p, c, m)) return false;

    // contig of the window
    uint32_t lo = 0, hi = p.n_contigs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (p.start_dev[mid] <= pos) lo = mid; else hi = mid;
    }
    const uint32_t cs = p.start_dev[lo], ce = p.start_dev[lo + 1] - 1;  // [cs, ce) are the contig's bases
    uint32_t meta = strand | ((uint32_t)__popc(m) << 1) | (p.P << 8);
    if (p.P == 0) {
        meta |= BC_META_PAM_OK | BC_META_PAM_FULL;
    } else {
        const bool right = (p.pam_dir == 0) == (strand == 0);
        const long long a = right ? (long long)pos + L : (long long)pos - (long long)p.P;
        if (a >= (long long)cs && a + (long long)p.P <= (long long)ce) {
            uint32_t codes = 0, amb = 0, ok = 1;
            // the P bases of the site as one funnel-shifted window per plane: six independent loads, then bit
            // operations only (base by base the loop was a chain of dependent loads: B, then H and Lo, P times)
            const uint32_t wB = bc_window(p.B, (uint32_t)a), wH = bc_window(p.H, (uint32_t)a), wL = bc_window(p.Lo, (uint32_t)a);
            for (uint32_t i = 0; i < p.P; i++) {
                const uint32_t s = strand ? p.P - 1 - i : i;
                if ((wB >> s) & 1u) { amb = 1; continue; }
                uint32_t code = (((wH >> s) & 1u) << 1) | ((wL >> s) & 1u);
                if (strand) code = 3u - code;
                codes |= code << (2 * i);
                if (!((p.pam_sets[i] >> code) & 1u)) ok = 0;
            }
            meta |= BC_META_PAM_FULL | (codes << 16);
            if (amb) meta |= BC_META_PAM_AMB;
            else if (ok) meta |= BC_META_PAM_OK;
        }
        if ((p.pam_flags & BC_PAM_GATE) && !(meta & (BC_META_PAM_OK | BC_META_PAM_AMB))) return false;
    }
    *out = make_uint4(sid + p.spacer_id_base, pos - lo, strand ? bc_rev_bits(m, L) : m, meta);
    return true;
}

static __device__ __noinline__ bool bc_make_hit(const SearchParams& p, uint32_t c, uint32_t pos, uint32_t e,
                                                uint32_t m, uint4* out) {
    return bc_make_hit_inl(p, c, pos, e, m, out);
}

// Hit records are staged per CTA in shared memory and flushed with ONE global atomic per flush:
// a single-address atomicAdd per hit serialises in L2 (~2.4 ns each, measured) and would cap
// cfg 4 (5.9e7 hits) at ~140 ms on its own.
#ifndef BC_STAGE_CAP
#define BC_STAGE_CAP 1024
#endif

struct HitStage {
    uint32_t n;
    uint32_t base;
    uint32_t pad[2];
    uint4 rec[BC_STAGE_CAP];
};

__device__ __forceinline__ void bc_stage_hit(const SearchParams& p, HitStage* st, const uint4& rec) {
    const uint32_t slot = atomicAdd(&st->n, 1u);
    if (slot < BC_STAGE_CAP) {
        st->rec[slot] = rec;
    } else {  // stage full: fall back to a direct append
        const unsigned long long g = atomicAdd(p.count, 1ull);
        if (g < p.cap) reinterpret_cast<uint4*>(p.hits)[g] = rec;
    }
}

// Called by ALL threads of the CTA (contains barriers).
__device__ __forceinline__ void bc_flush_hits(const SearchParams& p, HitStage* st) {
    __syncthreads();
    const uint32_t n = min(st->n, (uint32_t)BC_STAGE_CAP);
    if (n) {
        __shared__ unsigned long long s_base;
        if (threadIdx.x == 0) s_base = atomicAdd(p.count, (unsigned long long)n);
        __syncthreads();
        const unsigned long long base = s_base;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            if (base + i < p.cap) reinterpret_cast<uint4*>(p.hits)[base + i] = st->rec[i];
        __syncthreads();
        if (threadIdx.x == 0) st->n = 0;
    }
    __syncthreads();
}
