// bc_cjoin.cu - K3-cjoin: the bucket join of bc_join.cu with 8-byte window records, for spacers
// short enough that a record {dev position, x} holds everything the join needs (L <= ~20: CRISPR
// guides).  Replaces bowtie align (BowtieRunner.py:104-141) for dense libraries (cfg 3/4).
//
// Why a second form: at cfg 4 the step is the sort of ~10^9 (window, combination) records plus the
// all-pairs verification inside every slot.  Longer keys cut the pairs geometrically (9 nt: 3.6x
// fewer than 8 nt, 10 nt: 11x) but only pay if the sort stays cheap, and the sort is HBM traffic:
//   * a record carries only what is NOT implied by its place: the low key bits (until pass B has
//     used them) and the non-key ("rem") bit planes of the window, Rh | Rl << rem_nt - the key
//     bits are the slot.  8 bytes instead of 16 through every pass and into the verify kernel;
//   * the library index stores the same rem planes per entry, so verification is still
//     2 LOP3 + 1 POPC per pair and never reassembles the window;
//   * two shared-memory radix passes handle keys up to 11 nt (22 bits = up to 10 top bits in pass
//     A, up to 12 low bits in pass B), with 8192-record chunks so runs stay >= 64 bytes
//     (bench_kernels/sort_lab.cu: direct 8-byte scatters cost 25 ps per record, staged runs 3-5 ps).
//
//   k_cbincount     bin totals per combination from shared-memory histograms           -> bin_start (scan)
//   k_cbin          pass A: windows -> records grouped by bin (top key bits), runs      -> tmp
//   k_cslotcount    slot histogram of the binned records (pieces of a bin in shared mem) -> gdir (scan)
//   k_cplace_bulk   pass B: bin regions -> final slots (low key bits), bulk-async loads  -> gwin
//   k_ctile_*       tile list: one descriptor per (slot, <= 128 windows)
//   k_cverify<K>    first level: pipelined warp-tiles against the slot's library bucket  -> item queue
//   k_cfinish       second level + hit resolution (ownership, PAM, records)              -> hits
//   (k_ccount / k_cplace: the round-1 forms - one global RED per record, LDG-staged pass B - kept for A/B runs;
//    the library index is built with the same passes, template parameter LIB)
#include "bc_join.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bc_kernels.h"

struct CBucketParams {
    const uint32_t* H;
    const uint32_t* Lo;
    const uint32_t* B;
    const uint32_t* lib_dir;      // library directory (to skip windows whose bucket is empty)
    uint32_t pos_begin, pos_end;  // dev positions handled by this pass over the genome
    uint32_t n_words;             // words per plane
    uint32_t L, n_combos, prune, gate_first;
    uint32_t slot_lo, slot_hi;    // slot-range sharding
    uint32_t bin_aligned;         // slot_lo / slot_hi lie on pass-A bin boundaries: ownership is a test on the bin
    uint32_t P, pam_dir, pam_sets[8];
    // library-side sort (index build): entries instead of genome windows
    const uint32_t* qh;
    const uint32_t* ql;
    const uint32_t* sn;
    uint32_t n_entries, lib_has_n;
    ComboDesc combo[BC_MAX_COMBOS];
};

#define CJ_THREADS 512
#define CJ_ITEMS 16
#define CJ_CHUNK (CJ_THREADS * CJ_ITEMS)  // 8192 records per shared-memory sort
#define CJ_MAX_BINS 1024                  // pass-A bins per combination (top_bits <= 10)
#define CJ_MAX_SUB 4096                   // pass-B sub-slots per bin (low key bits <= 12)

// --------------------------------------------------------------------------------- permutation
// Key and rem planes of a window are bit gathers of its H / Lo planes.  Done run by run they cost
// ~300 instructions per (window, combination) and made pass A issue-bound (ncu: 11.5e9 warp
// instructions for 1.2e9 records); done through a byte-wise lookup table they cost ~10 per plane.
// lut[c][b][v] = contribution of byte b (value v) of a plane word to the PERMUTED word of
// combination c: key positions packed above the rem positions, both in ascending order.
// Round 2: two tables of 2^11 entries instead of four of 2^8 (the compact path only takes spacers of
// at most 22 nt): two look-ups per plane instead of three.  Pass A is bound by its shared-memory
// traffic (ncu: l1tex 79 %), and a random-index look-up costs ~3.4 wavefronts.
#ifndef CJ_LUT_BITS
#define CJ_LUT_BITS 11
#endif
#define CJ_LUT_TABS (CJ_LUT_BITS == 8 ? 4 : 2)
#define CJ_LUT_WORDS (CJ_LUT_TABS << CJ_LUT_BITS)
__global__ void k_clut_build(const __grid_constant__ CBucketParams gp, uint32_t* __restrict__ lut) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= gp.n_combos * CJ_LUT_WORDS) return;
    const ComboDesc& cd = gp.combo[g / CJ_LUT_WORDS];
    const uint32_t w = g % CJ_LUT_WORDS;
    const uint32_t word = (w & ((1u << CJ_LUT_BITS) - 1u)) << (CJ_LUT_BITS * (w >> CJ_LUT_BITS));
    lut[g] = (bc_combo_gather_key(cd, word) << cd.rem_nt) | bc_combo_rem(cd, word);
}

// v = an L-bit plane word (bits above L are zero)
__device__ __forceinline__ uint32_t cj_perm(const uint32_t* s_lut, uint32_t v, bool wide) {
#if CJ_LUT_BITS == 8
    uint32_t out = s_lut[v & 255u] | s_lut[256u + ((v >> 8) & 255u)] | s_lut[512u + ((v >> 16) & 255u)];
    if (wide) out |= s_lut[768u + (v >> 24)];  // spacers longer than 24 nt
    return out;
#else
    (void)wide;
    return s_lut[v & ((1u << CJ_LUT_BITS) - 1u)] | s_lut[(1u << CJ_LUT_BITS) + (v >> CJ_LUT_BITS)];
#endif
}

// ------------------------------------------------------------------------------------- count
// One RED per (window, combination).  Grid (x = genome chunks, y = combination): the CTAs of one
// combination run together, so its directory stays in L2.
template <bool LIB>
__global__ void __launch_bounds__(256) k_ccount(const __grid_constant__ CBucketParams gp, const uint32_t* __restrict__ lut,
                                                uint32_t* __restrict__ gdir) {
    __shared__ uint32_t s_lut[CJ_LUT_WORDS];
    const uint32_t lm = bc_lmask(gp.L);
    const ComboDesc& cd = gp.combo[blockIdx.y];
    if (!bc_combo_in_range(cd, gp.slot_lo, gp.slot_hi)) return;
    for (uint32_t i = threadIdx.x; i < CJ_LUT_WORDS; i += blockDim.x) s_lut[i] = lut[blockIdx.y * CJ_LUT_WORDS + i];
    __syncthreads();
    const uint32_t rem_nt = cd.rem_nt, key_nt = cd.key_nt;
    const bool wide = gp.L > 24;
    if (LIB) {  // library entries (index build)
        for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < gp.n_entries; e += gridDim.x * blockDim.x) {
            if (gp.lib_has_n) {
                uint32_t nm = gp.sn[e >> 1];
                if (e & 1u) nm = bc_rev_bits(nm, gp.L);
                if (nm & cd.key_mask) continue;  // a seed containing a non-ACGT base is never exact
            }
            const uint32_t ph = cj_perm(s_lut, gp.qh[e], wide), pl = cj_perm(s_lut, gp.ql[e], wide);
            const uint32_t slot = cd.dir_off + (((ph >> rem_nt) << key_nt) | (pl >> rem_nt));
            if (slot < gp.slot_lo || slot >= gp.slot_hi) continue;
            atomicAdd(&gdir[slot], 1u);
        }
        return;
    }
    PamGate gate;
    bc_gate_init(gate, gp.P, gp.L, gp.pam_dir, gp.pam_sets);
    for (uint32_t pos = gp.pos_begin + blockIdx.x * blockDim.x + threadIdx.x; pos < gp.pos_end;
         pos += gridDim.x * blockDim.x) {
        if (bc_window(gp.B, pos) & lm) continue;
        if (gp.gate_first && !bc_gate_window(gate, gp.H, gp.Lo, gp.B, pos)) continue;
        const uint32_t ph = cj_perm(s_lut, bc_window(gp.H, pos) & lm, wide), pl = cj_perm(s_lut, bc_window(gp.Lo, pos) & lm, wide);
        const uint32_t slot = cd.dir_off + (((ph >> rem_nt) << key_nt) | (pl >> rem_nt));
        if (slot < gp.slot_lo || slot >= gp.slot_hi) continue;
        if (gp.prune && gp.lib_dir[slot] == gp.lib_dir[slot + 1]) continue;
        atomicAdd(&gdir[slot], 1u);
    }
}

// bin_start[g] = first record of bin g (g over all combinations, + end sentinel); bin_cursor = copy;
// bin_combo[g] = its combination
__global__ void k_cbin_init(const __grid_constant__ CBucketParams gp, const uint32_t* __restrict__ gdir,
                            uint32_t n_bins, uint32_t n_slots, uint32_t* __restrict__ bin_start,
                            uint32_t* __restrict__ bin_cursor, uint8_t* __restrict__ bin_combo) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_bins) return;
    if (g == n_bins) {
        bin_start[g] = gdir[n_slots];
        return;
    }
    uint32_t c = 0;
    while (c + 1 < gp.n_combos && gp.combo[c + 1].bin_off <= g) c++;
    const ComboDesc& cd = gp.combo[c];
    const uint32_t low = 2u * cd.key_nt - cd.top_bits;
    const uint32_t v = gdir[cd.dir_off + ((g - cd.bin_off) << low)];
    bin_start[g] = v;
    bin_cursor[g] = v;
    bin_combo[g] = (uint8_t)c;
}

// ------------------------------------------------------------ counting without a RED per record
// k_ccount<win> costs one global RED per (window, combination): 1.5e9 of them at cfg 4 = 7.3 ms at
// the L2 atomic rate (ncu: lts 87 %), 13 % of the step - only to learn where every bin and slot
// starts.  Round 2 counts in shared memory instead, in two places:
//   k_cbincount   before pass A: per combination a shared-memory histogram over its <= 1024 bins,
//                 flushed once per (CTA, combination).  When the bin is a function of the H plane
//                 alone (top_bits <= key_nt, nothing pruned or sharded away) only that plane is
//                 permuted.  A thread walks 32 consecutive positions with one pair of words per plane.
//   k_cslotcount  after pass A: the records of a bin are contiguous, so the slot histogram of a piece of
//                 a bin (<= 65536 records) is a shared-memory histogram over the bin's sub-slots, flushed
//                 with one RED per non-empty (piece, sub-slot): ~2.5e7 REDs instead of 1.5e9.
#define CB_THREADS 512
template <bool LIB, bool PART>  // PART: the directory is sharded (slot range != everything): bins are range-checked
__global__ void __launch_bounds__(CB_THREADS, 2) k_cbincount(const __grid_constant__ CBucketParams gp,
                                                             const uint32_t* __restrict__ lut,
                                                             uint32_t* __restrict__ bin_count) {
    __shared__ uint32_t s_lut[CJ_LUT_WORDS];
    __shared__ uint32_t s_hist[CJ_MAX_BINS];
    const uint32_t lm = bc_lmask(gp.L);
    const uint32_t tid = threadIdx.x;
    PamGate gate;
    bc_gate_init(gate, gp.P, gp.L, gp.pam_dir, gp.pam_sets);
    // static split of the plane words [w_lo, w_hi) over the CTAs
    const uint32_t w_lo = gp.pos_begin >> 5, w_hi = (gp.pos_end + 31u) >> 5;
    const uint32_t per = (w_hi - w_lo + gridDim.x - 1) / gridDim.x;
    const uint32_t my_lo = w_lo + blockIdx.x * per, my_hi = min(w_hi, my_lo + per);
    for (uint32_t c = 0; c < gp.n_combos; c++) {
        const ComboDesc& cd = gp.combo[c];
        if (!bc_combo_in_range(cd, gp.slot_lo, gp.slot_hi)) continue;  // block-uniform
        const uint32_t n_bins = 1u << cd.top_bits, key_nt = cd.key_nt, rem_nt = cd.rem_nt;
        const uint32_t low = 2u * key_nt - cd.top_bits;
        const bool whole = cd.dir_off >= gp.slot_lo && cd.dir_off + (1u << (2u * key_nt)) <= gp.slot_hi;
        const bool h_only = cd.top_bits <= key_nt && (whole || gp.bin_aligned) && (LIB || !gp.prune);
        // bins of this combination the shard owns (exact when the range is bin aligned)
        const uint32_t bin_lo = (max(gp.slot_lo, cd.dir_off) - cd.dir_off) >> low;
        const uint32_t bin_hi = (min(gp.slot_hi, cd.dir_off + (1u << (2u * key_nt))) - cd.dir_off) >> low;
        __syncthreads();
        for (uint32_t j = tid; j < n_bins; j += CB_THREADS) s_hist[j] = 0;
        for (uint32_t j = tid; j < CJ_LUT_WORDS; j += CB_THREADS) s_lut[j] = lut[c * CJ_LUT_WORDS + j];
        __syncthreads();
        if (LIB) {  // library entries (index build): planes qh / ql, an entry whose key touches a non-ACGT character is skipped
            for (uint32_t e = blockIdx.x * CB_THREADS + tid; e < gp.n_entries; e += gridDim.x * CB_THREADS) {
                if (gp.lib_has_n) {
                    uint32_t nm = gp.sn[e >> 1];
                    if (e & 1u) nm = bc_rev_bits(nm, gp.L);
                    if (nm & cd.key_mask) continue;
                }
                const uint32_t ph = cj_perm(s_lut, gp.qh[e], false);
                uint32_t bin;
                if (h_only) {
                    bin = (ph >> rem_nt) >> (key_nt - cd.top_bits);
                    if (PART && bin - bin_lo >= bin_hi - bin_lo) continue;
                } else {
                    const uint32_t pl = cj_perm(s_lut, gp.ql[e], false);
                    const uint32_t key = ((ph >> rem_nt) << key_nt) | (pl >> rem_nt);
                    const uint32_t slot = cd.dir_off + key;
                    if (slot < gp.slot_lo || slot >= gp.slot_hi) continue;
                    bin = key >> low;
                }
                atomicAdd(&s_hist[bin], 1u);
            }
        }
        for (uint32_t w = my_lo + tid; !LIB && w < my_hi; w += CB_THREADS) {
            const uint32_t wn = min(w + 1u, gp.n_words - 1u);  // the planes are padded: the clamp only guards the very last word
            const uint32_t b0 = gp.B[w], b1 = gp.B[wn], h0 = gp.H[w], h1 = gp.H[wn];
            const uint32_t l0 = h_only ? 0u : gp.Lo[w], l1 = h_only ? 0u : gp.Lo[wn];
#pragma unroll 4
            for (uint32_t o = 0; o < 32; o++) {
                const uint32_t pos = (w << 5) + o;
                if (pos < gp.pos_begin || pos >= gp.pos_end) continue;
                if (__funnelshift_r(b0, b1, o) & lm) continue;
                if (gp.gate_first && !bc_gate_window(gate, gp.H, gp.Lo, gp.B, pos)) continue;
                const uint32_t ph = cj_perm(s_lut, __funnelshift_r(h0, h1, o) & lm, false);
                uint32_t bin;
                if (h_only) {
                    bin = (ph >> rem_nt) >> (key_nt - cd.top_bits);
                    if (PART && bin - bin_lo >= bin_hi - bin_lo) continue;
                } else {
                    const uint32_t pl = cj_perm(s_lut, __funnelshift_r(l0, l1, o) & lm, false);
                    const uint32_t key = ((ph >> rem_nt) << key_nt) | (pl >> rem_nt);
                    const uint32_t slot = cd.dir_off + key;
                    if (slot < gp.slot_lo || slot >= gp.slot_hi) continue;
                    if (gp.prune && gp.lib_dir[slot] == gp.lib_dir[slot + 1]) continue;
                    bin = key >> low;
                }
                atomicAdd(&s_hist[bin], 1u);
            }
        }
        __syncthreads();
        for (uint32_t j = tid; j < n_bins; j += CB_THREADS) {
            const uint32_t v = s_hist[j];
            if (v) atomicAdd(&bin_count[cd.bin_off + j], v);
        }
    }
}

// bin_start = the scanned bin counts (+ end sentinel), bin_combo[g] = combination of bin g
__global__ void k_cbin_init2(const __grid_constant__ CBucketParams gp, uint32_t n_bins, const uint32_t* __restrict__ bin_cursor,
                             uint32_t* __restrict__ bin_start, uint8_t* __restrict__ bin_combo) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_bins) return;
    bin_start[g] = bin_cursor[g];
    if (g == n_bins) return;
    uint32_t c = 0;
    while (c + 1 < gp.n_combos && gp.combo[c + 1].bin_off <= g) c++;
    bin_combo[g] = (uint8_t)c;
}

#define CS_THREADS 512
#define CS_SEG 65536u
__global__ void __launch_bounds__(CS_THREADS, 2) k_cslotcount(const __grid_constant__ CBucketParams gp,
                                                              const uint2* __restrict__ tmp,
                                                              const uint32_t* __restrict__ bin_start,
                                                              const uint8_t* __restrict__ bin_combo, uint32_t n_bins,
                                                              uint32_t* __restrict__ gdir, uint32_t* __restrict__ work) {
    extern __shared__ __align__(16) uint32_t cj_smem[];
    uint32_t* s_hist = cj_smem;  // [max_sub]
    __shared__ uint32_t s_unit;
    const uint32_t tid = threadIdx.x;
    const uint32_t n_rec = bin_start[n_bins];
    const uint32_t n_units = (uint32_t)(((uint64_t)n_rec + CS_SEG - 1) / CS_SEG);
    for (;;) {
        __syncthreads();
        if (tid == 0) s_unit = atomicAdd(work, 1u);
        __syncthreads();
        const uint32_t u = s_unit;
        if (u >= n_units) break;
        const uint32_t r0 = u * CS_SEG, r1 = r0 + min(CS_SEG, n_rec - r0);
        uint32_t g;
        {
            uint32_t lo = 0, hi = n_bins;  // bin_start[lo] <= r0 < bin_start[hi]
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(bin_start + mid) <= r0) lo = mid; else hi = mid;
            }
            g = lo;
        }
        uint32_t seg = r0;
        while (seg < r1) {
            while (g + 1 < n_bins && __ldg(bin_start + g + 1) <= seg) g++;  // skip empty bins
            const uint32_t s1 = min(r1, __ldg(bin_start + g + 1));
            const ComboDesc& cd = gp.combo[__ldg(bin_combo + g)];
            const uint32_t low = 2u * cd.key_nt - cd.top_bits, rem2 = 2u * cd.rem_nt;
            const uint32_t n_sub = 1u << low;
            const uint32_t slot0 = cd.dir_off + ((g - cd.bin_off) << low);
            if (low == 0) {  // the bin is one slot
                if (tid == 0) atomicAdd(&gdir[slot0], s1 - seg);
                seg = s1;
                continue;
            }
            __syncthreads();
            for (uint32_t j = tid; j < n_sub; j += CS_THREADS) s_hist[j] = 0;
            __syncthreads();
            // 8 independent loads per thread before the first one is used: with one load in flight per thread the kernel
            // ran at 2.5 TB/s (latency bound)
            for (uint32_t i0 = seg; i0 < s1; i0 += 8u * CS_THREADS) {
                uint32_t y[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t i = i0 + tid + (uint32_t)j * CS_THREADS;
                    y[j] = i < s1 ? __ldcs(&tmp[i].y) : 0xffffffffu;
                }
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (i0 + tid + (uint32_t)j * CS_THREADS < s1) atomicAdd(&s_hist[y[j] >> rem2], 1u);
            }
            __syncthreads();
            for (uint32_t j = tid; j < n_sub; j += CS_THREADS) {
                const uint32_t v = s_hist[j];
                if (v) atomicAdd(&gdir[slot0 + j], v);
            }
            seg = s1;
        }
    }
}

// exclusive scan of CJ_THREADS values, one per thread; s_warp: CJ_THREADS / 32 words
__device__ __forceinline__ uint32_t cj_block_scan(uint32_t v, uint32_t* s_warp) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < CJ_THREADS / 32; w++) before += (uint32_t)w < warp ? s_warp[w] : 0u;
    __syncthreads();
    return before + incl - v;
}

// ------------------------------------------------------------------------------------- pass A
// Position-major: a chunk of CJ_CHUNK window positions is validated once (ambiguity, PAM gate), its
// plane words are staged in shared memory, and then every combination in turn bins the chunk's
// windows: shared-memory histogram over the combination's bins, ONE global atomic per (chunk, bin),
// records grouped by bin in shared memory and written out as runs.
// LIB = true sorts the library entries instead (index build): "position" = entry number, the planes
// are the query planes qh / ql, and an entry whose key touches a non-ACGT spacer character is skipped.
template <bool LIB>
__global__ void __launch_bounds__(CJ_THREADS, 2) k_cbin(const __grid_constant__ CBucketParams gp,
                                                        const uint32_t* __restrict__ lut,
                                                        uint32_t* __restrict__ bin_cursor, uint2* __restrict__ tmp) {
    extern __shared__ __align__(16) uint32_t cj_smem[];
    uint32_t* s_hist = cj_smem;                      // [CJ_MAX_BINS]
    uint32_t* s_lstart = s_hist + CJ_MAX_BINS;       // [CJ_MAX_BINS]
    uint32_t* s_delta = s_lstart + CJ_MAX_BINS;      // [CJ_MAX_BINS] global base - local start
    uint32_t* s_H = s_delta + CJ_MAX_BINS;           // [CJ_CHUNK / 32 + 2]
    uint32_t* s_L = s_H + (CJ_CHUNK / 32 + 2);
    uint32_t* s_warp = s_L + (CJ_CHUNK / 32 + 2);    // [CJ_THREADS / 32]
    uint32_t* s_lut = s_warp + CJ_THREADS / 32;      // [CJ_LUT_WORDS] permutation table of the current combination
    uint2* s_rec = reinterpret_cast<uint2*>(s_lut + CJ_LUT_WORDS);  // [CJ_CHUNK]
    uint16_t* s_bin = reinterpret_cast<uint16_t*>(s_rec + CJ_CHUNK);    // [CJ_CHUNK]
    const uint32_t lm = bc_lmask(gp.L);
    const bool wide = gp.L > 24;
    PamGate gate;
    bc_gate_init(gate, gp.P, gp.L, gp.pam_dir, gp.pam_sets);
    const uint32_t tid = threadIdx.x;
    const uint64_t src_begin = LIB ? 0ull : (uint64_t)gp.pos_begin, src_end = LIB ? (uint64_t)gp.n_entries : (uint64_t)gp.pos_end;
    for (uint64_t c0 = src_begin + (uint64_t)blockIdx.x * CJ_CHUNK; c0 < src_end; c0 += (uint64_t)gridDim.x * CJ_CHUNK) {
        // plane words of the chunk (+ the word after it: a window spans two words).  The chunk
        // need not start on a word boundary: word 0 of the stage is word c0 >> 5 of the plane.
        const uint32_t w0 = (uint32_t)(c0 >> 5), sh0 = (uint32_t)(c0 & 31u);
        __syncthreads();
        uint32_t ok = 0;
        if (!LIB) {
            for (uint32_t i = tid; i < CJ_CHUNK / 32 + 2; i += CJ_THREADS) {
                const uint32_t w = min(w0 + i, gp.n_words - 1u);  // the planes are padded by less than a chunk
                s_H[i] = gp.H[w];
                s_L[i] = gp.Lo[w];
            }
#pragma unroll
            for (int i = 0; i < CJ_ITEMS; i++) {
                const uint64_t pos64 = c0 + tid + (uint32_t)i * CJ_THREADS;
                if (pos64 >= gp.pos_end) continue;
                const uint32_t pos = (uint32_t)pos64;
                if (bc_window(gp.B, pos) & lm) continue;
                if (gp.gate_first && !bc_gate_window(gate, gp.H, gp.Lo, gp.B, pos)) continue;
                ok |= 1u << i;
            }
        } else {
#pragma unroll
            for (int i = 0; i < CJ_ITEMS; i++)
                if (c0 + tid + (uint32_t)i * CJ_THREADS < src_end) ok |= 1u << i;
        }
        __syncthreads();
        for (uint32_t c = 0; c < gp.n_combos; c++) {
            const ComboDesc& cd = gp.combo[c];
            if (!bc_combo_in_range(cd, gp.slot_lo, gp.slot_hi)) continue;  // block-uniform
            const uint32_t low = 2u * cd.key_nt - cd.top_bits, rem_nt = cd.rem_nt;
            const uint32_t n_bins = 1u << cd.top_bits, key_nt = cd.key_nt;
            const uint32_t rm = (1u << rem_nt) - 1u, low_mask = (1u << low) - 1u;
            // the common case - key bits split evenly between the passes, nothing sharded or pruned away - needs
            // neither the assembled key nor the range / prune tests (pass A is issue- and shared-memory bound)
            const bool whole = cd.dir_off >= gp.slot_lo && cd.dir_off + (1u << (2u * key_nt)) <= gp.slot_hi;
            const bool even = low == key_nt && (LIB || !gp.prune) && (whole || gp.bin_aligned);
            const uint32_t bin_lo = (max(gp.slot_lo, cd.dir_off) - cd.dir_off) >> low;
            const uint32_t bin_hi = (min(gp.slot_hi, cd.dir_off + (1u << (2u * key_nt))) - cd.dir_off) >> low;
            for (uint32_t j = tid; j < n_bins; j += CJ_THREADS) s_hist[j] = 0;
            for (uint32_t j = tid; j < CJ_LUT_WORDS; j += CJ_THREADS) s_lut[j] = lut[c * CJ_LUT_WORDS + j];
            __syncthreads();
            uint32_t x[CJ_ITEMS], rb[CJ_ITEMS];
#pragma unroll
            for (int i = 0; i < CJ_ITEMS; i++) {
                rb[i] = 0xffffffffu;
                if (!((ok >> i) & 1u)) continue;
                uint32_t wh, wl;
                if (!LIB) {
                    const uint32_t t = sh0 + tid + (uint32_t)i * CJ_THREADS;  // bit offset inside the staged words
                    wh = __funnelshift_r(s_H[t >> 5], s_H[(t >> 5) + 1], t & 31u) & lm;
                    wl = __funnelshift_r(s_L[t >> 5], s_L[(t >> 5) + 1], t & 31u) & lm;
                } else {
                    const uint32_t e = (uint32_t)c0 + tid + (uint32_t)i * CJ_THREADS;
                    if (gp.lib_has_n) {  // a seed containing a non-ACGT spacer character is never exact
                        uint32_t nm = gp.sn[e >> 1];
                        if (e & 1u) nm = bc_rev_bits(nm, gp.L);
                        if (nm & cd.key_mask) continue;
                    }
                    wh = gp.qh[e];
                    wl = gp.ql[e];
                }
                const uint32_t ph = cj_perm(s_lut, wh, wide), pl = cj_perm(s_lut, wl, wide);
                uint32_t bin;
                if (even) {  // block-uniform: the bin is the H half of the key, and x = pl << rem_nt | Rh as it stands
                    bin = ph >> rem_nt;
                    if (bin - bin_lo >= bin_hi - bin_lo) continue;  // (all bins when the shard owns the whole combination)
                    x[i] = (pl << rem_nt) | (ph & rm);
                } else {
                    const uint32_t key = ((ph >> rem_nt) << key_nt) | (pl >> rem_nt);
                    if (cd.dir_off + key < gp.slot_lo || cd.dir_off + key >= gp.slot_hi) continue;
                    if (!LIB && gp.prune && gp.lib_dir[cd.dir_off + key] == gp.lib_dir[cd.dir_off + key + 1]) continue;
                    bin = key >> low;
                    x[i] = ((key & low_mask) << (2u * rem_nt)) | ((pl & rm) << rem_nt) | (ph & rm);
                }
                rb[i] = atomicAdd(&s_hist[bin], 1u) | (bin << 16);
            }
            __syncthreads();
            {   // two consecutive counters per thread (CJ_MAX_BINS == 2 * CJ_THREADS)
                const uint32_t j0 = 2u * tid, j1 = j0 + 1;
                const uint32_t c0n = j0 < n_bins ? s_hist[j0] : 0u, c1n = j1 < n_bins ? s_hist[j1] : 0u;
                const uint32_t before = cj_block_scan(c0n + c1n, s_warp);
                if (j0 < n_bins) {
                    s_lstart[j0] = before;
                    s_delta[j0] = c0n ? atomicAdd(&bin_cursor[cd.bin_off + j0], c0n) - before : 0u;
                }
                if (j1 < n_bins) {
                    s_lstart[j1] = before + c0n;
                    s_delta[j1] = c1n ? atomicAdd(&bin_cursor[cd.bin_off + j1], c1n) - (before + c0n) : 0u;
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < CJ_ITEMS; i++) {
                if (rb[i] == 0xffffffffu) continue;
                const uint32_t bin = rb[i] >> 16, at = s_lstart[bin] + (rb[i] & 0xffffu);
                s_rec[at] = make_uint2((uint32_t)(c0 + tid + (uint32_t)i * CJ_THREADS), x[i]);
                s_bin[at] = (uint16_t)bin;
            }
            __syncthreads();
            const uint32_t total = s_lstart[n_bins - 1] + s_hist[n_bins - 1];
            for (uint32_t i = tid; i < total; i += CJ_THREADS) tmp[i + s_delta[s_bin[i]]] = s_rec[i];
            __syncthreads();
        }
    }
}
static_assert(CJ_MAX_BINS == 2 * CJ_THREADS, "k_cbin scans two bins per thread");
static_assert(CJ_CHUNK <= 65536, "k_cbin packs the local rank into 16 bits");

#ifndef CJ_PLACE_FORM
#define CJ_PLACE_FORM 1   // pass B: 0 staged (LDG), 1 bulk-async with 4096-record chunks and 2 CTAs per SM, 2 bulk-async with 8192-record chunks (7.80 / 8.05 ms at cfg 4)
#endif
#define CJ_PLACE_CHUNK (CJ_PLACE_FORM == 1 ? CJ_CHUNK / 2 : CJ_CHUNK)  // records per pass-B chunk
// chunk_bin[ch] = bin that holds record ch * CJ_PLACE_CHUNK of the pass-A output
__global__ void k_cchunk_bins(const uint32_t* __restrict__ bin_start, uint32_t n_bins, const uint32_t* __restrict__ n_rec_ptr,
                              uint32_t* __restrict__ chunk_bin) {
    const uint32_t ch = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_rec = *n_rec_ptr;
    if ((uint64_t)ch * CJ_PLACE_CHUNK >= n_rec) return;
    const uint32_t r = ch * CJ_PLACE_CHUNK;
    uint32_t lo = 0, hi = n_bins;  // bin_start[lo] <= r < bin_start[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (bin_start[mid] <= r) lo = mid; else hi = mid;
    }
    chunk_bin[ch] = lo;
}

// ------------------------------------------------------------------------------------- pass B
// Chunks of the pass-A output, handed out through an atomic counter; a chunk that straddles bin
// regions is processed piece by piece.  Per piece: shared-memory histogram over the bin's
// sub-slots (low key bits), one global atomic per (piece, sub-slot) on the slot cursor, records
// grouped by sub-slot in shared memory and written to their final slots as runs.
// LIB = true writes the library index instead of window records: ent_hl[dst] = {Rh, Rl}, ent_id[dst] = entry.
template <bool LIB>
__global__ void __launch_bounds__(CJ_THREADS, 2) k_cplace(const __grid_constant__ CBucketParams gp,
                                                          const uint2* __restrict__ tmp,
                                                          const uint32_t* __restrict__ bin_start,
                                                          const uint8_t* __restrict__ bin_combo,
                                                          const uint32_t* __restrict__ chunk_bin, uint32_t n_bins,
                                                          uint32_t* __restrict__ gcursor, uint2* __restrict__ gwin,
                                                          uint32_t* __restrict__ out_id,
                                                          const uint32_t* __restrict__ n_rec_ptr,
                                                          uint32_t* __restrict__ work, uint32_t max_sub) {
    extern __shared__ __align__(16) uint32_t cj_smem[];
    uint32_t* s_hist = cj_smem;               // [max_sub]
    uint32_t* s_lstart = s_hist + max_sub;    // [max_sub]
    uint32_t* s_delta = s_lstart + max_sub;   // [max_sub]
    uint32_t* s_warp = s_delta + max_sub;     // [CJ_THREADS / 32]
    uint2* s_rec = reinterpret_cast<uint2*>(s_warp + CJ_THREADS / 32);  // [CJ_CHUNK]
    __shared__ uint32_t s_chunk;
    const uint32_t tid = threadIdx.x;
    const uint32_t n_rec = *n_rec_ptr;
    const uint32_t n_chunks = (uint32_t)(((uint64_t)n_rec + CJ_CHUNK - 1) / CJ_CHUNK);
    for (;;) {
        __syncthreads();
        if (tid == 0) s_chunk = atomicAdd(work, 1u);
        __syncthreads();
        const uint32_t ch = s_chunk;
        if (ch >= n_chunks) break;
        const uint32_t r0 = ch * CJ_CHUNK, r1 = r0 + min((uint32_t)CJ_CHUNK, n_rec - r0);
        uint32_t g = __ldg(chunk_bin + ch), seg = r0;
        while (seg < r1) {
            while (g + 1 < n_bins && __ldg(bin_start + g + 1) <= seg) g++;  // skip empty bins
            const uint32_t s1 = min(r1, __ldg(bin_start + g + 1));
            const ComboDesc& cd = gp.combo[__ldg(bin_combo + g)];
            const uint32_t low = 2u * cd.key_nt - cd.top_bits, rem2 = 2u * cd.rem_nt;
            const uint32_t n_sub = 1u << low;
            const uint32_t slot0 = cd.dir_off + ((g - cd.bin_off) << low);
            uint2 rec[CJ_ITEMS];
            uint32_t rank[CJ_ITEMS];
            const uint32_t rem_nt = cd.rem_nt, rm = (1u << rem_nt) - 1u;
            if (low == 0) {  // the bin is one slot: a plain copy
#pragma unroll
                for (int i = 0; i < CJ_ITEMS; i++) {
                    const uint32_t idx = seg + tid + (uint32_t)i * CJ_THREADS;
                    if (idx < s1) {
                        const uint2 r = __ldcs(tmp + idx);
                        if (LIB) {
                            gwin[idx] = make_uint2(r.y & rm, (r.y >> rem_nt) & rm);
                            out_id[idx] = r.x;
                        } else {
                            gwin[idx] = r;
                        }
                    }
                }
                seg = s1;
                continue;
            }
            for (uint32_t j = tid; j < n_sub; j += CJ_THREADS) s_hist[j] = 0;
            __syncthreads();
#pragma unroll
            for (int i = 0; i < CJ_ITEMS; i++) {
                const uint32_t idx = seg + tid + (uint32_t)i * CJ_THREADS;
                if (idx < s1) {
                    rec[i] = __ldcs(tmp + idx);
                    rank[i] = atomicAdd(&s_hist[rec[i].y >> rem2], 1u);
                }
            }
            __syncthreads();
            {   // n_sub / CJ_THREADS (>= 1 when n_sub >= CJ_THREADS) consecutive counters per thread
                const uint32_t per = (n_sub + CJ_THREADS - 1) / CJ_THREADS;
                const uint32_t j0 = tid * per;
                uint32_t sum = 0;
                for (uint32_t j = 0; j < per; j++) sum += j0 + j < n_sub ? s_hist[j0 + j] : 0u;
                uint32_t run = cj_block_scan(sum, s_warp);
                for (uint32_t j = 0; j < per && j0 + j < n_sub; j++) {
                    const uint32_t cnt = s_hist[j0 + j];
                    s_lstart[j0 + j] = run;
                    s_delta[j0 + j] = cnt ? atomicAdd(&gcursor[slot0 + j0 + j], cnt) - run : 0u;
                    run += cnt;
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < CJ_ITEMS; i++) {
                const uint32_t idx = seg + tid + (uint32_t)i * CJ_THREADS;
                if (idx < s1) s_rec[s_lstart[rec[i].y >> rem2] + rank[i]] = rec[i];
            }
            __syncthreads();
            const uint32_t n = s1 - seg;
            for (uint32_t i = tid; i < n; i += CJ_THREADS) {
                const uint2 r = s_rec[i];
                const uint32_t dst = i + s_delta[r.y >> rem2];
                if (LIB) {
                    gwin[dst] = make_uint2(r.y & rm, (r.y >> rem_nt) & rm);
                    out_id[dst] = r.x;
                } else {
                    gwin[dst] = r;
                }
            }
            __syncthreads();
            seg = s1;
        }
    }
}

// ------------------------------------------------------------------- pass B, bulk-async form
// Same pass with the chunk loads taken off the critical path: a chunk of the pass-A output is one
// contiguous run of bytes, so ONE elected thread fetches it with a 1-D bulk asynchronous copy
// (cp.async.bulk.shared::cluster.global, completion on an mbarrier) into one half of a double
// buffer while the CTA sorts the other half.  The chunk is sorted IN PLACE in its buffer (every
// thread holds its records in registers between the two barriers), so the second buffer costs no
// extra shared memory over the staged form above.  ncu on the staged form: 2.35 TB/s, long
// scoreboard 15 warps per issue - load, sort and store of a chunk were serialised per CTA.
__device__ __forceinline__ void pl_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void pl_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy accesses to the buffer are done
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                 "l"(gmem_src), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void pl_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b),
        "r"(parity)
        : "memory");
}

template <bool LIB, int ITEMS>
__global__ void __launch_bounds__(CJ_THREADS, ITEMS > 8 ? 1 : 2) k_cplace_bulk(const __grid_constant__ CBucketParams gp,
                                                                             const uint2* __restrict__ tmp,
                                                                             const uint32_t* __restrict__ bin_start,
                                                                             const uint8_t* __restrict__ bin_combo,
                                                                             const uint32_t* __restrict__ chunk_bin,
                                                                             uint32_t n_bins, uint32_t* __restrict__ gcursor,
                                                                             uint2* __restrict__ gwin, uint32_t* __restrict__ out_id,
                                                                             const uint32_t* __restrict__ n_rec_ptr,
                                                                             uint32_t* __restrict__ work, uint32_t max_sub) {
    constexpr uint32_t CH = CJ_THREADS * ITEMS;
    extern __shared__ __align__(128) uint32_t cj_smem[];
    uint2* s_buf = reinterpret_cast<uint2*>(cj_smem);            // [2][CH] chunk double buffer (sorted in place)
    uint32_t* s_hist = cj_smem + 4 * CH;                         // [max_sub]
    uint32_t* s_lstart = s_hist + max_sub;                       // [max_sub]
    uint32_t* s_delta = s_lstart + max_sub;                      // [max_sub]
    uint32_t* s_warp = s_delta + max_sub;                        // [CJ_THREADS / 32]
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_chunk[2], s_cbin[2];
    const uint32_t tid = threadIdx.x;
    const uint32_t n_rec = *n_rec_ptr;
    const uint32_t n_chunks = (uint32_t)(((uint64_t)n_rec + CH - 1) / CH);
    if (tid == 0) {
        pl_mbar_init(&s_bar[0], 1);
        pl_mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t ch = atomicAdd(work, 1u);
        s_chunk[0] = ch;
        if (CH == CJ_PLACE_CHUNK && ch < n_chunks) s_cbin[0] = __ldg(chunk_bin + ch);
        if (ch < n_chunks) {
            const uint32_t r0 = ch * CH, cnt = min(CH, n_rec - r0);
            pl_bulk_load(s_buf, tmp + r0, (cnt * 8u + 15u) & ~15u, &s_bar[0]);
        }
    }
    uint32_t cur = 0, parity0 = 0, parity1 = 0;
    for (;;) {
        __syncthreads();  // s_chunk[cur] is visible; everybody is done with buffer cur ^ 1
        const uint32_t ch = s_chunk[cur];
        if (ch >= n_chunks) break;
        if (tid == 0) {   // fetch the next chunk into the other buffer while this one is sorted
            const uint32_t nx = atomicAdd(work, 1u);
            s_chunk[cur ^ 1u] = nx;
            if (nx < n_chunks) {
                const uint32_t r0n = nx * CH, cntn = min(CH, n_rec - r0n);
                pl_bulk_load(s_buf + (cur ^ 1u) * CH, tmp + r0n, (cntn * 8u + 15u) & ~15u, &s_bar[cur ^ 1u]);
                if (CH == CJ_PLACE_CHUNK) s_cbin[cur ^ 1u] = __ldg(chunk_bin + nx);
            }
        }
        pl_mbar_wait(&s_bar[cur], cur ? parity1 : parity0);
        if (cur) parity1 ^= 1u; else parity0 ^= 1u;
        uint2* buf = s_buf + cur * CH;
        const uint32_t r0 = ch * CH, r1 = r0 + min(CH, n_rec - r0);
        // bin of the chunk's first record: from the chunk table (k_cchunk_bins), fetched together with the chunk by the
        // elected thread.  (Every warp searching bin_start itself - 14 dependent loads per chunk - was the top stall
        // site of this kernel: ncu, 15.8 % of the samples.)
        uint32_t g;
        if (CH == CJ_PLACE_CHUNK) {
            g = s_cbin[cur];
        } else {
            uint32_t lo = 0, hi = n_bins;  // bin_start[lo] <= r0 < bin_start[hi]
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(bin_start + mid) <= r0) lo = mid; else hi = mid;
            }
            g = lo;
        }
        uint32_t seg = r0;
        while (seg < r1) {
            while (g + 1 < n_bins && __ldg(bin_start + g + 1) <= seg) g++;  // skip empty bins
            const uint32_t s1 = min(r1, __ldg(bin_start + g + 1));
            const ComboDesc& cd = gp.combo[__ldg(bin_combo + g)];
            const uint32_t low = 2u * cd.key_nt - cd.top_bits, rem2 = 2u * cd.rem_nt;
            const uint32_t n_sub = 1u << low;
            const uint32_t slot0 = cd.dir_off + ((g - cd.bin_off) << low);
            const uint32_t rem_nt = cd.rem_nt, rm = (1u << rem_nt) - 1u;
            uint2* sb = buf + (seg - r0);
            const uint32_t n = s1 - seg;
            if (low != 0) {
                uint2 rec[ITEMS];
                uint32_t rank[ITEMS];
                for (uint32_t j = tid; j < n_sub; j += CJ_THREADS) s_hist[j] = 0;
                __syncthreads();
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    const uint32_t idx = tid + (uint32_t)i * CJ_THREADS;
                    if (idx < n) {
                        rec[i] = sb[idx];
                        rank[i] = atomicAdd(&s_hist[rec[i].y >> rem2], 1u);
                    }
                }
                __syncthreads();  // every record of the piece is in registers: the buffer may be overwritten
                {
                    const uint32_t per = (n_sub + CJ_THREADS - 1) / CJ_THREADS;
                    const uint32_t j0 = tid * per;
                    uint32_t sum = 0;
                    for (uint32_t j = 0; j < per; j++) sum += j0 + j < n_sub ? s_hist[j0 + j] : 0u;
                    uint32_t run = cj_block_scan(sum, s_warp);
                    // (holding the atomics' results in registers until the records are grouped, so that the cursors
                    // answer meanwhile, measured slower: 9.80 against 9.58 ms, 112 registers)
                    for (uint32_t j = 0; j < per && j0 + j < n_sub; j++) {
                        const uint32_t cnt = s_hist[j0 + j];
                        s_lstart[j0 + j] = run;
                        s_delta[j0 + j] = cnt ? atomicAdd(&gcursor[slot0 + j0 + j], cnt) - run : 0u;
                        run += cnt;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    const uint32_t idx = tid + (uint32_t)i * CJ_THREADS;
                    if (idx < n) sb[s_lstart[rec[i].y >> rem2] + rank[i]] = rec[i];
                }
                __syncthreads();
            }
            for (uint32_t i = tid; i < n; i += CJ_THREADS) {
                const uint2 r = sb[i];
                const uint32_t dst = low != 0 ? i + s_delta[r.y >> rem2] : seg + i;
                if (LIB) {
                    gwin[dst] = make_uint2(r.y & rm, (r.y >> rem_nt) & rm);
                    out_id[dst] = r.x;
                } else {
                    gwin[dst] = r;
                }
            }
            __syncthreads();
            seg = s1;
        }
        cur ^= 1u;
    }
}

// ------------------------------------------------------------------------------------- verify
#define CV_THREADS 256
#define CV_WARPS (CV_THREADS / 32)
#define CV_ITEMS 4          // window records per lane per warp-tile (at most)
#define CV_WQ 128           // per-warp candidate queue (entries)
#ifndef CV_QROWS
#define CV_QROWS 8
#endif
// CV_QROWS: per-warp item queue: CV_QROWS rows of 32 lane-private slots ({window, group} items for k_cfinish)
#define CV_GQ (32 * CV_QROWS)
#ifndef CV_STAGE
#define CV_STAGE 32         // library entries per shared-memory stage (x2 buffers per warp); buckets hold ~19 entries at cfg 4 (64: 11.35 ms, 32: 11.09 ms)
#endif
#define CV_GROUP 8          // entries per group (one ballot per group)
#ifndef CV_CHUNK_TILES
#define CV_CHUNK_TILES 32   // 128-record units per work chunk
#endif
#ifndef CV_MINBLOCKS
#define CV_MINBLOCKS 4
#endif
#ifndef CV_ITEM_POS
#define CV_ITEM_POS 1       // items handed to k_cfinish carry the window's dev position (fetched at flush time) instead of the record index
#endif
#define CV_INVALID 0xf0000000u  // rem plane of a missing window: 4 mismatches above the rem bits, never <= k

__device__ __forceinline__ void cv_cp_async8(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}

__device__ __forceinline__ uint32_t cv_combo_of_slot(const SearchParams& p, uint32_t slot) {
    uint32_t c = 0;
    while (c + 1 < p.n_combos && p.combo[c + 1].dir_off <= slot) c++;
    return c;
}

// Resolve up to 32 queued candidates {dev position, rem mismatch mask, index entry, slot} with all
// lanes: mask back to query positions, ownership, PAM annotation, ONE global atomic for the batch,
// coalesced store of the surviving records.
// ypos: the queue entries carry the dev position itself (k_cfinish: the verify kernel fetched it when it flushed
// the item, while the record's sector was still in L2) instead of the record index.
// The second level is inlined into k_cfinish (its only caller now that the verify kernel has no slow path): out of line
// the SearchParams reference is a generic pointer, and the loops over the combinations (cv_combo_of_slot, bc_owns) and the
// PAM annotation read every field with a dependent load instead of from the constant bank (ncu: those loops ran 25 million
// times per launch; k_cfinish 6.35 -> 4.92 ms).
#ifndef CV_FIN_INLINE
#define CV_FIN_INLINE __forceinline__
#define CV_FIN_MAKE_HIT bc_make_hit_inl
#endif
static __device__ CV_FIN_INLINE void cv_resolve(const SearchParams& p, const uint2* __restrict__ gwin, const uint4* q, uint32_t n,
                                               bool ypos) {
#ifdef CV_DEBUG_NO_RESOLVE  // timing experiment only: candidates are found but not turned into records
    if (p.cap != 1) return;
#endif
    const uint32_t lane = threadIdx.x & 31u;
    uint4 rec;
    bool ok = false;
    if (lane < n) {
        const uint4 qe = q[lane];  // x record index, y rem mismatch mask, z index entry, w slot
        const uint32_t c = cv_combo_of_slot(p, qe.w);
        const uint32_t m = bc_combo_rem_expand(p.combo[c], qe.y);
        // ownership is decided from the mask alone, before the position and the entry id (two random
        // DRAM sectors) are fetched - nearly half of the candidates are not owned
        if (p.lib_has_n || bc_owns(p, c, m)) {
            const uint32_t pos = ypos ? qe.x : __ldg(&gwin[qe.x].x), e = __ldg(p.ent_id + qe.z);
            ok = CV_FIN_MAKE_HIT(p, c, pos, e, m, &rec);
        }
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, ok);
    if (ballot == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(p.count, (unsigned long long)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (ok) {
        const unsigned long long dst = base + __popc(ballot & ((1u << lane) - 1u));
        if (dst < p.cap) reinterpret_cast<uint4*>(p.hits)[dst] = rec;
    }
}

static __device__ __noinline__ void cv_overflow(const SearchParams& p, const uint2* __restrict__ gwin, uint32_t rec_idx,
                                                uint32_t mr, uint32_t e, uint32_t slot, bool ypos) {
    uint4 rec;
    const uint32_t c = cv_combo_of_slot(p, slot);
    if (bc_make_hit(p, c, ypos ? rec_idx : __ldg(&gwin[rec_idx].x), p.ent_id[e], bc_combo_rem_expand(p.combo[c], mr), &rec)) {
        const unsigned long long g = atomicAdd(p.count, 1ull);
        if (g < p.cap) reinterpret_cast<uint4*>(p.hits)[g] = rec;
    }
}

__device__ __forceinline__ void cv_drain(const SearchParams& p, const uint2* __restrict__ gwin, uint4* q, uint32_t* qn,
                                         uint32_t lane, bool ypos = false) {
    __syncwarp();
    uint32_t nq = min(*qn, (uint32_t)CV_WQ);
    if (nq >= 32) {
        do {
            cv_resolve(p, gwin, q + (nq - 32), 32, ypos);
            nq -= 32;
        } while (nq >= 32);
        __syncwarp();
        if (lane == 0) *qn = nq;
    }
    __syncwarp();
}

// Second level.  A queue item {window planes (wh | wl << 16), record index, first entry of the group,
// slot} stands for ONE window x CV_GROUP entries of which at least one pair passed the filter.
// 32 items are re-examined at once, one per lane, so the per-pair compare + branch runs with full
// lanes; the window comes from the queue (no record re-read), only passing pairs fetch their
// position.  (The first version queued whole 4-window groups: at 9-10 nt keys 70-90 % of the groups
// contain a passing pair somewhere in the warp, and re-reading their records cost 13 GB of random
// sectors - ncu: 31.7 GB read for 11.5 GB algorithmic.)
static __device__ CV_FIN_INLINE void cv_resolve_groups(const SearchParams& p, const uint2* __restrict__ gwin,
                                                      const uint4* gq, uint32_t n, uint4* q, uint32_t* qn, bool ypos) {
    const uint32_t lane = threadIdx.x & 31u;
    const int k = (int)p.k;
    __syncwarp();
    const uint4 item = lane < n ? gq[lane] : make_uint4(0u, 0u, 0u, 0xffffffffu);  // x planes, y record index, z first entry, w slot
    if (item.w != 0xffffffffu) {
        const uint32_t wh = item.x & 0xffffu, wl = item.x >> 16;
        const uint32_t le = __ldg(p.dir + item.w + 1);  // end of the bucket: the last group may be ragged
        // all CV_GROUP entry loads are issued before the first one is used (one at a time, every lane
        // waited for CV_GROUP dependent DRAM round trips: 22 us per batch of 32 items in k_cfinish)
        uint2 qe[CV_GROUP];
#pragma unroll
        for (int j = 0; j < CV_GROUP; j++) qe[j] = __ldg(p.ent_hl + min(item.z + j, le - 1u));
#pragma unroll
        for (int j = 0; j < CV_GROUP; j++) {
            const uint32_t m_ = (wh ^ qe[j].x) | (wl ^ qe[j].y);
            if (item.z + j < le && __popc(m_) <= k) {
                const uint32_t qs = atomicAdd(qn, 1u);
                if (qs < CV_WQ) q[qs] = make_uint4(item.y, m_, item.z + j, item.w);
                else cv_overflow(p, gwin, item.y, m_, item.z + j, item.w, ypos);
            }
        }
    }
    cv_drain(p, gwin, q, qn, lane, ypos);
}

// Hand the queued items of the warp over to k_cfinish through the global item queue.  Re-examining them inside the verify kernel cost it 45 % of its time
// (timing experiment, cfg 4 at 9-nt keys: first level alone 22.6 ms, + second level 32.9 ms,
// + hit resolution 39.1 ms): the dependent loads of the slow path stall warps that should be
// feeding the POPC pipe.  The verify kernel therefore contains no slow path at all - not even as a
// fallback: a call inside its hot loop made ptxas keep loop state in local memory (324 bytes of
// spills; with 214 KB of the SM's 256 KB carved out as shared memory those reloads miss L1: ncu put
// 25 % of the stall samples on the instructions behind them).  If the queue is full the batch is
// dropped, the demand keeps counting, k_cfinish records it in count[5] (and processes nothing) and
// bc_search repeats the search with a queue of the demanded size (like a hit-buffer overflow).
//
// The per-warp queue is lane-private (round 2): lane l appends to column l (gq[row * 32 + l], its fill in a register),
// so queueing an item is four predicated instructions and needs no vote.  The ballot-compacted queue before it cost a
// VOTE + compare + branch per window and group plus ~12 instructions whenever any lane had an item - 60 % of the time at
// 10-nt keys: a quarter of all instructions of a kernel that ncu shows issue bound (issue active 78 %, no memory stalls
// left).  The queue is flushed when some lane could overflow with the next group: a warp prefix sum of the fills, ONE
// atomic for the warp, every lane copies its own items.
__device__ __forceinline__ void cv_flush_items(const SearchParams& p, const uint2* __restrict__ gwin, const uint4* gq, uint32_t cnt) {
    const uint32_t lane = threadIdx.x & 31u;
    __syncwarp();
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;  // warp-uniform
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(p.count + 4, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + total <= p.item_cap) {  // else: dropped, the demand keeps counting (see above)
        const unsigned long long dst = base + incl - cnt;
        const uint32_t rows = __reduce_max_sync(0xffffffffu, cnt);
        for (uint32_t r = 0; r < rows; r++) {
            if (r < cnt) {
                uint4 item = gq[r * 32 + lane];
#if CV_ITEM_POS
                // the item leaves with the window's dev position instead of its record index: the record's sector was read
                // a moment ago (L2 hit), in k_cfinish the same fetch is a random DRAM sector per candidate
                item.y = __ldg(&gwin[item.y].x);
#endif
                p.items[dst + r] = item;
            }
        }
    }
    __syncwarp();
}

// One warp-tile: `ITEMS` (1..CV_ITEMS, warp-uniform) resident windows per lane against the bucket
// [ls, le) of the library index, staged through shared memory (cp.async, CV_STAGE entries per
// stage, double buffered).  One group = CV_GROUP entries (16-byte broadcast LDS, two entries
// each) x ITEMS windows: independent LOP3/LOP3/POPC chains folded with min PER WINDOW, then one
// ballot per window; a lane whose minimum passes only QUEUES {window, group}.  With CV_ALU_PAIRS,
// 1 pair in 8 is tested on the ALU pipe instead of POPC (mismatch mask with its K lowest set bits
// cleared == 0): POPC alone saturates the XU pipe.
//
// Tile pipeline (round 2).  At 10-nt keys a tile is ~95 windows x ~19 entries: ~70 POPC per lane
// behind a chain of three dependent memory latencies (descriptor -> window records / bucket ->
// cp.async wait), 15.7 million times.  ncu on the unpipelined loop: long scoreboard 6.1 warps per
// issue, XU pipe 58 %.  A register software pipeline (descriptor t+2, records t+1 in flight) spilled
// more than it hid.  Here everything a tile needs travels through shared memory with cp.async and
// costs no registers: while tile t is verified, the window words and the first bucket stage of tile
// t+1 and the descriptor of tile t+2 are in flight; the switch to the next tile is three LDS.
#ifndef CV_ALU_PAIRS
#define CV_ALU_PAIRS 0   // (round 1's dense kernel saturated the XU pipe and won 3 % from this; the pipelined kernel does not: 11.9 vs 12.1 ms)
#endif
#define CV_WTILE (32 * CV_ITEMS)
#ifndef CV_TILE_INLINE
#define CV_TILE_INLINE __forceinline__
#endif
// bytes of dynamic shared memory per warp of k_cverify
#define CV_WARP_SMEM ((CV_GQ + 2) * 16 + 2 * CV_STAGE * 8 + 2 * CV_WTILE * 4 + 16)

__device__ __forceinline__ void cv_cp_async4(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cv_cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

// per-warp shared-memory staging of the tile pipeline
struct CvPipe {
    uint2* ent;        // [2][CV_STAGE] bucket stages (ping-pong by a running stage counter)
    uint32_t* win;     // [2][CV_WTILE] x words of the window records, by tile parity
    uint4* desc;       // [2] tile descriptors, by tile parity
    uint32_t* dslot;   // [2] slot of the tile, by tile parity
    const uint2* gwin;
    const uint2* ent_hl;
    const uint4* tile_desc;
    const uint32_t* tile_slot;
};

// lane 0 requests descriptor t (no commit)
__device__ __forceinline__ void cv_issue_desc(const CvPipe& pp, uint32_t t, uint32_t lane) {
    if (lane == 0) {
        cv_cp_async16(pp.desc + (t & 1u), pp.tile_desc + t);
        cv_cp_async4(pp.dslot + (t & 1u), pp.tile_slot + t);
    }
}
// window words of tile t (descriptor d) and stage `c` of its bucket into stage buffer `sb` (no commit)
__device__ __forceinline__ void cv_issue_windows(const CvPipe& pp, uint32_t t, const uint4& d, uint32_t lane) {
    const uint32_t n_win = d.y & 0xffu;
    uint32_t* dst = pp.win + (t & 1u) * CV_WTILE;
#pragma unroll
    for (int it = 0; it < CV_ITEMS; it++) {
        const uint32_t off = it * 32 + lane;
        if (off < n_win) cv_cp_async4(dst + off, &pp.gwin[d.x + off].y);
    }
}
__device__ __forceinline__ void cv_issue_stage(const CvPipe& pp, uint32_t ls, uint32_t le, uint32_t c, uint32_t sb, uint32_t lane) {
    uint2* dst = pp.ent + (sb & 1u) * CV_STAGE;
    const uint32_t n_ent = le - ls;
#pragma unroll
    for (int h = 0; h < CV_STAGE / 32; h++) {
        const uint32_t i = c * CV_STAGE + h * 32 + lane;
        if (i < n_ent) cv_cp_async8(dst + h * 32 + lane, pp.ent_hl + ls + i);
    }
}

// One group: UNITS x 2 entries (one 16-byte broadcast LDS each) against the ITEMS windows of the lane; pass[it] = some
// entry of the group is within K mismatches of window it.  In a full group one pair in eight goes to the ALU pipe
// (CV_ALU_PAIRS).
template <int K, int ITEMS, int UNITS>
__device__ __forceinline__ void cv_group(const uint4* sg, const uint32_t (&wh)[ITEMS], const uint32_t (&wl)[ITEMS], bool (&pass)[ITEMS]) {
    int best_[ITEMS];
    uint32_t rest_[ITEMS];
#pragma unroll
    for (int it = 0; it < ITEMS; it++) { best_[it] = 33; rest_[it] = 1u; }
#pragma unroll
    for (int j = 0; j < UNITS; j++) {
        const uint4 e2 = sg[j];
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            best_[it] = min(best_[it], __popc((wh[it] ^ e2.x) | (wl[it] ^ e2.y)));
            if (CV_ALU_PAIRS && UNITS == CV_GROUP / 2 && j == UNITS - 1) {
                uint32_t m = (wh[it] ^ e2.z) | (wl[it] ^ e2.w);
#pragma unroll
                for (int cc = 0; cc < K; cc++) m &= m - 1u;
                rest_[it] = m;
            } else {
                best_[it] = min(best_[it], __popc((wh[it] ^ e2.z) | (wl[it] ^ e2.w)));
            }
        }
    }
#pragma unroll
    for (int it = 0; it < ITEMS; it++) pass[it] = best_[it] <= K || rest_[it] == 0u;
}

// Verifies tile t (descriptor d, slot) whose window words and first bucket stage were requested
// earlier; before its last stage is computed the loads of tile t+1 (and descriptor t+2) are issued.
// sc = running stage counter (selects the bucket stage buffer); returns the queue fill.
template <int K, int ITEMS>
static __device__ CV_TILE_INLINE uint32_t cv_tile(const SearchParams& p, const CvPipe& pp, uint32_t t, uint32_t t_end, const uint4 d,
                                            uint32_t slot, uint32_t& sc, uint4* gq, uint32_t cnt) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t ls = d.z, le = d.w, first = d.x;
    const uint32_t n_ent = le - ls;
    const uint32_t n_stage = (n_ent + CV_STAGE - 1) / CV_STAGE;
    uint32_t wh[ITEMS], wl[ITEMS];
    for (uint32_t c = 0; c < n_stage; c++) {
        if (c + 1 < n_stage) {
            cv_issue_stage(pp, ls, le, c + 1, sc + c + 1, lane);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");  // this tile's data and the next descriptor have landed
            __syncwarp();
            if (t + 1 < t_end) {
                const uint4 dn = pp.desc[(t + 1) & 1u];
                cv_issue_windows(pp, t + 1, dn, lane);
                cv_issue_stage(pp, dn.z, dn.w, 0, sc + n_stage, lane);
                if (t + 2 < t_end) cv_issue_desc(pp, t + 2, lane);
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        }
        if (c == 0) {
            const uint32_t n_win = d.y & 0xffu, cmb = d.y >> 8;
            const uint32_t rem_nt = p.combo[cmb].rem_nt, rm = (1u << rem_nt) - 1u;
            const uint32_t* sw = pp.win + (t & 1u) * CV_WTILE;
#pragma unroll
            for (int it = 0; it < ITEMS; it++) {
                const bool have = it * 32 + lane < n_win;
                const uint32_t x = have ? sw[it * 32 + lane] : 0u;
                wh[it] = have ? (x & rm) : CV_INVALID;
                wl[it] = have ? ((x >> rem_nt) & rm) : 0u;
            }
        }
        const uint4* sb = reinterpret_cast<const uint4*>(pp.ent + ((sc + c) & 1u) * CV_STAGE);
        // groups of CV_GROUP entries; the last group of a bucket is tested in units of two entries (one LDS.128) instead
        // of being padded to eight: buckets hold ~19 entries at cfg 4, so padding cost 16 % of all POPCs
        const uint32_t n_here = min((uint32_t)CV_STAGE, n_ent - c * CV_STAGE);
        const uint32_t n_full = n_here / CV_GROUP, tail_units = ((n_here % CV_GROUP) + 1u) / 2u;
        const uint32_t ng = n_full + (tail_units ? 1u : 0u);
        const uint32_t ebase = ls + c * CV_STAGE;
        for (uint32_t g = 0; g < ng; g++) {
            bool pass_[ITEMS];
            const uint4* sg = sb + g * (CV_GROUP / 2);
            if (g < n_full || tail_units == CV_GROUP / 2) cv_group<K, ITEMS, CV_GROUP / 2>(sg, wh, wl, pass_);
            else if (tail_units == 1) cv_group<K, ITEMS, 1>(sg, wh, wl, pass_);
            else if (tail_units == 2) cv_group<K, ITEMS, 2>(sg, wh, wl, pass_);
            else cv_group<K, ITEMS, 3>(sg, wh, wl, pass_);
#pragma unroll
            for (int it = 0; it < ITEMS; it++) {
#ifdef CV_DEBUG_NO_SECOND  // timing experiment only: first level alone (no hits are produced)
                if (first != 0xffffffffu) pass_[it] = false;
#endif
                if (pass_[it]) {  // lane-private column of the queue: no vote, no compaction
                    gq[cnt * 32 + lane] = make_uint4(wh[it] | (wl[it] << 16), first + it * 32 + lane, ebase + g * CV_GROUP, slot);
                    cnt++;
                }
            }
            // some lane could overflow its column with the next group - of this tile or of the next one, which may hold more
            // windows per lane: hence CV_ITEMS, not ITEMS
            if (__any_sync(0xffffffffu, cnt > CV_QROWS - CV_ITEMS)) {
                cv_flush_items(p, pp.gwin, gq, cnt);
                cnt = 0;
            }
        }
        __syncwarp();  // every lane is done with this buffer before a later stage lands in it
    }
    sc += n_stage;
    return cnt;
}

// ------------------------------------------------------------------------------------ tile list
// Every slot with windows AND library entries is cut into warp-tiles of up to 128 window records
// aligned to the slot start, so all windows of a tile share one library bucket.  The tiles are
// listed up front ({first record, end record, bucket begin, bucket end} + slot): the verify kernel
// then needs ONE descriptor load per tile instead of walking the two directories with dependent
// loads, and carries no slot-walk state in registers (the walking version spilled 240 bytes per
// thread at 64 registers: ncu counted 1.3e9 local-memory load requests, 3x its global loads).
// (both kernels walk only the slots [s_lo, s_hi) this context owns: under slot-range sharding the rest of the directory
// is empty here, and walking / scanning all of it was a fixed cost that did not shrink with the number of GPUs)
__global__ void __launch_bounds__(256) k_ctile_count(const uint32_t* __restrict__ gdir, const uint32_t* __restrict__ dir,
                                                     uint32_t s_lo, uint32_t s_hi, uint32_t* __restrict__ tile_start,
                                                     unsigned long long* __restrict__ cand_out) {
    unsigned long long cand = 0;
    for (uint32_t s = s_lo + blockIdx.x * blockDim.x + threadIdx.x; s <= s_hi; s += gridDim.x * blockDim.x) {
        uint32_t tiles = 0;
        if (s < s_hi) {
            const uint32_t n_win = gdir[s + 1] - gdir[s], n_ent = dir[s + 1] - dir[s];
            if (n_win && n_ent) tiles = (n_win + CV_WTILE - 1) / CV_WTILE;
            cand += (unsigned long long)n_win * n_ent;
        }
        tile_start[s] = tiles;  // scanned in place afterwards; [s_hi] becomes the number of tiles
    }
    if (cand_out && cand) atomicAdd(cand_out, cand);
}

__global__ void __launch_bounds__(256) k_ctile_fill(const __grid_constant__ CBucketParams gp, const uint32_t* __restrict__ gdir,
                                                    const uint32_t* __restrict__ dir, uint32_t s_lo, uint32_t s_hi,
                                                    const uint32_t* __restrict__ tile_start, uint4* __restrict__ tile_desc,
                                                    uint32_t* __restrict__ tile_slot) {
    for (uint32_t s = s_lo + blockIdx.x * blockDim.x + threadIdx.x; s < s_hi; s += gridDim.x * blockDim.x) {
        const uint32_t t0 = tile_start[s], t1 = tile_start[s + 1];
        if (t0 == t1) continue;
        uint32_t c = 0;
        while (c + 1 < gp.n_combos && gp.combo[c + 1].dir_off <= s) c++;
        const uint32_t a = gdir[s], b = gdir[s + 1], ls = dir[s], le = dir[s + 1];
        for (uint32_t t = t0; t < t1; t++) {
            const uint32_t first = a + (t - t0) * CV_WTILE;
            // x first record, y windows in the tile (<= 128) | combination << 8, z / w bucket begin / end
            tile_desc[t] = make_uint4(first, min((uint32_t)CV_WTILE, b - first) | (c << 8), ls, le);
            tile_slot[t] = s;
        }
    }
}

// Verify, first level: warps take chunks of CV_CHUNK_TILES tiles from an atomic counter; per tile
// the (up to 4) windows of every lane stay in registers and the slot's bucket is streamed against
// them (cv_tile).  Passing {window, entry group} items go to the global queue of k_cfinish.
template <int K>
__global__ void __launch_bounds__(CV_THREADS, CV_MINBLOCKS) k_cverify(const __grid_constant__ SearchParams p,
                                                                      const uint2* __restrict__ gwin,
                                                                      const uint4* __restrict__ tile_desc,
                                                                      const uint32_t* __restrict__ tile_slot,
                                                                      const uint32_t* __restrict__ n_tiles_ptr,
                                                                      uint32_t* __restrict__ work, uint32_t slice,
                                                                      uint32_t frac_lo, uint32_t frac_hi) {
    // per warp: the item queue, two bucket stages, two tiles of window words, two descriptors + slots
    extern __shared__ __align__(16) uint4 cv_smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint4* gq = cv_smem + warp * (CV_WARP_SMEM / 16);                   // [CV_GQ]
    uint4* s_desc = gq + CV_GQ;                                         // [2]
    uint2* s_ent = reinterpret_cast<uint2*>(s_desc + 2);                // [2 * CV_STAGE]
    uint32_t* s_win = reinterpret_cast<uint32_t*>(s_ent + 2 * CV_STAGE);  // [2 * CV_WTILE]
    uint32_t* s_dslot = s_win + 2 * CV_WTILE;                           // [2] (+2 pad)
    uint32_t gn = 0;  // items this lane has queued in its column of gq (survives across tiles)
    CvPipe pp;
    pp.ent = s_ent; pp.win = s_win; pp.desc = s_desc; pp.dslot = s_dslot;
    pp.gwin = gwin; pp.ent_hl = p.ent_hl; pp.tile_desc = tile_desc; pp.tile_slot = tile_slot;
    const uint32_t n_tiles = *n_tiles_ptr;
    const uint32_t n_chunks = (n_tiles + CV_CHUNK_TILES - 1) / CV_CHUNK_TILES;
    const uint32_t ch_lo = (uint32_t)(((unsigned long long)n_chunks * frac_lo) >> 16);
    const uint32_t ch_hi = (uint32_t)(((unsigned long long)n_chunks * frac_hi) >> 16);
    uint32_t sc = 0;  // running bucket-stage counter
    for (;;) {
        uint32_t ch = 0;
        if (lane == 0) ch = ch_lo + atomicAdd(work + slice, 1u);
        ch = __shfl_sync(0xffffffffu, ch, 0);
        if (ch >= ch_hi) break;
        const uint32_t t_end = min((ch + 1) * CV_CHUNK_TILES, n_tiles);
        uint32_t t = ch * CV_CHUNK_TILES;
        // prologue of the chunk: the only exposed latency chain (descriptor, then its data)
        {
            const uint4 d0 = __ldg(tile_desc + t);
            if (lane == 0) {
                pp.desc[t & 1u] = d0;
                pp.dslot[t & 1u] = __ldg(tile_slot + t);
            }
            cv_issue_windows(pp, t, d0, lane);
            cv_issue_stage(pp, d0.z, d0.w, 0, sc, lane);
            if (t + 1 < t_end) cv_issue_desc(pp, t + 1, lane);
            asm volatile("cp.async.commit_group;" ::: "memory");
            __syncwarp();
        }
        for (; t < t_end; t++) {
            const uint4 d_cur = pp.desc[t & 1u];
            const uint32_t slot_cur = pp.dslot[t & 1u];
            switch (((d_cur.y & 0xffu) + 31u) / 32u) {
                case 1: gn = cv_tile<K, 1>(p, pp, t, t_end, d_cur, slot_cur, sc, gq, gn); break;
                case 2: gn = cv_tile<K, 2>(p, pp, t, t_end, d_cur, slot_cur, sc, gq, gn); break;
                case 3: gn = cv_tile<K, 3>(p, pp, t, t_end, d_cur, slot_cur, sc, gq, gn); break;
                default: gn = cv_tile<K, 4>(p, pp, t, t_end, d_cur, slot_cur, sc, gq, gn); break;
            }
        }
    }
    cv_flush_items(p, gwin, gq, gn);  // whatever is still queued
}

#ifndef CF_MINBLOCKS
#define CF_MINBLOCKS 4
#endif
// Second level + hit resolution as a kernel of their own: every warp takes batches of 32 items of
// the global queue, one per lane (8 entries each, tested against the item's window), queues the
// passing pairs and resolves them 32 at a time (ownership, PAM, one atomic per batch).  All of it
// is dependent random loads; here they overlap across ~10^8 items instead of stalling the POPC loop.
__global__ void __launch_bounds__(CV_THREADS, CF_MINBLOCKS) k_cfinish(const __grid_constant__ SearchParams p, const uint2* __restrict__ gwin) {
    __shared__ uint4 s_q[CV_WARPS][CV_WQ];
    __shared__ uint32_t s_qn[CV_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint4* q = s_q[warp];
    uint32_t* qn = &s_qn[warp];
    if (lane == 0) *qn = 0;
    __syncwarp();
    const unsigned long long queued = p.count[4];
    if (queued > p.item_cap && blockIdx.x == 0 && threadIdx.x == 0) atomicMax(p.count + 5, queued);  // batches were dropped: bc_search retries
    // after an overflow the queue has holes (a dropped reservation is never written) and the attempt is repeated anyway
    const unsigned long long n_items = queued <= p.item_cap ? queued : 0ull;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * CV_WARPS;
    for (unsigned long long b0 = ((unsigned long long)blockIdx.x * CV_WARPS + warp) * 32ull; b0 < n_items; b0 += n_warps * 32ull) {
        const uint32_t n = (uint32_t)min(32ull, n_items - b0);
        cv_resolve_groups(p, gwin, p.items + b0, n, q, qn, CV_ITEM_POS != 0);
    }
    __syncwarp();
    const uint32_t nq = min(*qn, (uint32_t)CV_WQ);
    if (nq) cv_resolve(p, gwin, q, nq, CV_ITEM_POS != 0);
}

// ------------------------------------------------------------------------------------------ host
// BC_WIN_COUNT=red (environment, A/B runs only): count the windows per slot with one global RED per record
// (k_ccount, the round-1 form) instead of k_cbincount + k_cslotcount
static bool cj_red_count() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("BC_WIN_COUNT");
        v = (e && !strcmp(e, "red")) ? 1 : 0;
    }
    return v != 0;
}


#define JCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            if (getenv("BC_DEBUG")) fprintf(stderr, "bc_cjoin.cu:%d: %s\n", __LINE__, cudaGetErrorString(e__)); \
            return e__;                                                                             \
        }                                                                                           \
    } while (0)

static cudaError_t cj_ensure_bins(JoinWorkspace& ws, uint32_t n_bins, uint64_t max_chunks) {
    const uint64_t bin_words = 2ull * (n_bins + 2) + max_chunks + (n_bins + 8) / 4 + 4;
    if (bin_words > ws.bin_cap) {
        if (ws.d_bin_cursor) cudaFree(ws.d_bin_cursor);
        ws.d_bin_cursor = nullptr;
        ws.bin_cap = 0;
        JCK(cudaMalloc(&ws.d_bin_cursor, bin_words * sizeof(uint32_t)));
        ws.bin_cap = bin_words;
    }
    return cudaSuccess;
}


template <bool LIB>
static cudaError_t cj_launch_place(const CBucketParams& gp, const uint2* tmp, const uint32_t* bin_start, const uint8_t* bin_combo,
                                   const uint32_t* chunk_bin, uint32_t n_bins, uint32_t* gcursor, uint2* out, uint32_t* out_id,
                                   const uint32_t* n_rec_ptr, uint32_t* work, uint32_t max_sub, size_t smem_b, int sm_count,
                                   cudaStream_t st) {
    if (CJ_PLACE_FORM == 0) {
        JCK(cudaFuncSetAttribute(k_cplace<LIB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
        k_cplace<LIB><<<(uint32_t)sm_count * 2u, CJ_THREADS, smem_b, st>>>(gp, tmp, bin_start, bin_combo, chunk_bin, n_bins, gcursor, out,
                                                                        out_id, n_rec_ptr, work, max_sub);
    } else if (CJ_PLACE_FORM == 1) {
        const size_t smem = 2 * (size_t)CJ_THREADS * 8 * 8 + (3 * (size_t)max_sub + CJ_THREADS / 32) * 4;
        JCK(cudaFuncSetAttribute(k_cplace_bulk<LIB, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_cplace_bulk<LIB, 8><<<(uint32_t)sm_count * 2u, CJ_THREADS, smem, st>>>(gp, tmp, bin_start, bin_combo, chunk_bin, n_bins, gcursor, out, out_id,
                                                                              n_rec_ptr, work, max_sub);
    } else {
        const size_t smem = 2 * (size_t)CJ_THREADS * 16 * 8 + (3 * (size_t)max_sub + CJ_THREADS / 32) * 4;
        JCK(cudaFuncSetAttribute(k_cplace_bulk<LIB, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_cplace_bulk<LIB, 16><<<(uint32_t)sm_count, CJ_THREADS, smem, st>>>(gp, tmp, bin_start, bin_combo, chunk_bin, n_bins, gcursor, out, out_id,
                                                                           n_rec_ptr, work, max_sub);
    }
    return cudaGetLastError();
}

static size_t cj_smem_a() {
    return (3 * CJ_MAX_BINS + 2 * (CJ_CHUNK / 32 + 2) + CJ_THREADS / 32 + CJ_LUT_WORDS) * 4 + (size_t)CJ_CHUNK * 8 +
           (size_t)CJ_CHUNK * 2;
}

// Library index of the compact join path, built with the same two radix passes as the window sort:
// count (RED) -> scan -> pass A (entries grouped by bin) -> pass B (final key order, written as
// ent_hl = rem planes {Rh, Rl} and ent_id = entry).  `tmp` must hold one uint2 per (entry, combination).
cudaError_t bc_cindex_build(JoinWorkspace& ws, const IndexParams& ip, uint32_t n_combos, uint32_t n_bins, uint32_t* d_dir,
                            uint64_t dir_slots, uint32_t* d_cursor, uint32_t* d_scan_tmp, uint2* tmp, uint2* ent_hl,
                            uint32_t* ent_id, int sm_count, cudaStream_t st) {
    const uint32_t n_slots = (uint32_t)(dir_slots - 1);
    JCK(cudaMemsetAsync(d_dir, 0, dir_slots * sizeof(uint32_t), st));
    if (ip.n_entries == 0) return cudaSuccess;
    CBucketParams gp;
    memset(&gp, 0, sizeof gp);
    gp.qh = ip.qh; gp.ql = ip.ql; gp.sn = ip.sn;
    gp.n_entries = ip.n_entries;
    gp.lib_has_n = ip.lib_has_n;
    gp.L = ip.L;
    gp.n_combos = n_combos;
    gp.slot_lo = ip.slot_lo; gp.slot_hi = ip.slot_hi;
    gp.bin_aligned = ip.bin_aligned;
    memcpy(gp.combo, ip.combo, sizeof gp.combo);
    uint32_t max_low = 0;
    for (uint32_t c = 0; c < n_combos; c++) {
        const uint32_t low = 2u * gp.combo[c].key_nt - gp.combo[c].top_bits;
        if (low > max_low) max_low = low;
    }
    const uint64_t max_chunks = ((uint64_t)ip.n_entries * n_combos + CJ_PLACE_CHUNK - 1) / CJ_PLACE_CHUNK + 1;
    JCK(cj_ensure_bins(ws, n_bins, max_chunks));
    uint32_t* d_bin_cursor = ws.d_bin_cursor;
    uint32_t* d_bin_start = d_bin_cursor + (n_bins + 2);
    uint32_t* d_chunk_bin = d_bin_start + (n_bins + 2);
    uint8_t* d_bin_combo = reinterpret_cast<uint8_t*>(d_chunk_bin + max_chunks);
    if (!ws.d_lut) JCK(cudaMalloc(&ws.d_lut, (size_t)BC_MAX_COMBOS * CJ_LUT_WORDS * sizeof(uint32_t)));
    if (!ws.d_work) JCK(cudaMalloc(&ws.d_work, BC_SINK_SLICES * sizeof(uint32_t)));
    if (CJ_LUT_BITS != 8 && ip.L > 2 * CJ_LUT_BITS) return cudaErrorInvalidValue;  // choose_scheme keeps such spacers off this path
    k_clut_build<<<(n_combos * CJ_LUT_WORDS + 255) / 256, 256, 0, st>>>(gp, ws.d_lut);
    JCK(cudaGetLastError());
    const size_t smem_a = cj_smem_a();
    const uint32_t max_sub = 1u << max_low;
    const size_t smem_b = (3 * (size_t)max_sub + CJ_THREADS / 32) * 4 + (size_t)CJ_CHUNK * 8;
    JCK(cudaFuncSetAttribute(k_cbin<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    uint32_t bx = (ip.n_entries + CJ_CHUNK - 1) / CJ_CHUNK;
    if (bx > (uint32_t)sm_count * 2u) bx = (uint32_t)sm_count * 2u;
    const uint32_t* n_rec_ptr;
    if (cj_red_count()) {  // round-1 form: one global RED per (entry, combination)
        uint32_t gx = (ip.n_entries + 255) / 256;
        if (gx > (uint32_t)sm_count * 8u) gx = (uint32_t)sm_count * 8u;
        k_ccount<true><<<dim3(gx, n_combos), 256, 0, st>>>(gp, ws.d_lut, d_dir);
        JCK(cudaGetLastError());
        JCK(bc_exclusive_scan(d_dir, dir_slots, d_scan_tmp, st));
        JCK(cudaMemcpyAsync(d_cursor, d_dir, dir_slots * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        k_cbin_init<<<(n_bins + 256) / 256, 256, 0, st>>>(gp, d_dir, n_bins, n_slots, d_bin_start, d_bin_cursor, d_bin_combo);
        JCK(cudaGetLastError());
        k_cbin<true><<<bx, CJ_THREADS, smem_a, st>>>(gp, ws.d_lut, d_bin_cursor, tmp);
        JCK(cudaGetLastError());
        n_rec_ptr = d_dir + n_slots;
    } else {               // bin totals and the slot histogram from shared-memory histograms (see k_cbincount)
        JCK(cudaMemsetAsync(d_bin_cursor, 0, (n_bins + 1) * sizeof(uint32_t), st));
        if (ip.slot_lo == 0 && ip.slot_hi == n_slots) k_cbincount<true, false><<<(uint32_t)sm_count * 2u, CB_THREADS, 0, st>>>(gp, ws.d_lut, d_bin_cursor);
        else k_cbincount<true, true><<<(uint32_t)sm_count * 2u, CB_THREADS, 0, st>>>(gp, ws.d_lut, d_bin_cursor);
        JCK(cudaGetLastError());
        JCK(bc_exclusive_scan(d_bin_cursor, (uint64_t)n_bins + 1, d_scan_tmp, st));
        k_cbin_init2<<<(n_bins + 256) / 256, 256, 0, st>>>(gp, n_bins, d_bin_cursor, d_bin_start, d_bin_combo);
        JCK(cudaGetLastError());
        k_cbin<true><<<bx, CJ_THREADS, smem_a, st>>>(gp, ws.d_lut, d_bin_cursor, tmp);
        JCK(cudaGetLastError());
        JCK(cudaMemsetAsync(ws.d_work, 0, BC_SINK_SLICES * sizeof(uint32_t), st));
        k_cslotcount<<<(uint32_t)sm_count * 2u, CS_THREADS, max_sub * sizeof(uint32_t), st>>>(gp, tmp, d_bin_start, d_bin_combo, n_bins, d_dir,
                                                                                        ws.d_work);
        JCK(cudaGetLastError());
        {   // only the owned slot range holds entries (the rest of d_dir stays zero)
            const uint64_t r_n = (uint64_t)(ip.slot_hi - ip.slot_lo) + 1;
            JCK(bc_exclusive_scan(d_dir + ip.slot_lo, r_n, d_scan_tmp, st));
            JCK(cudaMemcpyAsync(d_cursor + ip.slot_lo, d_dir + ip.slot_lo, r_n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        }
        n_rec_ptr = d_bin_start + n_bins;
    }
    k_cchunk_bins<<<(uint32_t)((max_chunks + 255) / 256), 256, 0, st>>>(d_bin_start, n_bins, n_rec_ptr, d_chunk_bin);
    JCK(cudaGetLastError());
    JCK(cudaMemsetAsync(ws.d_work, 0, BC_SINK_SLICES * sizeof(uint32_t), st));
    JCK(cj_launch_place<true>(gp, tmp, d_bin_start, d_bin_combo, d_chunk_bin, n_bins, d_cursor, ent_hl, ent_id, n_rec_ptr,
                              ws.d_work, max_sub, smem_b, sm_count, st));
    bc_launch_counter += 7;
    return cudaSuccess;
}

// CTAs of k_cverify per SM: CV_MINBLOCKS fills the register file; BC_VERIFY_CTAS (environment, experiments only)
// lowers it
static uint32_t cv_ctas_per_sm() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("BC_VERIFY_CTAS");
        v = e ? atoi(e) : CV_MINBLOCKS;
        if (v < 1 || v > CV_MINBLOCKS) v = CV_MINBLOCKS;
    }
    return (uint32_t)v;
}

cudaError_t bc_cjoin_search(JoinWorkspace& ws, const SearchParams& p, uint64_t dir_slots, uint32_t n_bins, int sm_count,
                            cudaStream_t st, uint32_t* launches, HitSink* sink) {
    const uint32_t launches0 = bc_launch_counter;
    ws.ms_join_kernels = ws.ms_bucket_kernels = 0;
    if (!ws.ev_a) JCK(cudaEventCreate(&ws.ev_a));
    if (!ws.ev_b) JCK(cudaEventCreate(&ws.ev_b));
    if (!ws.ev_c) JCK(cudaEventCreate(&ws.ev_c));
    for (int i = 0; i < 6; i++) {
        if (!ws.ev_k[i]) JCK(cudaEventCreate(&ws.ev_k[i]));
        ws.ms_kernel[i] = 0;
    }
    const uint32_t n_slots = (uint32_t)(dir_slots - 1);

    uint32_t max_low = 0;
    for (uint32_t c = 0; c < p.n_combos; c++) {
        const uint32_t low = 2u * p.combo[c].key_nt - p.combo[c].top_bits;
        if (low > max_low) max_low = low;
    }
    // genome positions per pass: two 8-byte record arrays within the workspace budget
    const uint64_t span = (uint64_t)p.pos_end - p.pos_begin;
    uint64_t chunk = span ? span : 1;
    if (p.join_chunk && chunk > p.join_chunk) chunk = p.join_chunk;
    const uint64_t cap8 = ws.gwin_cap * 2;  // the workspace is sized in 16-byte records; two uint2 per slot
    if (chunk * p.n_combos > cap8) {
        size_t free_b = 0, total_b = 0;
        JCK(cudaMemGetInfo(&free_b, &total_b));
        uint64_t budget = ((uint64_t)free_b + ((ws.d_gwin ? 1 : 0) + (ws.d_gtmp ? 1 : 0)) * ws.gwin_cap * sizeof(uint4)) / 2;
        if (budget > (96ull << 30)) budget = 96ull << 30;
        const uint64_t fit = budget / (2 * sizeof(uint2)) / p.n_combos;
        if (chunk > fit) chunk = fit;
        if (chunk < 1) chunk = 1;
    }
    if (chunk * p.n_combos >= (1ull << 32) - 65536) chunk = ((1ull << 32) - 65536) / p.n_combos;
    const uint64_t rec16_needed = (chunk * p.n_combos + 1) / 2 + 1;
    if (rec16_needed > ws.gwin_cap || !ws.d_gtmp) {
        const uint64_t want = rec16_needed > ws.gwin_cap ? rec16_needed : ws.gwin_cap;
        if (ws.d_gwin) cudaFree(ws.d_gwin);
        if (ws.d_gtmp) cudaFree(ws.d_gtmp);
        ws.d_gwin = ws.d_gtmp = nullptr;
        ws.gwin_cap = 0;
        JCK(cudaMalloc(&ws.d_gwin, (want + 1) * sizeof(uint4)));
        JCK(cudaMalloc(&ws.d_gtmp, (want + 1) * sizeof(uint4)));
        ws.gwin_cap = want;
    }
    // per-bin tables live in the bin cursor allocation: cursor | start (+1) | chunk table | combination bytes
    const uint64_t max_chunks = (chunk * p.n_combos + CJ_PLACE_CHUNK - 1) / CJ_PLACE_CHUNK + 1;
    JCK(cj_ensure_bins(ws, n_bins, max_chunks));
    uint32_t* d_bin_cursor = ws.d_bin_cursor;
    uint32_t* d_bin_start = d_bin_cursor + (n_bins + 2);
    uint32_t* d_chunk_bin = d_bin_start + (n_bins + 2);
    uint8_t* d_bin_combo = reinterpret_cast<uint8_t*>(d_chunk_bin + max_chunks);
    if (dir_slots > ws.gdir_cap) {
        if (ws.d_gdir) cudaFree(ws.d_gdir);
        if (ws.d_gcursor) cudaFree(ws.d_gcursor);
        ws.d_gdir = ws.d_gcursor = nullptr;
        ws.gdir_cap = 0;
        JCK(cudaMalloc(&ws.d_gdir, dir_slots * 4));
        JCK(cudaMalloc(&ws.d_gcursor, dir_slots * 4));
        ws.gdir_cap = dir_slots;
    }
    const uint64_t tmp_words = bc_scan_tmp_words(dir_slots);
    if (tmp_words > ws.scan_tmp_cap) {
        if (ws.d_scan_tmp) cudaFree(ws.d_scan_tmp);
        ws.d_scan_tmp = nullptr;
        ws.scan_tmp_cap = 0;
        JCK(cudaMalloc(&ws.d_scan_tmp, tmp_words * 4));
        ws.scan_tmp_cap = tmp_words;
    }
    if (!ws.d_work) JCK(cudaMalloc(&ws.d_work, BC_SINK_SLICES * sizeof(uint32_t)));
    // global item queue between k_cverify and k_cfinish: about one item per candidate pair.  Sized from the hit
    // buffer, or from the demand a previous attempt measured (ws.item_want, set by bc_search after an overflow)
    {
        uint64_t want = 2 * p.cap + (1ull << 16);
        if (want > (1ull << 29)) want = 1ull << 29;
        if (ws.item_want > want) want = ws.item_want;
        want = (want + 31ull) & ~31ull;
        if (want > ws.item_cap) {
            if (ws.d_items) cudaFree(ws.d_items);
            ws.d_items = nullptr;
            ws.item_cap = 0;
            JCK(cudaMalloc(&ws.d_items, want * sizeof(uint4)));
            ws.item_cap = want;
        }
    }
    // tile list: at most one ragged tile per non-empty slot plus the full ones
    {
        const uint64_t max_tiles = chunk * p.n_combos / CV_WTILE + n_slots + 2;
        if (max_tiles > ws.tile_cap) {
            if (ws.d_tile_desc) cudaFree(ws.d_tile_desc);
            if (ws.d_tile_slot) cudaFree(ws.d_tile_slot);
            ws.d_tile_desc = nullptr; ws.d_tile_slot = nullptr;
            ws.tile_cap = 0;
            JCK(cudaMalloc(&ws.d_tile_desc, max_tiles * sizeof(uint4)));
            JCK(cudaMalloc(&ws.d_tile_slot, max_tiles * sizeof(uint32_t)));
            ws.tile_cap = max_tiles;
        }
        if (dir_slots > ws.tile_start_cap) {
            if (ws.d_tile_start) cudaFree(ws.d_tile_start);
            ws.d_tile_start = nullptr;
            ws.tile_start_cap = 0;
            JCK(cudaMalloc(&ws.d_tile_start, dir_slots * sizeof(uint32_t)));
            ws.tile_start_cap = dir_slots;
        }
    }
    SearchParams pv = p;
    pv.items = ws.d_items;
    pv.item_cap = ws.item_cap;

    CBucketParams gp;
    memset(&gp, 0, sizeof gp);
    gp.H = p.H; gp.Lo = p.Lo; gp.B = p.B;
    gp.lib_dir = p.dir;
    gp.n_words = p.n_words;
    gp.L = p.L;
    gp.n_combos = p.n_combos;
    memcpy(gp.combo, p.combo, sizeof gp.combo);
    gp.prune = p.dir_entries < (dir_slots - 1) * 2 ? 1u : 0u;
    gp.gate_first = p.gate_first;
    gp.slot_lo = p.slot_lo; gp.slot_hi = p.slot_hi;
    gp.bin_aligned = p.bin_aligned;
    gp.P = p.P; gp.pam_dir = p.pam_dir;
    for (int i = 0; i < 8; i++) gp.pam_sets[i] = p.pam_sets[i];

    JCK(cudaFuncSetAttribute(k_cverify<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_WARPS * CV_WARP_SMEM));
    JCK(cudaFuncSetAttribute(k_cverify<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_WARPS * CV_WARP_SMEM));
    JCK(cudaFuncSetAttribute(k_cverify<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_WARPS * CV_WARP_SMEM));
    JCK(cudaFuncSetAttribute(k_cverify<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_WARPS * CV_WARP_SMEM));
    if (!ws.d_lut) JCK(cudaMalloc(&ws.d_lut, (size_t)BC_MAX_COMBOS * CJ_LUT_WORDS * sizeof(uint32_t)));
    if (CJ_LUT_BITS != 8 && p.L > 2 * CJ_LUT_BITS) return cudaErrorInvalidValue;
    k_clut_build<<<(p.n_combos * CJ_LUT_WORDS + 255) / 256, 256, 0, st>>>(gp, ws.d_lut);
    JCK(cudaGetLastError());
    bc_launch_counter += 1;
    const size_t smem_a = cj_smem_a();
    const uint32_t max_sub = 1u << max_low;
    const size_t smem_b = (3 * (size_t)max_sub + CJ_THREADS / 32) * 4 + (size_t)CJ_CHUNK * 8;
    JCK(cudaFuncSetAttribute(k_cbin<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    uint2* d_tmp = reinterpret_cast<uint2*>(ws.d_gtmp);
    uint2* d_win = reinterpret_cast<uint2*>(ws.d_gwin);
    const bool one_pass = max_low == 0;  // every bin is one slot: pass A writes the final array

    for (uint64_t begin = p.pos_begin; begin < p.pos_end; begin += chunk) {
        gp.pos_begin = (uint32_t)begin;
        gp.pos_end = (uint32_t)((begin + chunk < p.pos_end) ? begin + chunk : p.pos_end);
        const uint32_t npos = gp.pos_end - gp.pos_begin;
        uint32_t gx = (npos + 255) / 256;
        const uint32_t maxb = (uint32_t)sm_count * 8u;
        if (gx > maxb) gx = maxb;
        // the slots this context owns (everything unless the directory is sharded): the directory-sized passes below
        // (clear, scan, cursor copy, tile list) only touch this range (+ its end sentinel)
        const bool red_count = cj_red_count();
        const uint32_t r_lo = red_count ? 0u : p.slot_lo, r_hi = red_count ? n_slots : p.slot_hi;
        const uint64_t r_n = (uint64_t)(r_hi - r_lo) + 1;
        JCK(cudaMemsetAsync(ws.d_gdir + r_lo, 0, r_n * 4, st));
        JCK(cudaEventRecord(ws.ev_c, st));
        uint32_t bx = (npos + CJ_CHUNK - 1) / CJ_CHUNK;
        if (bx > (uint32_t)sm_count * 2u) bx = (uint32_t)sm_count * 2u;
        if (red_count) {  // round-1 form: one global RED per (window, combination), kept for A/B runs (BC_WIN_COUNT=red)
            k_ccount<false><<<dim3(gx, p.n_combos), 256, 0, st>>>(gp, ws.d_lut, ws.d_gdir);
            JCK(cudaGetLastError());
            JCK(cudaEventRecord(ws.ev_k[0], st));  // end of the count kernel
            JCK(bc_exclusive_scan(ws.d_gdir, dir_slots, ws.d_scan_tmp, st));
            if (!one_pass) JCK(cudaMemcpyAsync(ws.d_gcursor, ws.d_gdir, dir_slots * 4, cudaMemcpyDeviceToDevice, st));
            k_cbin_init<<<(n_bins + 256) / 256, 256, 0, st>>>(gp, ws.d_gdir, n_bins, n_slots, d_bin_start, d_bin_cursor, d_bin_combo);
            JCK(cudaGetLastError());
        } else {          // bin totals from shared-memory histograms; the slot histogram follows pass A
            JCK(cudaMemsetAsync(d_bin_cursor, 0, (n_bins + 1) * sizeof(uint32_t), st));
            if (p.slot_lo == 0 && p.slot_hi == n_slots) k_cbincount<false, false><<<(uint32_t)sm_count * 2u, CB_THREADS, 0, st>>>(gp, ws.d_lut, d_bin_cursor);
            else k_cbincount<false, true><<<(uint32_t)sm_count * 2u, CB_THREADS, 0, st>>>(gp, ws.d_lut, d_bin_cursor);
            JCK(cudaGetLastError());
            JCK(cudaEventRecord(ws.ev_k[0], st));  // end of the bin count
            JCK(bc_exclusive_scan(d_bin_cursor, (uint64_t)n_bins + 1, ws.d_scan_tmp, st));
            k_cbin_init2<<<(n_bins + 256) / 256, 256, 0, st>>>(gp, n_bins, d_bin_cursor, d_bin_start, d_bin_combo);
            JCK(cudaGetLastError());
        }
        JCK(cudaEventRecord(ws.ev_k[1], st));  // start of pass A
        k_cbin<false><<<bx, CJ_THREADS, smem_a, st>>>(gp, ws.d_lut, d_bin_cursor, one_pass ? d_win : d_tmp);
        JCK(cudaGetLastError());
        JCK(cudaEventRecord(ws.ev_k[2], st));  // end of pass A
        bc_launch_counter += 3;
        const uint32_t* n_rec_ptr = red_count ? ws.d_gdir + n_slots : d_bin_start + n_bins;
        if (!red_count) {   // slot histogram of the binned records, then the slot starts
            JCK(cudaMemsetAsync(ws.d_work, 0, BC_SINK_SLICES * sizeof(uint32_t), st));
            k_cslotcount<<<(uint32_t)sm_count * 2u, CS_THREADS, max_sub * sizeof(uint32_t), st>>>(gp, one_pass ? d_win : d_tmp, d_bin_start, d_bin_combo, n_bins,
                                                                                            ws.d_gdir, ws.d_work);
            JCK(cudaGetLastError());
            JCK(bc_exclusive_scan(ws.d_gdir + r_lo, r_n, ws.d_scan_tmp, st));
            JCK(cudaMemcpyAsync(ws.d_gcursor + r_lo, ws.d_gdir + r_lo, r_n * 4, cudaMemcpyDeviceToDevice, st));
            bc_launch_counter += 1;
        }
        JCK(cudaEventRecord(ws.ev_k[5], st));  // end of the slot count (start of pass B)
        if (!one_pass) {
            k_cchunk_bins<<<(uint32_t)((max_chunks + 255) / 256), 256, 0, st>>>(d_bin_start, n_bins, n_rec_ptr, d_chunk_bin);
            JCK(cudaGetLastError());
            JCK(cudaMemsetAsync(ws.d_work, 0, BC_SINK_SLICES * sizeof(uint32_t), st));
            JCK(cj_launch_place<false>(gp, d_tmp, d_bin_start, d_bin_combo, d_chunk_bin, n_bins, ws.d_gcursor, d_win, nullptr,
                                       n_rec_ptr, ws.d_work, max_sub, smem_b, sm_count, st));
            bc_launch_counter += 2;
        }
        JCK(cudaEventRecord(ws.ev_k[3], st));      // end of pass B
        {   // tile list of this pass
            const uint32_t tg = (uint32_t)sm_count * 8u;
            k_ctile_count<<<tg, 256, 0, st>>>(ws.d_gdir, p.dir, r_lo, r_hi, ws.d_tile_start, p.count_candidates ? p.count + 1 : nullptr);
            JCK(cudaGetLastError());
            JCK(bc_exclusive_scan(ws.d_tile_start + r_lo, r_n, ws.d_scan_tmp, st));
            k_ctile_fill<<<tg, 256, 0, st>>>(gp, ws.d_gdir, p.dir, r_lo, r_hi, ws.d_tile_start, ws.d_tile_desc, ws.d_tile_slot);
            JCK(cudaGetLastError());
            bc_launch_counter += 2;
        }
        JCK(cudaEventRecord(ws.ev_a, st));
        // Streamed delivery.  Short verify stages (a slot-range shard of an 8-GPU job runs ~3 ms) get fewer slices:
        // every slice costs a launch tail and a host round trip
        const double est_rec = (double)npos * p.n_combos * ((double)(p.slot_hi - p.slot_lo) / (double)(n_slots ? n_slots : 1));
        // Equal slices: at cfg 4 the verify + finish kernels (18 ms) take as long as the D2H copy of their 0.95 GB of
        // records, so every slice's copy hides behind the next slice's kernels and only the last one (1/8) is exposed.
        // (Round 1 halved the slices - 1/2, 1/4, ... - when verification took four times longer than the copy; with the
        // round-2 kernels the first half's copy alone outlasted the rest of the search: +8 ms end to end.)
        // A device destination (the peer merge: NVLink, ~10x the PCIe rate) gets fewer slices.
        const uint32_t n_slices = !sink ? 1u
                                  : (sink->host && sink->is_device && !sink->fn) ? (est_rec > 6e8 ? 4u : 2u)
                                  : est_rec > 6e8 ? (uint32_t)BC_SINK_SLICES : est_rec > 1e8 ? 4u : 2u;
        JCK(cudaMemsetAsync(ws.d_work, 0, BC_SINK_SLICES * sizeof(uint32_t), st));
        for (uint32_t s = 0; s < n_slices; s++) {
            const uint32_t f_lo = 65536u * s / n_slices, f_hi = 65536u * (s + 1) / n_slices;
            const uint32_t dgrid = (uint32_t)sm_count * cv_ctas_per_sm(), lo = f_lo;
            JCK(cudaMemsetAsync(p.count + 4, 0, sizeof(unsigned long long), st));
            switch (p.k) {  // (the dynamic shared-memory opt-in was set once, above)
                case 0: k_cverify<0><<<dgrid, CV_THREADS, CV_WARPS * CV_WARP_SMEM, st>>>(pv, d_win, ws.d_tile_desc, ws.d_tile_slot, ws.d_tile_start + r_hi, ws.d_work, s, lo, f_hi); break;
                case 1: k_cverify<1><<<dgrid, CV_THREADS, CV_WARPS * CV_WARP_SMEM, st>>>(pv, d_win, ws.d_tile_desc, ws.d_tile_slot, ws.d_tile_start + r_hi, ws.d_work, s, lo, f_hi); break;
                case 2: k_cverify<2><<<dgrid, CV_THREADS, CV_WARPS * CV_WARP_SMEM, st>>>(pv, d_win, ws.d_tile_desc, ws.d_tile_slot, ws.d_tile_start + r_hi, ws.d_work, s, lo, f_hi); break;
                default: k_cverify<3><<<dgrid, CV_THREADS, CV_WARPS * CV_WARP_SMEM, st>>>(pv, d_win, ws.d_tile_desc, ws.d_tile_slot, ws.d_tile_start + r_hi, ws.d_work, s, lo, f_hi); break;
            }
            JCK(cudaGetLastError());
            if (s + 1 == n_slices) JCK(cudaEventRecord(ws.ev_k[4], st));  // end of the (last) first-level kernel
            k_cfinish<<<(uint32_t)sm_count * CF_MINBLOCKS, CV_THREADS, 0, st>>>(pv, d_win);
            JCK(cudaGetLastError());
            bc_launch_counter += 1;
            if (sink) {
                JCK(cudaMemcpyAsync(sink->h_counts + s, p.count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                JCK(cudaEventRecord(sink->ev[s], st));
            }
        }
        JCK(cudaEventRecord(ws.ev_b, st));
        bc_launch_counter += n_slices;
        if (sink) JCK(bc_sink_deliver(sink, p, n_slices));
        JCK(cudaEventSynchronize(ws.ev_b));
        float ms = 0;
        JCK(cudaEventElapsedTime(&ms, ws.ev_a, ws.ev_b));
        ws.ms_join_kernels += ms;
        JCK(cudaEventElapsedTime(&ms, ws.ev_c, ws.ev_a));
        ws.ms_bucket_kernels += ms;
        // per-kernel split (count, pass A, pass B, tile list, first-level verify, finish); the verify
        // figures are those of the last slice, i.e. of the whole stage when nothing is streamed
        JCK(cudaEventElapsedTime(&ms, ws.ev_c, ws.ev_k[0])); ws.ms_kernel[0] += ms;
        JCK(cudaEventElapsedTime(&ms, ws.ev_k[1], ws.ev_k[2])); ws.ms_kernel[1] += ms;
        JCK(cudaEventElapsedTime(&ms, ws.ev_k[2], ws.ev_k[5])); ws.ms_kernel[0] += ms;  // slot count: part of "count"
        JCK(cudaEventElapsedTime(&ms, ws.ev_k[5], ws.ev_k[3])); ws.ms_kernel[2] += ms;
        JCK(cudaEventElapsedTime(&ms, ws.ev_k[3], ws.ev_a)); ws.ms_kernel[3] += ms;
        if (n_slices == 1) {
            JCK(cudaEventElapsedTime(&ms, ws.ev_a, ws.ev_k[4])); ws.ms_kernel[4] += ms;
            JCK(cudaEventElapsedTime(&ms, ws.ev_k[4], ws.ev_b)); ws.ms_kernel[5] += ms;
        }
    }
    *launches = bc_launch_counter - launches0;
    return cudaSuccess;
}
