// bc_join.cu - K3-join: the scan for libraries whose seed buckets are dense (cfg 3/4).
//
// The probe kernel walks a library bucket once per genome window at a random address, so with
// tens to hundreds of entries per bucket every window re-reads its bucket through L2/HBM.  Here
// the genome side is sorted by the same seed keys instead (a sort-merge join):
//
//   pass 1  k_bucket<0/1>   every genome window, for every seed combination c, becomes a 16-byte
//                           record {dev position, wh, wl, directory slot} placed in slot order:
//                           histogram -> exclusive scan -> scatter (counting sort, not stable).
//   pass 2  k_verify_dense  slots with many windows: warp-tiles of sorted records aligned to the
//                           slot start; the library bucket [dir[slot], dir[slot+1]) is staged
//                           through shared memory and streamed once per tile against the resident
//                           windows of every lane: 2 LOP3 + POPC + a min-reduce per pair.
//           k_verify_sparse the records of small slots, one record per lane.
//                           Pairs that pass are queued per warp and resolved 32 at a time
//                           (ownership, PAM, one atomic per batch, coalesced 16-byte records).
//
// Algorithmic HBM traffic: one 16 B write + one 16 B read per (window, combination), the three
// genome planes twice per combination, the library index once.
#include "bc_join.h"

#include <string.h>

#include "bc_kernels.h"

struct BucketParams {
    const uint32_t* H;
    const uint32_t* Lo;
    const uint32_t* B;
    const uint32_t* lib_dir;      // library directory (to skip windows whose bucket is empty)
    uint32_t pos_begin, pos_end;  // dev positions handled by this chunk of the genome
    uint32_t L, n_combos, prune, gate_first;
    uint32_t slot_lo, slot_hi;    // slot-range sharding
    uint32_t P, pam_dir, pam_sets[8];
    ComboDesc combo[BC_MAX_COMBOS];
};

// Pass 1.  PASS 0 counts (RED), PASS 1 scatters straight to the final slot (one returning atomic
// and one isolated 16-byte store per record).  Windows touching a non-ACGT base or a contig end are
// dropped here, so pass 2 never sees them.  PASS 1 is the fallback of the radix scatter below.
// The grid is (x = genome chunks, y = combination) and CTAs are dispatched x-fastest, so the
// persistent CTAs of ONE combination run at a time: its 4^key_nt write fronts (2 MB of sectors at
// cfg 4) stay in L2.  A position-major variant of PASS 1 (one thread = one window through all
// combinations, 5-10 independent atomics in flight) was measured at 35-38 ms against 27.7 ms for
// this form; a two-level coarse/fine scatter with returning atomics in both levels at 37 + 37 ms
// against 70 ms (cfg 4 at b=6).
template <int PASS>
__global__ void __launch_bounds__(256) k_bucket(const __grid_constant__ BucketParams gp,
                                                uint32_t* __restrict__ gdir_or_cursor, uint4* __restrict__ gwin) {
    const uint32_t lm = bc_lmask(gp.L);
    const uint32_t c = blockIdx.y;
    const ComboDesc& cd = gp.combo[c];
    if (!bc_combo_in_range(cd, gp.slot_lo, gp.slot_hi)) return;
    PamGate gate;
    bc_gate_init(gate, gp.P, gp.L, gp.pam_dir, gp.pam_sets);
    for (uint32_t pos = gp.pos_begin + blockIdx.x * blockDim.x + threadIdx.x; pos < gp.pos_end;
         pos += gridDim.x * blockDim.x) {
        if (bc_window(gp.B, pos) & lm) continue;
        if (gp.gate_first && !bc_gate_window(gate, gp.H, gp.Lo, gp.B, pos)) continue;
        const uint32_t wh = bc_window(gp.H, pos) & lm, wl = bc_window(gp.Lo, pos) & lm;
        const uint32_t slot = cd.dir_off + bc_combo_key(cd, wh, wl);
        if (slot < gp.slot_lo || slot >= gp.slot_hi) continue;
        if (gp.prune && gp.lib_dir[slot] == gp.lib_dir[slot + 1]) continue;
        if (PASS == 0) {
            atomicAdd(&gdir_or_cursor[slot], 1u);
        } else {
            const uint32_t dst = atomicAdd(&gdir_or_cursor[slot], 1u);
            gwin[dst] = make_uint4(pos, wh, wl, slot);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Radix form of the window scatter (BC_PARAM_WINDOW_SORT = 2, chosen automatically for dense
// directories).  The direct scatter above pays one returning global atomic and one isolated
// 16-byte store per record (22 ps per record at cfg 4, bound by the L2's atomic + partial-sector
// store rate, LSU queue full).  Here the exact slot offsets still come from the count pass, but
// the records reach their slots in two passes that sort RB_CHUNK records at a time in shared
// memory, so global atomics drop to one per (chunk, bin) and the stores leave as runs:
//   pass A  k_window_bin    windows -> records grouped by BIN = slot >> 8 (the top key bits), each
//                           bin region at its final place in a scratch array;
//   pass B  k_window_place  bin regions -> final slots (low 8 key bits), cursor per slot.
// Needs every combination's key to be at least 4 nt (directory offsets are multiples of 256 then,
// so `slot >> 8` is a global bin index) and at most 8 nt (<= 256 bins per combination).
#define RB_THREADS 256
#ifndef RB_ITEMS
#define RB_ITEMS 8
#endif
#define RB_CHUNK (RB_THREADS * RB_ITEMS)
#ifndef RB_MINBLOCKS
#define RB_MINBLOCKS 6
#endif

// exclusive scan of one value per thread over the CTA (256 threads); s_warp: 8 words of scratch
__device__ __forceinline__ uint32_t rb_block_scan(uint32_t v, uint32_t* s_warp) {
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
    for (int w = 0; w < RB_THREADS / 32; w++) before += (uint32_t)w < warp ? s_warp[w] : 0u;
    return before + incl - v;
}

__global__ void k_bin_init(const uint32_t* __restrict__ gdir, uint32_t* __restrict__ bin_cursor, uint32_t n_bins) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_bins) bin_cursor[g] = gdir[(size_t)g << 8];
}

__global__ void __launch_bounds__(RB_THREADS, RB_MINBLOCKS) k_window_bin(const __grid_constant__ BucketParams gp,
                                                                        uint32_t* __restrict__ bin_cursor,
                                                                        uint4* __restrict__ tmp) {
    __shared__ uint32_t s_hist[256], s_lstart[256], s_gbase[256], s_warp[RB_THREADS / 32];
    __shared__ uint4 s_rec[RB_CHUNK];
    const uint32_t lm = bc_lmask(gp.L);
    PamGate gate;
    bc_gate_init(gate, gp.P, gp.L, gp.pam_dir, gp.pam_sets);
    const uint32_t tid = threadIdx.x;
    // position-major: the windows of a chunk are read and validated once and then binned for
    // every combination in turn (they were 40 % of this kernel's instructions when every
    // combination re-read them)
    for (uint64_t c0 = (uint64_t)gp.pos_begin + (uint64_t)blockIdx.x * RB_CHUNK; c0 < gp.pos_end;
         c0 += (uint64_t)gridDim.x * RB_CHUNK) {
        uint32_t wh[RB_ITEMS], wl[RB_ITEMS], ok = 0;
#pragma unroll
        for (int i = 0; i < RB_ITEMS; i++) {
            const uint64_t pos64 = c0 + tid + (uint32_t)i * RB_THREADS;
            wh[i] = wl[i] = 0;
            if (pos64 >= gp.pos_end) continue;
            const uint32_t pos = (uint32_t)pos64;
            if (bc_window(gp.B, pos) & lm) continue;
            if (gp.gate_first && !bc_gate_window(gate, gp.H, gp.Lo, gp.B, pos)) continue;
            wh[i] = bc_window(gp.H, pos) & lm;
            wl[i] = bc_window(gp.Lo, pos) & lm;
            ok |= 1u << i;
        }
        for (uint32_t c = 0; c < gp.n_combos; c++) {
            const ComboDesc& cd = gp.combo[c];
            if (!bc_combo_in_range(cd, gp.slot_lo, gp.slot_hi)) continue;  // block-uniform
            const uint32_t bin0 = cd.dir_off >> 8;
            s_hist[tid] = 0;
            __syncthreads();
            uint32_t slot[RB_ITEMS], rank[RB_ITEMS];
#pragma unroll
            for (int i = 0; i < RB_ITEMS; i++) {
                slot[i] = 0xffffffffu;
                if (!((ok >> i) & 1u)) continue;
                const uint32_t sl = cd.dir_off + bc_combo_key(cd, wh[i], wl[i]);
                if (sl < gp.slot_lo || sl >= gp.slot_hi) continue;
                if (gp.prune && gp.lib_dir[sl] == gp.lib_dir[sl + 1]) continue;
                slot[i] = sl;
                rank[i] = atomicAdd(&s_hist[(sl >> 8) - bin0], 1u);
            }
            __syncthreads();
            const uint32_t cnt = s_hist[tid];
            s_lstart[tid] = rb_block_scan(cnt, s_warp);
            if (cnt) s_gbase[tid] = atomicAdd(&bin_cursor[bin0 + tid], cnt);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < RB_ITEMS; i++) {
                if (slot[i] != 0xffffffffu)
                    s_rec[s_lstart[(slot[i] >> 8) - bin0] + rank[i]] =
                        make_uint4((uint32_t)(c0 + tid + (uint32_t)i * RB_THREADS), wh[i], wl[i], slot[i]);
            }
            __syncthreads();
            const uint32_t total = s_lstart[255] + s_hist[255];
            for (uint32_t i = tid; i < total; i += RB_THREADS) {
                const uint4 r = s_rec[i];
                const uint32_t bl = (r.w >> 8) - bin0;
                tmp[s_gbase[bl] + (i - s_lstart[bl])] = r;
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(RB_THREADS, RB_MINBLOCKS) k_window_place(const uint4* __restrict__ tmp,
                                                            const uint32_t* __restrict__ gdir,
                                                            uint32_t* __restrict__ gcursor, uint4* __restrict__ gwin,
                                                            const uint32_t* __restrict__ n_rec_ptr,
                                                            uint32_t* __restrict__ work) {
    __shared__ uint32_t s_hist[256], s_lstart[256], s_gbase[256], s_warp[RB_THREADS / 32];
    __shared__ uint4 s_rec[RB_CHUNK];
    __shared__ uint32_t s_chunk;
    const uint32_t tid = threadIdx.x;
    const uint32_t n_rec = *n_rec_ptr;
    const uint32_t n_chunks = (n_rec + RB_CHUNK - 1) / RB_CHUNK;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_chunk = atomicAdd(work, 1u);
        __syncthreads();
        const uint32_t ch = s_chunk;
        if (ch >= n_chunks) break;
        const uint32_t r0 = ch * RB_CHUNK, r1 = r0 + min((uint32_t)RB_CHUNK, n_rec - r0);
        // a chunk of the scratch array may straddle bin regions: one shared-memory sort per piece
        uint32_t seg = r0;
        while (seg < r1) {
            const uint32_t g = __ldg(&tmp[seg].w) >> 8;
            const uint32_t s1 = min(r1, __ldg(gdir + (((size_t)g + 1) << 8)));
            s_hist[tid] = 0;
            __syncthreads();
            uint4 rec[RB_ITEMS];
            uint32_t rank[RB_ITEMS];
#pragma unroll
            for (int i = 0; i < RB_ITEMS; i++) {
                const uint32_t idx = seg + tid + (uint32_t)i * RB_THREADS;
                if (idx < s1) {
                    rec[i] = __ldcs(tmp + idx);
                    rank[i] = atomicAdd(&s_hist[rec[i].w & 255u], 1u);
                }
            }
            __syncthreads();
            const uint32_t cnt = s_hist[tid];
            s_lstart[tid] = rb_block_scan(cnt, s_warp);
            if (cnt) s_gbase[tid] = atomicAdd(&gcursor[((size_t)g << 8) + tid], cnt);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < RB_ITEMS; i++) {
                const uint32_t idx = seg + tid + (uint32_t)i * RB_THREADS;
                if (idx < s1) s_rec[s_lstart[rec[i].w & 255u] + rank[i]] = rec[i];
            }
            __syncthreads();
            const uint32_t n = s1 - seg;
            for (uint32_t i = tid; i < n; i += RB_THREADS) {
                const uint4 r = s_rec[i];
                const uint32_t sub = r.w & 255u;
                gwin[s_gbase[sub] + (i - s_lstart[sub])] = r;
            }
            __syncthreads();
            seg = s1;
        }
    }
}

#define MV_THREADS 256
#define MV_WARPS (MV_THREADS / 32)
#ifndef MV_ITEMS
#define MV_ITEMS 4    // records per lane per warp-tile
#endif
#define MV_WQ 128     // per-warp candidate queue (entries)
#define MV_GQ 96      // per-warp queue of dense groups awaiting re-examination (31 + 2 * 32)
#ifndef MV_DENSE_MIN
#define MV_DENSE_MIN 48     // window records a slot needs to be walked by the dense kernel
#endif
#ifndef MV_CHUNK_TILES
#define MV_CHUNK_TILES 16   // warp-tiles per work chunk of the dense kernel
#endif
#ifndef MV_ALU_PAIRS
#define MV_ALU_PAIRS 4      // of the MV_DENSE_ENTRIES x MV_ITEMS pairs of a group, how many are tested on the ALU pipe instead of POPC
#endif
#ifndef MV_STAGE
#define MV_STAGE 64         // library entries per shared-memory stage of the dense kernel (x2 buffers per warp)
#endif

__device__ __forceinline__ void mv_cp_async8(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
#ifndef MV_DENSE_ENTRIES
#define MV_DENSE_ENTRIES 8    // index entries per group in the dense kernel (one ballot per group)
#endif
#ifndef MV_DENSE_MINBLOCKS
#define MV_DENSE_MINBLOCKS 4  // occupancy target of the dense kernel (CTAs per SM): 32 warps (64 registers) beat 24 and 16
#endif

static_assert(MV_DENSE_ENTRIES % 2 == 0 && MV_STAGE % MV_DENSE_ENTRIES == 0,
              "the dense kernel reads the staged bucket two entries (16 bytes) at a time");

__device__ __forceinline__ uint32_t mv_combo_of_slot(const SearchParams& p, uint32_t slot) {
    uint32_t c = 0;
    while (c + 1 < p.n_combos && p.combo[c + 1].dir_off <= slot) c++;
    return c;
}

// Resolve up to 32 queued candidates {dev position, mismatch mask, index entry, directory slot}
// with all lanes of the warp: ownership, PAM annotation, then ONE global atomic for the whole
// batch and a coalesced store of the surviving records.
static __device__ __noinline__ void mv_resolve(const SearchParams& p, const uint4* q, uint32_t n) {
    const uint32_t lane = threadIdx.x & 31u;
    uint4 rec;
    bool ok = false;
    if (lane < n) {
        const uint4 qe = q[lane];
        ok = bc_make_hit(p, mv_combo_of_slot(p, qe.w), qe.x, p.ent_id[qe.z], qe.y, &rec);
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

// Queue full (pathological hit density): resolve one candidate on the spot.
static __device__ __noinline__ void mv_overflow(const SearchParams& p, uint32_t pos, uint32_t m, uint32_t e,
                                                uint32_t slot) {
    uint4 rec;
    if (bc_make_hit(p, mv_combo_of_slot(p, slot), pos, p.ent_id[e], m, &rec)) {
        const unsigned long long g = atomicAdd(p.count, 1ull);
        if (g < p.cap) reinterpret_cast<uint4*>(p.hits)[g] = rec;
    }
}

// Verify kernels, warp-autonomous (no block barriers): each warp owns warp-tiles of 32*MV_ITEMS
// sorted records.  Candidates that pass the popcount filter are NOT resolved where they are found
// - a single lane walking the slow path (dependent loads of the entry id, contig table and PAM
// bases) would stall its warp for microseconds, and ncu showed 11.6 active lanes per instruction
// when it did - they are pushed to a per-warp shared-memory queue and resolved 32 at a time.
//
// Two instantiations share the record stream (keeps registers and code size down):
//   DENSE = true   warp-tiles that lie entirely in ONE directory slot (cfg 4 at b=5: ~300 entries
//                  x ~1500 windows per slot): the bucket is walked once, 4 entries x MV_ITEMS
//                  windows = 16 independent LOP3/POPC chains per group, one broadcast load per
//                  MV_ITEMS candidates;
//   DENSE = false  all other warp-tiles: every lane walks the bucket of each of its records.
#define MV_CANDIDATE(E, Q)                                                                      \
    do {                                                                                        \
        const uint32_t m_ = (w.y ^ (Q).x) | (w.z ^ (Q).y);                                      \
        const uint32_t qs = atomicAdd(qn, 1u);                                                  \
        if (qs < MV_WQ) q[qs] = make_uint4(w.x, m_, (E), w.w);                                  \
        else mv_overflow(p, w.x, m_, (E), w.w);                                                 \
    } while (0)

// resolve full batches of 32 queued candidates; warp-uniform, called at warp-converged points
__device__ __forceinline__ void mv_drain(const SearchParams& p, uint4* q, uint32_t* qn, uint32_t lane) {
    __syncwarp();
    uint32_t nq = min(*qn, (uint32_t)MV_WQ);
    if (nq >= 32) {
        do {
            mv_resolve(p, q + (nq - 32), 32);
            nq -= 32;
        } while (nq >= 32);
        __syncwarp();
        if (lane == 0) *qn = nq;
    }
    __syncwarp();
}

// Dense kernel, second level: a queued GROUP = {first record of the lane that saw it, first entry
// of the group} stands for MV_ITEMS x MV_DENSE_ENTRIES pairs of which at least one passed the
// filter.  32 groups are re-examined at once, one per lane, so the per-pair compare+branch
// sequence runs with full lanes instead of the 1-2 lanes that found something (that sequence was
// ~35 % of the dense kernel's issued instructions when it ran where the group was found).
// Records past the end of the lane's slot (ragged last tile of a slot) and entries past the end
// of the bucket (ragged last group) are skipped here.
static __device__ __noinline__ void mv_resolve_groups(const SearchParams& p, const uint4* __restrict__ gwin,
                                                      const uint32_t* __restrict__ gdir, const uint2* gq, uint32_t n,
                                                      uint4* q, uint32_t* qn) {
    const uint32_t lane = threadIdx.x & 31u;
    const int k = (int)p.k;
    __syncwarp();
    if (lane < n) {
        const uint2 item = gq[lane];
        const uint32_t slot = __ldg(&gwin[item.x].w);
        const uint32_t slot_end = __ldg(gdir + slot + 1);
        const uint32_t n_e = min((uint32_t)MV_DENSE_ENTRIES, __ldg(p.dir + slot + 1) - item.y);
        uint4 w[MV_ITEMS];
#pragma unroll
        for (int it = 0; it < MV_ITEMS; it++) w[it] = __ldg(gwin + min(item.x + it * 32, slot_end - 1));
        for (uint32_t j = 0; j < n_e; j++) {
            const uint2 qe = __ldg(p.ent_hl + item.y + j);
#pragma unroll
            for (int it = 0; it < MV_ITEMS; it++) {
                if (item.x + it * 32 < slot_end && __popc((w[it].y ^ qe.x) | (w[it].z ^ qe.y)) <= k) {
                    const uint32_t m_ = (w[it].y ^ qe.x) | (w[it].z ^ qe.y);
                    const uint32_t qs = atomicAdd(qn, 1u);
                    if (qs < MV_WQ) q[qs] = make_uint4(w[it].x, m_, item.y + j, slot);
                    else mv_overflow(p, w[it].x, m_, item.y + j, slot);
                }
            }
        }
    }
    mv_drain(p, q, qn, lane);
}

// Dense verify: every slot with at least MV_DENSE_MIN window records is cut into warp-tiles of
// 32*MV_ITEMS records ALIGNED TO THE SLOT START, so all windows of a tile share one library bucket
// and the bucket is streamed once per tile as warp-uniform (broadcast) loads against MV_ITEMS
// resident windows per lane.  Only the last tile of a slot is ragged (its missing windows are
// copies that the second level skips); with fixed tiles over the record array every tile that
// crossed a slot boundary (8 % of cfg 4's) had to take the lane-per-record kernel, which is ~3x
// slower per pair.  Work distribution: the record array is cut into chunks of MV_CHUNK_TILES
// tiles dealt round-robin to the warps; a warp owns the slot-aligned tiles that START in its
// chunk and finds them from the first record of the chunk (its slot is in the record) and the
// record directory gdir, so no tile list has to be built.
template <int K>
__global__ void __launch_bounds__(MV_THREADS, MV_DENSE_MINBLOCKS) k_verify_dense(const __grid_constant__ SearchParams p,
                                                                                 const uint4* __restrict__ gwin,
                                                                                 const uint32_t* __restrict__ gdir,
                                                                                 const uint32_t* __restrict__ n_rec_ptr,
                                                                                 uint32_t* __restrict__ work, uint32_t slice,
                                                                                 uint32_t frac_lo, uint32_t frac_hi) {
    __shared__ uint4 s_q[MV_WARPS][MV_WQ];
    __shared__ uint2 s_gq[MV_WARPS][MV_GQ];
    __shared__ __align__(16) uint2 s_ent[MV_WARPS][2 * MV_STAGE];
    __shared__ uint32_t s_qn[MV_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint2* gq = s_gq[warp];
    uint2* sbuf = s_ent[warp];
    uint32_t gn = 0;  // queued groups of this warp (warp-uniform; survives across tiles)
    uint4* q = s_q[warp];
    uint32_t* qn = &s_qn[warp];
    if (lane == 0) *qn = 0;
    __syncwarp();
    const uint32_t n_rec = *n_rec_ptr;
    const uint32_t wtile = 32 * MV_ITEMS, chunk = wtile * MV_CHUNK_TILES;
    const uint32_t n_chunks = (n_rec + chunk - 1) / chunk;
    const int k = K;  // == p.k (host dispatch)
    const uint2* __restrict__ ent = p.ent_hl;
    unsigned long long cand = 0;
    // a launch handles the chunks [frac_lo, frac_hi) / 65536 of the record array (streamed result delivery)
    const uint32_t ch_lo = (uint32_t)(((unsigned long long)n_chunks * frac_lo) >> 16);
    const uint32_t ch_hi = (uint32_t)(((unsigned long long)n_chunks * frac_hi) >> 16);
    // Chunks are handed out through an atomic counter: with a static deal the warps the scheduler
    // favours (XU arbitration is by warp id) finished at ~60 % of the kernel and the rest could not
    // keep the POPC pipe full on their own (ncu: 24.5 of 32 warps resident on average).
    for (;;) {
        uint32_t ch = 0;
        if (lane == 0) ch = ch_lo + atomicAdd(work + slice, 1u);
        ch = __shfl_sync(0xffffffffu, ch, 0);
        if (ch >= ch_hi) break;
        const uint32_t r0 = ch * chunk, r1 = r0 + min(chunk, n_rec - r0);
        // all of these are warp-uniform (every lane loads the same words)
        uint32_t slot = __ldg(&gwin[r0].w);
        uint32_t a = __ldg(gdir + slot), b = __ldg(gdir + slot + 1);  // records of this slot
        uint32_t t = a + (r0 - a + wtile - 1) / wtile * wtile;         // first tile start >= r0
        for (;;) {
            if (t >= b || b - a < MV_DENSE_MIN) {  // slot finished, or left to the sparse kernel
                if (b >= r1) break;
                slot = __ldg(&gwin[b].w);  // next non-empty slot starts where this one ends
                a = b;
                b = __ldg(gdir + slot + 1);
                t = a;
                continue;
            }
            if (t >= r1) break;
            const uint32_t first = t;
            t += wtile;
            const uint32_t ls = __ldg(p.dir + slot), le = __ldg(p.dir + slot + 1);
            if (ls == le) continue;
            // lanes without a record of their own are switched off (threshold -1); the missing
            // higher items of a live lane are copies of its first window
            const bool live = first + lane < b;
            const int kl = live ? k : -1;
            uint4 wv[MV_ITEMS];
            wv[0] = __ldcs(gwin + (live ? first + lane : b - 1));
            uint32_t mine = live ? 1u : 0u;
#pragma unroll
            for (int it = 1; it < MV_ITEMS; it++) {
                const bool ok = first + it * 32 + lane < b;
                wv[it] = wv[0];
                if (ok) wv[it] = __ldcs(gwin + first + it * 32 + lane);
                mine += ok ? 1u : 0u;
            }
            cand += (unsigned long long)(le - ls) * mine;
            // The bucket is staged through shared memory in chunks of MV_STAGE entries, double
            // buffered with cp.async one whole chunk ahead (ncu: the bucket words came from L2/DRAM
            // more often than from L1 - global-load L1 hit rate 63 %, no reuse beyond the sector).
            // One group = MV_DENSE_ENTRIES entries (16-byte broadcast LDS, two entries each) x
            // MV_ITEMS windows = 32 independent LOP3/LOP3/POPC chains folded with 3-input integer
            // min, then ONE ballot: a lane whose minimum passes only QUEUES the group (one shared
            // store); mv_resolve_groups re-examines 32 groups at a time.  The last group of a
            // bucket may read stale entries of the stage buffer; the second level bounds them.
            const uint32_t n_ent = le - ls;
            const uint32_t n_stage = (n_ent + MV_STAGE - 1) / MV_STAGE;
            const uint2* bucket = ent + ls;
#define MV_ISSUE(C)                                                                            \
    do {                                                                                       \
        uint2* dst_ = sbuf + ((C) & 1u) * MV_STAGE;                                            \
        _Pragma("unroll") for (int h = 0; h < MV_STAGE / 32; h++) {                            \
            const uint32_t i_ = (C) * MV_STAGE + h * 32 + lane;                                \
            if (i_ < n_ent) mv_cp_async8(dst_ + h * 32 + lane, bucket + i_);                   \
        }                                                                                      \
        asm volatile("cp.async.commit_group;" ::: "memory");                                   \
    } while (0)
            MV_ISSUE(0u);
            for (uint32_t c = 0; c < n_stage; c++) {
                if (c + 1 < n_stage) {
                    MV_ISSUE(c + 1);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                __syncwarp();
                const uint4* sb = reinterpret_cast<const uint4*>(sbuf + (c & 1u) * MV_STAGE);
                const uint32_t ng = (min((uint32_t)MV_STAGE, n_ent - c * MV_STAGE) + MV_DENSE_ENTRIES - 1) / MV_DENSE_ENTRIES;
                const uint32_t ebase = ls + c * MV_STAGE;
                // The resolve calls sit OUTSIDE the group loop (a call inside it made ptxas spill
                // the window registers and reload them every iteration): the loop leaves when a
                // full batch is queued and is re-entered afterwards.
                uint32_t g = 0;
                for (;;) {
                    for (; g < ng && gn < 32; g++) {
                        int best_ = 33;
                        uint32_t rest_ = 0xffffffffu;  // min over the ALU-tested pairs of (mismatch mask with its K lowest bits cleared)
#pragma unroll
                        for (int j = 0; j < MV_DENSE_ENTRIES / 2; j++) {
                            const uint4 e2 = sb[g * (MV_DENSE_ENTRIES / 2) + j];
#pragma unroll
                            for (int it = 0; it < MV_ITEMS; it++) {
                                best_ = min(best_, __popc((wv[it].y ^ e2.x) | (wv[it].z ^ e2.y)));
                                if ((MV_DENSE_ENTRIES / 2 - 1 - j) * MV_ITEMS + it < MV_ALU_PAIRS) {
                                    // popc(m) <= K  <=>  m with its K lowest set bits cleared is 0: the same
                                    // test on the ALU pipe, which has head-room while POPC saturates the XU pipe
                                    uint32_t m = (wv[it].y ^ e2.z) | (wv[it].z ^ e2.w);
#pragma unroll
                                    for (int c = 0; c < K; c++) m &= m - 1u;
                                    rest_ = min(rest_, m);
                                } else {
                                    best_ = min(best_, __popc((wv[it].y ^ e2.z) | (wv[it].z ^ e2.w)));
                                }
                            }
                        }
                        const bool pass_ = best_ <= kl || (MV_ALU_PAIRS && kl >= 0 && rest_ == 0u);
                        const uint32_t hit_ = __ballot_sync(0xffffffffu, pass_);
                        if (hit_) {  // warp-uniform
                            if (pass_)
                                gq[gn + __popc(hit_ & lt_mask)] = make_uint2(first + lane, ebase + g * MV_DENSE_ENTRIES);
                            gn += __popc(hit_);
                        }
                    }
                    if (gn < 32) break;
                    do {  // warp-uniform; at most 31 + 32 groups are queued here
                        gn -= 32;
                        mv_resolve_groups(p, gwin, gdir, gq + gn, 32, q, qn);
                    } while (gn >= 32);
                }
                __syncwarp();  // every lane is done with this buffer before chunk c+2 lands in it
            }
#undef MV_ISSUE
            mv_drain(p, q, qn, lane);
        }
    }
    while (gn) {  // up to MV_GQ - 1 groups are still queued
        const uint32_t take = min(gn, 32u);
        gn -= take;
        mv_resolve_groups(p, gwin, gdir, gq + gn, take, q, qn);
    }
    __syncwarp();
    const uint32_t nq = min(*qn, (uint32_t)MV_WQ);
    if (nq) mv_resolve(p, q, nq);
    if (p.count_candidates) atomicAdd(p.count + 1, cand);
}

// Sparse verify: the records of slots with fewer than MV_DENSE_MIN windows, one record per lane,
// every lane walking the library bucket of its own record.
__global__ void __launch_bounds__(MV_THREADS, 3) k_verify_sparse(const __grid_constant__ SearchParams p,
                                                                 const uint4* __restrict__ gwin,
                                                                 const uint32_t* __restrict__ gdir,
                                                                 const uint32_t* __restrict__ n_rec_ptr) {
    __shared__ uint4 s_q[MV_WARPS][MV_WQ];
    __shared__ uint32_t s_qn[MV_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint4* q = s_q[warp];
    uint32_t* qn = &s_qn[warp];
    if (lane == 0) *qn = 0;
    __syncwarp();
    const uint32_t n_rec = *n_rec_ptr;
    const uint32_t wtile = 32 * MV_ITEMS;
    const uint32_t n_wtiles = (n_rec + wtile - 1) / wtile;
    const uint32_t n_warps = gridDim.x * MV_WARPS;
    const int k = (int)p.k, k1 = k + 1;
    const uint2* __restrict__ ent = p.ent_hl;
    unsigned long long cand = 0;
    for (uint32_t wt = blockIdx.x * MV_WARPS + warp; wt < n_wtiles; wt += n_warps) {
        // records are sorted by slot: a full tile whose first and last slots agree lies inside a
        // slot of >= 32*MV_ITEMS >= MV_DENSE_MIN records, which the dense kernel owns
        const uint32_t first = wt * wtile, last = first + min(wtile, n_rec - first) - 1;
        if (last - first + 1 == wtile && __ldg(&gwin[first].w) == __ldg(&gwin[last].w)) continue;
        uint4 wv[MV_ITEMS];
#pragma unroll
        for (int it = 0; it < MV_ITEMS; it++) wv[it] = __ldcs(gwin + min(first + it * 32 + lane, last));
        // Issue the directory loads of all MV_ITEMS records before any dependent work.
        uint32_t lsv[MV_ITEMS], lev[MV_ITEMS];
#pragma unroll
        for (int it = 0; it < MV_ITEMS; it++) {
            const bool mine = first + it * 32 + lane <= last &&
                              __ldg(gdir + wv[it].w + 1) - __ldg(gdir + wv[it].w) < MV_DENSE_MIN;
            lsv[it] = __ldg(p.dir + wv[it].w);
            lev[it] = mine ? __ldg(p.dir + wv[it].w + 1) : lsv[it];
        }
#pragma unroll
        for (int it = 0; it < MV_ITEMS; it++) {
            const uint4 w = wv[it];
            const uint32_t ls = lsv[it], le = lev[it];
            cand += le - ls;
            uint32_t e = ls;
            // branch-free batches of 4: the sign bits of (count - (k+1)) are OR-ed, only a batch
            // containing a candidate is re-examined
            for (; e + 4 <= le; e += 4) {
                const uint2 q0 = __ldg(ent + e), q1 = __ldg(ent + e + 1), q2 = __ldg(ent + e + 2),
                            q3 = __ldg(ent + e + 3);
                const int c0 = __popc((w.y ^ q0.x) | (w.z ^ q0.y));
                const int c1 = __popc((w.y ^ q1.x) | (w.z ^ q1.y));
                const int c2 = __popc((w.y ^ q2.x) | (w.z ^ q2.y));
                const int c3 = __popc((w.y ^ q3.x) | (w.z ^ q3.y));
                if (((c0 - k1) | (c1 - k1) | (c2 - k1) | (c3 - k1)) < 0) {
                    if (c0 <= k) MV_CANDIDATE(e, q0);
                    if (c1 <= k) MV_CANDIDATE(e + 1, q1);
                    if (c2 <= k) MV_CANDIDATE(e + 2, q2);
                    if (c3 <= k) MV_CANDIDATE(e + 3, q3);
                }
            }
            for (; e < le; e++) {
                const uint2 qq = __ldg(ent + e);
                if (__popc((w.y ^ qq.x) | (w.z ^ qq.y)) <= k) MV_CANDIDATE(e, qq);
            }
            mv_drain(p, q, qn, lane);
        }
    }
    __syncwarp();
    const uint32_t nq = min(*qn, (uint32_t)MV_WQ);
    if (nq) mv_resolve(p, q, nq);
    if (p.count_candidates) atomicAdd(p.count + 1, cand);
}
#undef MV_CANDIDATE

// ------------------------------------------------------------------------------------------ host
void bc_join_free(JoinWorkspace& ws) {
    if (ws.d_gdir) cudaFree(ws.d_gdir);
    if (ws.d_gcursor) cudaFree(ws.d_gcursor);
    if (ws.d_gwin) cudaFree(ws.d_gwin);
    if (ws.d_gtmp) cudaFree(ws.d_gtmp);
    if (ws.d_work) cudaFree(ws.d_work);
    if (ws.d_bin_cursor) cudaFree(ws.d_bin_cursor);
    if (ws.d_scan_tmp) cudaFree(ws.d_scan_tmp);
    if (ws.d_lut) cudaFree(ws.d_lut);
    if (ws.d_items) cudaFree(ws.d_items);
    if (ws.d_tile_desc) cudaFree(ws.d_tile_desc);
    if (ws.d_tile_slot) cudaFree(ws.d_tile_slot);
    if (ws.d_tile_start) cudaFree(ws.d_tile_start);
    if (ws.ev_a) cudaEventDestroy(ws.ev_a);
    if (ws.ev_b) cudaEventDestroy(ws.ev_b);
    if (ws.ev_c) cudaEventDestroy(ws.ev_c);
    for (int i = 0; i < 6; i++)
        if (ws.ev_k[i]) cudaEventDestroy(ws.ev_k[i]);
    ws = JoinWorkspace();
}

#define JCK(call)                           \
    do {                                    \
        cudaError_t e__ = (call);           \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

// Streamed delivery after the verify slices of one pass have been launched: hits are appended
// through one atomic cursor, so everything below the counter value read after slice s is final
// once that slice has finished.
cudaError_t bc_sink_deliver(HitSink* sink, const SearchParams& p, uint32_t n_slices) {
    for (uint32_t s = 0; s < n_slices; s++) {
        JCK(cudaEventSynchronize(sink->ev[s]));
        uint64_t done = sink->h_counts[s];
        if (done > p.cap) done = p.cap;
        if (sink->fn && done >= sink->reported) {  // also called for an empty part: callers count calls
            sink->fn(sink->fn_user, p.hits, sink->reported, done);
            sink->reported = done;
        }
        if (!sink->host) continue;
        if (done > sink->cap) done = sink->cap;
        if (done > sink->copied) {
            JCK(cudaMemcpyAsync(sink->host + sink->copied, p.hits + sink->copied,
                                (done - sink->copied) * sizeof(bc_hit), cudaMemcpyDefault, sink->stream));
            sink->copied = done;
        }
    }
    return cudaSuccess;
}

cudaError_t bc_join_search(JoinWorkspace& ws, const SearchParams& p, uint64_t dir_slots, int sm_count,
                           cudaStream_t st, uint32_t* launches, HitSink* sink) {
    const uint32_t launches0 = bc_launch_counter;
    ws.ms_join_kernels = ws.ms_bucket_kernels = 0;
    if (!ws.ev_a) JCK(cudaEventCreate(&ws.ev_a));
    if (!ws.ev_b) JCK(cudaEventCreate(&ws.ev_b));
    if (!ws.ev_c) JCK(cudaEventCreate(&ws.ev_c));

    // chunk the genome so the window records stay within the workspace budget; the (slow)
    // free-memory query only runs when the cached workspace cannot hold the whole range
    // Radix scatter: possible when every key is 4..8 nt; worth it when the bins are well filled
    // (each shared-memory sort then emits long runs).
    bool radix_ok = true;
    for (uint32_t c = 0; c < p.n_combos; c++) radix_ok = radix_ok && p.combo[c].key_nt >= 4 && p.combo[c].key_nt <= 8;
    const uint64_t span0 = (uint64_t)p.pos_end - p.pos_begin;
    const bool radix = radix_ok &&
                       (p.window_sort == 2 ||
                        (p.window_sort == 0 && span0 * p.n_combos / ((dir_slots >> 8) + 1) >= 16ull * RB_CHUNK));
    const uint32_t n_arrays = radix ? 2 : 1;  // record arrays (scratch + final)
    const uint64_t span = (uint64_t)p.pos_end - p.pos_begin;
    uint64_t chunk = span ? span : 1;
    if (p.join_chunk && chunk > p.join_chunk) chunk = p.join_chunk;  // forced small passes (tests of the multi-pass branch)
    if (chunk * p.n_combos > ws.gwin_cap) {
        size_t free_b = 0, total_b = 0;
        JCK(cudaMemGetInfo(&free_b, &total_b));
        uint64_t budget = ((uint64_t)free_b + n_arrays * ws.gwin_cap * sizeof(uint4)) / 2;
        if (budget > (96ull << 30)) budget = 96ull << 30;
        const uint64_t fit = budget / (n_arrays * sizeof(uint4)) / p.n_combos;
        if (chunk > fit) chunk = fit;
        if (chunk < 1) chunk = 1;
    }
    if (chunk * p.n_combos >= (1ull << 32) - 65536) chunk = ((1ull << 32) - 65536) / p.n_combos;  // 32-bit record indices (+ tile slack)
    const uint64_t rec_needed = chunk * p.n_combos;
    if (rec_needed > ws.gwin_cap) {
        if (ws.d_gwin) cudaFree(ws.d_gwin);
        if (ws.d_gtmp) cudaFree(ws.d_gtmp);
        ws.d_gwin = ws.d_gtmp = nullptr;
        ws.gwin_cap = 0;
        JCK(cudaMalloc(&ws.d_gwin, (chunk * p.n_combos + 1) * sizeof(uint4)));
        ws.gwin_cap = chunk * p.n_combos;
    }
    if (radix && !ws.d_gtmp) JCK(cudaMalloc(&ws.d_gtmp, (ws.gwin_cap + 1) * sizeof(uint4)));
    if (radix && (dir_slots >> 8) + 1 > ws.bin_cap) {
        if (ws.d_bin_cursor) cudaFree(ws.d_bin_cursor);
        ws.d_bin_cursor = nullptr;
        ws.bin_cap = 0;
        JCK(cudaMalloc(&ws.d_bin_cursor, ((dir_slots >> 8) + 1) * sizeof(uint32_t)));
        ws.bin_cap = (dir_slots >> 8) + 1;
    }
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
    BucketParams gp;
    memset(&gp, 0, sizeof gp);
    gp.H = p.H; gp.Lo = p.Lo; gp.B = p.B;
    gp.lib_dir = p.dir;
    gp.L = p.L;
    gp.n_combos = p.n_combos;
    memcpy(gp.combo, p.combo, sizeof gp.combo);
    // Skipping windows whose library bucket is empty only pays when most buckets are empty.
    gp.prune = p.dir_entries < (dir_slots - 1) * 2 ? 1u : 0u;
    gp.gate_first = p.gate_first;
    gp.slot_lo = p.slot_lo; gp.slot_hi = p.slot_hi;
    gp.P = p.P; gp.pam_dir = p.pam_dir;
    for (int i = 0; i < 8; i++) gp.pam_sets[i] = p.pam_sets[i];

    for (uint64_t begin = p.pos_begin; begin < p.pos_end; begin += chunk) {
        gp.pos_begin = (uint32_t)begin;
        gp.pos_end = (uint32_t)((begin + chunk < p.pos_end) ? begin + chunk : p.pos_end);
        uint32_t npos = gp.pos_end - gp.pos_begin;
        uint32_t gx = (npos + 255) / 256;
        uint32_t maxb = (uint32_t)sm_count * 8u;
        if (gx > maxb) gx = maxb;
        dim3 grid(gx, p.n_combos);
        JCK(cudaMemsetAsync(ws.d_gdir, 0, dir_slots * 4, st));
        JCK(cudaEventRecord(ws.ev_c, st));
        k_bucket<0><<<grid, 256, 0, st>>>(gp, ws.d_gdir, nullptr);
        JCK(cudaGetLastError());
        JCK(bc_exclusive_scan(ws.d_gdir, dir_slots, ws.d_scan_tmp, st));
        JCK(cudaMemcpyAsync(ws.d_gcursor, ws.d_gdir, dir_slots * 4, cudaMemcpyDeviceToDevice, st));
        if (radix) {
            const uint32_t n_bins = (uint32_t)((dir_slots - 1) >> 8);
            k_bin_init<<<(n_bins + 255) / 256, 256, 0, st>>>(ws.d_gdir, ws.d_bin_cursor, n_bins);
            JCK(cudaGetLastError());
            uint32_t bx = (npos + RB_CHUNK - 1) / RB_CHUNK;
            if (bx > (uint32_t)sm_count * RB_MINBLOCKS) bx = (uint32_t)sm_count * RB_MINBLOCKS;
            k_window_bin<<<bx, RB_THREADS, 0, st>>>(gp, ws.d_bin_cursor, ws.d_gtmp);
            JCK(cudaGetLastError());
            JCK(cudaMemsetAsync(ws.d_work, 0, BC_SINK_SLICES * sizeof(uint32_t), st));
            k_window_place<<<(uint32_t)sm_count * RB_MINBLOCKS, RB_THREADS, 0, st>>>(ws.d_gtmp, ws.d_gdir, ws.d_gcursor, ws.d_gwin,
                                                                        ws.d_gdir + (dir_slots - 1), ws.d_work);
            JCK(cudaGetLastError());
            bc_launch_counter += 2;
        } else {
            k_bucket<1><<<grid, 256, 0, st>>>(gp, ws.d_gcursor, ws.d_gwin);
            JCK(cudaGetLastError());
        }
        JCK(cudaEventRecord(ws.ev_a, st));
        // the last directory slot is the end sentinel: after the scan it holds the record count
        const uint32_t* n_rec_ptr = ws.d_gdir + (dir_slots - 1);
        k_verify_sparse<<<(uint32_t)sm_count * 6u, MV_THREADS, 0, st>>>(p, ws.d_gwin, ws.d_gdir, n_rec_ptr);
        JCK(cudaGetLastError());
        // Streamed delivery: the slices halve (1/2, 1/4, ... and the last one repeated), so most
        // of the result is on its way early, only 1/2^(S-1) of it is still to be copied when the
        // last kernel ends, and there are few launch tails to pay for.
        const uint32_t n_slices = sink ? 5u : 1u;  // halving slices (<= BC_SINK_SLICES)
        JCK(cudaMemsetAsync(ws.d_work, 0, BC_SINK_SLICES * sizeof(uint32_t), st));
        for (uint32_t s = 0; s < n_slices; s++) {
            const uint32_t f_lo = 65536u - (65536u >> s), f_hi = s + 1 == n_slices ? 65536u : 65536u - (65536u >> (s + 1));
            const uint32_t dgrid = (uint32_t)sm_count * MV_DENSE_MINBLOCKS, lo = n_slices == 1 ? 0u : f_lo;
            switch (p.k) {
                case 0: k_verify_dense<0><<<dgrid, MV_THREADS, 0, st>>>(p, ws.d_gwin, ws.d_gdir, n_rec_ptr, ws.d_work, s, lo, f_hi); break;
                case 1: k_verify_dense<1><<<dgrid, MV_THREADS, 0, st>>>(p, ws.d_gwin, ws.d_gdir, n_rec_ptr, ws.d_work, s, lo, f_hi); break;
                case 2: k_verify_dense<2><<<dgrid, MV_THREADS, 0, st>>>(p, ws.d_gwin, ws.d_gdir, n_rec_ptr, ws.d_work, s, lo, f_hi); break;
                default: k_verify_dense<3><<<dgrid, MV_THREADS, 0, st>>>(p, ws.d_gwin, ws.d_gdir, n_rec_ptr, ws.d_work, s, lo, f_hi); break;
            }
            JCK(cudaGetLastError());
            if (sink) {
                JCK(cudaMemcpyAsync(sink->h_counts + s, p.count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                JCK(cudaEventRecord(sink->ev[s], st));
            }
        }
        JCK(cudaEventRecord(ws.ev_b, st));
        bc_launch_counter += n_slices - 1;
        if (sink) JCK(bc_sink_deliver(sink, p, n_slices));
        bc_launch_counter += 4;
        // events are reused per chunk, so read them before the next record
        JCK(cudaEventSynchronize(ws.ev_b));
        float ms = 0;
        JCK(cudaEventElapsedTime(&ms, ws.ev_a, ws.ev_b));
        ws.ms_join_kernels += ms;
        JCK(cudaEventElapsedTime(&ms, ws.ev_c, ws.ev_a));
        ws.ms_bucket_kernels += ms;
    }
    *launches = bc_launch_counter - launches0;
    return cudaSuccess;
}
