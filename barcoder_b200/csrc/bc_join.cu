// bc_join.cu - K3-join: the scan for libraries whose seed buckets are dense (cfg 3/4).
//
// The probe kernel walks a library bucket once per genome window at a random address, so with
// tens to hundreds of entries per bucket every window re-reads its bucket through L2/HBM.  Here
// the genome side is sorted by the same seed keys instead (a sort-merge join):
//
//   pass 1  k_bucket<0/1>   every genome window, for every seed combination c, becomes a 16-byte
//                           record {dev position, wh, wl, directory slot} placed in slot order:
//                           histogram -> exclusive scan -> scatter (counting sort, not stable).
//   pass 2  k_merge_verify  one thread per sorted record.  Neighbouring lanes hold windows of the
//                           same slot, so the library bucket [dir[slot], dir[slot+1]) is read as
//                           warp-uniform (broadcast) 8-byte loads that hit L1, and one candidate
//                           costs 2 LOP3 + POPC + a min-reduce; only batches that contain a hit
//                           take the slow path (ownership, PAM, staged record).
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
    uint32_t L, n_combos, prune;
    ComboDesc combo[BC_MAX_COMBOS];
};

// Pass 1.  PASS 0 counts, PASS 1 scatters.  Windows touching a non-ACGT base or a contig end
// are dropped here, so pass 2 never sees them.
template <int PASS>
__global__ void __launch_bounds__(256) k_bucket(const __grid_constant__ BucketParams gp,
                                                uint32_t* __restrict__ gdir_or_cursor, uint4* __restrict__ gwin) {
    const uint32_t lm = bc_lmask(gp.L);
    const ComboDesc& cd = gp.combo[blockIdx.y];
    for (uint32_t pos = gp.pos_begin + blockIdx.x * blockDim.x + threadIdx.x; pos < gp.pos_end;
         pos += gridDim.x * blockDim.x) {
        if (bc_window(gp.B, pos) & lm) continue;
        const uint32_t wh = bc_window(gp.H, pos) & lm, wl = bc_window(gp.Lo, pos) & lm;
        const uint32_t slot = cd.dir_off + bc_combo_key(cd, wh, wl);
        if (gp.prune && gp.lib_dir[slot] == gp.lib_dir[slot + 1]) continue;
        if (PASS == 0) {
            atomicAdd(&gdir_or_cursor[slot], 1u);
        } else {
            const uint32_t dst = atomicAdd(&gdir_or_cursor[slot], 1u);
            __stcs(&gwin[dst], make_uint4(pos, wh, wl, slot));  // streaming: keep the directory in L2
        }
    }
}

#define MV_THREADS 256
#define MV_ITEMS 4
#define MV_TILE (MV_THREADS * MV_ITEMS)
#define MV_TILES_PER_ROUND 4  // tiles verified between two resolve/flush phases
// Candidates that pass the popcount filter are not resolved where they are found: a single lane
// walking the slow path (dependent loads of the entry id, contig table and PAM bases) would stall
// its whole warp for microseconds.  Candidates that this combination owns are queued in shared
// memory as {dev position, mismatch mask, index entry, combination} and resolved once per round
// by all threads at once.
#define MV_QCAP 1024

__device__ __forceinline__ uint32_t mv_combo_of_slot(const SearchParams& p, uint32_t slot) {
    uint32_t c = 0;
    while (c + 1 < p.n_combos && p.combo[c + 1].dir_off <= slot) c++;
    return c;
}

__global__ void __launch_bounds__(MV_THREADS) k_merge_verify(const __grid_constant__ SearchParams p,
                                                             const uint4* __restrict__ gwin,
                                                             const uint32_t* __restrict__ n_rec_ptr) {
    __shared__ HitStage stage;
    __shared__ uint4 s_q[MV_QCAP];
    __shared__ uint32_t s_qn;
    if (threadIdx.x == 0) { stage.n = 0; s_qn = 0; }
    __syncthreads();
    const uint32_t n_rec = *n_rec_ptr;
    const uint32_t n_tiles = (n_rec + MV_TILE - 1) / MV_TILE;
    const uint32_t n_rounds = (n_tiles + MV_TILES_PER_ROUND - 1) / MV_TILES_PER_ROUND;
    const int k = (int)p.k;
    const uint2* __restrict__ ent = p.ent_hl;
    unsigned long long cand = 0;
    // entry E of the index is within k mismatches of window w
#define MV_CANDIDATE(E, Q)                                                            \
    do {                                                                              \
        const uint32_t m_ = (w.y ^ (Q).x) | (w.z ^ (Q).y);                            \
        const uint32_t c_ = mv_combo_of_slot(p, w.w);                                 \
        if (p.lib_has_n || bc_owns(p, c_, m_)) {                                      \
            const uint32_t qs = atomicAdd(&s_qn, 1u);                                 \
            if (qs < MV_QCAP) s_q[qs] = make_uint4(w.x, m_, (E), c_);                 \
            else {                                                                    \
                uint4 rec_;                                                           \
                if (bc_make_hit(p, c_, w.x, p.ent_id[(E)], m_, &rec_)) bc_stage_hit(p, &stage, rec_); \
            }                                                                         \
        }                                                                             \
    } while (0)
    for (uint32_t round = blockIdx.x; round < n_rounds; round += gridDim.x) {
#pragma unroll 1
        for (uint32_t tr = 0; tr < MV_TILES_PER_ROUND; tr++) {
            const uint32_t tile = round * MV_TILES_PER_ROUND + tr;
            if (tile >= n_tiles) break;
            // Issue the loads of all MV_ITEMS records, then of their directory entries, before
            // any dependent work: three memory latencies per tile instead of three per record.
            uint4 wv[MV_ITEMS];
            uint32_t lsv[MV_ITEMS], lev[MV_ITEMS];
#pragma unroll
            for (int it = 0; it < MV_ITEMS; it++) {
                const uint32_t i = tile * MV_TILE + it * MV_THREADS + threadIdx.x;
                wv[it] = __ldcs(gwin + min(i, n_rec - 1));
            }
#pragma unroll
            for (int it = 0; it < MV_ITEMS; it++) {
                const uint32_t i = tile * MV_TILE + it * MV_THREADS + threadIdx.x;
                lsv[it] = __ldg(p.dir + wv[it].w);
                lev[it] = i < n_rec ? __ldg(p.dir + wv[it].w + 1) : lsv[it];
            }
#pragma unroll
            for (int it = 0; it < MV_ITEMS; it++) {
                const uint4 w = wv[it];
                const uint32_t ls = lsv[it], le = lev[it];
                cand += le - ls;
                uint32_t e = ls;
                // branch-free batches of 4: popcounts are min-reduced, only a batch containing a
                // candidate is re-examined entry by entry
                for (; e + 4 <= le; e += 4) {
                    const uint2 q0 = __ldg(ent + e), q1 = __ldg(ent + e + 1), q2 = __ldg(ent + e + 2),
                                q3 = __ldg(ent + e + 3);
                    const int c0 = __popc((w.y ^ q0.x) | (w.z ^ q0.y));
                    const int c1 = __popc((w.y ^ q1.x) | (w.z ^ q1.y));
                    const int c2 = __popc((w.y ^ q2.x) | (w.z ^ q2.y));
                    const int c3 = __popc((w.y ^ q3.x) | (w.z ^ q3.y));
                    if (min(min(c0, c1), min(c2, c3)) <= k) {
                        if (c0 <= k) MV_CANDIDATE(e, q0);
                        if (c1 <= k) MV_CANDIDATE(e + 1, q1);
                        if (c2 <= k) MV_CANDIDATE(e + 2, q2);
                        if (c3 <= k) MV_CANDIDATE(e + 3, q3);
                    }
                }
                for (; e < le; e++) {
                    const uint2 q = __ldg(ent + e);
                    if (__popc((w.y ^ q.x) | (w.z ^ q.y)) <= k) MV_CANDIDATE(e, q);
                }
            }
        }
        __syncthreads();
        const uint32_t nq = min(s_qn, (uint32_t)MV_QCAP);
        for (uint32_t j = threadIdx.x; j < nq; j += MV_THREADS) {
            const uint4 qe = s_q[j];
            uint4 rec;
            if (bc_make_hit(p, qe.w, qe.x, p.ent_id[qe.z], qe.y, &rec)) bc_stage_hit(p, &stage, rec);
        }
        bc_flush_hits(p, &stage);  // barriers inside
        if (threadIdx.x == 0) s_qn = 0;
        __syncthreads();
    }
#undef MV_CANDIDATE
    if (p.count_candidates) atomicAdd(p.count + 1, cand);
}

// ------------------------------------------------------------------------------------------ host
bool bc_join_supported(const ComboDesc*, uint32_t n_combos, uint64_t) { return n_combos > 0; }

void bc_join_free(JoinWorkspace& ws) {
    if (ws.d_gdir) cudaFree(ws.d_gdir);
    if (ws.d_gcursor) cudaFree(ws.d_gcursor);
    if (ws.d_gwin) cudaFree(ws.d_gwin);
    if (ws.d_scan_tmp) cudaFree(ws.d_scan_tmp);
    if (ws.ev_a) cudaEventDestroy(ws.ev_a);
    if (ws.ev_b) cudaEventDestroy(ws.ev_b);
    if (ws.ev_c) cudaEventDestroy(ws.ev_c);
    ws = JoinWorkspace();
}

#define JCK(call)                           \
    do {                                    \
        cudaError_t e__ = (call);           \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

cudaError_t bc_join_search(JoinWorkspace& ws, const SearchParams& p, uint64_t dir_slots, int sm_count,
                           cudaStream_t st, uint32_t* launches) {
    const uint32_t launches0 = bc_launch_counter;
    ws.ms_join_kernels = ws.ms_bucket_kernels = 0;
    if (!ws.ev_a) JCK(cudaEventCreate(&ws.ev_a));
    if (!ws.ev_b) JCK(cudaEventCreate(&ws.ev_b));
    if (!ws.ev_c) JCK(cudaEventCreate(&ws.ev_c));

    // chunk the genome so the window records stay within the workspace budget
    size_t free_b = 0, total_b = 0;
    JCK(cudaMemGetInfo(&free_b, &total_b));
    uint64_t budget = ((uint64_t)free_b + ws.gwin_cap * sizeof(uint4)) / 2;
    if (budget > (64ull << 30)) budget = 64ull << 30;
    uint64_t chunk = budget / sizeof(uint4) / p.n_combos;
    if (chunk > p.n_pos) chunk = p.n_pos;
    if (chunk < 1) chunk = 1;
    if (chunk * p.n_combos >= (1ull << 32)) chunk = ((1ull << 32) - 1) / p.n_combos;  // 32-bit record indices
    const uint64_t rec_needed = chunk * p.n_combos;
    if (rec_needed > ws.gwin_cap) {
        if (ws.d_gwin) cudaFree(ws.d_gwin);
        ws.d_gwin = nullptr;
        ws.gwin_cap = 0;
        JCK(cudaMalloc(&ws.d_gwin, (chunk * p.n_combos + 1) * sizeof(uint4)));
        ws.gwin_cap = chunk * p.n_combos;
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
    BucketParams gp;
    memset(&gp, 0, sizeof gp);
    gp.H = p.H; gp.Lo = p.Lo; gp.B = p.B;
    gp.lib_dir = p.dir;
    gp.L = p.L;
    gp.n_combos = p.n_combos;
    memcpy(gp.combo, p.combo, sizeof gp.combo);
    // Skipping windows whose library bucket is empty only pays when most buckets are empty.
    gp.prune = p.dir_entries < (dir_slots - 1) * 2 ? 1u : 0u;

    for (uint64_t begin = 0; begin < p.n_pos; begin += chunk) {
        gp.pos_begin = (uint32_t)begin;
        gp.pos_end = (uint32_t)((begin + chunk < p.n_pos) ? begin + chunk : p.n_pos);
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
        k_bucket<1><<<grid, 256, 0, st>>>(gp, ws.d_gcursor, ws.d_gwin);
        JCK(cudaGetLastError());
        JCK(cudaEventRecord(ws.ev_a, st));
        // the last directory slot is the end sentinel: after the scan it holds the record count
        k_merge_verify<<<(uint32_t)sm_count * 8u, MV_THREADS, 0, st>>>(p, ws.d_gwin, ws.d_gdir + (dir_slots - 1));
        JCK(cudaGetLastError());
        JCK(cudaEventRecord(ws.ev_b, st));
        bc_launch_counter += 3;
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
