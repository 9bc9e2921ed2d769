// bc_join.cu - K3-join: the scan for libraries whose seed buckets are dense (cfg 3/4).
//
// The probe kernel walks a library bucket once per genome window, so with tens to hundreds of
// entries per bucket every window re-reads its bucket through L2/HBM at random addresses.  Here
// the genome side is partitioned instead:
//
//   pass 1  k_part<0/1>   every genome window, for every seed combination c, is appended to the
//                         coarse partition (c, key >> d2[c]) as a 16-byte record
//                         {dev position, wh, wl, fine key}: histogram -> scan -> scatter.
//   pass 2  k_part_join   one CTA per partition: the partition's slice of the key-sorted library
//                         index (a few thousand {qh,ql} pairs) and its fine directory are staged in
//                         shared memory; the partition's windows stream through in chunks, each chunk
//                         is counting-sorted by fine key in shared memory, and then every thread
//                         verifies one window against its fine bucket: neighbouring lanes share the
//                         bucket, so the library words are shared-memory broadcasts and one candidate
//                         costs 2 LOP3 + POPC + compare.
//
// Algorithmic HBM traffic: one 16 B write + one 16 B read per (window, combination), the three
// genome planes twice per combination, the index once.
#include "bc_join.h"

#include <string.h>

#include "bc_kernels.h"

#define PJ_THREADS 512
#define PJ_CHUNK 2048        // windows sorted + verified per step (32 KB of shared memory)
#define PJ_PER_THREAD (PJ_CHUNK / PJ_THREADS)
#define PJ_LIB_CAP 4096      // library entries resident per partition (32 KB)
#define PJ_MAX_FINE_BITS 10  // fine buckets per partition <= 1024
#define PJ_MAX_FINE (1 << PJ_MAX_FINE_BITS)
#define PJ_LIB_TARGET 3072   // planned average library entries per partition

struct PartParams {
    const uint32_t* H;
    const uint32_t* Lo;
    const uint32_t* B;
    const uint32_t* lib_dir;      // library directory (to skip windows whose fine bucket is empty)
    uint32_t pos_begin, pos_end;  // dev positions handled by this chunk of the genome
    uint32_t L, n_combos, prune, d1;
    uint8_t d2[BC_MAX_COMBOS];
    ComboDesc combo[BC_MAX_COMBOS];
};

// Pass 1.  PASS 0 counts, PASS 1 scatters.  Windows touching a non-ACGT base or a contig end
// are dropped here, so pass 2 never sees them.
template <int PASS>
__global__ void __launch_bounds__(256) k_part(const __grid_constant__ PartParams gp,
                                              uint32_t* __restrict__ gdir_or_cursor, uint4* __restrict__ gwin) {
    const uint32_t lm = bc_lmask(gp.L);
    const uint32_t c = blockIdx.y;
    const ComboDesc& cd = gp.combo[c];
    const uint32_t d2 = gp.d2[c];
    for (uint32_t pos = gp.pos_begin + blockIdx.x * blockDim.x + threadIdx.x; pos < gp.pos_end;
         pos += gridDim.x * blockDim.x) {
        if (bc_window(gp.B, pos) & lm) continue;
        const uint32_t wh = bc_window(gp.H, pos) & lm, wl = bc_window(gp.Lo, pos) & lm;
        const uint32_t key = bc_combo_key(cd, wh, wl);
        if (gp.prune) {
            const uint32_t slot = cd.dir_off + key;
            if (gp.lib_dir[slot] == gp.lib_dir[slot + 1]) continue;
        }
        const uint32_t part = (c << gp.d1) | (key >> d2);
        if (PASS == 0) {
            atomicAdd(&gdir_or_cursor[part], 1u);
        } else {
            const uint32_t dst = atomicAdd(&gdir_or_cursor[part], 1u);
            gwin[dst] = make_uint4(pos, wh, wl, key & ((1u << d2) - 1u));
        }
    }
}

struct PartJoinParams {
    const uint32_t* gdir;  // [n_parts + 1] window ranges of the partitions
    const uint4* gwin;
    uint32_t n_parts, d1;
    uint8_t d2[BC_MAX_COMBOS];
};

// exclusive scan of s[0..n) in place, n <= 2 * PJ_THREADS, all threads of the CTA participate
__device__ __forceinline__ void pj_block_scan(uint32_t* s, uint32_t n, uint32_t* warp_sums) {
    const uint32_t t = threadIdx.x, lane = t & 31u, wid = t >> 5;
    const uint32_t i0 = 2 * t, i1 = 2 * t + 1;
    const uint32_t v0 = i0 < n ? s[i0] : 0, v1 = i1 < n ? s[i1] : 0;
    uint32_t x = v0 + v1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = lane < PJ_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < PJ_THREADS / 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= (uint32_t)o) ws += y;
        }
        if (lane < PJ_THREADS / 32) warp_sums[lane] = ws;
    }
    __syncthreads();
    const uint32_t base = (wid ? warp_sums[wid - 1] : 0) + x - (v0 + v1);
    if (i0 < n) s[i0] = base;
    if (i1 < n) s[i1] = base + v0;
    __syncthreads();
}

__global__ void __launch_bounds__(PJ_THREADS) k_part_join(const __grid_constant__ SearchParams p,
                                                          const __grid_constant__ PartJoinParams jp) {
    extern __shared__ __align__(16) unsigned char pj_smem[];
    uint4* s_win = reinterpret_cast<uint4*>(pj_smem);                      // PJ_CHUNK records
    uint2* s_lib = reinterpret_cast<uint2*>(s_win + PJ_CHUNK);             // PJ_LIB_CAP entries
    uint32_t* s_fdir = reinterpret_cast<uint32_t*>(s_lib + PJ_LIB_CAP);    // PJ_MAX_FINE + 1
    uint32_t* s_cnt = s_fdir + PJ_MAX_FINE + 1;                            // PJ_MAX_FINE
    uint32_t* s_warp = s_cnt + PJ_MAX_FINE;                                // PJ_THREADS / 32
    __shared__ HitStage stage;
    if (threadIdx.x == 0) stage.n = 0;
    const uint32_t tid = threadIdx.x;
    const int k = (int)p.k;
    unsigned long long cand = 0;

    for (uint32_t part = blockIdx.x; part < jp.n_parts; part += gridDim.x) {
        const uint32_t gs = jp.gdir[part], ge = jp.gdir[part + 1];
        if (gs == ge) continue;  // uniform across the CTA
        const uint32_t c = part >> jp.d1, kappa = part & ((1u << jp.d1) - 1u);
        const uint32_t d2 = jp.d2[c], nf = 1u << d2;
        const uint32_t s0 = p.combo[c].dir_off + (kappa << d2);
        const uint32_t ls = p.dir[s0], le = p.dir[s0 + nf];
        const uint32_t nl = le - ls;
        if (nl == 0) continue;
        __syncthreads();  // the previous partition is done with shared memory
        for (uint32_t f = tid; f <= nf; f += PJ_THREADS) s_fdir[f] = p.dir[s0 + f] - ls;
        const bool resident = nl <= PJ_LIB_CAP;
        if (resident)
            for (uint32_t i = tid; i < nl; i += PJ_THREADS) s_lib[i] = p.ent_hl[ls + i];

        for (uint32_t cb = gs; cb < ge; cb += PJ_CHUNK) {
            const uint32_t nw = min((uint32_t)PJ_CHUNK, ge - cb);
            for (uint32_t f = tid; f < nf; f += PJ_THREADS) s_cnt[f] = 0;
            __syncthreads();
            uint4 rec[PJ_PER_THREAD];
            uint32_t rank[PJ_PER_THREAD];
#pragma unroll
            for (int j = 0; j < PJ_PER_THREAD; j++) {
                const uint32_t i = j * PJ_THREADS + tid;
                if (i < nw) {
                    rec[j] = jp.gwin[cb + i];
                    rank[j] = atomicAdd(&s_cnt[rec[j].w], 1u);
                }
            }
            __syncthreads();
            pj_block_scan(s_cnt, nf, s_warp);
#pragma unroll
            for (int j = 0; j < PJ_PER_THREAD; j++) {
                const uint32_t i = j * PJ_THREADS + tid;
                if (i < nw) s_win[s_cnt[rec[j].w] + rank[j]] = rec[j];
            }
            __syncthreads();
            for (uint32_t lt = 0; lt < nl; lt += PJ_LIB_CAP) {
                const uint32_t tile_n = min((uint32_t)PJ_LIB_CAP, nl - lt);
                if (!resident) {
                    __syncthreads();
                    for (uint32_t i = tid; i < tile_n; i += PJ_THREADS) s_lib[i] = p.ent_hl[ls + lt + i];
                    __syncthreads();
                }
                for (uint32_t i = tid; i < nw; i += PJ_THREADS) {
                    const uint4 w = s_win[i];
                    const uint32_t a = max(s_fdir[w.w], lt), bnd = min(s_fdir[w.w + 1], lt + tile_n);
                    if (a >= bnd) continue;
                    cand += bnd - a;
                    const uint2* lib = s_lib + (a - lt);
                    const uint32_t n_e = bnd - a;
                    uint32_t e = 0;
                    // branch-free batches of 4: the popcounts are min-reduced and only a batch
                    // that contains a hit is re-examined entry by entry
                    for (; e + 4 <= n_e; e += 4) {
                        const uint2 q0 = lib[e], q1 = lib[e + 1], q2 = lib[e + 2], q3 = lib[e + 3];
                        const int c0 = __popc((w.y ^ q0.x) | (w.z ^ q0.y));
                        const int c1 = __popc((w.y ^ q1.x) | (w.z ^ q1.y));
                        const int c2 = __popc((w.y ^ q2.x) | (w.z ^ q2.y));
                        const int c3 = __popc((w.y ^ q3.x) | (w.z ^ q3.y));
                        if (min(min(c0, c1), min(c2, c3)) <= k) {
                            for (uint32_t j = e; j < e + 4; j++) {
                                const uint2 q = lib[j];
                                const uint32_t m = (w.y ^ q.x) | (w.z ^ q.y);
                                uint4 rec;
                                if (__popc(m) <= k && bc_make_hit(p, c, w.x, p.ent_id[ls + a + j], m, &rec))
                                    bc_stage_hit(p, &stage, rec);
                            }
                        }
                    }
                    for (; e < n_e; e++) {
                        const uint2 q = lib[e];
                        const uint32_t m = (w.y ^ q.x) | (w.z ^ q.y);
                        uint4 rec;
                        if (__popc(m) <= k && bc_make_hit(p, c, w.x, p.ent_id[ls + a + e], m, &rec))
                            bc_stage_hit(p, &stage, rec);
                    }
                }
            }
            bc_flush_hits(p, &stage);
            __syncthreads();  // before the next chunk reuses s_win / s_cnt
        }
    }
    if (p.count_candidates) atomicAdd(p.count + 1, cand);
}

// ------------------------------------------------------------------------------------------ host
static const size_t PJ_SMEM_BYTES = (size_t)PJ_CHUNK * sizeof(uint4) + (size_t)PJ_LIB_CAP * sizeof(uint2) +
                                    (size_t)(PJ_MAX_FINE + 1 + PJ_MAX_FINE + PJ_THREADS / 32) * sizeof(uint32_t);

// Coarse partition bits: partitions sized so that their slice of the library index fits the
// shared-memory tile, and no combination is left with more than 2^PJ_MAX_FINE_BITS fine buckets.
static bool plan_partition(const ComboDesc* combo, uint32_t n_combos, uint64_t entries_per_combo, uint32_t* d1_out,
                           uint8_t* d2_out) {
    if (n_combos == 0) return false;
    uint32_t kb_min = 64, kb_max = 0;
    for (uint32_t c = 0; c < n_combos; c++) {
        uint32_t kb = 2u * combo[c].key_nt;
        if (kb < kb_min) kb_min = kb;
        if (kb > kb_max) kb_max = kb;
    }
    uint32_t d1 = 0;
    while (d1 < 31 && (entries_per_combo >> d1) > PJ_LIB_TARGET) d1++;
    if (kb_max > PJ_MAX_FINE_BITS && d1 < kb_max - PJ_MAX_FINE_BITS) d1 = kb_max - PJ_MAX_FINE_BITS;
    if (d1 > kb_min) d1 = kb_min;
    if (kb_max - d1 > PJ_MAX_FINE_BITS) return false;
    if (((uint64_t)n_combos << d1) >= (1ull << 31)) return false;
    *d1_out = d1;
    for (uint32_t c = 0; c < n_combos; c++) d2_out[c] = (uint8_t)(2u * combo[c].key_nt - d1);
    return true;
}

bool bc_join_supported(const ComboDesc* combo, uint32_t n_combos, uint64_t entries_per_combo) {
    uint32_t d1;
    uint8_t d2[BC_MAX_COMBOS];
    return plan_partition(combo, n_combos, entries_per_combo, &d1, d2);
}

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
    if (!ws.smem_configured) {
        JCK(cudaFuncSetAttribute(k_part_join, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PJ_SMEM_BYTES));
        ws.smem_configured = true;
    }
    PartParams gp;
    memset(&gp, 0, sizeof gp);
    PartJoinParams jp;
    memset(&jp, 0, sizeof jp);
    const uint64_t entries_per_combo = p.dir_entries / p.n_combos;
    if (!plan_partition(p.combo, p.n_combos, entries_per_combo, &gp.d1, gp.d2)) return cudaErrorInvalidValue;
    const uint32_t n_parts = p.n_combos << gp.d1;
    const uint64_t gdir_slots = (uint64_t)n_parts + 1;

    // chunk the genome so the window records stay within the workspace budget
    size_t free_b = 0, total_b = 0;
    JCK(cudaMemGetInfo(&free_b, &total_b));
    uint64_t budget = (uint64_t)free_b + ws.gwin_cap * sizeof(uint4);
    budget = budget / 2;
    if (budget > (64ull << 30)) budget = 64ull << 30;
    uint64_t chunk = budget / sizeof(uint4) / p.n_combos;
    if (chunk > p.n_pos) chunk = p.n_pos;
    if (chunk < 1) chunk = 1;
    const uint64_t rec_needed = chunk * p.n_combos;
    if (rec_needed > ws.gwin_cap) {
        if (ws.d_gwin) cudaFree(ws.d_gwin);
        ws.d_gwin = nullptr;
        ws.gwin_cap = 0;
        JCK(cudaMalloc(&ws.d_gwin, (rec_needed + 1) * sizeof(uint4)));
        ws.gwin_cap = rec_needed;
    }
    if (gdir_slots > ws.gdir_cap) {
        if (ws.d_gdir) cudaFree(ws.d_gdir);
        if (ws.d_gcursor) cudaFree(ws.d_gcursor);
        ws.d_gdir = ws.d_gcursor = nullptr;
        ws.gdir_cap = 0;
        JCK(cudaMalloc(&ws.d_gdir, gdir_slots * 4));
        JCK(cudaMalloc(&ws.d_gcursor, gdir_slots * 4));
        ws.gdir_cap = gdir_slots;
    }
    const uint64_t tmp_words = bc_scan_tmp_words(gdir_slots);
    if (tmp_words > ws.scan_tmp_cap) {
        if (ws.d_scan_tmp) cudaFree(ws.d_scan_tmp);
        ws.d_scan_tmp = nullptr;
        ws.scan_tmp_cap = 0;
        JCK(cudaMalloc(&ws.d_scan_tmp, tmp_words * 4));
        ws.scan_tmp_cap = tmp_words;
    }
    gp.H = p.H; gp.Lo = p.Lo; gp.B = p.B;
    gp.lib_dir = p.dir;
    gp.L = p.L;
    gp.n_combos = p.n_combos;
    memcpy(gp.combo, p.combo, sizeof gp.combo);
    // Skipping windows whose library bucket is empty only pays when most buckets are empty.
    gp.prune = p.dir_entries < (dir_slots - 1) * 2 ? 1u : 0u;
    jp.gdir = ws.d_gdir;
    jp.gwin = ws.d_gwin;
    jp.n_parts = n_parts;
    jp.d1 = gp.d1;
    memcpy(jp.d2, gp.d2, sizeof jp.d2);

    for (uint64_t begin = 0; begin < p.n_pos; begin += chunk) {
        gp.pos_begin = (uint32_t)begin;
        gp.pos_end = (uint32_t)((begin + chunk < p.n_pos) ? begin + chunk : p.n_pos);
        uint32_t npos = gp.pos_end - gp.pos_begin;
        uint32_t gx = (npos + 255) / 256;
        uint32_t maxb = (uint32_t)sm_count * 8u;
        if (gx > maxb) gx = maxb;
        dim3 grid(gx, p.n_combos);
        JCK(cudaMemsetAsync(ws.d_gdir, 0, gdir_slots * 4, st));
        JCK(cudaEventRecord(ws.ev_c, st));
        k_part<0><<<grid, 256, 0, st>>>(gp, ws.d_gdir, nullptr);
        JCK(cudaGetLastError());
        JCK(bc_exclusive_scan(ws.d_gdir, gdir_slots, ws.d_scan_tmp, st));
        JCK(cudaMemcpyAsync(ws.d_gcursor, ws.d_gdir, gdir_slots * 4, cudaMemcpyDeviceToDevice, st));
        k_part<1><<<grid, 256, 0, st>>>(gp, ws.d_gcursor, ws.d_gwin);
        JCK(cudaGetLastError());
        JCK(cudaEventRecord(ws.ev_a, st));
        uint32_t vgrid = (uint32_t)sm_count * 2u;
        if (vgrid > n_parts) vgrid = n_parts;
        k_part_join<<<vgrid, PJ_THREADS, PJ_SMEM_BYTES, st>>>(p, jp);
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
