// bc_join.cu - K3-join: the scan for libraries whose seed buckets are dense (cfg 3/4).
//
// The probe kernel walks a library bucket once per genome window, so with ~300 entries per
// bucket every window re-reads kilobytes of index through L2.  Here the genome side is bucketed
// by the same seed keys instead (histogram -> scan -> scatter of {pos, wh, wl} records), and each
// (library bucket x genome bucket) pair is verified as a dense tile: the library bucket sits in
// shared memory and is broadcast to the warp, every lane keeps JOIN_R genome windows in
// registers, and one pair costs two LOP3, one POPC and a predicated compare.  Algorithmic HBM
// traffic is one write + one read of 16 B per (window, combination) plus the index, read once.
#include "bc_join.h"
#include "bc_kernels.h"

#define JOIN_THREADS 256
#define JOIN_WARPS (JOIN_THREADS / 32)
#define JOIN_R 4
#define JOIN_LIB_TILE 1024  // library entries staged per pass (8 KB of shared memory)

struct GenomeBucketParams {
    const uint32_t* H;
    const uint32_t* Lo;
    const uint32_t* B;
    const uint32_t* lib_dir;  // library directory (to skip windows whose library bucket is empty)
    uint32_t pos_begin, pos_end;  // dev positions handled by this chunk
    uint32_t L, n_combos, prune;
    ComboDesc combo[BC_MAX_COMBOS];
};

// One thread per genome window; every combination's key is counted (pass 0) or the window record
// is scattered to its slot (pass 1).  Windows touching a non-ACGT base or a contig end are
// dropped here, so the verify kernel never sees them.
template <int PASS>
__global__ void __launch_bounds__(256) k_genome_bucket(const __grid_constant__ GenomeBucketParams gp,
                                                       uint32_t* __restrict__ gdir_or_cursor,
                                                       uint4* __restrict__ gwin) {
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
            uint32_t dst = atomicAdd(&gdir_or_cursor[slot], 1u);
            gwin[dst] = make_uint4(pos, wh, wl, 0u);
        }
    }
}

// Verify kernel.  One CTA per (combination, key) bucket; warps split the genome bucket into
// chunks of 32*JOIN_R windows held in registers and sweep the library bucket from shared memory.
__global__ void __launch_bounds__(JOIN_THREADS) k_join_verify(const __grid_constant__ SearchParams p,
                                                              const uint32_t* __restrict__ gdir,
                                                              const uint4* __restrict__ gwin,
                                                              uint32_t n_slots) {
    __shared__ uint2 s_lib[JOIN_LIB_TILE];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int k = (int)p.k;
    unsigned long long cand = 0;
    for (uint32_t slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const uint32_t ls = p.dir[slot], le = p.dir[slot + 1];
        const uint32_t gs = gdir[slot], ge = gdir[slot + 1];
        if (ls == le || gs == ge) continue;  // uniform across the CTA
        uint32_t c = 0;  // combination that owns this directory slot
        while (c + 1 < p.n_combos && p.combo[c + 1].dir_off <= slot) c++;
        const uint32_t ng = ge - gs;
        for (uint32_t lt = ls; lt < le; lt += JOIN_LIB_TILE) {
            const uint32_t nl = min(le - lt, (uint32_t)JOIN_LIB_TILE);
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < nl; i += JOIN_THREADS) s_lib[i] = p.ent_hl[lt + i];
            __syncthreads();
            for (uint32_t base = warp * (32 * JOIN_R); base < ng; base += JOIN_WARPS * 32 * JOIN_R) {
                uint32_t gpos[JOIN_R], gh[JOIN_R], gl[JOIN_R];
#pragma unroll
                for (int r = 0; r < JOIN_R; r++) {
                    uint32_t i = base + r * 32 + lane;
                    // out-of-range lanes replicate the last window; their hits are masked below
                    uint4 w = gwin[gs + min(i, ng - 1)];
                    gpos[r] = i < ng ? w.x : 0xffffffffu;
                    gh[r] = w.y;
                    gl[r] = w.z;
                }
                if (p.count_candidates) {
#pragma unroll
                    for (int r = 0; r < JOIN_R; r++) cand += gpos[r] != 0xffffffffu ? nl : 0;
                }
#pragma unroll 2
                for (uint32_t e = 0; e < nl; e++) {
                    const uint2 q = s_lib[e];
                    uint32_t m[JOIN_R];
                    int best = 33;
#pragma unroll
                    for (int r = 0; r < JOIN_R; r++) {
                        m[r] = (gh[r] ^ q.x) | (gl[r] ^ q.y);
                        best = min(best, __popc(m[r]));
                    }
                    if (best <= k) {
                        const uint32_t id = p.ent_id[lt + e];
#pragma unroll
                        for (int r = 0; r < JOIN_R; r++)
                            if (__popc(m[r]) <= k && gpos[r] != 0xffffffffu) bc_emit_hit(p, c, gpos[r], id, m[r]);
                    }
                }
            }
        }
    }
    if (p.count_candidates) atomicAdd(p.count + 1, cand);
}

// ------------------------------------------------------------------------------------------ host

bool bc_join_supported(const ComboDesc*, uint32_t n_combos) { return n_combos > 0; }

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

#define JCK(call)                                 \
    do {                                          \
        cudaError_t e__ = (call);                 \
        if (e__ != cudaSuccess) return e__;       \
    } while (0)

cudaError_t bc_join_search(JoinWorkspace& ws, const SearchParams& p, uint64_t dir_slots, int sm_count,
                           cudaStream_t st, uint32_t* launches) {
    const uint32_t launches0 = bc_launch_counter;
    ws.ms_join_kernels = ws.ms_bucket_kernels = 0;
    if (!ws.ev_a) JCK(cudaEventCreate(&ws.ev_a));
    if (!ws.ev_b) JCK(cudaEventCreate(&ws.ev_b));
    if (!ws.ev_c) JCK(cudaEventCreate(&ws.ev_c));
    const uint32_t n_slots = (uint32_t)(dir_slots - 1);
    // chunk the genome so the bucketed window records stay within the workspace budget
    const uint64_t budget_records = (24ull << 30) / sizeof(uint4);
    uint64_t chunk = budget_records / p.n_combos;
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
    GenomeBucketParams gp;
    memset(&gp, 0, sizeof gp);
    gp.H = p.H; gp.Lo = p.Lo; gp.B = p.B;
    gp.lib_dir = p.dir;
    gp.L = p.L;
    gp.n_combos = p.n_combos;
    memcpy(gp.combo, p.combo, sizeof gp.combo);
    // Skipping windows whose library bucket is empty only pays when most buckets are empty.
    uint64_t entries = p.dir_entries;
    gp.prune = entries < (uint64_t)n_slots * 2 ? 1u : 0u;

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
        k_genome_bucket<0><<<grid, 256, 0, st>>>(gp, ws.d_gdir, nullptr);
        JCK(cudaGetLastError());
        JCK(bc_exclusive_scan(ws.d_gdir, dir_slots, ws.d_scan_tmp, st));
        JCK(cudaMemcpyAsync(ws.d_gcursor, ws.d_gdir, dir_slots * 4, cudaMemcpyDeviceToDevice, st));
        k_genome_bucket<1><<<grid, 256, 0, st>>>(gp, ws.d_gcursor, ws.d_gwin);
        JCK(cudaGetLastError());
        JCK(cudaEventRecord(ws.ev_a, st));
        uint32_t vgrid = (uint32_t)sm_count * 8u;
        if (vgrid > n_slots) vgrid = n_slots;
        k_join_verify<<<vgrid, JOIN_THREADS, 0, st>>>(p, ws.d_gdir, ws.d_gwin, n_slots);
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
