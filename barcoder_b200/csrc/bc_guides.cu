// bc_guides.cu - guide enumeration from PAM sites on the packed genome.
//
// Replaces the per-position Python loop of design_guides.find_sequences_with_barcode_and_pam
// (design_guides.py:22-49): every L-mer that is pure ACGT and sits next to a PAM match, on either
// strand of every contig, as a SET (duplicates removed).  One pass over the bit planes that are
// already resident for the search; de-duplication in a device hash set (open addressing,
// atomicCAS on 64-bit codes), then compaction.  A guide is returned as a 2-bit code,
// base j of the guide at bits [2j, 2j+2), A=0 C=1 G=2 T=3.
#include "bc_guides.h"

#include "bc_kernels.h"

#define GUIDE_EMPTY 0xffffffffffffffffull

struct GuideParams {
    const uint32_t* H;
    const uint32_t* Lo;
    const uint32_t* B;
    const uint32_t* start_dev;
    uint32_t n_pos, n_contigs;
    uint32_t L, P, upstream, quirks;
    uint32_t pam_sets[8];
};

__device__ __forceinline__ uint32_t g_bit(const uint32_t* pl, uint32_t d) { return (pl[d >> 5] >> (d & 31u)) & 1u; }

__device__ __forceinline__ uint64_t g_interleave(uint32_t h, uint32_t l) {
    // spread the bits of h and l to odd/even positions: code bits (2j+1, 2j) = (h_j, l_j)
    uint64_t x = h, y = l;
    x = (x | (x << 16)) & 0x0000ffff0000ffffull; y = (y | (y << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;  y = (y | (y << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;  y = (y | (y << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;  y = (y | (y << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;  y = (y | (y << 1)) & 0x5555555555555555ull;
    return (x << 1) | y;
}

// PAM test on the forward strand planes: positions [a, a+P) must be ACGT and match the sets;
// rc = the PAM is read on the reverse strand (position i of the PAM is base a+P-1-i complemented).
__device__ __forceinline__ bool g_pam(const GuideParams& gp, uint32_t a, bool rc) {
    for (uint32_t i = 0; i < gp.P; i++) {
        const uint32_t d = a + (rc ? gp.P - 1 - i : i);
        if (g_bit(gp.B, d)) return false;
        uint32_t code = (g_bit(gp.H, d) << 1) | g_bit(gp.Lo, d);
        if (rc) code = 3u - code;
        if (!((gp.pam_sets[i] >> code) & 1u)) return false;
    }
    return true;
}

__device__ __forceinline__ uint64_t g_mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// MODE 0: count guide occurrences; MODE 1: insert into the hash set
template <int MODE>
__global__ void __launch_bounds__(256) k_guides(const __grid_constant__ GuideParams gp, unsigned long long* count,
                                                unsigned long long* table, uint64_t table_mask, uint32_t* all_t) {
    const uint32_t L = gp.L, P = gp.P;
    const uint32_t lm = bc_lmask(L);
    unsigned long long local = 0;
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < gp.n_pos; pos += gridDim.x * blockDim.x) {
        if (bc_window(gp.B, pos) & lm) continue;  // not pure ACGT, or crosses a contig end
        // contig bounds of this window
        uint32_t lo = 0, hi = gp.n_contigs;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (gp.start_dev[mid] <= pos) lo = mid; else hi = mid;
        }
        const uint32_t cs = gp.start_dev[lo], ce = gp.start_dev[lo + 1] - 1;
        const uint32_t wh = bc_window(gp.H, pos) & lm, wl = bc_window(gp.Lo, pos) & lm;
#pragma unroll
        for (int strand = 0; strand < 2; strand++) {
            // side of the forward window the PAM sits on: right for (downstream,+) and (upstream,-)
            const bool right = (gp.upstream == 0) == (strand == 0);
            bool ok;
            if (right) {
                ok = pos + L + P <= ce;
                // design_guides.py:31 iterates i < len-L-P+1 even for the upstream PAM, so on the
                // reverse strand windows within P of the contig start are never visited
                if (gp.quirks && gp.upstream && pos < cs + P) ok = false;
            } else {
                ok = pos >= cs + P;
                if (gp.quirks && gp.upstream && pos + L + P > ce) ok = false;
            }
            if (!ok) continue;
            if (P && !g_pam(gp, right ? pos + L : pos - P, strand == 1)) continue;
            if (MODE == 0) { local++; continue; }
            uint64_t code;
            if (strand == 0) code = g_interleave(wh, wl);
            else code = g_interleave(bc_rev_bits(~wh & lm, L), bc_rev_bits(~wl & lm, L));
            if (code == GUIDE_EMPTY) { atomicOr(all_t, 1u); continue; }  // 32 x T collides with the sentinel
            uint64_t slot = g_mix(code) & table_mask;
            while (true) {
                const unsigned long long old = atomicCAS(&table[slot], GUIDE_EMPTY, (unsigned long long)code);
                if (old == GUIDE_EMPTY || old == code) break;
                slot = (slot + 1) & table_mask;
            }
        }
    }
    if (MODE == 0 && local) atomicAdd(count, local);
}

__global__ void __launch_bounds__(256) k_guides_compact(const unsigned long long* __restrict__ table, uint64_t n_slots,
                                                        unsigned long long* __restrict__ out,
                                                        unsigned long long* count) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long v = table[i];
        const bool keep = v != GUIDE_EMPTY;
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        if (!ballot) continue;
        const uint32_t lane = threadIdx.x & 31u;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) out[base + __popc(ballot & ((1u << lane) - 1u))] = v;
    }
}

#define GCK(call)                           \
    do {                                    \
        cudaError_t e__ = (call);           \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

cudaError_t bc_guides_enumerate(GuideWorkspace& ws, const uint32_t* H, const uint32_t* Lo, const uint32_t* B,
                                const uint32_t* start_dev, uint32_t n_pos, uint32_t n_contigs, uint32_t L,
                                uint32_t P, const uint32_t* pam_sets, int upstream, int quirks, int sm_count,
                                cudaStream_t st, uint64_t* n_out) {
    GuideParams gp;
    gp.H = H; gp.Lo = Lo; gp.B = B; gp.start_dev = start_dev;
    gp.n_pos = n_pos; gp.n_contigs = n_contigs;
    gp.L = L; gp.P = P; gp.upstream = upstream ? 1u : 0u; gp.quirks = quirks ? 1u : 0u;
    for (int i = 0; i < 8; i++) gp.pam_sets[i] = pam_sets[i];
    if (!ws.d_count) GCK(cudaMalloc(&ws.d_count, 4 * sizeof(unsigned long long)));
    GCK(cudaMemsetAsync(ws.d_count, 0, 4 * sizeof(unsigned long long), st));
    uint32_t grid = (n_pos + 255) / 256;
    const uint32_t maxb = (uint32_t)sm_count * 8u;
    if (grid > maxb) grid = maxb;
    if (grid == 0) grid = 1;
    k_guides<0><<<grid, 256, 0, st>>>(gp, ws.d_count, nullptr, 0, nullptr);
    GCK(cudaGetLastError());
    unsigned long long occurrences = 0;
    GCK(cudaMemcpyAsync(&occurrences, ws.d_count, sizeof occurrences, cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    uint64_t slots = 1024;
    while (slots < 2 * occurrences + 16) slots <<= 1;
    if (slots > ws.table_cap) {
        if (ws.d_table) cudaFree(ws.d_table);
        ws.d_table = nullptr;
        ws.table_cap = 0;
        GCK(cudaMalloc(&ws.d_table, slots * sizeof(unsigned long long)));
        ws.table_cap = slots;
    }
    if (occurrences + 2 > ws.out_cap) {
        if (ws.d_out) cudaFree(ws.d_out);
        ws.d_out = nullptr;
        ws.out_cap = 0;
        GCK(cudaMalloc(&ws.d_out, (occurrences + 2) * sizeof(unsigned long long)));
        ws.out_cap = occurrences + 2;
    }
    GCK(cudaMemsetAsync(ws.d_table, 0xff, slots * sizeof(unsigned long long), st));
    GCK(cudaMemsetAsync(ws.d_count, 0, 4 * sizeof(unsigned long long), st));
    uint32_t* d_all_t = reinterpret_cast<uint32_t*>(ws.d_count + 2);
    k_guides<1><<<grid, 256, 0, st>>>(gp, ws.d_count, ws.d_table, slots - 1, d_all_t);
    GCK(cudaGetLastError());
    uint64_t cgrid = (slots + 255) / 256;
    if (cgrid > maxb) cgrid = maxb;
    k_guides_compact<<<(uint32_t)cgrid, 256, 0, st>>>(ws.d_table, slots, ws.d_out, ws.d_count);
    GCK(cudaGetLastError());
    bc_launch_counter += 3;
    unsigned long long res[4];
    GCK(cudaMemcpyAsync(res, ws.d_count, sizeof res, cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    uint64_t n = res[0];
    if (reinterpret_cast<uint32_t*>(&res[2])[0]) {  // the all-T 32-mer, kept out of the table
        const unsigned long long v = GUIDE_EMPTY;
        GCK(cudaMemcpyAsync(ws.d_out + n, &v, sizeof v, cudaMemcpyHostToDevice, st));
        GCK(cudaStreamSynchronize(st));
        n++;
    }
    ws.n_guides = n;
    *n_out = n;
    return cudaSuccess;
}

void bc_guides_free(GuideWorkspace& ws) {
    if (ws.d_table) cudaFree(ws.d_table);
    if (ws.d_out) cudaFree(ws.d_out);
    if (ws.d_count) cudaFree(ws.d_count);
    ws = GuideWorkspace();
}
