// bc_sort.cu - K4: device-side ordering of the hit records (SURVEY.md 2a "hit sort").
//
// bowtie writes its alignments read by read and, with --best, fewest mismatches first
// (BowtieRunner.py:119); the CUDA search appends records in whatever order the warps find them.
// bc_sort_hits orders the device hit buffer before it is copied out, so the host does not have to
// (np.lexsort of cfg 4's 5.9e7 records takes tens of seconds against a 0.06 s search).
//
// Stable LSD radix sort of the 16-byte records, 4 bits per pass, over the key
//   order 0 "canonical": (spacer_id, gpos, strand)
//   order 1 "best"     : (spacer_id, mismatches, gpos, strand)
// Digits that are constant over the whole buffer (high bits of spacer_id / gpos) are skipped: an
// OR / AND reduction over the records finds them.  Per pass: per-tile digit histograms ->
// exclusive scan (digit-major) -> stable scatter.  Tiles are blocked (every thread owns 16
// consecutive records), which makes stability a per-thread running count.
#include "bc_kernels.h"

#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_BINS 16

struct SortPass {
    uint32_t field;  // 0 spacer_id, 1 gpos, 3 meta
    uint32_t shift, mask;
};

__device__ __forceinline__ uint32_t rs_digit(const uint4& r, const SortPass& sp) {
    const uint32_t w = sp.field == 0 ? r.x : sp.field == 1 ? r.y : r.w;
    return (w >> sp.shift) & sp.mask;
}

// OR and AND of the key words: out[0..2] = OR of (spacer_id, gpos, meta), out[3..5] = AND
__global__ void __launch_bounds__(256) k_rs_orand(const uint4* __restrict__ rec, uint64_t n, uint32_t* __restrict__ out) {
    uint32_t o0 = 0, o1 = 0, o2 = 0, a0 = ~0u, a1 = ~0u, a2 = ~0u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 r = rec[i];
        o0 |= r.x; o1 |= r.y; o2 |= r.w;
        a0 &= r.x; a1 &= r.y; a2 &= r.w;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        o0 |= __shfl_xor_sync(0xffffffffu, o0, d); o1 |= __shfl_xor_sync(0xffffffffu, o1, d); o2 |= __shfl_xor_sync(0xffffffffu, o2, d);
        a0 &= __shfl_xor_sync(0xffffffffu, a0, d); a1 &= __shfl_xor_sync(0xffffffffu, a1, d); a2 &= __shfl_xor_sync(0xffffffffu, a2, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicOr(out + 0, o0); atomicOr(out + 1, o1); atomicOr(out + 2, o2);
        atomicAnd(out + 3, a0); atomicAnd(out + 4, a1); atomicAnd(out + 5, a2);
    }
}

// hist[d * n_tiles + tile] = records of the tile whose digit is d
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint4* __restrict__ rec, uint64_t n, SortPass sp,
                                                        uint32_t n_tiles, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_cnt[RS_BINS];
    if (threadIdx.x < RS_BINS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)threadIdx.x * RS_ITEMS;
    uint32_t c[4] = {0, 0, 0, 0};  // 16 per-thread digit counters of 8 bits (at most RS_ITEMS = 16 each)
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        if (base + i < n) {
            const uint32_t d = rs_digit(rec[base + i], sp);
            c[d >> 2] += 1u << (8 * (d & 3u));
        }
    }
#pragma unroll
    for (int d = 0; d < RS_BINS; d++) {
        uint32_t v = (c[d >> 2] >> (8 * (d & 3))) & 0xffu;
#pragma unroll
        for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[d], v);
    }
    __syncthreads();
    if (threadIdx.x < RS_BINS) hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = s_cnt[threadIdx.x];
}

// stable scatter: position = scanned hist[d][tile] + (records of the tile with digit d before this one)
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint4* __restrict__ rec, uint64_t n, SortPass sp,
                                                           uint32_t n_tiles, const uint32_t* __restrict__ hist,
                                                           uint4* __restrict__ out) {
    __shared__ uint32_t s_warp[RS_BINS][RS_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)threadIdx.x * RS_ITEMS;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint4 r[RS_ITEMS];
    uint32_t dg[RS_ITEMS];
    uint32_t c[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        dg[i] = 0xffffffffu;
        if (base + i < n) {
            r[i] = rec[base + i];
            dg[i] = rs_digit(r[i], sp);
            c[dg[i] >> 2] += 1u << (8 * (dg[i] & 3u));
        }
    }
    // per digit: exclusive prefix of the per-thread counts over the CTA (threads in order)
    uint32_t before[RS_BINS];
#pragma unroll
    for (int d = 0; d < RS_BINS; d++) {
        const uint32_t v = (c[d >> 2] >> (8 * (d & 3))) & 0xffu;
        uint32_t incl = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= (uint32_t)s) incl += o;
        }
        if (lane == 31) s_warp[d][warp] = incl;
        before[d] = incl - v;
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < RS_BINS; d++) {
        uint32_t add = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; w++) add += (uint32_t)w < warp ? s_warp[d][w] : 0u;
        before[d] += add + hist[(size_t)d * n_tiles + blockIdx.x];
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        if (dg[i] == 0xffffffffu) continue;
        uint32_t pos = 0;
#pragma unroll
        for (int d = 0; d < RS_BINS; d++)  // register array indexed by a runtime digit: select, do not spill
            if (dg[i] == (uint32_t)d) { pos = before[d]; before[d] = pos + 1; }
        out[pos] = r[i];
    }
}

#define SCK(call)                           \
    do {                                    \
        cudaError_t e__ = (call);           \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

// Sorts rec[0..n) (16-byte bc_hit records).  *result = the buffer that holds the sorted records
// afterwards: rec itself or `scratch` (the caller swaps its pointers; both hold >= n records).
cudaError_t bc_sort_records(uint4* rec, uint4* scratch, uint64_t n, int order, uint32_t* d_hist, uint64_t hist_words,
                            uint32_t* d_scan_tmp, uint32_t* d_orand, int sm_count, cudaStream_t st, uint4** result,
                            uint32_t* passes_out) {
    *result = rec;
    if (passes_out) *passes_out = 0;
    if (n < 2) return cudaSuccess;
    if (n >= (1ull << 32)) return cudaErrorInvalidValue;
    const uint32_t n_tiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    if ((uint64_t)n_tiles * RS_BINS + 1 > hist_words) return cudaErrorInvalidValue;
    const uint32_t init[6] = {0, 0, 0, ~0u, ~0u, ~0u};
    SCK(cudaMemcpyAsync(d_orand, init, sizeof init, cudaMemcpyHostToDevice, st));
    k_rs_orand<<<(unsigned)sm_count * 4u, 256, 0, st>>>(rec, n, d_orand);
    SCK(cudaGetLastError());
    uint32_t h[6];
    SCK(cudaMemcpyAsync(h, d_orand, sizeof h, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    const uint32_t vary[3] = {h[0] ^ h[3], h[1] ^ h[4], h[2] ^ h[5]};  // bits that differ somewhere
    SortPass passes[24];
    uint32_t np = 0;
    auto add_field = [&](uint32_t field, uint32_t v, uint32_t lo_bit, uint32_t hi_bit) {
        for (uint32_t b = lo_bit; b < hi_bit; b += 4) {
            const uint32_t width = hi_bit - b < 4 ? hi_bit - b : 4, mask = (1u << width) - 1u;
            if ((v >> b) & mask) passes[np++] = SortPass{field, b, mask};
        }
    };
    // least significant first
    add_field(3, vary[2], 0, 1);              // strand
    add_field(1, vary[1], 0, 32);             // gpos
    if (order == 1) add_field(3, vary[2], 1, 3);  // mismatches
    add_field(0, vary[0], 0, 32);             // spacer_id
    uint4 *src = rec, *dst = scratch;
    for (uint32_t i = 0; i < np; i++) {
        k_rs_hist<<<n_tiles, RS_THREADS, 0, st>>>(src, n, passes[i], n_tiles, d_hist);
        SCK(cudaGetLastError());
        SCK(bc_exclusive_scan(d_hist, (uint64_t)n_tiles * RS_BINS, d_scan_tmp, st));
        k_rs_scatter<<<n_tiles, RS_THREADS, 0, st>>>(src, n, passes[i], n_tiles, d_hist, dst);
        SCK(cudaGetLastError());
        uint4* t = src; src = dst; dst = t;
    }
    bc_launch_counter += 1 + 2 * np;
    *result = src;
    if (passes_out) *passes_out = np;
    return cudaSuccess;
}

size_t bc_sort_hist_words(uint64_t n) { return (size_t)((n + RS_TILE - 1) / RS_TILE) * RS_BINS + 1; }
