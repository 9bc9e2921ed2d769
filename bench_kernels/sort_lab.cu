// sort_lab.cu - measurement-only: cost per record of the building blocks of the compact (8-byte
// record) window sort, to pick its shape before writing it.  Not product code.
//
//   A-direct   chunk of CH records -> NB bins: shared-memory histogram, ONE global atomic per
//              (chunk, bin), then every record stored straight from registers to
//              tmp[bin_base + rank] (8-byte stores; runs only form in L2)
//   A-staged   same, but the chunk is first grouped by bin in shared memory and leaves as runs
//   B-perbin   one CTA per bin region: shared-memory cursors (no global atomics), direct 8-byte
//              stores to out[slot_base + rank], NS sub-slots per bin
//   B-chunk    bin regions cut into chunks, shared-memory sort by sub-slot, one global atomic per
//              (chunk, sub-slot), run stores
// Keys are hashes of the record index (uniform).  Prints ps per record.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// ------------------------------------------------------------------ pass A
template <int THREADS, int ITEMS, bool STAGED>
__global__ void __launch_bounds__(THREADS) k_passA(uint32_t n, uint32_t nb_log, uint32_t* __restrict__ bin_cursor,
                                                   uint2* __restrict__ tmp) {
    extern __shared__ uint32_t sm[];
    const uint32_t NB = 1u << nb_log;
    uint32_t* s_hist = sm;            // [NB]
    uint32_t* s_gbase = sm + NB;      // [NB]
    uint32_t* s_lstart = sm + 2 * NB; // [NB] (staged only)
    uint2* s_rec = reinterpret_cast<uint2*>(sm + 3 * NB);
    const uint32_t CH = THREADS * ITEMS;
    for (uint32_t c0 = blockIdx.x * CH; c0 < n; c0 += gridDim.x * CH) {
        for (uint32_t j = threadIdx.x; j < NB; j += THREADS) s_hist[j] = 0;
        __syncthreads();
        uint32_t bin[ITEMS], rank[ITEMS], key[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const uint32_t idx = c0 + threadIdx.x + i * THREADS;
            key[i] = hash32(idx);
            bin[i] = idx < n ? key[i] >> (32 - nb_log) : 0xffffffffu;
            if (idx < n) rank[i] = atomicAdd(&s_hist[bin[i]], 1u);
        }
        __syncthreads();
        if (STAGED) {  // exclusive scan of the histogram by thread 0..: simple two-level
            // each thread owns NB/THREADS consecutive counters
            const uint32_t per = (NB + THREADS - 1) / THREADS;
            uint32_t s = 0;
            for (uint32_t j = 0; j < per; j++) { const uint32_t b = threadIdx.x * per + j; if (b < NB) s += s_hist[b]; }
            // block scan of s
            __shared__ uint32_t wsum[32];
            uint32_t incl = s;
            for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if ((threadIdx.x & 31) >= d) incl += o; }
            if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
            __syncthreads();
            uint32_t before = 0;
            for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) before += wsum[w];
            uint32_t run = before + incl - s;
            for (uint32_t j = 0; j < per; j++) { const uint32_t b = threadIdx.x * per + j; if (b < NB) { s_lstart[b] = run; run += s_hist[b]; } }
        }
        for (uint32_t j = threadIdx.x; j < NB; j += THREADS) {
            const uint32_t cnt = s_hist[j];
            if (cnt) s_gbase[j] = atomicAdd(&bin_cursor[j], cnt);
        }
        __syncthreads();
        if (!STAGED) {
#pragma unroll
            for (int i = 0; i < ITEMS; i++)
                if (bin[i] != 0xffffffffu) tmp[s_gbase[bin[i]] + rank[i]] = make_uint2(c0 + threadIdx.x + i * THREADS, key[i]);
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; i++)
                if (bin[i] != 0xffffffffu) s_rec[s_lstart[bin[i]] + rank[i]] = make_uint2(c0 + threadIdx.x + i * THREADS, key[i]);
            __syncthreads();
            const uint32_t total = min(CH, n - c0);
            for (uint32_t i = threadIdx.x; i < total; i += THREADS) {
                const uint2 r = s_rec[i];
                const uint32_t b = r.y >> (32 - nb_log);
                tmp[s_gbase[b] + (i - s_lstart[b])] = r;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ pass B, one CTA per bin
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_passB_perbin(const uint2* __restrict__ tmp, const uint32_t* __restrict__ bin_start,
                                                          uint32_t n_bins, uint32_t nb_log, uint32_t ns_log,
                                                          const uint32_t* __restrict__ slot_base, uint2* __restrict__ out,
                                                          uint32_t* __restrict__ work) {
    extern __shared__ uint32_t sm[];
    const uint32_t NS = 1u << ns_log;
    uint32_t* s_cur = sm;  // [NS]
    __shared__ uint32_t s_bin;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_bin = atomicAdd(work, 1u);
        __syncthreads();
        const uint32_t b = s_bin;
        if (b >= n_bins) break;
        for (uint32_t j = threadIdx.x; j < NS; j += THREADS) s_cur[j] = slot_base[(size_t)b * NS + j];
        __syncthreads();
        const uint32_t r0 = bin_start[b], r1 = bin_start[b + 1];
        for (uint32_t base = r0; base < r1; base += THREADS * ITEMS) {
            uint2 rec[ITEMS];
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const uint32_t idx = base + threadIdx.x + i * THREADS;
                if (idx < r1) rec[i] = __ldcs(tmp + idx);
            }
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const uint32_t idx = base + threadIdx.x + i * THREADS;
                if (idx < r1) {
                    const uint32_t sub = (rec[i].y >> (32 - nb_log - ns_log)) & (NS - 1);
                    const uint32_t dst = atomicAdd(&s_cur[sub], 1u);
                    out[dst] = rec[i];
                }
            }
        }
    }
}

// ------------------------------------------------------------------ pass B, chunked with shared-memory sort
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_passB_chunk(const uint2* __restrict__ tmp, const uint32_t* __restrict__ bin_start,
                                                         uint32_t n, uint32_t nb_log, uint32_t ns_log,
                                                         uint32_t* __restrict__ slot_cursor, uint2* __restrict__ out,
                                                         uint32_t* __restrict__ work) {
    extern __shared__ uint32_t sm[];
    const uint32_t NS = 1u << ns_log;
    uint32_t* s_hist = sm;
    uint32_t* s_gbase = sm + NS;
    uint32_t* s_lstart = sm + 2 * NS;
    uint2* s_rec = reinterpret_cast<uint2*>(sm + 3 * NS);
    __shared__ uint32_t s_chunk;
    __shared__ uint32_t wsum[32];
    const uint32_t CH = THREADS * ITEMS;
    const uint32_t n_chunks = (n + CH - 1) / CH;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(work, 1u);
        __syncthreads();
        const uint32_t ch = s_chunk;
        if (ch >= n_chunks) break;
        const uint32_t r0 = ch * CH, r1 = min(n, r0 + CH);
        uint32_t seg = r0;
        while (seg < r1) {
            const uint32_t g = __ldg(&tmp[seg].y) >> (32 - nb_log);
            const uint32_t s1 = min(r1, __ldg(bin_start + g + 1));
            for (uint32_t j = threadIdx.x; j < NS; j += THREADS) s_hist[j] = 0;
            __syncthreads();
            uint2 rec[ITEMS];
            uint32_t rank[ITEMS], sub[ITEMS];
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const uint32_t idx = seg + threadIdx.x + i * THREADS;
                if (idx < s1) {
                    rec[i] = __ldcs(tmp + idx);
                    sub[i] = (rec[i].y >> (32 - nb_log - ns_log)) & (NS - 1);
                    rank[i] = atomicAdd(&s_hist[sub[i]], 1u);
                }
            }
            __syncthreads();
            const uint32_t per = (NS + THREADS - 1) / THREADS;
            uint32_t s = 0;
            for (uint32_t j = 0; j < per; j++) { const uint32_t b = threadIdx.x * per + j; if (b < NS) s += s_hist[b]; }
            uint32_t incl = s;
            for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if ((threadIdx.x & 31) >= d) incl += o; }
            if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
            __syncthreads();
            uint32_t before = 0;
            for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) before += wsum[w];
            uint32_t run = before + incl - s;
            for (uint32_t j = 0; j < per; j++) {
                const uint32_t b = threadIdx.x * per + j;
                if (b < NS) {
                    s_lstart[b] = run;
                    const uint32_t cnt = s_hist[b];
                    run += cnt;
                    if (cnt) s_gbase[b] = atomicAdd(&slot_cursor[(size_t)g * NS + b], cnt);
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const uint32_t idx = seg + threadIdx.x + i * THREADS;
                if (idx < s1) s_rec[s_lstart[sub[i]] + rank[i]] = rec[i];
            }
            __syncthreads();
            const uint32_t cnt = s1 - seg;
            for (uint32_t i = threadIdx.x; i < cnt; i += THREADS) {
                const uint2 r = s_rec[i];
                const uint32_t sb = (r.y >> (32 - nb_log - ns_log)) & (NS - 1);
                out[s_gbase[sb] + (i - s_lstart[sb])] = r;
            }
            __syncthreads();
            seg = s1;
        }
    }
}

// exact counts for the synthetic keys
__global__ void k_count(uint32_t n, uint32_t bits, uint32_t* cnt) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&cnt[hash32(i) >> (32 - bits)], 1u);
}
__global__ void k_check(const uint2* out, uint32_t n, uint32_t bits, uint32_t* bad) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += gridDim.x * blockDim.x) {
        if ((out[i].y >> (32 - bits)) > (out[i + 1].y >> (32 - bits))) atomicAdd(bad, 1u);
        if (hash32(out[i].x) != out[i].y) atomicAdd(bad, 1u);
    }
}

static void exclusive_scan_host(uint32_t* d, size_t n) {
    uint32_t* h = (uint32_t*)malloc((n + 1) * 4);
    cudaMemcpy(h, d, n * 4, cudaMemcpyDeviceToHost);
    uint32_t run = 0;
    for (size_t i = 0; i < n; i++) { uint32_t c = h[i]; h[i] = run; run += c; }
    h[n] = run;
    cudaMemcpy(d, h, (n + 1) * 4, cudaMemcpyHostToDevice);
    free(h);
}

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atol(argv[1]) : 1000000000u;
    uint2 *tmp, *out;
    uint32_t *cntA, *curA, *cntS, *curS, *work, *bad;
    cudaMalloc(&tmp, (size_t)n * 8);
    cudaMalloc(&out, (size_t)n * 8);
    cudaMalloc(&cntA, ((1u << 12) + 1) * 4);
    cudaMalloc(&curA, ((1u << 12) + 1) * 4);
    cudaMalloc(&cntS, ((size_t)(1u << 24) + 1) * 4);
    cudaMalloc(&curS, ((size_t)(1u << 24) + 1) * 4);
    cudaMalloc(&work, 4);
    cudaMalloc(&bad, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float ms;
    int sms = 148;
    for (uint32_t nb_log = 8; nb_log <= 10; nb_log += 2) {
        const uint32_t NB = 1u << nb_log;
        cudaMemset(cntA, 0, (NB + 1) * 4);
        k_count<<<sms * 8, 256>>>(n, nb_log, cntA);
        exclusive_scan_host(cntA, NB);
#define RUN_A(TH, IT, ST, NAME)                                                                               \
    do {                                                                                                      \
        cudaMemcpy(curA, cntA, (NB + 1) * 4, cudaMemcpyDeviceToDevice);                                        \
        const size_t smem = 3 * NB * 4 + (ST ? (size_t)TH * IT * 8 : 0);                                      \
        cudaFuncSetAttribute(k_passA<TH, IT, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        int occ = 0;                                                                                          \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_passA<TH, IT, ST>, TH, smem);                   \
        cudaEventRecord(a);                                                                                   \
        k_passA<TH, IT, ST><<<sms * (occ ? occ : 1), TH, smem>>>(n, nb_log, curA, tmp);                       \
        cudaEventRecord(b);                                                                                   \
        cudaEventSynchronize(b);                                                                              \
        cudaEventElapsedTime(&ms, a, b);                                                                      \
        printf("A %-28s bins=%4u chunk=%5d occ=%d: %8.3f ms  %6.2f ps/rec  %s\n", NAME, NB, TH * IT, occ, ms, \
               ms * 1e9 / n, cudaGetErrorString(cudaGetLastError()));                                         \
    } while (0)
        RUN_A(256, 8, false, "direct 256x8");
        RUN_A(512, 8, false, "direct 512x8");
        RUN_A(512, 16, false, "direct 512x16");
        RUN_A(1024, 8, false, "direct 1024x8");
        RUN_A(256, 8, true, "staged 256x8");
        RUN_A(512, 8, true, "staged 512x8");
        RUN_A(512, 16, true, "staged 512x16");
        RUN_A(1024, 8, true, "staged 1024x8");
        // pass B on the output of the last pass A
        for (uint32_t ns_log = 8; ns_log <= 12; ns_log += 2) {
            if (nb_log + ns_log > 24) continue;
            const uint32_t NS = 1u << ns_log;
            const size_t nslots = (size_t)NB * NS;
            cudaMemset(cntS, 0, (nslots + 1) * 4);
            k_count<<<sms * 8, 256>>>(n, nb_log + ns_log, cntS);
            exclusive_scan_host(cntS, nslots);
#define RUN_BP(TH, IT, NAME)                                                                                    \
    do {                                                                                                        \
        cudaMemset(work, 0, 4);                                                                                 \
        const size_t smem = (size_t)NS * 4;                                                                     \
        cudaFuncSetAttribute(k_passB_perbin<TH, IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        int occ = 0;                                                                                            \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_passB_perbin<TH, IT>, TH, smem);                  \
        cudaEventRecord(a);                                                                                     \
        k_passB_perbin<TH, IT><<<sms * (occ ? occ : 1), TH, smem>>>(tmp, cntA, NB, nb_log, ns_log, cntS, out, work); \
        cudaEventRecord(b);                                                                                     \
        cudaEventSynchronize(b);                                                                                \
        cudaEventElapsedTime(&ms, a, b);                                                                        \
        cudaMemset(bad, 0, 4);                                                                                  \
        k_check<<<sms * 8, 256>>>(out, n, nb_log + ns_log, bad);                                                \
        uint32_t hbad = 0;                                                                                      \
        cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);                                                      \
        printf("B %-28s bins=%4u sub=%4u occ=%d: %8.3f ms  %6.2f ps/rec  bad=%u %s\n", NAME, NB, NS, occ, ms,   \
               ms * 1e9 / n, hbad, cudaGetErrorString(cudaGetLastError()));                                     \
    } while (0)
#define RUN_BC(TH, IT, NAME)                                                                                    \
    do {                                                                                                        \
        cudaMemset(work, 0, 4);                                                                                 \
        cudaMemcpy(curS, cntS, (nslots + 1) * 4, cudaMemcpyDeviceToDevice);                                      \
        const size_t smem = (size_t)3 * NS * 4 + (size_t)TH * IT * 8;                                           \
        cudaFuncSetAttribute(k_passB_chunk<TH, IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        int occ = 0;                                                                                            \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_passB_chunk<TH, IT>, TH, smem);                   \
        cudaEventRecord(a);                                                                                     \
        k_passB_chunk<TH, IT><<<sms * (occ ? occ : 1), TH, smem>>>(tmp, cntA, n, nb_log, ns_log, curS, out, work); \
        cudaEventRecord(b);                                                                                     \
        cudaEventSynchronize(b);                                                                                \
        cudaEventElapsedTime(&ms, a, b);                                                                        \
        cudaMemset(bad, 0, 4);                                                                                  \
        k_check<<<sms * 8, 256>>>(out, n, nb_log + ns_log, bad);                                                \
        uint32_t hbad = 0;                                                                                      \
        cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);                                                      \
        printf("B %-28s bins=%4u sub=%4u occ=%d: %8.3f ms  %6.2f ps/rec  bad=%u %s\n", NAME, NB, NS, occ, ms,   \
               ms * 1e9 / n, hbad, cudaGetErrorString(cudaGetLastError()));                                     \
    } while (0)
            RUN_BP(256, 8, "perbin 256x8");
            RUN_BP(512, 8, "perbin 512x8");
            RUN_BP(1024, 4, "perbin 1024x4");
            RUN_BC(256, 8, "chunk-sorted 256x8");
            RUN_BC(512, 16, "chunk-sorted 512x16");
            RUN_BC(1024, 8, "chunk-sorted 1024x8");
        }
    }
    // reference points: RED count pass and a plain copy
    cudaMemset(cntS, 0, ((size_t)(1u << 24) + 1) * 4);
    cudaEventRecord(a);
    k_count<<<sms * 8, 256>>>(n, 20, cntS);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    printf("RED count 2^20 slots: %8.3f ms  %6.2f ps/rec\n", ms, ms * 1e9 / n);
    cudaEventRecord(a);
    cudaMemcpyAsync(out, tmp, (size_t)n * 8, cudaMemcpyDeviceToDevice);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    printf("D2D copy 8 B records: %8.3f ms  %6.2f ps/rec\n", ms, ms * 1e9 / n);
    return 0;
}
