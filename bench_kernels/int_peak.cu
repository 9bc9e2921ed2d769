// int_peak.cu - measurement-only microbenchmarks (not part of the product ABI): the chip's
// sustained POPC issue rate and the rate of the verification atom of k_join_verify
// (2 LOP3 + POPC + min3 folding), used by bench.py as the integer-pipe roofline denominator.
#include <cuda_runtime.h>
#include <stdint.h>

#define CHAINS 8
#define ITERS 2048

__global__ void __launch_bounds__(256) k_popc_stream(uint32_t* out, uint32_t seed) {
    uint32_t v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) v[i] = seed * (threadIdx.x + 1u) + i * 0x9e3779b9u;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            uint32_t t;
            asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(v[i]));
            v[i] ^= t << 7;
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc ^= v[i];
    if (acc == 0x12345678u) out[0] = acc;
}

// the verify atom of k_verify_dense: two library entries per 16-byte shared-memory broadcast
// against CHAINS resident windows, 2 LOP3 + POPC per pair folded with integer min, one compare
// per group; every result feeds the compare so nothing can be hoisted
__global__ void __launch_bounds__(256) k_verify_atom(uint32_t* out, uint32_t seed, int k) {
    __shared__ uint4 s_lib[256];
    s_lib[threadIdx.x] = make_uint4(seed * (threadIdx.x + 3u), ~seed * (threadIdx.x + 7u),
                                    seed * (threadIdx.x + 11u), ~seed * (threadIdx.x + 13u));
    __syncthreads();
    uint32_t gh[CHAINS], gl[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        gh[i] = seed * (threadIdx.x + 1u) + i * 0x9e3779b9u;
        gl[i] = gh[i] * 0x85ebca6bu;
    }
    uint32_t hits = 0;
    for (int it = 0; it < ITERS / 2; it++) {
        const uint4 q = s_lib[it & 255];
        int best = 33;
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            best = min(best, __popc((gh[i] ^ q.x) | (gl[i] ^ q.y)));
            best = min(best, __popc((gh[i] ^ q.z) | (gl[i] ^ q.w)));
        }
        if (best <= k) hits += best + it;
    }
    if (hits == 77u) out[0] = hits;
}

extern "C" int ub_int_peak(int device, double* popc_per_s, double* atom_per_s, int* sm_count_out) {
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    uint32_t* d_out = nullptr;
    if (cudaMalloc(&d_out, 64) != cudaSuccess) return -1;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int grid = prop.multiProcessorCount * 8;
    const double ops = (double)grid * 256.0 * ITERS * CHAINS;
    double best[2] = {0, 0};
    for (int which = 0; which < 2; which++) {
        for (int rep = 0; rep < 6; rep++) {
            cudaEventRecord(a);
            if (which == 0) k_popc_stream<<<grid, 256>>>(d_out, 12345u + rep);
            else k_verify_atom<<<grid, 256>>>(d_out, 12345u + rep, 3);
            cudaEventRecord(b);
            if (cudaEventSynchronize(b) != cudaSuccess) return -2;
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            double rate = ops / (ms * 1e-3);
            if (rep >= 2 && rate > best[which]) best[which] = rate;  // first reps warm the clocks
        }
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_out);
    *popc_per_s = best[0];
    *atom_per_s = best[1];
    if (sm_count_out) *sm_count_out = prop.multiProcessorCount;
    return 0;
}

// ---------------------------------------------------------------------------------- random gathers
// The probe kernel (k_scan_probe, cfg 1/2/5) is bound by how many DIVERGENT global loads an SM can
// retire: every lane of a directory probe or bucket read touches its own 128-byte line, and the L1TEX
// unit works such a load off one line (wavefront) at a time.  This measures that rate: every thread
// issues independent 4-byte loads at hashed addresses of a table that stays resident in L2 (table_mb
// megabytes; 64 MB ~ the cfg-5 directories), UNROLL loads in flight per thread.
#define GATHER_UNROLL 8
__global__ void __launch_bounds__(256) k_gather(const uint32_t* __restrict__ table, uint32_t mask, uint32_t iters, uint32_t seed,
                                                uint32_t* out) {
    uint32_t x = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + seed;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; it++) {
        uint32_t v[GATHER_UNROLL];
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; u++) {
            x ^= x << 13; x ^= x >> 17; x ^= x << 5;  // xorshift32: independent of the loaded values
            v[u] = __ldg(table + (x & mask));
        }
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; u++) acc += v[u];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

extern "C" int ub_gather_peak(int device, int table_mb, double* gathers_per_s) {
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    size_t words = 1;
    while (words * 4 * 2 <= (size_t)table_mb << 20) words <<= 1;  // largest power of two that fits
    uint32_t *d_table = nullptr, *d_out = nullptr;
    if (cudaMalloc(&d_table, words * 4) != cudaSuccess) return -1;
    if (cudaMalloc(&d_out, 64) != cudaSuccess) return -1;
    cudaMemset(d_table, 1, words * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int grid = prop.multiProcessorCount * 8;
    const uint32_t iters = 256;
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(a);
        k_gather<<<grid, 256>>>(d_table, (uint32_t)(words - 1), iters, 99u + rep, d_out);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) return -2;
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double rate = (double)grid * 256.0 * iters * GATHER_UNROLL / (ms * 1e-3);
        if (rep >= 2 && rate > best) best = rate;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_table);
    cudaFree(d_out);
    *gathers_per_s = best;
    return 0;
}
