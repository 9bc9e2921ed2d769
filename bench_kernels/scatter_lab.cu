// scatter_lab.cu - measurement-only: what random atomics and random 16-byte stores cost on this GPU.
// Informs the design of the counting-sort scatter.  Not part of the product.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// mode 0: RED to table[h % slots]; 1: ATOM returning (result consumed); 2: random 16B store to out[h % n];
// 3: ATOM + dependent random 16B store (counting-sort scatter); 4: sequential 16B store; 5: random 32B store
__global__ void __launch_bounds__(256) k_lab(int mode, uint32_t n, uint32_t slots, uint32_t* table, uint4* out,
                                             uint32_t* sink) {
    uint32_t acc = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t h = hash32(i);
        const uint32_t s = h % slots;
        if (mode == 0) atomicAdd(&table[s], 1u);
        else if (mode == 1) acc += atomicAdd(&table[s], 1u);
        else if (mode == 2) out[hash32(h) % n] = make_uint4(i, h, s, 0);
        else if (mode == 3) {
            // slot s owns positions [s * (n / slots), ...): emulate cursor-based placement
            const uint32_t r = atomicAdd(&table[s], 1u);
            const uint64_t base = (uint64_t)s * (n / slots);
            out[(base + r) % n] = make_uint4(i, h, s, r);
        } else if (mode == 4) out[i] = make_uint4(i, h, s, 0);
        else if (mode == 5) {
            const uint32_t d = (hash32(h) % (n / 2)) * 2;
            out[d] = make_uint4(i, h, s, 0);
            out[d + 1] = make_uint4(i, h, s, 1);
        }
    }
    if (acc == 0x12345u) sink[0] = acc;
}

int main() {
    const uint32_t n = 1u << 30;  // 1.07e9 records, 16 GiB of output
    uint32_t *table, *sink;
    uint4* out;
    cudaMalloc(&table, (size_t)(1u << 24) * 4);
    cudaMalloc(&sink, 64);
    cudaMalloc(&out, (size_t)n * 16);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const char* names[] = {"RED", "ATOM(ret)", "store16 random", "ATOM+store16 (sort scatter)", "store16 sequential",
                           "store32 random"};
    const uint32_t slot_opts[] = {1u << 12, 1u << 16, 1u << 20, 1u << 24};
    for (int mode = 0; mode < 6; mode++) {
        for (int so = 0; so < 4; so++) {
            if ((mode == 2 || mode == 4 || mode == 5) && so > 0) continue;
            const uint32_t slots = slot_opts[so];
            cudaMemset(table, 0, (size_t)(1u << 24) * 4);
            k_lab<<<148 * 8, 256>>>(mode, 1u << 20, slots, table, out, sink);  // warm-up
            cudaMemset(table, 0, (size_t)(1u << 24) * 4);
            cudaEventRecord(a);
            k_lab<<<148 * 8, 256>>>(mode, n, slots, table, out, sink);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            printf("%-30s slots=2^%-2d  %8.2f ms  %7.2f ps/record  %6.1f Gops/s\n", names[mode],
                   31 - __builtin_clz(slots), ms, ms * 1e9 / n, n / ms / 1e6);
        }
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
