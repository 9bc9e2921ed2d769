// bc_kernels.h - host-callable launchers implemented in bc_kernels.cu / bc_join.cu.
#pragma once
#include "bc_device.cuh"

struct IndexParams {
    const uint32_t* qh;
    const uint32_t* ql;
    const uint32_t* sn;
    uint32_t n_entries;  // 2 * n spacers
    uint32_t L;
    uint32_t lib_has_n;
    ComboDesc combo[BC_MAX_COMBOS];
};

// Every launcher adds the kernels it launched here (bench.py reports gpu_launches from it).
extern thread_local uint32_t bc_launch_counter;

size_t bc_scan_tmp_words(uint64_t n);
cudaError_t bc_exclusive_scan(uint32_t* d_data, uint64_t n, uint32_t* d_tmp, cudaStream_t st);

cudaError_t bc_launch_pack_genome(const uint8_t* d_ascii, const uint64_t* d_coff, const uint32_t* d_start_dev,
                                  uint32_t n_contigs, uint32_t n_pos, uint32_t n_words, uint32_t* H, uint32_t* Lo,
                                  uint32_t* B, int sm_count, cudaStream_t st);
cudaError_t bc_launch_pack_library(const uint8_t* d_ascii, uint32_t n, uint32_t L, uint32_t* qh, uint32_t* ql,
                                   uint32_t* sn, uint32_t* any_n, cudaStream_t st);
cudaError_t bc_launch_index_build(const IndexParams& ip, uint32_t n_combos, uint32_t* d_dir, uint64_t dir_slots,
                                  uint32_t* d_cursor, uint32_t* d_scan_tmp, uint2* ent_hl, uint32_t* ent_id,
                                  int sm_count, cudaStream_t st);
cudaError_t bc_launch_scan_probe(const SearchParams& p, int sm_count, cudaStream_t st);
