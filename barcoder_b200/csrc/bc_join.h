// bc_join.h - partition-join scan (K3-join): partitions the genome windows by seed key and
// joins every partition with its slice of the library index in shared memory.  See bc_join.cu.
#pragma once
#include "bc_device.cuh"

struct JoinWorkspace {
    uint32_t* d_gdir = nullptr;     // genome-side directory: window range of every coarse partition
    uint32_t* d_gcursor = nullptr;
    uint4* d_gwin = nullptr;        // {dev position, wh, wl, slot} window records in slot order
    uint4* d_gtmp = nullptr;        // the same records after the coarse (level-1) scatter
    uint32_t* d_coarse_cursor = nullptr;
    uint32_t* d_scan_tmp = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
    uint64_t gdir_cap = 0, gwin_cap = 0, scan_tmp_cap = 0;
    bool smem_configured = false;
    float ms_join_kernels = 0;      // device time of the verify kernels of the last search
    float ms_bucket_kernels = 0;    // device time of the genome bucketing kernels (count, scan, scatter)
};

bool bc_join_supported(const ComboDesc* combo, uint32_t n_combos, uint64_t entries_per_combo);
cudaError_t bc_join_search(JoinWorkspace& ws, const SearchParams& p, uint64_t dir_slots, int sm_count,
                           cudaStream_t st, uint32_t* launches);
void bc_join_free(JoinWorkspace& ws);
