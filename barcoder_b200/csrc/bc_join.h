// bc_join.h - bucket-join scan (K3-join): sorts the genome windows by seed key and joins every
// slot with its bucket of the library index.  See bc_join.cu.
#pragma once
#include "bc_device.cuh"

struct JoinWorkspace {
    uint32_t* d_gdir = nullptr;     // genome-side directory: first window record of every slot (+ end sentinel)
    uint32_t* d_gcursor = nullptr;
    uint4* d_gwin = nullptr;        // {dev position, wh, wl, slot} window records in slot order
    uint4* d_gtmp = nullptr;        // radix window sort: the records grouped by bin (pass A output)
    uint32_t* d_bin_cursor = nullptr;  // radix window scatter: write cursor of every bin (slot >> 8)
    uint64_t bin_cap = 0;
    uint32_t* d_work = nullptr;      // per-slice chunk counters of the dense verify kernel (dynamic work distribution)
    uint32_t* d_scan_tmp = nullptr;
    uint4* d_tile_desc = nullptr;    // compact join: warp-tile list of the verify kernel ({first, end, bucket begin, bucket end})
    uint32_t* d_tile_slot = nullptr;
    uint32_t* d_tile_start = nullptr; // first tile of every slot (+ total)
    uint64_t tile_cap = 0, tile_start_cap = 0;
    uint4* d_items = nullptr;        // compact join: {window, entry group} items handed from k_cverify to k_cfinish
    uint64_t item_cap = 0;
    uint64_t item_want = 0;  // item-queue demand measured by an attempt that overflowed (bc_search)
    uint32_t* d_lut = nullptr;       // compact join: byte-wise bit-permutation tables, one per combination
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
    cudaEvent_t ev_k[6] = {nullptr};  // compact join: per-kernel split
    float ms_kernel[6] = {0};         // count, pass A, pass B, tile list, first-level verify, finish
    uint64_t gdir_cap = 0, gwin_cap = 0, scan_tmp_cap = 0;
    float ms_join_kernels = 0;      // device time of the verify kernels of the last search
    float ms_bucket_kernels = 0;    // device time of the genome bucketing kernels (count, scan, scatter)
};

// Streamed result delivery (bc_set_hit_sink / bc_set_slice_callback): while the verify kernels of
// later slices run, the hit records of finished slices are copied to the caller's host buffer on
// a second stream and/or reported to the caller's callback.
#define BC_SINK_SLICES 8
struct HitSink {
    bc_slice_fn fn = nullptr;          // bc_set_slice_callback: told about every finished part of the buffer
    void* fn_user = nullptr;
    uint64_t reported = 0;             // records already reported to fn in this search
    bc_hit* host = nullptr;            // caller's destination: pinned host memory, or device memory of this / a peer GPU
    bool is_device = false;            // the destination is device memory (this or a peer GPU: NVLink), not host memory (PCIe)
    uint64_t cap = 0;                  // records the destination can hold
    uint64_t copied = 0;               // records already queued for copy in this search
    cudaStream_t stream = nullptr;     // copy stream
    unsigned long long* h_counts = nullptr;  // pinned: hit counter after every slice
    cudaEvent_t ev[BC_SINK_SLICES] = {nullptr};
};

cudaError_t bc_join_search(JoinWorkspace& ws, const SearchParams& p, uint64_t dir_slots, int sm_count,
                           cudaStream_t st, uint32_t* launches, HitSink* sink);
cudaError_t bc_sink_deliver(HitSink* sink, const SearchParams& p, uint32_t n_slices);
// compact form (8-byte window records; bc_cjoin.cu): same workspace, the record arrays are viewed as uint2
cudaError_t bc_cjoin_search(JoinWorkspace& ws, const SearchParams& p, uint64_t dir_slots, uint32_t n_bins, int sm_count,
                            cudaStream_t st, uint32_t* launches, HitSink* sink);
struct IndexParams;
cudaError_t bc_cindex_build(JoinWorkspace& ws, const IndexParams& ip, uint32_t n_combos, uint32_t n_bins, uint32_t* d_dir,
                            uint64_t dir_slots, uint32_t* d_cursor, uint32_t* d_scan_tmp, uint2* tmp, uint2* ent_hl,
                            uint32_t* ent_id, int sm_count, cudaStream_t st);
void bc_join_free(JoinWorkspace& ws);
