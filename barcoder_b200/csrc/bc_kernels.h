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
    uint32_t slot_lo, slot_hi;  // slot-range sharding: only entries whose slot is in [slot_lo, slot_hi) are indexed
    uint32_t bin_aligned;       // compact join: slot_lo / slot_hi lie on pass-A bin boundaries
    uint32_t compact;    // 1: ent_hl holds the non-key (rem) planes {Rh, Rl} instead of the full query planes
    ComboDesc combo[BC_MAX_COMBOS];
};

// Two-level counting-sort scatter (shared by the library index and the genome-window sort).
// A one-level scatter of 16-byte records into ~10^6 slots keeps ~10^6 partially written 128-byte
// lines open at once - more than L2 holds - and ncu showed 3.4x the algorithmic DRAM traffic
// (sector read-modify-write).  Level 1 therefore scatters into <= 2^14 coarse partitions per
// combination (few open lines, fully merged in L2); level 2 streams the coarse array and places
// every record at its final slot, touching only the few hundred slots of the partitions in flight.
struct CoarsePlan {
    uint32_t n_combos;
    uint32_t n_coarse;
    uint32_t shift[BC_MAX_COMBOS];           // fine slots per coarse partition = 1 << shift
    uint32_t coarse_off[BC_MAX_COMBOS + 1];  // first coarse partition of each combination
    uint32_t dir_off[BC_MAX_COMBOS + 1];     // first directory slot of each combination
};

#define BC_COARSE_BITS 14

__device__ __forceinline__ uint32_t bc_coarse_of(const CoarsePlan& pl, uint32_t c, uint32_t slot) {
    return pl.coarse_off[c] + ((slot - pl.dir_off[c]) >> pl.shift[c]);
}

void bc_make_coarse_plan(const ComboDesc* combo, uint32_t n_combos, CoarsePlan* pl);
// coarse_cursor[j] = fine_dir[first slot of coarse partition j]  (fine_dir already scanned)
cudaError_t bc_launch_coarse_init(const CoarsePlan& pl, const uint32_t* fine_dir, uint32_t* coarse_cursor,
                                  cudaStream_t st);
// level 2: tmp records {a, b, c, slot} -> final position atomicAdd(fine_cursor[slot]).
//   mode 0: out_rec[dst] = record (genome windows);  mode 1: out_hl[dst] = {a, b}, out_id[dst] = c (index)
cudaError_t bc_launch_fine_scatter(int mode, const uint4* tmp, const uint32_t* n_rec_ptr, uint32_t* fine_cursor,
                                   uint4* out_rec, uint2* out_hl, uint32_t* out_id, int sm_count, cudaStream_t st);

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
                                  uint32_t* d_cursor, uint32_t* d_scan_tmp, uint4* ent_tmp, uint32_t* coarse_cursor,
                                  uint2* ent_hl, uint32_t* ent_id, int sm_count, cudaStream_t st);
#ifndef BC_PROBE_PIN_DIRECTORY
#define BC_PROBE_PIN_DIRECTORY 1   // probe path: keep the seed directory in the persisting L2 set-aside
#endif
cudaError_t bc_launch_scan_probe(const SearchParams& p, uint64_t dir_bytes, int sm_count, cudaStream_t st);
void bc_probe_release_l2();

// bc_sort.cu: stable LSD radix sort of the 16-byte hit records (order 0: spacer_id, gpos, strand;
// order 1: spacer_id, mismatches, gpos, strand).  *result = rec or scratch, whichever holds the output.
cudaError_t bc_sort_records(uint4* rec, uint4* scratch, uint64_t n, int order, uint32_t* d_hist, uint64_t hist_words,
                            uint32_t* d_scan_tmp, uint32_t* d_orand, int sm_count, cudaStream_t st, uint4** result,
                            uint32_t* passes_out);
size_t bc_sort_hist_words(uint64_t n);

// probe path: packed directory, one 4-byte load per probe (start | count << 26)
cudaError_t bc_launch_dir_pack(const uint32_t* dir, uint32_t n_slots, uint32_t* pdir, int sm_count, cudaStream_t st);
cudaError_t bc_launch_ent_h_pack(const uint2* ent_hl, uint64_t n, uint32_t* ent_h, int sm_count, cudaStream_t st);
