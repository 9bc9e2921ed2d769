// bc_guides.h - device-side guide enumeration from PAM sites (see bc_guides.cu).
#pragma once
#include "bc_device.cuh"

struct GuideWorkspace {
    unsigned long long* d_table = nullptr;  // hash set of 2-bit guide codes
    unsigned long long* d_out = nullptr;    // compacted distinct codes
    unsigned long long* d_count = nullptr;
    uint64_t table_cap = 0, out_cap = 0, n_guides = 0;
};

cudaError_t bc_guides_enumerate(GuideWorkspace& ws, const uint32_t* H, const uint32_t* Lo, const uint32_t* B,
                                const uint32_t* start_dev, uint32_t n_pos, uint32_t n_contigs, uint32_t L,
                                uint32_t P, const uint32_t* pam_sets, int upstream, int quirks, int sm_count,
                                cudaStream_t st, uint64_t* n_out);
void bc_guides_free(GuideWorkspace& ws);
