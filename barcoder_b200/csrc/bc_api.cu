// bc_api.cu - the C ABI declared in include/barcoder_b200.h: context, device memory,
// seed-scheme selection and kernel orchestration.  No CPU search path exists in this file:
// every data-touching step is a kernel launch; the host only plans and moves pointers.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "bc_kernels.h"
#include "bc_join.h"
#include "bc_guides.h"

struct bc_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    cudaEvent_t ev_i0 = nullptr, ev_i1 = nullptr;  // bc_build_index: enqueued without a host synchronisation, timed lazily
    bool index_timing_pending = false;
    std::string err;

    // genome
    uint64_t G = 0;
    uint32_t n_contigs = 0, n_pos = 0, n_words = 0;
    std::vector<uint64_t> coff;
    uint32_t *d_H = nullptr, *d_L = nullptr, *d_B = nullptr, *d_start_dev = nullptr;
    bool have_genome = false;
    uint64_t plane_cap = 0, start_cap = 0;       // allocated plane words / contig slots (grow-only)
    uint8_t* d_stage = nullptr;                  // grow-only staging buffer for host ASCII inputs
    uint64_t stage_cap = 0;
    uint64_t* d_coff = nullptr;
    uint64_t coff_cap = 0;

    // library
    uint32_t n = 0, L = 0;
    uint32_t *d_qh = nullptr, *d_ql = nullptr, *d_sn = nullptr, *d_any_n = nullptr;
    uint32_t lib_has_n = 0;
    bool have_library = false;
    uint64_t lib_cap = 0;                        // allocated spacers (grow-only)

    // PAM
    uint32_t P = 0, pam_dir = 0, pam_flags = 0, pam_sets[8] = {0};

    // params
    int64_t par_blocks = 0, par_path = 0, par_count = 0, par_hit_cap = 0, par_id_base = 0;
    int64_t par_scan_rank = 0, par_scan_world = 1, par_window_sort = 0, par_join_chunk = 0, par_key_nt = 0;
    int64_t par_slot_rank = 0, par_slot_world = 1, par_index_sort = 0, par_key_cap = 0, par_compact_dir = 0;

    // index
    bool have_index = false;
    int index_k = -1;
    uint32_t b = 0, n_combos = 0, n_bins = 0, slot_lo = 0, slot_hi = 0;
    uint32_t block_mask[BC_MAX_BLOCKS] = {0};
    ComboDesc combo[BC_MAX_COMBOS];
    uint64_t dir_slots = 0;
    uint32_t *d_dir = nullptr, *d_cursor = nullptr, *d_scan_tmp = nullptr, *d_ent_id = nullptr;
    uint2* d_ent_hl = nullptr;
    uint4* d_ent_tmp = nullptr;       // level-1 (coarse) output of the index scatter
    uint32_t* d_coarse_cursor = nullptr;
    uint64_t ent_cap = 0, dir_cap = 0, scan_tmp_cap = 0;
    uint32_t* d_pdir = nullptr;       // probe path: packed directory (bc_launch_dir_pack)
    uint64_t pdir_cap = 0;
    uint32_t* d_ent_h = nullptr;      // probe path: H planes of the index entries (bc_launch_ent_h_pack)
    uint64_t ent_h_cap = 0;
    bool have_ent_h = false;
    bool packed_dir = false;

    // join workspace
    JoinWorkspace join;
    GuideWorkspace guides;

    // results
    bc_hit* d_hits = nullptr;
    uint64_t hit_cap = 0, n_hits = 0;
    unsigned long long* d_count = nullptr;
    HitSink sink;                     // optional streamed delivery to host memory (bc_set_hit_sink)
    bc_hit* d_sort_scratch = nullptr; // bc_sort_hits: ping-pong buffer, histogram, scan scratch
    uint32_t *d_sort_hist = nullptr, *d_sort_tmp = nullptr, *d_sort_orand = nullptr;
    uint64_t sort_cap = 0, sort_hist_cap = 0, sort_tmp_cap = 0;

    bc_stats stats;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char buf__[512];                                                                       \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                     __FILE__, __LINE__);                                                          \
            ctx->err = buf__;                                                                      \
            return e__ == cudaErrorMemoryAllocation ? BC_ENOMEM : BC_ECUDA;                        \
        }                                                                                          \
    } while (0)

static int fail(bc_ctx* ctx, int code, const char* msg) {
    if (ctx) ctx->err = msg;
    return code;
}

template <class T>
static void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

static thread_local std::string g_create_err;

extern "C" int bc_abi_version(void) { return BC_ABI_VERSION; }

extern "C" int bc_device_count(void) {
    int count = 0;
    return cudaGetDeviceCount(&count) == cudaSuccess ? count : 0;
}

extern "C" int bc_create(bc_ctx** out, int device) {
    if (!out) return BC_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_err = std::string("no usable CUDA device: ") + cudaGetErrorString(e);
        return BC_ENODEV;
    }
    if (device < 0 || device >= count) {
        g_create_err = "device ordinal out of range";
        return BC_EINVAL;
    }
    bc_ctx* ctx = new bc_ctx();
    ctx->device = device;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    memset(ctx->combo, 0, sizeof ctx->combo);
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        cudaStreamCreate(&ctx->stream) != cudaSuccess || cudaEventCreate(&ctx->ev0) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_i0) != cudaSuccess || cudaEventCreate(&ctx->ev_i1) != cudaSuccess ||
        cudaEventCreate(&ctx->ev1) != cudaSuccess || cudaEventCreate(&ctx->ev2) != cudaSuccess ||
        cudaEventCreate(&ctx->ev3) != cudaSuccess ||
        cudaMalloc(&ctx->d_count, 8 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&ctx->d_any_n, sizeof(uint32_t)) != cudaSuccess) {
        g_create_err = std::string("CUDA context setup failed: ") + cudaGetErrorString(cudaGetLastError());
        bc_destroy(ctx);  // releases whatever was created before the failing call
        return BC_ECUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return BC_OK;
}

extern "C" void bc_destroy(bc_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    dfree(ctx->d_H); dfree(ctx->d_L); dfree(ctx->d_B); dfree(ctx->d_start_dev); dfree(ctx->d_stage); dfree(ctx->d_coff);
    dfree(ctx->d_qh); dfree(ctx->d_ql); dfree(ctx->d_sn); dfree(ctx->d_any_n);
    dfree(ctx->d_dir); dfree(ctx->d_cursor); dfree(ctx->d_scan_tmp); dfree(ctx->d_ent_id); dfree(ctx->d_ent_hl);
    dfree(ctx->d_ent_tmp); dfree(ctx->d_coarse_cursor);
    dfree(ctx->d_pdir);
    dfree(ctx->d_ent_h);
    dfree(ctx->d_hits); dfree(ctx->d_count);
    dfree(ctx->d_sort_scratch); dfree(ctx->d_sort_hist); dfree(ctx->d_sort_tmp); dfree(ctx->d_sort_orand);
    bc_join_free(ctx->join);
    bc_guides_free(ctx->guides);
    if (ctx->ev_i0) cudaEventDestroy(ctx->ev_i0);
    if (ctx->ev_i1) cudaEventDestroy(ctx->ev_i1);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->sink.stream) cudaStreamDestroy(ctx->sink.stream);
    if (ctx->sink.h_counts) cudaFreeHost(ctx->sink.h_counts);
    for (int i = 0; i < BC_SINK_SLICES; i++)
        if (ctx->sink.ev[i]) cudaEventDestroy(ctx->sink.ev[i]);
    delete ctx;
}

extern "C" const char* bc_last_error(bc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

// ---------------------------------------------------------------------------------------- genome
// bc_build_index returns as soon as its kernels are enqueued (the search that follows is enqueued behind them on the
// same stream, so the host never idles between the two: at 8 GPUs a step is ~7 ms and every host round trip shows).
// Its device time is read here, once something has synchronised with the stream anyway.
static int resolve_index_timing(bc_ctx* ctx) {
    if (!ctx->index_timing_pending) return BC_OK;
    ctx->index_timing_pending = false;
    CK(cudaEventSynchronize(ctx->ev_i1));
    CK(cudaEventElapsedTime(&ctx->stats.ms_build_index, ctx->ev_i0, ctx->ev_i1));
    return BC_OK;
}

static int set_genome_common(bc_ctx* ctx, const uint8_t* d_ascii, const uint64_t* contig_offsets,
                             uint32_t n_contigs, cudaStream_t st) {
    { int rc_ = resolve_index_timing(ctx); if (rc_ != BC_OK) return rc_; }  // index kernels may still read the old planes
    uint64_t G = contig_offsets[n_contigs];
    for (uint32_t c = 0; c < n_contigs; c++)
        if (contig_offsets[c + 1] < contig_offsets[c]) return fail(ctx, BC_EINVAL, "contig_offsets must be non-decreasing");
    if (contig_offsets[0] != 0) return fail(ctx, BC_EINVAL, "contig_offsets[0] must be 0");
    if (G + n_contigs + 8192 >= (1ull << 32)) return fail(ctx, BC_ELIMIT, "genome longer than 2^32 - 8192 positions");
    ctx->have_genome = false;
    ctx->have_index = false;  // the seed scheme and the scan path were costed with the previous genome size
    ctx->G = G;
    ctx->n_contigs = n_contigs;
    ctx->coff.assign(contig_offsets, contig_offsets + n_contigs + 1);
    ctx->n_pos = (uint32_t)(G + n_contigs);
    // whole probe tiles (2048 positions = 64 words) plus one halo tile: every kernel may read a
    // full tile and the word after it; the tail words are all-ambiguous padding
    ctx->n_words = ((ctx->n_pos + 2047) / 2048) * 64 + 64 + 4;
    // device buffers are grow-only: repeated bc_set_genome calls (the e2e loop) pay no cudaMalloc/cudaFree
    if (ctx->n_words > ctx->plane_cap) {
        dfree(ctx->d_H); dfree(ctx->d_L); dfree(ctx->d_B);
        ctx->plane_cap = 0;
        CK(cudaMalloc(&ctx->d_H, (size_t)ctx->n_words * 4));
        CK(cudaMalloc(&ctx->d_L, (size_t)ctx->n_words * 4));
        CK(cudaMalloc(&ctx->d_B, (size_t)ctx->n_words * 4));
        ctx->plane_cap = ctx->n_words;
    }
    if ((uint64_t)n_contigs + 1 > ctx->start_cap) {
        dfree(ctx->d_start_dev); dfree(ctx->d_coff);
        ctx->start_cap = 0;
        CK(cudaMalloc(&ctx->d_start_dev, (size_t)(n_contigs + 1) * 4));
        CK(cudaMalloc(&ctx->d_coff, (size_t)(n_contigs + 1) * 8));
        ctx->start_cap = (uint64_t)n_contigs + 1;
    }
    std::vector<uint32_t> start_dev(n_contigs + 1);
    for (uint32_t c = 0; c <= n_contigs; c++) start_dev[c] = (uint32_t)(contig_offsets[c] + c);
    uint64_t* d_coff = ctx->d_coff;
    CK(cudaMemcpyAsync(d_coff, contig_offsets, (size_t)(n_contigs + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_start_dev, start_dev.data(), (size_t)(n_contigs + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev0, st));
    CK(bc_launch_pack_genome(d_ascii, d_coff, ctx->d_start_dev, n_contigs, ctx->n_pos, ctx->n_words, ctx->d_H,
                             ctx->d_L, ctx->d_B, ctx->sm_count, st));
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaStreamSynchronize(st));  // start_dev / d_coff staging must outlive the kernel
    CK(cudaEventElapsedTime(&ctx->stats.ms_pack_genome, ctx->ev0, ctx->ev1));
    ctx->have_genome = true;
    ctx->stats.genome_bases = G;
    return BC_OK;
}

extern "C" int bc_set_genome_dev(bc_ctx* ctx, const uint8_t* d_ascii, const uint64_t* contig_offsets,
                                 uint32_t n_contigs, void* stream) {
    if (!ctx) return BC_EINVAL;
    if (!contig_offsets || n_contigs == 0) return fail(ctx, BC_EINVAL, "genome needs at least one contig");
    if (!d_ascii && contig_offsets[n_contigs] > 0) return fail(ctx, BC_EINVAL, "null genome pointer");
    CK(cudaSetDevice(ctx->device));
    return set_genome_common(ctx, d_ascii, contig_offsets, n_contigs, stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int bc_set_genome(bc_ctx* ctx, const uint8_t* ascii, const uint64_t* contig_offsets, uint32_t n_contigs) {
    if (!ctx) return BC_EINVAL;
    if (!contig_offsets || n_contigs == 0) return fail(ctx, BC_EINVAL, "genome needs at least one contig");
    uint64_t G = contig_offsets[n_contigs];
    if (!ascii && G > 0) return fail(ctx, BC_EINVAL, "null genome pointer");
    CK(cudaSetDevice(ctx->device));
    if (G + 1 > ctx->stage_cap) {
        dfree(ctx->d_stage);
        ctx->stage_cap = 0;
        CK(cudaMalloc(&ctx->d_stage, G + 1));
        ctx->stage_cap = G + 1;
    }
    if (G) CK(cudaMemcpyAsync(ctx->d_stage, ascii, G, cudaMemcpyHostToDevice, ctx->stream));
    return set_genome_common(ctx, ctx->d_stage, contig_offsets, n_contigs, ctx->stream);
}

// --------------------------------------------------------------------------------------- library
static int set_library_common(bc_ctx* ctx, const uint8_t* d_ascii, uint32_t n, uint32_t L, cudaStream_t st) {
    { int rc_ = resolve_index_timing(ctx); if (rc_ != BC_OK) return rc_; }  // index kernels may still read the old library
    if (L < 1 || L > 32) return fail(ctx, BC_ELIMIT, "spacer length must be 1..32");
    if (n >= (1u << 31)) return fail(ctx, BC_ELIMIT, "too many spacers");
    ctx->have_library = false;
    ctx->have_index = false;
    ctx->n = n;
    ctx->L = L;
    if ((uint64_t)n + 1 > ctx->lib_cap) {
        dfree(ctx->d_qh); dfree(ctx->d_ql); dfree(ctx->d_sn);
        ctx->lib_cap = 0;
        size_t ne = (size_t)2 * n + 2;
        CK(cudaMalloc(&ctx->d_qh, ne * 4));
        CK(cudaMalloc(&ctx->d_ql, ne * 4));
        CK(cudaMalloc(&ctx->d_sn, ((size_t)n + 1) * 4));
        ctx->lib_cap = (uint64_t)n + 1;
    }
    CK(cudaMemsetAsync(ctx->d_any_n, 0, 4, st));
    CK(cudaEventRecord(ctx->ev0, st));
    CK(bc_launch_pack_library(d_ascii, n, L, ctx->d_qh, ctx->d_ql, ctx->d_sn, ctx->d_any_n, st));
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaMemcpyAsync(&ctx->lib_has_n, ctx->d_any_n, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ctx->stats.ms_pack_library, ctx->ev0, ctx->ev1));
    ctx->have_library = true;
    ctx->stats.library_spacers = n;
    ctx->stats.spacer_len = L;
    return BC_OK;
}

extern "C" int bc_set_library_dev(bc_ctx* ctx, const uint8_t* d_ascii, uint32_t n, uint32_t L, void* stream) {
    if (!ctx) return BC_EINVAL;
    if (!d_ascii && n > 0) return fail(ctx, BC_EINVAL, "null library pointer");
    CK(cudaSetDevice(ctx->device));
    return set_library_common(ctx, d_ascii, n, L, stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int bc_set_library(bc_ctx* ctx, const uint8_t* ascii, uint32_t n, uint32_t L) {
    if (!ctx) return BC_EINVAL;
    if (!ascii && n > 0) return fail(ctx, BC_EINVAL, "null library pointer");
    if (L < 1 || L > 32) return fail(ctx, BC_ELIMIT, "spacer length must be 1..32");
    CK(cudaSetDevice(ctx->device));
    size_t bytes = (size_t)n * L;
    if (bytes + 1 > ctx->stage_cap) {
        dfree(ctx->d_stage);
        ctx->stage_cap = 0;
        CK(cudaMalloc(&ctx->d_stage, bytes + 1));
        ctx->stage_cap = bytes + 1;
    }
    if (bytes) CK(cudaMemcpyAsync(ctx->d_stage, ascii, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return set_library_common(ctx, ctx->d_stage, n, L, ctx->stream);
}

// ------------------------------------------------------------------------------------------- PAM
static uint32_t iupac_set(char ch, uint32_t flags) {
    switch (ch) {
        case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8; case 'N': return 15;
        default: break;
    }
    if (!(flags & BC_PAM_IUPAC)) return 0;  // literal letter: cannot equal an A/C/G/T base
    switch (ch) {
        case 'R': return 5; case 'Y': return 10; case 'S': return 6; case 'W': return 9; case 'K': return 12;
        case 'M': return 3; case 'B': return 14; case 'D': return 13; case 'H': return 11; case 'V': return 7;
        default: return 0;
    }
}

extern "C" int bc_set_pam(bc_ctx* ctx, const char* pam, int direction, uint32_t flags) {
    if (!ctx) return BC_EINVAL;
    if (!pam) pam = "";
    size_t P = strlen(pam);
    if (P > 8) return fail(ctx, BC_ELIMIT, "PAM longer than 8 letters");
    if (direction != 0 && direction != 1) return fail(ctx, BC_EINVAL, "direction must be 0 (downstream) or 1 (upstream)");
    ctx->P = (uint32_t)P;
    ctx->pam_dir = (uint32_t)direction;
    ctx->pam_flags = flags;
    for (size_t i = 0; i < 8; i++) ctx->pam_sets[i] = i < P ? iupac_set(pam[i], flags) : 0;
    return BC_OK;
}

extern "C" int bc_set_param(bc_ctx* ctx, int key, int64_t value) {
    if (!ctx) return BC_EINVAL;
    switch (key) {
        case BC_PARAM_BLOCKS: ctx->par_blocks = value; ctx->have_index = false; return BC_OK;
        case BC_PARAM_PATH:
            if (value < 0 || value > 3) return fail(ctx, BC_EINVAL, "path must be 0, 1, 2 or 3");
            ctx->par_path = value; ctx->have_index = false; return BC_OK;
        case BC_PARAM_COUNT_CANDIDATES: ctx->par_count = value; return BC_OK;
        case BC_PARAM_HIT_CAPACITY: ctx->par_hit_cap = value; return BC_OK;
        case BC_PARAM_SPACER_ID_BASE: ctx->par_id_base = value; return BC_OK;
        case BC_PARAM_SCAN_PART: {
            const int64_t rank = value & 0xffff, world = (value >> 16) & 0xffff;
            if (world < 1 || rank >= world) return fail(ctx, BC_EINVAL, "scan part must be rank | world << 16 with rank < world");
            ctx->par_scan_rank = rank; ctx->par_scan_world = world; return BC_OK;
        }
        case BC_PARAM_WINDOW_SORT:
            if (value < 0 || value > 2) return fail(ctx, BC_EINVAL, "window sort must be 0, 1 or 2");
            ctx->par_window_sort = value; return BC_OK;
        case BC_PARAM_SLOT_PART: {
            const int64_t rank = value & 0xffff, world = (value >> 16) & 0xffff;
            if (world < 1 || rank >= world) return fail(ctx, BC_EINVAL, "slot part must be rank | world << 16 with rank < world");
            ctx->par_slot_rank = rank; ctx->par_slot_world = world; ctx->have_index = false; return BC_OK;
        }
        case BC_PARAM_COMPACT_DIR:
            if (value < 0 || value > 1) return fail(ctx, BC_EINVAL, "packed directory must be 0 (auto) or 1 (off)");
            ctx->par_compact_dir = value; ctx->have_index = false; return BC_OK;
        case BC_PARAM_KEY_CAP:
            if (value < 0 || value > BC_KEY_MAX_NT) return fail(ctx, BC_EINVAL, "key cap must be 0..12");
            ctx->par_key_cap = value; ctx->have_index = false; return BC_OK;
        case BC_PARAM_INDEX_SORT:
            if (value < 0 || value > 2) return fail(ctx, BC_EINVAL, "index sort must be 0, 1 or 2");
            ctx->par_index_sort = value; ctx->have_index = false; return BC_OK;
        case BC_PARAM_KEY_NT:
            if (value < 0 || value > BC_KEY_MAX_NT) return fail(ctx, BC_EINVAL, "key length must be 0..12");
            ctx->par_key_nt = value; ctx->have_index = false; return BC_OK;
        case BC_PARAM_JOIN_CHUNK:
            if (value < 0) return fail(ctx, BC_EINVAL, "join chunk must be >= 0");
            ctx->par_join_chunk = value; return BC_OK;
        default: return fail(ctx, BC_EINVAL, "unknown parameter");
    }
}

// ----------------------------------------------------------------------------------- seed scheme
// A seed scheme is a family of position masks (one per seed combination) such that every set of
// k mismatch positions is avoided by at least one mask; the library is indexed, and the genome
// windows are probed or sorted, under each mask.  Two families are offered:
//   block schemes   generalised pigeonhole: the L query positions are split into b blocks and
//                   every (b-k)-subset of blocks is a mask (b = k+1 is the classic k+1-seed filter);
//   designs         covering designs found offline (tools/make_designs.py -> bc_designs.inc): fewer
//                   masks for the same key length (L=20, k=3: 15 masks of 10 nt against b=6's 20
//                   masks of 9..11 nt; 12 masks of 9 nt), so fewer records to sort and fewer pairs.
struct BcDesignRow {
    uint8_t L, k, key_nt, n_masks;
    uint16_t first;
};
#include "bc_designs.inc"

struct Scheme {
    uint32_t b, n_combos, key_nt_min, key_nt_max;
    uint32_t block_mask[BC_MAX_BLOCKS];
    ComboDesc combo[BC_MAX_COMBOS];
    uint64_t dir_slots;
    double cand_per_window;  // expected candidates per genome window on uniform data
    bool compact_ok;         // every combination fits the 8-byte record of the compact join path
    uint32_t n_bins;         // compact join path: pass-A bins over all combinations
};

// pieces (runs) of the key mask and of its complement, directory offsets, compact-path bit split
static bool finish_scheme(Scheme* s, uint32_t L, uint64_t n_entries) {
    uint64_t slots = 0;
    double cand = 0;
    uint32_t bins = 0;
    s->compact_ok = true;
    s->key_nt_min = 99;
    s->key_nt_max = 0;
    const uint32_t lm = L >= 32 ? 0xffffffffu : ((1u << L) - 1u);
    for (uint32_t c = 0; c < s->n_combos; c++) {
        ComboDesc& cd = s->combo[c];
        const uint32_t km = cd.key_mask & lm;
        cd.key_mask = km;
        uint32_t np = 0, nr = 0, knt = 0;
        for (uint32_t pos = 0; pos < L;) {
            const bool in_key = (km >> pos) & 1u;
            uint32_t end = pos;
            while (end < L && (((km >> end) & 1u) != 0) == in_key) end++;
            if (in_key) {
                if (np == BC_MAX_PIECES) return false;
                cd.start[np] = (uint8_t)pos;
                cd.len[np] = (uint8_t)(end - pos);
                knt += end - pos;
                np++;
            } else {
                if (nr == BC_MAX_PIECES + 1) return false;
                cd.rstart[nr] = (uint8_t)pos;
                cd.rlen[nr] = (uint8_t)(end - pos);
                nr++;
            }
            pos = end;
        }
        if (knt == 0 || knt > BC_KEY_MAX_NT) return false;
        cd.n_pieces = (uint8_t)np;
        cd.key_nt = (uint8_t)knt;
        cd.n_rem = (uint8_t)nr;
        cd.rem_nt = (uint8_t)(L - knt);
        if (knt < s->key_nt_min) s->key_nt_min = knt;
        if (knt > s->key_nt_max) s->key_nt_max = knt;
        if (slots >= (1ull << 32)) return false;
        cd.dir_off = (uint32_t)slots;
        slots += 1ull << (2 * knt);
        cand += (double)n_entries / (double)(1ull << (2 * knt));
        // compact join path: record = {dev position, x}; x = low key bits | rem planes (2 * rem_nt bits).
        // The key's top bits select the pass-A bin (<= 1024 bins per combination), the low bits the
        // sub-slot inside the bin (<= 4096); both passes like a balanced split.
        const uint32_t kb = 2 * knt, rem2 = 2 * (L - knt);
        uint32_t top = (kb + 1) / 2;
        if (kb <= 10) top = kb;
        while (top < kb && top < 10 && (kb - top) + rem2 > 32) top++;
        if (top > 10) top = 10;
        const uint32_t low = kb - top;
        if (low > 12 || low + rem2 > 32 || L - knt > 16 || L > 22) s->compact_ok = false;  // (22: two 11-bit permutation tables)
        cd.top_bits = (uint8_t)top;
        cd.bin_off = bins;
        bins += 1u << top;
    }
    s->n_bins = bins;
    s->dir_slots = slots + 1;  // +1: end sentinel of the last bucket
    s->cand_per_window = cand;
    return slots + 1 < (1ull << 32);
}

static bool make_block_scheme(uint32_t L, uint32_t k, uint32_t b, uint32_t key_cap_nt, uint64_t n_entries, Scheme* s) {
    if (b < k + 1 || b > BC_MAX_BLOCKS || b > L) return false;
    uint32_t pick = b - k;
    memset(s, 0, sizeof *s);
    s->b = b;
    uint32_t bstart[BC_MAX_BLOCKS + 1];
    for (uint32_t j = 0; j <= b; j++) bstart[j] = (uint32_t)((uint64_t)j * L / b);
    for (uint32_t j = 0; j < b; j++) {
        uint32_t len = bstart[j + 1] - bstart[j];
        s->block_mask[j] = (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << bstart[j];
    }
    uint32_t nc = 0;
    for (uint32_t mask = 0; mask < (1u << b); mask++) {  // ascending mask = deterministic combo order
        if ((uint32_t)__builtin_popcount(mask) != pick) continue;
        if (nc == BC_MAX_COMBOS) return false;
        ComboDesc& cd = s->combo[nc];
        cd.blocks_mask = mask;
        uint32_t budget = key_cap_nt;
        for (uint32_t j = 0; j < b && budget; j++) {  // the key = the first key_cap_nt bases of the chosen blocks
            if (!(mask >> j & 1u)) continue;
            uint32_t len = bstart[j + 1] - bstart[j];
            if (len > budget) len = budget;
            cd.key_mask |= (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << bstart[j];
            budget -= len;
        }
        nc++;
    }
    s->n_combos = nc;
    return finish_scheme(s, L, n_entries);
}

static bool make_design_scheme(uint32_t L, const BcDesignRow& row, uint64_t n_entries, Scheme* s) {
    if (row.n_masks > BC_MAX_COMBOS) return false;
    memset(s, 0, sizeof *s);
    s->b = 0;
    s->n_combos = row.n_masks;
    for (uint32_t c = 0; c < row.n_masks; c++) s->combo[c].key_mask = bc_design_masks[row.first + c];
    return finish_scheme(s, L, n_entries);
}

// Cost model in picoseconds of B200 time per unit, fitted to measurements (DESIGN.md section 4;
// bench_kernels/scatter_lab.cu and sort_lab.cu for the sort terms):
//   probe path   : a directory probe costs ~17 ps while the directories stay L2-resident and ~58 ps
//                  once they live in HBM; a candidate is an uncoalesced 8-byte read, ~3 ps;
//   join path    : 16-byte window records.  The sort costs ~20 ps per (window, combination) record
//                  when the radix scatter applies (keys of 4..8 nt), else ~28 ps with <= 2^16 slots
//                  per combination and ~6 ps more per doubling beyond that; verify ~0.25 ps per
//                  candidate plus ~3 ps per record in well-filled slots or ~20 ps per record otherwise;
//   compact join : 8-byte window records (L <= 20 or so), two shared-memory radix passes for keys up
//                  to 11 nt: ~5.3 ps (count) + ~4 ps (pass A) + ~8 ps (pass B) per record, verify as above;
//   all          : the library index costs ~32 ps per (entry, combination); directories are cleared,
//                  scanned and copied at ~5 ps per slot; fixed launch floor ~100 us.
static double scheme_cost(const bc_ctx* ctx, const Scheme& s, uint32_t path) {
    const uint64_t E = 2ull * ctx->n;
    const double windows = (double)(ctx->G ? ctx->G : 1);
    const double dir_bytes = 4.0 * (double)s.dir_slots;
    const double records = windows * s.n_combos;
    const double entries = (double)E * s.n_combos;
    const double cands = windows * s.cand_per_window;
    const double common = 32.0 * entries + 5.0 * (double)s.dir_slots;
    double slots_per_combo = (double)s.dir_slots / s.n_combos;
    const double c_rec = windows / slots_per_combo >= 64.0 ? 3.0 : 20.0;  // windows per slot: dense or sparse verify
    if (path == 1) {
        const double c_probe = dir_bytes < 100e6 ? 17.0 : 58.0;
        return c_probe * records + 3.0 * cands + common;
    }
    if (path == 2) {
        bool radix = true;
        for (uint32_t c = 0; c < s.n_combos; c++) radix = radix && s.combo[c].key_nt >= 4 && s.combo[c].key_nt <= 8;
        double c_sort = radix ? 20.0 : 28.0;
        while (!radix && slots_per_combo > 65536.0) { c_sort += 6.0; slots_per_combo *= 0.5; }
        return (c_sort + c_rec) * records + 0.25 * cands + common + 5.0 * (double)s.dir_slots + 1.0e8;
    }
    // compact join (measured at cfg 4, 9- and 10-nt designs, round-2 kernels): sort 14 ps per record (bin count 1.3,
    // pass A 6.1, slot count 1.3, pass B 5.2), index 18 ps per (entry, combination) with the radix builder, verify +
    // finish 0.25 ps per candidate plus ~5 ps per record of tile overhead when the slots are small (< 256 windows)
    const double c_sort = s.key_nt_max <= 5 ? 9.0 : 14.2;
    const double c_tile = windows / slots_per_combo >= 256.0 ? 1.0 : 5.0;
    return (c_sort + c_tile) * records + 0.25 * cands + 18.0 * entries + 10.0 * (double)s.dir_slots + 1.2e8;
}

static int choose_scheme(bc_ctx* ctx, uint32_t k, Scheme* best, uint32_t* path_out) {
    const uint64_t E = 2ull * ctx->n;
    uint32_t cap = 4;
    while (cap < BC_KEY_MAX_NT && (1ull << (2 * cap)) < 16 * (E ? E : 1)) cap++;
    if (ctx->par_key_cap && cap > (uint32_t)ctx->par_key_cap) cap = (uint32_t)ctx->par_key_cap;
    double best_cost = 0;
    bool found = false;
    uint32_t best_path = 1;
    auto consider = [&](const Scheme& s) {
        for (uint32_t path = 1; path <= 3; path++) {
            if (ctx->par_path && (uint32_t)ctx->par_path != path) continue;
            if (path == 3 && !s.compact_ok) continue;
            const double cost = scheme_cost(ctx, s, path);
            if (!found || cost < best_cost) {
                found = true;
                best_cost = cost;
                *best = s;
                best_path = path;
            }
        }
    };
    Scheme s;
    if (!ctx->par_key_nt) {
        for (uint32_t b = k + 1; b <= k + 4 && b <= BC_MAX_BLOCKS; b++) {
            if (ctx->par_blocks && (uint32_t)ctx->par_blocks != b) continue;
            if (make_block_scheme(ctx->L, k, b, cap, E, &s)) consider(s);
        }
    }
    if (!ctx->par_blocks) {
        for (size_t r = 0; r < sizeof bc_design_rows / sizeof bc_design_rows[0]; r++) {
            const BcDesignRow& row = bc_design_rows[r];
            if (row.L != ctx->L || row.k != k) continue;
            if (ctx->par_key_nt ? (uint32_t)ctx->par_key_nt != row.key_nt : row.key_nt > cap) continue;
            if (make_design_scheme(ctx->L, row, E, &s)) consider(s);
        }
    }
    if (!found) return fail(ctx, BC_EINVAL, "no feasible seed scheme for this (L, k, blocks / key length, path)");
    *path_out = best_path;
    return BC_OK;
}

extern "C" int bc_build_index(bc_ctx* ctx, int k) {
    if (!ctx) return BC_EINVAL;
    if (!ctx->have_library) return fail(ctx, BC_EINVAL, "bc_build_index: no library loaded");
    if (!ctx->have_genome) return fail(ctx, BC_EINVAL, "bc_build_index: no genome loaded");
    if (k < 0 || k > 3) return fail(ctx, BC_ELIMIT, "k must be 0..3 (bowtie -v limit)");
    CK(cudaSetDevice(ctx->device));
    { int rc_ = resolve_index_timing(ctx); if (rc_ != BC_OK) return rc_; }
    ctx->have_index = false;
    ctx->index_k = k;
    ctx->stats.k = (uint32_t)k;
    if ((uint32_t)k >= ctx->L || ctx->n == 0) {  // reads with L <= k are never aligned (oracle.c rule 7)
        ctx->n_combos = 0;
        ctx->b = 0;
        ctx->have_index = true;
        ctx->stats.ms_build_index = 0;
        return BC_OK;
    }
    Scheme s;
    uint32_t path = 1;
    int rc = choose_scheme(ctx, (uint32_t)k, &s, &path);
    if (rc != BC_OK) return rc;
    ctx->b = s.b;
    ctx->n_combos = s.n_combos;
    memcpy(ctx->block_mask, s.block_mask, sizeof s.block_mask);
    memcpy(ctx->combo, s.combo, sizeof s.combo);
    ctx->dir_slots = s.dir_slots;
    ctx->n_bins = s.n_bins;
    ctx->stats.blocks = s.b;
    ctx->stats.combos = s.n_combos;
    ctx->stats.path = path;
    ctx->stats.key_nt = s.key_nt_max;

    const uint64_t E = 2ull * ctx->n;
    const uint64_t ent_needed = E * s.n_combos;
    if (ent_needed >= (1ull << 32)) return fail(ctx, BC_ELIMIT, "library x combinations exceeds 2^32 index entries");
    if (ent_needed > ctx->ent_cap) {
        dfree(ctx->d_ent_hl); dfree(ctx->d_ent_id); dfree(ctx->d_ent_tmp);
        ctx->ent_cap = 0;
        CK(cudaMalloc(&ctx->d_ent_hl, (ent_needed + 1) * sizeof(uint2)));
        CK(cudaMalloc(&ctx->d_ent_id, (ent_needed + 1) * sizeof(uint32_t)));
        CK(cudaMalloc(&ctx->d_ent_tmp, (ent_needed + 1) * sizeof(uint4)));
        ctx->ent_cap = ent_needed;
    }
    if (s.dir_slots > ctx->dir_cap) {
        dfree(ctx->d_dir); dfree(ctx->d_cursor);
        ctx->dir_cap = 0;
        CK(cudaMalloc(&ctx->d_dir, s.dir_slots * 4));
        CK(cudaMalloc(&ctx->d_cursor, s.dir_slots * 4));
        ctx->dir_cap = s.dir_slots;
    }
    uint64_t tmp_words = bc_scan_tmp_words(s.dir_slots);
    if (tmp_words > ctx->scan_tmp_cap) {
        dfree(ctx->d_scan_tmp);
        ctx->scan_tmp_cap = 0;
        CK(cudaMalloc(&ctx->d_scan_tmp, tmp_words * 4));
        ctx->scan_tmp_cap = tmp_words;
    }
    IndexParams ip;
    ip.qh = ctx->d_qh; ip.ql = ctx->d_ql; ip.sn = ctx->d_sn;
    ip.n_entries = (uint32_t)E;
    ip.L = ctx->L;
    ip.lib_has_n = ctx->lib_has_n;
    // slot-range sharding: shard r owns [bound(r), bound(r + 1)).  On the compact join path the bounds are moved down to
    // a pass-A bin boundary (a multiple of 2^low slots inside their combination; a bin is ~1/15,000 of the directory), so
    // that "is this record mine" is a test on the bin and the kernels keep their one-plane fast paths in combinations a
    // shard owns only partly - at 8 shards over 15 combinations that is every shard.
    auto bound = [&](uint64_t r) -> uint32_t {
        if (r == 0) return 0u;
        if (r >= (uint64_t)ctx->par_slot_world) return (uint32_t)(s.dir_slots - 1);
        uint32_t raw = (uint32_t)((s.dir_slots - 1) * r / (uint64_t)ctx->par_slot_world);
        if (path == 3) {
            uint32_t c = 0;
            while (c + 1 < s.n_combos && s.combo[c + 1].dir_off <= raw) c++;
            const uint32_t low = 2u * s.combo[c].key_nt - s.combo[c].top_bits;
            raw = s.combo[c].dir_off + (((raw - s.combo[c].dir_off) >> low) << low);
        }
        return raw;
    };
    ip.slot_lo = bound((uint64_t)ctx->par_slot_rank);
    ip.slot_hi = bound((uint64_t)ctx->par_slot_rank + 1);
    ip.bin_aligned = path == 3 ? 1u : 0u;
    ctx->slot_lo = ip.slot_lo; ctx->slot_hi = ip.slot_hi;
    ip.compact = path == 3 ? 1u : 0u;  // compact join: the index stores the non-key (rem) planes of every entry
    memcpy(ip.combo, ctx->combo, sizeof ip.combo);
    const uint32_t launches0 = bc_launch_counter;
    CK(cudaEventRecord(ctx->ev_i0, ctx->stream));
    if (!ctx->d_coarse_cursor) CK(cudaMalloc(&ctx->d_coarse_cursor, ((size_t)BC_MAX_COMBOS << BC_COARSE_BITS) * 4 + 4));
    if (ip.compact && ctx->par_index_sort != 1)  // the two radix passes of the compact window sort, applied to the entries
        CK(bc_cindex_build(ctx->join, ip, s.n_combos, s.n_bins, ctx->d_dir, s.dir_slots, ctx->d_cursor, ctx->d_scan_tmp,
                           reinterpret_cast<uint2*>(ctx->d_ent_tmp), ctx->d_ent_hl, ctx->d_ent_id, ctx->sm_count, ctx->stream));
    else
        CK(bc_launch_index_build(ip, s.n_combos, ctx->d_dir, s.dir_slots, ctx->d_cursor, ctx->d_scan_tmp, ctx->d_ent_tmp,
                                 ctx->d_coarse_cursor, ctx->d_ent_hl, ctx->d_ent_id, ctx->sm_count, ctx->stream));
    ctx->packed_dir = false;
    if (path == 1 && ctx->par_compact_dir != 1 && ent_needed < (1ull << 26)) {
        // probe path: one 4-byte load per directory probe instead of two
        if (s.dir_slots > ctx->pdir_cap) {
            dfree(ctx->d_pdir);
            ctx->pdir_cap = 0;
            CK(cudaMalloc(&ctx->d_pdir, s.dir_slots * sizeof(uint32_t)));
            ctx->pdir_cap = s.dir_slots;
        }
        CK(bc_launch_dir_pack(ctx->d_dir, (uint32_t)(s.dir_slots - 1), ctx->d_pdir, ctx->sm_count, ctx->stream));
        ctx->packed_dir = true;
    }
    ctx->have_ent_h = false;
    if (path == 1 && ctx->par_compact_dir != 1) {
        // probe path: H planes of the entries as a dense array (tested first; halves the L2 working set of the entries)
        if (ent_needed + 4 > ctx->ent_h_cap) {
            dfree(ctx->d_ent_h);
            ctx->ent_h_cap = 0;
            CK(cudaMalloc(&ctx->d_ent_h, (ent_needed + 4) * sizeof(uint32_t)));
            ctx->ent_h_cap = ent_needed + 4;
        }
        CK(cudaMemsetAsync(ctx->d_ent_h + ent_needed, 0, 4 * sizeof(uint32_t), ctx->stream));
        CK(bc_launch_ent_h_pack(ctx->d_ent_hl, ent_needed, ctx->d_ent_h, ctx->sm_count, ctx->stream));
        ctx->have_ent_h = true;
    }
    CK(cudaEventRecord(ctx->ev_i1, ctx->stream));
    ctx->index_timing_pending = true;  // no host synchronisation here: see resolve_index_timing
    ctx->stats.index_launches = bc_launch_counter - launches0;
    ctx->have_index = true;
    return BC_OK;
}

// ---------------------------------------------------------------------------------------- search
static void fill_params(bc_ctx* ctx, SearchParams* p) {
    memset(p, 0, sizeof *p);
    p->H = ctx->d_H; p->Lo = ctx->d_L; p->B = ctx->d_B;
    p->start_dev = ctx->d_start_dev;
    p->n_pos = ctx->n_pos;
    p->n_words = ctx->n_words;
    p->pos_begin = (uint32_t)((uint64_t)ctx->n_pos * (uint64_t)ctx->par_scan_rank / (uint64_t)ctx->par_scan_world);
    p->pos_end = (uint32_t)((uint64_t)ctx->n_pos * (uint64_t)(ctx->par_scan_rank + 1) / (uint64_t)ctx->par_scan_world);
    p->n_contigs = ctx->n_contigs;
    p->slot_lo = ctx->slot_lo; p->slot_hi = ctx->slot_hi;
    p->bin_aligned = ctx->stats.path == 3 ? 1u : 0u;
    p->own_hash = ctx->par_slot_world > 1 ? 1u : 0u;  // balanced record counts across the slot-range shards
    p->sn = ctx->d_sn;
    p->lib_has_n = ctx->lib_has_n;
    p->L = ctx->L;
    p->k = (uint32_t)ctx->index_k;
    p->b = ctx->b;
    p->n_combos = ctx->n_combos;
    memcpy(p->block_mask, ctx->block_mask, sizeof p->block_mask);
    memcpy(p->combo, ctx->combo, sizeof p->combo);
    p->dir = ctx->d_dir;
    p->ent_hl = ctx->d_ent_hl;
    p->ent_id = ctx->d_ent_id;
    p->dir_entries = 2ull * ctx->n * ctx->n_combos;
    p->pdir = ctx->packed_dir ? ctx->d_pdir : nullptr;
    p->ent_h = ctx->have_ent_h ? ctx->d_ent_h : nullptr;
    p->P = ctx->P;
    p->pam_dir = ctx->pam_dir;
    p->pam_flags = ctx->pam_flags;
    memcpy(p->pam_sets, ctx->pam_sets, sizeof p->pam_sets);
    p->gate_first = ((ctx->pam_flags & BC_PAM_GATE) && ctx->P > 0) ? 1u : 0u;
    p->window_sort = (uint32_t)ctx->par_window_sort;
    p->join_chunk = (uint32_t)(ctx->par_join_chunk > 0xffffffffll ? 0xffffffffll : ctx->par_join_chunk);
    p->hits = ctx->d_hits;
    p->count = ctx->d_count;
    p->cap = ctx->hit_cap;
    p->count_candidates = ctx->par_count ? 1u : 0u;
    p->spacer_id_base = (uint32_t)ctx->par_id_base;
}

extern "C" int bc_search(bc_ctx* ctx, int k, uint64_t* n_hits_out) {
    if (!ctx) return BC_EINVAL;
    if (n_hits_out) *n_hits_out = 0;
    if (!ctx->have_library || !ctx->have_genome) return fail(ctx, BC_EINVAL, "bc_search: genome and library must be loaded first");
    if (!ctx->have_index || ctx->index_k != k) {
        int rc = bc_build_index(ctx, k);
        if (rc != BC_OK) return rc;
    }
    CK(cudaSetDevice(ctx->device));
    ctx->n_hits = 0;
    ctx->stats.hits = ctx->stats.candidates = ctx->stats.probes = 0;
    ctx->stats.ms_search = ctx->stats.ms_scan_kernel = ctx->stats.ms_genome_bucket = 0;
    ctx->stats.scan_launches = 0;
    if (ctx->n_combos == 0 || ctx->G < ctx->L) return BC_OK;

    uint64_t want = ctx->par_hit_cap > 0 ? (uint64_t)ctx->par_hit_cap : 8ull * ctx->n;
    if (ctx->par_hit_cap <= 0) {
        if (want < (1ull << 20)) want = 1ull << 20;
        if (want > (1ull << 28)) want = 1ull << 28;
    }
    if (ctx->hit_cap < want) {
        dfree(ctx->d_hits);
        ctx->hit_cap = 0;
        CK(cudaMalloc(&ctx->d_hits, want * sizeof(bc_hit)));
        ctx->hit_cap = want;
    }
    float ms_total = 0, ms_scan = 0, ms_bucket = 0;
    ctx->stats.search_attempts = 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        ctx->stats.search_attempts++;
        SearchParams p;
        fill_params(ctx, &p);
        CK(cudaMemsetAsync(ctx->d_count, 0, 8 * sizeof(unsigned long long), ctx->stream));
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        uint32_t launches = 0;
        ctx->sink.copied = ctx->sink.reported = 0;
        if (ctx->stats.path == 2) {
            CK(bc_join_search(ctx->join, p, ctx->dir_slots, ctx->sm_count, ctx->stream, &launches,
                              (ctx->sink.host || ctx->sink.fn) ? &ctx->sink : nullptr));
        } else if (ctx->stats.path == 3) {
            CK(bc_cjoin_search(ctx->join, p, ctx->dir_slots, ctx->n_bins, ctx->sm_count, ctx->stream, &launches,
                               (ctx->sink.host || ctx->sink.fn) ? &ctx->sink : nullptr));
        } else {
            CK(cudaEventRecord(ctx->ev2, ctx->stream));
            CK(bc_launch_scan_probe(p, ctx->dir_slots * 4ull, ctx->sm_count, ctx->stream));
            CK(cudaEventRecord(ctx->ev3, ctx->stream));
            launches = 1;
        }
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        unsigned long long counts[6] = {0, 0, 0, 0, 0, 0};
        CK(cudaMemcpyAsync(counts, ctx->d_count, sizeof counts, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->stats.path == 1) bc_probe_release_l2();
        float a = 0, b2 = 0;
        CK(cudaEventElapsedTime(&a, ctx->ev0, ctx->ev1));
        if (ctx->stats.path >= 2) { b2 = ctx->join.ms_join_kernels; ms_bucket += ctx->join.ms_bucket_kernels; }
        else CK(cudaEventElapsedTime(&b2, ctx->ev2, ctx->ev3));
        ms_total += a;
        ms_scan += b2;
        ctx->stats.scan_launches += launches;
        ctx->stats.candidates = counts[1];
        ctx->stats.probes = counts[2];
        if (ctx->stats.path == 3 && counts[5] > ctx->join.item_cap) {
            // compact join: the item queue between k_cverify and k_cfinish overflowed and batches were dropped.
            // The demand is known now: repeat the search with a queue of that size.
            if (ctx->sink.fn) {
                ctx->stats.hits = counts[0];
                return fail(ctx, BC_ELIMIT, "bc_search: verify item queue overflow with a slice callback installed (raise BC_PARAM_HIT_CAPACITY)");
            }
            if (attempt == 2) return fail(ctx, BC_ECUDA, "verify item queue kept overflowing");
            ctx->join.item_want = counts[5] + counts[5] / 8 + 1024;
            continue;
        }
        if (counts[0] <= ctx->hit_cap) {
            ctx->n_hits = counts[0];
            break;
        }
        // The buffer was too small: the kernel kept counting, so the exact size is known now.
        if (ctx->sink.fn) {  // parts of this attempt were already handed over: the caller restarts
            ctx->stats.hits = counts[0];
            return fail(ctx, BC_ELIMIT, "bc_search: hit buffer overflow with a slice callback installed (raise BC_PARAM_HIT_CAPACITY)");
        }
        uint64_t need = counts[0] + counts[0] / 16 + 1024;
        dfree(ctx->d_hits);
        ctx->hit_cap = 0;
        CK(cudaMalloc(&ctx->d_hits, need * sizeof(bc_hit)));
        ctx->hit_cap = need;
        if (attempt == 2) return fail(ctx, BC_ECUDA, "hit buffer kept overflowing");
    }
    ctx->stats.hits = ctx->n_hits;
    { int rc_ = resolve_index_timing(ctx); if (rc_ != BC_OK) return rc_; }  // (already complete: the stream was synchronised)
    ctx->stats.ms_search = ms_total;
    ctx->stats.ms_scan_kernel = ms_scan;
    ctx->stats.ms_genome_bucket = ms_bucket;
    ctx->stats.ms_win_count = ctx->stats.path == 3 ? ctx->join.ms_kernel[0] : 0;
    ctx->stats.ms_win_bin = ctx->stats.path == 3 ? ctx->join.ms_kernel[1] : 0;
    ctx->stats.ms_win_place = ctx->stats.path == 3 ? ctx->join.ms_kernel[2] : 0;
    ctx->stats.ms_finish = ctx->stats.path == 3 ? ctx->join.ms_kernel[5] : 0;
    if (n_hits_out) *n_hits_out = ctx->n_hits;
    if (ctx->sink.fn && ctx->n_hits > ctx->sink.reported) {  // the rest (all of it on the probe path)
        ctx->sink.fn(ctx->sink.fn_user, ctx->d_hits, ctx->sink.reported, ctx->n_hits);
        ctx->sink.reported = ctx->n_hits;
    }
    if (ctx->sink.host) {
        if (ctx->n_hits > ctx->sink.cap) return fail(ctx, BC_ELIMIT, "bc_search: more hits than the hit sink holds (use bc_copy_hits)");
        if (ctx->n_hits > ctx->sink.copied)  // whatever the slices did not deliver yet (all of it on the probe path)
            CK(cudaMemcpyAsync(ctx->sink.host + ctx->sink.copied, ctx->d_hits + ctx->sink.copied,
                               (ctx->n_hits - ctx->sink.copied) * sizeof(bc_hit), cudaMemcpyDefault, ctx->sink.stream));
        CK(cudaStreamSynchronize(ctx->sink.stream));
    }
    return BC_OK;
}

static int sink_resources(bc_ctx* ctx) {
    if (ctx->sink.stream) return BC_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamCreateWithFlags(&ctx->sink.stream, cudaStreamNonBlocking));
    CK(cudaHostAlloc((void**)&ctx->sink.h_counts, BC_SINK_SLICES * sizeof(unsigned long long), cudaHostAllocDefault));
    for (int i = 0; i < BC_SINK_SLICES; i++) CK(cudaEventCreateWithFlags(&ctx->sink.ev[i], cudaEventDisableTiming));
    return BC_OK;
}

extern "C" int bc_set_slice_callback(bc_ctx* ctx, bc_slice_fn fn, void* user) {
    if (!ctx) return BC_EINVAL;
    if (fn) {
        int rc = sink_resources(ctx);
        if (rc != BC_OK) return rc;
    }
    ctx->sink.fn = fn;
    ctx->sink.fn_user = fn ? user : nullptr;
    return BC_OK;
}

extern "C" int bc_set_hit_sink(bc_ctx* ctx, bc_hit* dst, uint64_t cap) {
    if (!ctx) return BC_EINVAL;
    if (dst && cap == 0) return fail(ctx, BC_EINVAL, "bc_set_hit_sink: zero capacity");
    if (dst) {
        int rc = sink_resources(ctx);
        if (rc != BC_OK) return rc;
    }
    ctx->sink.host = dst;
    ctx->sink.cap = dst ? cap : 0;
    ctx->sink.copied = 0;
    ctx->sink.is_device = false;
    if (dst) {  // a device destination (peer merge over NVLink) needs fewer, larger slices than a PCIe copy to the host
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, dst) == cudaSuccess) ctx->sink.is_device = attr.type == cudaMemoryTypeDevice;
        else (void)cudaGetLastError();
    }
    return BC_OK;
}

extern "C" int bc_peer_export(bc_ctx* ctx, uint64_t n_records, bc_hit** d_ptr, unsigned char handle[64]) {
    if (!ctx || !d_ptr || !handle || n_records == 0) return ctx ? fail(ctx, BC_EINVAL, "bc_peer_export: bad argument") : BC_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    CK(cudaSetDevice(ctx->device));
    void* p = nullptr;
    CK(cudaMalloc(&p, n_records * sizeof(bc_hit)));   // a whole allocation of its own: IPC handles name allocations
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        CK(e);
    }
    memcpy(handle, &h, sizeof h);
    *d_ptr = (bc_hit*)p;
    return BC_OK;
}

extern "C" int bc_peer_open(bc_ctx* ctx, const unsigned char handle[64], bc_hit** d_ptr) {
    if (!ctx || !d_ptr || !handle) return ctx ? fail(ctx, BC_EINVAL, "bc_peer_open: bad argument") : BC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = (bc_hit*)p;
    return BC_OK;
}

extern "C" int bc_peer_close(bc_ctx* ctx, bc_hit* d_ptr, int owner) {
    if (!ctx) return BC_EINVAL;
    if (!d_ptr) return BC_OK;
    CK(cudaSetDevice(ctx->device));
    if (owner) CK(cudaFree(d_ptr));
    else CK(cudaIpcCloseMemHandle(d_ptr));
    return BC_OK;
}

extern "C" int bc_copy_hits(bc_ctx* ctx, bc_hit* dst, uint64_t cap) {
    if (!ctx) return BC_EINVAL;
    if (cap < ctx->n_hits) return fail(ctx, BC_EINVAL, "bc_copy_hits: destination too small");
    if (ctx->n_hits == 0) return BC_OK;
    if (!dst) return fail(ctx, BC_EINVAL, "bc_copy_hits: null destination");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst, ctx->d_hits, ctx->n_hits * sizeof(bc_hit), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BC_OK;
}

extern "C" int bc_sort_hits(bc_ctx* ctx, int order) {
    if (!ctx) return BC_EINVAL;
    if (order != 0 && order != 1) return fail(ctx, BC_EINVAL, "bc_sort_hits: order must be 0 (canonical) or 1 (best)");
    if (ctx->n_hits < 2) return BC_OK;
    CK(cudaSetDevice(ctx->device));
    if (ctx->sort_cap < ctx->hit_cap) {
        dfree(ctx->d_sort_scratch);
        ctx->sort_cap = 0;
        CK(cudaMalloc(&ctx->d_sort_scratch, ctx->hit_cap * sizeof(bc_hit)));
        ctx->sort_cap = ctx->hit_cap;
    }
    const uint64_t hw = bc_sort_hist_words(ctx->n_hits);
    if (hw > ctx->sort_hist_cap) {
        dfree(ctx->d_sort_hist);
        ctx->sort_hist_cap = 0;
        CK(cudaMalloc(&ctx->d_sort_hist, hw * sizeof(uint32_t)));
        ctx->sort_hist_cap = hw;
    }
    const uint64_t tw = bc_scan_tmp_words(hw);
    if (tw > ctx->sort_tmp_cap) {
        dfree(ctx->d_sort_tmp);
        ctx->sort_tmp_cap = 0;
        CK(cudaMalloc(&ctx->d_sort_tmp, tw * sizeof(uint32_t)));
        ctx->sort_tmp_cap = tw;
    }
    if (!ctx->d_sort_orand) CK(cudaMalloc(&ctx->d_sort_orand, 8 * sizeof(uint32_t)));
    uint4* result = nullptr;
    uint32_t passes = 0;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    CK(bc_sort_records(reinterpret_cast<uint4*>(ctx->d_hits), reinterpret_cast<uint4*>(ctx->d_sort_scratch), ctx->n_hits, order,
                       ctx->d_sort_hist, ctx->sort_hist_cap, ctx->d_sort_tmp, ctx->d_sort_orand, ctx->sm_count, ctx->stream,
                       &result, &passes));
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->stats.ms_sort_hits, ctx->ev0, ctx->ev1));
    if (result != reinterpret_cast<uint4*>(ctx->d_hits)) {  // the sorted records ended up in the scratch buffer: swap roles
        bc_hit* t = ctx->d_hits;
        ctx->d_hits = ctx->d_sort_scratch;
        ctx->d_sort_scratch = t;
    }
    return BC_OK;
}

extern "C" int bc_hits_device(bc_ctx* ctx, const bc_hit** d_hits, uint64_t* n_hits) {
    if (!ctx || !d_hits || !n_hits) return BC_EINVAL;
    *d_hits = ctx->d_hits;
    *n_hits = ctx->n_hits;
    return BC_OK;
}

extern "C" int bc_get_stats(bc_ctx* ctx, bc_stats* out) {
    if (!ctx || !out) return BC_EINVAL;
    if (ctx->index_timing_pending) {
        cudaSetDevice(ctx->device);
        int rc_ = resolve_index_timing(ctx);
        if (rc_ != BC_OK) return rc_;
    }
    *out = ctx->stats;
    return BC_OK;
}

// -------------------------------------------------------------------------------- guide enumeration
extern "C" int bc_enumerate_guides(bc_ctx* ctx, uint32_t L, const char* pam, int direction, uint32_t flags,
                                   uint64_t* n_guides_out) {
    if (!ctx) return BC_EINVAL;
    if (n_guides_out) *n_guides_out = 0;
    if (!ctx->have_genome) return fail(ctx, BC_EINVAL, "bc_enumerate_guides: no genome loaded");
    if (L < 1 || L > 32) return fail(ctx, BC_ELIMIT, "guide length must be 1..32");
    if (!pam) pam = "";
    const size_t P = strlen(pam);
    if (P > 8) return fail(ctx, BC_ELIMIT, "PAM longer than 8 letters");
    if (direction != 0 && direction != 1) return fail(ctx, BC_EINVAL, "direction must be 0 or 1");
    uint32_t sets[8] = {0};
    for (size_t i = 0; i < P; i++) sets[i] = iupac_set(pam[i], flags);
    CK(cudaSetDevice(ctx->device));
    uint64_t n = 0;
    CK(bc_guides_enumerate(ctx->guides, ctx->d_H, ctx->d_L, ctx->d_B, ctx->d_start_dev, ctx->n_pos, ctx->n_contigs, L,
                           (uint32_t)P, sets, direction, (flags & BC_GUIDES_REFERENCE_RANGE) ? 1 : 0, ctx->sm_count,
                           ctx->stream, &n));
    if (n_guides_out) *n_guides_out = n;
    return BC_OK;
}

extern "C" int bc_copy_guides(bc_ctx* ctx, uint64_t* dst, uint64_t cap) {
    if (!ctx) return BC_EINVAL;
    const uint64_t n = ctx->guides.n_guides;
    if (cap < n) return fail(ctx, BC_EINVAL, "bc_copy_guides: destination too small");
    if (n == 0) return BC_OK;
    if (!dst) return fail(ctx, BC_EINVAL, "bc_copy_guides: null destination");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst, ctx->guides.d_out, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BC_OK;
}
