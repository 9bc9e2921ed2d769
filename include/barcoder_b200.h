/*
 * barcoder_b200.h - C ABI of the B200-native spacer->genome mismatch search.
 *
 * This is the drop-in boundary for ONE path of ryandward/barcoder: the
 * `bowtie -v k -a` call plus the PAM-adjacency check.  Each entry point cites the
 * reference interface (file:line under the reference checkout) it replaces.
 * Plain C: opaque context, raw pointers, sizes, integer return codes.  No torch,
 * no C++ types.  INTEGRATION.md shows the ctypes stub a reference maintainer adds.
 *
 * All functions return BC_OK (0) or a negative BC_E* code; bc_last_error() gives the
 * text.  There is no CPU fallback: without a usable CUDA device bc_create fails.
 * A context is bound to one GPU and must be used from one host thread at a time.
 * Side effects on the calling thread: every entry point makes the context's device current
 * (cudaSetDevice) and leaves it current; the probe kernel raises the device's persisting-L2 limit
 * for the duration of a search and restores the previous value afterwards.
 */
#ifndef BARCODER_B200_H
#define BARCODER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BC_ABI_VERSION 1

#define BC_OK 0
#define BC_EINVAL (-1)   /* bad argument / call order            */
#define BC_ECUDA (-2)    /* CUDA runtime error                   */
#define BC_ENODEV (-3)   /* no usable CUDA device                */
#define BC_ENOMEM (-4)   /* host or device allocation failed     */
#define BC_ELIMIT (-5)   /* size outside the supported envelope  */

typedef struct bc_ctx bc_ctx;

/*
 * One alignment.  Replaces one SAM line as consumed by PySamParser.py:26-48
 * (Chromosome/Start/End/Strand/Barcode/Mismatches) and one parse_sam_output row
 * (targets.py:354-410: spacer, target, mismatches, chr, tar_start, sp_dir, pam, diff).
 */
typedef struct {
    uint32_t spacer_id; /* index into the array given to bc_set_library            */
    uint32_t gpos;      /* 0-based leftmost '+'-strand position in the genome given to
                           bc_set_genome (contigs concatenated, no separators); the
                           contig is the last one whose offset is <= gpos           */
    uint32_t mm_mask;   /* bit i = mismatch at spacer position i (0-based, spacer
                           5'->3' orientation; targets.py:184-190); popcount == NM   */
    uint32_t meta;      /* BC_META_* below                                          */
} bc_hit;

#define BC_META_STRAND(m) ((m) & 1u)             /* 0 '+', 1 '-' (SAM flag 16)           */
#define BC_META_NMM(m) (((m) >> 1) & 3u)         /* mismatches, 0..3 (bowtie -v limit)   */
#define BC_META_PAM_OK (1u << 3)                 /* PAM full, unambiguous and matching   */
#define BC_META_PAM_FULL (1u << 4)               /* all PAM bases lie inside the contig  */
#define BC_META_PAM_AMB (1u << 5)                /* a PAM base is non-ACGT in the genome */
#define BC_META_PAM_LEN(m) (((m) >> 8) & 15u)    /* PAM length P (0..8)                  */
#define BC_META_PAM_CODE(m, i) (((m) >> (16 + 2 * (i))) & 3u) /* base i of the PAM in
                                                    spacer orientation: A0 C1 G2 T3      */

/* bc_set_pam flags */
#define BC_PAM_IUPAC 1u /* expand IUPAC letters (R,Y,...) in the pattern.  Without it only
                           N is a wildcard and other letters are literals, exactly like
                           PAMProcessor.py:7 and targets.py:224.                        */
#define BC_PAM_GATE 2u  /* report only hits whose PAM matches (plus ambiguous-PAM hits,
                           which the host resolves).  Off = bowtie's full hit set.      */

/* bc_set_param keys */
#define BC_PARAM_BLOCKS 1     /* force a block scheme: pigeonhole blocks b (k+1 <= b <= k+4); 0 = choose */
#define BC_PARAM_PATH 2       /* 0 auto, 1 probe kernel, 2 bucket-join kernel (16-byte window records),
                                 3 compact bucket-join kernel (8-byte window records, short spacers) */
#define BC_PARAM_COUNT_CANDIDATES 3 /* 1 = count verified candidates into bc_stats        */
#define BC_PARAM_HIT_CAPACITY 4     /* initial hit-buffer capacity (records)              */
#define BC_PARAM_SPACER_ID_BASE 5   /* added to every spacer_id (global ids of a library shard) */
#define BC_PARAM_SCAN_PART 6         /* genome-range sharding: value = rank | world << 16; this
                                       context scans only its 1/world slice of window starts  */
#define BC_PARAM_WINDOW_SORT 7       /* bucket-join path, genome-side sort: 0 auto, 1 direct scatter,
                                        2 two-pass shared-memory radix scatter                    */
#define BC_PARAM_KEY_NT 9            /* force a seed covering design with keys of this many bases (a row of
                                        the compiled-in design table for this L and k); 0 = choose      */
#define BC_PARAM_SLOT_PART 10         /* slot-range sharding (strong scaling of ONE library over several
                                        GPUs): value = rank | world << 16.  Every context holds the whole
                                        genome and library but indexes, sorts and verifies only its 1/world
                                        range of the seed directory; the contexts' hit sets are disjoint and
                                        their union is the whole result                                  */
#define BC_PARAM_KEY_CAP 12           /* upper bound on the seed key length of block schemes (bases; 0 = from the
                                        library size, at most 12): shorter keys = smaller directories    */
#define BC_PARAM_COMPACT_DIR 13       /* probe path: 0 auto (packed directory, one 4-byte load per probe, when the
                                        index has fewer than 2^26 entries), 1 off                      */
#define BC_PARAM_INDEX_SORT 11        /* compact join path, library-side sort: 0 auto (radix passes), 1 two-level
                                        atomic scatter (the builder of the other paths), 2 radix passes   */
#define BC_PARAM_JOIN_CHUNK 8        /* bucket-join path: upper bound on the window positions sorted per
                                        pass over the genome (0 = as many as the workspace holds); the
                                        passes append to one hit buffer                              */

typedef struct {
    uint64_t genome_bases;    /* G                                                   */
    uint64_t library_spacers; /* n                                                   */
    uint32_t spacer_len;      /* L                                                   */
    uint32_t k;
    uint32_t blocks;          /* b of the block scheme in use, 0 for a covering design */
    uint32_t combos;          /* seed combinations (C(b, k) for a block scheme)      */
    uint32_t path;            /* 1 probe, 2 join, 3 compact join                     */
    uint32_t scan_launches;   /* kernels launched by the last bc_search              */
    uint64_t hits;            /* records produced by the last bc_search              */
    uint64_t candidates;      /* verified (window, entry) pairs, if counting enabled */
    uint64_t probes;          /* directory look-ups, if counting enabled             */
    float ms_pack_genome;     /* device time of the last bc_set_genome               */
    float ms_pack_library;    /* device time of the last bc_set_library              */
    float ms_build_index;     /* device time of the last bc_build_index              */
    float ms_search;          /* device time of the last bc_search (all its kernels) */
    float ms_scan_kernel;     /* device time of the verify kernel(s) only            */
    float ms_genome_bucket;   /* join path: device time of the genome bucketing kernels */
    uint32_t index_launches;  /* kernels launched by the last bc_build_index          */
    uint32_t key_nt;          /* longest seed key of the scheme in use (bases)        */
    float ms_sort_hits;       /* device time of the last bc_sort_hits                 */
    float ms_win_count;       /* compact join: device time of the window count kernel (part of ms_genome_bucket) */
    float ms_win_bin;         /* compact join: ... of radix pass A                    */
    float ms_win_place;       /* compact join: ... of radix pass B                    */
    float ms_finish;          /* compact join: ... of k_cfinish (part of ms_scan_kernel; 0 when streaming) */
    uint32_t search_attempts; /* passes bc_search needed: 1, +1 per hit-buffer or verify-item-queue overflow */
    uint32_t reserved0;
} bc_stats;

int bc_abi_version(void);

/* Usable CUDA devices (0 when there is none).  The reference's `num_threads` (bowtie -p,
 * BowtieRunner.py:104,120) maps to this many GPUs at most. */
int bc_device_count(void);

/* Replaces `BowtieRunner()` / `__enter__` (BowtieRunner.py:14-20,49-50).  `device` is the
 * CUDA ordinal; the context owns a stream and all device memory it allocates. */
int bc_create(bc_ctx** out, int device);

/* Replaces `__exit__` -> temp_dir.cleanup() (BowtieRunner.py:52-53). */
void bc_destroy(bc_ctx* ctx);

/* Replaces make_fasta + bowtie-build (BowtieRunner.py:55-62,78-102): packs the genome
 * into 1-bit planes (hi, lo, ambiguity) resident in HBM.  `ascii` = all contigs
 * concatenated (any case; non-ACGT = ambiguous), `contig_offsets[n_contigs+1]` = start
 * of each contig and the total length.  The caller may free its buffers on return.
 * The _dev variant takes a device pointer for `ascii` (offsets stay on the host) and
 * enqueues on `stream` (a cudaStream_t, or NULL for the context's stream). */
int bc_set_genome(bc_ctx* ctx, const uint8_t* ascii, const uint64_t* contig_offsets, uint32_t n_contigs);
int bc_set_genome_dev(bc_ctx* ctx, const uint8_t* d_ascii, const uint64_t* contig_offsets,
                      uint32_t n_contigs, void* stream);

/* Replaces make_fastq (BowtieRunner.py:64-76): `ascii_spacers` = n spacers of L
 * characters each, back to back (1 <= L <= 32).  Non-ACGT characters mismatch
 * everything.  spacer_id in the results = index into this array. */
int bc_set_library(bc_ctx* ctx, const uint8_t* ascii_spacers, uint32_t n, uint32_t L);
int bc_set_library_dev(bc_ctx* ctx, const uint8_t* d_ascii_spacers, uint32_t n, uint32_t L, void* stream);

/* Replaces PAMFinder(records, pam, direction) (PAMProcessor.py:60-63) and the PAM
 * arguments of targets.py (:864-878).  `pam` up to 8 letters, "" = no PAM.
 * direction 0 = downstream (3' of the protospacer), 1 = upstream (5'). */
int bc_set_pam(bc_ctx* ctx, const char* pam, int direction, uint32_t flags);

int bc_set_param(bc_ctx* ctx, int key, int64_t value);

/* Replaces create_index (BowtieRunner.py:78-102) for the library side: builds the seed index for
 * <= k mismatches on device (0 <= k <= 3).  Both the library and the genome must be loaded: the seed
 * scheme (block scheme or covering design, key length) and the scan path are chosen by a cost model
 * that needs the genome size; a later bc_set_genome / bc_set_library invalidates the index and
 * bc_search rebuilds it.  The kernels are enqueued on the context's stream and the call returns
 * without waiting for them (a following bc_search runs behind them; device errors of the build surface
 * there); bc_stats.ms_build_index is filled in by the next bc_search or bc_get_stats. */
int bc_build_index(bc_ctx* ctx, int k);

/* Replaces align (BowtieRunner.py:104-141): every ungapped end-to-end alignment of every
 * spacer to both strands with <= k mismatches, PAM annotated in the same pass.  Results
 * stay on the device until copied.  Builds the index first if bc_build_index(k) was
 * not called. */
int bc_search(bc_ctx* ctx, int k, uint64_t* n_hits_out);

/* Replaces reading the SAM file (PySamParser.py:16-19).  Order is unspecified. */
int bc_copy_hits(bc_ctx* ctx, bc_hit* dst, uint64_t cap);

/* Streamed delivery of the same records: with a sink set, bc_search copies the hit records to
 * `dst` (host memory, `cap` records; page-locked memory lets the copies overlap the search)
 * WHILE the search runs, so that when bc_search returns dst[0..n_hits) is complete and no
 * bc_copy_hits is needed - the counterpart of bowtie writing the SAM file while it aligns
 * (BowtieRunner.py:111-136, `-S ... sam_path`).  `dst` may also be device memory of this or a
 * peer GPU (see bc_peer_open).  dst = NULL removes the sink.  If more than
 * `cap` hits are found bc_search fails with BC_ELIMIT; the records stay on the device and
 * bc_copy_hits still works. */
int bc_set_hit_sink(bc_ctx* ctx, bc_hit* dst, uint64_t cap);

/* Streamed hand-over on the device: `fn(user, d_hits, begin, end)` is called on the calling host
 * thread, from inside bc_search, every time another part [begin, end) of the device hit buffer
 * is final (the bucket-join path reports after each of its verify slices while the later slices
 * are still running; the probe path reports once).  `d_hits` is the buffer base; the records
 * stay valid until the next bc_search.  Used by the multi-GPU merge to send finished records to
 * the gathering rank while the search continues (the NCCL gather the reference has no
 * counterpart for; BASELINE north star).  With a callback installed bc_search does not repeat a
 * search whose hit buffer overflowed: it fails with BC_ELIMIT, bc_stats.hits holds the needed
 * capacity and the caller raises BC_PARAM_HIT_CAPACITY.  fn = NULL removes the callback. */
typedef void (*bc_slice_fn)(void* user, const bc_hit* d_hits, uint64_t begin, uint64_t end);
int bc_set_slice_callback(bc_ctx* ctx, bc_slice_fn fn, void* user);

/* Peer result buffers: the multi-GPU merge without a collective.  The gathering process exports a
 * device buffer (bc_peer_export: allocation + CUDA IPC handle, 64 bytes, to be shipped to the
 * other processes of the box by any means); every other process opens it (bc_peer_open) and
 * names ITS slice of it as its hit sink (bc_set_hit_sink accepts any unified-address pointer:
 * host memory or peer device memory).  Its records then cross NVLink through the copy engines
 * while its search is still running - no SMs, no rendezvous.  bc_peer_close releases a mapping
 * (owner = 0) or the allocation itself (owner = 1). */
int bc_peer_export(bc_ctx* ctx, uint64_t n_records, bc_hit** d_ptr, unsigned char handle[64]);
int bc_peer_open(bc_ctx* ctx, const unsigned char handle[64], bc_hit** d_ptr);
int bc_peer_close(bc_ctx* ctx, bc_hit* d_ptr, int owner);

/* Orders the device hit buffer of the last bc_search in place, before it is copied out: the counterpart
 * of bowtie writing its alignments read by read, and with --best fewest mismatches first
 * (BowtieRunner.py:119).  order 0: (spacer_id, gpos, strand) - the canonical order of the parity tests;
 * order 1: (spacer_id, mismatches, gpos, strand) - `bowtie --best`.  Records streamed earlier through a
 * hit sink are not re-sent. */
int bc_sort_hits(bc_ctx* ctx, int order);

/* Device-side view of the result buffer (valid until the next bc_search/bc_destroy),
 * for callers that gather across GPUs without a host round trip. */
int bc_hits_device(bc_ctx* ctx, const bc_hit** d_hits, uint64_t* n_hits);

int bc_get_stats(bc_ctx* ctx, bc_stats* out);

/* Replaces the per-position Python loops that enumerate guides next to PAM sites
 * (design_guides.py:22-49 find_sequences_with_barcode_and_pam; PAMProcessor.py:27-57): every
 * distinct pure-ACGT L-mer adjacent to a match of `pam` on either strand of every contig of
 * the resident genome, as a set.  direction 0: PAM 3' of the guide, 1: PAM 5'.  flags:
 * BC_PAM_IUPAC as in bc_set_pam; BC_GUIDES_REFERENCE_RANGE reproduces design_guides.py:31,
 * whose loop bound drops the last len(pam) start positions for an upstream PAM as well.
 * bc_copy_guides returns 2-bit codes: base j of the guide at bits [2j, 2j+2), A0 C1 G2 T3;
 * order unspecified (the reference builds a Python set). */
#define BC_GUIDES_REFERENCE_RANGE 4u
int bc_enumerate_guides(bc_ctx* ctx, uint32_t L, const char* pam, int direction, uint32_t flags,
                        uint64_t* n_guides_out);
int bc_copy_guides(bc_ctx* ctx, uint64_t* dst, uint64_t cap);

/* Replaces BowtieError.message (BowtieRunner.py:144-150). */
const char* bc_last_error(bc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* BARCODER_B200_H */
