/* rindex_gpu.h — the drop-in boundary: C ABI of librindex_gpu.so (hand-written sm_100a CUDA).
 *
 * The reference (nicolaprezza/r-index @ 7009b53) has no FFI layer: its hot path is the public
 * C++ surface of ri::r_index<>. Each entry point below names the reference interface it
 * replaces (paths under /root/reference). INTEGRATION.md shows the binding a maintainer of the
 * reference would add. Plain pointers and sizes only; no torch, no C++ types.
 *
 * Conventions (each checked against the reference source):
 *   - positions, ranges, counts: uint64_t                        (internal/definitions.hpp:39)
 *   - SA ranges are inclusive; the EMPTY range is the literal pair {1,0} (r_index.hpp:175,184)
 *   - patterns: N * m contiguous bytes, fixed length m           (ri-count.cpp:104-110)
 *   - locate order per pattern: SA[hi], SA[hi-1], ..., SA[lo]    (r_index.hpp:340-351)
 *   - never exits or throws: returns RIG_OK (0) or a negative code; rig_strerror() names it.
 */
#ifndef RINDEX_GPU_H_
#define RINDEX_GPU_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIG_OK 0
#define RIG_ERR_ARG -1      /* null pointer / inconsistent sizes */
#define RIG_ERR_CUDA -2     /* a CUDA runtime call failed; rig_last_cuda_error() has the text */
#define RIG_ERR_NO_DEVICE -3
#define RIG_ERR_CAPACITY -4 /* occurrence buffer too small; *occ_total holds the needed count */
#define RIG_ERR_INDEX -5    /* logical arrays are not a valid r-index (unsorted pred, bad sums, ...) */
#define RIG_ERR_NOMEM -6

/* Layout-free content of an r-index — what the reference keeps in
 *   F            internal/r_index.hpp:655   (here 257 entries, F[256] = n)
 *   bwt          internal/r_index.hpp:657   rle_string: run heads + run lengths (rle_string.hpp:52-124)
 *   samples_last internal/r_index.hpp:664   SA-1 (wrap to n-1) at the last position of every run
 *   pred         internal/r_index.hpp:663   sorted SA-1 at the first position of every run
 *   pred_to_run  internal/r_index.hpp:665
 * All arrays are HOST pointers, read during rig_index_create only. */
typedef struct rig_logical_view {
    uint64_t n; /* BWT length = text length + 1 */
    uint64_t r; /* number of BWT runs */
    const uint64_t* F;            /* [257] */
    const uint8_t* run_heads;     /* [r] */
    const uint64_t* run_lens;     /* [r] */
    const uint64_t* samples_last; /* [r] */
    const uint64_t* pred_pos;     /* [r] ascending */
    const uint64_t* pred_to_run;  /* [r] */
} rig_logical_view;

typedef struct rig_options {
    uint32_t runs_per_block;   /* 0 = default; 4, 8 or 16: lanes cooperating on one rank query */
    uint32_t lf_bucket_log2;   /* 0 = auto: log2 of (directory buckets per run block) */
    uint32_t phi_bucket_log2;  /* 0 = auto: log2 of (directory buckets per Phi sample) */
    uint32_t expand_threads;   /* 0 = default block size of the Phi expansion kernel */
    uint32_t reserved[4];      /* reserved[0] = phi_jump D: 0 = auto (4), 1/2/4/8 = occurrences produced per Phi record lookup, 6 = six per
                                  32-byte entry (n < 2^32-1 only; RIG_ERR_ARG when the entry's next/count fields do not fit 24 + 8 bits);
                                  reserved[1] bit0 = force 64-bit position words (testing the n >= 2^32 paths);
                                  reserved[2] = SEG of the two-pass Phi expansion: 0 = auto, 1 = off (single pass),
                                  16/32/64/128/256 = occurrences per seed-table hop (window size in output slots);
                                  reserved[3]: unused, must be 0 */
} rig_options;

typedef struct rig_index_info {
    uint64_t n, r, sigma;        /* sigma = distinct BWT symbols (terminator included) */
    uint64_t device_bytes;       /* HBM footprint of the flattened index */
    uint64_t lf_blocks, lf_buckets, phi_buckets;
    uint32_t runs_per_block, lf_shift, phi_shift, device;
    uint32_t sm_count, phi_jump;  /* phi_jump = D: each Phi record holds the deltas of Phi^1..Phi^D */
    uint64_t phi_jump_pieces;     /* pieces of the refined Phi^1..Phi^D translation (<= D*r) */
    uint32_t words32;             /* 1: n < 2^32-1, position words and Phi entries are 32-bit */
    uint32_t lf_record_bytes;     /* bytes of one block record of the backward-search structure (what one rank query reads besides its directory sector) */
    uint32_t seed_jump, seed_shift; /* seed table Phi^SEG of the two-pass expansion (0 = none) and its bucket shift */
    uint64_t seed_pieces, seed_bytes;
} rig_index_info;

/* CUDA-event timings (ms) of the phases of the most recent batch call on this index. */
typedef struct rig_timing {
    float h2d_ms;     /* pattern upload (host-buffer entry points only) */
    float search_ms;  /* backward-search kernel (count or count+toehold) */
    float scan_ms;    /* exclusive scans of n_occ / chain counts (locate only) */
    float expand_ms;  /* Phi expansion kernel (locate only) */
    float d2h_ms;     /* result download (host-buffer entry points only) */
    uint32_t launches;      /* kernels launched by the call */
    uint32_t slices;        /* locate: kernels of the Phi expansion: 1 = the fused producer/consumer kernel (or the single-pass
                               walk), 2 = seed pass + window pass (seed_ms / window_ms then time them separately), 0 = none ran */
    uint64_t lf_steps;      /* executed LF steps (early exits excluded), r_index.hpp:297 */
    uint64_t occ_total;     /* occurrences written */
    uint64_t chains;        /* independent Phi chains the ranges were split into */
    float seed_ms;          /* two kernels: pass 1 (toeholds, chain heads, one seed per output window); fused: ~0 */
    float window_ms;        /* two kernels: pass 2 (one lane per output window); fused: the whole kernel; inside expand_ms */
} rig_timing;

typedef struct rig_index rig_index;

int rig_device_count(void);
const char* rig_strerror(int code);
const char* rig_last_cuda_error(void);
const char* rig_version(void);

/* Flatten the logical arrays into HBM-resident word arrays with interleaved rank directories
 * and upload them to `device`. Replaces r_index<>::load (internal/r_index.hpp:407-422). */
int rig_index_create(const rig_logical_view* view, int device, rig_index** out);
int rig_index_create_ex(const rig_logical_view* view, int device, const rig_options* opt, rig_index** out);
void rig_index_destroy(rig_index* idx);
/* The FLATTENED index as a file: what rig_index_create[_ex] builds in HBM, byte for byte, so that the next load is a
 * file read + upload instead of a flatten (config C4, r = 2.4e7: a minute on the host -> seconds). The reference's
 * "Load time" (ri-count.cpp:126-127) is the cost this shortens. rig_index_load_flat: `check` (may be NULL) = the
 * logical index the file must belong to (n, r and a 64-bit digest of its arrays are compared); RIG_ERR_INDEX if the
 * file is not a flat index of this library version or belongs to another index. */
int rig_index_save_flat(const rig_index* idx, const char* path);
int rig_index_load_flat(const char* path, const rig_logical_view* check, int device, rig_index** out);
int rig_index_info_get(const rig_index* idx, rig_index_info* info);

/* Replaces N calls of r_index<>::count (internal/r_index.hpp:292-302; LF :171-190;
 * rle_string::rank rle_string.hpp:170-218). HOST buffers; copies are part of the call. */
int rig_count_batch(rig_index* idx, const uint8_t* patterns, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi);

/* Replaces N calls of r_index<>::locate_all (internal/r_index.hpp:328-355; count_and_get_occ
 * :482-545; Phi :195-221). HOST buffers. occ_offsets has N+1 entries; pattern p's occurrences
 * are occ[occ_offsets[p] .. occ_offsets[p+1]). If occ == NULL or occ_capacity (in entries) is
 * too small, lo/hi/occ_offsets/occ_total are still filled and RIG_ERR_CAPACITY is returned. */
int rig_locate_batch(rig_index* idx, const uint8_t* patterns, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi,
                     uint64_t* occ_offsets, uint64_t* occ, uint64_t occ_capacity, uint64_t* occ_total);

/* Same operations on DEVICE buffers (pointers valid on the index's device); `stream` is a
 * cudaStream_t (NULL = the index's own stream). Count is fully asynchronous. Locate queues the
 * search AND the expansion, then waits only for the batch totals (occ_total is returned on the
 * host): when it returns, the expansion may still be running on `stream`. The expansion kernels
 * take the totals from device memory and skip themselves when occ_total > occ_capacity. */
int rig_count_batch_dev(rig_index* idx, const uint8_t* d_patterns, uint64_t N, uint64_t m, uint64_t* d_lo,
                        uint64_t* d_hi, void* stream);
int rig_locate_batch_dev(rig_index* idx, const uint8_t* d_patterns, uint64_t N, uint64_t m, uint64_t* d_lo,
                         uint64_t* d_hi, uint64_t* d_occ_offsets, uint64_t* d_occ, uint64_t occ_capacity,
                         uint64_t* occ_total, void* stream);

/* Order-independent + order-dependent digest of a device u64 array (for parity at sizes where
 * copying every occurrence back is pointless): out[0] = sum, out[1] = sum of v*(i+1) (mod 2^64). */
int rig_digest_dev(rig_index* idx, const uint64_t* d_values, uint64_t count, uint64_t out[2], void* stream);

/* ---- ri-locate's post-processing options on the device ----
 * -o (ri-locate.cpp:146-152): every pattern's occurrences sorted ascending (std::sort per pattern).
 * -c (ri-locate.cpp:156-190): per pattern, the number of occurrences found in the text by brute force must equal
 *    the number located, and text[o, o+m) must equal the pattern for every located o. On the device the N
 *    string::find scans are one hash join between the text's m-grams and the patterns. */
typedef struct rig_check_report {
    uint64_t patterns_checked;
    uint64_t wrong_count_patterns;   /* patterns whose brute-force count differs from hi-lo+1 ("wrong number of located occurrences") */
    uint64_t wrong_occurrences;      /* located positions whose text does not match ("wrong occurrence") */
    uint64_t unsorted_or_duplicate;  /* adjacent positions of one pattern not strictly ascending (after the sort: repeats) */
    uint64_t first_bad_pattern;      /* smallest offending pattern index, ~0 if none */
    uint64_t first_bad_position;     /* smallest offending text position, ~0 if none */
} rig_check_report;

#define RIG_LOCATE_SORT 1u   /* -o: per-pattern ascending order instead of locate_all order */
#define RIG_LOCATE_CHECK 2u  /* -c: needs rig_text_attach; implies the sort, as in the reference */
#define RIG_LOCATE_DEVICE_ONLY 4u /* rig_locate_batch_ex: the occurrences stay in the library's own device buffer (occ and
                                     occ_capacity are ignored; the buffer grows to the batch as needed, without repeating the
                                     search); ranges, offsets and *occ_total come back as usual. What ri-locate does with its
                                     occurrences when neither -o nor -c is given: it drops them (ri-locate.cpp:144). Fetch any
                                     part afterwards with rig_fetch_occurrences. */

/* Copy the indexed text (len = n - 1 bytes, no terminator) into HBM for rig_check_dev / RIG_LOCATE_CHECK. */
int rig_text_attach(rig_index* idx, const uint8_t* text, uint64_t len);
/* In-place segmented sort of a device occurrence array (segments = d_occ_offsets[p] .. d_occ_offsets[p+1]). */
int rig_sort_occurrences_dev(rig_index* idx, uint64_t N, const uint64_t* d_occ_offsets, uint64_t* d_occ, uint64_t total,
                             void* stream);
/* The -c check on device buffers; synchronises and fills *report. `sorted` != 0 also checks strict ascent. */
int rig_check_dev(rig_index* idx, const uint8_t* d_patterns, uint64_t N, uint64_t m, const uint64_t* d_lo,
                  const uint64_t* d_hi, const uint64_t* d_occ_offsets, const uint64_t* d_occ, uint64_t total, int sorted,
                  rig_check_report* report, void* stream);
/* rig_locate_batch with post-processing flags (HOST buffers). report may be NULL without RIG_LOCATE_CHECK. */
int rig_locate_batch_ex(rig_index* idx, const uint8_t* patterns, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi,
                        uint64_t* occ_offsets, uint64_t* occ, uint64_t occ_capacity, uint64_t* occ_total, uint32_t flags,
                        rig_check_report* report);

/* Occurrences [first, first + count) of this index's most recent RIG_LOCATE_DEVICE_ONLY call, copied to a HOST buffer
 * (after RIG_LOCATE_SORT / _CHECK: in per-pattern ascending order). */
int rig_fetch_occurrences(rig_index* idx, uint64_t first, uint64_t count, uint64_t* out);
/* Page-locked host memory for pattern / result buffers (cudaMallocHost / cudaFreeHost): downloads into pageable memory
 * run at a fraction of the PCIe rate. NULL on failure. */
void* rig_host_alloc(uint64_t bytes);
void rig_host_free(void* p);

/* rig_locate_batch with 32-bit positions, for indexes with n <= 2^32 (RIG_ERR_ARG otherwise): the same values as
 * rig_locate_batch, written as uint32_t by the expansion kernels themselves (default table form, no sort; otherwise
 * narrowed on the device afterwards): half the store sectors and half the bytes over PCIe (the host-buffer call is
 * PCIe-bound: 7.8 ms instead of 15.5 ms on config C2). flags: 0 or RIG_LOCATE_SORT. The reference's locate_all
 * returns 64-bit ulint; this entry point is an addition for callers that store 32-bit positions anyway. */
int rig_locate_batch32(rig_index* idx, const uint8_t* patterns, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi,
                       uint64_t* occ_offsets, uint32_t* occ, uint64_t occ_capacity, uint64_t* occ_total, uint32_t flags);

/* ---- single-position navigation as batches (no reference CLI calls these) ----
 *   RIG_NAV_BWT   r_index<>::operator[](i)  bwt[i] (the terminator row holds 0x01)      internal/r_index.hpp:162-164
 *   RIG_NAV_LF    r_index<>::LF(i)          F[c] + bwt.rank(i,c), c = bwt[i]            internal/r_index.hpp:224-229
 *   RIG_NAV_FL    r_index<>::FL(i)          bwt.select(i - F[c], c), c = F_at(i)        internal/r_index.hpp:232-243
 *   RIG_NAV_F_AT  r_index<>::F_at(i)        symbol of column F at row i                 internal/r_index.hpp:263-271
 * out[k] = op(positions[k]) as uint64_t; positions >= n give ~0. */
#define RIG_NAV_BWT 0
#define RIG_NAV_LF 1
#define RIG_NAV_FL 2
#define RIG_NAV_F_AT 3
int rig_navigate_batch(rig_index* idx, int op, const uint64_t* positions, uint64_t N, uint64_t* out);
int rig_navigate_batch_dev(rig_index* idx, int op, const uint64_t* d_positions, uint64_t N, uint64_t* d_out, void* stream);
/* r_index<>::get_bwt (internal/r_index.hpp:375-377, rle_string::toString) restricted to [from, from+len): the BWT
 * as bytes, terminator row = 0x01 (HOST buffer). */
int rig_get_bwt(rig_index* idx, uint64_t from, uint64_t len, uint8_t* out);

/* rle_string::break_range (internal/rle_string.hpp:261-302) for N queries (lo[k], hi[k], c[k]): the maximal sub-ranges
 * of [lo, hi] that hold only c, in ascending order; requires bwt[lo] == bwt[hi] == c as the reference does (it asserts;
 * here such a query yields no range). Query k's ranges are [out_first[j], out_last[j]] for j in
 * [out_offsets[k], out_offsets[k+1]). HOST buffers; two-call protocol: with capacity (in ranges) too small,
 * out_offsets and *total are still filled and RIG_ERR_CAPACITY is returned. */
int rig_break_range_batch(rig_index* idx, const uint64_t* lo, const uint64_t* hi, const uint8_t* c, uint64_t N,
                          uint64_t* out_offsets, uint64_t* out_first, uint64_t* out_last, uint64_t capacity, uint64_t* total);
/* rle_string::closest_run_break (internal/rle_string.hpp:455-493) for N queries: bwt[lo] == c -> the last position of
 * lo's run, otherwise the first position after lo that holds c (~0 if none: the reference asserts). hi is the
 * range's end as in the reference's signature (its value does not enter the result). */
int rig_closest_run_break_batch(rig_index* idx, const uint64_t* lo, const uint64_t* hi, const uint8_t* c, uint64_t N, uint64_t* out);

/* ---- a job sharded over several GPUs, planned on every device (index replicated) ----
 * rig_plan_batch_dev: the backward search of the WHOLE batch on this device — ranges, toeholds and the exclusive prefix of
 * the occurrence counts, i.e. the first half of rig_locate_batch_dev — and, from that prefix, the cut points of `shards`
 * contiguous shards of equal work (work(p) = n_occ(p) + per_pattern_cost; the integer rule of rig_balanced_cuts_dev).
 * cuts: HOST array of shards + 1 entries (cuts[0] = 0, cuts[shards] = N). No communication: every device reaches the
 * same cuts from the same batch, and a count pass over the whole batch costs less than a collective's latency for
 * batches up to a few million LF steps (beyond that: count shards, exchange the counts, rig_balanced_cuts_dev).
 * rig_expand_shard_dev: the Phi expansion of patterns [c0, c1) of the batch most recently planned on this index (same N
 * and device arrays; any other count / locate / plan call on the index in between invalidates the plan: RIG_ERR_ARG):
 * the occurrences of pattern p go to d_occ[d_occ_offsets[p] - d_occ_offsets[c0] ...], in locate_all order.
 * RIG_ERR_CAPACITY (and *shard_total) when occ_capacity is too small; several shards may be expanded after one plan. Replaces, for a rank's shard, the
 * per-pattern loop of ri-locate.cpp:126-145. */
int rig_plan_batch_dev(rig_index* idx, const uint8_t* d_patterns, uint64_t N, uint64_t m, uint64_t* d_lo, uint64_t* d_hi,
                       uint64_t* d_occ_offsets, uint32_t shards, uint64_t per_pattern_cost, uint64_t* cuts,
                       uint64_t* occ_total, void* stream);
int rig_expand_shard_dev(rig_index* idx, uint64_t N, uint64_t c0, uint64_t c1, const uint64_t* d_lo, const uint64_t* d_hi,
                         const uint64_t* d_occ_offsets, uint64_t* d_occ, uint64_t occ_capacity, uint64_t* shard_total,
                         void* stream);

/* ---- multi-GPU fan-out on device buffers ----
 * The index is replicated, the patterns are sharded; a locate job first COUNTS on equal-count shards, the hosts
 * all-gather the per-pattern counts (8 bytes per pattern: NCCL), and the batch is re-cut into contiguous shards of
 * near-equal work, work(p) = n_occ(p) + per_pattern_cost. rig_counts_dev: n_occ from the ranges of a count call
 * (occ(), r_index.hpp:307-313). rig_balanced_cuts_dev: the shards + 1 ascending cut points (HOST array; synchronises),
 * cuts[0] = 0, cuts[shards] = N; rank k takes patterns [cuts[k], cuts[k+1]). The rule is integer arithmetic and the
 * same in r-index_b200/_shard.py and host/cli_common.hpp. */
int rig_counts_dev(rig_index* idx, const uint64_t* d_lo, const uint64_t* d_hi, uint64_t N, uint64_t* d_nocc, void* stream);
int rig_balanced_cuts_dev(rig_index* idx, const uint64_t* d_nocc, uint64_t N, uint32_t shards, uint64_t per_pattern_cost,
                          uint64_t* cuts, void* stream);

int rig_last_timing(const rig_index* idx, rig_timing* t);

#ifdef __cplusplus
}
#endif
#endif
