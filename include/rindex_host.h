/* rindex_host.h — C ABI of the HOST side (index construction, container I/O, workload
 * generators). Plain g++ code, no CUDA. The query path is NOT here: see rindex_gpu.h.
 *
 * Reference interfaces replaced (paths under /root/reference):
 *   rih_build_from_text   r_index<>::r_index(string&, bool)        internal/r_index.hpp:42-150
 *                         (+ sufsort :553-634, rle_string ctor rle_string.hpp:52-124)
 *   rih_save / rih_load   r_index<>::serialize / load              internal/r_index.hpp:382-422
 *                         (own container; byte-compat with SDSL blobs is out of scope)
 *   rih_view              read access to what the reference keeps in F / bwt / pred /
 *                         samples_last / pred_to_run                internal/r_index.hpp:655-665
 */
#ifndef RINDEX_HOST_H_
#define RINDEX_HOST_H_
#include <stdint.h>
#include "rindex_gpu.h" /* rig_logical_view */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rih_index rih_index;

#define RIH_OK 0
#define RIH_ERR_RESERVED_CHARS -1 /* text contains 0x00 or 0x01 (r_index.hpp:46-51) */
#define RIH_ERR_IO -2
#define RIH_ERR_FORMAT -3
#define RIH_ERR_ARG -4

int rih_build_from_text(const uint8_t* text, uint64_t len, rih_index** out);
/* Scalable construction by prefix-free parsing (r-index_b200/host/pfp_builder.hpp): same
 * result as rih_build_from_text, memory proportional to the dictionary and the parse of the text instead of
 * 5-9 bytes per symbol. w = window length (0 = 10), p = trigger modulus (0 = 100). stats_out (may be NULL):
 * distinct phrases, dictionary bytes, parse length, suffix groups, BWT rows emitted as whole blocks, rows merged
 * one by one. */
int rih_build_from_text_pfp(const uint8_t* text, uint64_t len, uint32_t w, uint32_t p, uint64_t stats_out[6], rih_index** out);
/* What ri-build uses: prefix-free parsing from 16 MB up (SA-IS fallback), SA-IS below; env RIB_BUILDER=sais|pfp forces one. */
int rih_build_auto(const uint8_t* text, uint64_t len, int* used_pfp, rih_index** out);
void rih_destroy(rih_index* idx);
/* Borrowed pointers, valid until rih_destroy. */
int rih_view(const rih_index* idx, rig_logical_view* view);
/* `with_flag_byte` != 0 writes/skips the leading 1-byte `fast` flag the reference CLIs use
 * (ri-build.cpp:133, ri-count.cpp:155-158). */
int rih_save(const rih_index* idx, const char* path, int with_flag_byte);
int rih_load(const char* path, int with_flag_byte, rih_index** out);

/* Synthetic workloads of SURVEY.md §8d. `kind`: 0 = dna_drift (C2), 1 = dna_indep (C5),
 * 2 = versioned_doc (C3), 3 = pangenome (C4). p0/p1 are kind-specific:
 *   0: p0 = base length, p1 = SNPs per copy          1: p0 = base length, p1 = SNP rate * 1e9
 *   2: p0 = base length, p1 = sigma (p_edit = 0.25)  3: p0 = base length, p1 = variant sites
 * Writes exactly n bytes into out. */
int rih_gen_text(int kind, uint64_t n, uint64_t p0, uint64_t p1, uint64_t seed, uint8_t* out);
/* N*m pattern bytes sampled from text at uniform starts in [0, start_limit-m]. */
int rih_gen_patterns(const uint8_t* text, uint64_t text_len, uint64_t N, uint64_t m, uint64_t start_limit,
                     uint64_t seed, uint8_t* out);
/* Suffix array (int64) of text·\0, for tests (SA[0] = len). */
int rih_suffix_array(const uint8_t* text, uint64_t len, int64_t* sa_out);

#ifdef __cplusplus
}
#endif
#endif
