/* ri_oracle.c — CPU restatement ("port") of the reference's count + locate path, in plain C99.
 *
 * TEST INFRASTRUCTURE ONLY. Imported/linked only by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg. The product path never touches it (and has no CPU fallback).
 *
 * Reference: nicolaprezza/r-index @ 7009b53. Each function cites the file:line it follows.
 * The third-party layer the reference delegates to (simongog/sdsl-lite, version UNPINNED by the
 * reference: no submodule, README.md:34 only links the repo) is not restated structure by
 * structure here; its primitives are replaced by their mathematical definitions over plain
 * sorted arrays: sd_vector rank(i) = #ones in [0,i), select(k) = k-th one; wt_huff rank(i,c),
 * select(k,c), access. Results are fixed by those definitions.
 *
 * PINNING: the reference ships no golden vectors. This oracle is pinned by
 *   (1) brute-force substring search on the text (rio_brute_count; the ri-locate -c idea),
 *   (2) the reference's own code run here (oracle/_ref, reference headers + SDSL-API shim),
 *       whose outputs are committed as fixtures under tests/golden/ by tests/golden/make_golden.py,
 *   (3) the brute-force known answers of SURVEY.md §4 for the bundled datasets.
 * tests/test_oracle.py runs all three.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef uint64_t u64;
typedef uint8_t u8;

typedef struct {
    u64 n, r;
    u64 F[257];         /* r_index.hpp:72-82 (F[256]=n added) */
    u8* head;           /* [r]   run heads            rle_string.hpp:68-99 */
    u64* start;         /* [r+1] run starts, start[r]=n */
    u64* samples_last;  /* [r]   r_index.hpp:131-139 */
    u64* pred_pos;      /* [r]   r_index.hpp:108-124 */
    u64* pred_to_run;   /* [r]   r_index.hpp:141-146 */
    /* per-letter view of the runs = content of runs_per_letter[c] + run_heads rank/select
       (rle_string.hpp:73-92,115-119) */
    u64 ccount[256];    /* number of c-runs */
    u64* crun[256];     /* [ccount[c]] global run ids of the c-runs, ascending */
    u64* ccum[256];     /* [ccount[c]+1] total length of the first k c-runs */
} rio_index;

/* ---------------------------------------------------------------- suffix sorting (small inputs) */
static const u64* g_rank; static u64 g_k, g_n;
static int cmp_sfx(const void* a, const void* b) {
    u64 x = *(const u64*)a, y = *(const u64*)b;
    if (g_rank[x] != g_rank[y]) return g_rank[x] < g_rank[y] ? -1 : 1;
    u64 rx = x + g_k < g_n ? g_rank[x + g_k] + 1 : 0, ry = y + g_k < g_n ? g_rank[y + g_k] + 1 : 0;
    return rx < ry ? -1 : (rx > ry ? 1 : 0);
}
/* Prefix doubling, O(n log^2 n): suffix array of s[0..n) where s[n-1] is the unique smallest byte. */
static void naive_sa(const u8* s, u64 n, int64_t* sa) {
    u64* rank = (u64*)malloc(n * 8), *tmp = (u64*)malloc(n * 8), *idx = (u64*)malloc(n * 8);
    for (u64 i = 0; i < n; ++i) { idx[i] = i; rank[i] = s[i]; }
    for (u64 k = 1;; k <<= 1) {
        g_rank = rank; g_k = k; g_n = n;
        qsort(idx, n, 8, cmp_sfx);
        tmp[idx[0]] = 0;
        for (u64 i = 1; i < n; ++i) tmp[idx[i]] = tmp[idx[i - 1]] + (cmp_sfx(&idx[i - 1], &idx[i]) < 0);
        memcpy(rank, tmp, n * 8);
        if (rank[idx[n - 1]] == n - 1 || k >= n) break;
    }
    for (u64 i = 0; i < n; ++i) sa[i] = (int64_t)idx[i];
    free(rank); free(tmp); free(idx);
}

/* Linear-time verification that sa is the suffix array of text+\0 (n = len+1 entries). 0 = ok. */
int rio_check_sa(const u8* text, u64 len, const int64_t* sa) {
    u64 n = len + 1;
    u64* inv = (u64*)malloc(n * 8);
    int ok = 0;
    for (u64 i = 0; i < n; ++i) inv[i] = ~(u64)0;
    for (u64 i = 0; i < n; ++i) {
        if (sa[i] < 0 || (u64)sa[i] >= n || inv[sa[i]] != ~(u64)0) { ok = 1; goto done; }
        inv[sa[i]] = i;
    }
    if ((u64)sa[0] != len) { ok = 2; goto done; }
    for (u64 i = 1; i < n; ++i) {
        u64 a = (u64)sa[i - 1], b = (u64)sa[i];
        unsigned ca = a < len ? text[a] : 0, cb = b < len ? text[b] : 0;
        if (ca > cb) { ok = 3; goto done; }
        if (ca == cb) {
            if (a == len || b == len) { ok = 4; goto done; }
            if (inv[a + 1] >= inv[b + 1]) { ok = 5; goto done; }
        }
    }
done:
    free(inv);
    return ok;
}

/* ---------------------------------------------------------------- construction */
/* Follows r_index<>::sufsort (r_index.hpp:553-634): BWT[x] = T[SA[x]-1] or 0x01 when SA[x]=0
 * (:587-590); samples are SA-1 with SA=0 -> n-1 (:599,604,614,619); then the F column (:72-82)
 * and the Phi predecessor arrays sorted by text position (:108-146). sa may be NULL. */
rio_index* rio_build(const u8* text, u64 len, const int64_t* sa_in) {
    for (u64 i = 0; i < len; ++i) if (text[i] == 0 || text[i] == 1) return NULL; /* r_index.hpp:46-51 */
    u64 n = len + 1;
    int64_t* sa = NULL;
    if (!sa_in) {
        u8* s = (u8*)malloc(n);
        memcpy(s, text, len); s[len] = 0;
        sa = (int64_t*)malloc(n * 8);
        naive_sa(s, n, sa);
        free(s);
        sa_in = sa;
    }
    rio_index* R = (rio_index*)calloc(1, sizeof(rio_index));
    R->n = n;
    /* pass 1: count runs */
    u64 r = 0; unsigned prev = 256;
    for (u64 x = 0; x < n; ++x) {
        unsigned c = sa_in[x] > 0 ? text[sa_in[x] - 1] : 1;
        if (c != prev) { ++r; prev = c; }
    }
    R->r = r;
    R->head = (u8*)malloc(r); R->start = (u64*)malloc((r + 1) * 8);
    R->samples_last = (u64*)malloc(r * 8);
    u64* first = (u64*)malloc(r * 8);
    u64 hist[256]; memset(hist, 0, sizeof(hist));
    u64 j = 0; prev = 256;
    for (u64 x = 0; x < n; ++x) {
        unsigned c = sa_in[x] > 0 ? text[sa_in[x] - 1] : 1;
        u64 smp = sa_in[x] > 0 ? (u64)sa_in[x] - 1 : n - 1;
        hist[c]++;
        if (c != prev) { R->head[j] = (u8)c; R->start[j] = x; first[j] = smp; ++j; prev = c; }
        R->samples_last[j - 1] = smp; /* overwritten until the run's last position */
    }
    R->start[r] = n;
    u64 acc = 0;
    for (int c = 0; c < 256; ++c) { R->F[c] = acc; acc += hist[c]; }
    R->F[256] = n;
    /* sort run-first samples by text position (r_index.hpp:108); positions are distinct */
    R->pred_pos = (u64*)malloc(r * 8); R->pred_to_run = (u64*)malloc(r * 8);
    {
        u64* key = (u64*)malloc(r * 16);
        for (u64 k = 0; k < r; ++k) { key[2 * k] = first[k]; key[2 * k + 1] = k; }
        /* simple LSD radix sort on 64-bit keys, 8 passes */
        u64* buf = (u64*)malloc(r * 16);
        for (int pass = 0; pass < 8; ++pass) {
            u64 cnt[257]; memset(cnt, 0, sizeof(cnt));
            for (u64 k = 0; k < r; ++k) cnt[((key[2 * k] >> (8 * pass)) & 255) + 1]++;
            for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
            for (u64 k = 0; k < r; ++k) { u64 d = cnt[(key[2 * k] >> (8 * pass)) & 255]++; buf[2 * d] = key[2 * k]; buf[2 * d + 1] = key[2 * k + 1]; }
            u64* t = key; key = buf; buf = t;
        }
        for (u64 k = 0; k < r; ++k) { R->pred_pos[k] = key[2 * k]; R->pred_to_run[k] = key[2 * k + 1]; }
        free(key); free(buf);
    }
    free(first);
    /* per-letter run lists (content of runs_per_letter[c] and of the WT over run heads) */
    for (u64 q = 0; q < r; ++q) R->ccount[R->head[q]]++;
    for (int c = 0; c < 256; ++c) {
        R->crun[c] = (u64*)malloc((R->ccount[c] + 1) * 8);
        R->ccum[c] = (u64*)malloc((R->ccount[c] + 1) * 8);
        R->ccum[c][0] = 0; R->ccount[c] = 0;
    }
    for (u64 q = 0; q < r; ++q) {
        int c = R->head[q]; u64 k = R->ccount[c]++;
        R->crun[c][k] = q; R->ccum[c][k + 1] = R->ccum[c][k] + (R->start[q + 1] - R->start[q]);
    }
    free(sa);
    return R;
}

void rio_free(rio_index* R) {
    if (!R) return;
    free(R->head); free(R->start); free(R->samples_last); free(R->pred_pos); free(R->pred_to_run);
    for (int c = 0; c < 256; ++c) { free(R->crun[c]); free(R->ccum[c]); }
    free(R);
}
u64 rio_n(const rio_index* R) { return R->n; }
u64 rio_r(const rio_index* R) { return R->r; }

void rio_extract(const rio_index* R, u64* F, u8* heads, u64* lens, u64* sl, u64* pp, u64* ptr) {
    memcpy(F, R->F, 257 * 8);
    for (u64 j = 0; j < R->r; ++j) {
        heads[j] = R->head[j]; lens[j] = R->start[j + 1] - R->start[j];
        sl[j] = R->samples_last[j]; pp[j] = R->pred_pos[j]; ptr[j] = R->pred_to_run[j];
    }
}

/* ---------------------------------------------------------------- primitives */
/* number of elements of sorted a[0..k) that are < x  (= sd_vector rank, sparse_sd_vector.hpp:107-112) */
static u64 lower_bound(const u64* a, u64 k, u64 x) {
    u64 lo = 0, hi = k;
    while (lo < hi) { u64 mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}
/* run containing BWT position i, 0 <= i < n  (rle_string::run_of_position, rle_string.hpp:223-256) */
static u64 run_of_position(const rio_index* R, u64 i) {
    /* largest j with start[j] <= i */
    return lower_bound(R->start, R->r, i + 1) - 1;
}
/* rle_string::rank(i,c), rle_string.hpp:170-218: #c in bwt[0,i), 0 <= i <= n */
u64 rio_rank(const rio_index* R, u64 i, u8 c) {
    if (R->ccount[c] == 0) return 0;                    /* :175 letter absent */
    if (i == R->n) return R->ccum[c][R->ccount[c]];     /* :177 */
    u64 j = run_of_position(R, i);                      /* :179-201 block scan -> run holding i */
    u64 dist = i - R->start[j];
    u64 rk = lower_bound(R->crun[c], R->ccount[c], j);  /* :207 run_heads.rank(run,c) */
    u64 tail = (R->head[j] == c) ? dist : 0;            /* :210 */
    return R->ccum[c][rk] + tail;                       /* :214-216 */
}
/* rle_string::select(i,c), rle_string.hpp:136-165: position of the i-th c, i 0-based */
static u64 rle_select(const rio_index* R, u64 i, u8 c) {
    /* j = number of complete c-runs before the i-th c (:143) */
    u64 lo = 0, hi = R->ccount[c];
    while (hi - lo > 1) { u64 mid = (lo + hi) >> 1; if (R->ccum[c][mid] <= i) lo = mid; else hi = mid; }
    u64 before = i - R->ccum[c][lo];                    /* :147 */
    return R->start[R->crun[c][lo]] + before;           /* :150-164 */
}
/* r_index::LF(range,c), r_index.hpp:171-190. Empty range = {1,0}. */
static void LF(const rio_index* R, u64 lo, u64 hi, u8 c, u64* olo, u64* ohi) {
    if ((c == 255 && R->F[c] == R->n) || R->F[c] >= R->F[c + 1]) { *olo = 1; *ohi = 0; return; } /* :174 */
    u64 c_before = rio_rank(R, lo, c);                  /* :178 */
    u64 c_inside = rio_rank(R, hi + 1, c) - c_before;   /* :181 */
    if (c_inside == 0) { *olo = 1; *ohi = 0; return; }  /* :184 */
    u64 l = R->F[c] + c_before;                         /* :186 */
    *olo = l; *ohi = l + c_inside - 1;                  /* :188 */
}
/* r_index::Phi, r_index.hpp:195-221, with predecessor_rank_circular sparse_sd_vector.hpp:153-157 */
u64 rio_phi(const rio_index* R, u64 i) {
    u64 rk = lower_bound(R->pred_pos, R->r, i);         /* pred.rank(i): samples < i */
    u64 jr = rk == 0 ? R->r - 1 : rk - 1;               /* :200 */
    u64 j = R->pred_pos[jr];                            /* :205 */
    u64 delta = j < i ? i - j : i + 1;                  /* :210 */
    u64 prev = R->samples_last[R->pred_to_run[jr] - 1]; /* :217 */
    return (prev + delta) % R->n;                       /* :219 */
}
/* r_index::count, r_index.hpp:292-302 */
static void count_one(const rio_index* R, const u8* P, u64 m, u64* olo, u64* ohi) {
    u64 lo = 0, hi = R->n - 1;                          /* full_range :155-160 */
    for (u64 i = 0; i < m && hi >= lo; ++i) LF(R, lo, hi, P[m - i - 1], &lo, &hi);
    *olo = lo; *ohi = hi;
}
/* r_index::count_and_get_occ, r_index.hpp:482-545: range + toehold k = SA[range.second] */
static void count_and_get_occ(const rio_index* R, const u8* P, u64 m, u64* olo, u64* ohi, u64* ok) {
    u64 lo = 0, hi = R->n - 1;
    u64 k = (R->samples_last[R->r - 1] + 1) % R->n;     /* :489 */
    for (u64 i = 0; i < m && hi >= lo; ++i) {
        u8 c = P[m - i - 1];
        u64 lo1, hi1;
        LF(R, lo, hi, c, &lo1, &hi1);                   /* :499 */
        if (lo1 <= hi1) {                               /* :502 */
            if (R->head[run_of_position(R, hi)] == c) { /* :505 bwt[range.second]==c */
                k--;                                    /* :509 */
            } else {
                u64 rnk = rio_rank(R, hi, c);           /* :516 */
                rnk--;                                  /* :522 */
                u64 j = rle_select(R, rnk, c);          /* :525 */
                u64 run_of_j = run_of_position(R, j);   /* :531 */
                k = R->samples_last[run_of_j];          /* :533 */
            }
        }
        lo = lo1; hi = hi1;                             /* :539 */
    }
    *olo = lo; *ohi = hi; *ok = k;
}

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

double rio_count_batch(const rio_index* R, const u8* patt, u64 N, u64 m, u64* lo, u64* hi) {
    double t0 = now_s();
    for (u64 p = 0; p < N; ++p) count_one(R, patt + p * m, m, &lo[p], &hi[p]);
    return now_s() - t0;
}
/* r_index::locate_all, r_index.hpp:328-355: SA[R], Phi(SA[R]), ... (n_occ values) per pattern,
 * written at occ[occ_offsets[p]..]. */
double rio_locate_batch(const rio_index* R, const u8* patt, u64 N, u64 m, const u64* occ_offsets, u64* occ, u64* total) {
    double t0 = now_s();
    u64 tot = 0;
    for (u64 p = 0; p < N; ++p) {
        u64 L, Rr, k;
        count_and_get_occ(R, patt + p * m, m, &L, &Rr, &k);
        u64 n_occ = Rr >= L ? Rr - L + 1 : 0;           /* :338 */
        u64* out = occ + occ_offsets[p];
        if (n_occ > 0) {
            out[0] = k;                                 /* :342 */
            for (u64 i = 1; i < n_occ; ++i) { k = rio_phi(R, k); out[i] = k; } /* :344-349 */
        }
        tot += n_occ;
    }
    if (total) *total = tot;
    return now_s() - t0;
}

/* SDSL-independent witness: overlapping occurrences by direct comparison (ri-locate.cpp:156-190 idea). */
u64 rio_brute_count(const u8* text, u64 len, const u8* p, u64 m) {
    u64 c = 0;
    if (m == 0) return len + 1;
    if (m > len) return 0;
    for (u64 i = 0; i + m <= len; ++i)
        if (text[i] == p[0] && memcmp(text + i, p, m) == 0) ++c;
    return c;
}
