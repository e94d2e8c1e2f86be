// rindex_gpu.cu — C ABI of librindex_gpu.so (include/rindex_gpu.h): index upload, batch
// count / locate launches, timing. Build: nvcc -gencode arch=compute_100a,code=sm_100a.
// There is no CPU path in this library: every query entry point launches CUDA kernels.
#include "../../include/rindex_gpu.h"
#include "flat_layout.hpp"
#include "search_kernels.cuh"
#include "phi_kernels.cuh"
#include "post_kernels.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <new>
#include <algorithm>

using rigk::FlatDev;
typedef unsigned long long ull;

static thread_local std::string g_cuda_err;

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            g_cuda_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                      \
            return RIG_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

// inside rig_index_create_ex, once the handle exists: a failing CUDA call releases everything built so far
#define CU_TRY_IX(expr)                                                                           \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            g_cuda_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                      \
            rig_index_destroy(ix);                                                                \
            return RIG_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    uint64_t generation = 0;   // bumped by every (re)allocation: the new block holds garbage, whatever its address
    int ensure(size_t bytes) {
        if (bytes <= cap) return RIG_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0; ++generation;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { g_cuda_err = std::string("cudaMalloc: ") + cudaGetErrorString(e); p = nullptr; return RIG_ERR_NOMEM; }
        cap = want;
        return RIG_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; ++generation; }
};

}  // namespace

struct rig_index {
    int device = 0;
    int sm_count = 0;
    FlatDev d{};
    rig_index_info info{};
    rig_options opt{};
    void* arena = nullptr;  // one allocation holding every flat array
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_scan = nullptr;    // the totals of a locate call have reached the host
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [6] = between the two expansion passes, [7] = end of the window pass
    // workspace (grow-only)
    DevBuf toe, jl, nch, nocc, choff, sums, patt, lo, hi, occoff, occ, items;
    DevBuf occ32;             // rig_locate_batch32: narrowed positions
    DevBuf text, big1, big2, ctable, crep, cfound;  // post-processing (-o / -c): attached text, sort tiers, hash join
    uint64_t text_len = 0;
    bool has_text = false;
    bool sort_attr_done = false;  // dynamic shared memory opt-in of the sort kernels (per device)
    ull* d_post = nullptr;      // [0] big1 count [1] big2 count [2..7] check report
    ull* h_post = nullptr;      // pinned mirror
    ull* d_counters = nullptr;  // [0] lf_steps [2..3] totals (occurrences, chains: RIG_CTR_TOTAL / _CHAINS) [4..5] digest [6] items of the seed pass (RIG_CTR_ITEMS)
    ull* h_counters = nullptr;  // pinned mirror
    rig_timing timing{};
    int variant = 0;  // see rig_index_create_ex
    size_t phi_bytes = 0;        // rec + pent span in the arena (warmed into L2 before each expansion)
    size_t lf_bytes = 0;         // span of the backward-search arrays at the head of the arena (warmed into L2 before each search)
    size_t l2_window_bytes = 0;  // persisting-L2 access policy window over the Phi records (0 = unsupported)
    float l2_hit_ratio = 1.f;
    bool timing_pending = false;
    size_t arena_bytes = 0;
    uint64_t digest = 0;           // logical_digest() of the index this handle was made from
    // the batch most recently planned on this index (rig_plan_batch_dev): cut points and the offsets at them
    uint64_t plan_N = 0;
    bool plan_valid = false;
    std::vector<uint64_t> plan_cuts, plan_occ, plan_ch;
    DevBuf d_plan;
    ull* h_plan = nullptr;         // pinned, 3 * 1025 words
    bool warned_fused = false;     // the fused kernel's cooperative launch has failed once (reported on stderr)
    uint32_t epoch = 0;            // fused expansion: tag of the current call's items (1..65535; the list is zeroed when it wraps)
    uint64_t items_zeroed = ~0ull; // generation of the item-list allocation that has been zero-filled. A fresh cudaMalloc holds garbage —
                                   // possibly another index's items with a matching epoch tag — and can come back at the address of
                                   // the block just freed, so the pointer does not tell
    uint64_t last_items_cap = 0;   // item-list capacity the most recent expansion was queued with
    bool last_two_pass = false;
    uint64_t kept_total = 0;       // occurrences of the most recent RIG_LOCATE_DEVICE_ONLY call, still in `occ`
    bool ev_valid[8] = {false, false, false, false, false, false, false, false};
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// streams, events, counters of a new handle (the arena is in place)
static int finish_create(rig_index* ix) {
    CU_TRY(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreateWithFlags(&ix->ev_scan, cudaEventDisableTiming));
    for (auto& ev : ix->ev) CU_TRY(cudaEventCreate(&ev));
    CU_TRY(cudaMalloc((void**)&ix->d_counters, 2 * RIG_CTR_BLOCK * sizeof(ull)));   // the call's block + a shard's block (rig_expand_shard_dev)
    CU_TRY(cudaMemset(ix->d_counters, 0, 2 * RIG_CTR_BLOCK * sizeof(ull)));
    CU_TRY(cudaMallocHost((void**)&ix->h_counters, 16 * sizeof(ull)));
    std::memset(ix->h_counters, 0, 16 * sizeof(ull));
    CU_TRY(cudaMalloc((void**)&ix->d_post, 8 * sizeof(ull)));
    CU_TRY(cudaMemset(ix->d_post, 0, 8 * sizeof(ull)));
    CU_TRY(cudaMallocHost((void**)&ix->h_post, 8 * sizeof(ull)));
    return RIG_OK;
}

// 64-bit digest of an index's logical content (FNV-1a over 8-byte words, four interleaved lanes): ties a file of
// the flattened form to the logical index it was made from
static uint64_t logical_digest(const rig_logical_view& v) {
    uint64_t h[4] = {0xcbf29ce484222325ull, 0x84222325cbf29ce4ull, 0x9e3779b97f4a7c15ull, 0xc2b2ae3d27d4eb4full};
    auto mix = [&](const void* p, size_t bytes) {
        const uint8_t* b = (const uint8_t*)p;
        size_t i = 0, k = 0;
        for (; i + 8 <= bytes; i += 8, ++k) { uint64_t w; std::memcpy(&w, b + i, 8); h[k & 3] = (h[k & 3] ^ w) * 0x100000001b3ull; }
        for (; i < bytes; ++i) h[0] = (h[0] ^ b[i]) * 0x100000001b3ull;
    };
    mix(&v.n, 8); mix(&v.r, 8); mix(v.F, 257 * 8); mix(v.run_heads, v.r); mix(v.run_lens, v.r * 8);
    mix(v.samples_last, v.r * 8); mix(v.pred_pos, v.r * 8); mix(v.pred_to_run, v.r * 8);
    return ((h[0] * 31 + h[1]) * 31 + h[2]) * 31 + h[3];
}

// every device pointer of a FlatDev, for rebasing between "address" and "offset into the arena"
template <class Fn>
static void for_each_pointer(FlatDev& d, Fn fn) {
    fn((const void*&)d.F); fn((const void*&)d.sid); fn(d.start); fn((const void*&)d.blk); fn(d.last); fn(d.bstart);
    fn((const void*&)d.bdir); fn(d.samples_last); fn(d.phi.rec); fn(d.phi.pent); fn(d.seed.rec); fn(d.seed.pent);
}


extern "C" {

int rig_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* rig_strerror(int code) {
    switch (code) {
        case RIG_OK: return "ok";
        case RIG_ERR_ARG: return "invalid argument";
        case RIG_ERR_CUDA: return "CUDA runtime error";
        case RIG_ERR_NO_DEVICE: return "no such CUDA device";
        case RIG_ERR_CAPACITY: return "occurrence buffer too small";
        case RIG_ERR_INDEX: return "logical arrays are not a valid r-index";
        case RIG_ERR_NOMEM: return "out of device memory";
        default: return "unknown error";
    }
}

const char* rig_last_cuda_error(void) { return g_cuda_err.c_str(); }
const char* rig_version(void) { return "rindex_b200 0.1 (sm_100a)"; }

int rig_index_create(const rig_logical_view* view, int device, rig_index** out) {
    rig_options opt;
    std::memset(&opt, 0, sizeof(opt));
    return rig_index_create_ex(view, device, &opt, out);
}

int rig_index_create_ex(const rig_logical_view* view, int device, const rig_options* optp, rig_index** out) {
    if (!view || !out) return RIG_ERR_ARG;
    *out = nullptr;
    rig_options opt;
    std::memset(&opt, 0, sizeof(opt));
    if (optp) opt = *optp;
    int ndev = rig_device_count();
    if (device < 0 || device >= ndev) return RIG_ERR_NO_DEVICE;
    CU_TRY(cudaSetDevice(device));
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b));

    int variant = 0;  // env RIG_VARIANT: bit0 persisting-L2 experiment, bit1 no evict_last hint, bit2 no L2 warm-up, bit3 forces the 64-bit code paths (as for n >= 2^32), bit5 single-pass expansion, bit12 no L2 warm-up of the search structures, bit13 seed pass and window pass as two kernels instead of the fused producer/consumer kernel, bit7 cooperative (group-per-pattern) search kernel, bit14 one lane per pattern instead of two (search_lane_kernel), bits 15-17 window-pass timing diagnostics (builds with -DRIG_WINDOW_DIAG only), bit18 rig_locate_batch32 through 64-bit positions + narrowing pass (A/B switch)
    if (const char* ev = getenv("RIG_VARIANT")) variant = atoi(ev);
    if (variant & 8) opt.reserved[1] |= 1;
    if (variant & 512) opt.reserved[1] |= 2 | 4;  // bit9: 64-bit words inside the block records even when n < 2^40 (A/B switch)
    rigf::FlatHost f;
    int rc = rigf::flatten(*view, opt, f, (uint64_t)(free_b * 0.9));
    if (rc != RIG_OK) return rc;
    if (f.bytes() + (64u << 20) > free_b) return RIG_ERR_NOMEM;

    rig_index* ix = new (std::nothrow) rig_index();
    if (!ix) return RIG_ERR_NOMEM;
    ix->device = device;
    ix->opt = opt;
    ix->variant = variant;
    cudaDeviceProp prop;
    CU_TRY_IX(cudaGetDeviceProperties(&prop, device));
    ix->sm_count = prop.multiProcessorCount;

    // one arena, every array 256-byte aligned
    std::vector<uint32_t> rec32, pent32, start32, bstart32, last32, sl32, srec32, spent32;
    auto narrow = [](const std::vector<uint64_t>& src, std::vector<uint32_t>& dst) {
        dst.resize(src.size());
        for (size_t i = 0; i < src.size(); ++i) dst[i] = (uint32_t)src[i];
    };
    if (f.w32) {  // 32-bit words: ~0 sentinels truncate to 0xFFFFFFFF, everything else is <= n < 2^32-1
        narrow(f.start, start32); narrow(f.bstart, bstart32); narrow(f.last, last32); narrow(f.samples_last, sl32);
        rec32.resize(f.phi.rec.size());
        for (size_t i = 0; i < rec32.size(); ++i) rec32[i] = (uint32_t)f.phi.rec[i];
        pent32.resize(f.phi.pent.size());
        for (size_t i = 0; i < pent32.size(); ++i) pent32[i] = (uint32_t)f.phi.pent[i];
        narrow(f.seed.rec, srec32); narrow(f.seed.pent, spent32);
    }
    struct Part { const void* src; size_t bytes; size_t off; };
    std::vector<Part> parts = {
        {f.F.data(), f.F.size() * 8, 0},            {f.sid.data(), f.sid.size() * 2, 0},
        {f.start.data(), f.start.size() * 8, 0},    {f.blk.data(), f.blk.size(), 0},
        {f.bstart.data(), f.bstart.size() * 8, 0},  {f.last.data(), f.last.size() * 8, 0},
        {f.bdir.data(), f.bdir.size() * 4, 0},      {f.samples_last.data(), f.samples_last.size() * 8, 0}};
    if (f.w32) {
        parts[2] = {start32.data(), start32.size() * 4, 0};
        parts[4] = {bstart32.data(), bstart32.size() * 4, 0};
        parts[5] = {last32.data(), last32.size() * 4, 0};
        parts[7] = {sl32.data(), sl32.size() * 4, 0};
    }
    if (f.w32) {
        parts.push_back({rec32.data(), rec32.size() * 4, 0});
        parts.push_back({pent32.data(), pent32.size() * 4, 0});
        parts.push_back({srec32.data(), srec32.size() * 4, 0});
        parts.push_back({spent32.data(), spent32.size() * 4, 0});
    } else {
        if (f.phi_packed) {
            parts.push_back({f.phi_rec_p.data(), f.phi_rec_p.size(), 0});
            parts.push_back({f.phi_pent_p.data(), f.phi_pent_p.size(), 0});
        } else {
            parts.push_back({f.phi.rec.data(), f.phi.rec.size() * 8, 0});
            parts.push_back({f.phi.pent.data(), f.phi.pent.size() * 8, 0});
        }
        parts.push_back({f.seed.rec.data(), f.seed.rec.size() * 8, 0});
        parts.push_back({f.seed.pent.data(), f.seed.pent.size() * 8, 0});
    }
    size_t total = 0;
    for (auto& p : parts) { p.off = total; total += align_up(p.bytes + 128, 256); }
    cudaError_t e = cudaMalloc(&ix->arena, total);
    if (e != cudaSuccess) { g_cuda_err = std::string("cudaMalloc(arena): ") + cudaGetErrorString(e); ix->arena = nullptr; rig_index_destroy(ix); return RIG_ERR_NOMEM; }
    CU_TRY_IX(cudaMemset(ix->arena, 0, total));
    for (auto& p : parts)
        if (p.bytes) CU_TRY_IX(cudaMemcpy((char*)ix->arena + p.off, p.src, p.bytes, cudaMemcpyHostToDevice));
    char* A = (char*)ix->arena;
    FlatDev& d = ix->d;
    d.n = f.n; d.r = f.r; d.nblk = f.nblk; d.toe0 = f.toe0;
    d.K = f.K; d.S = f.S; d.lf_shift = f.lf_shift; d.pad0 = 0;
    d.F = (const ull*)(A + parts[0].off);
    d.sid = (const uint16_t*)(A + parts[1].off);
    d.start = (const void*)(A + parts[2].off);
    d.blk = (const char*)(A + parts[3].off);
    d.blk_stride = f.blk_stride; d.off_head = f.off_head; d.off_cum = f.off_cum; d.rec_w = f.rec_w;
    d.bstart = (const void*)(A + parts[4].off);
    d.last = (const void*)(A + parts[5].off);
    d.bdir = (const uint32_t*)(A + parts[6].off);
    d.samples_last = (const void*)(A + parts[7].off);
    d.phi.rec = (const void*)(A + parts[8].off);
    d.phi.pent = (const void*)(A + parts[9].off);
    d.phi.shift = f.phi.shift; d.phi.D = f.phi.D;
    d.phi.packed = f.phi_packed ? 1u : 0u;
    d.phi.esz = f.phi_packed ? 32u : f.phi.RW * (f.w32 ? 4u : 8u);
    d.seed.rec = (const void*)(A + parts[10].off);
    d.seed.pent = (const void*)(A + parts[11].off);
    d.seed.shift = f.seed.shift; d.seed.J = f.seed.J > 1 ? f.seed.J : 0;
    ix->phi_bytes = (parts[9].off + parts[9].bytes) - parts[8].off;
    ix->lf_bytes = parts[8].off;  // F, sid, start, block records, bstart, last, bdir, samples_last
    d.w32 = f.w32 ? 1u : 0u; d.pad = 0;
#ifdef RIG_WINDOW_DIAG   // window-pass diagnostics (bits 15-17): stores redirected / no stores / no dependent lookups
    d.pad = (variant & 32768) ? 1u : ((variant & 65536) ? 2u : ((variant & 131072) ? 3u : 0u));
#endif
    d.dbg = nullptr;
    if (d.pad) { void* q = nullptr; if (cudaMalloc(&q, 32u << 20) == cudaSuccess) d.dbg = (ull*)q; else d.pad = 0; }

    // L2 persistence for the Phi records: reserve the largest carve-out the device allows (device-wide
    // limit; harmless for other users of the context) and size the window / hit ratio to it.
    {
        const size_t rec_bytes = ix->phi_bytes;  // bucket records + piece entries of the Phi^1..D table
        // Measured on C2 (B200, 79 MiB max carve-out): reserving persisting L2 made the expansion kernel
        // SLOWER (0.51 -> 0.91 ms; the carve-out shrinks the L2 left for the output stream and the other
        // arrays), so it is off unless RIG_VARIANT bit0 asks for the experiment.
        if ((variant & 1) && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0 && rec_bytes > 0) {
            size_t carve = (rec_bytes + (rec_bytes >> 2) + (1u << 20)) & ~(size_t)((1u << 20) - 1);  // table + 25%
            if (carve > (size_t)prop.persistingL2CacheMaxSize) carve = (size_t)prop.persistingL2CacheMaxSize;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
                ix->l2_window_bytes = rec_bytes < (size_t)prop.accessPolicyMaxWindowSize ? rec_bytes : (size_t)prop.accessPolicyMaxWindowSize;
                const double ratio = (double)carve / (double)ix->l2_window_bytes;
                ix->l2_hit_ratio = ratio >= 1.0 ? 1.f : (float)ratio;
            } else {
                cudaGetLastError();
            }
        }
    }
    if ((rc = finish_create(ix)) != RIG_OK) { rig_index_destroy(ix); return rc; }
    ix->arena_bytes = total;
    ix->digest = logical_digest(*view);

    rig_index_info& I = ix->info;
    std::memset(&I, 0, sizeof(I));
    I.n = f.n; I.r = f.r; I.sigma = f.S; I.device_bytes = total;
    I.lf_blocks = f.nblk; I.lf_buckets = f.lf_nbkt; I.phi_buckets = f.phi.nbkt;
    I.runs_per_block = f.K; I.lf_shift = f.lf_shift; I.phi_shift = f.phi.shift;
    I.phi_jump = f.phi.D; I.phi_jump_pieces = f.phi.pieces(); I.words32 = f.w32 ? 1u : 0u; I.lf_record_bytes = f.blk_stride;
    I.device = (uint32_t)device; I.sm_count = (uint32_t)ix->sm_count;
    I.seed_jump = d.seed.J; I.seed_shift = f.seed.shift; I.seed_pieces = f.seed.pieces(); I.seed_bytes = f.seed.bytes(f.w32);
    *out = ix;
    return RIG_OK;
}

// ---- the flattened index as a file: flatten once, load in seconds ---------------------------------------------
// Layout: FlatFileHeader | the device arena, byte for byte. The FlatDev in the header holds OFFSETS into the arena
// instead of addresses. A file is accepted only by the library version that wrote it (magic + struct sizes).
namespace {
struct FlatFileHeader {
    char magic[8];               // "RIGFLAT2"
    uint32_t header_bytes, flatdev_bytes, info_bytes, reserved;
    uint64_t arena_bytes, digest, phi_bytes, lf_bytes;
    rig_options opt;
    rig_index_info info;
    FlatDev d;                   // pointers rebased to offsets
};
}  // namespace

int rig_index_save_flat(const rig_index* ix, const char* path) {
    if (!ix || !path) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    FlatFileHeader h;
    std::memset(&h, 0, sizeof(h));
    std::memcpy(h.magic, "RIGFLAT2", 8);
    h.header_bytes = sizeof(FlatFileHeader); h.flatdev_bytes = sizeof(FlatDev); h.info_bytes = sizeof(rig_index_info);
    h.arena_bytes = ix->arena_bytes; h.digest = ix->digest; h.phi_bytes = ix->phi_bytes; h.lf_bytes = ix->lf_bytes;
    h.opt = ix->opt; h.info = ix->info; h.d = ix->d;
    const char* A = (const char*)ix->arena;
    for_each_pointer(h.d, [&](const void*& p) { p = (const void*)(uintptr_t)((const char*)p - A); });
    FILE* f = std::fopen(path, "wb");
    if (!f) return RIG_ERR_ARG;
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1;
    const size_t CH = 64u << 20;
    void* stage = nullptr;
    if (cudaMallocHost(&stage, CH) != cudaSuccess) { cudaGetLastError(); std::fclose(f); return RIG_ERR_NOMEM; }
    for (size_t off = 0; ok && off < ix->arena_bytes; off += CH) {
        const size_t k = std::min(CH, ix->arena_bytes - off);
        if (cudaMemcpy(stage, A + off, k, cudaMemcpyDeviceToHost) != cudaSuccess) { ok = false; break; }
        ok = std::fwrite(stage, 1, k, f) == k;
    }
    cudaFreeHost(stage);
    ok = (std::fclose(f) == 0) && ok;
    return ok ? RIG_OK : RIG_ERR_CUDA;
}

int rig_index_load_flat(const char* path, const rig_logical_view* check, int device, rig_index** out) {
    if (!path || !out) return RIG_ERR_ARG;
    *out = nullptr;
    int ndev = rig_device_count();
    if (device < 0 || device >= ndev) return RIG_ERR_NO_DEVICE;
    FILE* f = std::fopen(path, "rb");
    if (!f) return RIG_ERR_ARG;
    FlatFileHeader h;
    if (std::fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "RIGFLAT2", 8) != 0 || h.header_bytes != sizeof(FlatFileHeader) ||
        h.flatdev_bytes != sizeof(FlatDev) || h.info_bytes != sizeof(rig_index_info) || h.arena_bytes == 0) {
        std::fclose(f);
        return RIG_ERR_INDEX;
    }
    if (check && (h.d.n != check->n || h.d.r != check->r || h.digest != logical_digest(*check))) { std::fclose(f); return RIG_ERR_INDEX; }
    // every offset must lie inside the arena
    bool sane = true;
    for_each_pointer(h.d, [&](const void*& p) { if ((uintptr_t)p >= h.arena_bytes) sane = false; });
    if (!sane || h.phi_bytes > h.arena_bytes || h.lf_bytes > h.arena_bytes) { std::fclose(f); return RIG_ERR_INDEX; }
    CU_TRY(cudaSetDevice(device));
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b));
    if (h.arena_bytes + (64u << 20) > free_b) { std::fclose(f); return RIG_ERR_NOMEM; }
    rig_index* ix = new (std::nothrow) rig_index();
    if (!ix) { std::fclose(f); return RIG_ERR_NOMEM; }
    ix->device = device; ix->opt = h.opt;
    if (const char* ev = getenv("RIG_VARIANT")) ix->variant = atoi(ev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); std::fclose(f); delete ix; return RIG_ERR_CUDA; }
    ix->sm_count = prop.multiProcessorCount;
    if (cudaMalloc(&ix->arena, h.arena_bytes) != cudaSuccess) { cudaGetLastError(); std::fclose(f); delete ix; return RIG_ERR_NOMEM; }
    const size_t CH = 64u << 20;
    void* stage[2] = {nullptr, nullptr};
    bool ok = cudaMallocHost(&stage[0], CH) == cudaSuccess && cudaMallocHost(&stage[1], CH) == cudaSuccess;
    cudaStream_t st = nullptr;
    ok = ok && cudaStreamCreate(&st) == cudaSuccess;
    cudaEvent_t evs[2] = {nullptr, nullptr};
    ok = ok && cudaEventCreate(&evs[0]) == cudaSuccess && cudaEventCreate(&evs[1]) == cudaSuccess;
    int b = 0;
    for (size_t off = 0; ok && off < h.arena_bytes; off += CH, b ^= 1) {   // read chunk k+1 from the file while chunk k uploads
        const size_t k = std::min(CH, (size_t)h.arena_bytes - off);
        ok = cudaEventSynchronize(evs[b]) == cudaSuccess;                  // the buffer's previous upload has drained
        ok = ok && std::fread(stage[b], 1, k, f) == k;
        ok = ok && cudaMemcpyAsync((char*)ix->arena + off, stage[b], k, cudaMemcpyHostToDevice, st) == cudaSuccess;
        ok = ok && cudaEventRecord(evs[b], st) == cudaSuccess;
    }
    if (st) { ok = (cudaStreamSynchronize(st) == cudaSuccess) && ok; cudaStreamDestroy(st); }
    for (auto& e : evs) if (e) cudaEventDestroy(e);
    for (auto& p : stage) if (p) cudaFreeHost(p);
    std::fclose(f);
    if (!ok) { cudaGetLastError(); rig_index_destroy(ix); return RIG_ERR_CUDA; }
    ix->d = h.d;
    char* A = (char*)ix->arena;
    for_each_pointer(ix->d, [&](const void*& p) { p = (const void*)(A + (uintptr_t)p); });
    ix->info = h.info; ix->info.device = (uint32_t)device; ix->info.sm_count = (uint32_t)ix->sm_count;
    ix->phi_bytes = h.phi_bytes; ix->lf_bytes = h.lf_bytes; ix->arena_bytes = h.arena_bytes; ix->digest = h.digest;
    int rc = finish_create(ix);
    if (rc != RIG_OK) { rig_index_destroy(ix); return rc; }
    *out = ix;
    return RIG_OK;
}

void rig_index_destroy(rig_index* ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    for (DevBuf* b : {&ix->toe, &ix->jl, &ix->nch, &ix->nocc, &ix->choff, &ix->sums, &ix->patt, &ix->lo, &ix->hi,
                      &ix->occoff, &ix->occ, &ix->items, &ix->occ32, &ix->text, &ix->big1, &ix->big2, &ix->ctable, &ix->crep, &ix->cfound})
        b->release();
    if (ix->arena) cudaFree(ix->arena);
    if (ix->d_counters) cudaFree(ix->d_counters);
    if (ix->h_counters) cudaFreeHost(ix->h_counters);
    if (ix->h_plan) cudaFreeHost(ix->h_plan);
    ix->d_plan.release();
    if (ix->d_post) cudaFree(ix->d_post);
    if (ix->h_post) cudaFreeHost(ix->h_post);
    for (auto& ev : ix->ev) if (ev) cudaEventDestroy(ev);
    if (ix->ev_scan) cudaEventDestroy(ix->ev_scan);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
}

int rig_index_info_get(const rig_index* ix, rig_index_info* info) {
    if (!ix || !info) return RIG_ERR_ARG;
    *info = ix->info;
    return RIG_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
namespace {

// workspace of the search kernel's fused offset scan: ticket + RIG_TILE_WORDS words per tile of 128 patterns
size_t tile_ws_words(uint64_t N) { return 2 + RIG_TILE_WORDS * ((N + 127) / 128) + 2; }

// Start of a batch call: zero the counters and (locate) the offset-scan workspace, warm the backward-search
// structures (the head of the arena: F, sid, run starts, block records, directories, samples) into L2 when they are
// small enough to stay there (<= 48 MB) — one kernel (rigk::prep_kernel). RIG_VARIANT bit 12: no warm-up (A/B switch).
int prep_call(rig_index* ix, uint64_t N, bool locate, cudaStream_t st) {
    const bool lane_kernel = ix->d.K == 4 && !(ix->variant & 128);
    ull* ws = (locate && lane_kernel) ? (ull*)ix->sums.p : nullptr;
    const uint64_t nws = ws ? tile_ws_words(N) : 0;
    const uint64_t warm = ((ix->variant & 4096) || ix->lf_bytes > (48u << 20)) ? 0 : ix->lf_bytes;
    const uint64_t work = std::max<uint64_t>(std::max<uint64_t>(nws, 16), (warm + 127) / 128);
    const unsigned nb = (unsigned)std::min<uint64_t>((work + 255) / 256, (uint64_t)ix->sm_count * 8);
    rigk::prep_kernel<<<nb, 256, 0, st>>>(ix->d_counters, 16, ws, nws, (const char*)ix->arena, warm);
    CU_TRY(cudaGetLastError());
    ix->timing.launches += 1;
    return RIG_OK;
}

template <bool LOCATE>
int launch_search(rig_index* ix, const uint8_t* d_patt, uint64_t N, uint64_t m, ull* d_lo, ull* d_hi, ull* d_occoff,
                  cudaStream_t st) {
    const uint32_t G = ix->d.K;
    ull* toe = (ull*)ix->toe.p; ull* jl = (ull*)ix->jl.p; ull* nch = (ull*)ix->nch.p; ull* nocc = (ull*)ix->nocc.p;
    ull* steps = ix->d_counters + 0;
    const bool n32 = ix->d.w32 != 0;
    if (G == 4 && !(ix->variant & 128)) {  // one lane per pattern over the K = 4 block records (bit7: cooperative kernel, A/B switch)
        const int lt = 128;
        const uint64_t lb = (N + lt - 1) / lt;
        if (lb > 0x7fffffffull) return RIG_ERR_ARG;
        ull* ws = (ull*)ix->sums.p; ull* choff = (ull*)ix->choff.p; ull* totals = ix->d_counters + RIG_CTR_TOTAL;
        if (ix->variant & 16384) {   // one lane per pattern (A/B switch); default: two lanes per pattern, 256-thread CTAs
            if (n32) rigk::search_lane_kernel<LOCATE, uint32_t><<<(unsigned)lb, lt, 0, st>>>(ix->d, d_patt, N, m, d_lo, d_hi, toe, jl, choff, d_occoff, steps, ws, totals);
            else rigk::search_lane_kernel<LOCATE, ull><<<(unsigned)lb, lt, 0, st>>>(ix->d, d_patt, N, m, d_lo, d_hi, toe, jl, choff, d_occoff, steps, ws, totals);
        } else {
            if (n32) rigk::search_pair_kernel<LOCATE, uint32_t><<<(unsigned)lb, 2 * lt, 0, st>>>(ix->d, d_patt, N, m, d_lo, d_hi, toe, jl, choff, d_occoff, steps, ws, totals);
            else rigk::search_pair_kernel<LOCATE, ull><<<(unsigned)lb, 2 * lt, 0, st>>>(ix->d, d_patt, N, m, d_lo, d_hi, toe, jl, choff, d_occoff, steps, ws, totals);
        }
        CU_TRY(cudaGetLastError());
        ix->timing.launches += 1;
        return RIG_OK;
    }
    const int threads = 256;
    const uint64_t ppw = 32 / (2 * G);
    const uint64_t warps = (N + ppw - 1) / ppw;
    const uint64_t blocks = (warps + (threads / 32) - 1) / (threads / 32);
    if (blocks > 0x7fffffffull) return RIG_ERR_ARG;
#define RIG_LAUNCH(GG)                                                                                          \
    do {                                                                                                        \
        if (n32) rigk::search_kernel<GG, LOCATE, uint32_t><<<(unsigned)blocks, threads, 0, st>>>(               \
            ix->d, d_patt, N, m, d_lo, d_hi, toe, jl, nch, nocc, steps);                                        \
        else rigk::search_kernel<GG, LOCATE, ull><<<(unsigned)blocks, threads, 0, st>>>(                        \
            ix->d, d_patt, N, m, d_lo, d_hi, toe, jl, nch, nocc, steps);                                        \
    } while (0)
    switch (G) {
        case 4: RIG_LAUNCH(4); break;
        case 8: RIG_LAUNCH(8); break;
        case 16: RIG_LAUNCH(16); break;
        default: return RIG_ERR_ARG;
    }
#undef RIG_LAUNCH
    CU_TRY(cudaGetLastError());
    ix->timing.launches += 1;
    if (LOCATE) {  // the cooperative kernel leaves n_occ / chain counts per pattern: three-kernel scan
        const uint64_t ntiles = (N + RIG_SCAN_TILE - 1) / RIG_SCAN_TILE;
        ull* sums = (ull*)ix->sums.p;
        rigk::scan_tile_sums<<<(unsigned)ntiles, RIG_SCAN_THREADS, 0, st>>>(nocc, nch, N, sums, ntiles);
        rigk::scan_sums_inplace<<<1, 1024, 0, st>>>(sums, ntiles, ix->d_counters + RIG_CTR_TOTAL);
        rigk::scan_tiles<<<(unsigned)ntiles, RIG_SCAN_THREADS, 0, st>>>(nocc, nch, N, sums, ntiles, d_occoff, (ull*)ix->choff.p,
                                                                          ix->d_counters + RIG_CTR_TOTAL, nullptr);
        CU_TRY(cudaGetLastError());
        ix->timing.launches += 3;
    }
    return RIG_OK;
}

int finish_timing(rig_index* ix) {
    if (!ix->timing_pending) return RIG_OK;
    CU_TRY(cudaSetDevice(ix->device));
    // the last recorded event closes the call
    for (int i = 5; i >= 0; --i)  // ev[6], ev[7] sit between ev[3] and ev[4] in stream order
        if (ix->ev_valid[i]) { CU_TRY(cudaEventSynchronize(ix->ev[i])); break; }
    auto el = [&](int a, int b, float& dst) -> int {
        dst = 0.f;
        if (ix->ev_valid[a] && ix->ev_valid[b]) CU_TRY(cudaEventElapsedTime(&dst, ix->ev[a], ix->ev[b]));
        return RIG_OK;
    };
    int rc;
    if ((rc = el(0, 1, ix->timing.h2d_ms))) return rc;
    if ((rc = el(1, 2, ix->timing.search_ms))) return rc;
    if ((rc = el(2, 3, ix->timing.scan_ms))) return rc;
    if ((rc = el(3, 4, ix->timing.expand_ms))) return rc;
    if ((rc = el(4, 5, ix->timing.d2h_ms))) return rc;
    if ((rc = el(3, 6, ix->timing.seed_ms))) return rc;
    if ((rc = el(6, ix->ev_valid[7] ? 7 : 4, ix->timing.window_ms))) return rc;
    ix->timing.lf_steps = ix->h_counters[0];
    ix->timing_pending = false;
    return RIG_OK;
}

void begin_call(rig_index* ix, bool keep_plan = false) {
    if (!keep_plan) ix->plan_valid = false;   // any other batch call reuses the buffers a planned batch lives in
    std::memset(&ix->timing, 0, sizeof(ix->timing));
    for (bool& b : ix->ev_valid) b = false;
    ix->timing_pending = true;
}

int rec(rig_index* ix, int i, cudaStream_t st) {
    CU_TRY(cudaEventRecord(ix->ev[i], st));
    ix->ev_valid[i] = true;
    return RIG_OK;
}

// count on device buffers; events 1..2 bracket the kernel
int count_dev(rig_index* ix, const uint8_t* d_patt, uint64_t N, uint64_t m, ull* d_lo, ull* d_hi, cudaStream_t st) {
    int rc;
    if ((rc = prep_call(ix, N, false, st))) return rc;
    if ((rc = rec(ix, 1, st))) return rc;
    if (N && (rc = launch_search<false>(ix, d_patt, N, m, d_lo, d_hi, nullptr, st))) return rc;
    if ((rc = rec(ix, 2, st))) return rc;
    CU_TRY(cudaMemcpyAsync(ix->h_counters, ix->d_counters, 8 * sizeof(ull), cudaMemcpyDeviceToHost, st));
    return RIG_OK;
}

int sort_dev(rig_index* ix, uint64_t N, const ull* d_off, ull* d_occ, uint64_t total, cudaStream_t st);
int check_dev(rig_index* ix, const uint8_t* d_patt, uint64_t N, uint64_t m, const ull* d_lo, const ull* d_hi,
              const ull* d_off, const ull* d_occ, uint64_t total, int sorted, rig_check_report* report, cudaStream_t st);

// resident CTAs of a kernel on this device (persistent grids are sized SMs x this)
template <typename K>
int resident_ctas(K kernel, int threads) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
    return nb;
}

// Phi expansion of the batch whose offsets and totals the search left on the device: seed pass + window pass, or the
// single-pass walk. Both kernels are persistent and decide on the device whether to run (expansion_enabled): the
// host queues them WITHOUT knowing the totals. `items_cap`: entries the item list can hold.
int launch_expansion(rig_index* ix, uint64_t N, const ull* d_lo, const ull* d_hi, const ull* d_occoff, void* d_occ_v,
                     uint64_t cap, bool two_pass, uint64_t items_cap, cudaStream_t st, bool out32, uint64_t p0 = 0,
                     ull* ctr = nullptr) {   // p0, ctr: a shard of a planned batch (d_lo / d_hi / d_occoff already point at pattern p0)
    int rc;
    const int threads = ix->opt.expand_threads ? (int)ix->opt.expand_threads : 128;  // measured: 0.445 ms (128) vs 0.463 ms (256) on C2
    if (threads < 32 || threads > 256 || (threads & 31)) return RIG_ERR_ARG;
    const bool w32 = ix->d.w32 != 0;
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3((unsigned)threads); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (ix->l2_window_bytes) {  // persisting-L2 experiment (RIG_VARIANT bit 0)
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(ix->d.phi.rec);
        attr[0].val.accessPolicyWindow.num_bytes = ix->l2_window_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = ix->l2_hit_ratio;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    const ull* a_choff = (const ull*)ix->choff.p + p0;
    const ull* a_toe = (const ull*)ix->toe.p + p0; const ull* a_jl = (const ull*)ix->jl.p + p0;
    ull* a_ctr = ctr ? ctr : ix->d_counters;
    ull a_N = N, a_cap = cap, a_icap = items_cap;
    ull* a_items = (ull*)ix->items.p;
    const bool keep = (ix->variant & 2) == 0;  // L2::evict_last on the Phi entry loads (bit1 disables: A/B switch)
    const uint32_t SEG = ix->d.seed.J;
    uint32_t seg_shift = 0;
    while ((1u << seg_shift) < SEG) ++seg_shift;
    // the chains / items of the batch are not known here: persistent grids, bounded by what the batch can hold
    // (a chain per occurrence at most; total <= cap when the kernels run at all)
    const uint64_t max_chains = cap ? cap : 1;
    const uint64_t max_items = items_cap ? items_cap : 1;
    const int wthreads = 256;
    // Fused expansion (default): one persistent kernel produces the items and consumes them as they appear.
    // RIG_VARIANT bit 13 keeps the two kernels apart (A/B switch, and the per-pass timings of rig_timing).
    const bool fused = two_pass && !(ix->variant & 8192) && ix->d.n < (1ull << 48);
    ull a_tag = 0;
    bool fused_ok = false;
    uint32_t a_pmod = 1;   // every CTA produces its share first (measured: dedicating 1/2, 1/4, 1/8 of the CTAs to production is slower; RIG_FUSED_PROD overrides)
    if (const char* ev = getenv("RIG_FUSED_PROD")) { int v = atoi(ev); if (v >= 1 && v <= 64) a_pmod = (uint32_t)v; }
    if (fused) {
        if (ix->items_zeroed != ix->items.generation || ((ix->epoch + 1) & 0xFFFFu) == 0) {   // fresh allocation, or the 16-bit tag wraps
            CU_TRY(cudaMemsetAsync(ix->items.p, 0, ix->items.cap, st));
            ix->items_zeroed = ix->items.generation;
            ix->epoch = 0;
        }
        ix->epoch += 1;
        a_tag = (ull)ix->epoch << 48;
    }
#define RIG_EXPAND2(W, DD, KP, OTT)                                                                             \
    do {                                                                                                        \
        OTT* d_occ = (OTT*)d_occ_v;                                                                             \
        if (fused) {                                                                                            \
            auto kf = rigk::phi_fused_kernel<W, DD, KP, (sizeof(W) == 4 ? 5 : 3), OTT>;                           \
            const int per_sm = resident_ctas(kf, wthreads);                                                     \
            uint64_t gf = (uint64_t)ix->sm_count * per_sm;                                                      \
            cudaLaunchConfig_t cfgf = cfg;                                                                      \
            cfgf.blockDim = dim3((unsigned)wthreads);                                                           \
            /* consumers wait for producers: the grid must be co-resident — a COOPERATIVE launch guarantees it or fails */ \
            cudaLaunchAttribute fattr[2];                                                                       \
            unsigned nfa = 0;                                                                                   \
            if (cfg.numAttrs) fattr[nfa++] = attr[0];                                                           \
            fattr[nfa].id = cudaLaunchAttributeCooperative; fattr[nfa].val.cooperative = 1; ++nfa;              \
            cfgf.attrs = fattr; cfgf.numAttrs = nfa;                                                            \
            if ((rc = rec(ix, 6, st))) return rc;                                                               \
            cudaError_t fe = cudaErrorUnknown;                                                                  \
            /* "too large" can come from resources the occupancy query does not see: retry with one CTA per SM less */ \
            for (int per = per_sm; per >= 1 && per + 2 >= per_sm; --per) {                                      \
                gf = (uint64_t)ix->sm_count * per;                                                              \
                cfgf.gridDim = dim3((unsigned)gf);                                                              \
                fe = cudaLaunchKernelEx(&cfgf, kf, ix->d, a_N, a_choff, d_occoff, d_lo, d_hi, a_toe, a_jl, d_occ, \
                                        a_ctr, a_cap, a_items, a_icap, seg_shift, a_tag, a_pmod);               \
                if (fe != cudaErrorCooperativeLaunchTooLarge) break;                                            \
                cudaGetLastError();                                                                             \
            }                                                                                                   \
            if (fe == cudaSuccess) {                                                                            \
                ix->timing.launches += 1; ix->timing.slices = 1;                                                \
                if ((rc = rec(ix, 7, st))) return rc;                                                           \
                fused_ok = true;                                                                                \
            } else {                                                                                            \
                /* not co-resident on this device right now: the two-kernel form below */                       \
                if (!ix->warned_fused) {                                                                        \
                    fprintf(stderr, "[rindex_gpu] fused expansion kernel not launched (%s, grid %llu x %d); "   \
                            "falling back to seed pass + window pass\n", cudaGetErrorString(fe),                \
                            (unsigned long long)gf, wthreads);                                                  \
                    ix->warned_fused = true;                                                                    \
                }                                                                                               \
                cudaGetLastError();                                                                             \
            }                                                                                                   \
        }                                                                                                       \
        if (fused_ok) {                                                                                         \
        } else if (two_pass) {                                                                                  \
            auto k1 = rigk::phi_expand_kernel<W, DD, KP, true, OTT>;                                             \
            auto k2 = rigk::phi_window_batch_kernel<W, DD, KP, (sizeof(W) == 4 ? 6 : 4), OTT>;                  \
            uint64_t g1 = (uint64_t)ix->sm_count * resident_ctas(k1, threads);                                  \
            uint64_t g2 = (uint64_t)ix->sm_count * resident_ctas(k2, wthreads);                                 \
            g1 = std::min<uint64_t>(g1, (max_chains + threads - 1) / threads);                                  \
            g2 = std::min<uint64_t>(g2, (max_items + wthreads - 1) / wthreads);                                 \
            cfg.gridDim = dim3((unsigned)g1);                                                                   \
            CU_TRY(cudaLaunchKernelEx(&cfg, k1, ix->d, a_N, a_choff, d_occoff, d_lo, d_hi, a_toe, a_jl, d_occ,  \
                                      a_ctr, a_cap, a_items, a_icap, seg_shift));                               \
            if ((rc = rec(ix, 6, st))) return rc;                                                               \
            cudaLaunchConfig_t cfg2 = cfg;                                                                      \
            cfg2.gridDim = dim3((unsigned)g2); cfg2.blockDim = dim3((unsigned)wthreads);                        \
            const ull* c_items = a_items; const ull* c_ctr = a_ctr;                                             \
            CU_TRY(cudaLaunchKernelEx(&cfg2, k2, ix->d, c_items, c_ctr, d_occ, a_cap, a_icap, seg_shift));      \
            ix->timing.launches += 2; ix->timing.slices = 2;                                                    \
            if ((rc = rec(ix, 7, st))) return rc;                                                               \
        } else {                                                                                                \
            auto k1 = rigk::phi_expand_kernel<W, DD, KP, false, OTT>;                                            \
            uint64_t g1 = (uint64_t)ix->sm_count * resident_ctas(k1, threads);                                  \
            g1 = std::min<uint64_t>(g1, (max_chains + threads - 1) / threads);                                  \
            cfg.gridDim = dim3((unsigned)g1);                                                                   \
            CU_TRY(cudaLaunchKernelEx(&cfg, k1, ix->d, a_N, a_choff, d_occoff, d_lo, d_hi, a_toe, a_jl, d_occ,  \
                                      a_ctr, a_cap, a_items, a_icap, seg_shift));                               \
            ix->timing.launches += 1;                                                                           \
        }                                                                                                       \
    } while (0)
#define RIG_EXPAND(W, DD)                                                                                     \
    do {                                                                                                      \
        if (keep) RIG_EXPAND2(W, DD, true, ull);                                                              \
        else RIG_EXPAND2(W, DD, false, ull);                                                                  \
    } while (0)
    if (out32) {   // native 32-bit positions (rig_locate_batch32): the default table form only
        if (!(ix->d.phi.D == 4 && w32)) return RIG_ERR_ARG;
        if (keep) RIG_EXPAND2(uint32_t, 4, true, uint32_t);
        else RIG_EXPAND2(uint32_t, 4, false, uint32_t);
    } else
    switch (ix->d.phi.D * 2 + (w32 ? 1 : 0)) {
        case 2: RIG_EXPAND(ull, 1); break;
        case 3: RIG_EXPAND(uint32_t, 1); break;
        case 4: RIG_EXPAND(ull, 2); break;
        case 5: RIG_EXPAND(uint32_t, 2); break;
        case 8: RIG_EXPAND(ull, 4); break;
        case 9: RIG_EXPAND(uint32_t, 4); break;
        case 13: RIG_EXPAND(uint32_t, 6); break;
        case 16: RIG_EXPAND(ull, 8); break;
        case 17: RIG_EXPAND(uint32_t, 8); break;
        default: return RIG_ERR_ARG;
    }
#undef RIG_EXPAND2
#undef RIG_EXPAND
    CU_TRY(cudaGetLastError());
    return RIG_OK;
}

// search (+ offsets) and expansion, queued back to back; the host then waits for the totals only (they are copied
// out right after the search), while the expansion is already running: events 1..4 on `st`.
// own_occ (host-buffer entry points with RIG_LOCATE_DEVICE_ONLY): the occurrences go to the library's own buffer,
// which is grown to the batch total when it is too small — only the expansion is queued again, not the search.
// out32: d_occ is an array of 32-bit positions (n < 2^32; the default Phi table form), cap counts its elements.
int locate_dev(rig_index* ix, const uint8_t* d_patt, uint64_t N, uint64_t m, ull* d_lo, ull* d_hi, ull* d_occoff,
               void* d_occ, uint64_t cap, uint64_t* occ_total, cudaStream_t st, DevBuf* own_occ = nullptr, bool out32 = false) {
    int rc;
    if ((rc = ix->toe.ensure((N + 1) * 8)) || (rc = ix->jl.ensure((N + 1) * 8)) || (rc = ix->nch.ensure((N + 1) * 8)) ||
        (rc = ix->nocc.ensure((N + 1) * 8)) || (rc = ix->choff.ensure((N + 4) * 8)) ||
        (rc = ix->sums.ensure((std::max<uint64_t>(tile_ws_words(N), 2 * ((N + RIG_SCAN_TILE - 1) / RIG_SCAN_TILE) + 8)) * 8)))
        return rc;
    if (own_occ) { d_occ = own_occ->p; cap = own_occ->cap / 8; }
    const uint32_t SEG = ix->d.seed.J;
    auto items_bound = [&](uint64_t total, uint64_t chains) { return total / (SEG ? SEG : 1) + chains; };
    // Two passes when the index has a seed table and the output array is line-aligned (the window kernel writes
    // whole 128-byte lines); otherwise the single-pass walk. The item list is sized BEFORE the totals are known,
    // from the capacity the caller offers (total <= cap or nothing runs), the list kept from earlier calls, and a
    // guess of two chains per pattern; the kernels check the exact bound on the device, and the host queues them
    // again below if the guess was short. `exact` = the totals are known (re-launch).
    auto queue_expansion = [&](uint64_t exact_items) -> int {
        const bool two_pass = SEG > 1 && !(ix->variant & 32) && ((reinterpret_cast<uintptr_t>(d_occ) & 127) == 0);
        uint64_t items_cap = 0;
        if (two_pass) {
            const uint64_t want_items = exact_items ? exact_items + 64 : items_bound(std::min<uint64_t>(cap, 1ull << 33), 2 * N + 1024) + 32;
            if (exact_items && want_items >= (1ull << 31)) return RIG_ERR_ARG;  // 32-bit item indices in the window pass: > 1.3e11 occurrences in one call
            if (ix->items.cap < want_items * 16) {
                int r2 = ix->items.ensure(exact_items ? want_items * 16 : std::min<uint64_t>(want_items * 16, 1ull << 30));
                if (r2) return r2;
            }
            items_cap = ix->items.cap / 16 - 32;
        }
        ix->last_items_cap = items_cap; ix->last_two_pass = two_pass;
        return launch_expansion(ix, N, d_lo, d_hi, d_occoff, d_occ, cap, two_pass, items_cap, st, out32);
    };
    if ((rc = prep_call(ix, N, true, st))) return rc;
    if ((rc = rec(ix, 1, st))) return rc;
    if (N) {
        if ((rc = launch_search<true>(ix, d_patt, N, m, d_lo, d_hi, d_occoff, st))) return rc;
    } else {
        CU_TRY(cudaMemsetAsync(d_occoff, 0, 8, st));
    }
    if ((rc = rec(ix, 2, st))) return rc;
    CU_TRY(cudaMemcpyAsync(ix->h_counters, ix->d_counters, 8 * sizeof(ull), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(ix->ev_scan, st));
    if ((rc = rec(ix, 3, st))) return rc;
    bool queued = false;
    if (d_occ != nullptr && cap > 0 && N > 0) {
        // warm the Phi tables into L2 (RIG_VARIANT bit2 disables: A/B switch)
        if (!(ix->variant & 4) && ix->phi_bytes) {
            const uint64_t lines = (ix->phi_bytes + 127) / 128;
            rigk::l2_warm_kernel<<<(unsigned)((lines + 255) / 256), 256, 0, st>>>((const char*)ix->d.phi.rec, ix->phi_bytes);
            ix->timing.launches += 1;
        }
        if ((rc = queue_expansion(0))) return rc;
        queued = true;
    }
    CU_TRY(cudaEventSynchronize(ix->ev_scan));  // the expansion is queued (or running) behind it
    const uint64_t total = ix->h_counters[RIG_CTR_TOTAL], chains = ix->h_counters[RIG_CTR_CHAINS];
    bool fits = !(total > cap || (total && !d_occ));
    if (!fits && own_occ && total) {   // grow the library's buffer to this batch; the search results stay valid on the device
        if ((rc = own_occ->ensure(total * 8 + 128))) return rc;
        d_occ = own_occ->p; cap = own_occ->cap / 8;
        fits = true; queued = false;
    }
    const bool need_items = fits && total && ix->d.seed.J > 1 && !(ix->variant & 32) && ((reinterpret_cast<uintptr_t>(d_occ) & 127) == 0);
    if (fits && total && (!queued || (need_items && items_bound(total, chains) > ix->last_items_cap))) {
        // not queued yet (buffer just grown), or the kernels returned at once because the item list was too short
        // (same test on the device): queue the expansion (again) with exact sizes
        CU_TRY(cudaMemsetAsync(ix->d_counters + RIG_CTR_ITEMS, 0, 3 * sizeof(ull), st));  // items, producers done, items taken
        if ((rc = queue_expansion(need_items ? items_bound(total, chains) : 0))) return rc;
    }
    ix->timing.occ_total = total;
    ix->timing.chains = chains;
    if (occ_total) *occ_total = total;
    if ((rc = rec(ix, 4, st))) return rc;
    return fits ? RIG_OK : RIG_ERR_CAPACITY;
}

// Replicated planning of a job sharded over several GPUs: search + offsets of the WHOLE batch here, cut points of equal
// work from the offsets (a binary search per cut), one small D2H. The search results stay in the index's buffers.
int plan_dev(rig_index* ix, const uint8_t* d_patt, uint64_t N, uint64_t m, ull* d_lo, ull* d_hi, ull* d_occoff, uint32_t shards,
             uint64_t cost, uint64_t* cuts, uint64_t* occ_total, cudaStream_t st) {
    int rc;
    if ((rc = ix->toe.ensure((N + 1) * 8)) || (rc = ix->jl.ensure((N + 1) * 8)) || (rc = ix->nch.ensure((N + 1) * 8)) ||
        (rc = ix->nocc.ensure((N + 1) * 8)) || (rc = ix->choff.ensure((N + 4) * 8)) ||
        (rc = ix->sums.ensure((std::max<uint64_t>(tile_ws_words(N), 2 * ((N + RIG_SCAN_TILE - 1) / RIG_SCAN_TILE) + 8)) * 8)) ||
        (rc = ix->d_plan.ensure(3 * 1025 * 8)))
        return rc;
    if (!ix->h_plan) CU_TRY(cudaMallocHost((void**)&ix->h_plan, 3 * 1025 * sizeof(ull)));
    if ((rc = prep_call(ix, N, true, st))) return rc;
    if ((rc = rec(ix, 1, st))) return rc;
    if (N) {
        if ((rc = launch_search<true>(ix, d_patt, N, m, d_lo, d_hi, d_occoff, st))) return rc;
    } else {
        CU_TRY(cudaMemsetAsync(d_occoff, 0, 8, st));
        CU_TRY(cudaMemsetAsync(ix->choff.p, 0, 8, st));
    }
    if ((rc = rec(ix, 2, st))) return rc;
    rigk::cuts_from_offsets_kernel<<<1, 32, 0, st>>>(d_occoff, (const ull*)ix->choff.p, N, shards, cost, (ull*)ix->d_plan.p);
    CU_TRY(cudaGetLastError());
    ix->timing.launches += 1;
    CU_TRY(cudaMemcpyAsync(ix->h_counters, ix->d_counters, 8 * sizeof(ull), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(ix->h_plan, ix->d_plan.p, 3 * (shards + 1) * sizeof(ull), cudaMemcpyDeviceToHost, st));
    if ((rc = rec(ix, 3, st)) || (rc = rec(ix, 4, st))) return rc;
    CU_TRY(cudaStreamSynchronize(st));
    {   // the cut rule multiplies the batch's work by the shard count in 64 bits
        const unsigned __int128 work = (unsigned __int128)(N ? ix->h_counters[RIG_CTR_TOTAL] : 0) + (unsigned __int128)cost * N;
        if (work * shards > (unsigned __int128)~0ull) return RIG_ERR_ARG;
    }
    ix->plan_N = N; ix->plan_valid = true;
    ix->plan_cuts.assign(ix->h_plan, ix->h_plan + shards + 1);
    ix->plan_occ.assign(ix->h_plan + shards + 1, ix->h_plan + 2 * (shards + 1));
    ix->plan_ch.assign(ix->h_plan + 2 * (shards + 1), ix->h_plan + 3 * (shards + 1));
    for (uint32_t k = 0; k <= shards; ++k) cuts[k] = ix->plan_cuts[k];
    ix->timing.occ_total = N ? ix->h_counters[RIG_CTR_TOTAL] : 0;
    ix->timing.chains = N ? ix->h_counters[RIG_CTR_CHAINS] : 0;
    if (occ_total) *occ_total = ix->timing.occ_total;
    return RIG_OK;
}

// Expansion of patterns [c0, c1) of the planned batch into d_occ (the shard's occurrences from slot 0).
int expand_shard_dev(rig_index* ix, uint64_t N, uint64_t c0, uint64_t c1, const ull* d_lo, const ull* d_hi, const ull* d_occoff,
                     ull* d_occ, uint64_t cap, uint64_t* shard_total, cudaStream_t st) {
    int rc;
    if (!ix->plan_valid || N != ix->plan_N || c0 > c1 || c1 > N) return RIG_ERR_ARG;   // no planned batch (or another call since)
    uint64_t o0 = 0, o1 = 0, h0 = 0, h1 = 0;
    bool known0 = false, known1 = false;
    for (size_t k = 0; k < ix->plan_cuts.size(); ++k) {
        if (ix->plan_cuts[k] == c0 && !known0) { o0 = ix->plan_occ[k]; h0 = ix->plan_ch[k]; known0 = true; }
        if (ix->plan_cuts[k] == c1) { o1 = ix->plan_occ[k]; h1 = ix->plan_ch[k]; known1 = true; }
    }
    if (!known0 || !known1) {   // not cut points of the plan: fetch the four offsets
        ull v[4];
        CU_TRY(cudaMemcpyAsync(&v[0], d_occoff + c0, 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(&v[1], d_occoff + c1, 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(&v[2], (const ull*)ix->choff.p + c0, 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(&v[3], (const ull*)ix->choff.p + c1, 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        o0 = v[0]; o1 = v[1]; h0 = v[2]; h1 = v[3];
    }
    const uint64_t total = o1 - o0, chains = h1 - h0;
    if (shard_total) *shard_total = total;
    ix->timing.occ_total = total; ix->timing.chains = chains;
    if ((rc = rec(ix, 3, st))) return rc;
    if (total > cap || (total && !d_occ)) { if ((rc = rec(ix, 4, st))) return rc; return RIG_ERR_CAPACITY; }
    if (total) {
        ull* ctr = ix->d_counters + RIG_CTR_BLOCK;
        rigk::shard_setup_kernel<<<1, 32, 0, st>>>(ctr, d_occoff, (const ull*)ix->choff.p, c0, c1);
        CU_TRY(cudaGetLastError());
        ix->timing.launches += 1;
        if (!(ix->variant & 4) && ix->phi_bytes) {
            const uint64_t lines = (ix->phi_bytes + 127) / 128;
            rigk::l2_warm_kernel<<<(unsigned)((lines + 255) / 256), 256, 0, st>>>((const char*)ix->d.phi.rec, ix->phi_bytes);
            ix->timing.launches += 1;
        }
        const uint32_t SEG = ix->d.seed.J;
        const bool two_pass = SEG > 1 && !(ix->variant & 32) && ((reinterpret_cast<uintptr_t>(d_occ) & 127) == 0);
        uint64_t items_cap = 0;
        if (two_pass) {
            const uint64_t want_items = total / SEG + chains + 64;
            if (want_items >= (1ull << 31)) return RIG_ERR_ARG;
            if (ix->items.cap < (want_items + 32) * 16 && (rc = ix->items.ensure((want_items + 32) * 16))) return rc;
            items_cap = ix->items.cap / 16 - 32;
        }
        ix->last_items_cap = items_cap; ix->last_two_pass = two_pass;
        if ((rc = launch_expansion(ix, c1 - c0, d_lo + c0, d_hi + c0, d_occoff + c0, d_occ, cap, two_pass, items_cap, st, false, c0, ctr)))
            return rc;
    }
    if ((rc = rec(ix, 4, st))) return rc;
    return RIG_OK;
}

// The host-buffer locate calls: upload the patterns, locate, optional -o / -c post-processing on the device, download.
// occ32: narrow the positions to 32 bits on the device before the download (rig_locate_batch32).
int locate_host(rig_index* ix, const uint8_t* patterns, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi,
                uint64_t* occ_offsets, void* occ, bool occ32, uint64_t occ_capacity, uint64_t* occ_total, uint32_t flags,
                rig_check_report* report) {
    if (!ix || !occ_offsets || (N && (!lo || !hi)) || (N && m && !patterns)) return RIG_ERR_ARG;
    if ((flags & RIG_LOCATE_CHECK) && (!report || !ix->has_text || occ32)) return RIG_ERR_ARG;
    if (occ32 && ix->d.n > 0xFFFFFFFFull) return RIG_ERR_ARG;  // positions would not fit
    const bool devonly = (flags & RIG_LOCATE_DEVICE_ONLY) != 0;
    if (devonly && occ32) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    int rc;
    if ((rc = ix->patt.ensure(N * m + 16)) || (rc = ix->lo.ensure((N + 1) * 8)) || (rc = ix->hi.ensure((N + 1) * 8)) ||
        (rc = ix->occoff.ensure((N + 2) * 8)))
        return rc;
    const bool want = devonly || (occ && occ_capacity);
    // 32-bit positions are written by the expansion kernels themselves (half the store sectors, no narrowing pass) when the
    // index has the default table form and no sort is asked for (the sort works on 64-bit keys); otherwise narrowed after
    const bool native32 = occ32 && ix->d.w32 && ix->d.phi.D == 4 && !(flags & RIG_LOCATE_SORT) && !(ix->variant & 262144);
    if (!devonly && want && !native32 && (rc = ix->occ.ensure(occ_capacity * 8))) return rc;
    if (occ32 && want && (rc = ix->occ32.ensure(occ_capacity * 4 + 128))) return rc;
    begin_call(ix);
    if ((rc = rec(ix, 0, st))) return rc;
    if (N * m) CU_TRY(cudaMemcpyAsync(ix->patt.p, patterns, N * m, cudaMemcpyHostToDevice, st));
    uint64_t total = 0;
    int lrc = locate_dev(ix, (const uint8_t*)ix->patt.p, N, m, (ull*)ix->lo.p, (ull*)ix->hi.p, (ull*)ix->occoff.p,
                         want ? (native32 ? ix->occ32.p : ix->occ.p) : nullptr, want ? occ_capacity : 0, &total, st,
                         devonly ? &ix->occ : nullptr, native32);
    if (lrc != RIG_OK && lrc != RIG_ERR_CAPACITY) return lrc;
    if (occ_total) *occ_total = total;
    ix->kept_total = (lrc == RIG_OK && devonly) ? total : 0;
    if (lrc == RIG_OK && total && (flags & (RIG_LOCATE_SORT | RIG_LOCATE_CHECK))) {  // the reference sorts for -o and for -c (:147,:159)
        if ((rc = sort_dev(ix, N, (const ull*)ix->occoff.p, (ull*)ix->occ.p, total, st))) return rc;
    }
    if (lrc == RIG_OK && (flags & RIG_LOCATE_CHECK)) {
        if ((rc = check_dev(ix, (const uint8_t*)ix->patt.p, N, m, (const ull*)ix->lo.p, (const ull*)ix->hi.p,
                            (const ull*)ix->occoff.p, (const ull*)ix->occ.p, total, 1, report, st)))
            return rc;
    }
    if (lrc == RIG_OK && total && occ32 && !native32) {
        const uint64_t nb = ((total + 3) / 4 + 255) / 256;
        if (nb > 0x7fffffffull) return RIG_ERR_ARG;
        rigk::narrow_kernel<<<(unsigned)nb, 256, 0, st>>>((const ull*)ix->occ.p, (uint32_t*)ix->occ32.p, total);
        CU_TRY(cudaGetLastError());
        ix->timing.launches += 1;
    }
    if (N) {
        CU_TRY(cudaMemcpyAsync(lo, ix->lo.p, N * 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(hi, ix->hi.p, N * 8, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaMemcpyAsync(occ_offsets, ix->occoff.p, (N + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (lrc == RIG_OK && total && !devonly)
        CU_TRY(cudaMemcpyAsync(occ, occ32 ? ix->occ32.p : ix->occ.p, total * (occ32 ? 4 : 8), cudaMemcpyDeviceToHost, st));
    if ((rc = rec(ix, 5, st))) return rc;
    CU_TRY(cudaStreamSynchronize(st));
    return lrc;
}

}  // namespace

extern "C" {

int rig_count_batch_dev(rig_index* ix, const uint8_t* d_patterns, uint64_t N, uint64_t m, uint64_t* d_lo,
                        uint64_t* d_hi, void* stream) {
    if (!ix || (N && (!d_patterns && m)) || (N && (!d_lo || !d_hi))) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    begin_call(ix);
    return count_dev(ix, d_patterns, N, m, (ull*)d_lo, (ull*)d_hi, st);
}

int rig_locate_batch_dev(rig_index* ix, const uint8_t* d_patterns, uint64_t N, uint64_t m, uint64_t* d_lo,
                         uint64_t* d_hi, uint64_t* d_occ_offsets, uint64_t* d_occ, uint64_t occ_capacity,
                         uint64_t* occ_total, void* stream) {
    if (!ix || !d_occ_offsets || (N && (!d_patterns && m)) || (N && (!d_lo || !d_hi))) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    begin_call(ix);
    return locate_dev(ix, d_patterns, N, m, (ull*)d_lo, (ull*)d_hi, (ull*)d_occ_offsets, (ull*)d_occ, occ_capacity,
                      occ_total, st);
}

int rig_plan_batch_dev(rig_index* ix, const uint8_t* d_patterns, uint64_t N, uint64_t m, uint64_t* d_lo, uint64_t* d_hi,
                       uint64_t* d_occ_offsets, uint32_t shards, uint64_t per_pattern_cost, uint64_t* cuts, uint64_t* occ_total,
                       void* stream) {
    if (!ix || !cuts || !d_occ_offsets || shards < 1 || shards > 1024 || (N && (!d_patterns && m)) || (N && (!d_lo || !d_hi)))
        return RIG_ERR_ARG;
    if (N >= (1ull << 32) || per_pattern_cost > (1ull << 20)) return RIG_ERR_ARG;   // total * shards must stay below 2^64
    CU_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    begin_call(ix);
    return plan_dev(ix, d_patterns, N, m, (ull*)d_lo, (ull*)d_hi, (ull*)d_occ_offsets, shards, per_pattern_cost, cuts, occ_total, st);
}

int rig_expand_shard_dev(rig_index* ix, uint64_t N, uint64_t c0, uint64_t c1, const uint64_t* d_lo, const uint64_t* d_hi,
                         const uint64_t* d_occ_offsets, uint64_t* d_occ, uint64_t occ_capacity, uint64_t* shard_total, void* stream) {
    if (!ix || !d_occ_offsets || !d_lo || !d_hi) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    begin_call(ix, true);
    return expand_shard_dev(ix, N, c0, c1, (const ull*)d_lo, (const ull*)d_hi, (const ull*)d_occ_offsets, (ull*)d_occ, occ_capacity,
                            shard_total, st);
}

int rig_count_batch(rig_index* ix, const uint8_t* patterns, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi) {
    if (!ix || (N && (!lo || !hi)) || (N && m && !patterns)) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    int rc;
    if ((rc = ix->patt.ensure(N * m + 16)) || (rc = ix->lo.ensure((N + 1) * 8)) || (rc = ix->hi.ensure((N + 1) * 8))) return rc;
    begin_call(ix);
    if ((rc = rec(ix, 0, st))) return rc;
    if (N * m) CU_TRY(cudaMemcpyAsync(ix->patt.p, patterns, N * m, cudaMemcpyHostToDevice, st));
    if ((rc = count_dev(ix, (const uint8_t*)ix->patt.p, N, m, (ull*)ix->lo.p, (ull*)ix->hi.p, st))) return rc;
    if ((rc = rec(ix, 3, st)) || (rc = rec(ix, 4, st))) return rc;
    if (N) {
        CU_TRY(cudaMemcpyAsync(lo, ix->lo.p, N * 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(hi, ix->hi.p, N * 8, cudaMemcpyDeviceToHost, st));
    }
    if ((rc = rec(ix, 5, st))) return rc;
    CU_TRY(cudaStreamSynchronize(st));
    return RIG_OK;
}

int rig_locate_batch(rig_index* ix, const uint8_t* patterns, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi,
                     uint64_t* occ_offsets, uint64_t* occ, uint64_t occ_capacity, uint64_t* occ_total) {
    return locate_host(ix, patterns, N, m, lo, hi, occ_offsets, occ, false, occ ? occ_capacity : 0, occ_total, 0, nullptr);
}

int rig_digest_dev(rig_index* ix, const uint64_t* d_values, uint64_t count, uint64_t out[2], void* stream) {
    if (!ix || !out || (count && !d_values)) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    ull* acc = ix->d_counters + 4;
    CU_TRY(cudaMemsetAsync(acc, 0, 2 * sizeof(ull), st));
    if (count) {
        uint64_t blocks = (count + 255) / 256;
        uint64_t maxb = (uint64_t)ix->sm_count * 8;
        rigk::digest_kernel<<<(unsigned)(blocks < maxb ? blocks : maxb), 256, 0, st>>>((const ull*)d_values, count, acc);
        CU_TRY(cudaGetLastError());
    }
    ull h[2];
    CU_TRY(cudaMemcpyAsync(h, acc, sizeof(h), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    out[0] = h[0]; out[1] = h[1];
    return RIG_OK;
}

// ---- ri-locate -o / -c on the device -------------------------------------------------
}  // extern "C"
namespace {

int sort_dev(rig_index* ix, uint64_t N, const ull* d_off, ull* d_occ, uint64_t total, cudaStream_t st) {
    if (N == 0 || total < 2) return RIG_OK;
    if (N >= 0x7fffffffull) return RIG_ERR_ARG;
    const bool k32 = ix->d.w32 != 0;  // n < 2^32-1: every position fits 32 bits, half the shared memory per key
    const uint32_t cap1 = 4096, cap2 = k32 ? 32768u : 16384u;
    const size_t ksz = k32 ? 4 : 8;
    const uint64_t nbig1 = std::min<uint64_t>(N, total / cap1 + 1), nbig2 = std::min<uint64_t>(N, total / cap2 + 1);
    int rc;
    if ((rc = ix->big1.ensure(nbig1 * 4)) || (rc = ix->big2.ensure(nbig2 * 4))) return rc;
    CU_TRY(cudaMemsetAsync(ix->d_post, 0, 2 * sizeof(ull), st));
    uint32_t* big1 = (uint32_t*)ix->big1.p; uint32_t* big2 = (uint32_t*)ix->big2.p;
    if (!ix->sort_attr_done) {
        CU_TRY(cudaFuncSetAttribute(rigk::segsort_smem_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
        CU_TRY(cudaFuncSetAttribute(rigk::segsort_smem_kernel<ull>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        ix->sort_attr_done = true;
    }
    // tier 1: one CTA per pattern, segments up to cap1 keys; tier 2: the longer ones, up to cap2; tier 3: the rest
    if (k32) {
        rigk::segsort_smem_kernel<uint32_t><<<(unsigned)N, 256, cap1 * ksz, st>>>(d_off, d_occ, N, cap1, nullptr, nullptr, big1, ix->d_post + 0);
        rigk::segsort_smem_kernel<uint32_t><<<(unsigned)nbig1, 1024, cap2 * ksz, st>>>(d_off, d_occ, N, cap2, big1, ix->d_post + 0, big2, ix->d_post + 1);
    } else {
        rigk::segsort_smem_kernel<ull><<<(unsigned)N, 256, cap1 * ksz, st>>>(d_off, d_occ, N, cap1, nullptr, nullptr, big1, ix->d_post + 0);
        rigk::segsort_smem_kernel<ull><<<(unsigned)nbig1, 1024, cap2 * ksz, st>>>(d_off, d_occ, N, cap2, big1, ix->d_post + 0, big2, ix->d_post + 1);
    }
    rigk::segsort_global_kernel<<<(unsigned)nbig2, 1024, 0, st>>>(d_off, d_occ, big2, ix->d_post + 1);
    CU_TRY(cudaGetLastError());
    return RIG_OK;
}

int check_dev(rig_index* ix, const uint8_t* d_patt, uint64_t N, uint64_t m, const ull* d_lo, const ull* d_hi,
              const ull* d_off, const ull* d_occ, uint64_t total, int sorted, rig_check_report* report, cudaStream_t st) {
    if (!ix->has_text) return RIG_ERR_ARG;
    if (N >= 0x7fffffffull) return RIG_ERR_ARG;
    uint64_t tsize = 16;
    while (tsize < 2 * N) tsize <<= 1;
    int rc;
    if ((rc = ix->ctable.ensure(tsize * 4)) || (rc = ix->crep.ensure((N + 1) * 4)) || (rc = ix->cfound.ensure((N + 1) * 8))) return rc;
    rigk::CheckReport* rp = reinterpret_cast<rigk::CheckReport*>(ix->d_post + 2);
    CU_TRY(cudaMemsetAsync(ix->ctable.p, 0, tsize * 4, st));
    CU_TRY(cudaMemsetAsync(ix->cfound.p, 0, (N + 1) * 8, st));
    ix->h_post[2] = N; ix->h_post[3] = 0; ix->h_post[4] = 0; ix->h_post[5] = 0; ix->h_post[6] = ~0ull; ix->h_post[7] = ~0ull;
    CU_TRY(cudaMemcpyAsync(ix->d_post + 2, ix->h_post + 2, 6 * sizeof(ull), cudaMemcpyHostToDevice, st));
    const uint8_t* text = (const uint8_t*)ix->text.p;
    const uint64_t len = ix->text_len;
    if (N) {
        const unsigned nb = (unsigned)((N + 255) / 256);
        if (m) {
            rigk::check_build_kernel<<<nb, 256, 0, st>>>(d_patt, N, m, (uint32_t*)ix->ctable.p, tsize - 1, (uint32_t*)ix->crep.p);
            if (len >= m) {
                const uint64_t pos = len - m + 1;
                if ((pos + 255) / 256 > 0x7fffffffull) return RIG_ERR_ARG;
                rigk::check_scan_kernel<<<(unsigned)((pos + 255) / 256), 256, 0, st>>>(text, len, d_patt, m, (const uint32_t*)ix->ctable.p,
                                                                                       tsize - 1, (ull*)ix->cfound.p);
            }
        }
        rigk::check_counts_kernel<<<nb, 256, 0, st>>>(d_lo, d_hi, (const uint32_t*)ix->crep.p, (const ull*)ix->cfound.p, N, m, len, rp);
        if (total) {
            if ((total + 255) / 256 > 0x7fffffffull) return RIG_ERR_ARG;
            rigk::check_positions_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(text, len, d_patt, N, m, d_off, d_occ, total, sorted, rp);
        }
        CU_TRY(cudaGetLastError());
    }
    CU_TRY(cudaMemcpyAsync(ix->h_post + 2, ix->d_post + 2, 6 * sizeof(ull), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (report) std::memcpy(report, ix->h_post + 2, sizeof(rig_check_report));
    return RIG_OK;
}

}  // namespace
extern "C" {

int rig_text_attach(rig_index* ix, const uint8_t* text, uint64_t len) {
    if (!ix || (len && !text)) return RIG_ERR_ARG;
    if (len + 1 != ix->d.n) return RIG_ERR_ARG;  // must be the indexed text: n = |T| + 1 (r_index.hpp:88)
    CU_TRY(cudaSetDevice(ix->device));
    int rc;
    if ((rc = ix->text.ensure(len + 16))) return rc;
    if (len) CU_TRY(cudaMemcpy(ix->text.p, text, len, cudaMemcpyHostToDevice));
    ix->text_len = len; ix->has_text = true;
    return RIG_OK;
}

int rig_sort_occurrences_dev(rig_index* ix, uint64_t N, const uint64_t* d_occ_offsets, uint64_t* d_occ, uint64_t total,
                             void* stream) {
    if (!ix || (N && !d_occ_offsets) || (total && !d_occ)) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    return sort_dev(ix, N, (const ull*)d_occ_offsets, (ull*)d_occ, total, stream ? (cudaStream_t)stream : ix->stream);
}

int rig_check_dev(rig_index* ix, const uint8_t* d_patterns, uint64_t N, uint64_t m, const uint64_t* d_lo,
                  const uint64_t* d_hi, const uint64_t* d_occ_offsets, const uint64_t* d_occ, uint64_t total, int sorted,
                  rig_check_report* report, void* stream) {
    if (!ix || !report || (N && (!d_lo || !d_hi || !d_occ_offsets || (m && !d_patterns))) || (total && !d_occ)) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    return check_dev(ix, d_patterns, N, m, (const ull*)d_lo, (const ull*)d_hi, (const ull*)d_occ_offsets, (const ull*)d_occ,
                     total, sorted, report, stream ? (cudaStream_t)stream : ix->stream);
}

int rig_locate_batch_ex(rig_index* ix, const uint8_t* patterns, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi,
                        uint64_t* occ_offsets, uint64_t* occ, uint64_t occ_capacity, uint64_t* occ_total, uint32_t flags,
                        rig_check_report* report) {
    return locate_host(ix, patterns, N, m, lo, hi, occ_offsets, occ, false, occ ? occ_capacity : 0, occ_total, flags, report);
}

// ---- 32-bit positions for texts below 4 GiB: half the bytes over PCIe -------------------------------------
int rig_locate_batch32(rig_index* ix, const uint8_t* patterns, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi,
                       uint64_t* occ_offsets, uint32_t* occ, uint64_t occ_capacity, uint64_t* occ_total, uint32_t flags) {
    if (flags & (RIG_LOCATE_CHECK | RIG_LOCATE_DEVICE_ONLY)) return RIG_ERR_ARG;
    return locate_host(ix, patterns, N, m, lo, hi, occ_offsets, occ, true, occ ? occ_capacity : 0, occ_total, flags, nullptr);
}

// occurrences [first, first + count) of the most recent RIG_LOCATE_DEVICE_ONLY call -> host
int rig_fetch_occurrences(rig_index* ix, uint64_t first, uint64_t count, uint64_t* out) {
    if (!ix || (count && !out) || first > ix->kept_total || count > ix->kept_total - first) return RIG_ERR_ARG;
    if (!count) return RIG_OK;
    CU_TRY(cudaSetDevice(ix->device));
    CU_TRY(cudaMemcpyAsync(out, (const ull*)ix->occ.p + first, count * 8, cudaMemcpyDeviceToHost, ix->stream));
    CU_TRY(cudaStreamSynchronize(ix->stream));
    return RIG_OK;
}

// page-locked host memory for result buffers (a download into pageable memory runs at a fraction of the link rate)
void* rig_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void rig_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- single-position navigation as batches --------------------------------------------
int rig_navigate_batch_dev(rig_index* ix, int op, const uint64_t* d_positions, uint64_t N, uint64_t* d_out, void* stream) {
    if (!ix || op < RIG_NAV_BWT || op > RIG_NAV_F_AT || (N && (!d_positions || !d_out))) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    if (!N) return RIG_OK;
    const uint64_t nb = (N + 255) / 256;
    if (nb > 0x7fffffffull) return RIG_ERR_ARG;
    if (ix->d.w32) rigk::navigate_kernel<uint32_t><<<(unsigned)nb, 256, 0, st>>>(ix->d, op, (const ull*)d_positions, N, (ull*)d_out);
    else rigk::navigate_kernel<ull><<<(unsigned)nb, 256, 0, st>>>(ix->d, op, (const ull*)d_positions, N, (ull*)d_out);
    CU_TRY(cudaGetLastError());
    return RIG_OK;
}

int rig_navigate_batch(rig_index* ix, int op, const uint64_t* positions, uint64_t N, uint64_t* out) {
    if (!ix || op < RIG_NAV_BWT || op > RIG_NAV_F_AT || (N && (!positions || !out))) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    int rc;
    if ((rc = ix->lo.ensure((N + 1) * 8)) || (rc = ix->hi.ensure((N + 1) * 8))) return rc;
    cudaStream_t st = ix->stream;
    if (N) CU_TRY(cudaMemcpyAsync(ix->lo.p, positions, N * 8, cudaMemcpyHostToDevice, st));
    if ((rc = rig_navigate_batch_dev(ix, op, (const uint64_t*)ix->lo.p, N, (uint64_t*)ix->hi.p, st))) return rc;
    if (N) CU_TRY(cudaMemcpyAsync(out, ix->hi.p, N * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return RIG_OK;
}

int rig_get_bwt(rig_index* ix, uint64_t from, uint64_t len, uint8_t* out) {
    if (!ix || (len && !out) || from > ix->d.n || len > ix->d.n - from) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    if (!len) return RIG_OK;
    int rc;
    if ((rc = ix->patt.ensure(len + 16))) return rc;
    cudaStream_t st = ix->stream;
    const uint64_t nb = ((len + 15) / 16 + 255) / 256;
    if (nb > 0x7fffffffull) return RIG_ERR_ARG;
    if (ix->d.w32) rigk::bwt_range_kernel<uint32_t><<<(unsigned)nb, 256, 0, st>>>(ix->d, from, len, (uint8_t*)ix->patt.p);
    else rigk::bwt_range_kernel<ull><<<(unsigned)nb, 256, 0, st>>>(ix->d, from, len, (uint8_t*)ix->patt.p);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out, ix->patt.p, len, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return RIG_OK;
}

// ---- range utilities of the RLBWT as batches: break_range, closest_run_break --------------------
namespace {
int range_nav_upload(rig_index* ix, const uint64_t* lo, const uint64_t* hi, const uint8_t* c, uint64_t N, cudaStream_t st) {
    int rc;
    if ((rc = ix->lo.ensure((N + 1) * 8)) || (rc = ix->hi.ensure((N + 1) * 8)) || (rc = ix->patt.ensure(N + 16)) ||
        (rc = ix->occoff.ensure((N + 2) * 8)) || (rc = ix->nocc.ensure((N + 1) * 8)))
        return rc;
    CU_TRY(cudaMemcpyAsync(ix->lo.p, lo, N * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(ix->hi.p, hi, N * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(ix->patt.p, c, N, cudaMemcpyHostToDevice, st));
    return RIG_OK;
}
int range_nav_launch(rig_index* ix, int op, uint64_t N, const ull* off, ull* out_a, ull* out_b, cudaStream_t st) {
    const uint64_t nb = (N + 255) / 256;
    if (nb > 0x7fffffffull) return RIG_ERR_ARG;
    if (ix->d.w32) rigk::range_nav_kernel<uint32_t><<<(unsigned)nb, 256, 0, st>>>(ix->d, op, (const ull*)ix->lo.p, (const ull*)ix->hi.p,
                                                                                  (const uint8_t*)ix->patt.p, N, off, out_a, out_b);
    else rigk::range_nav_kernel<ull><<<(unsigned)nb, 256, 0, st>>>(ix->d, op, (const ull*)ix->lo.p, (const ull*)ix->hi.p,
                                                                   (const uint8_t*)ix->patt.p, N, off, out_a, out_b);
    CU_TRY(cudaGetLastError());
    return RIG_OK;
}
}  // namespace

int rig_closest_run_break_batch(rig_index* ix, const uint64_t* lo, const uint64_t* hi, const uint8_t* c, uint64_t N, uint64_t* out) {
    if (!ix || (N && (!lo || !hi || !c || !out))) return RIG_ERR_ARG;
    if (!N) return RIG_OK;
    CU_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    int rc;
    if ((rc = range_nav_upload(ix, lo, hi, c, N, st))) return rc;
    if ((rc = range_nav_launch(ix, rigk::RANGE_CLOSEST, N, nullptr, (ull*)ix->nocc.p, nullptr, st))) return rc;
    CU_TRY(cudaMemcpyAsync(out, ix->nocc.p, N * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return RIG_OK;
}

int rig_break_range_batch(rig_index* ix, const uint64_t* lo, const uint64_t* hi, const uint8_t* c, uint64_t N,
                          uint64_t* out_offsets, uint64_t* out_first, uint64_t* out_last, uint64_t capacity, uint64_t* total) {
    if (!ix || !out_offsets || (N && (!lo || !hi || !c))) return RIG_ERR_ARG;
    out_offsets[0] = 0;
    if (total) *total = 0;
    if (!N) return RIG_OK;
    CU_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    int rc;
    if ((rc = range_nav_upload(ix, lo, hi, c, N, st))) return rc;
    if ((rc = range_nav_launch(ix, rigk::RANGE_BREAK_COUNT, N, nullptr, (ull*)ix->nocc.p, nullptr, st))) return rc;
    CU_TRY(cudaMemcpyAsync(out_offsets + 1, ix->nocc.p, N * 8, cudaMemcpyDeviceToHost, st));  // counts, turned into offsets below
    CU_TRY(cudaStreamSynchronize(st));
    for (uint64_t k = 1; k <= N; ++k) out_offsets[k] += out_offsets[k - 1];
    const uint64_t tot = out_offsets[N];
    if (total) *total = tot;
    if (tot > capacity || (tot && (!out_first || !out_last))) return RIG_ERR_CAPACITY;
    if (!tot) return RIG_OK;
    if ((rc = ix->big1.ensure(tot * 8)) || (rc = ix->big2.ensure(tot * 8))) return rc;
    CU_TRY(cudaMemcpyAsync(ix->occoff.p, out_offsets, (N + 1) * 8, cudaMemcpyHostToDevice, st));
    if ((rc = range_nav_launch(ix, rigk::RANGE_BREAK_FILL, N, (const ull*)ix->occoff.p, (ull*)ix->big1.p, (ull*)ix->big2.p, st))) return rc;
    CU_TRY(cudaMemcpyAsync(out_first, ix->big1.p, tot * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(out_last, ix->big2.p, tot * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return RIG_OK;
}

// ---- multi-GPU fan-out helpers on device buffers ----------------------------------------------------
int rig_counts_dev(rig_index* ix, const uint64_t* d_lo, const uint64_t* d_hi, uint64_t N, uint64_t* d_nocc, void* stream) {
    if (!ix || (N && (!d_lo || !d_hi || !d_nocc))) return RIG_ERR_ARG;
    if (!N) return RIG_OK;
    CU_TRY(cudaSetDevice(ix->device));
    const uint64_t nb = (N + 255) / 256;
    if (nb > 0x7fffffffull) return RIG_ERR_ARG;
    rigk::counts_kernel<<<(unsigned)nb, 256, 0, stream ? (cudaStream_t)stream : ix->stream>>>((const ull*)d_lo, (const ull*)d_hi, N, (ull*)d_nocc);
    CU_TRY(cudaGetLastError());
    return RIG_OK;
}

int rig_balanced_cuts_dev(rig_index* ix, const uint64_t* d_nocc, uint64_t N, uint32_t shards, uint64_t per_pattern_cost,
                          uint64_t* cuts, void* stream) {
    if (!ix || !cuts || shards < 1 || shards > 1024 || (N && !d_nocc) || N >= (1ull << 40) / shards) return RIG_ERR_ARG;
    CU_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    cuts[0] = 0;
    for (uint32_t k = 1; k <= shards; ++k) cuts[k] = N;
    if (shards == 1 || N == 0) return RIG_OK;
    int rc;
    if ((rc = ix->crep.ensure(1024 * 8))) return rc;
    rigk::balanced_cuts_kernel<<<1, 1024, 0, st>>>((const ull*)d_nocc, N, shards, per_pattern_cost, (ull*)ix->crep.p);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(ix->h_post, ix->crep.p, std::min<uint32_t>(shards - 1, 8) * 8, cudaMemcpyDeviceToHost, st));
    std::vector<ull> big;
    if (shards - 1 > 8) { big.resize(shards - 1); CU_TRY(cudaMemcpyAsync(big.data(), ix->crep.p, (shards - 1) * 8, cudaMemcpyDeviceToHost, st)); }
    CU_TRY(cudaStreamSynchronize(st));
    for (uint32_t k = 1; k < shards; ++k) {
        ull c = shards - 1 > 8 ? big[k - 1] : ix->h_post[k - 1];
        if (c == ~0ull) c = N;
        cuts[k] = std::min<uint64_t>(std::max<uint64_t>(c, cuts[k - 1]), N);
    }
    return RIG_OK;
}

int rig_last_timing(const rig_index* cix, rig_timing* t) {
    if (!cix || !t) return RIG_ERR_ARG;
    rig_index* ix = const_cast<rig_index*>(cix);
    int rc = finish_timing(ix);
    if (rc) return rc;
    *t = ix->timing;
    return RIG_OK;
}

}  // extern "C"
