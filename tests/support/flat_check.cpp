// flat_check.cpp — TEST DOUBLE, not shipped: walks the flattened arrays of
// r-index_b200/csrc/flat_layout.hpp with scalar CPU code that mirrors, step by step, what the
// CUDA kernels in kernels.cuh do (bdir -> bstart search -> block scan; toehold; chain cutting at
// run boundaries; phi_dir -> phi_ent). It lets the CPU-only test tier validate the flatten step
// and the index arithmetic against the oracle before any GPU time is spent. It is compiled into
// tests/support/libflat_check.so by tests/conftest.py and is never loaded by the product.
#include "../../r-index_b200/csrc/flat_layout.hpp"
#include <cstring>

using rigf::FlatHost;
typedef uint64_t u64;
typedef uint32_t u32;

namespace {

struct Q { u64 cnt, run, prev_c_run; bool head_is_c; };

// mirrors rigk::block_query, reading the interleaved block records exactly as the device does
u64 rd(const uint8_t* p, u32 W) { u64 x = 0; memcpy(&x, p, W); return x; }  // W = 4, 5 (40-bit packed) or 8 bytes, little endian
Q block_query(const FlatHost& f, u64 x, uint8_t c, u32 sidc) {
    const u32 K = f.K, W = f.rec_w;
    u64 q = x >> f.lf_shift;
    u64 b0 = f.bdir[q], b1 = f.bdir[q + 1];
    while (b1 > b0) {  // same G-ary narrowing as the kernel, G = K probes per round
        u64 span = b1 - b0, step = (span + K - 1) / K, k = 0;
        for (u32 g = 0; g < K; ++g) {
            u64 probe = b0 + (u64)(g + 1) * step;
            if (probe > b1) probe = b1;
            if (f.bstart[probe] <= x) ++k;
        }
        if (k == 0) b1 = std::min(b1, b0 + step - 1);
        else { u64 nb0 = std::min(b1, b0 + k * step); b1 = std::min(b1, nb0 + step - 1); b0 = nb0; }
    }
    const uint8_t* R = &f.blk[b0 * f.blk_stride];
    u64 base = b0 * K;
    u64 st[16];
    for (u32 g = 0; g < K; ++g) st[g] = rd(R + g * W, W);
    int t = -1;
    for (u32 g = 0; g < K; ++g) if (st[g] <= x) ++t;
    Q r;
    u64 sum = 0; u32 mc = 0;
    for (u32 g = 0; g < K; ++g) {
        bool isc = R[f.off_head + g] == c;
        if (isc) mc |= 1u << g;
        if (isc) sum += ((int)g < t) ? (st[g + 1] - st[g]) : (((int)g == t) ? (x - st[g] + 1) : 0);
    }
    r.cnt = rd(R + f.off_cum + sidc * W, W) + sum;
    r.head_is_c = (mc >> t) & 1u;
    u32 below = mc & ((1u << t) - 1u);
    r.prev_c_run = below ? base + (31 - __builtin_clz(below)) : f.last[b0 * f.S + sidc];
    r.run = base + t;
    return r;
}

// mirrors the packed branch of rigk::load_entry (32-byte entries of a 64-bit index: 4 x 40-bit deltas | 40-bit
// s1/start | 32-bit nxt | 24-bit cnt)
void unpack32(const uint8_t* p, u64* w) {
    u64 q0, q1, q2, q3;
    memcpy(&q0, p, 8); memcpy(&q1, p + 8, 8); memcpy(&q2, p + 16, 8); memcpy(&q3, p + 24, 8);
    const u64 m40 = 0xFFFFFFFFFFull;
    w[0] = q0 & m40;
    w[1] = ((q0 >> 40) | (q1 << 24)) & m40;
    w[2] = (q1 >> 16) & m40;
    w[3] = ((q1 >> 56) | (q2 << 8)) & m40;
    w[4] = ((q2 >> 32) | (q3 << 32)) & m40;
    if (w[4] == m40) w[4] = ~(u64)0;
    w[5] = (q3 >> 8) & 0xFFFFFFFFull;
    w[6] = q3 >> 40;
    w[7] = 0;
}

// mirrors the per-lane state machine of rigk::phi_expand_kernel for ONE lookup: returns
// e[t] = Phi^(t+1)(i), reading only rec[] / pent[] entries (with the 32-bit truncation when w32).
// *loads counts the entry loads the lane needed.
void phi_lookup(const FlatHost& f, u64 i, u64* e, u64* loads = nullptr) {
    const rigf::PhiTable& T = f.phi;
    const u32 D = T.D, RW = T.RW;
    const u64 mask = f.w32 ? 0xFFFFFFFFull : ~(u64)0;
    bool searching = false;
    u64 slo = 0, shi = 0, nload = 0;
    for (;;) {
        const u64 probe = (slo < shi) ? ((slo + shi + 1) >> 1) : slo;
        const u64* w = searching ? &T.pent[probe * RW] : &T.rec[(i >> T.shift) * RW];
        u64 unpacked[8];
        if (f.phi_packed) {  // what the device reads
            unpack32(searching ? &f.phi_pent_p[probe * 32] : &f.phi_rec_p[(i >> T.shift) * 32], unpacked);
            w = unpacked;
        }
        ++nload;
        bool emit;
        if (!searching) {
            emit = i < (w[D] & mask);
            if (!emit) { slo = rigf::PhiTable::nxt_of(w, D) & mask; shi = slo + (rigf::PhiTable::cnt_of(w, D) & mask) - 1; searching = true; }
        } else if (slo == shi) emit = true;
        else if ((w[D] & mask) <= i) { slo = probe; emit = (slo == shi); }
        else { shi = probe - 1; emit = false; }
        if (emit) {
            for (u32 j = 0; j < D; ++j) { u64 v = i + (w[j] & mask); if (v >= f.n) v -= f.n; e[j] = v; }
            if (loads) *loads += nload;
            return;
        }
    }
}

// mirrors rigk::seed_hop: Phi^SEG(i) reading only seed.rec[] / seed.pent[] (32-bit truncation when w32)
u64 seed_hop(const FlatHost& f, u64 i) {
    const rigf::JumpTable& T = f.seed;
    const u64 mask = f.w32 ? 0xFFFFFFFFull : ~(u64)0;
    const u64* w = &T.rec[(i >> T.shift) * rigf::SEED_RW];
    u64 d = w[0] & mask;
    for (u32 k = 0; k < rigf::SEED_INLINE; ++k)
        if (i >= (w[1 + 2 * k] & mask)) d = w[2 + 2 * k] & mask;
    if ((w[14] & mask) > rigf::SEED_INLINE && i >= (w[11] & mask)) {
        u64 lo = (w[13] & mask) + 5, hi = (w[13] & mask) + (w[14] & mask) - 1;
        while (lo < hi) {
            const u64 mid = (lo + hi + 1) >> 1;
            if ((T.pent[2 * mid + 1] & mask) <= i) lo = mid; else hi = mid - 1;
        }
        d = T.pent[2 * lo] & mask;
    }
    u64 v = i + d;
    if (v >= f.n) v -= f.n;
    return v;
}

// mirrors rigk::walk_chain: `remaining` occurrences after v, written from o on; returns the last value
u64 walk_chain(const FlatHost& f, u64 v, u64* occ, u64 slot, u64 remaining, bool& misaligned) {
    const u32 D = f.phi.D;
    u64 e[8];
    u64* o = occ + slot;
    const u32 mis = (u32)(slot % D);
    if (D > 1 && mis != 0 && remaining > 0) {
        const u32 cnt = (u32)std::min<u64>(D - mis, remaining);
        phi_lookup(f, v, e);
        for (u32 t = 0; t < cnt; ++t) { o[t] = e[t]; v = e[t]; }
        o += cnt; remaining -= cnt;
    }
    while (remaining >= D) {
        phi_lookup(f, v, e);
        if (((u64)(o - occ)) % D != 0) misaligned = true;  // vector stores must be aligned
        for (u32 t = 0; t < D; ++t) o[t] = e[t];
        v = e[D - 1]; o += D; remaining -= D;
    }
    if (D > 1 && remaining > 0) {
        phi_lookup(f, v, e);
        for (u32 t = 0; t < remaining; ++t) { o[t] = e[t]; v = e[t]; }
    }
    return v;
}

// mirrors rigk::phi_window_kernel for one item (first slot << 8 | occurrences after the seed): groups
// [v, Phi(v), .., Phi^(D-1)(v)]; whole 16-slot lines are written as lines, the rest group by group
int item_fill(const FlatHost& f, u64* occ, u64 item, u64 seed) {
    const u32 D = f.phi.D;
    u64 left = (item & 255) + 1;
    u64 slot = item >> 8;
    if (slot % 16 != 0) return 1;  // items start on a 128-byte line of the output
    u64 v = seed, e[8];
    occ[slot] = v;
    while (left > 1) {
        phi_lookup(f, v, e);
        const u64 cnt = std::min<u64>(left, D);
        occ[slot] = v;
        for (u32 t = 1; t < cnt; ++t) occ[slot + t] = e[t - 1];
        v = e[D - 1]; slot += cnt; left -= cnt;
    }
    if (left == 1) occ[slot] = v;
    return 0;
}

// mirrors rigk::search_kernel (one pattern)
void search(const FlatHost& f, const uint8_t* P, u64 m, bool locate, u64& lo, u64& hi, u64& k) {
    lo = 0; hi = f.n - 1; k = f.toe0;
    for (u64 i = 0; i < m; ++i) {
        uint8_t c = P[m - 1 - i];
        u64 Fc = f.F[c], Fc1 = f.F[c + 1];
        if (!(Fc < Fc1)) { lo = 1; hi = 0; return; }
        u32 sidc = f.sid[c];
        u64 A = lo > 0 ? block_query(f, lo - 1, c, sidc).cnt : 0;
        Q qb = block_query(f, hi, c, sidc);
        if (qb.cnt == A) { lo = 1; hi = 0; return; }
        if (locate) { if (qb.head_is_c) k -= 1; else k = f.samples_last[qb.prev_c_run]; }
        lo = Fc + A; hi = Fc + qb.cnt - 1;
    }
}

}  // namespace

extern "C" {

void* fc_create(const rig_logical_view* v, uint32_t K, uint32_t lf_log2, uint32_t phi_log2, uint32_t jump, uint32_t force_wide, int* rc_out, uint32_t seed_jump) {
    rig_options opt;
    std::memset(&opt, 0, sizeof(opt));
    opt.runs_per_block = K; opt.lf_bucket_log2 = lf_log2; opt.phi_bucket_log2 = phi_log2; opt.reserved[0] = jump;
    opt.reserved[1] = force_wide;  // bit0: 64-bit position words, bit1: 64-bit (not 40-bit packed) words inside the block records
    opt.reserved[2] = seed_jump;
    FlatHost* f = new FlatHost();
    int rc = rigf::flatten(*v, opt, *f);
    if (rc_out) *rc_out = rc;
    if (rc != RIG_OK) { delete f; return nullptr; }
    return f;
}
void fc_destroy(void* h) { delete (FlatHost*)h; }
uint64_t fc_bytes(void* h) { return ((FlatHost*)h)->bytes(); }
uint64_t fc_jump(void* h) { return ((FlatHost*)h)->phi.D; }
int fc_phi_packed(void* h) { return ((FlatHost*)h)->phi_packed ? 1 : 0; }
uint64_t fc_seed_jump(void* h) { return ((FlatHost*)h)->seed.J; }
uint64_t fc_seed_pieces(void* h) { return ((FlatHost*)h)->seed.pieces(); }
// Phi^SEG through the seed table (scalar and bucket-record lookup) == SEG applications of Phi
int fc_check_seed(void* h, uint64_t i) {
    const FlatHost& f = *(FlatHost*)h;
    if (f.seed.J < 2) return -1;
    u64 a = i;
    for (u32 j = 0; j < f.seed.J; ++j) a = f.phi.apply(a, 1, f.n);
    if (a != f.seed.apply(i, f.n)) return 1;
    if (a != seed_hop(f, i)) return 2;
    return 0;
}
uint64_t fc_pieces(void* h) { return ((FlatHost*)h)->phi.pieces(); }
int fc_w32(void* h) { return ((FlatHost*)h)->w32 ? 1 : 0; }
// Phi^j(i), j = 1..D, evaluated three ways must agree: j applications of Phi^1 through the scalar
// table, one scalar application of delta_j, and the bucket-record lookup the kernel performs.
int fc_check_jump(void* h, uint64_t i) {
    const FlatHost& f = *(FlatHost*)h;
    u64 e[8];
    phi_lookup(f, i, e);
    u64 a = i;
    for (u32 j = 1; j <= f.phi.D; ++j) {
        a = f.phi.apply(a, 1, f.n);
        if (a != f.phi.apply(i, j, f.n) || a != e[j - 1]) return (int)j;
    }
    return 0;
}

void fc_count(void* h, const uint8_t* patt, u64 N, u64 m, u64* lo, u64* hi) {
    const FlatHost& f = *(FlatHost*)h;
    for (u64 p = 0; p < N; ++p) { u64 k; search(f, patt + p * m, m, false, lo[p], hi[p], k); }
}

// Same decomposition as the GPU: ranges cut into one chain per overlapped run.
// occ_off must hold N+1 exclusive prefix sums; returns the number of chains.
uint64_t fc_locate(void* h, const uint8_t* patt, u64 N, u64 m, u64* lo, u64* hi, const u64* occ_off, u64* occ) {
    const FlatHost& f = *(FlatHost*)h;
    u64 chains = 0;
    std::vector<u64> items;
    for (u64 p = 0; p < N; ++p) {
        u64 k;
        search(f, patt + p * m, m, true, lo[p], hi[p], k);
        if (hi[p] < lo[p]) continue;
        u64 L = lo[p], H = hi[p];
        u64 jL = block_query(f, L, 0, 0).run, jR = block_query(f, H, 0, 0).run;
        for (u64 j = jL; j <= jR; ++j) {
            u64 sj = f.start[j], ej = f.start[j + 1] - 1;
            u64 top = std::min(H, ej), bot = std::max(L, sj);
            // mirrors rigk::phi_expand_kernel: toehold, then the whole chain (single pass) or the chain head
            // up to the next SEG-aligned output slot followed by one seed per window (two passes)
            u64 v = (top == H) ? k : (f.samples_last[j] + 1) % f.n;
            const u64 g0 = occ_off[p] + (H - top);
            occ[g0] = v;
            bool mis = false;
            if (f.seed.J < 2) {
                walk_chain(f, v, occ, g0 + 1, top - bot, mis);
            } else {
                const u64 SEG = f.seed.J, glast = g0 + (top - bot), a1 = (g0 + 15) / 16 * 16;
                v = walk_chain(f, v, occ, g0 + 1, std::min(a1, glast) - g0, mis);
                if (a1 <= glast) {
                    u64 sl = a1;
                    for (;;) {
                        items.push_back((sl << 8) | std::min<u64>(SEG - 1, glast - sl));
                        items.push_back(v);
                        sl += SEG;
                        if (sl > glast) break;
                        v = seed_hop(f, v);
                    }
                }
            }
            if (mis) return ~(u64)0;
            ++chains;
        }
    }
    if (f.seed.J >= 2 && items.size() / 2 > occ_off[N] / f.seed.J + chains) return ~(u64)0 - 1;  // the kernel's buffer bound
    for (size_t k = items.size() / 2; k-- > 0;)  // any order: items are independent
        if (item_fill(f, occ, items[2 * k], items[2 * k + 1])) return ~(u64)0 - 2;
    return chains;
}

}  // extern "C"
