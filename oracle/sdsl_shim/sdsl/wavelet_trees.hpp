// sdsl/wavelet_trees.hpp — stand-in for the ONE sdsl-lite header the reference includes
// (reference: internal/huff_string.hpp:17). TEST INFRASTRUCTURE ONLY: this lets the unmodified
// reference headers under /root/reference compile here (sdsl-lite is neither vendored nor
// installed) so the reference's own rle_string::rank / LF / count_and_get_occ /
// Phi logic can serve as oracle and CPU baseline. Nothing in the product links this.
//
// The primitives are restated from sdsl-lite's published algorithms (Elias-Fano sd_vector with
// sampled select on the high bits; Huffman-shaped wavelet tree with a rank_support_v style
// two-word directory per 512-bit block; packed int_vector), written from scratch. They are
// honest constant-time-ish structures, not binary-search toys, because the CPU baseline is
// timed through them. Results do not depend on these internals: they are fixed
// by the definitions rank(i) = #ones in [0,i), select(k) = position of the k-th one (1-based),
// wt.rank(i,c) = #c in [0,i), wt.select(k,c) = position of the k-th c (1-based), wt[i].
//
// Exact API surface provided = the reference's complete list of SDSL call sites.
#pragma once
#include <cstdint>
#include <cstring>
#include <cassert>
#include <chrono>   // the reference gets <chrono> transitively through SDSL (ri-count.cpp:53-55)
#include <string>
#include <vector>
#include <map>
#include <queue>
#include <memory>
#include <iostream>
#include <sstream>
#include <algorithm>
#include <immintrin.h>
#include "sais.hpp"  // product header (suffix sorting); its output is validated independently by tests/

namespace sdsl {

namespace shim {
inline uint64_t popcnt(uint64_t x) { return (uint64_t)__builtin_popcountll(x); }
inline unsigned select_in_word(uint64_t w, unsigned k) {  // position of the k-th (0-based) set bit
#if defined(__BMI2__)
    return (unsigned)__builtin_ctzll(_pdep_u64(uint64_t(1) << k, w));
#else
    for (unsigned i = 0; i < k; ++i) w &= w - 1;
    return (unsigned)__builtin_ctzll(w);
#endif
}
template <class T> inline uint64_t write_pod(std::ostream& out, const T& v) { out.write((const char*)&v, sizeof(T)); return sizeof(T); }
template <class T> inline void read_pod(std::istream& in, T& v) { in.read((char*)&v, sizeof(T)); }
inline uint64_t write_words(std::ostream& out, const std::vector<uint64_t>& w) {
    uint64_t k = w.size();
    out.write((const char*)&k, 8);
    out.write((const char*)w.data(), (std::streamsize)(k * 8));
    return 8 + k * 8;
}
inline void read_words(std::istream& in, std::vector<uint64_t>& w) {
    uint64_t k = 0;
    in.read((char*)&k, 8);
    w.resize(k);
    in.read((char*)w.data(), (std::streamsize)(k * 8));
}
}  // namespace shim

// ---------------------------------------------------------------- int_vector<W>
template <uint8_t W = 0>
class int_vector {
public:
    typedef uint64_t value_type;
    typedef uint64_t size_type;
    class reference {
        int_vector* v; size_type i;
    public:
        reference(int_vector* v_, size_type i_) : v(v_), i(i_) {}
        operator uint64_t() const { return v->get(i); }
        reference& operator=(uint64_t x) { v->set(i, x); return *this; }
        reference& operator=(const reference& o) { v->set(i, (uint64_t)o); return *this; }
    };
    int_vector() : m_size(0), m_width(W ? W : 64) {}
    int_vector(size_type n, uint64_t def = 0, uint8_t width = (W ? W : 64)) : m_size(n), m_width(W ? W : width) {
        if (m_width == 0) m_width = 1;
        m_data.assign((n * m_width + 63) / 64 + 1, 0);
        if (def) for (size_type i = 0; i < n; ++i) set(i, def);
    }
    size_type size() const { return m_size; }
    uint8_t width() const { return m_width; }
    bool empty() const { return m_size == 0; }
    inline uint64_t get(size_type i) const {
        uint64_t bit = i * m_width, w = bit >> 6, o = bit & 63;
        uint64_t x = m_data[w] >> o;
        if (o + m_width > 64) x |= m_data[w + 1] << (64 - o);
        return m_width == 64 ? x : (x & ((uint64_t(1) << m_width) - 1));
    }
    inline void set(size_type i, uint64_t x) {
        uint64_t bit = i * m_width, w = bit >> 6, o = bit & 63;
        uint64_t mask = m_width == 64 ? ~uint64_t(0) : ((uint64_t(1) << m_width) - 1);
        x &= mask;
        m_data[w] = (m_data[w] & ~(mask << o)) | (x << o);
        if (o + m_width > 64) {
            unsigned sh = 64 - (unsigned)o;
            m_data[w + 1] = (m_data[w + 1] & ~(mask >> sh)) | (x >> sh);
        }
    }
    uint64_t operator[](size_type i) const { return get(i); }
    reference operator[](size_type i) { return reference(this, i); }
    void resize(size_type n) { m_size = n; m_data.resize((n * m_width + 63) / 64 + 1, 0); }
    void push_back_value(uint64_t x) { resize(m_size + 1); set(m_size - 1, x); }
    const uint64_t* data() const { return m_data.data(); }
    size_type serialize(std::ostream& out) const {
        size_type w = shim::write_pod(out, m_size);
        w += shim::write_pod(out, m_width);
        w += shim::write_words(out, m_data);
        return w;
    }
    void load(std::istream& in) {
        shim::read_pod(in, m_size);
        shim::read_pod(in, m_width);
        shim::read_words(in, m_data);
    }
private:
    size_type m_size;
    uint8_t m_width;
    std::vector<uint64_t> m_data;
};
typedef int_vector<1> bit_vector;

// ---------------------------------------------------------------- plain bit array with rank/select
// (used for the high part of sd_vector and for the wavelet-tree levels)
namespace shim {
class RankedBits {
public:
    std::vector<uint64_t> bits;   // raw bits, padded to a multiple of 8 words
    std::vector<uint64_t> dir;    // 2 words per 512-bit block: absolute rank, 7 x 9-bit in-block ranks
    std::vector<uint32_t> sel1;   // block index holding every 512-th one
    std::vector<uint32_t> sel0;   // block index holding every 512-th zero
    uint64_t len = 0, ones = 0;
    void init(uint64_t nbits) { len = nbits; bits.assign(((nbits + 511) / 512 + 1) * 8, 0); }
    inline void set(uint64_t i) { bits[i >> 6] |= uint64_t(1) << (i & 63); }
    inline bool get(uint64_t i) const { return (bits[i >> 6] >> (i & 63)) & 1; }
    void build(bool with_select0) {
        uint64_t nb = bits.size() / 8;
        dir.assign(nb * 2, 0);
        sel1.clear(); sel0.clear();
        uint64_t acc = 0, zacc = 0;
        for (uint64_t b = 0; b < nb; ++b) {
            dir[2 * b] = acc;
            uint64_t rel = 0, packed = 0;
            for (unsigned w = 0; w < 8; ++w) {
                if (w) packed |= rel << (9 * (w - 1));
                rel += popcnt(bits[8 * b + w]);
            }
            dir[2 * b + 1] = packed;
            // select samples: first block whose cumulative count reaches k*512+1
            while ((uint64_t)sel1.size() * 512 < acc + rel && (uint64_t)sel1.size() * 512 >= acc) sel1.push_back((uint32_t)b);
            if (with_select0) {
                uint64_t zrel = 512 - rel;
                while ((uint64_t)sel0.size() * 512 < zacc + zrel && (uint64_t)sel0.size() * 512 >= zacc) sel0.push_back((uint32_t)b);
                zacc += zrel;
            }
            acc += rel;
        }
        ones = 0;
        for (uint64_t i = 0; i < (len + 63) / 64; ++i) ones += popcnt(bits[i]);
    }
    inline uint64_t rank1(uint64_t i) const {  // ones in [0,i)
        uint64_t b = i >> 9, w = (i >> 6) & 7;
        uint64_t r = dir[2 * b] + (w ? ((dir[2 * b + 1] >> (9 * (w - 1))) & 511) : 0);
        uint64_t o = i & 63;
        return r + (o ? popcnt(bits[i >> 6] & ((uint64_t(1) << o) - 1)) : 0);
    }
    inline uint64_t block_rank1(uint64_t b) const { return dir[2 * b]; }
    inline uint64_t block_rank0(uint64_t b) const { return b * 512 - dir[2 * b]; }
    // position of the k-th one, k 0-based, k < ones
    inline uint64_t select1(uint64_t k) const {
        uint64_t nb = dir.size() / 2;
        uint64_t lo = sel1[k >> 9], hi = ((k >> 9) + 1 < sel1.size()) ? sel1[(k >> 9) + 1] + 1 : nb;
        while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (block_rank1(mid) <= k) lo = mid; else hi = mid; }
        uint64_t rem = k - block_rank1(lo);
        for (unsigned w = 0; w < 8; ++w) {
            uint64_t c = popcnt(bits[8 * lo + w]);
            if (rem < c) return lo * 512 + w * 64 + select_in_word(bits[8 * lo + w], (unsigned)rem);
            rem -= c;
        }
        assert(false); return 0;
    }
    // position of the k-th zero, k 0-based (zeros in the padding count; callers stay below len)
    inline uint64_t select0(uint64_t k) const {
        uint64_t nb = dir.size() / 2;
        uint64_t lo = sel0[k >> 9], hi = ((k >> 9) + 1 < sel0.size()) ? sel0[(k >> 9) + 1] + 1 : nb;
        while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (block_rank0(mid) <= k) lo = mid; else hi = mid; }
        uint64_t rem = k - block_rank0(lo);
        for (unsigned w = 0; w < 8; ++w) {
            uint64_t c = 64 - popcnt(bits[8 * lo + w]);
            if (rem < c) return lo * 512 + w * 64 + select_in_word(~bits[8 * lo + w], (unsigned)rem);
            rem -= c;
        }
        assert(false); return 0;
    }
    uint64_t serialize(std::ostream& out) const { uint64_t w = write_pod(out, len); return w + write_words(out, bits); }
    void load(std::istream& in, bool with_select0) { read_pod(in, len); read_words(in, bits); build(with_select0); }
};
}  // namespace shim

// ---------------------------------------------------------------- sd_vector (Elias-Fano)
template <class A = void, class B = void, class C = void>
class sd_vector {
public:
    typedef uint64_t size_type;
    class rank_1_type {
        const sd_vector* v;
    public:
        rank_1_type(const sd_vector* v_ = nullptr) : v(v_) {}
        size_type operator()(size_type i) const { return v->rank1(i); }
    };
    class select_1_type {
        const sd_vector* v;
    public:
        select_1_type(const sd_vector* v_ = nullptr) : v(v_) {}
        size_type operator()(size_type k) const { return v->select1(k); }  // 1-based
    };
    sd_vector() {}
    explicit sd_vector(const bit_vector& bv) {
        m_u = bv.size();
        m_m = 0;
        for (size_type i = 0; i < m_u; ++i) m_m += bv[i];
        m_wl = 0;
        if (m_m) { uint64_t q = m_u / m_m; while ((uint64_t(2) << m_wl) <= q) ++m_wl; }
        m_low = int_vector<>(m_m, 0, m_wl ? m_wl : 1);
        m_high.init(m_m + (m_u >> m_wl) + 2);
        size_type k = 0;
        for (size_type i = 0; i < m_u; ++i)
            if (bv[i]) {
                if (m_wl) m_low[k] = i & ((uint64_t(1) << m_wl) - 1);
                m_high.set((i >> m_wl) + k);
                ++k;
            }
        m_high.build(true);
    }
    size_type size() const { return m_u; }
    size_type ones() const { return m_m; }
    bool operator[](size_type i) const { return rank1(i + 1) - rank1(i); }
    // ones in [0,i), 0 <= i <= size()
    inline size_type rank1(size_type i) const {
        if (m_m == 0) return 0;
        if (i >= m_u) return m_m;
        uint64_t h = i >> m_wl, l = m_wl ? (i & ((uint64_t(1) << m_wl) - 1)) : 0;
        // elements with high part h sit between the (h-1)-th and h-th zero of `high` (0-based zeros)
        uint64_t p_end = m_high.select0(h);           // position of the zero closing bucket h
        uint64_t idx_end = p_end - h;                 // #ones before it = #elements with high <= h
        uint64_t p = p_end;
        // walk back over the ones of bucket h while their low part is >= l
        while (p > 0 && m_high.get(p - 1)) {
            uint64_t idx = idx_end - (p_end - p) - 1;
            uint64_t lowv = m_wl ? m_low.get(idx) : 0;
            if (lowv < l) break;
            --p;
        }
        return idx_end - (p_end - p);
    }
    inline size_type select1(size_type k) const {  // k 1-based
        uint64_t hp = m_high.select1(k - 1);
        uint64_t hv = hp - (k - 1);
        return (hv << m_wl) | (m_wl ? m_low.get(k - 1) : 0);
    }
    size_type serialize(std::ostream& out) const {
        size_type w = shim::write_pod(out, m_u);
        w += shim::write_pod(out, m_m);
        w += shim::write_pod(out, m_wl);
        w += m_low.serialize(out);
        w += m_high.serialize(out);
        return w;
    }
    void load(std::istream& in) {
        shim::read_pod(in, m_u); shim::read_pod(in, m_m); shim::read_pod(in, m_wl);
        m_low.load(in);
        m_high.load(in, true);
    }
private:
    size_type m_u = 0, m_m = 0;
    uint8_t m_wl = 0;
    int_vector<> m_low;
    shim::RankedBits m_high;
};

// hyb_vector<> is only instantiated (never exercised) by the reference (ri-count.cpp:162-165,
// `hyb=false`); an alias of the Elias-Fano shim satisfies the compiler.
template <uint32_t K = 16>
class hyb_vector : public sd_vector<> {
public:
    hyb_vector() {}
    explicit hyb_vector(const bit_vector& bv) : sd_vector<>(bv) {}
    class rank_1_type : public sd_vector<>::rank_1_type {
    public:
        rank_1_type(const hyb_vector* v = nullptr) : sd_vector<>::rank_1_type(v) {}
    };
    class select_1_type : public sd_vector<>::select_1_type {
    public:
        select_1_type(const hyb_vector* v = nullptr) : sd_vector<>::select_1_type(v) {}
    };
};

// ---------------------------------------------------------------- wt_huff (Huffman-shaped wavelet tree over bytes)
template <class A = void, class B = void, class C = void, class D = void>
class wt_huff {
public:
    typedef uint64_t size_type;
    typedef uint8_t value_type;
    wt_huff() {}
    size_type size() const { return m_size; }

    void build(const uint8_t* s, size_type n) {
        m_size = n;
        uint64_t freq[256] = {0};
        for (size_type i = 0; i < n; ++i) freq[s[i]]++;
        // Huffman tree; ties broken by (weight, smallest symbol in subtree) for determinism.
        struct N { uint64_t w; int minsym; int left, right, sym; };
        std::vector<N> nodes;
        typedef std::pair<std::pair<uint64_t, int>, int> QE;
        std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
        for (int c = 0; c < 256; ++c)
            if (freq[c]) { nodes.push_back({freq[c], c, -1, -1, c}); pq.push({{freq[c], c}, (int)nodes.size() - 1}); }
        m_sigma = (uint32_t)nodes.size();
        for (int c = 0; c < 256; ++c) { m_code_len[c] = 0; m_code[c] = 0; m_leaf_node[c] = -1; }
        m_nodes.clear();
        if (m_sigma == 0) { m_bits.init(0); m_bits.build(true); return; }
        while (pq.size() > 1) {
            QE a = pq.top(); pq.pop();
            QE b = pq.top(); pq.pop();
            nodes.push_back({a.first.first + b.first.first, std::min(a.first.second, b.first.second), a.second, b.second, -1});
            pq.push({{nodes.back().w, nodes.back().minsym}, (int)nodes.size() - 1});
        }
        int root = pq.top().second;
        // Lay the internal nodes out in BFS order, each owning a contiguous bit range.
        struct Q { int hn; int parent; int bit; };
        std::vector<Q> order;
        order.push_back({root, -1, 0});
        std::vector<int> hn_to_wn(nodes.size(), -1);
        uint64_t off = 0;
        for (size_t qi = 0; qi < order.size(); ++qi) {
            Q q = order[qi];
            const N& hn = nodes[q.hn];
            WNode wn;
            wn.parent = q.parent < 0 ? -1 : hn_to_wn[q.parent];
            wn.parent_bit = (uint8_t)q.bit;
            wn.sym = hn.sym;
            wn.child[0] = wn.child[1] = -1;
            wn.off = off; wn.len = hn.w; wn.ones_before = 0;
            if (hn.sym < 0) off += hn.w;
            hn_to_wn[q.hn] = (int)m_nodes.size();
            if (wn.parent >= 0) m_nodes[wn.parent].child[q.bit] = (int)m_nodes.size();
            m_nodes.push_back(wn);
            if (hn.sym < 0) { order.push_back({hn.left, q.hn, 0}); order.push_back({hn.right, q.hn, 1}); }
            else m_leaf_node[hn.sym] = hn_to_wn[q.hn];
        }
        // codes
        for (int c = 0; c < 256; ++c) {
            if (m_leaf_node[c] < 0) continue;
            uint64_t code = 0; unsigned len = 0;
            std::vector<uint8_t> path;
            for (int v = m_leaf_node[c]; m_nodes[v].parent >= 0; v = m_nodes[v].parent) path.push_back(m_nodes[v].parent_bit);
            for (size_t k = path.size(); k-- > 0;) { code |= (uint64_t)path[k] << len; ++len; }
            m_code[c] = code; m_code_len[c] = (uint8_t)len;  // bit t of code = branch taken at depth t
        }
        // fill the bits: every symbol writes one bit in each internal node on its path
        m_bits.init(off);
        std::vector<uint64_t> cursor(m_nodes.size());
        for (size_t v = 0; v < m_nodes.size(); ++v) cursor[v] = m_nodes[v].off;
        for (size_type i = 0; i < n; ++i) {
            int v = 0;
            uint64_t code = m_code[s[i]];
            for (unsigned d = 0; d < m_code_len[s[i]]; ++d) {
                unsigned b = (code >> d) & 1;
                if (b) m_bits.set(cursor[v]);
                cursor[v]++;
                v = m_nodes[v].child[b];
            }
        }
        m_bits.build(true);
        for (auto& wn : m_nodes) if (wn.sym < 0) wn.ones_before = m_bits.rank1(wn.off);
    }

    value_type operator[](size_type i) const {
        int v = 0;
        while (m_nodes[v].sym < 0) {
            const WNode& w = m_nodes[v];
            bool b = m_bits.get(w.off + i);
            uint64_t r1 = m_bits.rank1(w.off + i) - w.ones_before;
            i = b ? r1 : i - r1;
            v = w.child[b];
        }
        return (value_type)m_nodes[v].sym;
    }
    size_type rank(size_type i, value_type c) const {  // #c in [0,i)
        if (m_leaf_node[c] < 0) return 0;
        int v = 0;
        uint64_t code = m_code[c];
        for (unsigned d = 0; d < m_code_len[c] && i > 0; ++d) {
            const WNode& w = m_nodes[v];
            unsigned b = (code >> d) & 1;
            uint64_t r1 = m_bits.rank1(w.off + i) - w.ones_before;
            i = b ? r1 : i - r1;
            v = w.child[b];
        }
        return i;
    }
    size_type select(size_type k, value_type c) const {  // position of the k-th c, k 1-based
        int v = m_leaf_node[c];
        uint64_t pos = k - 1;  // 0-based index inside the current node
        while (m_nodes[v].parent >= 0) {
            const WNode& p = m_nodes[m_nodes[v].parent];
            if (m_nodes[v].parent_bit) pos = m_bits.select1(p.ones_before + pos) - p.off;
            else pos = m_bits.select0((p.off - p.ones_before) + pos) - p.off;
            v = m_nodes[v].parent;
        }
        return pos;
    }
    size_type serialize(std::ostream& out) const {
        size_type w = shim::write_pod(out, m_size);
        w += shim::write_pod(out, m_sigma);
        uint64_t nn = m_nodes.size();
        w += shim::write_pod(out, nn);
        for (const auto& x : m_nodes) w += shim::write_pod(out, x);
        out.write((const char*)m_code, sizeof(m_code)); w += sizeof(m_code);
        out.write((const char*)m_code_len, sizeof(m_code_len)); w += sizeof(m_code_len);
        out.write((const char*)m_leaf_node, sizeof(m_leaf_node)); w += sizeof(m_leaf_node);
        w += m_bits.serialize(out);
        return w;
    }
    void load(std::istream& in) {
        shim::read_pod(in, m_size); shim::read_pod(in, m_sigma);
        uint64_t nn = 0; shim::read_pod(in, nn);
        m_nodes.resize(nn);
        for (auto& x : m_nodes) shim::read_pod(in, x);
        in.read((char*)m_code, sizeof(m_code));
        in.read((char*)m_code_len, sizeof(m_code_len));
        in.read((char*)m_leaf_node, sizeof(m_leaf_node));
        m_bits.load(in, true);
    }
private:
    struct WNode { uint64_t off, len, ones_before; int parent; int child[2]; int sym; uint8_t parent_bit; };
    size_type m_size = 0;
    uint32_t m_sigma = 0;
    std::vector<WNode> m_nodes;
    uint64_t m_code[256];
    uint8_t m_code_len[256];
    int m_leaf_node[256];
    shim::RankedBits m_bits;
};

// construct_im(wt, c_string, 1): byte sequence up to the terminating NUL (huff_string.hpp:33)
template <class WT>
inline void construct_im(WT& wt, const char* s, uint8_t num_bytes) {
    (void)num_bytes;
    wt.build((const uint8_t*)s, std::strlen(s));
}

// ---------------------------------------------------------------- suffix-array construction plumbing
// (reference: internal/r_index.hpp:557-575, 629-630). The reference streams the SA from SDSL's
// on-disk cache; here the cache_config object owns the text and the SA in memory.
struct cache_config {
    std::vector<uint8_t> text;
    std::vector<int32_t> sa32;
    std::vector<int64_t> sa64;
    bool wide = false;
};
namespace conf {
static const char* const KEY_TEXT = "text";
static const char* const KEY_SA = "sa";
}
enum byte_sa_algo_type { LIBDIVSUFSORT, SE_SAIS };
struct construct_config { static byte_sa_algo_type byte_algo_sa; };
inline byte_sa_algo_type construct_config::byte_algo_sa = LIBDIVSUFSORT;

inline void append_zero_symbol(int_vector<8>& text) { text.push_back_value(0); }
inline void store_to_cache(const int_vector<8>& text, const char* key, cache_config& cc) {
    (void)key;
    cc.text.resize(text.size());
    for (uint64_t i = 0; i < text.size(); ++i) cc.text[i] = (uint8_t)text[i];
}
template <uint8_t W>
inline void construct_sa(cache_config& cc) {
    uint64_t n = cc.text.size();
    if (n < (uint64_t(1) << 31) - 2) {
        cc.wide = false; cc.sa32.resize(n);
        rib::suffix_array_with_sentinel<int32_t>(cc.text.data(), (int32_t)n, cc.sa32.data());
    } else {
        cc.wide = true; cc.sa64.resize(n);
        rib::suffix_array_with_sentinel<int64_t>(cc.text.data(), (int64_t)n, cc.sa64.data());
    }
}
inline std::string cache_file_name(const char* key, const cache_config& cc) {
    std::ostringstream ss;
    ss << "shim:" << (const void*)&cc << ":" << key;
    return ss.str();
}
inline cache_config* shim_parse_cc(const std::string& name, std::string* key = nullptr) {
    void* p = nullptr;
    size_t a = name.find(':'), b = name.rfind(':');
    std::istringstream ss(name.substr(a + 1, b - a - 1));
    ss >> p;
    if (key) *key = name.substr(b + 1);
    return (cache_config*)p;
}
inline int remove(const std::string& name) {
    std::string key;
    cache_config* cc = shim_parse_cc(name, &key);
    if (key == "text") std::vector<uint8_t>().swap(cc->text);
    else { std::vector<int32_t>().swap(cc->sa32); std::vector<int64_t>().swap(cc->sa64); }
    return 0;
}
template <uint8_t W = 0>
class int_vector_buffer {
    cache_config* cc;
public:
    explicit int_vector_buffer(const std::string& name) : cc(shim_parse_cc(name)) {}
    uint64_t size() const { return cc->wide ? cc->sa64.size() : cc->sa32.size(); }
    uint64_t operator[](uint64_t i) const { return cc->wide ? (uint64_t)cc->sa64[i] : (uint64_t)cc->sa32[i]; }
};

}  // namespace sdsl
