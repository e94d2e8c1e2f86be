// sais.hpp — suffix array construction by induced sorting (SA-IS, Nong/Zhang/Chan 2009).
//
// Replaces the SDSL call the reference makes at build time
// (reference: internal/r_index.hpp:571-572, `construct_sa<8>(cc)` with SE_SAIS / LIBDIVSUFSORT):
// the suffix array of text·\0, where \0 is a unique smallest sentinel, so SA[0] = n-1.
//
// Own implementation, templated on the symbol type (bytes at level 0, names in the
// recursion) and on the index type (int32_t while n < 2^31, int64_t above: configs 4-5).
// Memory: the SA itself + n/8 bytes of type bits + O(alphabet) buckets per level.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <type_traits>

namespace rib {
namespace sais_detail {

struct TypeBits {
    std::vector<uint64_t> w;
    explicit TypeBits(size_t n) : w((n + 63) / 64, 0) {}
    inline bool get(size_t i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    inline void set(size_t i, bool b) {
        if (b) w[i >> 6] |= (uint64_t(1) << (i & 63));
        else   w[i >> 6] &= ~(uint64_t(1) << (i & 63));
    }
};

template <class S, class I>
static void bucket_bounds(const S* s, std::vector<I>& bkt, I n, I K, bool ends) {
    for (I i = 0; i <= K; ++i) bkt[i] = 0;
    for (I i = 0; i < n; ++i) bkt[(I)s[i]]++;
    I sum = 0;
    for (I i = 0; i <= K; ++i) {
        sum += bkt[i];
        bkt[i] = ends ? sum : sum - bkt[i];
    }
}

// s has n symbols in [0,K]; s[n-1] is the unique smallest symbol. SA has room for n entries.
template <class S, class I>
static void sais_rec(const S* s, I* SA, I n, I K) {
    static_assert(std::is_signed<I>::value, "index type must be signed (-1 marks empty slots)");
    TypeBits t((size_t)n);  // 1 = S-type, 0 = L-type
    t.set(n - 1, true);
    if (n >= 2) t.set(n - 2, false);
    for (I i = n - 3; i >= 0; --i)
        t.set(i, (s[i] < s[i + 1]) || (s[i] == s[i + 1] && t.get(i + 1)));
    auto is_lms = [&](I i) { return i > 0 && t.get(i) && !t.get(i - 1); };

    std::vector<I> bkt((size_t)K + 1);
    auto induce_L = [&]() {
        bucket_bounds(s, bkt, n, K, false);
        for (I i = 0; i < n; ++i) {
            I j = SA[i] - 1;
            if (SA[i] > 0 && !t.get(j)) SA[bkt[(I)s[j]]++] = j;
        }
    };
    auto induce_S = [&]() {
        bucket_bounds(s, bkt, n, K, true);
        for (I i = n - 1; i >= 0; --i) {
            I j = SA[i] - 1;
            if (SA[i] > 0 && t.get(j)) SA[--bkt[(I)s[j]]] = j;
        }
    };

    // Stage 1: sort the LMS substrings.
    bucket_bounds(s, bkt, n, K, true);
    for (I i = 0; i < n; ++i) SA[i] = -1;
    for (I i = 1; i < n; ++i)
        if (is_lms(i)) SA[--bkt[(I)s[i]]] = i;
    induce_L();
    induce_S();

    // Compact the sorted LMS substrings, then name them.
    I n1 = 0;
    for (I i = 0; i < n; ++i)
        if (is_lms(SA[i])) SA[n1++] = SA[i];
    for (I i = n1; i < n; ++i) SA[i] = -1;
    I name = 0, prev = -1;
    for (I i = 0; i < n1; ++i) {
        I pos = SA[i];
        bool diff = false;
        if (prev < 0) diff = true;
        else {
            for (I d = 0;; ++d) {
                if (s[pos + d] != s[prev + d] || t.get(pos + d) != t.get(prev + d)) { diff = true; break; }
                if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) break;
            }
        }
        if (diff) { ++name; prev = pos; }
        SA[n1 + pos / 2] = name - 1;
    }
    for (I i = n - 1, j = n - 1; i >= n1; --i)
        if (SA[i] >= 0) SA[j--] = SA[i];

    // Stage 2: suffix array of the reduced string.
    I* SA1 = SA;
    I* s1 = SA + (n - n1);
    if (name < n1) sais_rec<I, I>(s1, SA1, n1, name - 1);
    else for (I i = 0; i < n1; ++i) SA1[s1[i]] = i;

    // Stage 3: induce the full SA from the sorted LMS suffixes.
    bucket_bounds(s, bkt, n, K, true);
    for (I i = 1, j = 0; i < n; ++i)
        if (is_lms(i)) s1[j++] = i;
    for (I i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
    for (I i = n1; i < n; ++i) SA[i] = -1;
    for (I i = n1 - 1; i >= 0; --i) {
        I j = SA[i];
        SA[i] = -1;
        SA[--bkt[(I)s[j]]] = j;
    }
    induce_L();
    induce_S();
}

}  // namespace sais_detail

// Suffix array of text[0..len) followed by a virtual 0 byte (the text must not contain 0x00).
// Output: SA has len+1 entries, SA[0] == len. `buf` must already hold text plus the trailing 0.
template <class I>
inline void suffix_array_with_sentinel(const uint8_t* text_with_zero, I n_with_zero, I* SA) {
    if (n_with_zero == 1) { SA[0] = 0; return; }
    sais_detail::sais_rec<uint8_t, I>(text_with_zero, SA, n_with_zero, (I)255);
}

}  // namespace rib
