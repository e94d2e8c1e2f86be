// textgen.hpp — deterministic synthetic repetitive texts + Pizza&Chili pattern files for the
// BASELINE.json configs (concretised in SURVEY.md §8d). splitmix64-seeded xoshiro256**, no
// std:: distributions (their output is library-specific), so CPU and GPU boxes see identical bytes.
//
// Pattern file shape follows the reference's reader (utils.hpp:57-91, ri-count.cpp:86-110):
// one header line "# number=N length=M file=F forbidden=\n", then N*M raw bytes.
#pragma once
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <stdexcept>

namespace rib {

struct Rng {
    uint64_t s[4];
    explicit Rng(uint64_t seed) {
        uint64_t z = seed;
        for (int i = 0; i < 4; ++i) {  // splitmix64
            z += 0x9E3779B97F4A7C15ull;
            uint64_t x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
            s[i] = x ^ (x >> 31);
        }
    }
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    inline uint64_t next() {  // xoshiro256**
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    inline uint64_t below(uint64_t bound) {  // multiply-shift, bias < 2^-64*bound: irrelevant here
        return (uint64_t)(((unsigned __int128)next() * bound) >> 64);
    }
    inline double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

static const char kDna[4] = {'A', 'C', 'G', 'T'};

// C2: base of `base_len` iid ACGT; copies until `n` bytes; copy k = copy k-1 with `snps` random
// single-base substitutions (cumulative drift).
inline std::vector<uint8_t> gen_dna_drift(uint64_t n, uint64_t base_len, uint64_t snps, uint64_t seed) {
    Rng g(seed);
    std::vector<uint8_t> cur(base_len), out;
    out.reserve(n);
    for (auto& c : cur) c = kDna[g.below(4)];
    while (out.size() < n) {
        uint64_t take = std::min<uint64_t>(base_len, n - out.size());
        out.insert(out.end(), cur.begin(), cur.begin() + take);
        for (uint64_t k = 0; k < snps; ++k) {
            uint64_t p = g.below(base_len);
            uint8_t c;
            do c = kDna[g.below(4)]; while (c == cur[p]);
            cur[p] = c;
        }
    }
    return out;
}

// C5: every copy = base with independent substitutions at rate `rate` per base.
inline std::vector<uint8_t> gen_dna_indep(uint64_t n, uint64_t base_len, double rate, uint64_t seed) {
    Rng g(seed);
    std::vector<uint8_t> base(base_len), out;
    out.reserve(n);
    for (auto& c : base) c = kDna[g.below(4)];
    while (out.size() < n) {
        uint64_t take = std::min<uint64_t>(base_len, n - out.size());
        size_t off = out.size();
        out.insert(out.end(), base.begin(), base.begin() + take);
        // number of SNPs ~ Binomial(base_len, rate) approximated by floor(mean) + Bernoulli(frac)
        double mean = rate * (double)base_len;
        uint64_t k = (uint64_t)mean + (g.unit() < (mean - std::floor(mean)) ? 1 : 0);
        for (uint64_t t = 0; t < k; ++t) {
            uint64_t p = g.below(base_len);
            if (p >= take) continue;
            uint8_t c;
            do c = kDna[g.below(4)]; while (c == base[p]);
            out[off + p] = c;
        }
    }
    return out;
}

// C3: versioned document over `sigma` printable symbols (0x20.. and '\n'), Zipf(1.0) symbol
// distribution; each version = previous with probability `p_edit` of one edit (50% substitute,
// 25% insert, 25% delete of a 1-8 byte span).
inline std::vector<uint8_t> gen_versioned_doc(uint64_t n, uint64_t base_len, unsigned sigma, double p_edit,
                                              uint64_t seed) {
    if (sigma < 2 || sigma > 96) throw std::invalid_argument("sigma must be in [2,96]");
    Rng g(seed);
    std::vector<uint8_t> alpha(sigma);
    for (unsigned i = 0; i < sigma; ++i) alpha[i] = (i + 1 == sigma && sigma == 96) ? 0x0A : (uint8_t)(0x20 + i);
    std::vector<double> cdf(sigma);
    double h = 0;
    for (unsigned i = 0; i < sigma; ++i) { h += 1.0 / (double)(i + 1); cdf[i] = h; }
    auto sym = [&]() -> uint8_t {
        double u = g.unit() * h;
        unsigned i = (unsigned)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
        return alpha[std::min(i, sigma - 1)];
    };
    std::vector<uint8_t> cur(base_len), out;
    out.reserve(n + base_len + 16);
    for (auto& c : cur) c = sym();
    while (out.size() < n) {
        out.insert(out.end(), cur.begin(), cur.end());
        if (g.unit() < p_edit) {
            uint64_t span = 1 + g.below(8);
            uint64_t kind = g.below(4);  // 0,1 substitute; 2 insert; 3 delete
            uint64_t p = g.below(cur.size() > span ? cur.size() - span : 1);
            if (kind <= 1) { for (uint64_t t = 0; t < span && p + t < cur.size(); ++t) cur[p + t] = sym(); }
            else if (kind == 2) { std::vector<uint8_t> ins(span); for (auto& c : ins) c = sym(); cur.insert(cur.begin() + p, ins.begin(), ins.end()); }
            else if (cur.size() > 2 * span) cur.erase(cur.begin() + p, cur.begin() + p + span);
        }
    }
    out.resize(n);
    return out;
}

// C4: pan-genome. Base of `base_len` iid ACGT; pool of `n_sites` variant sites (90% SNP, 10%
// 1-10 bp indel), each with allele frequency ~ Beta(0.2,0.8) (Joehnk); haplotypes carry each
// variant independently with that frequency and are separated by 'N'.
inline std::vector<uint8_t> gen_pangenome(uint64_t n, uint64_t base_len, uint64_t n_sites, uint64_t seed) {
    Rng g(seed);
    std::vector<uint8_t> base(base_len);
    for (auto& c : base) c = kDna[g.below(4)];
    struct Site { uint64_t pos; uint8_t kind; uint8_t len; uint8_t alt[10]; double freq; };
    std::vector<Site> sites(n_sites);
    for (auto& s : sites) {
        s.pos = g.below(base_len);
        bool snp = g.below(10) != 0;
        if (snp) { s.kind = 0; s.len = 1; do s.alt[0] = kDna[g.below(4)]; while (s.alt[0] == base[s.pos]); }
        else { s.kind = (uint8_t)(1 + g.below(2)); s.len = (uint8_t)(1 + g.below(10)); for (int t = 0; t < s.len; ++t) s.alt[t] = kDna[g.below(4)]; }
        double x, y;
        do { x = std::pow(g.unit(), 1.0 / 0.2); y = std::pow(g.unit(), 1.0 / 0.8); } while (x + y > 1.0 || x + y == 0.0);
        s.freq = x / (x + y);
    }
    std::sort(sites.begin(), sites.end(), [](const Site& a, const Site& b) { return a.pos < b.pos; });
    std::vector<uint8_t> out;
    out.reserve(n + base_len + 64);
    while (out.size() < n) {
        uint64_t p = 0;
        for (const auto& s : sites) {
            if (s.pos < p) continue;
            if (g.unit() >= s.freq) continue;
            out.insert(out.end(), base.begin() + p, base.begin() + s.pos);
            p = s.pos;
            if (s.kind == 0) { out.push_back(s.alt[0]); p += 1; }
            else if (s.kind == 1) { out.insert(out.end(), s.alt, s.alt + s.len); }       // insertion
            else { p = std::min<uint64_t>(base_len, p + s.len); }                         // deletion
        }
        out.insert(out.end(), base.begin() + p, base.end());
        out.push_back('N');
    }
    out.resize(n);
    return out;
}

// N patterns of length m, starts uniform in [0, limit-m] (limit = text length, or the first copy
// for C5). Returns header + bodies; `bodies_off` receives the offset of the first pattern byte.
inline std::vector<uint8_t> gen_patterns(const uint8_t* text, uint64_t text_len, uint64_t N, uint64_t m,
                                         uint64_t start_limit, uint64_t seed, const std::string& file_label,
                                         uint64_t* bodies_off = nullptr) {
    if (start_limit == 0 || start_limit > text_len) start_limit = text_len;
    if (m > start_limit) throw std::invalid_argument("pattern longer than text");
    Rng g(seed);
    std::string hdr = "# number=" + std::to_string(N) + " length=" + std::to_string(m) + " file=" + file_label + " forbidden=\n";
    std::vector<uint8_t> out(hdr.begin(), hdr.end());
    if (bodies_off) *bodies_off = out.size();
    out.reserve(out.size() + N * m);
    for (uint64_t i = 0; i < N; ++i) {
        uint64_t p = g.below(start_limit - m + 1);
        out.insert(out.end(), text + p, text + p + m);
    }
    return out;
}

}  // namespace rib
