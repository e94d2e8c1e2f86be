// rindex_host.cpp — C ABI over the host-side builder / container / generators (include/rindex_host.h).
#include "../../include/rindex_host.h"
#include "logical_index.hpp"
#include "pfp_builder.hpp"
#include "textgen.hpp"
#include <new>

struct rih_index {
    rib::LogicalIndex L;
};

extern "C" {

int rih_build_from_text(const uint8_t* text, uint64_t len, rih_index** out) {
    if (!out || (!text && len)) return RIH_ERR_ARG;
    try {
        rih_index* h = new rih_index();
        h->L = rib::build_logical_index(text, len);
        *out = h;
        return RIH_OK;
    } catch (const std::invalid_argument&) {
        return RIH_ERR_RESERVED_CHARS;
    } catch (const std::bad_alloc&) {
        return RIH_ERR_ARG;
    }
}

int rih_build_from_text_pfp(const uint8_t* text, uint64_t len, uint32_t w, uint32_t p, uint64_t stats_out[6], rih_index** out) {
    if (!out || (!text && len)) return RIH_ERR_ARG;
    try {
        rih_index* h = new rih_index();
        rib::pfp::Params prm;
        if (w) prm.w = w;
        if (p) prm.p = p;
        rib::pfp::Stats st;
        h->L = rib::pfp::build(text, len, prm, &st);
        if (stats_out) {
            stats_out[0] = st.phrases; stats_out[1] = st.dict_bytes; stats_out[2] = st.parse_len;
            stats_out[3] = st.groups; stats_out[4] = st.uniform_rows; stats_out[5] = st.merged_rows;
        }
        *out = h;
        return RIH_OK;
    } catch (const std::invalid_argument&) {
        return RIH_ERR_RESERVED_CHARS;
    } catch (const std::exception&) {
        return RIH_ERR_ARG;
    }
}

int rih_build_auto(const uint8_t* text, uint64_t len, int* used_pfp, rih_index** out) {
    if (!out || (!text && len)) return RIH_ERR_ARG;
    try {
        rih_index* h = new rih_index();
        bool pf = false;
        h->L = rib::build_logical_index_auto(text, len, &pf);
        if (used_pfp) *used_pfp = pf ? 1 : 0;
        *out = h;
        return RIH_OK;
    } catch (const std::invalid_argument&) {
        return RIH_ERR_RESERVED_CHARS;
    } catch (const std::exception&) {
        return RIH_ERR_ARG;
    }
}

void rih_destroy(rih_index* idx) { delete idx; }

int rih_view(const rih_index* idx, rig_logical_view* v) {
    if (!idx || !v) return RIH_ERR_ARG;
    const rib::LogicalIndex& L = idx->L;
    v->n = L.n; v->r = L.r; v->F = L.F;
    v->run_heads = L.run_heads.data(); v->run_lens = L.run_lens.data();
    v->samples_last = L.samples_last.data(); v->pred_pos = L.pred_pos.data(); v->pred_to_run = L.pred_to_run.data();
    return RIH_OK;
}

int rih_save(const rih_index* idx, const char* path, int with_flag_byte) {
    if (!idx || !path) return RIH_ERR_ARG;
    std::ofstream out(path, std::ios::binary);
    if (!out) return RIH_ERR_IO;
    if (with_flag_byte) { bool fast = false; out.write((char*)&fast, sizeof(fast)); }
    rib::serialize(idx->L, out);
    return out ? RIH_OK : RIH_ERR_IO;
}

int rih_load(const char* path, int with_flag_byte, rih_index** out) {
    if (!path || !out) return RIH_ERR_ARG;
    std::ifstream in(path, std::ios::binary);
    if (!in) return RIH_ERR_IO;
    if (with_flag_byte) { bool fast; in.read((char*)&fast, sizeof(fast)); }
    rih_index* h = new rih_index();
    if (!rib::load(h->L, in)) { delete h; return RIH_ERR_FORMAT; }
    *out = h;
    return RIH_OK;
}

int rih_gen_text(int kind, uint64_t n, uint64_t p0, uint64_t p1, uint64_t seed, uint8_t* out) {
    if (!out) return RIH_ERR_ARG;
    try {
        std::vector<uint8_t> t;
        switch (kind) {
            case 0: t = rib::gen_dna_drift(n, p0, p1, seed); break;
            case 1: t = rib::gen_dna_indep(n, p0, (double)p1 * 1e-9, seed); break;
            // p1 = sigma + 1000 * (edit probability per version in permille; 0 = SURVEY 8d's 0.25)
            case 2: t = rib::gen_versioned_doc(n, p0, (unsigned)(p1 % 1000), p1 >= 1000 ? (double)(p1 / 1000) * 1e-3 : 0.25, seed); break;
            case 3: t = rib::gen_pangenome(n, p0, p1, seed); break;
            default: return RIH_ERR_ARG;
        }
        if (t.size() != n) return RIH_ERR_ARG;
        std::memcpy(out, t.data(), n);
        return RIH_OK;
    } catch (...) {
        return RIH_ERR_ARG;
    }
}

int rih_gen_patterns(const uint8_t* text, uint64_t text_len, uint64_t N, uint64_t m, uint64_t start_limit,
                     uint64_t seed, uint8_t* out) {
    if (!text || !out) return RIH_ERR_ARG;
    try {
        uint64_t off = 0;
        std::vector<uint8_t> f = rib::gen_patterns(text, text_len, N, m, start_limit, seed, "synthetic", &off);
        std::memcpy(out, f.data() + off, N * m);
        return RIH_OK;
    } catch (...) {
        return RIH_ERR_ARG;
    }
}

int rih_suffix_array(const uint8_t* text, uint64_t len, int64_t* sa_out) {
    if (!sa_out || (!text && len)) return RIH_ERR_ARG;
    std::vector<uint8_t> buf(len + 1);
    if (len) std::memcpy(buf.data(), text, len);
    buf[len] = 0;
    rib::suffix_array_with_sentinel<int64_t>(buf.data(), (int64_t)(len + 1), sa_out);
    return RIH_OK;
}

}  // extern "C"
