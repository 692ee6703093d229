// projectultra_b200/csrc/ldpc_code.cpp — see ldpc_code.h.
#include "ldpc_code.h"

#include <algorithm>
#include <numeric>
#include <random>
#include <stdexcept>

namespace pu {

void ldpc_code_params(int rate, int* k, int* m) {
    static const struct { int rate, k; } tab[] = {{0, 162}, {2, 324}, {3, 432}, {4, 486}, {5, 540}};
    *k = 324;
    for (const auto& t : tab)
        if (t.rate == rate) *k = t.k;
    *m = kLdpcN - *k;
}

LdpcCode build_ldpc_code(int rate) {
    LdpcCode c;
    c.rate = rate;
    ldpc_code_params(rate, &c.k, &c.m);
    const int k = c.k, m = c.m;
    c.rows.assign(m, {});

    // The pseudo-random graph is grown one information bit at a time from a generator seeded with the rate
    // enum (ldpc_decoder.cpp:72).  std::mt19937 is bit-exact across standard libraries; the shuffle is the
    // reference's hand-rolled Fisher-Yates (`rng() % i`, :100-103), not std::shuffle.
    std::mt19937 gen(static_cast<uint32_t>(0x12345678 + rate));
    const int per_bit = std::min(std::max(3, (4 * m) / k), m / 2);   // target_var_degree, :82-84
    std::vector<int> load(m, 0);
    std::vector<int> open;
    open.reserve(m);
    for (int bit = 0; bit < k; ++bit) {
        open.clear();
        for (int chk = 0; chk < m; ++chk)
            if (load[chk] < kMaxInfoEdgesPerCheck) open.push_back(chk);
        for (size_t top = open.size(); top > 1; --top) std::swap(open[top - 1], open[gen() % top]);
        const int take = std::min<int>(per_bit, static_cast<int>(open.size()));
        for (int d = 0; d < take; ++d) {
            c.rows[open[d]].push_back(bit);
            ++load[open[d]];
        }
    }
    for (int chk = 0; chk < m; ++chk)                 // empty-row fix-up draws from the same stream, :115-121
        if (c.rows[chk].empty()) c.rows[chk].push_back(static_cast<int>(gen() % k));
    std::vector<int> vdeg(k, 0);
    for (int chk = 0; chk < m; ++chk) {
        for (int v : c.rows[chk]) ++vdeg[v];
        c.rows[chk].push_back(k + chk);               // identity part, :124-128
        c.n_edges += static_cast<int>(c.rows[chk].size());
    }
    c.max_var_degree = *std::max_element(vdeg.begin(), vdeg.end());
    return c;
}

std::vector<uint8_t> ldpc_encode(const LdpcCode& code, const uint8_t* data, size_t n_bytes) {
    const size_t k = code.k, m = code.m, n = k + m, total = n_bytes * 8;
    std::vector<uint8_t> out;
    std::vector<uint8_t> word(n);
    for (size_t start = 0; start < total; start += k) {     // k bits per block, zero padded (:214-219)
        for (size_t j = 0; j < k; ++j) {
            const size_t b = start + j;
            word[j] = b < total ? (data[b >> 3] >> (7 - (b & 7))) & 1 : 0;
        }
        for (size_t i = 0; i < m; ++i) {                    // parity_i = XOR of row i's info bits (:221-230)
            uint8_t acc = 0;
            const auto& row = code.rows[i];
            for (size_t e = 0; e + 1 < row.size(); ++e) acc ^= word[row[e]];
            word[k + i] = acc;
        }
        for (size_t base = 0; base < n; base += 8) {        // MSB-first, 81 bytes per block (:237-250)
            uint8_t byte = 0;
            for (size_t b = 0; b < 8; ++b) byte = static_cast<uint8_t>((byte << 1) | (base + b < n ? word[base + b] : 0));
            out.push_back(byte);
        }
    }
    return out;
}

LdpcHostTables make_ldpc_tables(const LdpcCode& code) {
    LdpcHostTables t;
    t.k = code.k;
    t.m = code.m;
    const int k = code.k, m = code.m;
    std::vector<int> slot_of(m), order(m);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return code.rows[a].size() > code.rows[b].size(); });
    t.cn_ninfo.resize(m);
    t.cn_check.resize(m);
    t.cn_var.assign(static_cast<size_t>(kMaxInfoEdgesPerCheck) * m, 0xFFFF);
    for (int p = 0; p < m; ++p) {
        const int chk = order[p];
        slot_of[chk] = p;
        const auto& row = code.rows[chk];
        const int ninfo = static_cast<int>(row.size()) - 1;
        if (ninfo > kMaxInfoEdgesPerCheck) throw std::logic_error("check degree exceeds table width");
        t.cn_ninfo[p] = static_cast<uint8_t>(ninfo);
        t.cn_check[p] = static_cast<uint16_t>(chk);
        for (int e = 0; e < ninfo; ++e) t.cn_var[static_cast<size_t>(e) * m + p] = static_cast<uint16_t>(row[e]);
    }
    // variable side: walk checks in ascending index so each list is already in the reference's summation order
    std::vector<std::vector<uint16_t>> per_var(k);
    for (int chk = 0; chk < m; ++chk) {
        const auto& row = code.rows[chk];
        for (size_t e = 0; e + 1 < row.size(); ++e)
            per_var[row[e]].push_back(static_cast<uint16_t>(e * m + slot_of[chk]));
    }
    t.dv_max = 0;
    for (const auto& v : per_var) t.dv_max = std::max<int>(t.dv_max, static_cast<int>(v.size()));
    t.vn_deg.resize(k);
    t.vn_slot.assign(static_cast<size_t>(std::max(t.dv_max, 1)) * k, 0xFFFF);
    for (int j = 0; j < k; ++j) {
        t.vn_deg[j] = static_cast<uint8_t>(per_var[j].size());
        for (size_t d = 0; d < per_var[j].size(); ++d) t.vn_slot[d * k + j] = per_var[j][d];
    }
    return t;
}

}  // namespace pu
