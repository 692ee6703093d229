// projectultra_b200/csrc/ldpc_code.h — host-side construction of the reference's LDPC code and of the
// device tables the decoder kernel walks.
#pragma once
#include <cstdint>
#include <vector>

namespace pu {

constexpr int kLdpcN = 648;
constexpr int kMaxInfoEdgesPerCheck = 6;  // max_check_degree, src/fec/ldpc_decoder.cpp:87

struct LdpcCode {
    int rate = 0, k = 0, m = 0, n_edges = 0;
    // rows[i] = variable indices of check i in the reference's stored order: info bits in construction
    // order, optional fix-up bit, then the identity column k+i (src/fec/ldpc_decoder.cpp:89-128)
    std::vector<std::vector<int>> rows;
    int max_var_degree = 0;
};

// CodeRate enum -> (k, m); unknown rates fall back to R1/2 sizes (getCodeParams, ldpc_decoder.cpp:21-36)
void ldpc_code_params(int rate, int* k, int* m);
// H = [H_data | I] exactly as LDPCDecoder::Impl::buildMatrix (ldpc_decoder.cpp:64-137)
LdpcCode build_ldpc_code(int rate);
// systematic encoder, LDPCEncoder::encode (src/fec/ldpc_encoder.cpp:193-257)
std::vector<uint8_t> ldpc_encode(const LdpcCode& code, const uint8_t* data, size_t n_bytes);

// Flat tables in the layout ldpc_decode.cu expects.  Checks are visited in "slot" order p (sorted by
// descending info degree so that a warp sees one loop bound); messages of slot p, edge e live at e*m + p.
struct LdpcHostTables {
    int k = 0, m = 0, dv_max = 0;
    std::vector<uint8_t> cn_ninfo;   // [m]      info edges of slot p
    std::vector<uint16_t> cn_check;  // [m]      original check index of slot p (parity variable = k + check)
    std::vector<uint16_t> cn_var;    // [6][m]   info variable of (edge e, slot p), 0xFFFF if absent
    std::vector<uint8_t> vn_deg;     // [k]      degree of info variable j
    std::vector<uint16_t> vn_slot;   // [dv_max][k] message slot of j's d-th edge in ASCENDING CHECK ORDER
                                     //          (the accumulation order of ldpc_decoder.cpp:208-213)
};
LdpcHostTables make_ldpc_tables(const LdpcCode& code);

}  // namespace pu
