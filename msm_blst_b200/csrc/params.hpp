// Host-side parameter layer: the run-time form of the reference's compile-time configuration and the CHES
// bucket-set machinery. Replaces (reference, host C++):
//   ches_config_files/config_file_n_exp_*.h:5-17     -> kConfigs / find_config
//   auxiliaryfunc.h:234-288 omega2/omega3/construct_bucket_set   -> build_bucket_set (array sieve, O(q))
//   main_p1.cpp:134-152 BUCKET_VALUE_TO_ITS_INDEX + DIGIT_CONVERSION_HASH_TABLE  -> build_digit_table (packed u32)
//   src/multi_scalar.c:268-275 pippenger_window_size
// Everything here runs once at context creation; the outputs are uploaded to HBM.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/msm_b200.h"

namespace msmb200 {

struct NamedConfig { const char *name; msmb200_config c; };
static const NamedConfig kConfigs[] = {
    {"8", {8, 12, 22, 7, 6, 857, 10, 26}},           {"9", {9, 13, 20, 231, 6, 1725, 11, 24}},
    {"10", {10, 13, 20, 231, 6, 1725, 12, 22}},      {"11", {11, 14, 19, 7, 6, 3417, 13, 20}},
    {"12", {12, 14, 19, 7, 6, 3417, 13, 20}},        {"13", {13, 16, 16, 29677, 6, 18343, 15, 17}},
    {"14", {14, 16, 16, 29677, 6, 18343, 15, 17}},   {"15", {15, 16, 16, 29677, 6, 18343, 16, 16}},
    {"16", {16, 19, 14, 231, 6, 109244, 17, 15}},    {"16_beta", {16, 18, 15, 7, 6, 54618, 17, 15}},
    {"17", {17, 20, 13, 29677, 6, 220931, 17, 15}},  {"17_beta", {17, 19, 14, 231, 6, 109244, 17, 15}},
    {"18", {18, 20, 13, 29677, 6, 220931, 19, 14}},  {"19", {19, 20, 13, 29677, 6, 220931, 20, 13}},
    {"20", {20, 22, 12, 7419, 6, 874437, 20, 13}},   {"20_beta", {20, 20, 13, 29677, 6, 220931, 20, 13}},
    {"21", {21, 22, 12, 7419, 6, 874437, 22, 12}},
};

inline const msmb200_config *find_config(const char *name) {
    for (const NamedConfig &n : kConfigs)
        if (strcmp(n.name, name) == 0) return &n.c;
    return nullptr;
}

inline int omega23_parity(int i) {  // (omega2(i) + omega3(i)) & 1
    int c = 0;
    while ((i & 1) == 0) { i >>= 1; c++; }
    while (i % 3 == 0) { i /= 3; c++; }
    return c & 1;
}

// B.1 of SURVEY App. B; the erasure passes are sequential and data dependent, exactly as in the reference.
inline std::vector<int> build_bucket_set(int q, int ah) {
    std::vector<uint8_t> in(q / 2 + 2, 0);
    in[0] = in[1] = 1;
    for (int i = 2; i <= q / 2; ++i) in[i] = omega23_parity(i) == 0;
    for (int i = q / 4; i < q / 2; ++i)
        if (in[i] && q - 2 * i <= q / 2 && in[q - 2 * i]) in[q - 2 * i] = 0;
    for (int i = q / 6; i < q / 4; ++i)
        if (in[i] && q - 3 * i <= q / 2 && in[q - 3 * i]) in[q - 3 * i] = 0;
    for (int i = 1; i <= ah + 1 && i <= q / 2; ++i)
        if (omega23_parity(i) == 0) in[i] = 1;
    std::vector<int> B;
    for (int i = 0; i <= q / 2; ++i)
        if (in[i]) B.push_back(i);
    return B;
}

// Parameter-search side of the construction: check_bucket_set_validity + max_gap_in_bucket_set of the
// reference's tool (main_bucket_set_construction.cpp:74-113, :115-122) on bitmaps instead of std::set.
//   leading: every top digit 0..ah+1 is m*b with b in B, m in {1,2,3}            (:80-90)
//   cover:   every digit 0..q is m*b or q - m*b with m*b <= q                     (:92-110)
struct BucketSetCheck { bool leading_ok, cover_ok; long size, max_gap, first_uncovered; };
inline BucketSetCheck check_bucket_set(const std::vector<int> &B, long q, long ah) {
    BucketSetCheck r{true, true, (long)B.size(), 0, -1};
    std::vector<uint8_t> lead((size_t)ah + 2, 0), cov((size_t)q + 1, 0);
    lead[0] = 1;
    cov[0] = 1;
    for (size_t j = 0; j < B.size(); ++j) {
        const long b = B[j];
        if (j + 1 < B.size() && B[j + 1] - b > r.max_gap) r.max_gap = B[j + 1] - b;
        for (long m = 1; m <= 3; ++m) {
            const long v = m * b;
            if (v <= ah + 1) lead[(size_t)v] = 1;
            if (v <= q) { cov[(size_t)v] = 1; cov[(size_t)(q - v)] = 1; }
        }
    }
    for (long d = 0; d <= ah + 1; ++d) r.leading_ok = r.leading_ok && lead[(size_t)d];
    for (long d = 0; d <= q; ++d)
        if (!cov[(size_t)d]) { r.cover_ok = false; if (r.first_uncovered < 0) r.first_uncovered = d; }
    return r;
}

// Packed digit table: entry d in [0, q] -> bucket index (bits 0..21) | (m-1) << 22 | alpha << 24.
// Same precedence as the reference's two passes (later writes win: alpha = 0 preferred, then the largest m).
constexpr uint32_t DT_IDX_MASK = (1u << 22) - 1;
constexpr int DT_M_SHIFT = 22;
constexpr int DT_A_SHIFT = 24;
constexpr uint32_t DT_INVALID = 0xffffffffu;

inline std::vector<uint32_t> build_digit_table(int q, const std::vector<int> &B) {
    std::vector<uint32_t> T((size_t)q + 1, DT_INVALID);
    for (int alpha = 1; alpha >= 0; --alpha)
        for (int m = 1; m <= 3; ++m)
            for (size_t k = 0; k < B.size(); ++k) {
                long v = (long)m * B[k];
                if (v > q) break;  // B ascending
                size_t d = alpha ? (size_t)(q - v) : (size_t)v;
                T[d] = (uint32_t)k | (uint32_t)(m - 1) << DT_M_SHIFT | (uint32_t)alpha << DT_A_SHIFT;
            }
    return T;
}

inline size_t pippenger_window_size(size_t npoints) {
    size_t wbits;
    for (wbits = 0; npoints >>= 1; wbits++) ;
    return wbits > 12 ? wbits - 3 : (wbits > 4 ? wbits - 2 : (wbits ? 2 : 1));
}

}  // namespace msmb200
