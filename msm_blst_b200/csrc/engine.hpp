// Host-side state of one MSM context (one per GPU / per point shard) and the per-group entry table.
// The context is the device-resident form of the globals of the reference driver (main_p1.cpp:41-50):
// FIX_POINTS_LIST, BUCKET_SET, BUCKET_VALUE_TO_ITS_INDEX, DIGIT_CONVERSION_HASH_TABLE and the two
// PRECOMPUTATION_POINTS_LISTs live in HBM; pointer arrays into host tables become table indices.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "params.hpp"

namespace msmb200 {

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct GroupOps;

// Static plan of the digit-splitting bucket reduction (list_sum_kernel in msm_kernels.cuh). Built once on the host
// from the bucket values of one window; identical for every window of a layout.
struct ListPlan {
    DevBuf start, idx;      // start[nlists + 1], idx[members]: members of list i are idx[start[i] .. start[i+1])
    uint32_t nlists = 0;
    uint32_t tl = 1;        // team per list: lanes (list_sum_kernel) or quads (list_sum_coop_kernel), a power of two
    size_t members = 0;
};
struct ReducePlan {
    bool valid = false;
    size_t key_nbw = 0;     // cache key: number of buckets per window and largest value (dense plans)
    uint32_t c_lo = 0;      // value = lo + 2^c_lo * hi
    uint32_t nbits_w = 0;   // bit positions per window after stage 2
    bool s1_coop = false;   // stage 1 by quads (list_sum_coop_kernel MODE 0) instead of one lane per slice
    // stage 1: slices of the (digit, value) lists over buckets, one lane each; 1b: slices -> digit lists;
    // 2a: (bit, slice) lists over the digit sums; 2b: slices -> bit lists
    ListPlan s1, s1b, s2a, s2b;
};

struct Ctx {
    int group = 0;
    msmb200_config cfg{};
    size_t n = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    const GroupOps *ops = nullptr;
    std::string err;

    // parameters
    std::vector<int> bucket_set;  // host copy of BUCKET_SET
    int q = 0;
    int pip_window = 0, pip_tiles = 0;
    int *d_bucket_vals = nullptr;     // BUCKET_SET
    int *d_v2i = nullptr;             // BUCKET_VALUE_TO_ITS_INDEX  [q/2+1]
    uint32_t *d_dtab = nullptr;       // packed DIGIT_CONVERSION_HASH_TABLE [q+1]
    int *d_chunk_first = nullptr;     // bucket-reduction chunk boundaries of BUCKET_SET (value-aligned chunks)
    uint32_t red_vspan = 0, red_nchunks = 0;

    // points and tables (affine, Montgomery)
    void *d_points = nullptr;      bool have_points = false;
    // the context's own precomputation tables: entries padded to whole 128-byte lines (stride 128 / 256 bytes) unless
    // MSMB200_PACKED_TABLES was set at creation; downloads and table files always use the reference's packed layout
    uint32_t table_stride = 0;
    void *d_table_ches = nullptr;  bool have_ches = false;
    void *d_table_bgmw = nullptr;  bool have_bgmw = false;

    // workspace (grow-only)
    DevBuf scalars, keys, vals, ranks, sorted, count, packed, scanned, tile_sums, seg_start, item_start, cursor,
        item_begin, item_cnt, order, len_hist, len_start, len_cursor, partial, chunk_a, chunk_b, result,
        flat, signs, pidx, heavy, light, medium, maxcount;
    // batch-affine accumulation (batch_affine.cuh): per-round totals / scans / slot descriptors, ping-pong point buffers, the
    // prefix-product scratch, one affine sum per bucket, identity index for the reducers
    DevBuf ba_totals, ba_tile_sums, ba_bases, ba_adesc, ba_cdesc, ba_heavy, pts_a, pts_b, ba_scratch, bucket_sum, iota, ba_counters, ba_sm_arrivals;
    uint32_t *h_totals = nullptr;   // pinned read-back of ba_totals
    size_t iota_n = 0;
    int ba_resident = 0;            // co-resident ba_round_kernel blocks per SM (occupancy query, cached)
    int vspan_env = 0;              // MSMB200_VSPAN at creation (0: automatic)
    bool no_overlap = false;        // MSMB200_NO_OVERLAP at creation: upload all scalars before the first kernel
    int ba_batch_max = 110, ba_batch_fixed = 0, ba_stagger = 1;  // slots per lane per round: upper bound / forced value (MSMB200_BA_BATCH at creation)
    int sms = 148;                  // cudaDevAttrMultiProcessorCount of the context's device
    int item_len_fixed = 0;         // XYZZ work-item length override (0 = automatic)
    int accum_env = 0, reduce_env = 0;  // MSMB200_ACCUM / MSMB200_REDUCE read once at context creation (0 = not set)
    ReducePlan plan_ches, plan_bgmw, plan_pip;  // digit-splitting reduction plans (sparse CHES set / dense windows)
    uint32_t plan_bgmw_windows = 0, plan_pip_windows = 0;
    DevBuf red_a, red_b, red_c, red_d;          // outputs of the list-sum stages
    int reduce_mode = 0;                        // 0 = default (digit splitting), 1 = chunked running sums
    std::vector<int> h_chunk_first;   // host copy of d_chunk_first
    int shard_rank = 0, shard_world = 1;  // bucket-range sharding: this context owns 1/world of the reduction chunks
    int accum_mode = 0;  // 0 = default, 1 = XYZZ work items, 2 = batch-affine rounds
    void *h_result = nullptr;  // pinned staging for the result
    void *bits_out = nullptr;  // device buffer for the per-bit XYZZ sums (msmb200_msm_bits_device) instead of a finalised point
    // host-to-host CHES calls: chunked upload on a second stream overlapped with the digit kernel (msmb200_msm)
    const void *h_scalars_pending = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_chunk[8] = {};

    // timing
    cudaEvent_t ev[8] = {};
    float last_ms[6] = {0, 0, 0, 0, 0, 0};
    int launches = 0;
    int last_accum = 0;             // accumulator the last MSM used: 1 XYZZ work items, 2 batch-affine rounds
};

struct ShimAux {
    const int *in = nullptr; int *out_b = nullptr; unsigned char *signs = nullptr; unsigned long long *ptrs = nullptr; size_t m = 0;
    const int *triples = nullptr; unsigned long long base = 0; unsigned entry_bytes = 0; unsigned long long entries = 0;
    uint32_t *pidx = nullptr, *bad = nullptr;
};

struct GroupOps {
    size_t aff_bytes, jac_bytes, xyzz_bytes;
    // method 1..4; scalars in device memory; writes Jacobian partial to d_out_jac (device) and/or affine to
    // ctx->h_result (host, synchronises the stream) when want_affine.
    int (*msm)(Ctx *, int method, const void *d_scalars, void *d_out_jac, bool want_affine);
    int (*generate_fix_points)(Ctx *, size_t first);
    int (*table_build)(Ctx *, int which);  // 0 CHES 3nh, 1 BGMW95
    int (*sum_partials)(Ctx *, const void *d_partials, int count);
    // multi-GPU combine over all-gathered per-bit sums (world x nwindows x nbits_w XYZZ points): sum, Horner, to_affine
    int (*combine_bits)(Ctx *, const void *d_gathered, int world, uint32_t nwindows, uint32_t nbits_w, uint32_t wbits);
    // generic tile: device arrays of bucket index (or value when v2i given) / sign / point index into d_table
    int (*tile)(Ctx *, const void *d_table, const int *d_bvals, const unsigned char *d_signs, const uint32_t *d_pidx,
                size_t m, const int *d_v2i, const int *d_bucket_vals, size_t nbuckets, int d_max, const int *d_chunk_first, uint32_t vspan,
                uint32_t nchunks, const ReducePlan *plan, void *d_out_jac);
    int (*pippenger)(Ctx *, const void *d_points, size_t npoints, const void *d_scalars, int nbits, void *d_out_jac,
                     bool want_affine, int wbits_table, int tile_bit0, int tile_window);
    // device halves of the literal shims on the caller's arrays: what 0 = construct_nh on {m, b, alpha} triples, 1 = host
    // pointers -> table indices (*bad counts pointers outside the table), 2 = max of an int array, 3 = min / max pointer
    int (*shim_aux)(Ctx *, int what, const ShimAux &x);
    int (*field_op)(int field, int op, const void *a, const void *b, void *out, size_t n);
    int (*point_op)(int op, const void *a, const void *b, const unsigned char *flags, void *out, size_t n);
    int (*digits)(Ctx *, int kind, const void *d_scalars, size_t n, uint32_t *d_keys, uint32_t *d_vals);
    // resident 128-thread blocks per SM of the list-sum kernels (0: list_sum_kernel stage 1, 1: list_sum_coop_kernel):
    // the reduction plan sizes its grids to whole waves; which == 2: lane groups per warp of the cooperative kernels
    int (*resident_blocks)(int which);
    // table[i * 2^(wbits-1) + k] = (k + 1) * P_i (blst_pNs_mult_wbits_precompute), device buffers
    int (*wbits_precompute)(Ctx *, const void *d_points, size_t npoints, int wbits, void *d_table);
    // table persistence: dir 0 = affine entries -> blst_pN_affine_serialize bytes; dir 1 = bytes (serialized != 0) or raw
    // Montgomery entries (serialized == 0) -> validated affine entries, *d_bad += number of invalid entries
    int (*table_io)(Ctx *, int dir, int serialized, const void *d_src, void *d_dst, size_t n, uint32_t *d_bad);
    // *d_out += position-weighted 64-bit checksum of `bytes` bytes of device memory (bytes a multiple of 8)
    int (*checksum)(Ctx *, const void *d_src, size_t bytes, unsigned long long *d_out);
};

int measure_peaks(double out[4]);
const GroupOps *group_ops_g1();
const GroupOps *group_ops_g2();

// chunk_first[c] = first index l >= 1 with values[l] > c * vspan (c = 0..nchunks), values ascending
std::vector<int> build_chunk_first(const int *values, size_t count, uint32_t vspan, uint32_t *nchunks_out);
uint32_t pick_vspan_host(size_t max_value, uint32_t nwindows);
// values: ascending bucket values of one window (values[0] = 0 unused) or nullptr for dense (value == index)
int build_reduce_plan(Ctx *c, ReducePlan &plan, const int *values, size_t nbw, uint32_t nwindows);
void free_reduce_plan(ReducePlan &plan);
int ctx_fail(Ctx *c, int code, const std::string &msg);
int ensure(Ctx *c, DevBuf &b, size_t bytes);

#define MSM_CUDA(c, call)                                                                              \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return ctx_fail((c), MSMB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));  \
    } while (0)

}  // namespace msmb200
