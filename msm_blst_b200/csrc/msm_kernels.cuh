// CUDA kernels of the bucket-method MSM pipeline (sm_100a), shared by all four methods of the reference:
//
//   scalars --digits_*--> (bucket key, table index | sign) pairs + per-bucket histogram      [subsystem 2]
//           --scan / scatter / itemize / order--> per-bucket contiguous segments, work items sorted by length
//           --accumulate--> one XYZZ partial per work item (gathers 96 B / 192 B affine table entries)  [3]
//           --combine / reduce_chunks / sum_groups--> one XYZZ sum per window                           [4]
//           --finalize--> Horner over windows, Jacobian partial or canonical affine result               [5]
//
// Reference functions replaced are cited at each kernel. The field/curve math is in fp.cuh/fp2.cuh/ec.cuh.
#pragma once
#include <cstdint>
#include "ec.cuh"
#include "coop.cuh"
#include "batch_affine.cuh"

namespace msmb200 {

constexpr uint32_t KEY_SKIP = 0xffffffffu;
constexpr uint32_t DTAB_IDX_MASK = (1u << 22) - 1;

// ------------------------------------------------------------------------------------------------
// 256-bit scalar helpers (8 x u32 little-endian) — uint256_t of src_from_aztec replaced by plain limbs
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t scalar_bits(const uint32_t s[8], int off, int nbits) {
    // bits [off, off+nbits), nbits <= 25, zero beyond bit 255
    int w = off >> 5, sh = off & 31;
    uint64_t lo = w < 8 ? s[w] : 0u;
    uint64_t hi = (w + 1) < 8 ? s[w + 1] : 0u;
    uint64_t v = (lo | (hi << 32)) >> sh;
    return (uint32_t)v & ((1u << nbits) - 1u);
}
__device__ __forceinline__ void load_scalar(uint32_t s[8], const uint32_t *scalars, size_t i) {
    const uint4 *p = reinterpret_cast<const uint4 *>(scalars + 8 * i);
    uint4 a = p[0], b = p[1];
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
}

// ------------------------------------------------------------------------------------------------
// [2] digit decomposition
// ------------------------------------------------------------------------------------------------

// CHES "MB radix-q" digits. Replaces trans_uint256_t_to_MB_radixq_expr (auxiliaryfunc.h:92-118) and the
// per-digit bookkeeping of pippenger_variant_q_over_5_CHES (main_p1.cpp:208-228): one thread per scalar,
// h table lookups with the alpha carry; emits the bucket INDEX (BUCKET_VALUE_TO_ITS_INDEX already folded
// into the packed table) and table entry 3(i*h+j)+m-1 with the sign in bit 31; entries whose bucket value
// is 0 are skipped like `if(booth_idx)` in src/multi_scalar.c:445,:457.
static __global__ void digits_ches_kernel(const uint32_t *__restrict__ scalars, size_t n, int h, int e,
                                   const uint32_t *__restrict__ dtab, uint32_t *__restrict__ keys,
                                   uint32_t *__restrict__ vals, uint32_t *__restrict__ count, uint32_t *__restrict__ ranks, int digit_major,
                                   uint32_t lo, uint32_t hi, size_t i0, size_t cnt) {
    // scalars [i0, i0 + cnt) of the n: the host-to-host entry point launches one slice per uploaded chunk so that the
    // digit decomposition overlaps the PCIe copy of the next chunk
    // ranks[at] = position of the entry inside its bucket (the value returned by the histogram atomic), so the
    // scatter pass needs no second round of atomics; may be null (parity-test hook).
    // [lo, hi): bucket-index range owned by this context (bucket-range sharding over GPUs); others are skipped
    size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i0 + cnt || i >= n) return;
    uint32_t s[8];
    load_scalar(s, scalars, i);
    uint32_t carry = 0;
    for (int j = 0; j < h; j++) {
        uint32_t d = scalar_bits(s, e * j, e) + carry;
        uint32_t ent = dtab[d];
        uint32_t idx = ent & DTAB_IDX_MASK;
        uint32_t m1 = (ent >> 22) & 3u;
        uint32_t alpha = (ent >> 24) & 1u;
        carry = alpha;
        size_t slot = i * h + j;
        uint32_t key = KEY_SKIP, rank = 0;
        if (idx != 0 && idx >= lo && idx < hi) {
            key = idx;
            rank = atomicAdd(&count[idx], 1u);
        }
        // digit-major placement (j*n + i) makes the warp's stores contiguous; the table index keeps the reference's i*h + j
        const size_t at = digit_major ? (size_t)j * n + i : slot;
        keys[at] = key;
        if (ranks) ranks[at] = rank;
        vals[at] = (uint32_t)(3 * slot + m1) | (alpha << 31);
    }
}

// Integral-scalar-conversion front end, step (a): standard q-ary digits into one flat int array
// (trans_uint256_t_to_standard_q_ary_expr, auxiliaryfunc.h:83-90; main_p1.cpp:259-266).
static __global__ void digits_std_kernel(const uint32_t *__restrict__ scalars, size_t n, int h, int e, int *__restrict__ flat) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    load_scalar(s, scalars, i);
    for (int j = 0; j < h; j++) flat[i * h + j] = (int)scalar_bits(s, e * j, e);
}
// step (b): in-place conversion of the flat digit array, replaces blst_p1_construct_nh_scalars_nh_points
// (src/multi_scalar.c:748-775): slot k becomes the bucket VALUE b, booth_signs[k] = alpha, the point
// "pointer" becomes table index 3k+m-1, and alpha carries into slot k+1. The top digit of a scalar never
// carries (bucket-set step 4), so scalars are independent and one thread walks the h slots of one scalar.
static __global__ void construct_nh_kernel(int *__restrict__ flat, unsigned char *__restrict__ signs,
                                    uint32_t *__restrict__ pidx, size_t n, int h,
                                    const uint32_t *__restrict__ dtab, const int *__restrict__ bucket_vals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int carry = 0;
    for (int j = 0; j < h; j++) {
        size_t k = i * h + j;
        uint32_t ent = dtab[flat[k] + carry];
        uint32_t idx = ent & DTAB_IDX_MASK;
        carry = (ent >> 24) & 1u;
        flat[k] = bucket_vals[idx];
        signs[k] = (unsigned char)carry;
        pidx[k] = (uint32_t)(3 * k + ((ent >> 22) & 3u));
    }
}
// Literal blst_pN_construct_nh_scalars_nh_points (src/multi_scalar.c:748-775) on the caller's own arrays: the digit table is
// the reference's array of {m, b, alpha} triples (bindings/blst.h:253); slot k becomes the bucket VALUE b, booth_signs[k] =
// alpha, alpha carries into slot k + 1, and the output "pointer" is the HOST address table_base + (3k + m - 1) * entry size,
// exactly what the reference stores. The reference is one sequential pass over all slots (it does not know where a scalar
// ends); here every slot finds its own carry-in: it walks back to the nearest slot whose alpha does not depend on ITS
// carry-in (alpha(H[d]) == alpha(H[d + 1]) — every scalar's small top digit is one) and replays forward from there.
__device__ __forceinline__ int tri_alpha(const int *__restrict__ triples, int d) { return triples[3 * (size_t)d + 2]; }
static __global__ void construct_nh_triples_kernel(const int *__restrict__ in, int *__restrict__ out_b, unsigned char *__restrict__ signs,
                                                   unsigned long long *__restrict__ ptrs, size_t m, const int *__restrict__ triples,
                                                   unsigned long long table_base, unsigned entry_bytes) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    long long j = (long long)k - 1;
    int carry = 0;
    while (j >= 0) {
        const int d = in[j], a0 = tri_alpha(triples, d), a1 = tri_alpha(triples, d + 1);
        if (a0 == a1) { carry = a0; break; }
        j--;
    }
    for (long long t = j + 1; t < (long long)k; t++) carry = tri_alpha(triples, in[t] + carry);
    const int d = in[k] + carry;
    const int mm = triples[3 * (size_t)d], b = triples[3 * (size_t)d + 1], alpha = triples[3 * (size_t)d + 2];
    out_b[k] = b;
    signs[k] = (unsigned char)alpha;
    ptrs[k] = table_base + (unsigned long long)(3 * k + (size_t)mm - 1) * entry_bytes;
}
// largest value of an int array (sizes the upload of the caller's digit table) / lowest and highest host pointer
static __global__ void max_int_kernel(const int *__restrict__ v, size_t m, int *__restrict__ out) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int x = k < m ? v[k] : 0;
    x = __reduce_max_sync(0xffffffffu, x);
    if ((threadIdx.x & 31) == 0) atomicMax(out, x);
}
static __global__ void minmax_ptr_kernel(const unsigned long long *__restrict__ p, size_t m, unsigned long long *__restrict__ out /* [min, max] */) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    atomicMin(&out[0], p[k]);
    atomicMax(&out[1], p[k]);
}
// host pointers into a registered table -> table indices; anything outside the table raises *bad
static __global__ void ptrs_to_index_kernel(const unsigned long long *__restrict__ ptrs, size_t m, unsigned long long base, unsigned entry_bytes,
                                            unsigned long long entries, uint32_t *__restrict__ pidx, uint32_t *__restrict__ bad) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const unsigned long long off = ptrs[k] - base, idx = off / entry_bytes;
    if (ptrs[k] < base || idx >= entries || off % entry_bytes) { atomicAdd(bad, 1u); pidx[k] = 0; return; }
    pidx[k] = (uint32_t)idx;
}
// Front half of blst_p1_tile_pippenger_d_CHES (src/multi_scalar.c:437-461): booth_idx =
// bucket_value_to_its_index[scalars[k]], skip when 0. Also serves the literal blst-named shim.
static __global__ void tile_lookup_kernel(const int *__restrict__ bvals, const unsigned char *__restrict__ signs,
                                   const uint32_t *__restrict__ pidx, size_t m, const int *__restrict__ v2i,
                                   uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ count,
                                   uint32_t *__restrict__ ranks, uint32_t lo, uint32_t hi) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    int idx = v2i ? v2i[bvals[k]] : bvals[k];
    uint32_t key = KEY_SKIP, rank = 0;
    if (idx != 0 && (uint32_t)idx >= lo && (uint32_t)idx < hi) {
        key = (uint32_t)idx;
        rank = atomicAdd(&count[idx], 1u);
    }
    keys[k] = key;
    ranks[k] = rank;
    vals[k] = pidx[k] | ((uint32_t)(signs[k] != 0) << 31);
}

// BGMW95 signed radix-q' digits. Replaces trans_uint256_t_to_qhalf_expr (auxiliaryfunc.h:130-145) and the
// front end of pippenger_variant_BGMW95 (main_p1.cpp:311-375) including the r - a switch for the
// configurations with e'*h' == 255 (`trick`). r = group order (auxiliaryfunc.h:5-7).
static __global__ void digits_bgmw_kernel(const uint32_t *__restrict__ scalars, size_t n, int h, int e, int trick,
                                   uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ count,
                                   uint32_t *__restrict__ ranks, int digit_major, uint32_t lo, uint32_t hi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    load_scalar(s, scalars, i);
    uint32_t flip = 0;
    if (trick) {
        // data[3] > 2^62  (64-bit top limb = s[7]:s[6])
        bool cond = (s[7] > 0x40000000u) || (s[7] == 0x40000000u && s[6] != 0u);
        if (cond) {
            const uint32_t r[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
            uint32_t borrow = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint64_t t = (uint64_t)r[k] - s[k] - borrow;
                s[k] = (uint32_t)t;
                borrow = (uint32_t)(t >> 32) & 1u;
            }
            flip = 1;
        }
    }
    const int q = 1 << e, qhalf = q >> 1;
    int carry = 0;
    for (int j = 0; j < h; j++) {
        int d = (int)scalar_bits(s, e * j, e) + carry;
        carry = 0;
        if (j < h - 1 && d > qhalf) { d -= q; carry = 1; }
        uint32_t sign = d < 0 ? 1u : 0u;
        int mag = d < 0 ? -d : d;
        if (mag > qhalf) mag = qhalf;  // unreachable for scalars < r except with prob 2^-62; clamp
        size_t slot = i * h + j;
        uint32_t key = KEY_SKIP, rank = 0;
        if (mag != 0 && (uint32_t)mag >= lo && (uint32_t)mag < hi) {
            key = (uint32_t)mag;
            rank = atomicAdd(&count[mag], 1u);
        }
        const size_t at = digit_major ? (size_t)j * n + i : slot;
        keys[at] = key;
        if (ranks) ranks[at] = rank;
        vals[at] = (uint32_t)slot | ((sign ^ flip) << 31);
    }
}

// blst Pippenger signed windows. Replaces get_wval_limb + booth_encode (src/ec_mult.h:23-56) as driven by
// POINTonE1s_mult_pippenger / s_tile_pippenger (src/multi_scalar.c:549-576,:383-419): tile t covers bits
// [t*w, t*w + wb) plus the bit below it; the top tile has wb = nbits % w (possibly 0) and is unsigned.
// key = t * (2^(w-1) + 1) + |digit|; all tiles are emitted at once (ntiles entries per scalar).
static __global__ void digits_booth_kernel(const uint32_t *__restrict__ scalars, size_t n, int nbits, int w, int ntiles,
                                    uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ count,
                                    uint32_t *__restrict__ ranks, int digit_major, uint32_t lo, uint32_t hi, uint32_t table_nwin,
                                    int tile_bit0) {
    // tile_bit0 >= 0: ONE tile starting at that bit (blst_p1s_tile_pippenger, src/multi_scalar.c:383-419,:587-600;
    // ntiles must be 1): the tile is the top one when bit0 + w > nbits, exactly like the export's wrapper.
    // table_nwin != 0: blst_p1s_mult_wbits mode (src/multi_scalar.c:176-261, same window recoding): the digit selects
    // row entry |d| - 1 of point i's precomputed row (table_nwin = 2^(w-1) multiples per point) and every window has
    // ONE bucket, so key = 2t + 1 and val = i * table_nwin + |d| - 1.
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    load_scalar(s, scalars, i);
    const uint32_t nbw = table_nwin ? 2u : (1u << (w - 1)) + 1u;
    for (int t = 0; t < ntiles; t++) {
        const int bit0 = tile_bit0 >= 0 ? tile_bit0 : t * w;
        const bool top = tile_bit0 >= 0 ? (bit0 + w > nbits) : (t == ntiles - 1);
        int wb = top ? (nbits - bit0) : w;  // bits in this tile
        int cbits = top ? wb + 1 : w;
        uint32_t wval;
        if (bit0 == 0) wval = (scalar_bits(s, 0, wb) << 1);
        else wval = scalar_bits(s, bit0 - 1, wb + 1);
        uint32_t sign = (wval >> cbits) & 1u;
        int d = (int)((wval + 1u) >> 1) - (sign ? (1 << cbits) : 0);
        int mag = d < 0 ? -d : d;
        size_t slot = i * ntiles + t;
        uint32_t key = KEY_SKIP, rank = 0;
        if (mag != 0 && (uint32_t)mag >= lo && (uint32_t)mag < hi) {
            key = (uint32_t)t * nbw + (table_nwin ? 1u : (uint32_t)mag);
            rank = atomicAdd(&count[key], 1u);
        }
        const size_t at = digit_major ? (size_t)t * n + i : slot;
        keys[at] = key;
        if (ranks) ranks[at] = rank;
        vals[at] = (table_nwin ? (uint32_t)(i * table_nwin + (size_t)(mag > 0 ? mag - 1 : 0)) : (uint32_t)i) | (sign << 31);
    }
}

// ------------------------------------------------------------------------------------------------
// sort by bucket: histogram (fused above) -> exclusive scan -> scatter; work items; order by length
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// packs (count, nitems) into one u64 so one scan yields segment starts and item starts
// Also records the largest bucket.
static __global__ void prep_counts_kernel(const uint32_t *__restrict__ count, uint64_t *__restrict__ packed, size_t nb, uint32_t item_len,
                                          uint32_t *__restrict__ max_count) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = b < nb ? count[b] : 0;
    uint32_t wmax = __reduce_max_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && wmax) atomicMax(max_count, wmax);
    if (b >= nb) return;
    uint32_t items = (c + item_len - 1) / item_len;
    packed[b] = (uint64_t)c | ((uint64_t)items << 32);
}
__device__ __forceinline__ uint64_t block_exclusive_scan_u64(uint64_t v, uint64_t *total) {
    __shared__ uint64_t warp_sums[SCAN_THREADS / 32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint64_t s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    uint64_t base = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return base + x - v;
}
static __global__ void scan_tiles_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, uint64_t *__restrict__ tile_sums, size_t n) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : 0; sum += v[k]; }
    uint64_t total;
    uint64_t ex = block_exclusive_scan_u64(sum, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// single block: exclusive scan of the tile sums in place; grand total to tile_sums[ntiles]
static __global__ void scan_sums_kernel(uint64_t *tile_sums, size_t ntiles) {
    uint64_t carry = 0;
    for (size_t base = 0; base < ntiles; base += SCAN_THREADS) {
        size_t i = base + threadIdx.x;
        uint64_t v = i < ntiles ? tile_sums[i] : 0;
        uint64_t total;
        uint64_t ex = block_exclusive_scan_u64(v, &total);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = carry;
}
// adds tile offsets and splits the packed scan into seg_start / item_start; also clears the cursors
static __global__ void scan_finish_kernel(const uint64_t *__restrict__ scanned, const uint64_t *__restrict__ tile_sums,
                                   uint32_t *__restrict__ seg_start, uint32_t *__restrict__ item_start,
                                   uint32_t *__restrict__ cursor, size_t nb) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint64_t v = scanned[b] + tile_sums[b / SCAN_TILE];
    seg_start[b] = (uint32_t)v;
    item_start[b] = (uint32_t)(v >> 32);
    cursor[b] = 0;
}
static __global__ void scatter_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, size_t m,
                               const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ ranks,
                               uint32_t *__restrict__ sorted) {
    // counting-sort scatter; the position inside the bucket was fixed by the histogram atomic (ranks), so no atomics here
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    uint32_t key = keys[k];
    if (key == KEY_SKIP) return;
    uint32_t pos = seg_start[key] + ranks[k];
    sorted[pos] = vals[k];
}
// One work item = up to item_len consecutive entries of one bucket (a bucket with a huge count is split so
// that no thread serialises more than item_len additions). Also builds the histogram of item lengths.
static __global__ void itemize_kernel(const uint32_t *__restrict__ count, const uint32_t *__restrict__ seg_start,
                               const uint32_t *__restrict__ item_start, size_t nb, uint32_t item_len,
                               uint32_t *__restrict__ item_begin, uint32_t *__restrict__ item_cnt,
                               uint32_t *__restrict__ len_hist, uint32_t *__restrict__ heavy /* [0] = count, then bucket ids */,
                               uint32_t heavy_items, uint32_t *__restrict__ light /* [0] = count, then bucket ids */,
                               uint32_t medium_items, uint32_t *__restrict__ medium /* same */) {
    // buckets split into 2..heavy_items work items are folded by ONE lane group each, heavy_items+1..medium_items by one
    // WARP of lane groups each (combine_light_kernel), more by a block each (combine_heavy_kernel)
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t c = count[b];
    if (c == 0) return;
    if ((uint64_t)c > (uint64_t)item_len * medium_items) heavy[1 + atomicAdd(&heavy[0], 1u)] = (uint32_t)b;
    else if (c > item_len * heavy_items) medium[1 + atomicAdd(&medium[0], 1u)] = (uint32_t)b;
    else if (c > item_len) light[1 + atomicAdd(&light[0], 1u)] = (uint32_t)b;
    uint32_t s = seg_start[b], it = item_start[b];
    for (uint32_t off = 0; off < c; off += item_len, it++) {
        uint32_t len = min(item_len, c - off);
        item_begin[it] = s + off;
        item_cnt[it] = len;
        // warp-aggregated histogram update
        uint32_t peers = __match_any_sync(__activemask(), len);
        if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&len_hist[len], (uint32_t)__popc(peers));
    }
}
// single block: len_start[len] = number of items strictly longer than len (descending order => longest first)
static __global__ void len_scan_kernel(const uint32_t *__restrict__ len_hist, uint32_t *__restrict__ len_start,
                                uint32_t *__restrict__ len_cursor, uint32_t item_len) {
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int l = (int)item_len; l >= 0; l--) { len_start[l] = acc; acc += len_hist[l]; len_cursor[l] = 0; }
    }
}
static __global__ void order_items_kernel(const uint32_t *__restrict__ item_cnt, const uint64_t *__restrict__ totals,
                                   const uint32_t *__restrict__ len_start, uint32_t *__restrict__ len_cursor,
                                   uint32_t *__restrict__ order) {
    size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= (size_t)(*totals >> 32)) return;
    uint32_t len = item_cnt[it];
    uint32_t peers = __match_any_sync(__activemask(), len);
    int lane = threadIdx.x & 31;
    int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&len_cursor[len], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    order[len_start[len] + base + rank] = (uint32_t)it;
}

// ------------------------------------------------------------------------------------------------
// [3] bucket accumulation (XYZZ path): replaces the hot loop of POINTonE1_tile_pippenger_d_CHES /
// _BGMW95 / s_tile_pippenger (src/multi_scalar.c:437-461,:522-544,:397-417 -> xyzz_dadd_affine).
// One thread per work item; items sorted longest-first so the warps of a block have equal trip counts.
// Each table entry (96 B / 192 B) is fetched with 16-byte vector loads. Entries lie `stride16` 16-byte units apart: 6 / 12
// in the reference's packed layout (caller-owned tables, the fixed points), 8 / 16 in the context's own precomputation
// tables, which are padded to whole 128-byte lines (the memory system fetches lines: a packed 96-byte entry straddles two
// of them half of the time, profiles/r2_gather_microbench.json).
// ------------------------------------------------------------------------------------------------
template <class F> __device__ __forceinline__ const aff_t<F> *table_entry(const aff_t<F> *table, size_t idx, uint32_t stride16) {
    return reinterpret_cast<const aff_t<F> *>(reinterpret_cast<const uint4 *>(table) + idx * stride16);
}
template <class F> __device__ __forceinline__ aff_t<F> *table_entry(aff_t<F> *table, size_t idx, uint32_t stride16) {
    return reinterpret_cast<aff_t<F> *>(reinterpret_cast<uint4 *>(table) + idx * stride16);
}
template <class F>
__device__ __forceinline__ void load_affine(aff_t<F> &p, const aff_t<F> *__restrict__ table, uint32_t idx, uint32_t stride16) {
    const uint4 *src = reinterpret_cast<const uint4 *>(table) + (size_t)idx * stride16;
    uint4 *dst = reinterpret_cast<uint4 *>(&p);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(aff_t<F>) / 16); k++) dst[k] = __ldg(src + k);
}
template <class F>
static __global__ void __launch_bounds__(128) accumulate_kernel(const aff_t<F> *__restrict__ table, uint32_t stride16, const uint32_t *__restrict__ sorted,
                                                         const uint32_t *__restrict__ item_begin, const uint32_t *__restrict__ item_cnt,
                                                         const uint32_t *__restrict__ order, const uint64_t *__restrict__ totals,
                                                         xyzz_t<F> *__restrict__ partial) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)(*totals >> 32)) return;  // number of work items (high half of the packed scan total)
    uint32_t it = order[t];
    uint32_t beg = item_begin[it], cnt = item_cnt[it];
    xyzz_t<F> acc;
    xyzz_set_inf(acc);
#pragma unroll 1
    for (uint32_t k = 0; k < cnt; k++) {
        uint32_t v = sorted[beg + k];
        aff_t<F> p;
        load_affine(p, table, v & 0x7fffffffu, stride16);
        xyzz_add_affine(acc, p, (v >> 31) != 0);
    }
    partial[it] = acc;
}
template <class F> __device__ __noinline__ void xyzz_add_cold(xyzz_t<F> &acc, const xyzz_t<F> &q);
template <class F> __device__ __forceinline__ void load_xyzz(xyzz_t<F> &p, const xyzz_t<F> *src_);
// Buckets that were split into several work items ("heavy": more than item_len entries, e.g. the top window of
// blst's Pippenger or adversarially equal scalars): one BLOCK per heavy bucket folds its partials into the first
// one with a strided pass and a shared-memory tree, so the dependent chain is items/blockDim + log2(blockDim).
template <class F>
static __global__ void __launch_bounds__(128) combine_heavy_kernel(const uint32_t *__restrict__ count, const uint32_t *__restrict__ item_start,
                                                                   const uint32_t *__restrict__ heavy, uint32_t item_len,
                                                                   xyzz_t<F> *partial) {
    // tree in place in global memory (no shared memory: a different carve-out than the neighbouring kernels costs
    // an SM reconfiguration per launch, measured ~0.3 ms even when there is nothing to combine)
    const uint32_t nheavy = heavy[0];
    for (uint32_t hidx = blockIdx.x; hidx < nheavy; hidx += gridDim.x) {
        const uint32_t b = heavy[1 + hidx], t = threadIdx.x;
        const uint32_t c = count[b], items = (c + item_len - 1) / item_len, it = item_start[b];
        // strided pass: thread t folds items t, t+128, ... into slot t
        if (t < items) {
            xyzz_t<F> acc;
            load_xyzz(acc, partial + it + t);
            for (uint32_t k = t + blockDim.x; k < items; k += blockDim.x) {
                xyzz_t<F> q;
                load_xyzz(q, partial + it + k);
                xyzz_add_cold(acc, q);
            }
            partial[it + t] = acc;
        }
        __threadfence_block();
        __syncthreads();
        const uint32_t live = min(items, (uint32_t)blockDim.x);
        for (uint32_t s = 64; s > 0; s >>= 1) {
            if (t < s && t + s < live) {
                xyzz_t<F> a, q;
                load_xyzz(a, partial + it + t);
                load_xyzz(q, partial + it + t + s);
                xyzz_add_cold(a, q);
                partial[it + t] = a;
            }
            __threadfence_block();
            __syncthreads();
        }
    }
}

// [3'] bucket accumulation by batch-affine pairwise rounds (default): batch_affine.cuh

// ------------------------------------------------------------------------------------------------
// [4] bucket reduction. Replaces POINTonE1_integrate_buckets_accumulation_d_CHES (src/multi_scalar.c:301-321)
// and POINTonE1_integrate_buckets (:281-297) with the chunked form of SURVEY App. B.6: window `w` has nbw
// buckets (local index 0 unused) with ascending values val(l) (bucket_vals[l] for CHES, l itself when
// bucket_vals == nullptr); chunk c covers locals [1 + c*chunk, 1 + (c+1)*chunk) and produces
//     sum_{l in chunk} val(l) * S_l = W + base * T,   base = val(first-1),
// W by the reference's running sums over gaps (<= d_max, kept in tmp_d[]), base*T by double-and-add.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_GAP = 8;

// out-of-line copies for the cold paths (keeps the instruction footprint of the hot loops small)
template <class F> __device__ __noinline__ void xyzz_add_cold(xyzz_t<F> &acc, const xyzz_t<F> &q) { xyzz_add(acc, q); }

template <class F>
__device__ __forceinline__ void load_xyzz(xyzz_t<F> &p, const xyzz_t<F> *src_) {
    const uint4 *src = reinterpret_cast<const uint4 *>(src_);
    uint4 *dst = reinterpret_cast<uint4 *>(&p);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(xyzz_t<F>) / 16); k++) dst[k] = src[k];
}
template <class F, bool DENSE, bool AFFINE_IN>
static __global__ void __launch_bounds__(64) reduce_chunks_kernel(const void *__restrict__ bucket_points, const uint32_t *__restrict__ count,
                                                                   const uint32_t *__restrict__ item_start, const int *__restrict__ bucket_vals,
                                                                   const int *__restrict__ chunk_first, uint32_t nbw, uint32_t nwindows,
                                                                   uint32_t vspan, uint32_t chunks_per_window, uint32_t chunk_lo, int d_max,
                                                                   xyzz_t<F> *__restrict__ out) {
    // chunk c of window w covers the buckets whose VALUE lies in (c*vspan, (c+1)*vspan]  (vspan a power of two).
    //   DENSE : value == local index           -> locals [c*vspan + 1, (c+1)*vspan]
    //   sparse: values = bucket_vals[] (CHES)  -> locals [chunk_first[c], chunk_first[c+1])
    // out rows per window: row 0 = W_c = sum (val - c*vspan) * S, row 1 = c * T_c  (T_c = sum S); the caller adds
    // sum(row0) + vspan * sum(row1).
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nwindows * chunks_per_window) return;
    // this context owns chunks [chunk_lo, chunk_lo + chunks_per_window) of every window (all of them unless bucket-range sharded)
    uint32_t w = t / chunks_per_window, cl = t % chunks_per_window, c = chunk_lo + cl;
    uint32_t lo, hi;
    if (DENSE) { lo = 1 + c * vspan; hi = min(nbw, lo + vspan); }
    else { lo = (uint32_t)chunk_first[c]; hi = (uint32_t)chunk_first[c + 1]; }
    const int base = (int)(c * vspan);
    xyzz_t<F> tmp, W;
    xyzz_set_inf(tmp);
    xyzz_set_inf(W);
    xyzz_t<F> tmp_d[DENSE ? 1 : MAX_GAP + 1];
    if (!DENSE) {
        for (int g = 0; g <= d_max; g++) xyzz_set_inf(tmp_d[g]);
    }
    if (lo < hi) {
#pragma unroll 1
        for (uint32_t l = hi; l-- > lo;) {
            size_t b = (size_t)w * nbw + l;
            if (count[b] != 0) {
                // bucket sum: XYZZ partial of the bucket's first work item, or (batch-affine path) ONE affine point
                if (AFFINE_IN) {
                    aff_t<F> a;
                    load_affine(a, (const aff_t<F> *)bucket_points, item_start[b], (uint32_t)(sizeof(aff_t<F>) / 16));
                    xyzz_add_affine(tmp, a, false);
                } else {
                    xyzz_t<F> s;
                    load_xyzz(s, (const xyzz_t<F> *)bucket_points + item_start[b]);
                    xyzz_add(tmp, s);
                }
            }
            if (DENSE) {
                xyzz_add_cold(W, tmp);
            } else {
                int gap = bucket_vals[l] - (l > lo ? bucket_vals[l - 1] : base);
                xyzz_add_cold(tmp_d[gap], tmp);
            }
        }
        if (!DENSE) {
            xyzz_t<F> acc;
            xyzz_set_inf(acc);
#pragma unroll 1
            for (int g = d_max; g > 0; g--) {
                xyzz_add_cold(acc, tmp_d[g]);
                xyzz_add_cold(W, acc);
            }
        }
    }
    // row 1: c * T_c by MSB-first double-and-add (c < 2^17)
    xyzz_t<F> r;
    xyzz_set_inf(r);
    if (c != 0 && !xyzz_is_inf(tmp)) {
        int top = 31 - __clz((int)c);
#pragma unroll 1
        for (int bit = top; bit >= 0; bit--) {
            if (!xyzz_is_inf(r)) { xyzz_t<F> a = r; xyzz_double(r, a); }
            if ((c >> bit) & 1) xyzz_add_cold(r, tmp);
        }
    }
    size_t o = (size_t)w * 2 * chunks_per_window + cl;
    out[o] = W;
    out[o + chunks_per_window] = r;
}
// out[w*groups + g] = sum of in[w*per_window + g*r .. +r)
template <class F>
static __global__ void __launch_bounds__(128) sum_groups_kernel(const xyzz_t<F> *__restrict__ in, uint32_t per_window, uint32_t nwindows,
                                                         uint32_t r, uint32_t groups, xyzz_t<F> *__restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nwindows * groups) return;
    uint32_t w = t / groups, g = t % groups;
    uint32_t lo = g * r, hi = min(per_window, lo + r);
    xyzz_t<F> acc;
    xyzz_set_inf(acc);
    for (uint32_t k = lo; k < hi; k++) {
        xyzz_t<F> q;
        load_xyzz(q, in + (size_t)w * per_window + k);
        xyzz_add(acc, q);
    }
    out[t] = acc;
}

// last levels of the tree in ONE launch: block `row` sums in[row*per .. +per) (per <= a few thousand) with a
// shared-memory tree: log2(blockDim) dependent additions instead of one kernel launch per level
template <class F>
static __global__ void __launch_bounds__(256) tree_tail_kernel(xyzz_t<F> *in, uint32_t per, xyzz_t<F> *__restrict__ out) {
    // block `row` sums in[row*per .. +per) in place (global memory tree, see combine_heavy_kernel for why not shared memory)
    const uint32_t row = blockIdx.x, t = threadIdx.x;
    xyzz_t<F> *base = in + (size_t)row * per;
    if (t < per) {
        xyzz_t<F> acc;
        load_xyzz(acc, base + t);
        for (uint32_t k = t + blockDim.x; k < per; k += blockDim.x) {
            xyzz_t<F> q;
            load_xyzz(q, base + k);
            xyzz_add_cold(acc, q);
        }
        base[t] = acc;
    }
    __threadfence_block();
    __syncthreads();
    const uint32_t live = min(per, (uint32_t)blockDim.x);
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (t < s && t + s < live) {
            xyzz_t<F> a, q;
            load_xyzz(a, base + t);
            load_xyzz(q, base + t + s);
            xyzz_add_cold(a, q);
            base[t] = a;
        }
        __threadfence_block();
        __syncthreads();
    }
    if (t == 0) out[row] = base[0];
}

// ------------------------------------------------------------------------------------------------
// [4'] bucket reduction by DIGIT SPLITTING (default). Computes the same sum as the reference's
// POINTonE1_integrate_buckets_accumulation_d_CHES (src/multi_scalar.c:301-321) / integrate_buckets (:281-297),
// i.e. sum_l val(l) * S_l per window, but without any long running-sum chain:
//   stage 1   val = lo + 2^c_lo * hi: every bucket sum is added to the list of its lo digit and to the list of
//             its hi digit (2 additions per bucket, the reference's count) -> two dense arrays of <= 2^11 sums;
//   stage 2   every dense entry v is added to the list of each set bit of v (sliced, then the slices summed);
//   final     Horner over the bit positions (bits_finalize_coop_kernel).
// Every stage is a LIST SUM: a static plan (built once on the host from the bucket values, ReducePlan in engine.hpp)
// names the members of every list. Stage 1 cuts the digit lists into equal slices and gives each slice to one lane
// (list_sum_kernel); the slices of a list, and all later stages, are summed by quads (list_sum_coop_kernel).
//   MODE 0: members are buckets, sums are XYZZ partials at src[item_start[b]] (skipped when count[b] == 0)
//   MODE 1: same, but the sums are affine points (batch-affine accumulation)
//   MODE 2: members index a dense XYZZ array (output of the previous stage)
// ------------------------------------------------------------------------------------------------
template <class F, int MODE>
static __global__ void __launch_bounds__(128) list_sum_kernel(const void *__restrict__ src, const uint32_t *__restrict__ count,
                                                              const uint32_t *__restrict__ item_start, uint32_t in_stride,
                                                              const uint32_t *__restrict__ start, const uint32_t *__restrict__ idx,
                                                              uint32_t nlists, uint32_t nwindows, xyzz_t<F> *__restrict__ out) {
    // one LANE per list (stage 1: the lists are equal-length slices, so the lanes of a warp have equal trip counts);
    // the accumulator never has its address taken, so it lives in registers
    const uint32_t gl = blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= nlists * nwindows) return;
    const uint32_t w = gl / nlists, li = gl % nlists;
    xyzz_t<F> acc;
    xyzz_set_inf(acc);
    const uint32_t e1 = start[li + 1];
#pragma unroll 1
    for (uint32_t e = start[li]; e < e1; e++) {
        const size_t b = (size_t)w * in_stride + idx[e];
        if (MODE == 2) {
            xyzz_t<F> s;
            load_xyzz(s, (const xyzz_t<F> *)src + b);
            xyzz_add(acc, s);
        } else if (count[b] != 0) {
            if (MODE == 1) {
                aff_t<F> a;
                load_affine(a, (const aff_t<F> *)src, item_start[b], (uint32_t)(sizeof(aff_t<F>) / 16));
                xyzz_add_affine(acc, a, false);
            } else {
                xyzz_t<F> s;
                load_xyzz(s, (const xyzz_t<F> *)src + item_start[b]);
                xyzz_add(acc, s);
            }
        }
    }
    out[gl] = acc;
}


// Lane-cooperative variants (coop.cuh: a point is distributed over a GROUP of 4 lanes for G1, 8 lanes for G2): list
// sums with one group per team slot (tq groups per list, tq a power of two <= groups per warp), and the Horner pass
// with the bit positions spread over the groups of one warp. Control flow is warp-uniform; trip counts are the warp
// maximum and idle groups add infinity.
//   MODE 0: members are buckets; their sums are the XYZZ partials src[item_start[b]] (skipped when count[b] == 0)
//   MODE 2: members index a dense XYZZ array
template <class F, int MODE>
static __global__ void __launch_bounds__(128) list_sum_coop_kernel(const xyzz_t<F> *__restrict__ src, const uint32_t *__restrict__ count,
                                                                   const uint32_t *__restrict__ item_start, uint32_t in_stride,
                                                                   const uint32_t *__restrict__ start, const uint32_t *__restrict__ idx,
                                                                   uint32_t nlists, uint32_t nwindows, uint32_t tq, xyzz_t<F> *__restrict__ out) {
    using C = typename coop_of<F>::type;
    constexpr uint32_t GL = coop_group_lanes<C>();
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gq = gt / GL, gl = gq / tq, sq = gq % tq;
    const uint32_t total = nlists * nwindows;
    const bool live = gl < total;
    uint32_t e0 = 0, e1 = 0;
    size_t base = 0;
    if (live) {
        const uint32_t w = gl / nlists, li = gl % nlists;
        e0 = start[li] + sq;
        e1 = start[li + 1];
        base = (size_t)w * in_stride;
    }
    const uint32_t my_iters = e1 > e0 ? (e1 - e0 + tq - 1) / tq : 0u;
    const uint32_t iters = __reduce_max_sync(0xffffffffu, my_iters);
    C acc;
    f_set_zero(acc);
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it++) {
        const uint32_t e = e0 + it * tq;
        const bool have = e < e1;
        C s;
        f_set_zero(s);
        if (MODE == 2) {
            if (have) dq_load(s, src + base + idx[e]);
        } else if (have) {
            const size_t b = base + idx[e];
            if (count[b] != 0) dq_load(s, src + item_start[b]);
        }
        if (it == 0) acc = s;
        else dq_add(acc, s);
    }
#pragma unroll 1
    for (uint32_t o = tq >> 1; o > 0; o >>= 1) {
        C other;
        dq_shfl_down(other, acc, (int)(GL * o));
        if (sq + o >= tq) f_set_zero(other);
        dq_add(acc, other);
    }
    if (live && sq == 0) dq_store(out + gl, acc);
}

// Buckets split into several work items: tq lane groups per bucket (1 for the light list, a whole warp for the medium
// list) stride over the partials and fold them into the first one.
template <class F>
static __global__ void __launch_bounds__(128) combine_light_kernel(const uint32_t *__restrict__ count, const uint32_t *__restrict__ item_start,
                                                                   const uint32_t *__restrict__ list, uint32_t item_len, uint32_t tq,
                                                                   xyzz_t<F> *partial) {
    using C = typename coop_of<F>::type;
    constexpr uint32_t GL = coop_group_lanes<C>();
    const uint32_t nlist = list[0];
    const uint32_t gq = (blockIdx.x * blockDim.x + threadIdx.x) / GL, gb = gq / tq, sq = gq % tq;
    if (__all_sync(0xffffffffu, gb >= nlist)) return;
    uint32_t first = 0, nitems = 0;
    if (gb < nlist) {
        const uint32_t b = list[1 + gb];
        first = item_start[b];
        nitems = (count[b] + item_len - 1) / item_len;
    }
    const uint32_t my_iters = nitems > sq ? (nitems - sq + tq - 1) / tq : 0u;
    const uint32_t iters = __reduce_max_sync(0xffffffffu, my_iters);
    C acc;
    f_set_zero(acc);
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it++) {
        const uint32_t k = sq + it * tq;
        C s;
        f_set_zero(s);
        if (k < nitems) dq_load(s, partial + first + k);
        if (it == 0) acc = s;
        else dq_add(acc, s);
    }
#pragma unroll 1
    for (uint32_t o = tq >> 1; o > 0; o >>= 1) {
        C other;
        dq_shfl_down(other, acc, (int)(GL * o));
        if (sq + o >= tq) f_set_zero(other);
        dq_add(acc, other);
    }
    if (nitems && sq == 0) dq_store(partial + first, acc);
}

template <class C> __device__ __forceinline__ void dq_shift(C &acc, uint32_t my_doublings) {
    const uint32_t maxd = __reduce_max_sync(0xffffffffu, my_doublings);
#pragma unroll 1
    for (uint32_t i = 0; i < maxd; i++) {
        C d;
        dq_double(d, acc);
        if (i < my_doublings) acc = d;
    }
}
// Horner over bit positions: L[w * nbits_w + k] is the sum of everything whose value has bit k set in window w, i.e. the
// result is sum_w sum_k 2^(w * wbits + k) L[w][k]. Replaces the tail of integrate_buckets and the outer loop of
// POINTonE1s_mult_pippenger (src/multi_scalar.c:565-575), xyzz_to_Jacobian and (want_affine) blst_p1_to_affine.
// One warp of NG groups. Entries in descending bit position: t = 0..G-1 <-> (w = nwindows-1 - t / nbits_w,
// k = nbits_w-1 - t % nbits_w), position w * wbits + k. Group j runs Horner over entries [j*len, (j+1)*len); the NG
// partial results are combined by a tree in which the group holding the higher positions is doubled down to its
// partner's lowest position.
template <class F>
static __global__ void __launch_bounds__(32) bits_finalize_coop_kernel(const xyzz_t<F> *__restrict__ L, uint32_t nwindows, uint32_t nbits_w,
                                                                      uint32_t wbits, xyzz_t<F> *__restrict__ scratch,
                                                                      jac_t<F> *__restrict__ out_jac, aff_t<F> *__restrict__ out_aff) {
    using C = typename coop_of<F>::type;
    constexpr uint32_t GL = coop_group_lanes<C>(), NG = 32 / GL;
    const uint32_t G = nwindows * nbits_w, j = threadIdx.x / GL;
    const uint32_t len = (G + NG - 1) / NG;
    const uint32_t t0 = min(G, j * len), t1 = min(G, t0 + len);
    C acc;
    f_set_zero(acc);
    uint32_t low = 0;  // lowest position folded into acc so far (0 for an empty group)
#pragma unroll 1
    for (uint32_t s = 0; s < len; s++) {
        const uint32_t t = t0 + s;
        C e;
        f_set_zero(e);
        uint32_t dbl = 0;
        if (t < t1) {
            const uint32_t w = nwindows - 1 - t / nbits_w, k = nbits_w - 1 - t % nbits_w;
            dq_load(e, L + (size_t)w * nbits_w + k);
            const uint32_t p = w * wbits + k;
            if (s > 0) dbl = low - p;
            low = p;
        }
        if (s == 0) { acc = e; continue; }  // warp-uniform
        dq_shift(acc, dbl);
        dq_add(acc, e);
    }
    if (t0 >= t1) low = 0;
#pragma unroll 1
    for (uint32_t o = 1; o < NG; o <<= 1) {
        C other;
        dq_shfl_down(other, acc, (int)(GL * o));
        const uint32_t low_other = __shfl_down_sync(0xffffffffu, low, GL * o);
        const bool recv = (j % (2 * o)) == 0;
        if (!recv) f_set_zero(other);
        dq_shift(acc, recv ? low - low_other : 0u);
        dq_add(acc, other);
        if (recv) low = low_other;
    }
    dq_shift(acc, j == 0 ? low : 0u);
    // gather the distributed result of group 0 into one thread for the (single) inversion
    if (j == 0) dq_store(scratch, acc);
    __syncwarp();
    if (threadIdx.x != 0) return;
    xyzz_t<F> full;
    load_xyzz(full, scratch);
    jac_t<F> jj;
    if (xyzz_is_inf(full)) jac_set_inf(jj);
    else xyzz_to_jac(jj, full);
    if (out_jac) *out_jac = jj;
    if (out_aff) {
        aff_t<F> a;
        jac_to_affine(a, jj);
        *out_aff = a;
    }
}

// ------------------------------------------------------------------------------------------------
// [5] window combination + output. Replaces the outer loop of POINTonE1s_mult_pippenger
// (src/multi_scalar.c:565-575: ret = (ret + tile) * 2^window, top tile first), xyzz_to_Jacobian
// (src/ec_ops.h:771-777) and, when want_affine, blst_p1_to_affine (src/e1.c:80-92). One thread.
// ------------------------------------------------------------------------------------------------
template <class F>
static __global__ void finalize_kernel(const xyzz_t<F> *__restrict__ row_sums, uint32_t nwindows, uint32_t wbits, uint32_t vshift,
                                       jac_t<F> *__restrict__ out_jac, aff_t<F> *__restrict__ out_aff) {
    // row_sums[2w] = sum of the W rows, row_sums[2w+1] = sum of the c*T rows of window w: window sum = W + 2^vshift * P
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    xyzz_t<F> acc;
    xyzz_set_inf(acc);
    for (int w = (int)nwindows - 1; w >= 0; w--) {
        xyzz_t<F> s = row_sums[2 * w + 1];
        for (uint32_t k = 0; k < vshift; k++)
            if (!xyzz_is_inf(s)) { xyzz_t<F> a = s; xyzz_double(s, a); }
        xyzz_t<F> wrow = row_sums[2 * w];
        xyzz_add_cold(s, wrow);
        xyzz_add_cold(acc, s);
        if (w > 0)
            for (uint32_t k = 0; k < wbits; k++)
                if (!xyzz_is_inf(acc)) { xyzz_t<F> a = acc; xyzz_double(acc, a); }
    }
    jac_t<F> j;
    if (xyzz_is_inf(acc)) jac_set_inf(j);
    else xyzz_to_jac(j, acc);
    if (out_jac) *out_jac = j;
    if (out_aff) {
        aff_t<F> a;
        jac_to_affine(a, j);
        *out_aff = a;
    }
}

// Lane-cooperative form: one warp of NG groups, group j folds partials j, j+NG, ... (Jacobian (X, Y, Z) read as XYZZ
// (X, Y, Z^3, Z^2)), then a tree over the groups: ~14 multiplication levels for 8 GPUs instead of 7 dependent 16M additions.
template <class F>
static __global__ void __launch_bounds__(32) sum_partials_coop_kernel(const jac_t<F> *__restrict__ partials, int count, xyzz_t<F> *__restrict__ scratch,
                                                                     aff_t<F> *__restrict__ out_aff) {
    using C = typename coop_of<F>::type;
    constexpr int LPS = coop_traits<C>::LPS, GL = coop_group_lanes<C>(), NG = 32 / GL;
    const int l = coop_slot<C>(), j = threadIdx.x / GL, sub = threadIdx.x & (LPS - 1);
    C acc;
    f_set_zero(acc);
#pragma unroll 1
    for (int base = 0; base < count; base += NG) {
        const int k = base + j;
        C c, z;
        f_set_zero(c);
        f_set_zero(z);
        if (k < count) {
            const fp_t *p = reinterpret_cast<const fp_t *>(partials + k);  // pieces: x, y, z (each LPS x 48 bytes)
            const fp_t zp = p[2 * LPS + sub], cp = p[(l < 2 ? l : 2) * LPS + sub];
            *reinterpret_cast<fp_t *>(&z) = zp;
            *reinterpret_cast<fp_t *>(&c) = cp;
        }
        C zz, zzz;
        f_sqr(zz, z);
        f_mul(zzz, zz, z);
        fq_sel(c, l == 2, zzz, c);
        fq_sel(c, l == 3, zz, c);
        if (base == 0) acc = c;
        else dq_add(acc, c);
    }
#pragma unroll 1
    for (int o = 1; o < NG; o <<= 1) {
        C other;
        dq_shfl_down(other, acc, GL * o);
        if (j + o >= NG) f_set_zero(other);
        dq_add(acc, other);
    }
    if (j == 0) dq_store(scratch, acc);
    __syncwarp();
    if (threadIdx.x != 0) return;
    xyzz_t<F> full;
    load_xyzz(full, scratch);
    jac_t<F> jj;
    if (xyzz_is_inf(full)) jac_set_inf(jj);
    else xyzz_to_jac(jj, full);
    aff_t<F> a;
    jac_to_affine(a, jj);
    *out_aff = a;
}

// Multi-GPU combine, first half: entry e of the result = sum over the ranks of gathered[g * entries + e] (the per-bit
// sums every rank contributes, all-gathered); one lane group per entry. The Horner pass + to_affine that follow run ONCE
// (bits_finalize_coop_kernel) instead of once per rank plus a second inversion after the gather.
template <class F>
static __global__ void __launch_bounds__(128) sum_ranks_coop_kernel(const xyzz_t<F> *__restrict__ gathered, uint32_t entries, uint32_t world,
                                                                    xyzz_t<F> *__restrict__ out) {
    using C = typename coop_of<F>::type;
    constexpr uint32_t GL = coop_group_lanes<C>();
    const uint32_t gq = (blockIdx.x * blockDim.x + threadIdx.x) / GL;
    const bool live = gq < entries;
    C acc;
    f_set_zero(acc);
#pragma unroll 1
    for (uint32_t g = 0; g < world; g++) {
        C s;
        f_set_zero(s);
        if (live) dq_load(s, gathered + (size_t)g * entries + gq);
        if (g == 0) acc = s;
        else dq_add(acc, s);
    }
    if (live) dq_store(out + gq, acc);
}

// ------------------------------------------------------------------------------------------------
// precomputation tables and fixed points
// ------------------------------------------------------------------------------------------------
// Normalise up to 3 Jacobian points with one inversion (Montgomery's trick); Z == 0 -> (0,0).
template <class F>
__device__ __forceinline__ void to_affine3(aff_t<F> *o0, const jac_t<F> &p0, aff_t<F> *o1, const jac_t<F> &p1,
                                           aff_t<F> *o2, const jac_t<F> &p2) {
    F one;
    f_set_one(one);
    bool z0 = f_is_zero(p0.z), z1 = f_is_zero(p1.z), z2 = f_is_zero(p2.z);
    F a = z0 ? one : p0.z, b = z1 ? one : p1.z, c = z2 ? one : p2.z;
    F ab, abc, inv, t, ia, ib, ic;
    f_mul(ab, a, b);
    f_mul(abc, ab, c);
    f_inv(inv, abc);
    f_mul(ic, inv, ab);    // 1/c
    f_mul(t, inv, c);      // 1/(ab)
    f_mul(ib, t, a);       // 1/b
    f_mul(ia, t, b);       // 1/a
    F zi2;
    f_sqr(zi2, ia); f_mul(o0->x, p0.x, zi2); f_mul(zi2, zi2, ia); f_mul(o0->y, p0.y, zi2);
    f_sqr(zi2, ib); f_mul(o1->x, p1.x, zi2); f_mul(zi2, zi2, ib); f_mul(o1->y, p1.y, zi2);
    f_sqr(zi2, ic); f_mul(o2->x, p2.x, zi2); f_mul(zi2, zi2, ic); f_mul(o2->y, p2.y, zi2);
    if (z0) { f_set_zero(o0->x); f_set_zero(o0->y); }
    if (z1) { f_set_zero(o1->x); f_set_zero(o1->y); }
    if (z2) { f_set_zero(o2->x); f_set_zero(o2->y); }
}
// blst_p1s_to_affine / blst_p2s_to_affine (src/multi_scalar.c:17-59): batched normalisation. The reference shares one
// inversion over the whole batch (Montgomery's trick, serial); here every thread normalises 3 consecutive points with
// one inversion, all threads in parallel. Affine output is canonical, so the bytes equal the reference's.
template <class F>
static __global__ void __launch_bounds__(128) batch_to_affine_kernel(const jac_t<F> *__restrict__ in, size_t n, aff_t<F> *__restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, i = 3 * t;
    if (i >= n) return;
    jac_t<F> p0 = in[i], p1, p2;
    jac_set_inf(p1);
    jac_set_inf(p2);
    if (i + 1 < n) p1 = in[i + 1];
    if (i + 2 < n) p2 = in[i + 2];
    aff_t<F> a0, a1, a2;
    to_affine3(&a0, p0, &a1, p1, &a2, p2);
    out[i] = a0;
    if (i + 1 < n) out[i + 1] = a1;
    if (i + 2 < n) out[i + 2] = a2;
}
template <class F> __device__ __forceinline__ void jac_from_affine(jac_t<F> &j, const aff_t<F> &a) {
    j.x = a.x; j.y = a.y;
    if (aff_is_inf(a)) f_set_zero(j.z); else f_set_one(j.z);
}
// Table build. Replaces the loops of init_pippenger_CHES_q_over_5 (main_p1.cpp:156-172, nmult = 3:
// T[3(i*h+j)+m-1] = m q^j P_i) and init_pippenger_BGMW95 (main_p1.cpp:108-115, nmult = 1: T[i*h+j] = q^j P_i),
// i.e. h * (e doublings + 2 additions) per point instead of single_scalar_multiplication + one inversion per
// entry; one thread per fixed point, one shared inversion per (i, j).
template <class F>
static __global__ void __launch_bounds__(128) table_build_kernel(const aff_t<F> *__restrict__ points, size_t n, int h, int e, int nmult,
                                                          aff_t<F> *__restrict__ table, uint32_t stride16) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    aff_t<F> q = points[i];
#pragma unroll 1
    for (int j = 0; j < h; j++) {
        size_t base = (i * h + j) * nmult;
        *table_entry(table, base, stride16) = q;
        jac_t<F> jq, d, t3;
        jac_from_affine(jq, q);
        jac_double(d, jq);                        // 2Q (Z stays 0 for infinity)
        if (nmult == 3) jac_add(t3, d, jq);       // 3Q
        else jac_set_inf(t3);
        jac_t<F> nx = d;                          // q * Q = 2^(e-1) * 2Q
#pragma unroll 1
        for (int k = 1; k < e; k++) jac_double(nx, nx);
        aff_t<F> a2, a3, an;
        to_affine3(&a2, d, &a3, t3, &an, nx);
        if (nmult == 3) { *table_entry(table, base + 1, stride16) = a2; *table_entry(table, base + 2, stride16) = a3; }
        q = an;
    }
}
// out[i] = k_i * G_i for 256-bit scalars (MSB-first double-and-add), Jacobian; used to seed the fixed-point
// chains and for test vectors.
template <class F>
static __global__ void __launch_bounds__(128) scalar_mul_kernel(const aff_t<F> *__restrict__ bases, int base_stride,
                                                         const uint32_t *__restrict__ scalars, size_t n, jac_t<F> *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    load_scalar(s, scalars, i);
    jac_t<F> b, acc;
    jac_from_affine(b, bases[i * base_stride]);
    jac_set_inf(acc);
#pragma unroll 1
    for (int bit = 255; bit >= 0; bit--) {
        jac_double(acc, acc);
        if ((s[bit >> 5] >> (bit & 31)) & 1) jac_add(acc, acc, b);
    }
    out[i] = acc;
}
// blst_p1s_mult_wbits_precompute (src/multi_scalar.c:81-158): table[i * nwin + k] = (k + 1) * P_i, affine, nwin = 2^(wbits-1).
// One thread per entry (MSB-first double-and-add over the <= wbits bits of k + 1, then to_affine); affine points are
// canonical, so the bytes equal the reference's table. An infinity input gives a row of infinities.
template <class F>
static __global__ void __launch_bounds__(128) wbits_precompute_kernel(const aff_t<F> *__restrict__ points, size_t npoints, int wbits,
                                                                      aff_t<F> *__restrict__ table) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nwin = (size_t)1 << (wbits - 1);
    if (t >= npoints * nwin) return;
    const uint32_t mult = (uint32_t)(t % nwin) + 1u;
    jac_t<F> b, acc;
    jac_from_affine(b, points[t / nwin]);
    jac_set_inf(acc);
#pragma unroll 1
    for (int bit = 31 - __clz((int)mult); bit >= 0; bit--) {
        jac_double(acc, acc);
        if ((mult >> bit) & 1u) jac_add(acc, acc, b);
    }
    aff_t<F> a;
    jac_to_affine(a, acc);
    table[t] = a;
}
// init_fix_point_list (main_p1.cpp:52-66) in parallel: thread t starts from seeds[t] = 2^(first + t*chunk) G and
// emits the next `chunk` doublings, each normalised (3 at a time) to canonical affine.
template <class F>
static __global__ void __launch_bounds__(128) fix_points_kernel(const jac_t<F> *__restrict__ seeds, size_t nthreads, uint32_t chunk, size_t n,
                                                         aff_t<F> *__restrict__ points) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    jac_t<F> cur = seeds[t];
    size_t i0 = t * chunk;
#pragma unroll 1
    for (uint32_t k = 0; k < chunk; k += 3) {
        jac_t<F> p0, p1, p2;
        jac_double(p0, cur);
        jac_double(p1, p0);
        jac_double(p2, p1);
        cur = p2;
        aff_t<F> a0, a1, a2;
        to_affine3(&a0, p0, &a1, p1, &a2, p2);
        if (k + 0 < chunk && i0 + k + 0 < n) points[i0 + k + 0] = a0;
        if (k + 1 < chunk && i0 + k + 1 < n) points[i0 + k + 1] = a1;
        if (k + 2 < chunk && i0 + k + 2 < n) points[i0 + k + 2] = a2;
    }
}

// ------------------------------------------------------------------------------------------------
// batched building blocks for parity tests (msmb200_test_field_op / msmb200_test_point_op)
// ------------------------------------------------------------------------------------------------
template <class F>
static __global__ void field_op_kernel(int op, const F *__restrict__ a, const F *__restrict__ b, F *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (op == 7) {  // the per-lane branch-free inversion of the batch-affine rounds (inv.cuh): every lane takes part in the vote
        F x, r;
        if (i < n) x = a[i]; else f_set_one(x);
        f_inv_warp(r, x);
        if (i < n) out[i] = r;
        return;
    }
    if (i >= n) return;
    F x = a[i], y, r;
    if (b) y = b[i]; else f_set_zero(y);
    switch (op) {
    case 0: f_mul(r, x, y); break;
    case 1: f_sqr(r, x); break;
    case 2: f_add(r, x, y); break;
    case 3: f_sub(r, x, y); break;
    case 4: f_cneg(r, x, true); break;
    case 5: f_mul3(r, x); break;
    case 6: f_inv(r, x); break;
    default: f_set_zero(r);
    }
    out[i] = r;
}
// ops 2,3 (the XYZZ additions of the two hot loops) are instantiated with the hot field type, the rest cold
template <class F>
static __global__ void __launch_bounds__(128) point_op_xyzz_kernel(int op, const void *__restrict__ a, const void *__restrict__ b,
                                                                   const unsigned char *__restrict__ flags, void *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op == 2) {
        xyzz_t<F> x = ((const xyzz_t<F> *)a)[i];
        aff_t<F> y = ((const aff_t<F> *)b)[i];
        xyzz_add_affine(x, y, flags && flags[i]);
        ((xyzz_t<F> *)out)[i] = x;
    } else {
        xyzz_t<F> x = ((const xyzz_t<F> *)a)[i], y = ((const xyzz_t<F> *)b)[i];
        xyzz_add(x, y);
        ((xyzz_t<F> *)out)[i] = x;
    }
}
// ------------------------------------------------------------------------------------------------
// table persistence: blst_p1_affine_serialize / blst_p2_affine_serialize for every entry
// (src/e1.c:139-162, src/e2.c:176-203: from Montgomery, 48-byte big-endian X | Y, G2 as X.im | X.re | Y.im | Y.re,
// infinity = 0x40 then zeros) and the inverse with the checks of blst_pN_deserialize (src/e1.c:303-330: flags,
// coordinates < p, y^2 = x^3 + B with B = 4 resp. 4 + 4i, src/e1.c:14, src/e2.c:14).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fp_to_be48(uint32_t out[12], const fp_t &mont) {
    fp_t v;
    fp_from_mont(v, mont);
#pragma unroll
    for (int w = 0; w < 12; w++) out[w] = __byte_perm(v.l[11 - w], 0, 0x0123);
}
__device__ __forceinline__ bool fp_from_be48(fp_t &mont, const uint32_t in[12]) {  // false when the value is >= p
    fp_t v, rr;
#pragma unroll
    for (int w = 0; w < 12; w++) { v.l[11 - w] = __byte_perm(in[w], 0, 0x0123); rr.l[w] = fp_rr_limb(w); }
    uint32_t t = sub_cc(v.l[0], fp_p_limb(0));
#pragma unroll
    for (int i = 1; i < 12; i++) t = subc_cc(v.l[i], fp_p_limb(i));
    (void)t;
    const bool below = subc(0, 0) != 0;
    fp_mul(mont, v, rr);
    return below;
}
template <class F> struct fp_count;
template <> struct fp_count<fp_t> { static constexpr int N = 1; };
template <> struct fp_count<fpc_t> { static constexpr int N = 1; };
template <> struct fp_count<fp2_t> { static constexpr int N = 2; };
// y^2 == x^3 + B
__device__ __forceinline__ bool on_curve(const aff_t<fpc_t> &p) {
    fpc_t x2, x3, y2, b;
    f_sqr(x2, p.x); f_mul(x3, x2, p.x); f_sqr(y2, p.y);
    f_set_one(b); f_dbl(b, b); f_dbl(b, b);
    f_add(x3, x3, b);
    return f_eq(x3, y2);
}
__device__ __forceinline__ bool on_curve(const aff_t<fp2_t> &p) {
    fp2_t x2, x3, y2, b;
    f_sqr(x2, p.x); f_mul(x3, x2, p.x); f_sqr(y2, p.y);
    fp_set_one(b.c0); fp_dbl(b.c0, b.c0); fp_dbl(b.c0, b.c0);
    b.c1 = b.c0;
    f_add(x3, x3, b);
    return f_eq(x3, y2);
}
template <class F>
static __global__ void __launch_bounds__(128) table_serialize_kernel(const aff_t<F> *__restrict__ in, size_t n, uint32_t *__restrict__ out) {
    constexpr int NF = fp_count<F>::N;  // Fp elements per coordinate
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    aff_t<F> p = in[i];
    uint32_t *o = out + i * (24 * NF);
    if (aff_is_inf(p)) {
#pragma unroll
        for (int w = 0; w < 24 * NF; w++) o[w] = 0;
        o[0] = 0x40u;  // first BYTE 0x40 (little-endian word)
        return;
    }
    const fp_t *coord = reinterpret_cast<const fp_t *>(&p);  // G1: x, y; G2: x.c0, x.c1, y.c0, y.c1
#pragma unroll
    for (int k = 0; k < 2 * NF; k++) {
        const int src = NF == 1 ? k : (k ^ 1);  // G2: imaginary part first
        uint32_t be[12];
        fp_to_be48(be, coord[src]);
#pragma unroll
        for (int w = 0; w < 12; w++) o[12 * k + w] = be[w];
    }
}
// position-weighted 64-bit checksum of an array of 64-bit words (table files bind to the fixed points they were built from)
static __global__ void checksum_kernel(const unsigned long long *__restrict__ w, size_t n, unsigned long long *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += w[i] * (0x9E3779B97F4A7C15ull * (i + 1) | 1ull);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
// which entries fail is not needed: the count of bad entries (flags, range, curve equation) is accumulated
template <class F>
static __global__ void __launch_bounds__(128) table_deserialize_kernel(const uint32_t *__restrict__ in, size_t n, int serialized,
                                                                       aff_t<F> *__restrict__ out, uint32_t *__restrict__ bad) {
    constexpr int NF = fp_count<F>::N;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    aff_t<F> p;
    bool ok = true;
    if (serialized) {
        const uint32_t *s = in + i * (24 * NF);
        const uint32_t flags = s[0] & 0xe0u;
        if (flags & 0x40u) {  // infinity: everything else must be zero
            uint32_t acc = s[0] & ~0x40u;
#pragma unroll
            for (int w = 1; w < 24 * NF; w++) acc |= s[w];
            ok = acc == 0 && flags == 0x40u;
            fp_t *c = reinterpret_cast<fp_t *>(&p);
#pragma unroll
            for (int k = 0; k < 2 * NF; k++) fp_set_zero(c[k]);
            if (ok) { out[i] = p; return; }
        } else if (flags) {
            ok = false;  // compressed encodings are not table entries
        }
        fp_t *c = reinterpret_cast<fp_t *>(&p);
#pragma unroll
        for (int k = 0; k < 2 * NF; k++) {
            const int dst = NF == 1 ? k : (k ^ 1);
            uint32_t be[12];
#pragma unroll
            for (int w = 0; w < 12; w++) be[w] = s[12 * k + w];
            ok = fp_from_be48(c[dst], be) && ok;
        }
    } else {
        p = reinterpret_cast<const aff_t<F> *>(in)[i];
        const fp_t *c = reinterpret_cast<const fp_t *>(&p);
#pragma unroll
        for (int k = 0; k < 2 * NF; k++) {  // Montgomery limbs must be fully reduced
            uint32_t t = sub_cc(c[k].l[0], fp_p_limb(0));
#pragma unroll
            for (int j = 1; j < 12; j++) t = subc_cc(c[k].l[j], fp_p_limb(j));
            (void)t;
            ok = (subc(0, 0) != 0) && ok;
        }
    }
    if (ok && !aff_is_inf(p)) ok = on_curve(p);
    if (!ok) atomicAdd(bad, 1u);
    out[i] = p;
}

// ops 6,7: the lane-cooperative XYZZ addition / doubling (coop.cuh), one group (4 lanes G1, 8 lanes G2) per element
template <class F>
static __global__ void __launch_bounds__(128) point_op_coop_kernel(int op, const void *__restrict__ a, const void *__restrict__ b,
                                                                   void *__restrict__ out, size_t n) {
    using C = typename coop_of<F>::type;
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / coop_group_lanes<C>();
    C x, y;
    f_set_zero(x);
    f_set_zero(y);
    if (i < n) {
        dq_load(x, (const xyzz_t<F> *)a + i);
        if (op == 6) dq_load(y, (const xyzz_t<F> *)b + i);
    }
    if (op == 6) dq_add(x, y);
    else { C d; dq_double(d, x); x = d; }
    if (i < n) dq_store((xyzz_t<F> *)out + i, x);
}
template <class F>
static __global__ void __launch_bounds__(128) point_op_misc_kernel(int op, const void *__restrict__ a, const void *__restrict__ b,
                                                                   void *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (op) {
    case 0: { jac_t<F> x = ((const jac_t<F> *)a)[i], y = ((const jac_t<F> *)b)[i], r; jac_add(r, x, y); ((jac_t<F> *)out)[i] = r; break; }
    case 1: { jac_t<F> x = ((const jac_t<F> *)a)[i], r; jac_double(r, x); ((jac_t<F> *)out)[i] = r; break; }
    case 4: { xyzz_t<F> x = ((const xyzz_t<F> *)a)[i]; jac_t<F> r; if (xyzz_is_inf(x)) jac_set_inf(r); else xyzz_to_jac(r, x); ((jac_t<F> *)out)[i] = r; break; }
    case 5: { jac_t<F> x = ((const jac_t<F> *)a)[i]; aff_t<F> r; jac_to_affine(r, x); ((aff_t<F> *)out)[i] = r; break; }
    case 8: { xyzz_t<F> x = ((const xyzz_t<F> *)a)[i], r; xyzz_double(r, x); ((xyzz_t<F> *)out)[i] = r; break; }
    }
}

}  // namespace msmb200
