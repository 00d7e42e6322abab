// Bucket accumulation by BATCH-AFFINE pairwise rounds (sm_100a) — the GPU analogue of the reference's
// POINTonE{1,2}s_accumulate with its HEAD / TAIL macros (src/bulk_addition.c:51-143): affine + affine -> affine, the
// slope denominators of a whole batch inverted together by Montgomery's trick (5M + 1S per addition plus the shared
// inversion, :27), doubling folded into the same batch with denominator 2y (:63-74), P + (-P) and infinity operands
// resolved without arithmetic (:92-104).
//
// Shape of the computation. Every bucket is a list of k points; round r turns it into ceil(k / 2) points by adding
// neighbours (2i, 2i + 1); an odd last element is copied. After ceil(log2(max k)) rounds every bucket is ONE affine
// point (bucket_sum[b]). All additions of a round, over all buckets, are independent "slots".
//   * The slots of a round are enumerated by 16-byte RESOLVED descriptors {p, q, out, 0} (operand indices in the round's
//     source array: the precomputation table in round 0, bit 31 = negate; the previous round's points later; out = position
//     in the next round's buffer or FINAL | bucket), built before any arithmetic from the bucket histogram alone
//     (ba_totals / ba_scan_* / ba_emit_*): the arithmetic kernel never sees bucket boundaries, only real additions.
//   * ba_round_kernel is persistent: every WARP takes batches of up to ~110 rows (a row = 32 consecutive slots, lane t
//     handles slot 32 j + t, so descriptor, round-buffer and scratch accesses of a warp are contiguous) from an atomic
//     counter. Forward pass: denominator d_j, running product; the product BEFORE d_j goes to a global scratch array. One
//     inversion PER LANE by the branch-free safegcd of inv.cuh (all lanes run the same instruction stream, mostly on the
//     otherwise idle ALU pipe). Backward pass: slopes and results, 4M + 1S; the forward pass costs 1M.
//   * Operands reach the lanes through shared memory (cp.async), one slot ahead: 36 KB per block, 4 blocks per SM for Fp.
//     Round 0 gathers table entries warp-cooperatively (ba_coop_gather): the memory system is bound by REQUESTS there.
// Montgomery-form, fully reduced affine results: the bytes equal what any other correct summation order produces once
// normalised. What was measured on the way (and what lost) is in profiles/experiments/README.md.
#pragma once
#include <cstdint>
#include "ec.cuh"
#include "inv.cuh"

namespace msmb200 {

constexpr int BA_RMAX = 32;                   // rounds: counts are < 2^32
constexpr uint32_t BA_FINAL = 0x80000000u;    // output descriptor: bucket_sum[b] instead of the next round's buffer
constexpr uint32_t BA_HEAVY = 2048;           // buckets with more entries are planned by a whole block
constexpr int BA_THREADS = 128;
// resident blocks per SM the round kernel is compiled for (register budget 65536 / (128 x blocks) per thread)
#ifndef BA_MIN_BLOCKS_FP
#define BA_MIN_BLOCKS_FP 4
#endif
#ifndef BA_MIN_BLOCKS_FP2
#define BA_MIN_BLOCKS_FP2 2
#endif
template <class F> struct ba_cfg { static constexpr int MIN_BLOCKS = sizeof(F) > 48 ? BA_MIN_BLOCKS_FP2 : BA_MIN_BLOCKS_FP; };

// filled by the host once the per-round totals are known
struct BaRounds {
    uint32_t R;                                     // rounds with at least one addition
    uint32_t aoff[BA_RMAX + 1], coff[BA_RMAX + 1];  // first add / copy descriptor of round r
};

__device__ __forceinline__ uint32_t ba_len(uint32_t count, uint32_t r) { return count ? ((count - 1u) >> r) + 1u : 0u; }
// per-round quantities of a bucket with `c` entries: additions, copy flag, elements held in the round's input buffer
__device__ __forceinline__ void ba_row_values(uint32_t c, uint32_t r, uint32_t &adds, uint32_t &copy, uint32_t &elems) {
    const uint32_t k = ba_len(c, r);
    adds = k >= 2 ? k >> 1 : 0u;
    copy = ((k >= 3 && (k & 1u)) || (r == 0 && c == 1)) ? 1u : 0u;
    elems = r == 0 ? c : (k >= 2 ? k : 0u);
}

// totals[3 r + {0, 1, 2}] = sum over buckets of (adds, copies, elems) of round r; totals[3 BA_RMAX] = largest count
static __global__ void __launch_bounds__(256) ba_totals_kernel(const uint32_t *__restrict__ count, size_t nb, uint32_t *__restrict__ totals) {
    __shared__ uint32_t s[3 * BA_RMAX + 1];
    for (int i = threadIdx.x; i < 3 * BA_RMAX + 1; i += blockDim.x) s[i] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nb_pad = (nb + 31) & ~(size_t)31;   // whole warps take part in the votes
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb_pad; b += stride) {
        const uint32_t c = b < nb ? count[b] : 0u;
        const uint32_t rl = c >= 2 ? 32u - (uint32_t)__clz((int)(c - 1u)) : (c ? 1u : 0u);
        const uint32_t rw = __reduce_max_sync(0xffffffffu, rl);
        const uint32_t cmax = __reduce_max_sync(0xffffffffu, c);
        for (uint32_t r = 0; r < rw; r++) {
            uint32_t a, cp, e;
            ba_row_values(c, r, a, cp, e);
            a = __reduce_add_sync(0xffffffffu, a);
            cp = __reduce_add_sync(0xffffffffu, cp);
            e = __reduce_add_sync(0xffffffffu, e);
            if ((threadIdx.x & 31) == 0) {
                if (a) atomicAdd(&s[3 * r], a);
                if (cp) atomicAdd(&s[3 * r + 1], cp);
                if (e) atomicAdd(&s[3 * r + 2], e);
            }
        }
        if ((threadIdx.x & 31) == 0) atomicMax(&s[3 * BA_RMAX], cmax);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * BA_RMAX; i += blockDim.x)
        if (s[i]) atomicAdd(&totals[i], s[i]);
    if (threadIdx.x == 0) atomicMax(&totals[3 * BA_RMAX], s[3 * BA_RMAX]);
}

// ---- exclusive scans of the 3R rows over the buckets; one WARP per tile of 128 buckets, 4 consecutive buckets per lane ----
constexpr int BA_TILE = 128;
// tile_sums[tile * NR + row]
static __global__ void __launch_bounds__(256) ba_scan_tiles_kernel(const uint32_t *__restrict__ count, size_t nb, uint32_t R,
                                                                   uint32_t *__restrict__ tile_sums, size_t ntiles) {
    const size_t tile = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tile >= ntiles) return;
    const uint32_t lane = threadIdx.x & 31, NR = 3 * R;
    const size_t b0 = tile * BA_TILE + 4 * lane;
    uint32_t c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) c[k] = b0 + k < nb ? count[b0 + k] : 0u;
    for (uint32_t r = 0; r < R; r++) {
        uint32_t sa = 0, sc = 0, se = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t a, cp, e;
            ba_row_values(c[k], r, a, cp, e);
            sa += a; sc += cp; se += e;
        }
        sa = __reduce_add_sync(0xffffffffu, sa);
        sc = __reduce_add_sync(0xffffffffu, sc);
        se = __reduce_add_sync(0xffffffffu, se);
        if (lane == 0) {
            tile_sums[tile * NR + 3 * r] = sa;
            tile_sums[tile * NR + 3 * r + 1] = sc;
            tile_sums[tile * NR + 3 * r + 2] = se;
        }
    }
}
// block `row`: exclusive scan of tile_sums[. * NR + row] over the tiles, in place
static __global__ void __launch_bounds__(256) ba_scan_sums_kernel(uint32_t *tile_sums, size_t ntiles, uint32_t NR) {
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t carry_s;
    const uint32_t row = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (size_t base = 0; base < ntiles; base += 256) {
        const size_t i = base + threadIdx.x;
        const uint32_t v = i < ntiles ? tile_sums[i * NR + row] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)lane >= o) x += y;
        }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        uint32_t wbase = 0;
        for (uint32_t w = 0; w < wid; w++) wbase += wsum[w];
        const uint32_t carry = carry_s;
        if (i < ntiles) tile_sums[i * NR + row] = carry + wbase + x - v;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = carry + wbase + x;
        __syncthreads();
    }
}
// bases[row * nb_stride + b] = exclusive prefix of the row at bucket b
static __global__ void __launch_bounds__(256) ba_scan_apply_kernel(const uint32_t *__restrict__ count, size_t nb, uint32_t R,
                                                                   const uint32_t *__restrict__ tile_offs, size_t ntiles,
                                                                   uint32_t *__restrict__ bases, size_t nb_stride) {
    const size_t tile = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tile >= ntiles) return;
    const uint32_t lane = threadIdx.x & 31, NR = 3 * R;
    const size_t b0 = tile * BA_TILE + 4 * lane;
    uint32_t c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) c[k] = b0 + k < nb ? count[b0 + k] : 0u;
    for (uint32_t r = 0; r < R; r++) {
        uint32_t v[3][4], s[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            ba_row_values(c[k], r, v[0][k], v[1][k], v[2][k]);
            s[0] += v[0][k]; s[1] += v[1][k]; s[2] += v[2][k];
        }
#pragma unroll
        for (int w = 0; w < 3; w++) {
            uint32_t x = s[w];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if ((int)lane >= o) x += y;
            }
            uint32_t run = tile_offs[tile * NR + 3 * r + w] + x - s[w];
            uint32_t *dst = bases + (size_t)(3 * r + w) * nb_stride + b0;   // nb_stride is a multiple of 4: aligned 16-byte store
            uint4 o4;
            o4.x = run; run += v[w][0];
            o4.y = run; run += v[w][1];
            o4.z = run; run += v[w][2];
            o4.w = run;
            if (b0 < nb_stride) *reinterpret_cast<uint4 *>(dst) = o4;
        }
    }
}

// ---- descriptors: one thread per bucket (block per heavy bucket) writes every round's slots of its bucket ----
// Descriptors are RESOLVED: {p, q, out, 0} with p, q = index of the operand in the round's source array (the
// precomputation table in round 0, bit 31 = negate; the previous round's points later) — the arithmetic kernel needs no
// further indirection. Copy descriptors are {p, out}.
__device__ __forceinline__ void ba_emit_round(uint32_t b, uint32_t c, uint32_t r, const uint32_t *__restrict__ bases, size_t nbs,
                                              const BaRounds &rd, const uint32_t *__restrict__ sorted, uint4 *__restrict__ adesc,
                                              uint2 *__restrict__ cdesc, uint32_t i0, uint32_t istep) {
    const uint32_t k = ba_len(c, r), a = k >> 1, k1 = (k + 1) >> 1;
    const uint32_t eb = bases[(size_t)(3 * r + 2) * nbs + b];
    const uint32_t ob = k1 == 1 ? (BA_FINAL | b) : bases[(size_t)(3 * r + 5) * nbs + b];
    const uint32_t ab = rd.aoff[r] + bases[(size_t)(3 * r) * nbs + b];
    for (uint32_t i = i0; i < a; i += istep) {
        const uint32_t in = eb + 2 * i;
        adesc[ab + i] = make_uint4(r == 0 ? sorted[in] : in, r == 0 ? sorted[in + 1] : in + 1, k1 == 1 ? ob : ob + i, 0u);
    }
    if (i0 == 0 && (k & 1u)) {  // k >= 3 here
        const uint32_t in = eb + k - 1;
        cdesc[rd.coff[r] + bases[(size_t)(3 * r + 1) * nbs + b]] = make_uint2(r == 0 ? sorted[in] : in, ob + a);
    }
}
static __global__ void __launch_bounds__(256) ba_emit_kernel(const uint32_t *__restrict__ count, size_t nb, const uint32_t *__restrict__ bases, size_t nbs,
                                                             BaRounds rd, const uint32_t *__restrict__ sorted, uint4 *__restrict__ adesc,
                                                             uint2 *__restrict__ cdesc, uint32_t *__restrict__ heavy /* [0] = count, then bucket ids */,
                                                             uint4 *__restrict__ bucket_sum16, uint32_t aff_chunks) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t c = count[b];
    if (c == 0) {  // an empty bucket sums to infinity (the reduction may read it)
        for (uint32_t k = 0; k < aff_chunks; k++) bucket_sum16[b * aff_chunks + k] = make_uint4(0, 0, 0, 0);
        return;
    }
    if (c == 1) {  // a single entry goes straight to the bucket sum (copy list of round 0)
        cdesc[rd.coff[0] + bases[(size_t)1 * nbs + b]] = make_uint2(sorted[bases[(size_t)2 * nbs + b]], BA_FINAL | (uint32_t)b);
        return;
    }
    if (c > BA_HEAVY) { heavy[1 + atomicAdd(&heavy[0], 1u)] = (uint32_t)b; return; }
    for (uint32_t r = 0; r < rd.R && ba_len(c, r) >= 2; r++) ba_emit_round((uint32_t)b, c, r, bases, nbs, rd, sorted, adesc, cdesc, 0, 1);
}
static __global__ void __launch_bounds__(256) ba_emit_heavy_kernel(const uint32_t *__restrict__ count, const uint32_t *__restrict__ bases, size_t nbs,
                                                                   BaRounds rd, const uint32_t *__restrict__ sorted, uint4 *__restrict__ adesc,
                                                                   uint2 *__restrict__ cdesc, const uint32_t *__restrict__ heavy) {
    const uint32_t nheavy = heavy[0];
    for (uint32_t hi = blockIdx.x; hi < nheavy; hi += gridDim.x) {
        const uint32_t b = heavy[1 + hi], c = count[b];
        for (uint32_t r = 0; r < rd.R && ba_len(c, r) >= 2; r++) ba_emit_round(b, c, r, bases, nbs, rd, sorted, adesc, cdesc, threadIdx.x, blockDim.x);
    }
}

// ---- arithmetic ----
struct s30_warp_vote {
    MSMB200_HD bool operator()(bool done) const {
#if defined(__CUDA_ARCH__)
        return __all_sync(0xffffffffu, done) != 0;
#else
        return done;
#endif
    }
};
// 1/a, Montgomery form, 0 -> 0 (reciprocal_fp, src/recip.c:58-92). EVERY lane of the warp must call it (warp vote).
static __device__ __noinline__ fp_t fp_inv_warp_fn(fp_t a) {
    fp_t r;
    s30_inverse_words(r.l, a.l, s30_warp_vote());
    return r;
}
__device__ __forceinline__ void f_inv_warp(fp_t &r, const fp_t &a) { r = fp_inv_warp_fn(a); }
// 1/(a + b i) = (a - b i) / (a^2 + b^2)   (reciprocal_fp2, src/recip.c:100-114)
template <class F2> __device__ __forceinline__ void f_inv_warp_fp2(F2 &r, const F2 &a) {
    fp_t t0, t1;
    fp_sqr(t0, a.c0);
    fp_sqr(t1, a.c1);
    fp_add(t0, t0, t1);
    t1 = fp_inv_warp_fn(t0);
    fp_mul(r.c0, a.c0, t1);
    fp_mul(t0, a.c1, t1);
    fp_neg(r.c1, t0);
}
__device__ __forceinline__ void f_inv_warp(fp2_t &r, const fp2_t &a) { f_inv_warp_fp2(r, a); }

template <class F> __device__ __forceinline__ void f_ld(F &r, const F *p) {
    const uint4 *s = reinterpret_cast<const uint4 *>(p);
    uint4 *d = reinterpret_cast<uint4 *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) d[k] = __ldg(s + k);
}
template <class F> __device__ __forceinline__ void f_st(F *p, const F &v) {
    uint4 *d = reinterpret_cast<uint4 *>(p);
    const uint4 *s = reinterpret_cast<const uint4 *>(&v);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) d[k] = s[k];
}
// prefix-product scratch: 16-byte parts of slot s at part * stride + s (a warp's accesses are contiguous)
template <class F> __device__ __forceinline__ void ba_scratch_st(uint4 *scratch, size_t stride, size_t s, const F &v) {
    const uint4 *src = reinterpret_cast<const uint4 *>(&v);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) scratch[(size_t)k * stride + s] = src[k];
}
template <class F> __device__ __forceinline__ void ba_scratch_ld(F &v, const uint4 *scratch, size_t stride, size_t s) {
    uint4 *dst = reinterpret_cast<uint4 *>(&v);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) dst[k] = scratch[(size_t)k * stride + s];
}

enum { BA_TAG_ADD = 0, BA_TAG_DBL = 1, BA_TAG_COPY_P = 2, BA_TAG_COPY_Q = 3, BA_TAG_INF = 4 };

// rare operand patterns (an infinity operand, equal x): decided from the full points, identically in both passes.
// Everything by value: nothing of the hot loops has its address taken (no local memory).
template <class F> struct ba_cold_t { F d; int tag; };
template <class F>
static __device__ __noinline__ ba_cold_t<F> ba_classify_cold(const F *px, const F *py, bool sp, const F *qx, const F *qy, bool sq) {
    ba_cold_t<F> r;
    F x1, y1, x2, y2;
    f_ld(x1, px); f_ld(y1, py); f_ld(x2, qx); f_ld(y2, qy);
    f_cneg(y1, y1, sp);
    f_cneg(y2, y2, sq);
    f_sub(r.d, x2, x1);
    if (f_is_zero(x1) && f_is_zero(y1)) r.tag = BA_TAG_COPY_Q;
    else if (f_is_zero(x2) && f_is_zero(y2)) r.tag = BA_TAG_COPY_P;
    else if (!f_is_zero(r.d)) r.tag = BA_TAG_ADD;
    else if (f_eq(y1, y2) && !f_is_zero(y1)) { f_dbl(r.d, y1); r.tag = BA_TAG_DBL; }
    else r.tag = BA_TAG_INF;
    return r;
}

// Where a round reads its operands and writes its results. Round 0 reads the precomputation table (array of {x, y});
// between rounds the points live in SEPARATE x[] and y[] arrays, so the forward pass, which needs only x, streams half the
// bytes. Final results (one per bucket) go to bucket_sum as {x, y} for the reducers.
template <class F> struct ba_io {
    const aff_t<F> *table;       // round 0 only
    uint32_t tstride16;          // distance between table entries in 16-byte units (packed: sizeof(aff_t<F>) / 16; padded own tables: 8 / 16)
    const F *in_x, *in_y;        // rounds >= 1
    F *out_x, *out_y;
    aff_t<F> *bucket_sum;
};
template <class F> __device__ __forceinline__ const aff_t<F> *ba_entry(const ba_io<F> &io, uint32_t i) {
    return reinterpret_cast<const aff_t<F> *>(reinterpret_cast<const uint4 *>(io.table) + (size_t)i * io.tstride16);
}
template <class F, bool FIRST> __device__ __forceinline__ const F *ba_px(const ba_io<F> &io, uint32_t i) { return FIRST ? &ba_entry(io, i)->x : io.in_x + i; }
template <class F, bool FIRST> __device__ __forceinline__ const F *ba_py(const ba_io<F> &io, uint32_t i) { return FIRST ? &ba_entry(io, i)->y : io.in_y + i; }
template <class F> __device__ __forceinline__ void ba_store_point(uint32_t out, const ba_io<F> &io, const F &x, const F &y) {
    if (out & BA_FINAL) {
        aff_t<F> *dst = io.bucket_sum + (out & ~BA_FINAL);
        f_st(&dst->x, x);
        f_st(&dst->y, y);
    } else {
        f_st(io.out_x + out, x);
        f_st(io.out_y + out, y);
    }
}

// ---- asynchronous staging of slot operands in shared memory (cp.async, SASS LDGSTS) ----
// A lane does only 1 (forward) or 5 (backward) multiplications per slot, each an out-of-line call, and a call waits for
// every register load still in flight — so nothing the hot loops need may be a pending register load. Descriptors and
// operands are copied global -> shared ahead of use instead: thread t owns 16-byte chunk k of a stage at
// (k * 128 + t) * 16 (conflict-free). In rounds >= 1 a thread reads back only what it copied itself, so
// cp.async.wait_group is the only synchronisation; in round 0 the lanes of a warp fill each other's stages
// (ba_coop_gather) and a __syncwarp follows the wait. Forward: descriptors three slots ahead, x coordinates one slot
// ahead. Backward: the descriptor of slot j - 2 is requested at the top of slot j, the operands of slot j - 1 in the middle
// of slot j, right after slot j has taken its own into registers (three multiplications of lead).
__device__ __forceinline__ uint32_t ba_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ba_cp16(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the same past L1 (table entries are read once per pass and never again by this SM). Measured (gpurun_out/r2aq log): Fp2
// gains (G2 n=2^18 accumulate 3.39 -> 3.31 ms, n=2^20 11.48 -> 11.32 ms), Fp does not (7.2-7.3 ms either way), so only the
// Fp2 gather bypasses L1.
template <bool BYPASS_L1> __device__ __forceinline__ void ba_cp16_stream(uint32_t dst, const void *src) {
    if (BYPASS_L1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else ba_cp16(dst, src);
}
__device__ __forceinline__ void ba_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void ba_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t ba_chunk(int k) { return (uint32_t)((k * BA_THREADS + threadIdx.x) * 16); }
template <class F> __device__ __forceinline__ void ba_cp_field(uint32_t base, int chunk0, const F *src) {
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) ba_cp16(base + ba_chunk(chunk0 + k), reinterpret_cast<const uint4 *>(src) + k);
}
template <class F> __device__ __forceinline__ void ba_ld_stage(F &v, const unsigned char *base, int chunk0) {
    uint4 *dst = reinterpret_cast<uint4 *>(&v);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) dst[k] = *reinterpret_cast<const uint4 *>(base + ba_chunk(chunk0 + k));
}

// ---- round 0: warp-cooperative gather of table entries ----
// Measured on B200 (profiles/r2_gather_microbench.json): the memory system serves about 41 G REQUESTS per second that miss
// L2, whatever their size (32 .. 128 bytes of one line), and always fetches the whole 128-byte line from DRAM. A lane
// that pulls its own entry with 16-byte copies makes one request per 32-byte sector (3-4 per 96-byte entry, twice per
// round): 12.4 M slots x ~11 requests = 3.3 ms, the whole duration of round 0. Here the lanes of a warp fetch each
// other's operands instead: NCH consecutive lanes copy the NCH consecutive 16-byte chunks of ONE entry, so the piece of
// a line an entry occupies travels as one request. The owner's descriptor is read from the shared-memory ring (copied
// there by the owner; visible after cp.async.wait_group + __syncwarp), the chunks land in the owner's stage.
template <class F, int NCH>
__device__ __forceinline__ void ba_coop_gather(const aff_t<F> *__restrict__ table, uint32_t stride16, const unsigned char *shared, uint32_t sbase, int desc_chunk,
                                               int first_chunk, int dst_p, int dst_q) {
    constexpr int EPI = 32 / NCH;                 // entries per warp instruction
    constexpr int NI = (64 + EPI - 1) / EPI;      // 32 slots x 2 operands
    const uint32_t lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u, sub = lane / NCH, c = lane % NCH;
#pragma unroll
    for (int i = 0; i < NI; i++) {
        const uint32_t e = (uint32_t)i * EPI + sub;
        if (sub < (uint32_t)EPI && e < 64u) {
            const uint32_t o = e >> 1, which = e & 1u;
            const uint32_t idx = *reinterpret_cast<const uint32_t *>(shared + ((size_t)desc_chunk * BA_THREADS + wbase + o) * 16 + which * 4) & 0x7fffffffu;
            const uint4 *src = reinterpret_cast<const uint4 *>(table) + (size_t)idx * stride16 + first_chunk + c;
            ba_cp16_stream<(NCH > 3)>(sbase + (uint32_t)((((which ? dst_q : dst_p) + (int)c) * BA_THREADS + (int)(wbase + o)) * 16), src);
        }
    }
}
#ifdef MSMB200_BA_TIMING   // developer builds only: per-block phase time stamps (clock64) of every round
constexpr int BA_DBG_BLOCKS = 4096;
static __device__ unsigned long long ba_dbg[BA_RMAX][BA_DBG_BLOCKS][4];
#define BA_STAMP(k) do { if (threadIdx.x == 0 && blockIdx.x < BA_DBG_BLOCKS) ba_dbg[round][blockIdx.x][k] += clock64(); } while (0)
#else
#define BA_STAMP(k) do { } while (0)
#endif
// Staging depth (measured, gpurun_out/r2ap_stages.log): ONE stage ahead is enough in both passes — a slot is 1 (forward) or 5
// (backward) multiplications long — and the shallow rings are faster than two / three stages (G1 n=2^21 accumulate 7.29 vs
// 7.50 ms, G2 n=2^18 3.39 vs 3.45 ms): 36 KB instead of 54 KB per block leave the L1 more room for the descriptor, prefix
// and round-buffer lines that pass through it.
#ifndef BA_FWD_STAGES
#define BA_FWD_STAGES 2
#endif
#ifndef BA_BWD_X_STAGES
#define BA_BWD_X_STAGES 1
#endif
template <class F> struct ba_smem {
    static constexpr int FCH = (int)(sizeof(F) / 16);        // 16-byte chunks per field element
    static constexpr int FWD_DESC = 4, FWD_STAGES = BA_FWD_STAGES;   // descriptor ring / data stages (x1, x2, descriptor copy)
    static constexpr int FWD_STAGE_CH = 2 * FCH + 1;
    static constexpr int BWD_DESC = 3, BWD_X_CH = 3 * FCH;   // descriptor ring; one x stage: prefix product, x1, x2; one y stage: y1, y2
    static constexpr int BWD_X_STAGES = BA_BWD_X_STAGES;
    static constexpr int FWD_CH = FWD_DESC + FWD_STAGES * FWD_STAGE_CH, BWD_CH = BWD_DESC + BWD_X_STAGES * BWD_X_CH + 2 * FCH;
    // Fp: 18 / 18 chunks of 2 KB -> 36 KB per block, 4 blocks (16 warps) per SM by registers; Fp2: 30 / 33 chunks -> 66 KB, 2 blocks
    // by registers (three blocks at 168 registers were measured slower: profiles/experiments/README.md)
    static constexpr int BYTES = (FWD_CH > BWD_CH ? FWD_CH : BWD_CH) * BA_THREADS * 16;
};

// Work distribution of one round. The kernel is PERSISTENT (`grid` = the co-resident blocks) and every WARP is an
// independent worker: it takes batches of n slots per lane (32 n consecutive slots) from an atomic counter until the round
// is exhausted — one inversion per lane per batch. In a round with several batches per warp the first batch of a warp is
// shortened ((k % 3 + 1) / 3 of it, k = arrival order of its block on the SM), so the warps sharing a scheduler sit in
// different phases: the forward pass is DRAM-bound (one multiplication per slot), the inversion leaves the multiplier
// idle, the backward pass saturates it. Towards the end of the round the batches shrink to a third so that all warps
// finish close together. Rounds with a single batch per warp run unstaggered.
struct BaSched {
    uint32_t *counter;      // slots per lane (in units of 32-slot rows) handed out so far; zeroed before the launch
    uint32_t *sm_arrivals;  // per-SM block arrival count (never reset: only its value mod 3 matters)
    uint32_t rows;          // ceil(nadds / 32)
    uint32_t batch;         // rows per full batch
    uint32_t stagger;       // 1: shortened first batches + shrinking tail
};
__device__ __forceinline__ uint32_t ba_smid() { uint32_t r; asm("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

template <class F, bool FIRST>
static __global__ void __launch_bounds__(BA_THREADS, ba_cfg<F>::MIN_BLOCKS) ba_round_kernel(ba_io<F> io, const uint4 *__restrict__ adesc, uint32_t nadds,
                                                                     const uint2 *__restrict__ cdesc, uint32_t ncopies,
                                                                     uint4 *__restrict__ scratch, size_t scratch_stride, BaSched sched, uint32_t round) {
    using SM = ba_smem<F>;
    constexpr int FCH = SM::FCH;
    constexpr uint32_t IDX = 0x7fffffffu;
    extern __shared__ __align__(16) unsigned char ba_shared[];
    __shared__ uint32_t sh_k;
    if (threadIdx.x == 0) sh_k = sched.stagger ? atomicAdd(&sched.sm_arrivals[ba_smid()], 1u) % 3u : 2u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, nwarps = gridDim.x * (BA_THREADS / 32), last = nadds ? nadds - 1 : 0;
    const uint32_t sbase = ba_smem_addr(ba_shared);
    bool first_batch = true;
#pragma unroll 1
    for (; nadds != 0;) {
        uint32_t start = 0, n = 0;
        if (lane == 0) {
            n = sched.batch;
            if (sched.stagger) {
                if (first_batch) n = max(1u, (n * (sh_k + 1u)) / 3u);
                const uint32_t seen = *(volatile uint32_t *)sched.counter;
                const uint32_t left = seen < sched.rows ? sched.rows - seen : 0u;
                if (left < nwarps * sched.batch) n = min(n, max((sched.batch + 2u) / 3u, left / nwarps));
            }
            start = atomicAdd(sched.counter, n);
        }
        first_batch = false;
        start = __shfl_sync(0xffffffffu, start, 0);
        n = __shfl_sync(0xffffffffu, n, 0);
        if (start >= sched.rows) break;
        const uint32_t niter = min(n, sched.rows - start);
        const size_t tile0 = (size_t)start * 32 + lane;
        auto slot_of = [&](uint32_t j) { return tile0 + (size_t)j * 32; };
        auto desc_src = [&](uint32_t j) { const size_t s = slot_of(j); return adesc + (s < nadds ? s : last); };  // clamped: copies stay in bounds
        BA_STAMP(0);
        // ---- forward: denominators and their running product ----
        F run;
        f_set_one(run);
        {
            constexpr int ND = SM::FWD_DESC, NS = SM::FWD_STAGES, SCH = SM::FWD_STAGE_CH;
            auto stage_ch = [&](uint32_t j) { return ND + (int)(j % NS) * SCH; };
            auto issue_desc = [&](uint32_t j) { ba_cp16(sbase + ba_chunk((int)(j % ND)), desc_src(j)); };
            auto issue_data = [&](uint32_t j) {   // needs the descriptor of slot j in the ring
                const uint4 m = *reinterpret_cast<const uint4 *>(ba_shared + ba_chunk((int)(j % ND)));
                const int c0 = stage_ch(j);
                if (FIRST) ba_coop_gather<F, FCH>(io.table, io.tstride16, ba_shared, sbase, (int)(j % ND), 0, c0, c0 + FCH);   // x = chunks 0 .. FCH - 1 of an entry
                else {
                    ba_cp_field(sbase, c0, ba_px<F, FIRST>(io, m.x & IDX));
                    ba_cp_field(sbase, c0 + FCH, ba_px<F, FIRST>(io, m.y & IDX));
                }
                *reinterpret_cast<uint4 *>(ba_shared + ba_chunk(c0 + 2 * FCH)) = m;
            };
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)ND; j++) issue_desc(j);
            ba_cp_commit();
            ba_cp_wait<0>();
            if (FIRST) __syncwarp();        // round 0: the lanes read each other's descriptors and fill each other's stages
            constexpr int D = NS - 1;       // data of slot j + D is requested while slot j is processed
#pragma unroll
            for (int k = 0; k < D; k++) { issue_data((uint32_t)k); ba_cp_commit(); }
#pragma unroll 1
            for (uint32_t j = 0; j < niter; j++) {
                ba_cp_wait<D - 1>();        // everything but the D - 1 most recent groups has landed: the data of slot j
                if (FIRST) __syncwarp();
                issue_data(j + D);          // its descriptor arrived with the group of iteration j - 2 (or the prologue)
                issue_desc(j + D + 2);
                ba_cp_commit();
                const size_t s = slot_of(j);
                if (s < nadds) {
                    const int c0 = stage_ch(j);
                    F x1, d;
                    ba_ld_stage(x1, ba_shared, c0);
                    ba_ld_stage(d, ba_shared, c0 + FCH);
                    const bool odd = f_is_zero(x1) || f_is_zero(d);
                    f_sub(d, d, x1);
                    int tag = BA_TAG_ADD;
                    if (odd || f_is_zero(d)) {
                        const uint4 m = *reinterpret_cast<const uint4 *>(ba_shared + ba_chunk(c0 + 2 * FCH));
                        const ba_cold_t<F> c = ba_classify_cold(ba_px<F, FIRST>(io, m.x & IDX), ba_py<F, FIRST>(io, m.x & IDX), FIRST && (m.x >> 31),
                                                                ba_px<F, FIRST>(io, m.y & IDX), ba_py<F, FIRST>(io, m.y & IDX), FIRST && (m.y >> 31));
                        tag = c.tag;
                        d = c.d;
                    }
                    if (tag <= BA_TAG_DBL) {
                        ba_scratch_st(scratch, scratch_stride, s, run);
                        f_mul(run, run, d);
                    }
                }
            }
            ba_cp_wait<0>();
        }
        BA_STAMP(1);
        // ---- one inversion per lane (branch-free; every lane of the warp takes part) ----
        F inv;
        f_inv_warp(inv, run);
        BA_STAMP(2);
        // ---- backward: slopes and results ----
        {
            // shared-memory plan: descriptor ring of 3 (slots j, j - 1 live; j - 2 arriving), NXS stages of {prefix product,
            // x1, x2} and ONE stage of {y1, y2}: the y coordinates of slot j - 1 are requested in the middle of slot j, right
            // after slot j has taken its own into registers — three multiplications before they are needed. With a single x
            // stage (Fp2) the x part of slot j - 1 is requested at the same point; with two (Fp) at the top of slot j.
            constexpr int ND = SM::BWD_DESC, XCH = SM::BWD_X_CH, NXS = SM::BWD_X_STAGES, X0 = ND, Y0 = ND + NXS * XCH;
            auto xstage = [&](uint32_t j) { return X0 + (int)(j % NXS) * XCH; };
            auto desc_at = [&](uint32_t j) { return *reinterpret_cast<const uint4 *>(ba_shared + ba_chunk((int)(j % ND))); };
            auto issue_desc = [&](uint32_t j) { ba_cp16(sbase + ba_chunk((int)(j % ND)), desc_src(j)); };
            auto issue_x = [&](uint32_t j) {   // needs the descriptor of slot j in the ring
                const uint4 m = desc_at(j);
                const int c0 = xstage(j);
                const size_t s = min(slot_of(j), (size_t)last);
#pragma unroll
                for (int k = 0; k < FCH; k++) ba_cp16(sbase + ba_chunk(c0 + k), scratch + (size_t)k * scratch_stride + s);
                if (FIRST) ba_coop_gather<F, FCH>(io.table, io.tstride16, ba_shared, sbase, (int)(j % ND), 0, c0 + FCH, c0 + 2 * FCH);
                else {
                    ba_cp_field(sbase, c0 + FCH, ba_px<F, FIRST>(io, m.x & IDX));
                    ba_cp_field(sbase, c0 + 2 * FCH, ba_px<F, FIRST>(io, m.y & IDX));
                }
            };
            auto issue_y = [&](uint32_t j) {
                if (FIRST) { ba_coop_gather<F, FCH>(io.table, io.tstride16, ba_shared, sbase, (int)(j % ND), FCH, Y0, Y0 + FCH); return; }
                const uint4 m = desc_at(j);
                ba_cp_field(sbase, Y0, ba_py<F, FIRST>(io, m.x & IDX));
                ba_cp_field(sbase, Y0 + FCH, ba_py<F, FIRST>(io, m.y & IDX));
            };
            issue_desc(niter - 1);
            if (niter >= 2) issue_desc(niter - 2);
            ba_cp_commit();
            ba_cp_wait<0>();
            issue_x(niter - 1);
            issue_y(niter - 1);
            ba_cp_commit();
#pragma unroll 1
            for (uint32_t j = niter; j-- > 0;) {
                ba_cp_wait<0>();            // operands of slot j and the descriptor of slot j - 1 have landed
                if (FIRST) __syncwarp();    // round 0: ... for every lane of the warp (cooperative gather)
                if (NXS == 2 && j >= 1) issue_x(j - 1);
                if (j >= 2) issue_desc(j - 2);
                ba_cp_commit();
                const size_t s = slot_of(j);
                const int c0 = xstage(j);
                const uint4 m = desc_at(j);
                F lam, x1, x2, y1, t, d;
                // The single y stage is refilled for slot j - 1 once slot j has taken its y coordinates into registers.
                if (FIRST) {
                    // Round 0 fills the stages cooperatively: EVERY lane (also one past the end of the round) reads both y
                    // coordinates, the warp synchronises, and only then the refill is issued.
                    const bool live = s < nadds;
                    int tag = BA_TAG_INF;
                    if (live) {
                        ba_ld_stage(t, ba_shared, c0);
                        f_mul(lam, inv, t);            // 1 / d when the slot is an addition or a doubling (unused otherwise)
                        ba_ld_stage(x1, ba_shared, c0 + FCH);
                        ba_ld_stage(x2, ba_shared, c0 + 2 * FCH);
                        f_sub(d, x2, x1);
                        tag = BA_TAG_ADD;
                        if (f_is_zero(x1) || f_is_zero(x2) || f_is_zero(d)) {
                            const ba_cold_t<F> c = ba_classify_cold(ba_px<F, FIRST>(io, m.x & IDX), ba_py<F, FIRST>(io, m.x & IDX), (m.x >> 31) != 0,
                                                                    ba_px<F, FIRST>(io, m.y & IDX), ba_py<F, FIRST>(io, m.y & IDX), (m.y >> 31) != 0);
                            tag = c.tag;
                            d = c.d;
                        }
                        if (tag <= BA_TAG_DBL && j != 0) f_mul(inv, inv, d);
                    }
                    ba_ld_stage(y1, ba_shared, Y0);
                    ba_ld_stage(t, ba_shared, Y0 + FCH);
                    __syncwarp();
                    if (j >= 1) {
                        if (NXS == 1) issue_x(j - 1);
                        issue_y(j - 1);
                    }
                    ba_cp_commit();
                    if (!live) continue;
                    f_cneg(y1, y1, (m.x >> 31) != 0);
                    f_cneg(t, t, (m.y >> 31) != 0);
                    if (tag <= BA_TAG_DBL) {
                        if (tag == BA_TAG_ADD) {
                            f_sub(t, t, y1);
                        } else {
                            f_sqr(t, x1);
                            f_mul3(t, t);
                        }
                        f_mul(lam, lam, t);            // (y2 - y1) / (x2 - x1)   or   3 x1^2 / (2 y1)
                        f_sqr(t, lam);
                        f_sub(t, t, x1);
                        f_sub(t, t, x2);               // x3 = lambda^2 - x1 - x2
                        f_sub(x1, x1, t);
                        f_mul(x1, x1, lam);
                        f_sub(x1, x1, y1);             // y3 = lambda (x1 - x3) - y1
                        ba_store_point(m.z, io, t, x1);
                    } else if (tag == BA_TAG_INF) {
                        f_set_zero(t);
                        ba_store_point(m.z, io, t, t);
                    } else if (tag == BA_TAG_COPY_P) {
                        ba_store_point(m.z, io, x1, y1);
                    } else {
                        ba_store_point(m.z, io, x2, t);
                    }
                    continue;
                }
                // later rounds: every lane copies for itself and reads only what its slot's case needs
                auto next_y = [&]() {       // exactly once per iteration, once the stages have been read (or are not needed)
                    if (j >= 1) {
                        if (NXS == 1) issue_x(j - 1);
                        issue_y(j - 1);
                    }
                    ba_cp_commit();
                };
                if (s >= nadds) { next_y(); continue; }
                ba_ld_stage(t, ba_shared, c0);
                f_mul(lam, inv, t);
                ba_ld_stage(x1, ba_shared, c0 + FCH);
                ba_ld_stage(x2, ba_shared, c0 + 2 * FCH);
                f_sub(d, x2, x1);
                int tag = BA_TAG_ADD;
                if (f_is_zero(x1) || f_is_zero(x2) || f_is_zero(d)) {
                    const ba_cold_t<F> c = ba_classify_cold(ba_px<F, FIRST>(io, m.x & IDX), ba_py<F, FIRST>(io, m.x & IDX), false,
                                                            ba_px<F, FIRST>(io, m.y & IDX), ba_py<F, FIRST>(io, m.y & IDX), false);
                    tag = c.tag;
                    d = c.d;
                }
                if (tag <= BA_TAG_DBL) {
                    if (j != 0) f_mul(inv, inv, d);
                    ba_ld_stage(y1, ba_shared, Y0);
                    if (tag == BA_TAG_ADD) {
                        ba_ld_stage(t, ba_shared, Y0 + FCH);
                        next_y();
                        f_sub(t, t, y1);
                    } else {
                        next_y();
                        f_sqr(t, x1);
                        f_mul3(t, t);
                    }
                    f_mul(lam, lam, t);
                    f_sqr(t, lam);
                    f_sub(t, t, x1);
                    f_sub(t, t, x2);
                    f_sub(x1, x1, t);
                    f_mul(x1, x1, lam);
                    f_sub(x1, x1, y1);
                    ba_store_point(m.z, io, t, x1);
                } else if (tag == BA_TAG_INF) {
                    next_y();
                    f_set_zero(t);
                    ba_store_point(m.z, io, t, t);
                } else {
                    const bool cp = tag == BA_TAG_COPY_P;
                    ba_ld_stage(y1, ba_shared, Y0 + (cp ? 0 : FCH));
                    next_y();
                    ba_store_point(m.z, io, cp ? x1 : x2, y1);
                }
            }
            ba_cp_wait<0>();
        }
        BA_STAMP(3);
    }
    // ---- copies (odd last elements, single-entry buckets): no arithmetic; done last, in the
    // shadow of the blocks that are still adding ----
    for (size_t ci = (size_t)blockIdx.x * BA_THREADS + threadIdx.x; ci < ncopies; ci += (size_t)gridDim.x * BA_THREADS) {
        const uint2 dsc = cdesc[ci];
        F x, y;
        f_ld(x, ba_px<F, FIRST>(io, dsc.x & IDX));
        f_ld(y, ba_py<F, FIRST>(io, dsc.x & IDX));
        if (FIRST) f_cneg(y, y, (dsc.x >> 31) != 0);
        ba_store_point(dsc.y, io, x, y);
    }
}

}  // namespace msmb200
