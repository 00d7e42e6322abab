// Bucket accumulation by BATCH-AFFINE pairwise rounds (sm_100a) — the GPU analogue of the reference's
// POINTonE{1,2}s_accumulate with its HEAD / TAIL macros (src/bulk_addition.c:51-143): affine + affine -> affine, the
// slope denominators of a whole batch inverted together by Montgomery's trick (5M + 1S per addition plus the shared
// inversion, :27), doubling folded into the same batch with denominator 2y (:63-74), P + (-P) and infinity operands
// resolved without arithmetic (:92-104).
//
// Shape of the computation. Every bucket is a list of k points; round r turns it into ceil(k / 2) points by adding
// neighbours (2i, 2i + 1); an odd last element is copied. After ceil(log2(max k)) rounds every bucket is ONE affine
// point (bucket_sum[b]). All additions of a round, over all buckets, are independent "slots".
//   * The slots of a round are enumerated by 8-byte descriptors (input position, output position | FINAL bucket), built
//     before any arithmetic from the bucket histogram alone (ba_totals / ba_scan_* / ba_emit_*): the arithmetic kernel
//     never sees bucket boundaries, only real additions — no lane is spent on padding.
//   * ba_round_kernel: a block owns B x 128 consecutive slots; lane t handles slots j * 128 + t (j < B), so descriptor,
//     point and scratch accesses of a warp are contiguous. Forward pass: denominator d_j, running product, the product
//     BEFORE d_j goes to a global scratch array (the DRAM system is nearly idle in this phase). One inversion PER LANE by
//     the branch-free safegcd of inv.cuh (all lanes run the same instruction stream, mostly on the idle ALU pipe).
//     Backward pass: slopes and results, 4M + 1S; the forward pass costs 1M.
//   * The same engine sums the digit lists of the bucket REDUCTION (src/multi_scalar.c:301-321), whose plan is static.
// Montgomery-form, fully reduced affine results: the bytes equal what any other correct summation order produces once
// normalised.
#pragma once
#include <cstdint>
#include "ec.cuh"
#include "inv.cuh"

namespace msmb200 {

constexpr int BA_RMAX = 32;                   // rounds: counts are < 2^32
constexpr uint32_t BA_FINAL = 0x80000000u;    // output descriptor: bucket_sum[b] instead of the next round's buffer
constexpr uint32_t BA_HEAVY = 2048;           // buckets with more entries are planned by a whole block
constexpr int BA_THREADS = 128;

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
__device__ __forceinline__ void ba_emit_round(uint32_t b, uint32_t c, uint32_t r, const uint32_t *__restrict__ bases, size_t nbs,
                                              const BaRounds &rd, uint2 *__restrict__ adesc, uint2 *__restrict__ cdesc, uint32_t i0, uint32_t istep) {
    const uint32_t k = ba_len(c, r), a = k >> 1, k1 = (k + 1) >> 1;
    const uint32_t eb = bases[(size_t)(3 * r + 2) * nbs + b];
    const uint32_t ob = k1 == 1 ? (BA_FINAL | b) : bases[(size_t)(3 * r + 5) * nbs + b];
    const uint32_t ab = rd.aoff[r] + bases[(size_t)(3 * r) * nbs + b];
    for (uint32_t i = i0; i < a; i += istep) adesc[ab + i] = make_uint2(eb + 2 * i, k1 == 1 ? ob : ob + i);
    if (i0 == 0 && (k & 1u)) cdesc[rd.coff[r] + bases[(size_t)(3 * r + 1) * nbs + b]] = make_uint2(eb + k - 1, ob + a);  // k >= 3 here
}
static __global__ void __launch_bounds__(256) ba_emit_kernel(const uint32_t *__restrict__ count, size_t nb, const uint32_t *__restrict__ bases, size_t nbs,
                                                             BaRounds rd, uint2 *__restrict__ adesc, uint2 *__restrict__ cdesc,
                                                             uint32_t *__restrict__ heavy /* [0] = count, then bucket ids */) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t c = count[b];
    if (c == 0) return;
    if (c == 1) {  // a single entry goes straight to the bucket sum (copy list of round 0)
        cdesc[rd.coff[0] + bases[(size_t)1 * nbs + b]] = make_uint2(bases[(size_t)2 * nbs + b], BA_FINAL | (uint32_t)b);
        return;
    }
    if (c > BA_HEAVY) { heavy[1 + atomicAdd(&heavy[0], 1u)] = (uint32_t)b; return; }
    for (uint32_t r = 0; r < rd.R && ba_len(c, r) >= 2; r++) ba_emit_round((uint32_t)b, c, r, bases, nbs, rd, adesc, cdesc, 0, 1);
}
static __global__ void __launch_bounds__(256) ba_emit_heavy_kernel(const uint32_t *__restrict__ count, const uint32_t *__restrict__ bases, size_t nbs,
                                                                   BaRounds rd, uint2 *__restrict__ adesc, uint2 *__restrict__ cdesc,
                                                                   const uint32_t *__restrict__ heavy) {
    const uint32_t nheavy = heavy[0];
    for (uint32_t hi = blockIdx.x; hi < nheavy; hi += gridDim.x) {
        const uint32_t b = heavy[1 + hi], c = count[b];
        for (uint32_t r = 0; r < rd.R && ba_len(c, r) >= 2; r++) ba_emit_round(b, c, r, bases, nbs, rd, adesc, cdesc, threadIdx.x, blockDim.x);
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
__device__ __forceinline__ void f_inv_warp(fp2v_t &r, const fp2v_t &a) { f_inv_warp_fp2(r, a); }

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

// where the two operands of a slot live: round 0 reads the precomputation table through the bucket-sorted
// (index | sign << 31) references, later rounds read the previous round's points
template <class F, bool FIRST> struct ba_operands {
    const aff_t<F> *p, *q;
    bool sp, sq;
    __device__ __forceinline__ ba_operands(const aff_t<F> *__restrict__ src, const uint32_t *__restrict__ sorted, uint32_t in, bool two) {
        if (FIRST) {
            const uint32_t v0 = sorted[in], v1 = two ? sorted[in + 1] : v0;
            p = src + (v0 & 0x7fffffffu); sp = (v0 >> 31) != 0;
            q = src + (v1 & 0x7fffffffu); sq = (v1 >> 31) != 0;
        } else {
            p = src + in; q = src + in + (two ? 1 : 0);
            sp = sq = false;
        }
    }
};
// rare operand patterns (an infinity operand, equal x): decided from the full points, identically in both passes
template <class F>
static __device__ __noinline__ int ba_classify_cold(const aff_t<F> *p, bool sp, const aff_t<F> *q, bool sq, F &d) {
    F x1, y1, x2, y2;
    f_ld(x1, &p->x); f_ld(y1, &p->y); f_ld(x2, &q->x); f_ld(y2, &q->y);
    f_cneg(y1, y1, sp);
    f_cneg(y2, y2, sq);
    if (f_is_zero(x1) && f_is_zero(y1)) return BA_TAG_COPY_Q;
    if (f_is_zero(x2) && f_is_zero(y2)) return BA_TAG_COPY_P;
    f_sub(d, x2, x1);
    if (!f_is_zero(d)) return BA_TAG_ADD;
    if (f_eq(y1, y2) && !f_is_zero(y1)) { f_dbl(d, y1); return BA_TAG_DBL; }
    return BA_TAG_INF;
}

template <class F> __device__ __forceinline__ void ba_store_point(uint32_t out, aff_t<F> *__restrict__ pts_out, aff_t<F> *__restrict__ bucket_sum,
                                                                  const F &x, const F &y) {
    aff_t<F> *dst = (out & BA_FINAL) ? bucket_sum + (out & ~BA_FINAL) : pts_out + out;
    f_st(&dst->x, x);
    f_st(&dst->y, y);
}

template <class F, bool FIRST>
static __global__ void __launch_bounds__(BA_THREADS) ba_round_kernel(const aff_t<F> *__restrict__ src, const uint32_t *__restrict__ sorted,
                                                                     const uint2 *__restrict__ adesc, uint32_t nadds,
                                                                     const uint2 *__restrict__ cdesc, uint32_t ncopies,
                                                                     aff_t<F> *__restrict__ pts_out, aff_t<F> *__restrict__ bucket_sum,
                                                                     uint4 *__restrict__ scratch, size_t scratch_stride, uint32_t B) {
    // ---- copies (odd last elements, single-entry buckets): no arithmetic ----
    for (size_t ci = (size_t)blockIdx.x * BA_THREADS + threadIdx.x; ci < ncopies; ci += (size_t)gridDim.x * BA_THREADS) {
        const uint2 dsc = cdesc[ci];
        ba_operands<F, FIRST> op(src, sorted, dsc.x, false);
        F x, y;
        f_ld(x, &op.p->x);
        f_ld(y, &op.p->y);
        if (FIRST) f_cneg(y, y, op.sp);
        ba_store_point(dsc.y, pts_out, bucket_sum, x, y);
    }
    const size_t tile0 = (size_t)blockIdx.x * B * BA_THREADS;
    if (tile0 >= nadds) return;
    // ---- forward: denominators and their running product ----
    F run;
    f_set_one(run);
#pragma unroll 1
    for (uint32_t j = 0; j < B; j++) {
        const size_t s = tile0 + (size_t)j * BA_THREADS + threadIdx.x;
        if (s < nadds) {
            const uint2 dsc = adesc[s];
            ba_operands<F, FIRST> op(src, sorted, dsc.x, true);
            F x1, d;
            f_ld(x1, &op.p->x);
            f_ld(d, &op.q->x);
            const bool odd = f_is_zero(x1) || f_is_zero(d);
            f_sub(d, d, x1);
            int tag = BA_TAG_ADD;
            if (odd || f_is_zero(d)) tag = ba_classify_cold(op.p, op.sp, op.q, op.sq, d);
            if (tag <= BA_TAG_DBL) {
                ba_scratch_st(scratch, scratch_stride, s, run);
                f_mul(run, run, d);
            }
        }
    }
    // ---- one inversion per lane (branch-free; every lane of the warp takes part) ----
    F inv;
    f_inv_warp(inv, run);
    // ---- backward: slopes and results ----
#pragma unroll 1
    for (uint32_t j = B; j-- > 0;) {
        const size_t s = tile0 + (size_t)j * BA_THREADS + threadIdx.x;
        if (s >= nadds) continue;
        const uint2 dsc = adesc[s];
        ba_operands<F, FIRST> op(src, sorted, dsc.x, true);
        F x1, x2, d;
        f_ld(x1, &op.p->x);
        f_ld(x2, &op.q->x);
        const bool odd = f_is_zero(x1) || f_is_zero(x2);
        f_sub(d, x2, x1);
        int tag = BA_TAG_ADD;
        if (odd || f_is_zero(d)) tag = ba_classify_cold(op.p, op.sp, op.q, op.sq, d);
        F y1, lam, t;
        if (tag <= BA_TAG_DBL) {
            F pre;
            ba_scratch_ld(pre, scratch, scratch_stride, s);
            f_mul(lam, inv, pre);          // 1 / d
            if (j != 0) f_mul(inv, inv, d);
            f_ld(y1, &op.p->y);
            if (FIRST) f_cneg(y1, y1, op.sp);
            if (tag == BA_TAG_ADD) {
                f_ld(t, &op.q->y);
                if (FIRST) f_cneg(t, t, op.sq);
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
            ba_store_point(dsc.y, pts_out, bucket_sum, t, x1);
        } else if (tag == BA_TAG_INF) {
            f_set_zero(t);
            ba_store_point(dsc.y, pts_out, bucket_sum, t, t);
        } else {
            const aff_t<F> *c = tag == BA_TAG_COPY_P ? op.p : op.q;
            const bool sc = tag == BA_TAG_COPY_P ? op.sp : op.sq;
            f_ld(t, &c->x);
            f_ld(y1, &c->y);
            if (FIRST) f_cneg(y1, y1, sc);
            ba_store_point(dsc.y, pts_out, bucket_sum, t, y1);
        }
    }
}

}  // namespace msmb200
