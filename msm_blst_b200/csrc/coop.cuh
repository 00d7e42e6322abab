// Lane-cooperative XYZZ arithmetic for the low-parallelism stages of the pipeline (sm_100a).
//
// Why: one Montgomery multiplication is ~300 IMAD.WIDE on the heavy FMA pipe, which issues one warp instruction per
// 4 cycles per SM sub-partition whether 1 or 32 lanes are active. A lone thread therefore needs >= 1200 cycles per
// multiplication, and the tails of the bucket reduction (list sums over a few thousand points, Horner over bit
// positions, the sum of the per-GPU partials) are chains of such lone additions. An XYZZ addition (12M+2S,
// src/ec_ops.h:642-702) has only 4 dependent multiplication LEVELS and a doubling (6M+3S) has 3, so a GROUP of lanes
// computes one level per multiplication time.
//
// Layout: the point is DISTRIBUTED over the group — coordinate j of {x, y, zzz, zz} (the memory order of
// blst_p1xyzz, bindings/blst.h:251) lives in "slot" j of the group, so a group loads/stores one point with coalesced
// accesses and each lane keeps only a handful of Fp elements in registers (no local memory).
//   G1 (Fp) : a slot is ONE lane, a group is a QUAD (8 points per warp).
//   G2 (Fp2): a slot is a lane PAIR — the even lane holds the real halves, the odd lane the imaginary halves
//             (fp2h_t); an Fp2 multiplication is two rounds of one Fp multiplication per lane
//             ((a0 b0, a1 b1), then (a0 b1, a1 b0)) with three 12-word exchanges inside the pair. A group is 8 lanes.
// Operands that live in another slot are fetched with group-width shuffles (7 exchanges per addition, 3 per
// doubling); results land in their owner slots. Same formulas as xyzz_add / xyzz_double (ec.cuh), hence the same
// fully reduced field values, bit for bit.
//
// All 32 lanes of a warp must call these functions together (full-mask shuffles / ballots, warp-uniform control
// flow); the special cases of the reference (infinity, equal points, opposite points) are resolved by selects.
#pragma once
#include "ec.cuh"

namespace msmb200 {

// ---- the half of an Fp2 element held by one lane of a pair (even lane: c0, odd lane: c1) ----
struct fp2h_t { fp_t h; };

template <class C> struct coop_traits;                       // LPS = lanes per slot, field = element type in memory
template <> struct coop_traits<fp_t> { static constexpr int LPS = 1; using field = fp_t; };
template <> struct coop_traits<fpc_t> { static constexpr int LPS = 1; using field = fpc_t; };
template <> struct coop_traits<fp2h_t> { static constexpr int LPS = 2; using field = fp2_t; };
template <class F> struct coop_of { using type = F; };       // memory field type -> lane type
template <> struct coop_of<fp2_t> { using type = fp2h_t; };

__device__ __forceinline__ void fp_shfl(fp_t &r, const fp_t &a, int src, int width) {
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = __shfl_sync(0xffffffffu, a.l[k], src, width);
}
__device__ __forceinline__ void fp_shfl_xor1(fp_t &r, const fp_t &a) {
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = __shfl_xor_sync(0xffffffffu, a.l[k], 1);
}

// Fp2 arithmetic on pair-distributed elements (spec: src/no_asm.h:566-579)
__device__ __forceinline__ void f_add(fp2h_t &r, const fp2h_t &a, const fp2h_t &b) { fp_add(r.h, a.h, b.h); }
__device__ __forceinline__ void f_sub(fp2h_t &r, const fp2h_t &a, const fp2h_t &b) { fp_sub(r.h, a.h, b.h); }
__device__ __forceinline__ void f_dbl(fp2h_t &r, const fp2h_t &a) { fp_dbl(r.h, a.h); }
__device__ __forceinline__ void f_mul3(fp2h_t &r, const fp2h_t &a) { fp_mul3(r.h, a.h); }
__device__ __forceinline__ void f_set_zero(fp2h_t &r) { fp_set_zero(r.h); }
__device__ __forceinline__ bool f_is_zero(const fp2h_t &a) {  // both halves zero; same answer in both lanes
    const int z = fp_is_zero(a.h) ? 1 : 0;
    const int zo = __shfl_xor_sync(0xffffffffu, z, 1);  // unconditional: every lane takes part in the exchange
    return (z & zo) != 0;
}
// (a0 + a1 i)(b0 + b1 i): even lane ends with a0 b0 - a1 b1, odd lane with a0 b1 + a1 b0
__device__ __forceinline__ void f_mul(fp2h_t &r, const fp2h_t &a, const fp2h_t &b) {
    const bool odd = threadIdx.x & 1;
    fp_t bo, p1, p2, t;
    fp_shfl_xor1(bo, b.h);             // even: b1, odd: b0
    fp_mul(p1, a.h, b.h);              // even: a0 b0, odd: a1 b1
    fp_mul(p2, a.h, bo);               // even: a0 b1, odd: a1 b0
    // even needs the partner's p1 (a1 b1), odd needs the partner's p2 (a0 b1): one exchange
    fp_t pub;
#pragma unroll
    for (int k = 0; k < 12; k++) pub.l[k] = odd ? p1.l[k] : p2.l[k];
    fp_shfl_xor1(t, pub);              // even: a1 b1, odd: a0 b1
    fp_t s, d;
    fp_sub(d, p1, t);                  // even: a0 b0 - a1 b1
    fp_add(s, p2, t);                  // odd:  a1 b0 + a0 b1
#pragma unroll
    for (int k = 0; k < 12; k++) r.h.l[k] = odd ? s.l[k] : d.l[k];
}
// (a0 + a1)(a0 - a1) + 2 a0 a1 i: one multiplication round
__device__ __forceinline__ void f_sqr(fp2h_t &r, const fp2h_t &a) {
    const bool odd = threadIdx.x & 1;
    fp_t o, s, d, x, y, p;
    fp_shfl_xor1(o, a.h);              // partner's half
    fp_add(s, a.h, o);                 // a0 + a1
    fp_sub(d, a.h, o);                 // even: a0 - a1
#pragma unroll
    for (int k = 0; k < 12; k++) { x.l[k] = odd ? a.h.l[k] : s.l[k]; y.l[k] = odd ? o.l[k] : d.l[k]; }
    fp_mul(p, x, y);                   // even: (a0 + a1)(a0 - a1), odd: a1 a0
    fp_dbl(s, p);
#pragma unroll
    for (int k = 0; k < 12; k++) r.h.l[k] = odd ? s.l[k] : p.l[k];
}

// ---- group geometry ----
template <class C> __device__ __forceinline__ int coop_slot() { return (threadIdx.x / coop_traits<C>::LPS) & 3; }
template <class C> __device__ __forceinline__ int coop_group_base() { return threadIdx.x & 31 & ~(4 * coop_traits<C>::LPS - 1); }
template <class C> __host__ __device__ constexpr int coop_group_lanes() { return 4 * coop_traits<C>::LPS; }

// every lane publishes `pub`; a lane reads what the same sub-lane of slot `src` of its own group published
__device__ __forceinline__ void fq_read(fp_t &r, const fp_t &pub, int src) { fp_shfl(r, pub, src, 4); }
__device__ __forceinline__ void fq_read(fpc_t &r, const fpc_t &pub, int src) { fp_shfl((fp_t &)r, (const fp_t &)pub, src, 4); }
__device__ __forceinline__ void fq_read(fp2h_t &r, const fp2h_t &pub, int src) { fp_shfl(r.h, pub.h, 2 * src + (threadIdx.x & 1), 8); }

__device__ __forceinline__ void fq_sel(fp_t &r, bool c, const fp_t &a, const fp_t &b) {  // r = c ? a : b
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = c ? a.l[k] : b.l[k];
}
__device__ __forceinline__ void fq_sel(fpc_t &r, bool c, const fpc_t &a, const fpc_t &b) { fq_sel((fp_t &)r, c, (const fp_t &)a, (const fp_t &)b); }
__device__ __forceinline__ void fq_sel(fp2h_t &r, bool c, const fp2h_t &a, const fp2h_t &b) { fq_sel(r.h, c, a.h, b.h); }

// point <-> group: slot j holds coordinate j of {x, y, zzz, zz}
template <class C> __device__ __forceinline__ void dq_load(C &c, const xyzz_t<typename coop_traits<C>::field> *p) {
    // 48-byte piece index: coordinate slot (and, for Fp2, the half held by this lane)
    const int piece = coop_traits<C>::LPS == 1 ? (int)(threadIdx.x & 3) : (int)(((threadIdx.x >> 1) & 3) * 2 + (threadIdx.x & 1));
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const fp_t *>(p) + piece);
    uint4 *dst = reinterpret_cast<uint4 *>(&c);
#pragma unroll
    for (int k = 0; k < 3; k++) dst[k] = src[k];
}
template <class C> __device__ __forceinline__ void dq_store(xyzz_t<typename coop_traits<C>::field> *p, const C &c) {
    const int piece = coop_traits<C>::LPS == 1 ? (int)(threadIdx.x & 3) : (int)(((threadIdx.x >> 1) & 3) * 2 + (threadIdx.x & 1));
    uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<fp_t *>(p) + piece);
    const uint4 *src = reinterpret_cast<const uint4 *>(&c);
#pragma unroll
    for (int k = 0; k < 3; k++) dst[k] = src[k];
}
// bits (zzz zero, zz zero) of this lane's group out of a ballot of f_is_zero(coordinate)
template <class C> __device__ __forceinline__ unsigned dq_zbits(unsigned ballot) {
    constexpr int LPS = coop_traits<C>::LPS;
    const int gb = coop_group_base<C>();
    return ((ballot >> (gb + 2 * LPS)) & 1u) | (((ballot >> (gb + 3 * LPS)) & 1u) << 1);
}
template <class C> __device__ __forceinline__ bool dq_is_inf(const C &c) {
    return dq_zbits<C>(__ballot_sync(0xffffffffu, f_is_zero(c))) == 3u;
}

// r = 2p (dbl-2008-s-1, the formulas of xyzz_double): 3 multiplication levels, 3 exchanges. p infinite -> r infinite.
template <class C> __device__ __forceinline__ void dq_double(C &r, const C &p) {
    const int l = coop_slot<C>();
    C U, a, b, t, p1, p2, p3, M;
    f_dbl(U, p);                       // slot 1: U = 2y
    fq_sel(a, l == 1, U, p);
    f_sqr(p1, a);                      // slot 0: XX = x^2, slot 1: V = U^2
    f_mul3(M, p1);                     // slot 0: M = 3 XX
    fq_sel(a, l == 0, M, p1);
    fq_read(t, a, l == 2 ? 0 : 1);     // slots 0, 1, 3: V; slot 2: M
    fq_sel(a, l == 1, U, p);
    fq_sel(a, l == 2, t, a);
    f_mul(p2, a, t);                   // slot 0: S = x V, slot 1: W = U V, slot 2: MM = M^2, slot 3: ZZ3 = zz V
    fq_read(t, p2, l == 0 ? 2 : 1);    // slot 0: MM; slot 2: W
    C d, X3, Sm;
    f_dbl(d, p2);
    f_sub(X3, t, d);                   // slot 0: X3 = MM - 2S
    f_sub(Sm, p2, X3);                 // slot 0: S - X3
    fq_sel(a, l == 0, M, p);
    fq_sel(a, l == 1, p2, a);          // slot 0: M, slot 1: W, slot 2: zzz
    fq_sel(b, l == 0, Sm, p);
    fq_sel(b, l == 2, t, b);           // slot 0: S - X3, slot 1: y, slot 2: W
    f_mul(p3, a, b);                   // slot 0: T = M (S - X3), slot 1: W y, slot 2: ZZZ3 = zzz W
    fq_read(t, p3, 0);                 // slot 1: T
    f_sub(d, t, p3);                   // slot 1: Y3 = T - W y
    fq_sel(r, l == 0, X3, p3);
    fq_sel(r, l == 1, d, r);
    fq_sel(r, l == 3, p2, r);
}

// acc += q (add-2008-s, the formulas and case analysis of xyzz_add): 4 multiplication levels, 7 exchanges
// (+ a doubling when some group of the warp adds a point to itself)
template <class C> __device__ __forceinline__ void dq_add(C &acc, const C &q) {
    const int l = coop_slot<C>();
    const unsigned za = dq_zbits<C>(__ballot_sync(0xffffffffu, f_is_zero(acc))), zq = dq_zbits<C>(__ballot_sync(0xffffffffu, f_is_zero(q)));
    const bool a_inf = za == 3u, q_inf = zq == 3u;
    C t, p1, D, a, b, p2, tP, t4, p3, t5, W1, p4;
    fq_read(t, q, 3 - l);              // slot 0: q.zz, slot 1: q.zzz, slot 2: q.y, slot 3: q.x
    f_mul(p1, acc, t);                 // slot 0: U = x q.zz, slot 1: S = y q.zzz, slot 2: q.y zzz, slot 3: q.x zz
    fq_read(t, p1, 3 - l);             // slot 2: S, slot 3: U
    f_sub(D, p1, t);                   // slot 2: R, slot 3: P
    const unsigned zd = dq_zbits<C>(__ballot_sync(0xffffffffu, f_is_zero(D)));
    const bool r_zero = zd & 1u, p_zero = (zd >> 1) & 1u;
    fq_read(tP, D, l == 1 ? 2 : 3);    // slot 0: P, slot 1: R, slot 2: P
    fq_sel(a, l < 2, tP, acc);
    fq_sel(b, l < 2, tP, q);
    f_mul(p2, a, b);                   // slot 0: PP, slot 1: RR, slot 2: B = zzz q.zzz, slot 3: A = zz q.zz
    fq_sel(a, l == 2, D, p2);
    fq_read(t4, a, l == 0 ? 2 : 0);    // slot 0: R; slots 2, 3: PP
    fq_sel(a, l == 0, p1, p2);
    fq_sel(a, l == 2, tP, a);          // slot 0: U, slot 2: P, slot 3: A
    fq_sel(b, l == 0, p2, t4);         // slot 0: PP, slots 2, 3: PP
    f_mul(p3, a, b);                   // slot 0: Q = U PP, slot 2: PPP = P PP, slot 3: ZZ3 = A PP
    fq_read(t5, p3, 2);                // PPP (used by slot 1)
    f_sub(W1, p2, t5);                 // slot 1: RR - PPP
    fq_read(t, W1, 1);                 // slot 0: RR - PPP
    C d, X3, Qm;
    f_dbl(d, p3);
    f_sub(X3, t, d);                   // slot 0: X3 = RR - PPP - 2Q
    f_sub(Qm, p3, X3);                 // slot 0: Q - X3
    fq_sel(a, l == 0, Qm, p2);
    fq_sel(a, l == 1, p1, a);          // slot 0: Q - X3, slot 1: S, slot 2: B
    fq_sel(b, l == 0, t4, p3);
    fq_sel(b, l == 1, t5, b);          // slot 0: R, slot 1: PPP, slot 2: PPP
    f_mul(p4, a, b);                   // slot 0: T = (Q - X3) R, slot 1: S PPP, slot 2: ZZZ3 = B PPP
    fq_read(t, p4, 0);                 // slot 1: T
    f_sub(d, t, p4);                   // slot 1: Y3 = T - S PPP
    C res;
    fq_sel(res, l == 0, X3, p4);
    fq_sel(res, l == 1, d, res);
    fq_sel(res, l == 3, p3, res);
    const bool finite = !a_inf && !q_inf;
    const bool need_dbl = finite && p_zero && r_zero;
    if (__any_sync(0xffffffffu, need_dbl)) {
        C dd;
        dq_double(dd, acc);
        fq_sel(res, need_dbl, dd, res);
    }
    if (finite && p_zero && !r_zero) f_set_zero(res);
    if (q_inf) res = acc;
    else if (a_inf) res = q;
    acc = res;
}

// move a distributed point between groups of a warp: lane i reads lane i + o (o a multiple of the group size)
__device__ __forceinline__ void dq_shfl_down(fp_t &r, const fp_t &a, int o) {
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = __shfl_down_sync(0xffffffffu, a.l[k], o);
}
__device__ __forceinline__ void dq_shfl_down(fpc_t &r, const fpc_t &a, int o) { dq_shfl_down((fp_t &)r, (const fp_t &)a, o); }
__device__ __forceinline__ void dq_shfl_down(fp2h_t &r, const fp2h_t &a, int o) { dq_shfl_down(r.h, a.h, o); }

}  // namespace msmb200
