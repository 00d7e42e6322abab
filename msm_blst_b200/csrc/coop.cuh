// Quad-cooperative XYZZ arithmetic for the low-parallelism stages of the pipeline (sm_100a).
//
// Why: one Montgomery multiplication is ~300 IMAD.WIDE on the heavy FMA pipe, which issues one warp instruction per
// 4 cycles per SM sub-partition whether 1 or 32 lanes are active. A lone thread therefore needs >= 1200 cycles per
// multiplication, and the tails of the bucket reduction (list sums over a few thousand points, Horner over bit
// positions) are chains of such lone additions. An XYZZ addition (12M+2S, src/ec_ops.h:642-702) has only 4 dependent
// multiplication LEVELS and a doubling (6M+3S) has 3, so a QUAD of 4 adjacent lanes computes one level per
// multiplication time.
//
// Layout: the point is DISTRIBUTED over the quad — lane j (threadIdx.x & 3) owns coordinate j of {x, y, zzz, zz}
// (the memory order of blst_p1xyzz, bindings/blst.h:251), so a quad loads/stores one point with four coalesced
// F-sized accesses and each lane keeps only a handful of field elements in registers. Operands that live in another
// lane are fetched with width-4 shuffles (7 exchanges per addition, 3 per doubling); results land in their owner lanes.
// Same formulas as xyzz_add / xyzz_double (ec.cuh), hence the same fully reduced field values, bit for bit.
//
// All 32 lanes of a warp must call these functions together (full-mask shuffles / ballots, warp-uniform control
// flow); the special cases of the reference (infinity, equal points, opposite points) are resolved by selects.
#pragma once
#include "ec.cuh"

namespace msmb200 {

// every lane publishes `pub`; lane reads the value published by lane `src` (0..3) of its own quad
__device__ __forceinline__ void fq_read(fp_t &r, const fp_t &pub, int src) {
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = __shfl_sync(0xffffffffu, pub.l[k], src, 4);
}
__device__ __forceinline__ void fq_read(fpc_t &r, const fpc_t &pub, int src) { fq_read((fp_t &)r, (const fp_t &)pub, src); }
__device__ __forceinline__ void fq_read(fp2_t &r, const fp2_t &pub, int src) { fq_read(r.c0, pub.c0, src); fq_read(r.c1, pub.c1, src); }

__device__ __forceinline__ void fq_sel(fp_t &r, bool c, const fp_t &a, const fp_t &b) {  // r = c ? a : b
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = c ? a.l[k] : b.l[k];
}
__device__ __forceinline__ void fq_sel(fpc_t &r, bool c, const fpc_t &a, const fpc_t &b) { fq_sel((fp_t &)r, c, (const fp_t &)a, (const fp_t &)b); }
__device__ __forceinline__ void fq_sel(fp2_t &r, bool c, const fp2_t &a, const fp2_t &b) { fq_sel(r.c0, c, a.c0, b.c0); fq_sel(r.c1, c, a.c1, b.c1); }

// point <-> quad: lane j holds coordinate j of {x, y, zzz, zz}
template <class F> __device__ __forceinline__ void dq_load(F &c, const xyzz_t<F> *p) {
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const F *>(p) + (threadIdx.x & 3));
    uint4 *dst = reinterpret_cast<uint4 *>(&c);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) dst[k] = src[k];
}
template <class F> __device__ __forceinline__ void dq_store(xyzz_t<F> *p, const F &c) {
    uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<F *>(p) + (threadIdx.x & 3));
    const uint4 *src = reinterpret_cast<const uint4 *>(&c);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) dst[k] = src[k];
}
// infinity <=> zzz == 0 and zz == 0 (lanes 2 and 3 of the quad); same answer in all four lanes
template <class F> __device__ __forceinline__ bool dq_is_inf(const F &c) {
    const unsigned z = __ballot_sync(0xffffffffu, f_is_zero(c));
    return ((z >> ((threadIdx.x & 28) + 2)) & 3u) == 3u;
}

// r = 2p (dbl-2008-s-1, the formulas of xyzz_double): 3 multiplication levels, 3 exchanges. p infinite -> r infinite.
template <class F> __device__ __forceinline__ void dq_double(F &r, const F &p) {
    const int l = threadIdx.x & 3;
    F U, a, b, t, p1, p2, p3, M;
    f_dbl(U, p);                       // lane 1: U = 2y
    fq_sel(a, l == 1, U, p);
    f_sqr(p1, a);                      // lane 0: XX = x^2, lane 1: V = U^2
    f_mul3(M, p1);                     // lane 0: M = 3 XX
    fq_sel(a, l == 0, M, p1);
    fq_read(t, a, l == 2 ? 0 : 1);     // lanes 0, 1, 3: V; lane 2: M
    fq_sel(a, l == 1, U, p);
    fq_sel(a, l == 2, t, a);
    f_mul(p2, a, t);                   // lane 0: S = x V, lane 1: W = U V, lane 2: MM = M^2, lane 3: ZZ3 = zz V
    fq_read(t, p2, l == 0 ? 2 : 1);    // lane 0: MM; lane 2: W
    F d, X3, Sm;
    f_dbl(d, p2);
    f_sub(X3, t, d);                   // lane 0: X3 = MM - 2S
    f_sub(Sm, p2, X3);                 // lane 0: S - X3
    fq_sel(a, l == 0, M, p);
    fq_sel(a, l == 1, p2, a);          // lane 0: M, lane 1: W, lane 2: zzz
    fq_sel(b, l == 0, Sm, p);
    fq_sel(b, l == 2, t, b);           // lane 0: S - X3, lane 1: y, lane 2: W
    f_mul(p3, a, b);                   // lane 0: T = M (S - X3), lane 1: W y, lane 2: ZZZ3 = zzz W
    fq_read(t, p3, 0);                 // lane 1: T
    f_sub(d, t, p3);                   // lane 1: Y3 = T - W y
    fq_sel(r, l == 0, X3, p3);
    fq_sel(r, l == 1, d, r);
    fq_sel(r, l == 3, p2, r);
}

// acc += q (add-2008-s, the formulas and case analysis of xyzz_add): 4 multiplication levels, 7 exchanges
// (+ a doubling when some quad of the warp adds a point to itself)
template <class F> __device__ __forceinline__ void dq_add(F &acc, const F &q) {
    const int l = threadIdx.x & 3;
    const int qb = threadIdx.x & 28;
    const unsigned za = __ballot_sync(0xffffffffu, f_is_zero(acc)), zq = __ballot_sync(0xffffffffu, f_is_zero(q));
    const bool a_inf = ((za >> (qb + 2)) & 3u) == 3u, q_inf = ((zq >> (qb + 2)) & 3u) == 3u;
    F t, p1, D, a, b, p2, tP, t4, p3, t5, W1, p4;
    fq_read(t, q, 3 - l);              // lane 0: q.zz, lane 1: q.zzz, lane 2: q.y, lane 3: q.x
    f_mul(p1, acc, t);                 // lane 0: U = x q.zz, lane 1: S = y q.zzz, lane 2: q.y zzz, lane 3: q.x zz
    fq_read(t, p1, 3 - l);             // lane 2: S, lane 3: U
    f_sub(D, p1, t);                   // lane 2: R, lane 3: P
    const unsigned zd = __ballot_sync(0xffffffffu, f_is_zero(D));
    const bool r_zero = (zd >> (qb + 2)) & 1u, p_zero = (zd >> (qb + 3)) & 1u;
    fq_read(tP, D, l == 1 ? 2 : 3);    // lane 0: P, lane 1: R, lane 2: P
    fq_sel(a, l < 2, tP, acc);
    fq_sel(b, l < 2, tP, q);
    f_mul(p2, a, b);                   // lane 0: PP, lane 1: RR, lane 2: B = zzz q.zzz, lane 3: A = zz q.zz
    fq_sel(a, l == 2, D, p2);
    fq_read(t4, a, l == 0 ? 2 : 0);    // lane 0: R; lanes 2, 3: PP
    fq_sel(a, l == 0, p1, p2);
    fq_sel(a, l == 2, tP, a);          // lane 0: U, lane 2: P, lane 3: A
    fq_sel(b, l == 0, p2, t4);         // lane 0: PP, lanes 2, 3: PP
    f_mul(p3, a, b);                   // lane 0: Q = U PP, lane 2: PPP = P PP, lane 3: ZZ3 = A PP
    fq_read(t5, p3, 2);                // PPP (used by lane 1)
    f_sub(W1, p2, t5);                 // lane 1: RR - PPP
    fq_read(t, W1, 1);                 // lane 0: RR - PPP
    F d, X3, Qm;
    f_dbl(d, p3);
    f_sub(X3, t, d);                   // lane 0: X3 = RR - PPP - 2Q
    f_sub(Qm, p3, X3);                 // lane 0: Q - X3
    fq_sel(a, l == 0, Qm, p2);
    fq_sel(a, l == 1, p1, a);          // lane 0: Q - X3, lane 1: S, lane 2: B
    fq_sel(b, l == 0, t4, p3);
    fq_sel(b, l == 1, t5, b);          // lane 0: R, lane 1: PPP, lane 2: PPP
    f_mul(p4, a, b);                   // lane 0: T = (Q - X3) R, lane 1: S PPP, lane 2: ZZZ3 = B PPP
    fq_read(t, p4, 0);                 // lane 1: T
    f_sub(d, t, p4);                   // lane 1: Y3 = T - S PPP
    F res;
    fq_sel(res, l == 0, X3, p4);
    fq_sel(res, l == 1, d, res);
    fq_sel(res, l == 3, p3, res);
    const bool finite = !a_inf && !q_inf;
    const bool need_dbl = finite && p_zero && r_zero;
    if (__any_sync(0xffffffffu, need_dbl)) {
        F dd;
        dq_double(dd, acc);
        fq_sel(res, need_dbl, dd, res);
    }
    if (finite && p_zero && !r_zero) f_set_zero(res);
    if (q_inf) res = acc;
    else if (a_inf) res = q;
    acc = res;
}

// move a distributed point between quads of a warp: lane i reads lane i + o (o a multiple of 4)
__device__ __forceinline__ void dq_shfl_down(fp_t &r, const fp_t &a, int o) {
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = __shfl_down_sync(0xffffffffu, a.l[k], o);
}
__device__ __forceinline__ void dq_shfl_down(fpc_t &r, const fpc_t &a, int o) { dq_shfl_down((fp_t &)r, (const fp_t &)a, o); }
__device__ __forceinline__ void dq_shfl_down(fp2_t &r, const fp2_t &a, int o) { dq_shfl_down(r.c0, a.c0, o); dq_shfl_down(r.c1, a.c1, o); }

}  // namespace msmb200
