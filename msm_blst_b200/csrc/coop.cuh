// Quad-cooperative XYZZ arithmetic for the serial tails of the pipeline (sm_100a).
//
// Why: one Montgomery multiplication is ~300 IMAD.WIDE on the heavy FMA pipe, which issues one warp instruction per
// 4 cycles per SM sub-partition whether 1 or 32 lanes are active. A lone thread therefore needs ~1200 cycles per
// multiplication, and the tails of the bucket reduction (stage-2 lists, Horner over bit positions, the sum of the
// per-GPU partials) are chains of such lone additions. An XYZZ addition (12M+2S, src/ec_ops.h:642-702) has only 4
// dependent multiplication LEVELS and a doubling (6M+3S) has 3, so a QUAD of 4 adjacent lanes that all hold the same
// operands computes one level per multiplication time: each lane multiplies one operand pair, then the four products
// are exchanged with width-4 shuffles. Same formulas, same fully reduced field results, bit for bit.
//
// All 32 lanes of the warp must call these functions together (full-mask shuffles, warp-uniform control flow); the
// special cases of the reference (infinity, equal points, opposite points) are resolved by selects afterwards.
#pragma once
#include "ec.cuh"

namespace msmb200 {

__device__ __forceinline__ void fq_gather(fp_t &r, const fp_t &a, int src) {
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = __shfl_sync(0xffffffffu, a.l[k], src, 4);
}
__device__ __forceinline__ void fq_gather(fpc_t &r, const fpc_t &a, int src) { fq_gather((fp_t &)r, (const fp_t &)a, src); }
__device__ __forceinline__ void fq_gather(fp2_t &r, const fp2_t &a, int src) { fq_gather(r.c0, a.c0, src); fq_gather(r.c1, a.c1, src); }

__device__ __forceinline__ void fq_select(fp_t &r, int q, const fp_t &a0, const fp_t &a1, const fp_t &a2, const fp_t &a3) {
#pragma unroll
    for (int k = 0; k < 12; k++) {
        uint32_t lo = (q & 1) ? a1.l[k] : a0.l[k], hi = (q & 1) ? a3.l[k] : a2.l[k];
        r.l[k] = (q & 2) ? hi : lo;
    }
}
__device__ __forceinline__ void fq_select(fpc_t &r, int q, const fpc_t &a0, const fpc_t &a1, const fpc_t &a2, const fpc_t &a3) {
    fq_select((fp_t &)r, q, (const fp_t &)a0, (const fp_t &)a1, (const fp_t &)a2, (const fp_t &)a3);
}
__device__ __forceinline__ void fq_select(fp2_t &r, int q, const fp2_t &a0, const fp2_t &a1, const fp2_t &a2, const fp2_t &a3) {
    fq_select(r.c0, q, a0.c0, a1.c0, a2.c0, a3.c0);
    fq_select(r.c1, q, a0.c1, a1.c1, a2.c1, a3.c1);
}

// r_j = a_j * b_j for j = 0..3, lane (threadIdx.x & 3) == j computing product j; every lane of the quad gets all four
template <class F>
__device__ __noinline__ void quad_mul4(F &r0, F &r1, F &r2, F &r3, const F &a0, const F &b0, const F &a1, const F &b1, const F &a2,
                                       const F &b2, const F &a3, const F &b3) {
    const int q = threadIdx.x & 3;
    F a, b, r;
    fq_select(a, q, a0, a1, a2, a3);
    fq_select(b, q, b0, b1, b2, b3);
    f_mul(r, a, b);
    fq_gather(r0, r, 0);
    fq_gather(r1, r, 1);
    fq_gather(r2, r, 2);
    fq_gather(r3, r, 3);
}

// r = 2p (dbl-2008-s-1, same operation order as xyzz_double): 3 multiplication levels. p infinite -> r infinite.
template <class F> __device__ __forceinline__ void quad_xyzz_double(xyzz_t<F> &r, const xyzz_t<F> &p) {
    F U, V, XX, W, S, M, MM, T, Wy, t, d0, d1;
    xyzz_t<F> o;
    f_dbl(U, p.y);
    quad_mul4(V, XX, d0, d1, U, U, p.x, p.x, U, U, p.x, p.x);
    f_mul3(M, XX);
    quad_mul4(W, S, MM, o.zz, V, U, p.x, V, M, M, p.zz, V);
    f_dbl(t, S);
    f_sub(o.x, MM, t);
    f_sub(S, S, o.x);
    quad_mul4(T, Wy, o.zzz, d0, S, M, W, p.y, p.zzz, W, S, M);
    f_sub(o.y, T, Wy);
    r = o;
}

// acc += q (both XYZZ, add-2008-s with the reference's case analysis): 4 multiplication levels (+3 when some quad of
// the warp hits the doubling case)
template <class F> __device__ __forceinline__ void quad_xyzz_add(xyzz_t<F> &acc, const xyzz_t<F> &q) {
    const bool q_inf = xyzz_is_inf(q), a_inf = xyzz_is_inf(acc);
    F U, S, P, R, PP, RR, A, B, PPP, Q, T, Y1, t, d0;
    xyzz_t<F> o;
    quad_mul4(U, S, P, R, acc.x, q.zz, acc.y, q.zzz, q.x, acc.zz, q.y, acc.zzz);
    f_sub(P, P, U);
    f_sub(R, R, S);
    const bool p_zero = f_is_zero(P), r_zero = f_is_zero(R);
    quad_mul4(PP, RR, A, B, P, P, R, R, acc.zz, q.zz, acc.zzz, q.zzz);
    quad_mul4(PPP, Q, o.zz, d0, PP, P, U, PP, A, PP, A, PP);
    f_dbl(t, Q);
    f_sub(o.x, RR, PPP);
    f_sub(o.x, o.x, t);
    f_sub(Q, Q, o.x);
    quad_mul4(T, Y1, o.zzz, d0, Q, R, S, PPP, B, PPP, B, PPP);
    f_sub(o.y, T, Y1);
    const bool finite = !q_inf && !a_inf;
    const bool need_dbl = finite && p_zero && r_zero;
    if (__any_sync(0xffffffffu, need_dbl)) {
        xyzz_t<F> d;
        quad_xyzz_double(d, acc);
        if (need_dbl) o = d;
    }
    if (finite && p_zero && !r_zero) xyzz_set_inf(o);
    if (q_inf) o = acc;
    else if (a_inf) o = q;
    acc = o;
}

// exchange of whole points between quads of a warp (every lane of the source quad holds the same value)
template <class F> __device__ __forceinline__ void f_shfl_down_any(F &r, const F &a, int o);
template <> __device__ __forceinline__ void f_shfl_down_any<fp_t>(fp_t &r, const fp_t &a, int o) {
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = __shfl_down_sync(0xffffffffu, a.l[k], o);
}
template <> __device__ __forceinline__ void f_shfl_down_any<fpc_t>(fpc_t &r, const fpc_t &a, int o) {
    f_shfl_down_any<fp_t>((fp_t &)r, (const fp_t &)a, o);
}
template <> __device__ __forceinline__ void f_shfl_down_any<fp2_t>(fp2_t &r, const fp2_t &a, int o) {
    f_shfl_down_any<fp_t>(r.c0, a.c0, o);
    f_shfl_down_any<fp_t>(r.c1, a.c1, o);
}
template <class F> __device__ __forceinline__ void xyzz_shfl_down_any(xyzz_t<F> &r, const xyzz_t<F> &a, int o) {
    f_shfl_down_any<F>(r.x, a.x, o);
    f_shfl_down_any<F>(r.y, a.y, o);
    f_shfl_down_any<F>(r.zzz, a.zzz, o);
    f_shfl_down_any<F>(r.zz, a.zz, o);
}

}  // namespace msmb200
