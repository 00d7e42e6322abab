// Curve arithmetic for BLS12-381 G1 (F = fp_t) and G2 (F = fp2_t) on sm_100a: affine, Jacobian and XYZZ
// points, every special case (infinity, doubling, cancellation) handled because the synthetic fixed points
// P_i = 2^(i+1) G make P+P and P-P really occur inside buckets.
//
// Replaces: POINTonE{1,2}xyzz_dadd_affine  reference src/ec_ops.h:710-769   (8M+2S)
//           POINTonE{1,2}xyzz_dadd         reference src/ec_ops.h:642-702   (12M+2S)
//           POINTonE{1,2}xyzz_to_Jacobian  reference src/ec_ops.h:771-777
//           POINTonE{1,2}_dadd / _double   reference src/ec_ops.h:40-100,:299-327
//           POINTonE{1,2}_from_Jacobian / to_affine   reference src/e1.c:60-92, src/e2.c:97-128
// Memory layouts equal the reference structs (bindings/blst.h:164-165,:191-192,:251-252): Montgomery limbs, LE.
#pragma once
#include "fp2.cuh"

namespace msmb200 {

template <class F> struct aff_t { F x, y; };
template <class F> struct jac_t { F x, y, z; };
template <class F> struct xyzz_t { F x, y, zzz, zz; };

template <class F> __device__ __forceinline__ bool aff_is_inf(const aff_t<F> &p) { return f_is_zero(p.x) && f_is_zero(p.y); }
template <class F> __device__ __forceinline__ bool xyzz_is_inf(const xyzz_t<F> &p) { return f_is_zero(p.zz) && f_is_zero(p.zzz); }
template <class F> __device__ __forceinline__ void xyzz_set_inf(xyzz_t<F> &p) { f_set_zero(p.x); f_set_zero(p.y); f_set_zero(p.zzz); f_set_zero(p.zz); }
template <class F> __device__ __forceinline__ void jac_set_inf(jac_t<F> &p) { f_set_zero(p.x); f_set_zero(p.y); f_set_zero(p.z); }

// 2*(x,y) affine -> XYZZ (mdbl-2008-s-1); sign folded into ZZZ like the reference (ec_ops.h:749-766)
template <class F> __device__ __noinline__ void xyzz_double_affine(xyzz_t<F> &r, const F &x, const F &y, bool subtract) {
    F U, S, M, t;
    f_dbl(U, y);
    f_sqr(r.zz, U);
    f_mul(r.zzz, r.zz, U);
    f_mul(S, x, r.zz);
    f_sqr(M, x);
    f_mul3(M, M);
    f_sqr(r.x, M);
    f_dbl(t, S);
    f_sub(r.x, r.x, t);
    f_mul(r.y, r.zzz, y);
    f_sub(S, S, r.x);
    f_mul(S, S, M);
    f_sub(r.y, S, r.y);
    f_cneg(r.zzz, r.zzz, subtract);
}
// 2*p, p in XYZZ (dbl-2008-s-1)
template <class F> __device__ __noinline__ void xyzz_double(xyzz_t<F> &r, const xyzz_t<F> &p) {
    F U, V, W, S, M, t;
    xyzz_t<F> o;
    f_dbl(U, p.y);
    f_sqr(V, U);
    f_mul(W, V, U);
    f_mul(S, p.x, V);
    f_sqr(M, p.x);
    f_mul3(M, M);
    f_sqr(o.x, M);
    f_dbl(t, S);
    f_sub(o.x, o.x, t);
    f_mul(o.y, W, p.y);
    f_sub(S, S, o.x);
    f_mul(S, S, M);
    f_sub(o.y, S, o.y);
    f_mul(o.zz, p.zz, V);
    f_mul(o.zzz, p.zzz, W);
    r = o;
}

// acc += (subtract ? -p : p), p affine.   Hot loop #1 of the reference (src/multi_scalar.c:437-461).
template <class F> __device__ __forceinline__ void xyzz_add_affine(xyzz_t<F> &acc, const aff_t<F> &p, bool subtract) {
    if (aff_is_inf(p)) return;
    if (xyzz_is_inf(acc)) {
        acc.x = p.x;
        acc.y = p.y;
        F one;
        f_set_one(one);
        f_cneg(acc.zzz, one, subtract);
        acc.zz = one;
        return;
    }
    F P, R;
    f_mul(P, p.x, acc.zz);
    f_mul(R, p.y, acc.zzz);
    f_cneg(R, R, subtract);
    f_sub(P, P, acc.x);
    f_sub(R, R, acc.y);
    if (!f_is_zero(P)) {
        F PP, PPP, Q, t;
        f_sqr(PP, P);
        f_mul(PPP, PP, P);
        f_mul(Q, acc.x, PP);
        f_sqr(acc.x, R);
        f_dbl(t, Q);
        f_sub(acc.x, acc.x, PPP);
        f_sub(acc.x, acc.x, t);
        f_sub(Q, Q, acc.x);
        f_mul(Q, Q, R);
        f_mul(acc.y, acc.y, PPP);
        f_sub(acc.y, Q, acc.y);
        f_mul(acc.zz, acc.zz, PP);
        f_mul(acc.zzz, acc.zzz, PPP);
    } else if (f_is_zero(R)) {
        // rare (P + P): run out of line on COPIES so that neither `acc` nor `p` has its address taken —
        // otherwise the compiler keeps the hot accumulator in local memory for the whole loop
        xyzz_t<F> t;
        F px = p.x, py = p.y;
        xyzz_double_affine(t, px, py, subtract);
        acc = t;
    } else {
        xyzz_set_inf(acc);
    }
}

// acc += q, both XYZZ.   Hot loop #2 of the reference (src/multi_scalar.c:301-321).
template <class F> __device__ __forceinline__ void xyzz_add(xyzz_t<F> &acc, const xyzz_t<F> &q) {
    if (xyzz_is_inf(q)) return;
    if (xyzz_is_inf(acc)) { acc = q; return; }
    F U, S, P, R;
    f_mul(U, acc.x, q.zz);
    f_mul(S, acc.y, q.zzz);
    f_mul(P, q.x, acc.zz);
    f_mul(R, q.y, acc.zzz);
    f_sub(P, P, U);
    f_sub(R, R, S);
    if (!f_is_zero(P)) {
        F PP, PPP, Q, t;
        f_sqr(PP, P);
        f_mul(PPP, PP, P);
        f_mul(Q, U, PP);
        f_sqr(acc.x, R);
        f_dbl(t, Q);
        f_sub(acc.x, acc.x, PPP);
        f_sub(acc.x, acc.x, t);
        f_sub(Q, Q, acc.x);
        f_mul(Q, Q, R);
        f_mul(acc.y, S, PPP);
        f_sub(acc.y, Q, acc.y);
        f_mul(acc.zz, acc.zz, q.zz);
        f_mul(acc.zzz, acc.zzz, q.zzz);
        f_mul(acc.zz, acc.zz, PP);
        f_mul(acc.zzz, acc.zzz, PPP);
    } else if (f_is_zero(R)) {
        xyzz_t<F> t, a = acc;  // copies: keep `acc` out of local memory (see xyzz_add_affine)
        xyzz_double(t, a);
        acc = t;
    } else {
        xyzz_set_inf(acc);
    }
}

template <class F> __device__ __forceinline__ void xyzz_to_jac(jac_t<F> &o, const xyzz_t<F> &in) {
    f_mul(o.x, in.x, in.zz);
    f_mul(o.y, in.y, in.zzz);
    o.z = in.zz;
}
template <class F> __device__ __forceinline__ void jac_to_xyzz(xyzz_t<F> &o, const jac_t<F> &in) {
    o.x = in.x;
    o.y = in.y;
    f_sqr(o.zz, in.z);
    f_mul(o.zzz, o.zz, in.z);
}

// Jacobian doubling, a = 0 (dbl-2009-l)
template <class F> __device__ __forceinline__ void jac_double(jac_t<F> &r, const jac_t<F> &p) {
    F A, B, C, t;
    jac_t<F> o;
    f_sqr(A, p.x);
    f_sqr(B, p.y);
    f_sqr(C, B);
    f_add(B, B, p.x);
    f_sqr(B, B);
    f_sub(B, B, A);
    f_sub(B, B, C);
    f_dbl(B, B);
    f_mul3(A, A);
    f_sqr(o.x, A);
    f_sub(o.x, o.x, B);
    f_sub(o.x, o.x, B);
    f_dbl(o.z, p.z);
    f_mul(o.z, o.z, p.y);
    f_dbl(C, C); f_dbl(C, C); f_dbl(C, C);
    f_sub(t, B, o.x);
    f_mul(t, t, A);
    f_sub(o.y, t, C);
    r = o;
}
// Jacobian add-or-double with infinities (same case analysis as reference POINT_DADD_IMPL)
template <class F> __device__ __noinline__ void jac_add(jac_t<F> &r, const jac_t<F> &p1, const jac_t<F> &p2) {
    if (f_is_zero(p2.z)) { r = p1; return; }
    if (f_is_zero(p1.z)) { r = p2; return; }
    F Z1Z1, Z2Z2, U1, U2, S1, S2, H, R;
    f_sqr(Z1Z1, p1.z);
    f_sqr(Z2Z2, p2.z);
    f_mul(U1, p1.x, Z2Z2);
    f_mul(U2, p2.x, Z1Z1);
    f_mul(S1, p1.y, p2.z);
    f_mul(S1, S1, Z2Z2);
    f_mul(S2, p2.y, p1.z);
    f_mul(S2, S2, Z1Z1);
    f_sub(H, U2, U1);
    f_sub(R, S2, S1);
    if (f_is_zero(H)) {
        if (f_is_zero(R)) { jac_double(r, p1); return; }
        jac_set_inf(r);
        return;
    }
    F HH, HHH, V, t;
    jac_t<F> o;
    f_sqr(HH, H);
    f_mul(HHH, HH, H);
    f_mul(V, U1, HH);
    f_sqr(o.x, R);
    f_sub(o.x, o.x, HHH);
    f_dbl(t, V);
    f_sub(o.x, o.x, t);
    f_sub(t, V, o.x);
    f_mul(t, t, R);
    f_mul(S1, S1, HHH);
    f_sub(o.y, t, S1);
    f_mul(o.z, p1.z, p2.z);
    f_mul(o.z, o.z, H);
    r = o;
}

// (X/Z^2, Y/Z^3); infinity -> (0,0). Output limbs fully reduced: canonical Montgomery encoding.
template <class F> __device__ __forceinline__ void jac_to_affine(aff_t<F> &o, const jac_t<F> &p) {
    if (f_is_zero(p.z)) { f_set_zero(o.x); f_set_zero(o.y); return; }
    F zi, zi2;
    f_inv(zi, p.z);
    f_sqr(zi2, zi);
    f_mul(o.x, p.x, zi2);
    f_mul(zi2, zi2, zi);
    f_mul(o.y, p.y, zi2);
}

}  // namespace msmb200
