// Fp2 = Fp[i]/(i^2+1) tower for G2 on sm_100a, built on fp.cuh.
// Replaces mul_fp2 / sqr_fp2 / add_fp2 / sub_fp2 / cneg_fp2 / mul_by_3_fp2 (reference src/fields.h:54-82;
// mul_mont_384x Karatsuba spec src/no_asm.h:566-579; inversion by norm src/recip.c:100-114).
// Also defines the overloaded f_* interface the curve templates (ec.cuh) are written against.
#pragma once
#include "fp.cuh"

namespace msmb200 {

struct __align__(16) fp2_t { fp_t c0, c1; };

// ---- uniform field interface: Fp ----
__device__ __forceinline__ void f_mul(fp_t &r, const fp_t &a, const fp_t &b) { fp_mul(r, a, b); }
__device__ __forceinline__ void f_sqr(fp_t &r, const fp_t &a) { fp_sqr(r, a); }
__device__ __forceinline__ void f_add(fp_t &r, const fp_t &a, const fp_t &b) { fp_add(r, a, b); }
__device__ __forceinline__ void f_sub(fp_t &r, const fp_t &a, const fp_t &b) { fp_sub(r, a, b); }
__device__ __forceinline__ void f_cneg(fp_t &r, const fp_t &a, bool f) { fp_cneg(r, a, f); }
__device__ __forceinline__ void f_dbl(fp_t &r, const fp_t &a) { fp_dbl(r, a); }
__device__ __forceinline__ void f_mul3(fp_t &r, const fp_t &a) { fp_mul3(r, a); }
__device__ __forceinline__ bool f_is_zero(const fp_t &a) { return fp_is_zero(a); }
__device__ __forceinline__ bool f_eq(const fp_t &a, const fp_t &b) { return fp_eq(a, b); }
__device__ __forceinline__ void f_set_zero(fp_t &r) { fp_set_zero(r); }
__device__ __forceinline__ void f_set_one(fp_t &r) { fp_set_one(r); }
__device__ __forceinline__ void f_inv(fp_t &r, const fp_t &a) { fp_inv(r, a); }

// ---- Fp, "cold" flavour: identical layout and results, but multiplication/squaring are out-of-line calls.
// Kernels off the hot path are instantiated with F = fpc_t so that each translation unit contains ONE copy
// of the ~350-instruction Montgomery multiplier instead of one per call site (compile time, I-cache).
struct __align__(16) fpc_t : fp_t {};
__device__ __forceinline__ void fp_mul_call(fp_t &r, const fp_t &a, const fp_t &b) { r = fp_mul_fn(a, b); }
__device__ __forceinline__ void f_mul(fpc_t &r, const fpc_t &a, const fpc_t &b) { fp_mul_call(r, a, b); }
__device__ __forceinline__ void f_sqr(fpc_t &r, const fpc_t &a) { fp_sqr(r, a); }
__device__ __forceinline__ void f_add(fpc_t &r, const fpc_t &a, const fpc_t &b) { fp_add(r, a, b); }
__device__ __forceinline__ void f_sub(fpc_t &r, const fpc_t &a, const fpc_t &b) { fp_sub(r, a, b); }
__device__ __forceinline__ void f_cneg(fpc_t &r, const fpc_t &a, bool f) { fp_cneg(r, a, f); }
__device__ __forceinline__ void f_dbl(fpc_t &r, const fpc_t &a) { fp_dbl(r, a); }
__device__ __forceinline__ void f_mul3(fpc_t &r, const fpc_t &a) { fp_mul3(r, a); }
__device__ __forceinline__ bool f_is_zero(const fpc_t &a) { return fp_is_zero(a); }
__device__ __forceinline__ bool f_eq(const fpc_t &a, const fpc_t &b) { return fp_eq(a, b); }
__device__ __forceinline__ void f_set_zero(fpc_t &r) { fp_set_zero(r); }
__device__ __forceinline__ void f_set_one(fpc_t &r) { fp_set_one(r); }
__device__ __forceinline__ void f_inv(fpc_t &r, const fpc_t &a) { fp_inv(r, a); }

// ---- Fp2 ----
__device__ __forceinline__ void f_add(fp2_t &r, const fp2_t &a, const fp2_t &b) { fp_add(r.c0, a.c0, b.c0); fp_add(r.c1, a.c1, b.c1); }
__device__ __forceinline__ void f_sub(fp2_t &r, const fp2_t &a, const fp2_t &b) { fp_sub(r.c0, a.c0, b.c0); fp_sub(r.c1, a.c1, b.c1); }
__device__ __forceinline__ void f_cneg(fp2_t &r, const fp2_t &a, bool f) { fp_cneg(r.c0, a.c0, f); fp_cneg(r.c1, a.c1, f); }
__device__ __forceinline__ void f_dbl(fp2_t &r, const fp2_t &a) { fp_dbl(r.c0, a.c0); fp_dbl(r.c1, a.c1); }
__device__ __forceinline__ void f_mul3(fp2_t &r, const fp2_t &a) { fp_mul3(r.c0, a.c0); fp_mul3(r.c1, a.c1); }
__device__ __forceinline__ bool f_is_zero(const fp2_t &a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
__device__ __forceinline__ bool f_eq(const fp2_t &a, const fp2_t &b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
__device__ __forceinline__ void f_set_zero(fp2_t &r) { fp_set_zero(r.c0); fp_set_zero(r.c1); }
__device__ __forceinline__ void f_set_one(fp2_t &r) { fp_set_one(r.c0); fp_set_zero(r.c1); }

// (a0 + a1 i)(b0 + b1 i) = (a0b0 - a1b1) + ((a0+a1)(b0+b1) - a0b0 - a1b1) i : 3 Fp multiplications.
// Out of line: a G2 point addition is a sequence of calls (operands live in L1-cached local memory), which
// keeps the register count of the G2 kernels low enough for several warps per scheduler.
static __device__ __noinline__ void f_mul(fp2_t &r, const fp2_t &a, const fp2_t &b) {
    fp_t aa, bb, v0, v1;
    fp_add(aa, a.c0, a.c1);
    fp_add(bb, b.c0, b.c1);
    fp_mul(bb, bb, aa);
    fp_mul(v0, a.c0, b.c0);
    fp_mul(v1, a.c1, b.c1);
    fp_sub(r.c0, v0, v1);
    fp_sub(bb, bb, v0);
    fp_sub(r.c1, bb, v1);
}
// (a0+a1)(a0-a1) + 2 a0 a1 i : 2 Fp multiplications
static __device__ __noinline__ void f_sqr(fp2_t &r, const fp2_t &a) {
    fp_t s, d, m;
    fp_add(s, a.c0, a.c1);
    fp_sub(d, a.c0, a.c1);
    fp_mul(m, a.c0, a.c1);
    fp_mul(r.c0, s, d);
    fp_add(r.c1, m, m);
}
// 1/(a + b i) = (a - b i)/(a^2 + b^2)
static __device__ __noinline__ void f_inv(fp2_t &r, const fp2_t &a) {
    fp_t t0, t1;
    fp_sqr(t0, a.c0);
    fp_sqr(t1, a.c1);
    fp_add(t0, t0, t1);
    fp_inv(t1, t0);
    fp_mul(r.c0, a.c0, t1);
    fp_mul(t0, a.c1, t1);
    fp_neg(r.c1, t0);
}

}  // namespace msmb200
