// 384-bit Montgomery field arithmetic for BLS12-381 Fp on sm_100a: 12 x u32 limbs per thread, held in
// registers, multiply-accumulate rows issued as mad.lo.cc/madc.hi.cc pairs that ptxas fuses into
// IMAD.WIDE.U32(.X) carry chains.
//
// Replaces (semantics: result fully reduced, < p, Montgomery radix R = 2^384):
//   mul_fp / sqr_fp   -> mulx_mont_384 / sqrx_mont_384   reference src/fields.h:36-40, spec src/no_asm.h:29-82
//   add_fp / sub_fp   -> add_mod_384 / sub_mod_384       reference src/fields.h:15-19, spec src/no_asm.h:104-161
//   cneg_fp, mul_by_3_fp                                  reference src/fields.h:21,:42
// Constants: reference src/consts.c:10-26, src/consts.h:12-22 (same bytes, read as 32-bit limbs).
#pragma once
#include "inv.cuh"
#include <cstdint>

namespace msmb200 {

struct __align__(16) fp_t { uint32_t l[12]; };

// p = BLS12_381_P as 12 little-endian 32-bit limbs
#define FP_P0 0xffffaaabu
#define FP_P1 0xb9feffffu
#define FP_P2 0xb153ffffu
#define FP_P3 0x1eabfffeu
#define FP_P4 0xf6b0f624u
#define FP_P5 0x6730d2a0u
#define FP_P6 0xf38512bfu
#define FP_P7 0x64774b84u
#define FP_P8 0x434bacd7u
#define FP_P9 0x4b1ba7b6u
#define FP_P10 0x397fe69au
#define FP_P11 0x1a0111eau
#define FP_INV32 0xfffcfffdu  // -1/p mod 2^32 (low half of src/consts.h:12 p0)

__device__ __forceinline__ uint32_t fp_p_limb(int i) {
    switch (i) {
    case 0: return FP_P0; case 1: return FP_P1; case 2: return FP_P2; case 3: return FP_P3;
    case 4: return FP_P4; case 5: return FP_P5; case 6: return FP_P6; case 7: return FP_P7;
    case 8: return FP_P8; case 9: return FP_P9; case 10: return FP_P10; default: return FP_P11;
    }
}
// ONE_MONT_P (src/consts.h:17-22)
__device__ __forceinline__ uint32_t fp_one_limb(int i) {
    switch (i) {
    case 0: return 0x0002fffdu; case 1: return 0x76090000u; case 2: return 0xc40c0002u; case 3: return 0xebf4000bu;
    case 4: return 0x53c758bau; case 5: return 0x5f489857u; case 6: return 0x70525745u; case 7: return 0x77ce5853u;
    case 8: return 0xa256ec6du; case 9: return 0x5c071a97u; case 10: return 0xfa80e493u; default: return 0x15f65ec3u;
    }
}

// ---- single-instruction carry-chain primitives (CC register carried between consecutive asm volatile) ----
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

// ---- basic predicates / moves ----
__device__ __forceinline__ bool fp_is_zero(const fp_t &a) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc |= a.l[i];
    return acc == 0;
}
__device__ __forceinline__ bool fp_eq(const fp_t &a, const fp_t &b) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc |= a.l[i] ^ b.l[i];
    return acc == 0;
}
__device__ __forceinline__ void fp_set_zero(fp_t &r) {
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = 0;
}
__device__ __forceinline__ void fp_set_one(fp_t &r) {
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = fp_one_limb(i);
}

// r = t - p if t >= p else t     (t < 2p)
__device__ __forceinline__ void fp_final_sub(fp_t &r, const uint32_t t[12]) {
    uint32_t u[12];
    u[0] = sub_cc(t[0], fp_p_limb(0));
#pragma unroll
    for (int i = 1; i < 12; i++) u[i] = subc_cc(t[i], fp_p_limb(i));
    uint32_t borrow = subc(0, 0);  // 0xffffffff when t < p
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = borrow ? t[i] : u[i];
}

__device__ __forceinline__ void fp_add(fp_t &r, const fp_t &a, const fp_t &b) {
    uint32_t t[12];
    t[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) t[i] = addc_cc(a.l[i], b.l[i]);
    t[11] = addc(a.l[11], b.l[11]);  // < 2p < 2^382: no carry out
    fp_final_sub(r, t);
}
__device__ __forceinline__ void fp_sub(fp_t &r, const fp_t &a, const fp_t &b) {
    uint32_t t[12];
    t[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 12; i++) t[i] = subc_cc(a.l[i], b.l[i]);
    uint32_t mask = subc(0, 0);  // 0xffffffff on borrow
    r.l[0] = add_cc(t[0], fp_p_limb(0) & mask);
#pragma unroll
    for (int i = 1; i < 11; i++) r.l[i] = addc_cc(t[i], fp_p_limb(i) & mask);
    r.l[11] = addc(t[11], fp_p_limb(11) & mask);
}
// r = flag ? -a : a  (0 stays 0), reference cneg_fp
__device__ __forceinline__ void fp_cneg(fp_t &r, const fp_t &a, bool flag) {
    uint32_t t[12];
    bool z = fp_is_zero(a);
    t[0] = sub_cc(fp_p_limb(0), a.l[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) t[i] = subc_cc(fp_p_limb(i), a.l[i]);
    t[11] = subc(fp_p_limb(11), a.l[11]);
    bool take = flag && !z;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = take ? t[i] : a.l[i];
}
__device__ __forceinline__ void fp_neg(fp_t &r, const fp_t &a) { fp_cneg(r, a, true); }
__device__ __forceinline__ void fp_dbl(fp_t &r, const fp_t &a) { fp_add(r, a, a); }
__device__ __forceinline__ void fp_mul3(fp_t &r, const fp_t &a) { fp_t t; fp_add(t, a, a); fp_add(r, t, a); }

// ---- Montgomery multiplication, CIOS with even/odd-aligned accumulators -------------------------------------
// T = ev + (od << 32); ev[k] sits at limb position k, od[k] at k+1. One row adds a*b_i and m*p and drops the
// (zero) lowest limb; the roles of the two arrays swap every row so no register moves are needed.
// first row: ev/od are written, not accumulated.
__device__ __forceinline__ void fp_row_reduce(uint32_t *ev, uint32_t *od) {
    uint32_t m = ev[0] * FP_INV32;
    od[0] = mad_lo_cc(FP_P1, m, od[0]);
    od[1] = madc_hi_cc(FP_P1, m, od[1]);
#pragma unroll
    for (int k = 2; k < 10; k += 2) {
        od[k] = madc_lo_cc(fp_p_limb(k + 1), m, od[k]);
        od[k + 1] = madc_hi_cc(fp_p_limb(k + 1), m, od[k + 1]);
    }
    od[10] = madc_lo_cc(FP_P11, m, od[10]);
    od[11] = madc_hi(FP_P11, m, od[11]);
    ev[0] = mad_lo_cc(FP_P0, m, ev[0]);
    ev[1] = madc_hi_cc(FP_P0, m, ev[1]);
#pragma unroll
    for (int k = 2; k < 12; k += 2) {
        ev[k] = madc_lo_cc(fp_p_limb(k), m, ev[k]);
        ev[k + 1] = madc_hi_cc(fp_p_limb(k), m, ev[k + 1]);
    }
    od[11] = addc(od[11], 0);
}
// generic row i >= 1. On entry `ev` holds the previous row's odd-aligned array and `od` the previous row's
// even-aligned one (whose limb 0 is zero and limb 1 still has to be folded in).
__device__ __forceinline__ void fp_row(uint32_t *ev, uint32_t *od, const uint32_t *a, uint32_t bi) {
    ev[0] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int k = 0; k < 10; k += 2) {
        od[k] = madc_lo_cc(a[k + 1], bi, od[k + 2]);
        od[k + 1] = madc_hi_cc(a[k + 1], bi, od[k + 3]);
    }
    od[10] = madc_lo_cc(a[11], bi, 0);
    od[11] = madc_hi(a[11], bi, 0);
    ev[0] = mad_lo_cc(a[0], bi, ev[0]);
    ev[1] = madc_hi_cc(a[0], bi, ev[1]);
#pragma unroll
    for (int k = 2; k < 12; k += 2) {
        ev[k] = madc_lo_cc(a[k], bi, ev[k]);
        ev[k + 1] = madc_hi_cc(a[k], bi, ev[k + 1]);
    }
    od[11] = addc(od[11], 0);
    fp_row_reduce(ev, od);
}

__device__ __forceinline__ void fp_mul_inline(fp_t &r, const fp_t &a, const fp_t &b) {
    uint32_t A[12], B[12];
    {
        uint32_t b0 = b.l[0];
#pragma unroll
        for (int k = 0; k < 12; k += 2) {
            uint64_t e = (uint64_t)a.l[k] * b0;
            uint64_t o = (uint64_t)a.l[k + 1] * b0;
            A[k] = (uint32_t)e; A[k + 1] = (uint32_t)(e >> 32);
            B[k] = (uint32_t)o; B[k + 1] = (uint32_t)(o >> 32);
        }
        fp_row_reduce(A, B);
    }
#pragma unroll
    for (int i = 1; i < 12; i += 2) {
        fp_row(B, A, a.l, b.l[i]);
        if (i + 1 < 12) fp_row(A, B, a.l, b.l[i + 1]);
    }
    // after row 11: even-aligned = B (limb 0 is zero), odd-aligned = A
    uint32_t t[12];
    t[0] = add_cc(A[0], B[1]);
#pragma unroll
    for (int k = 1; k < 11; k++) t[k] = addc_cc(A[k], B[k + 1]);
    t[11] = addc(A[11], 0);
    fp_final_sub(r, t);
}
// ONE out-of-line copy of the multiplier per translation unit, operands and result passed BY VALUE: the CUDA ABI
// keeps 12-word structs in registers across the call (cuobjdump shows no LDL/STL), so a point addition is ~10
// CALLs into 5.6 KB of code instead of 60+ KB of straight-line code. ncu on the fully inlined version showed
// `stalled_no_instruction` as a top stall reason (instruction-cache misses: every multiplication instance was
// fetched again on every loop iteration).
static __device__ __noinline__ fp_t fp_mul_fn(fp_t a, fp_t b) {
    fp_t r;
    fp_mul_inline(r, a, b);
    return r;
}
__device__ __forceinline__ void fp_mul(fp_t &r, const fp_t &a, const fp_t &b) { r = fp_mul_fn(a, b); }

// ---- dedicated Montgomery squaring (sqr_fp -> sqrx_mont_384, reference src/fields.h:40) --------------------------
// 66 off-diagonal products a_i a_j (i < j) + 12 squares instead of 144 products, then the same 12 reduction rounds:
// 234 instead of 300 wide multiply-accumulates on the heavy pipe. Products with even / odd (i + j) go to two
// accumulator arrays at absolute word positions (i+j, i+j+1), so every row is two carry chains: the one ending at
// j = 11 finishes in a fresh word, the one ending at j = 10 spills its carry into the next (fresh) word. The sum is
// doubled, the squares are added in one chain, and the 24-word value is reduced with fp_row_reduce on a 12-word
// window that slides down one limb per round (the next high word is injected at the top). The carry logic is the one
// of the limb-level model in tests/fp_sqr_model.py (3000 random + edge values against x^2 R^-1 mod p).
__device__ __forceinline__ void fp_shift_inject(uint32_t *ev, uint32_t *od, uint32_t w) {
    // window >> 32: ev = previous odd-aligned array (now even-aligned), od = previous even-aligned one (limb 0 is zero)
    ev[0] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int k = 0; k < 10; k++) od[k] = addc_cc(od[k + 2], 0);
    od[10] = addc_cc(w, 0);
    od[11] = addc(0, 0);
}
__device__ __forceinline__ void fp_sqr_inline(fp_t &r, const fp_t &a_) {
    uint32_t a[12], X[2][24];
#pragma unroll
    for (int k = 0; k < 12; k++) a[k] = a_.l[k];
#pragma unroll
    for (int k = 0; k < 24; k++) { X[0][k] = 0; X[1][k] = 0; }
#pragma unroll
    for (int i = 0; i < 11; i++) {
        {   // chain over j = 11, 9, ... (ascending), last word fresh
            constexpr int JEND = 11;
            const int par = (i + JEND) & 1;
            bool first = true;
#pragma unroll
            for (int j = 1; j <= JEND; j++) {
                if (j <= i || ((JEND - j) & 1)) continue;
                const int p = i + j;
                X[par][p] = first ? mad_lo_cc(a[i], a[j], X[par][p]) : madc_lo_cc(a[i], a[j], X[par][p]);
                first = false;
                if (j == JEND) X[par][p + 1] = madc_hi(a[i], a[j], 0);
                else X[par][p + 1] = madc_hi_cc(a[i], a[j], X[par][p + 1]);
            }
        }
        if (i < 10) {  // chain over j = 10, 8, ... (ascending), carry into the next (fresh) word
            constexpr int JEND = 10;
            const int par = (i + JEND) & 1;
            bool first = true;
#pragma unroll
            for (int j = 1; j <= JEND; j++) {
                if (j <= i || ((JEND - j) & 1)) continue;
                const int p = i + j;
                X[par][p] = first ? mad_lo_cc(a[i], a[j], X[par][p]) : madc_lo_cc(a[i], a[j], X[par][p]);
                first = false;
                X[par][p + 1] = madc_hi_cc(a[i], a[j], X[par][p + 1]);
            }
            X[par][i + 12] = addc(0, 0);
        }
    }
    // t = X[0] + X[1], doubled, plus the squares
    uint32_t t[24], d[24];
    t[0] = add_cc(X[0][0], X[1][0]);
#pragma unroll
    for (int k = 1; k < 23; k++) t[k] = addc_cc(X[0][k], X[1][k]);
    t[23] = addc(X[0][23], X[1][23]);
    d[0] = t[0] << 1;
#pragma unroll
    for (int k = 1; k < 24; k++) d[k] = __funnelshift_l(t[k - 1], t[k], 1);
    d[0] = mad_lo_cc(a[0], a[0], d[0]);
    d[1] = madc_hi_cc(a[0], a[0], d[1]);
#pragma unroll
    for (int i = 1; i < 11; i++) {
        d[2 * i] = madc_lo_cc(a[i], a[i], d[2 * i]);
        d[2 * i + 1] = madc_hi_cc(a[i], a[i], d[2 * i + 1]);
    }
    d[22] = madc_lo_cc(a[11], a[11], d[22]);
    d[23] = madc_hi(a[11], a[11], d[23]);
    // Montgomery reduction on a sliding window
    uint32_t A[12], B[12];
#pragma unroll
    for (int k = 0; k < 12; k++) { A[k] = d[k]; B[k] = 0; }
    fp_row_reduce(A, B);
#pragma unroll
    for (int s = 1; s < 12; s += 2) {
        fp_shift_inject(B, A, d[11 + s]);
        fp_row_reduce(B, A);
        if (s + 1 < 12) {
            fp_shift_inject(A, B, d[12 + s]);
            fp_row_reduce(A, B);
        }
    }
    // after round 11: even-aligned = B (limb 0 is zero), odd-aligned = A; last shift injects d[23]
    uint32_t u[12];
    u[0] = add_cc(A[0], B[1]);
#pragma unroll
    for (int k = 1; k < 11; k++) u[k] = addc_cc(A[k], B[k + 1]);
    u[11] = addc(A[11], d[23]);
    fp_final_sub(r, u);
}
#ifndef MSMB200_NO_FAST_SQR  // default: dedicated squaring (measured: accumulate 8.92 -> 8.65 ms at G1 n=2^21)
static __device__ __noinline__ fp_t fp_sqr_fn(fp_t a) {
    fp_t r;
    fp_sqr_inline(r, a);
    return r;
}
__device__ __forceinline__ void fp_sqr(fp_t &r, const fp_t &a) { r = fp_sqr_fn(a); }
#else
__device__ __forceinline__ void fp_sqr(fp_t &r, const fp_t &a) { r = fp_mul_fn(a, a); }
#endif

// from Montgomery form: a * 1 * R^-1
__device__ __forceinline__ void fp_from_mont(fp_t &r, const fp_t &a) {
    fp_t one;
    fp_set_zero(one);
    one.l[0] = 1;
    fp_mul(r, a, one);
}

// RR = 2^768 mod p (reference src/consts.c:22-26), 32-bit limbs
__device__ __forceinline__ uint32_t fp_rr_limb(int i) {
    switch (i) {
    case 0: return 0x1c341746u; case 1: return 0xf4df1f34u; case 2: return 0x09d104f1u; case 3: return 0x0a76e6a6u;
    case 4: return 0x4c95b6d5u; case 5: return 0x8de5476cu; case 6: return 0x939d83c0u; case 7: return 0x67eb88a9u;
    case 8: return 0xb519952du; case 9: return 0x9a793e85u; case 10: return 0x92cae3aau; default: return 0x11988fe5u;
    }
}
// 1/a in Montgomery form, 0 -> 0 like the reference's reciprocal_fp (src/recip.c:58-92): the branch-free safegcd of
// inv.cuh (about 25 K instructions, no data-dependent inner loops). Each thread stops when ITS g reaches zero; the
// batch-affine kernels use the warp-voting variant (batch_affine.cuh). The earlier shift / subtract binary GCD took about
// 45 K instructions with divergent inner loops and made the single-thread finalize 0.06 ms slower.
struct s30_thread_vote { MSMB200_HD bool operator()(bool done) const { return done; } };
static __device__ __noinline__ void fp_inv(fp_t &r, const fp_t &a) { s30_inverse_words(r.l, a.l, s30_thread_vote()); }

}  // namespace msmb200
