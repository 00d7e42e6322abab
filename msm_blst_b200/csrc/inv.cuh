// Branch-free modular inversion in Fp for the batch-affine kernels (sm_100a): Bernstein-Yang "safegcd" division steps
// in the half-delta variant, processed 30 steps at a time on the low words of (f, g) with the resulting 2x2 transition
// matrix applied to the full-width values held as 13 signed 30-bit limbs.
//
// Replaces reciprocal_fp (reference src/recip.c:58-92: constant-time binary GCD in assembly, ct_inverse_mod_383, with a
// Fermat fallback). Same contract: input and output in Montgomery form, 0 -> 0, result fully reduced.
//
// Why this form: every lane of a warp inverts its OWN value (the running product of its batch of slope denominators)
// and all lanes execute the identical instruction stream — no data-dependent branch, unlike the shift/subtract GCD of
// fp_inv (fp.cuh) whose inner while-loops diverge. ~27 outer iterations x (30 x ~20 ALU instructions on 32-bit words +
// ~130 IMAD.WIDE + ~160 ALU for the matrix application): about 12 Montgomery-multiplication-equivalents on the heavy
// FMA pipe and ~25 K instructions on the otherwise idle ALU pipe per inversion, versus ~460 multiplications for
// Fermat's a^(p-2).
//
// Written as plain C++ (no PTX) so that the SAME source compiles for the host: tests/test_inv_model.py builds it with
// g++ and checks it against Python's pow(x, -1, p) (edge values included) without a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MSMB200_HD __host__ __device__ __forceinline__
#else
#define MSMB200_HD inline
#endif

namespace msmb200 {

struct s30_t { int32_t v[13]; };                 // value = sum v[i] 2^(30 i); limbs 0..11 in [0, 2^30) once normalised, v[12] signed
struct trans30_t { int32_t u, v, q, r; };        // 2^30 [f', g'] = [[u, v], [q, r]] [f, g]

constexpr int32_t S30_M = 0x3fffffff;
// p, R^2 mod p (R = 2^384) as 30-bit limbs; p^-1 mod 2^30
#define MSMB200_P30 {0x3fffaaab, 0x27fbffff, 0x153ffffb, 0x2affffac, 0x30f6241e, 0x034a83da, 0x112bf673, 0x12e13ce1, 0x2cd76477, 0x1ed90d2e, 0x29a4b1ba, 0x3a8e5ff9, 0x001a0111}
#define MSMB200_RR30 {0x1c341746, 0x137c7cd0, 0x1d104f1f, 0x1db9a982, 0x15b6d50a, 0x151db132, 0x183c08de, 0x222a64e7, 0x152d67eb, 0x3a16d466, 0x3aa9a793, 0x3964b2b8, 0x0011988f}
constexpr uint32_t S30_PINV = 0x30003u;

MSMB200_HD int32_t s30_p(int i) {
    constexpr int32_t P[13] = MSMB200_P30;
    return P[i];
}
MSMB200_HD int32_t s30_rr(int i) {
    constexpr int32_t RR[13] = MSMB200_RR30;
    return RR[i];
}

// 12 x u32 (little-endian words of a 384-bit value) <-> 13 x 30-bit limbs
MSMB200_HD void s30_from_words(s30_t &r, const uint32_t *x) {
#pragma unroll
    for (int i = 0; i < 13; i++) {
        const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
        uint32_t lo = x[w] >> sh;
        if (sh > 2 && w + 1 < 12) lo |= x[w + 1] << (32 - sh);
        r.v[i] = (int32_t)(lo & (uint32_t)S30_M);
    }
}
MSMB200_HD void s30_to_words(uint32_t *x, const s30_t &a) {  // a normalised: all limbs in [0, 2^30), value < 2^384
#pragma unroll
    for (int w = 0; w < 12; w++) {
        const int bit = 32 * w, i = bit / 30, sh = bit % 30;  // word w starts inside limb i at bit sh
        uint32_t v = (uint32_t)a.v[i] >> sh;
        v |= (uint32_t)a.v[i + 1] << (30 - sh);
        if (sh > 28 && i + 2 < 13) v |= (uint32_t)a.v[i + 2] << (60 - sh);  // only 30 - sh < 2 bits came from limb i
        x[w] = v;
    }
}

// 30 division steps on the low words; zeta = -(delta + 1/2)
MSMB200_HD int32_t s30_divsteps(int32_t zeta, uint32_t f0, uint32_t g0, trans30_t &t) {
    uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll 2
    for (int i = 0; i < 30; i++) {
        uint32_t c1 = (uint32_t)(zeta >> 31);        // zeta < 0
        const uint32_t c2 = 0u - (g & 1u);           // g odd
        const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;  // conditionally negated f, u, v
        g += x & c2; q += y & c2; r += z & c2;
        c1 &= c2;
        zeta = (int32_t)((uint32_t)zeta ^ c1) - 1;   // -zeta - 2 or zeta - 1
        f += g & c1; u += q & c1; v += r & c1;
        g >>= 1; u <<= 1; v <<= 1;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return zeta;
}
// [f, g] <- t [f, g] / 2^30 (exact)
MSMB200_HD void s30_update_fg(s30_t &f, s30_t &g, const trans30_t &t) {
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    int64_t cf = u * f.v[0] + v * g.v[0], cg = q * f.v[0] + r * g.v[0];
    cf >>= 30; cg >>= 30;
#pragma unroll
    for (int i = 1; i < 13; i++) {
        const int64_t fi = f.v[i], gi = g.v[i];
        cf += u * fi + v * gi;
        cg += q * fi + r * gi;
        f.v[i - 1] = (int32_t)cf & S30_M; cf >>= 30;
        g.v[i - 1] = (int32_t)cg & S30_M; cg >>= 30;
    }
    f.v[12] = (int32_t)cf;
    g.v[12] = (int32_t)cg;
}
// [d, e] <- t [d, e] / 2^30 mod p, inputs and outputs in (-2p, p)
MSMB200_HD void s30_update_de(s30_t &d, s30_t &e, const trans30_t &t) {
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    const int32_t sd = d.v[12] >> 31, se = e.v[12] >> 31;
    int32_t md = (t.u & sd) + (t.v & se), me = (t.q & sd) + (t.r & se);
    int64_t cd = u * d.v[0] + v * e.v[0], ce = q * d.v[0] + r * e.v[0];
    // make the low 30 bits of t [d, e] + p [md, me] vanish
    md -= (int32_t)((S30_PINV * (uint32_t)cd + (uint32_t)md) & (uint32_t)S30_M);
    me -= (int32_t)((S30_PINV * (uint32_t)ce + (uint32_t)me) & (uint32_t)S30_M);
    cd += (int64_t)s30_p(0) * md;
    ce += (int64_t)s30_p(0) * me;
    cd >>= 30; ce >>= 30;
#pragma unroll
    for (int i = 1; i < 13; i++) {
        const int64_t di = d.v[i], ei = e.v[i];
        cd += u * di + v * ei;
        ce += q * di + r * ei;
        cd += (int64_t)s30_p(i) * md;
        ce += (int64_t)s30_p(i) * me;
        d.v[i - 1] = (int32_t)cd & S30_M; cd >>= 30;
        e.v[i - 1] = (int32_t)ce & S30_M; ce >>= 30;
    }
    d.v[12] = (int32_t)cd;
    e.v[12] = (int32_t)ce;
}
// r in (-2p, p) -> [0, p), negated first when sign < 0
MSMB200_HD void s30_normalize(s30_t &r, int32_t sign) {
    int32_t cond_add = r.v[12] >> 31;
    const int32_t cond_negate = sign >> 31;
#pragma unroll
    for (int i = 0; i < 13; i++) {
        r.v[i] += s30_p(i) & cond_add;
        r.v[i] = (r.v[i] ^ cond_negate) - cond_negate;
    }
#pragma unroll
    for (int i = 0; i < 12; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= S30_M; }
    cond_add = r.v[12] >> 31;
#pragma unroll
    for (int i = 0; i < 13; i++) r.v[i] += s30_p(i) & cond_add;
#pragma unroll
    for (int i = 0; i < 12; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= S30_M; }
}
MSMB200_HD bool s30_is_zero(const s30_t &a) {
    int32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 13; i++) acc |= a.v[i];
    return acc == 0;
}

// Number of 30-step rounds that certainly suffice for a 381-bit modulus in the half-delta variant:
// floor((45907 * 381 + 26313) / 19929) = 878 division steps <= 30 * 30. The device loop stops as soon as g == 0 in every
// lane of the warp (further steps are the identity on f and d), so the cap is a guard, not the usual trip count.
constexpr int S30_MAX_ROUNDS = 31;

// out = x^-1 * 2^768 mod p as 12 words (x, out: plain 384-bit little-endian words; for x = a R this is a^-1 R, i.e. the
// Montgomery-form inverse of a Montgomery-form input). x = 0 -> 0. `all_done(done)` must return true when every
// cooperating caller is done (warp vote on the device, identity on the host); returns the number of rounds executed.
template <class Vote>
MSMB200_HD int s30_inverse_words(uint32_t *out, const uint32_t *x, Vote all_done) {
    s30_t f, g, d, e;
#pragma unroll
    for (int i = 0; i < 13; i++) { f.v[i] = s30_p(i); d.v[i] = 0; e.v[i] = s30_rr(i); }
    s30_from_words(g, x);
    int32_t zeta = -1;
    int rounds = 0;
#pragma unroll 1
    while (rounds < S30_MAX_ROUNDS) {
        if (all_done(s30_is_zero(g))) break;
        trans30_t t;
        zeta = s30_divsteps(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        s30_update_de(d, e, t);
        s30_update_fg(f, g, t);
        rounds++;
    }
    s30_normalize(d, f.v[12]);
    s30_to_words(out, d);
    return rounds;
}

}  // namespace msmb200
