// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the fixed-base MSM path of LuoGuiwen/MSM_blst (a blst 0.3.10 fork).
// Nothing in the product (msm_blst_b200/, include/) may link, import or call this file;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// Parity pin: this restatement is checked (tests/test_oracle_*.py) against
//   (1) the reference itself, compiled in place from /root/reference into oracle/_ref/
//       (libblst_ref.so = src/server.c + build/assembly.S; refdrv_p{1,2}.so = main_p{1,2}.cpp),
//   (2) the known-answer vectors of SURVEY.md App. C (tests/golden/kat_appc.json),
//   (3) golden vectors generated from (1) by tests/golden/make_golden.py.
//
// Every function cites the reference file:line it follows. 64-bit limbs + unsigned __int128,
// i.e. the arithmetic of src/no_asm.h with LIMB_T_BITS=64.
//
// Build: make -C oracle   (g++ -O2 -shared -fPIC)

#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <set>
#include <string>
#include <thread>
#include <chrono>
#include <algorithm>

typedef unsigned __int128 u128;
typedef uint64_t limb_t;

// ---------------------------------------------------------------------------------------------
// Fp — src/consts.c:10-26, src/consts.h:12-22
// ---------------------------------------------------------------------------------------------
struct Fp { limb_t l[6]; };

static const Fp FP_P = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                         0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
static const limb_t FP_P0 = 0x89f3fffcfffcfffdULL;  // -1/P mod 2^64, src/consts.h:12
static const Fp FP_ONE = {{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                           0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}};
static const Fp FP_RR = {{0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                          0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL}};
static const Fp FP_ZERO = {{0, 0, 0, 0, 0, 0}};

static inline bool is_zero(const Fp &a) {
    limb_t acc = 0;
    for (int i = 0; i < 6; i++) acc |= a.l[i];
    return acc == 0;
}
static inline bool is_equal(const Fp &a, const Fp &b) { return memcmp(&a, &b, sizeof(Fp)) == 0; }

// src/no_asm.h:29-82 mul_mont_n (CIOS, final conditional subtraction; result < p)
static Fp mul(const Fp &a, const Fp &b) {
    limb_t tmp[8];
    u128 limbx;
    limb_t mask, borrow, mx, hi, carry;
    const int n = 6;
    size_t i, j;

    mx = b.l[0];
    hi = 0;
    for (i = 0; i < (size_t)n; i++) {
        limbx = (u128)mx * a.l[i] + hi;
        tmp[i] = (limb_t)limbx;
        hi = (limb_t)(limbx >> 64);
    }
    mx = FP_P0 * tmp[0];
    tmp[i] = hi;
    for (carry = 0, j = 0;;) {
        limbx = (u128)mx * FP_P.l[0] + tmp[0];
        hi = (limb_t)(limbx >> 64);
        for (i = 1; i < (size_t)n; i++) {
            limbx = (u128)mx * FP_P.l[i] + hi + tmp[i];
            tmp[i - 1] = (limb_t)limbx;
            hi = (limb_t)(limbx >> 64);
        }
        limbx = (u128)tmp[i] + hi + carry;
        tmp[i - 1] = (limb_t)limbx;
        carry = (limb_t)(limbx >> 64);
        if (++j == (size_t)n) break;
        for (mx = b.l[j], hi = 0, i = 0; i < (size_t)n; i++) {
            limbx = (u128)mx * a.l[i] + hi + tmp[i];
            tmp[i] = (limb_t)limbx;
            hi = (limb_t)(limbx >> 64);
        }
        mx = FP_P0 * tmp[0];
        limbx = (u128)hi + carry;
        tmp[i] = (limb_t)limbx;
        carry = (limb_t)(limbx >> 64);
    }
    Fp ret;
    for (borrow = 0, i = 0; i < (size_t)n; i++) {
        limbx = (u128)tmp[i] - FP_P.l[i] - borrow;
        ret.l[i] = (limb_t)limbx;
        borrow = (limb_t)(limbx >> 64) & 1;
    }
    mask = carry - borrow;  // all-ones when tmp < p
    for (i = 0; i < (size_t)n; i++) ret.l[i] = (ret.l[i] & ~mask) | (tmp[i] & mask);
    return ret;
}
static inline Fp sqr(const Fp &a) { return mul(a, a); }  // src/fields.h:39

// src/no_asm.h:104-137 add_mod_n
static Fp add(const Fp &a, const Fp &b) {
    limb_t tmp[6], carry = 0, borrow = 0, mask;
    u128 limbx;
    for (int i = 0; i < 6; i++) {
        limbx = (u128)a.l[i] + b.l[i] + carry;
        tmp[i] = (limb_t)limbx;
        carry = (limb_t)(limbx >> 64);
    }
    Fp ret;
    for (int i = 0; i < 6; i++) {
        limbx = (u128)tmp[i] - FP_P.l[i] - borrow;
        ret.l[i] = (limb_t)limbx;
        borrow = (limb_t)(limbx >> 64) & 1;
    }
    mask = carry - borrow;
    for (int i = 0; i < 6; i++) ret.l[i] = (ret.l[i] & ~mask) | (tmp[i] & mask);
    return ret;
}
// src/no_asm.h:139-161 sub_mod_n
static Fp sub(const Fp &a, const Fp &b) {
    limb_t borrow = 0, carry = 0, mask;
    u128 limbx;
    Fp ret;
    for (int i = 0; i < 6; i++) {
        limbx = (u128)a.l[i] - b.l[i] - borrow;
        ret.l[i] = (limb_t)limbx;
        borrow = (limb_t)(limbx >> 64) & 1;
    }
    mask = 0 - borrow;
    for (int i = 0; i < 6; i++) {
        limbx = (u128)ret.l[i] + (FP_P.l[i] & mask) + carry;
        ret.l[i] = (limb_t)limbx;
        carry = (limb_t)(limbx >> 64);
    }
    return ret;
}
// src/fields.h:42 cneg_fp -> cneg_mod_n (src/no_asm.h:231-253): 0 stays 0
static Fp cneg(const Fp &a, bool flag) {
    if (!flag || is_zero(a)) return a;
    return sub(FP_ZERO, a);  // 0 - a mod p
}
static inline Fp neg(const Fp &a) { return cneg(a, true); }
static inline Fp mul3(const Fp &a) { return add(add(a, a), a); }        // src/fields.h:21
static inline Fp mul8(const Fp &a) { Fp t = add(a, a); t = add(t, t); return add(t, t); }
static inline Fp from_mont(const Fp &a) { Fp one = {{1, 0, 0, 0, 0, 0}}; return mul(a, one); }
static inline Fp to_mont(const Fp &a) { return mul(a, FP_RR); }

// src/recip.c:58-92 reciprocal_fp. The reference uses a constant-time binary GCD with a Fermat
// fallback (a^(p-2), recip.c:29-56); both give the unique inverse, 0 -> 0. Restated as Fermat.
static Fp inv(const Fp &a) {
    // exponent p-2
    limb_t e[6];
    memcpy(e, FP_P.l, sizeof(e));
    e[0] -= 2;
    Fp acc = FP_ONE;
    for (int i = 383; i >= 0; i--) {
        acc = sqr(acc);
        if ((e[i / 64] >> (i % 64)) & 1) acc = mul(acc, a);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Fp2 = Fp[i]/(i^2+1) — src/fields.h:54-82, src/no_asm.h:566-579 (mul), :638-688 (sqr)
// ---------------------------------------------------------------------------------------------
struct Fp2 { Fp c[2]; };
static inline bool is_zero(const Fp2 &a) { return is_zero(a.c[0]) && is_zero(a.c[1]); }
static inline bool is_equal(const Fp2 &a, const Fp2 &b) { return memcmp(&a, &b, sizeof(Fp2)) == 0; }
static Fp2 mul(const Fp2 &a, const Fp2 &b) {  // src/no_asm.h:566-579
    Fp aa = add(a.c[0], a.c[1]);
    Fp bb = add(b.c[0], b.c[1]);
    bb = mul(bb, aa);
    aa = mul(a.c[0], b.c[0]);
    Fp cc = mul(a.c[1], b.c[1]);
    Fp2 r;
    r.c[0] = sub(aa, cc);
    r.c[1] = sub(bb, aa);
    r.c[1] = sub(r.c[1], cc);
    return r;
}
static Fp2 sqr(const Fp2 &a) {  // (a0+a1)(a0-a1) + 2 a0 a1 i, src/no_asm.h:581-597
    Fp t0 = add(a.c[0], a.c[1]);
    Fp t1 = sub(a.c[0], a.c[1]);
    Fp2 r;
    r.c[1] = mul(a.c[0], a.c[1]);
    r.c[1] = add(r.c[1], r.c[1]);
    r.c[0] = mul(t0, t1);
    return r;
}
static inline Fp2 add(const Fp2 &a, const Fp2 &b) { Fp2 r; r.c[0] = add(a.c[0], b.c[0]); r.c[1] = add(a.c[1], b.c[1]); return r; }
static inline Fp2 sub(const Fp2 &a, const Fp2 &b) { Fp2 r; r.c[0] = sub(a.c[0], b.c[0]); r.c[1] = sub(a.c[1], b.c[1]); return r; }
static inline Fp2 cneg(const Fp2 &a, bool f) { Fp2 r; r.c[0] = cneg(a.c[0], f); r.c[1] = cneg(a.c[1], f); return r; }
static inline Fp2 neg(const Fp2 &a) { return cneg(a, true); }
static inline Fp2 mul3(const Fp2 &a) { Fp2 r; r.c[0] = mul3(a.c[0]); r.c[1] = mul3(a.c[1]); return r; }
static inline Fp2 mul8(const Fp2 &a) { Fp2 r; r.c[0] = mul8(a.c[0]); r.c[1] = mul8(a.c[1]); return r; }
static Fp2 inv(const Fp2 &a) {  // src/recip.c:100-114
    Fp t0 = sqr(a.c[0]);
    Fp t1 = sqr(a.c[1]);
    t0 = add(t0, t1);
    t1 = inv(t0);
    Fp2 r;
    r.c[0] = mul(a.c[0], t1);
    r.c[1] = neg(mul(a.c[1], t1));
    return r;
}

template <class F> static F field_one();
template <> Fp field_one<Fp>() { return FP_ONE; }
template <> Fp2 field_one<Fp2>() { Fp2 r; r.c[0] = FP_ONE; r.c[1] = FP_ZERO; return r; }
template <class F> static F field_zero() { F z; memset(&z, 0, sizeof(F)); return z; }

// ---------------------------------------------------------------------------------------------
// Points — bindings/blst.h:164-165,:191-192,:251-252 (struct layouts), src/ec_ops.h
// ---------------------------------------------------------------------------------------------
template <class F> struct Aff { F x, y; };
template <class F> struct Jac { F x, y, z; };
template <class F> struct Xyzz { F x, y, zzz, zz; };

template <class F> static bool aff_is_inf(const Aff<F> &p) { return is_zero(p.x) && is_zero(p.y); }
template <class F> static bool xyzz_is_inf(const Xyzz<F> &p) { return is_zero(p.zzz) && is_zero(p.zz); }

// src/ec_ops.h:299-327 POINT_DOUBLE_IMPL_A0 (dbl-2009-l)
template <class F> static Jac<F> jac_double(const Jac<F> &p1) {
    Jac<F> p3;
    F A = sqr(p1.x), B = sqr(p1.y), C = sqr(B);
    B = add(B, p1.x);
    B = sqr(B);
    B = sub(B, A);
    B = sub(B, C);
    B = add(B, B);
    A = mul3(A);
    p3.x = sqr(A);
    p3.x = sub(p3.x, B);
    p3.x = sub(p3.x, B);
    p3.z = add(p1.z, p1.z);
    p3.z = mul(p3.z, p1.y);
    C = mul8(C);
    p3.y = sub(B, p3.x);
    p3.y = mul(p3.y, A);
    p3.y = sub(p3.y, C);
    return p3;
}

// src/ec_ops.h:40-100 POINT_DADD_IMPL with a4 == NULL (add-or-double, handles infinities)
template <class F> static Jac<F> jac_dadd(const Jac<F> &p1, const Jac<F> &p2) {
    Jac<F> p3;
    F dsx = add(p1.x, p1.x);
    F dR = mul3(sqr(p1.x));
    F dH = add(p1.y, p1.y);
    bool p2inf = is_zero(p2.z);
    p3.x = sqr(p2.z);
    p3.z = mul(p1.z, p2.z);
    bool p1inf = is_zero(p1.z);
    F aH = sqr(p1.z);
    p3.y = mul(p1.y, p2.z);
    p3.y = mul(p3.y, p3.x);   // S1
    F aR = mul(p2.y, p1.z);
    aR = mul(aR, aH);         // S2
    aR = sub(aR, p3.y);       // R = S2-S1
    p3.x = mul(p3.x, p1.x);   // U1
    aH = mul(aH, p2.x);       // U2
    F asx = add(aH, p3.x);
    aH = sub(aH, p3.x);       // H
    bool is_dbl = is_zero(aH) && is_zero(aR);
    if (is_dbl) { p3 = p1; aH = dH; aR = dR; asx = dsx; }
    p3.z = mul(p3.z, aH);
    F HH = sqr(aH);
    F HHH = mul(HH, aH);
    HHH = mul(HHH, p3.y);
    p3.y = mul(HH, p3.x);
    HH = mul(HH, asx);
    p3.x = sqr(aR);
    p3.x = sub(p3.x, HH);
    p3.y = sub(p3.y, p3.x);
    p3.y = mul(p3.y, aR);
    p3.y = sub(p3.y, HHH);
    if (p2inf) p3 = p1;
    if (p1inf) p3 = p2;
    return p3;
}

// src/ec_ops.h:710-769 POINTXYZZ_DADD_AFFINE_IMPL
template <class F> static void xyzz_dadd_affine(Xyzz<F> &p3, const Xyzz<F> &p1, const Aff<F> &p2, bool subtract) {
    if (aff_is_inf(p2)) { p3 = p1; return; }
    if (xyzz_is_inf(p1)) {
        Xyzz<F> r;
        r.x = p2.x; r.y = p2.y;
        r.zzz = cneg(field_one<F>(), subtract);
        r.zz = field_one<F>();
        p3 = r;
        return;
    }
    F P = mul(p2.x, p1.zz);
    F R = mul(p2.y, p1.zzz);
    R = cneg(R, subtract);
    P = sub(P, p1.x);
    R = sub(R, p1.y);
    Xyzz<F> r;
    if (!is_zero(P)) {
        F PP = sqr(P), PPP = mul(PP, P), Q = mul(p1.x, PP);
        r.x = sqr(R);
        P = add(Q, Q);
        r.x = sub(r.x, PPP);
        r.x = sub(r.x, P);
        Q = sub(Q, r.x);
        Q = mul(Q, R);
        r.y = mul(p1.y, PPP);
        r.y = sub(Q, r.y);
        r.zz = mul(p1.zz, PP);
        r.zzz = mul(p1.zzz, PPP);
    } else if (is_zero(R)) {
        F U = add(p2.y, p2.y);
        r.zz = sqr(U);
        r.zzz = mul(r.zz, U);
        F S = mul(p2.x, r.zz);
        F M = mul3(sqr(p2.x));
        r.x = sqr(M);
        U = add(S, S);
        r.x = sub(r.x, U);
        r.y = mul(r.zzz, p2.y);
        S = sub(S, r.x);
        S = mul(S, M);
        r.y = sub(S, r.y);
        r.zzz = cneg(r.zzz, subtract);
    } else {
        r = p1;  // X,Y keep p1's values when p3 aliases p1 (all reference call sites)
        r.zzz = field_zero<F>();
        r.zz = field_zero<F>();
    }
    p3 = r;
}

// src/ec_ops.h:642-702 POINTXYZZ_DADD_IMPL
template <class F> static void xyzz_dadd(Xyzz<F> &p3, const Xyzz<F> &p1, const Xyzz<F> &p2) {
    if (xyzz_is_inf(p2)) { p3 = p1; return; }
    if (xyzz_is_inf(p1)) { p3 = p2; return; }
    F U = mul(p1.x, p2.zz);
    F S = mul(p1.y, p2.zzz);
    F P = mul(p2.x, p1.zz);
    F R = mul(p2.y, p1.zzz);
    P = sub(P, U);
    R = sub(R, S);
    Xyzz<F> r;
    if (!is_zero(P)) {
        F PP = sqr(P), PPP = mul(PP, P), Q = mul(U, PP);
        r.x = sqr(R);
        P = add(Q, Q);
        r.x = sub(r.x, PPP);
        r.x = sub(r.x, P);
        Q = sub(Q, r.x);
        Q = mul(Q, R);
        r.y = mul(S, PPP);
        r.y = sub(Q, r.y);
        r.zz = mul(p1.zz, p2.zz);
        r.zzz = mul(p1.zzz, p2.zzz);
        r.zz = mul(r.zz, PP);
        r.zzz = mul(r.zzz, PPP);
    } else if (is_zero(R)) {
        U = add(p1.y, p1.y);
        F V = sqr(U), W = mul(V, U);
        S = mul(p1.x, V);
        F M = mul3(sqr(p1.x));
        r.x = sqr(M);
        U = add(S, S);
        r.x = sub(r.x, U);
        r.y = mul(W, p1.y);
        S = sub(S, r.x);
        S = mul(S, M);
        r.y = sub(S, r.y);
        r.zz = mul(p1.zz, V);
        r.zzz = mul(p1.zzz, W);
    } else {
        r = p1;
        r.zzz = field_zero<F>();
        r.zz = field_zero<F>();
    }
    p3 = r;
}

// src/ec_ops.h:771-777
template <class F> static Jac<F> xyzz_to_jacobian(const Xyzz<F> &in) {
    Jac<F> o;
    o.x = mul(in.x, in.zz);
    o.y = mul(in.y, in.zzz);
    o.z = in.zz;
    return o;
}

// src/e1.c:60-92 / src/e2.c:97-128: from_Jacobian + to_affine (infinity -> all-zero affine)
template <class F> static Aff<F> jac_to_affine(const Jac<F> &in) {
    Aff<F> o;
    if (is_equal(in.z, field_one<F>())) { o.x = in.x; o.y = in.y; return o; }
    F Z = inv(in.z);
    F ZZ = sqr(Z);
    o.x = mul(in.x, ZZ);
    ZZ = mul(ZZ, Z);
    o.y = mul(in.y, ZZ);
    return o;
}
template <class F> static Jac<F> jac_from_affine(const Aff<F> &a) {  // src/e1.c:94-101
    Jac<F> o;
    o.x = a.x; o.y = a.y;
    o.z = aff_is_inf(a) ? field_zero<F>() : field_one<F>();
    return o;
}
template <class F> static Jac<F> jac_infinity() { Jac<F> o; memset(&o, 0, sizeof(o)); return o; }

// Generators — src/e1.c:20-32, src/e2.c:23-47 (Montgomery form)
template <class F> static Jac<F> generator();
template <> Jac<Fp> generator<Fp>() {
    Jac<Fp> g;
    g.x = {{0x5cb38790fd530c16ULL, 0x7817fc679976fff5ULL, 0x154f95c7143ba1c1ULL,
            0xf0ae6acdf3d0e747ULL, 0xedce6ecc21dbf440ULL, 0x120177419e0bfb75ULL}};
    g.y = {{0xbaac93d50ce72271ULL, 0x8c22631a7918fd8eULL, 0xdd595f13570725ceULL,
            0x51ac582950405194ULL, 0x0e1c8c3fad0059c0ULL, 0x0bbc3efc5008a26aULL}};
    g.z = FP_ONE;
    return g;
}
template <> Jac<Fp2> generator<Fp2>() {
    Jac<Fp2> g;
    g.x.c[0] = {{0xf5f28fa202940a10ULL, 0xb3f5fb2687b4961aULL, 0xa1a893b53e2ae580ULL,
                 0x9894999d1a3caee9ULL, 0x6f67b7631863366bULL, 0x058191924350bcd7ULL}};
    g.x.c[1] = {{0xa5a9c0759e23f606ULL, 0xaaa0c59dbccd60c3ULL, 0x3bb17e18e2867806ULL,
                 0x1b1ab6cc8541b367ULL, 0xc2b6ed0ef2158547ULL, 0x11922a097360edf3ULL}};
    g.y.c[0] = {{0x4c730af860494c4aULL, 0x597cfa1f5e369c5aULL, 0xe7e6856caa0a635aULL,
                 0xbbefb5e96e0d495fULL, 0x07d3a975f0ef25a2ULL, 0x0083fd8e7e80dae5ULL}};
    g.y.c[1] = {{0xadc0fc92df64b05dULL, 0x18aa270a2b1461dcULL, 0x86adac6a3be4eba0ULL,
                 0x79495c4ec93da33aULL, 0xe7175850a43ccaedULL, 0x0b2bc2a163de1bf2ULL}};
    g.z = field_one<Fp2>();
    return g;
}

// Serialisation — src/e1.c:139-162, src/e2.c:176-203, src/bytes.h:35-44
static void fp_be_bytes(unsigned char out[48], const Fp &mont) {
    Fp t = from_mont(mont);
    for (int i = 0; i < 48; i++) out[i] = (unsigned char)(t.l[(47 - i) / 8] >> (8 * ((47 - i) % 8)));
}
static void affine_serialize(unsigned char *out, const Aff<Fp> &a) {
    if (aff_is_inf(a)) { memset(out, 0, 96); out[0] = 0x40; return; }
    fp_be_bytes(out, a.x);
    fp_be_bytes(out + 48, a.y);
}
static void affine_serialize(unsigned char *out, const Aff<Fp2> &a) {
    if (aff_is_inf(a)) { memset(out, 0, 192); out[0] = 0x40; return; }
    fp_be_bytes(out, a.x.c[1]);
    fp_be_bytes(out + 48, a.x.c[0]);
    fp_be_bytes(out + 96, a.y.c[1]);
    fp_be_bytes(out + 144, a.y.c[0]);
}

// ---------------------------------------------------------------------------------------------
// Scalars: 4 x u64 little-endian limbs (src_from_aztec uint256_t::data[4]); group order
// auxiliaryfunc.h:5-7
// ---------------------------------------------------------------------------------------------
struct U256 { uint64_t d[4]; };
static const U256 R_ORDER = {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL}};
static inline bool u256_is_zero(const U256 &a) { return (a.d[0] | a.d[1] | a.d[2] | a.d[3]) == 0; }
static inline int u256_cmp(const U256 &a, const U256 &b) {
    for (int i = 3; i >= 0; i--) if (a.d[i] != b.d[i]) return a.d[i] < b.d[i] ? -1 : 1;
    return 0;
}
static inline U256 u256_shr(const U256 &a, unsigned s) {  // 0 < s < 64
    U256 r;
    for (int i = 0; i < 4; i++) r.d[i] = (a.d[i] >> s) | (i < 3 ? a.d[i + 1] << (64 - s) : 0);
    return r;
}
static inline U256 u256_sub(const U256 &a, const U256 &b) {
    U256 r; limb_t borrow = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)a.d[i] - b.d[i] - borrow; r.d[i] = (limb_t)t; borrow = (limb_t)(t >> 64) & 1; }
    return r;
}

// splitmix64 scalar stream of SURVEY.md App. C (harness definition; the reference's own
// generator, auxiliaryfunc.h:178-207, is unseeded OpenSSL RAND_bytes -> SHA256)
static inline uint64_t splitmix_next(uint64_t &state) {
    state += 0x9E3779B97F4A7C15ULL;
    uint64_t z = state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// ---------------------------------------------------------------------------------------------
// Configurations — ches_config_files/config_file_n_exp_*.h:5-17
// ---------------------------------------------------------------------------------------------
struct Config { const char *name; int n_exp, e, h, a, d, bsize, e_bgmw, h_bgmw; };
static const Config CONFIGS[] = {
    {"8", 8, 12, 22, 7, 6, 857, 10, 26},          {"9", 9, 13, 20, 231, 6, 1725, 11, 24},
    {"10", 10, 13, 20, 231, 6, 1725, 12, 22},     {"11", 11, 14, 19, 7, 6, 3417, 13, 20},
    {"12", 12, 14, 19, 7, 6, 3417, 13, 20},       {"13", 13, 16, 16, 29677, 6, 18343, 15, 17},
    {"14", 14, 16, 16, 29677, 6, 18343, 15, 17},  {"15", 15, 16, 16, 29677, 6, 18343, 16, 16},
    {"16", 16, 19, 14, 231, 6, 109244, 17, 15},   {"16_beta", 16, 18, 15, 7, 6, 54618, 17, 15},
    {"17", 17, 20, 13, 29677, 6, 220931, 17, 15}, {"17_beta", 17, 19, 14, 231, 6, 109244, 17, 15},
    {"18", 18, 20, 13, 29677, 6, 220931, 19, 14}, {"19", 19, 20, 13, 29677, 6, 220931, 20, 13},
    {"20", 20, 22, 12, 7419, 6, 874437, 20, 13},  {"20_beta", 20, 20, 13, 29677, 6, 220931, 20, 13},
    {"21", 21, 22, 12, 7419, 6, 874437, 22, 12},
};
static const Config *find_config(const char *name) {
    for (const Config &c : CONFIGS) if (strcmp(c.name, name) == 0) return &c;
    return nullptr;
}

// auxiliaryfunc.h:234-254
static int omega2(int n) { int e = 0; while (n % 2 == 0) { e++; n >>= 1; } return e; }
static int omega3(int n) { int e = 0; while (n % 3 == 0) { e++; n /= 3; } return e; }

// auxiliaryfunc.h:257-288 construct_bucket_set
static std::vector<int> construct_bucket_set(int q, int ah) {
    std::set<int> B = {0, 1};
    for (int i = 2; i <= q / 2; ++i) if (((omega2(i) + omega3(i)) % 2) == 0) B.insert(i);
    for (int i = q / 4; i < q / 2; ++i) if (B.count(i) && B.count(q - 2 * i)) B.erase(q - 2 * i);
    for (int i = q / 6; i < q / 4; ++i) if (B.count(i) && B.count(q - 3 * i)) B.erase(q - 3 * i);
    for (int i = 1; i <= ah + 1; ++i) if (((omega2(i) + omega3(i)) % 2) == 0) B.insert(i);
    return std::vector<int>(B.begin(), B.end());
}

struct Digit { int m, b, alpha; };  // bindings/blst.h:253 digit_decomposition

// ---------------------------------------------------------------------------------------------
// The MSM context: globals of main_p1.cpp:41-50 gathered in a struct
// ---------------------------------------------------------------------------------------------
template <class F> struct Ctx {
    Config cfg;
    size_t n;
    std::vector<Aff<F>> fix_points;         // FIX_POINTS_LIST
    std::vector<int> bucket_set;            // BUCKET_SET
    std::vector<int> value_to_index;        // BUCKET_VALUE_TO_ITS_INDEX
    std::vector<Digit> hash;                // DIGIT_CONVERSION_HASH_TABLE
    std::vector<Aff<F>> table_3nh;          // PRECOMPUTATION_POINTS_LIST_3nh
    std::vector<Aff<F>> table_bgmw;         // PRECOMPUTATION_POINTS_LIST_BGMW95
    int threads = 1;
};

// main_p1.cpp:52-66 init_fix_point_list: P_i = 2^(i+1) G, i = first .. first+n-1
template <class F> static void init_fix_point_list(Ctx<F> &c, size_t first) {
    c.fix_points.resize(c.n);
    Jac<F> t = generator<F>();
    for (size_t i = 0; i < first; i++) t = jac_double(t);
    for (size_t i = 0; i < c.n; i++) {
        t = jac_double(t);
        c.fix_points[i] = jac_to_affine(t);
    }
}

// main_p1.cpp:72-91 single_scalar_multiplication (LSB-first double-and-add, then to_affine)
template <class F> static Aff<F> single_scalar_multiplication(uint64_t scalar, const Aff<F> &Q) {
    Jac<F> ret = jac_infinity<F>();
    Jac<F> xyzQ = jac_from_affine(Q);
    while (scalar > 0) {
        if (scalar & 1) ret = jac_dadd(ret, xyzQ);
        xyzQ = jac_dadd(xyzQ, xyzQ);
        scalar >>= 1;
    }
    return jac_to_affine(ret);
}

// main_p1.cpp:128-153: bucket set, index map, digit hash (two passes, later writes win)
template <class F> static void init_ches_params(Ctx<F> &c) {
    int q = 1 << c.cfg.e;
    c.bucket_set = construct_bucket_set(q, c.cfg.a);
    c.value_to_index.assign(q / 2 + 1, 0);
    for (size_t i = 0; i < c.bucket_set.size(); i++) c.value_to_index[c.bucket_set[i]] = (int)i;
    c.hash.assign(q + 1, Digit{0, 0, 0});
    for (int m = 1; m <= 3; m++)
        for (int b : c.bucket_set) if ((long)m * b <= q) c.hash[q - m * b] = {m, b, 1};
    for (int m = 1; m <= 3; m++)
        for (int b : c.bucket_set) if ((long)m * b <= q) c.hash[m * b] = {m, b, 0};
}

// main_p1.cpp:156-172 table loop
template <class F> static void build_table_3nh(Ctx<F> &c) {
    int h = c.cfg.h;
    uint64_t q = 1ull << c.cfg.e;
    c.table_3nh.resize(3 * c.n * h);
    auto work = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            Aff<F> qjQi = c.fix_points[i];
            for (int j = 0; j < h; j++) {
                for (int m = 1; m <= 3; m++) {
                    size_t idx = 3 * (i * h + j) + m - 1;
                    c.table_3nh[idx] = (m == 1) ? qjQi : single_scalar_multiplication<F>(m, qjQi);
                }
                qjQi = single_scalar_multiplication<F>(q, qjQi);
            }
        }
    };
    int T = c.threads > 1 ? c.threads : 1;
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(work, c.n * t / T, c.n * (t + 1) / T);
    for (auto &x : th) x.join();
}
// main_p1.cpp:94-122 init_pippenger_BGMW95 (incl. the copy shortcut :99-106)
template <class F> static void build_table_bgmw(Ctx<F> &c) {
    int h = c.cfg.h_bgmw;
    uint64_t q = 1ull << c.cfg.e_bgmw;
    c.table_bgmw.resize(c.n * h);
    if (!c.table_3nh.empty() && c.cfg.e == c.cfg.e_bgmw) {
        for (size_t i = 0; i < c.n; i++)
            for (int j = 0; j < h; j++) { size_t idx = i * h + j; c.table_bgmw[idx] = c.table_3nh[3 * idx]; }
        return;
    }
    auto work = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            Aff<F> qjQi = c.fix_points[i];
            for (int j = 0; j < h; j++) {
                c.table_bgmw[i * h + j] = qjQi;
                qjQi = single_scalar_multiplication<F>(q, qjQi);
            }
        }
    };
    int T = c.threads > 1 ? c.threads : 1;
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(work, c.n * t / T, c.n * (t + 1) / T);
    for (auto &x : th) x.join();
}

// auxiliaryfunc.h:83-90 trans_uint256_t_to_standard_q_ary_expr
static void std_q_ary(int *out, const U256 &a, int e, int h) {
    U256 tmp = a;
    uint32_t mask = (1u << e) - 1;
    for (int i = 0; i < h; i++) { out[i] = (int)(tmp.d[0] & mask); tmp = u256_shr(tmp, e); }
}
// auxiliaryfunc.h:92-118 trans_uint256_t_to_MB_radixq_expr: out_m signed (+-m), out_b = b
template <class F> static void mb_radixq(const Ctx<F> &c, int *out_m, int *out_b, const U256 &a) {
    int h = c.cfg.h;
    std::vector<int> d(h + 1, 0);
    std_q_ary(d.data(), a, c.cfg.e, h);
    for (int i = 0; i < h; i++) {
        Digit t = c.hash[d[i]];
        if (t.alpha == 0) { out_m[i] = t.m; out_b[i] = t.b; }
        else { out_m[i] = -t.m; out_b[i] = t.b; d[i + 1] += 1; }
    }
}
// auxiliaryfunc.h:130-145 trans_uint256_t_to_qhalf_expr
static void qhalf_expr(int *out, const U256 &a, int e, int h) {
    std_q_ary(out, a, e, h);
    int q = 1 << e, qhalf = q >> 1;
    for (int i = 0; i < h - 1; i++) if (out[i] > qhalf) { out[i] -= q; out[i + 1] += 1; }
}

// src/multi_scalar.c:301-321 integrate_buckets_accumulation_d_CHES
template <class F> static Jac<F> integrate_buckets_d_ches(std::vector<Xyzz<F>> &buckets, const std::vector<int> &bset, int d_max) {
    Xyzz<F> tmp, tmp1;
    std::vector<Xyzz<F>> tmp_d(d_max + 1);
    memset(&tmp, 0, sizeof(tmp));
    memset(tmp_d.data(), 0, sizeof(Xyzz<F>) * (d_max + 1));
    for (size_t i = bset.size() - 1; i > 0; --i) {
        xyzz_dadd(tmp, tmp, buckets[i]);
        int differ = bset[i] - bset[i - 1];
        xyzz_dadd(tmp_d[differ], tmp_d[differ], tmp);
    }
    memset(&tmp, 0, sizeof(tmp));
    memset(&tmp1, 0, sizeof(tmp1));
    for (int i = d_max; i > 0; --i) {
        xyzz_dadd(tmp, tmp, tmp_d[i]);
        xyzz_dadd(tmp1, tmp1, tmp);
    }
    return xyzz_to_jacobian(tmp1);
}
// src/multi_scalar.c:281-297 integrate_buckets: sum (i+1)*buckets[i], i < 2^wbits
template <class F> static Jac<F> integrate_buckets(Xyzz<F> *buckets, size_t wbits) {
    Xyzz<F> ret, acc;
    size_t n = (size_t)1 << wbits;
    acc = buckets[--n];
    ret = buckets[n];
    memset(&buckets[n], 0, sizeof(buckets[n]));
    while (n--) {
        xyzz_dadd(acc, acc, buckets[n]);
        xyzz_dadd(ret, ret, acc);
        memset(&buckets[n], 0, sizeof(buckets[n]));
    }
    return xyzz_to_jacobian(ret);
}

// src/multi_scalar.c:421-463 tile_pippenger_d_CHES. `faithful_bug` reproduces the guard at :461
// (tests the previous entry's index); otherwise the mathematically intended sum.
template <class F>
static Jac<F> tile_pippenger_d_ches(const Aff<F> *const *points, size_t npoints, const int *scalars,
                                    const unsigned char *signs, const std::vector<int> &bset,
                                    const std::vector<int> &v2i, int d_max, bool faithful_bug) {
    std::vector<Xyzz<F>> buckets(bset.size());
    memset(buckets.data(), 0, sizeof(Xyzz<F>) * bset.size());
    for (size_t k = 0; k < npoints; k++) {
        int idx = v2i[scalars[k]];
        bool take = idx != 0;
        if (faithful_bug && k == npoints - 1 && npoints >= 2) {
            take = v2i[scalars[k - 1]] != 0;  // :461 `if(booth_idx)` with the stale index
        }
        if (take) xyzz_dadd_affine(buckets[idx], buckets[idx], *points[k], signs[k] != 0);
    }
    return integrate_buckets_d_ches(buckets, bset, d_max);
}

// src/multi_scalar.c:506-547 tile_pippenger_BGMW95
template <class F>
static Jac<F> tile_pippenger_bgmw95(const Aff<F> *const *points, size_t npoints, const int *scalars,
                                    const unsigned char *signs, size_t q_exponent) {
    size_t bsz = ((size_t)1 << (q_exponent - 1)) + 1;
    std::vector<Xyzz<F>> buckets(bsz);
    memset(buckets.data(), 0, sizeof(Xyzz<F>) * bsz);
    for (size_t k = 0; k < npoints; k++) {
        int idx = scalars[k];
        if (idx || k == npoints - 1) xyzz_dadd_affine(buckets[idx], buckets[idx], *points[k], signs[k] != 0);
    }
    return integrate_buckets(buckets.data() + 1, q_exponent - 1);
}

// main_p1.cpp:192-246 pippenger_variant_q_over_5_CHES
template <class F> static Aff<F> method_ches(const Ctx<F> &c, const U256 *sc, bool faithful_bug) {
    int h = c.cfg.h;
    size_t npoints = c.n * h;
    std::vector<int> scalars(npoints + 1, 0);
    std::vector<unsigned char> signs(npoints);
    std::vector<const Aff<F> *> ptr(npoints);
    std::vector<int> em(h), eb(h);
    for (size_t i = 0; i < c.n; i++) {
        mb_radixq(c, em.data(), eb.data(), sc[i]);
        for (int j = 0; j < h; j++) {
            size_t idx = i * h + j;
            int m = em[j];
            scalars[idx] = eb[j];
            if (m > 0) { ptr[idx] = &c.table_3nh[3 * idx + m - 1]; signs[idx] = 0; }
            else { ptr[idx] = &c.table_3nh[3 * idx - m - 1]; signs[idx] = 1; }
        }
    }
    Jac<F> ret = tile_pippenger_d_ches<F>(ptr.data(), npoints, scalars.data(), signs.data(), c.bucket_set,
                                          c.value_to_index, c.cfg.d, faithful_bug);
    return jac_to_affine(ret);
}
// main_p1.cpp:249-291 + src/multi_scalar.c:748-775 construct_nh_scalars_nh_points
template <class F> static Aff<F> method_ches_integral(const Ctx<F> &c, const U256 *sc, bool faithful_bug) {
    int h = c.cfg.h;
    size_t npoints = c.n * h;
    std::vector<int> scalars(npoints + 2, 0);
    for (size_t i = 0; i < c.n; i++) std_q_ary(&scalars[i * h], sc[i], c.cfg.e, h);
    std::vector<unsigned char> signs(npoints);
    std::vector<const Aff<F> *> ptr(npoints);
    for (size_t i = 0; i < npoints; i++) {  // the reference peels the last iteration (:769-774)
        Digit t = c.hash[scalars[i]];
        scalars[i] = t.b;
        signs[i] = (unsigned char)t.alpha;
        if (t.alpha && i + 1 < npoints) ++scalars[i + 1];
        ptr[i] = &c.table_3nh[3 * i + t.m - 1];
    }
    Jac<F> ret = tile_pippenger_d_ches<F>(ptr.data(), npoints, scalars.data(), signs.data(), c.bucket_set,
                                          c.value_to_index, c.cfg.d, faithful_bug);
    return jac_to_affine(ret);
}
// main_p1.cpp:294-398 pippenger_variant_BGMW95
template <class F> static Aff<F> method_bgmw95(const Ctx<F> &c, const U256 *sc) {
    int h = c.cfg.h_bgmw, e = c.cfg.e_bgmw;
    size_t npoints = c.n * h;
    std::vector<int> scalars(npoints);
    std::vector<unsigned char> signs(npoints);
    std::vector<const Aff<F> *> ptr(npoints);
    std::vector<int> ex(h);
    bool trick = (c.cfg.n_exp == 13 || c.cfg.n_exp == 14 || c.cfg.n_exp == 16 || c.cfg.n_exp == 17);
    for (size_t i = 0; i < c.n; i++) {
        U256 aa = sc[i];
        bool cond = trick && (aa.d[3] > (1ull << 62));
        if (cond) aa = u256_sub(R_ORDER, aa);
        qhalf_expr(ex.data(), aa, e, h);
        for (int j = 0; j < h; j++) {
            size_t idx = i * h + j;
            int v = ex[j];
            ptr[idx] = &c.table_bgmw[idx];
            if (v > 0) { scalars[idx] = v; signs[idx] = cond ? 1 : 0; }
            else { scalars[idx] = -v; signs[idx] = cond ? 0 : 1; }
        }
    }
    Jac<F> ret = tile_pippenger_bgmw95<F>(ptr.data(), npoints, scalars.data(), signs.data(), e);
    return jac_to_affine(ret);
}

// src/multi_scalar.c:268-275
static size_t pippenger_window_size(size_t npoints) {
    size_t wbits;
    for (wbits = 0; npoints >>= 1; wbits++) ;
    return wbits > 12 ? wbits - 3 : (wbits > 4 ? wbits - 2 : (wbits ? 2 : 1));
}
// src/ec_mult.h:23-40 get_wval_limb (up to 25 bits), restated without the branch-free masking
static limb_t get_wval_limb(const unsigned char *d, size_t off, size_t bits) {
    size_t top = (off + bits - 1) / 8;
    d += off / 8;
    top -= off / 8 - 1;  // number of bytes to read (1..4)
    limb_t ret = 0;
    for (size_t i = 0; i < 4 && i < top; i++) ret |= (limb_t)d[i] << (8 * i);
    return ret >> (off % 8);
}
// src/ec_mult.h:47-56 booth_encode
static limb_t booth_encode(limb_t wval, size_t sz) {
    limb_t mask = 0 - (wval >> sz);
    wval = (wval + 1) >> 1;
    wval = (wval & ~mask) | ((0 - wval) & mask);
    return wval;
}
// src/multi_scalar.c:347-356 ptype##_bucket
template <class F> static void pip_bucket(Xyzz<F> *buckets, limb_t booth_idx, size_t wbits, const Aff<F> &p) {
    bool sign = (booth_idx >> wbits) & 1;
    booth_idx &= ((limb_t)1 << wbits) - 1;
    if (booth_idx--) xyzz_dadd_affine(buckets[booth_idx], buckets[booth_idx], p, sign);
}
// src/multi_scalar.c:383-419 s_tile_pippenger (contiguous points / 32-byte scalars)
template <class F>
static Jac<F> s_tile_pippenger(const Aff<F> *points, size_t npoints, const unsigned char *scalars, size_t nbits,
                               Xyzz<F> *buckets, size_t bit0, size_t wbits, size_t cbits) {
    size_t nbytes = (nbits + 7) / 8;
    limb_t wmask = ((limb_t)1 << (wbits + 1)) - 1;
    size_t z = bit0 == 0;
    bit0 -= z ^ 1; wbits += z ^ 1;
    for (size_t i = 0; i < npoints; i++) {
        limb_t wval = (get_wval_limb(scalars + i * nbytes, bit0, wbits) << z) & wmask;
        wval = booth_encode(wval, cbits);
        pip_bucket(buckets, wval, cbits, points[i]);
    }
    return integrate_buckets(buckets, cbits - 1);
}
// src/multi_scalar.c:549-576 s_mult_pippenger
template <class F>
static Jac<F> s_mult_pippenger(const Aff<F> *points, size_t npoints, const unsigned char *scalars, size_t nbits, size_t window) {
    size_t wbits, cbits, bit0 = nbits;
    window = window ? window : pippenger_window_size(npoints);
    std::vector<Xyzz<F>> buckets((size_t)1 << (window - 1));
    memset(buckets.data(), 0, sizeof(Xyzz<F>) * buckets.size());
    Jac<F> ret = jac_infinity<F>();
    wbits = nbits % window;
    cbits = wbits + 1;
    while (bit0 -= wbits) {
        Jac<F> tile = s_tile_pippenger<F>(points, npoints, scalars, nbits, buckets.data(), bit0, wbits, cbits);
        ret = jac_dadd(ret, tile);
        for (size_t i = 0; i < window; i++) ret = jac_double(ret);
        cbits = wbits = window;
    }
    Jac<F> tile = s_tile_pippenger<F>(points, npoints, scalars, nbits, buckets.data(), 0, wbits, cbits);
    return jac_dadd(ret, tile);
}
// main_p1.cpp:400-436 pippenger_blst_built_in
template <class F> static Aff<F> method_pippenger(const Ctx<F> &c, const U256 *sc) {
    // blst_scalar_from_uint64 (src/exports.c:356-375): 4 LE u64 -> 32 LE bytes == memory image on x86
    Jac<F> ret = s_mult_pippenger<F>(c.fix_points.data(), c.n, (const unsigned char *)sc, 255, 0);
    return jac_to_affine(ret);
}

// Closed form for the synthetic points P_i = 2^(i+1) G: k = sum s_i 2^(first+i+1) mod r,
// result = k*G by MSB-first double-and-add. Independent of every bucket structure above.
static U256 addmod_r(const U256 &a, const U256 &b) {
    U256 s; limb_t carry = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)a.d[i] + b.d[i] + carry; s.d[i] = (limb_t)t; carry = (limb_t)(t >> 64); }
    if (carry || u256_cmp(s, R_ORDER) >= 0) s = u256_sub(s, R_ORDER);
    return s;
}
static U256 closed_form_scalar(const U256 *sc, size_t n, size_t first) {
    // Horner from the top: k = (((s_{n-1})*2 + s_{n-2})*2 + ... + s_0) * 2^(first+1)
    U256 k = {{0, 0, 0, 0}};
    for (size_t i = n; i-- > 0;) { k = addmod_r(k, k); k = addmod_r(k, sc[i]); }
    for (size_t i = 0; i < first + 1; i++) k = addmod_r(k, k);
    return k;
}
template <class F> static Aff<F> scalar_mul_generator(const U256 &k) {
    Jac<F> acc = jac_infinity<F>();
    Jac<F> g = generator<F>();
    for (int i = 255; i >= 0; i--) {
        acc = jac_dadd(acc, acc);
        if ((k.d[i / 64] >> (i % 64)) & 1) acc = jac_dadd(acc, g);
    }
    return jac_to_affine(acc);
}

// ---------------------------------------------------------------------------------------------
// C exports for ctypes (tests / bench cpu_baseline only)
// ---------------------------------------------------------------------------------------------
template <class F> static int field_op_impl(int op, const F *a, const F *b, F *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        switch (op) {
        case 0: out[i] = mul(a[i], b[i]); break;
        case 1: out[i] = sqr(a[i]); break;
        case 2: out[i] = add(a[i], b[i]); break;
        case 3: out[i] = sub(a[i], b[i]); break;
        case 4: out[i] = neg(a[i]); break;
        case 5: out[i] = mul3(a[i]); break;
        case 6: out[i] = inv(a[i]); break;
        case 7: out[i] = mul8(a[i]); break;
        default: return -1;
        }
    }
    return 0;
}
// point ops: 0 jac_dadd(a:Jac,b:Jac)->Jac, 1 jac_double(a)->Jac, 2 xyzz_dadd_affine(a:Xyzz,b:Aff,sub=flag)->Xyzz,
// 3 xyzz_dadd(a:Xyzz,b:Xyzz)->Xyzz, 4 xyzz_to_jacobian(a)->Jac, 5 jac_to_affine(a)->Aff
template <class F> static int point_op_impl(int op, const void *a, const void *b, const unsigned char *flags, void *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        switch (op) {
        case 0: ((Jac<F> *)out)[i] = jac_dadd(((const Jac<F> *)a)[i], ((const Jac<F> *)b)[i]); break;
        case 1: ((Jac<F> *)out)[i] = jac_double(((const Jac<F> *)a)[i]); break;
        case 2: xyzz_dadd_affine(((Xyzz<F> *)out)[i], ((const Xyzz<F> *)a)[i], ((const Aff<F> *)b)[i], flags && flags[i]); break;
        case 3: xyzz_dadd(((Xyzz<F> *)out)[i], ((const Xyzz<F> *)a)[i], ((const Xyzz<F> *)b)[i]); break;
        case 4: ((Jac<F> *)out)[i] = xyzz_to_jacobian(((const Xyzz<F> *)a)[i]); break;
        case 5: ((Aff<F> *)out)[i] = jac_to_affine(((const Jac<F> *)a)[i]); break;
        default: return -1;
        }
    }
    return 0;
}

struct OracleHandle { int group; Ctx<Fp> *g1; Ctx<Fp2> *g2; };

template <class F> static int ctx_msm(const Ctx<F> &c, int method, const U256 *sc, Aff<F> *out, int faithful_bug) {
    switch (method) {
    case 1: *out = method_ches(c, sc, faithful_bug != 0); return 0;
    case 2: *out = method_ches_integral(c, sc, faithful_bug != 0); return 0;
    case 3: *out = method_bgmw95(c, sc); return 0;
    case 4: *out = method_pippenger(c, sc); return 0;
    }
    return -1;
}

extern "C" {

int oracle_fp_op(int op, const void *a, const void *b, void *out, size_t n) { return field_op_impl<Fp>(op, (const Fp *)a, (const Fp *)b, (Fp *)out, n); }
int oracle_fp2_op(int op, const void *a, const void *b, void *out, size_t n) { return field_op_impl<Fp2>(op, (const Fp2 *)a, (const Fp2 *)b, (Fp2 *)out, n); }
int oracle_point_op(int group, int op, const void *a, const void *b, const unsigned char *flags, void *out, size_t n) {
    return group == 1 ? point_op_impl<Fp>(op, a, b, flags, out, n) : point_op_impl<Fp2>(op, a, b, flags, out, n);
}
void oracle_fp_to_mont(const void *a, void *out, size_t n) { for (size_t i = 0; i < n; i++) ((Fp *)out)[i] = to_mont(((const Fp *)a)[i]); }
void oracle_fp_from_mont(const void *a, void *out, size_t n) { for (size_t i = 0; i < n; i++) ((Fp *)out)[i] = from_mont(((const Fp *)a)[i]); }

// seeded scalars: n x 4 u64 LE limbs
void oracle_gen_scalars(uint64_t seed, size_t n, uint64_t *out) {
    uint64_t state = seed;
    for (size_t i = 0; i < n; i++) {
        U256 s;
        do {
            for (int k = 0; k < 4; k++) s.d[k] = splitmix_next(state);
            s.d[3] >>= 1;
        } while (u256_cmp(s, R_ORDER) >= 0);
        memcpy(out + 4 * i, s.d, 32);
    }
}

int oracle_config(const char *name, int *out9) {
    const Config *c = find_config(name);
    if (!c) return -1;
    int v[9] = {c->n_exp, c->e, c->h, c->a, c->d, c->bsize, c->e_bgmw, c->h_bgmw, (int)pippenger_window_size((size_t)1 << c->n_exp)};
    memcpy(out9, v, sizeof(v));
    return 0;
}
// bucket set for (e, a): returns |B|; fills out[] if cap allows
long oracle_bucket_set(int e, int a, int *out, long cap) {
    std::vector<int> B = construct_bucket_set(1 << e, a);
    if (out && (long)B.size() <= cap) memcpy(out, B.data(), B.size() * sizeof(int));
    return (long)B.size();
}
// main_bucket_set_construction.cpp:74-113 check_bucket_set_validity; :115-122 max gap. returns max gap or -1
int oracle_bucket_set_check(int e, int a) {
    int q = 1 << e;
    std::vector<int> B = construct_bucket_set(q, a);
    std::set<int> S(B.begin(), B.end());
    for (int i = 0; i <= q; i++) {
        bool ok = false;
        for (int m = 1; m <= 3 && !ok; m++) {
            if (i % m == 0 && S.count(i / m)) ok = true;
            if (!ok && (q - i) >= 0 && (q - i) % m == 0 && S.count((q - i) / m)) ok = true;
        }
        if (!ok) return -1;
    }
    for (int i = 0; i <= a + 1; i++) {
        bool ok = false;
        for (int m = 1; m <= 3 && !ok; m++) if (i % m == 0 && S.count(i / m)) ok = true;
        if (!ok) return -1;
    }
    int gap = 0;
    for (size_t i = 1; i < B.size(); i++) if (B[i] - B[i - 1] > gap) gap = B[i] - B[i - 1];
    return gap;
}

void *oracle_ctx_create(int group, const char *cfgname, size_t n) {
    const Config *c = find_config(cfgname);
    if (!c || (group != 1 && group != 2)) return nullptr;
    OracleHandle *h = new OracleHandle{group, nullptr, nullptr};
    if (group == 1) { h->g1 = new Ctx<Fp>(); h->g1->cfg = *c; h->g1->n = n; init_ches_params(*h->g1); }
    else { h->g2 = new Ctx<Fp2>(); h->g2->cfg = *c; h->g2->n = n; init_ches_params(*h->g2); }
    return h;
}
void oracle_ctx_destroy(void *hv) {
    OracleHandle *h = (OracleHandle *)hv;
    if (!h) return;
    delete h->g1; delete h->g2; delete h;
}
void oracle_ctx_set_threads(void *hv, int t) { OracleHandle *h = (OracleHandle *)hv; if (h->g1) h->g1->threads = t; else h->g2->threads = t; }
// P_i = 2^(first+i+1) G
void oracle_ctx_init_fix_points(void *hv, size_t first) {
    OracleHandle *h = (OracleHandle *)hv;
    if (h->g1) init_fix_point_list(*h->g1, first); else init_fix_point_list(*h->g2, first);
}
void oracle_ctx_set_points(void *hv, const void *pts) {
    OracleHandle *h = (OracleHandle *)hv;
    if (h->g1) { h->g1->fix_points.resize(h->g1->n); memcpy(h->g1->fix_points.data(), pts, h->g1->n * sizeof(Aff<Fp>)); }
    else { h->g2->fix_points.resize(h->g2->n); memcpy(h->g2->fix_points.data(), pts, h->g2->n * sizeof(Aff<Fp2>)); }
}
const void *oracle_ctx_points(void *hv) { OracleHandle *h = (OracleHandle *)hv; return h->g1 ? (const void *)h->g1->fix_points.data() : (const void *)h->g2->fix_points.data(); }
// which: 0 = 3nh (CHES), 1 = BGMW95
void oracle_ctx_build_table(void *hv, int which) {
    OracleHandle *h = (OracleHandle *)hv;
    if (h->g1) { if (which == 0) build_table_3nh(*h->g1); else build_table_bgmw(*h->g1); }
    else { if (which == 0) build_table_3nh(*h->g2); else build_table_bgmw(*h->g2); }
}
size_t oracle_ctx_table_len(void *hv, int which) {
    OracleHandle *h = (OracleHandle *)hv;
    if (h->g1) return which == 0 ? h->g1->table_3nh.size() : h->g1->table_bgmw.size();
    return which == 0 ? h->g2->table_3nh.size() : h->g2->table_bgmw.size();
}
const void *oracle_ctx_table(void *hv, int which) {
    OracleHandle *h = (OracleHandle *)hv;
    if (h->g1) return which == 0 ? (const void *)h->g1->table_3nh.data() : (const void *)h->g1->table_bgmw.data();
    return which == 0 ? (const void *)h->g2->table_3nh.data() : (const void *)h->g2->table_bgmw.data();
}
// import a table computed elsewhere (e.g. downloaded from the GPU after spot checks)
void oracle_ctx_load_table(void *hv, int which, const void *data, size_t count) {
    OracleHandle *h = (OracleHandle *)hv;
    if (h->g1) { auto &t = which == 0 ? h->g1->table_3nh : h->g1->table_bgmw; t.resize(count); memcpy(t.data(), data, count * sizeof(Aff<Fp>)); }
    else { auto &t = which == 0 ? h->g2->table_3nh : h->g2->table_bgmw; t.resize(count); memcpy(t.data(), data, count * sizeof(Aff<Fp2>)); }
}
long oracle_ctx_bucket_set(void *hv, int *out, long cap) {
    OracleHandle *h = (OracleHandle *)hv;
    const std::vector<int> &B = h->g1 ? h->g1->bucket_set : h->g2->bucket_set;
    if (out && (long)B.size() <= cap) memcpy(out, B.data(), B.size() * sizeof(int));
    return (long)B.size();
}
// digit hash table as (q+1) x 3 ints (m, b, alpha)
void oracle_ctx_hash(void *hv, int *out) {
    OracleHandle *h = (OracleHandle *)hv;
    const std::vector<Digit> &H = h->g1 ? h->g1->hash : h->g2->hash;
    memcpy(out, H.data(), H.size() * sizeof(Digit));
}
// per-scalar digits. kind 0: CHES (out_m signed +-m, out_b), kind 1: BGMW95 signed digits in out_m
// (after the r-a trick; out_b[0] = 1 if the trick fired)
void oracle_ctx_digits(void *hv, int kind, const uint64_t *scalar, int *out_m, int *out_b) {
    OracleHandle *h = (OracleHandle *)hv;
    U256 s; memcpy(s.d, scalar, 32);
    const Config &cfg = h->g1 ? h->g1->cfg : h->g2->cfg;
    if (kind == 0) { if (h->g1) mb_radixq(*h->g1, out_m, out_b, s); else mb_radixq(*h->g2, out_m, out_b, s); }
    else {
        bool trick = (cfg.n_exp == 13 || cfg.n_exp == 14 || cfg.n_exp == 16 || cfg.n_exp == 17);
        bool cond = trick && s.d[3] > (1ull << 62);
        if (cond) s = u256_sub(R_ORDER, s);
        qhalf_expr(out_m, s, cfg.e_bgmw, cfg.h_bgmw);
        out_b[0] = cond;
    }
}
// methods 1..4; out = affine struct (96 B G1 / 192 B G2, Montgomery limbs)
int oracle_ctx_msm(void *hv, int method, const uint64_t *scalars, void *out_affine, int faithful_bug) {
    OracleHandle *h = (OracleHandle *)hv;
    if (h->g1) return ctx_msm(*h->g1, method, (const U256 *)scalars, (Aff<Fp> *)out_affine, faithful_bug);
    return ctx_msm(*h->g2, method, (const U256 *)scalars, (Aff<Fp2> *)out_affine, faithful_bug);
}
// does scalar set hit SURVEY App. D-1 case (i)?  (penultimate entry -> bucket 0, last entry not)
int oracle_ctx_hits_last_element_bug(void *hv, const uint64_t *scalars) {
    OracleHandle *h = (OracleHandle *)hv;
    const Config &cfg = h->g1 ? h->g1->cfg : h->g2->cfg;
    size_t n = h->g1 ? h->g1->n : h->g2->n;
    std::vector<int> em(cfg.h), eb(cfg.h);
    U256 s; memcpy(s.d, scalars + 4 * (n - 1), 32);
    if (h->g1) mb_radixq(*h->g1, em.data(), eb.data(), s); else mb_radixq(*h->g2, em.data(), eb.data(), s);
    return eb[cfg.h - 2] == 0 && eb[cfg.h - 1] != 0;
}
void oracle_affine_serialize(int group, const void *aff, unsigned char *out) {
    if (group == 1) affine_serialize(out, *(const Aff<Fp> *)aff); else affine_serialize(out, *(const Aff<Fp2> *)aff);
}
// closed form for synthetic points: out affine = (sum s_i 2^(first+i+1) mod r) * G
void oracle_closed_form(int group, const uint64_t *scalars, size_t n, size_t first, void *out_affine, uint64_t *k_out) {
    U256 k = closed_form_scalar((const U256 *)scalars, n, first);
    if (k_out) memcpy(k_out, k.d, 32);
    if (group == 1) *(Aff<Fp> *)out_affine = scalar_mul_generator<Fp>(k); else *(Aff<Fp2> *)out_affine = scalar_mul_generator<Fp2>(k);
}
// generic naive MSM over arbitrary affine points (double-and-add each, sum) for small n
void oracle_naive_msm(int group, const void *points, const uint64_t *scalars, size_t n, void *out_affine) {
    auto run = [&](auto tag) {
        typedef decltype(tag) F;
        const Aff<F> *P = (const Aff<F> *)points;
        Jac<F> acc = jac_infinity<F>();
        for (size_t i = 0; i < n; i++) {
            Jac<F> t = jac_infinity<F>(), b = jac_from_affine(P[i]);
            for (int k = 255; k >= 0; k--) {
                t = jac_dadd(t, t);
                if ((scalars[4 * i + k / 64] >> (k % 64)) & 1) t = jac_dadd(t, b);
            }
            acc = jac_dadd(acc, t);
        }
        *(Aff<F> *)out_affine = jac_to_affine(acc);
    };
    if (group == 1) run(Fp{}); else run(Fp2{});
}
// sum of Jacobian partials (multi-GPU gather check): out affine
void oracle_sum_partials(int group, const void *partials, size_t count, void *out_affine) {
    if (group == 1) { Jac<Fp> a = jac_infinity<Fp>(); for (size_t i = 0; i < count; i++) a = jac_dadd(a, ((const Jac<Fp> *)partials)[i]); *(Aff<Fp> *)out_affine = jac_to_affine(a); }
    else { Jac<Fp2> a = jac_infinity<Fp2>(); for (size_t i = 0; i < count; i++) a = jac_dadd(a, ((const Jac<Fp2> *)partials)[i]); *(Aff<Fp2> *)out_affine = jac_to_affine(a); }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Reference arm: the SAME driver-level glue as above (digit conversion, pointer arrays: main_p1.cpp:192-436)
// but with the hot loops executed by the COMPILED REFERENCE (oracle/_ref/libblst_ref.so: x86-64 ADX assembly
// field arithmetic, blst_pN_tile_pippenger_d_CHES / _BGMW95 / blst_pNs_mult_pippenger / blst_pN_to_affine),
// loaded with dlopen. Used as bench.py's CPU baseline (kind "reference") and to cross-check the restatement.
// The reference is single-threaded; `threads` > 1 shards the points over host threads the way the upstream
// Rust/Go bindings do (bindings/rust/src/lib.rs:1840-1866): each thread runs the reference tile on its own
// contiguous slice with its own bucket array, the partial sums are added with blst_pN_add_or_double.
// ---------------------------------------------------------------------------------------------
#include <dlfcn.h>
struct RefFns {
    void *lib = nullptr;
    void (*tile_ches[2])(void *, const void *const *, size_t, const int *, const unsigned char *, void *, int *, int *, size_t, int);
    void (*tile_bgmw[2])(void *, const void *const *, size_t, const int *, const unsigned char *, void *, size_t);
    void (*mult_pip[2])(void *, const void *const *, size_t, const unsigned char *const *, size_t, void *);
    size_t (*pip_scratch[2])(size_t);
    void (*to_affine[2])(void *, const void *);
    void (*add_or_double[2])(void *, const void *, const void *);
    void (*integrate_ches[2])(void *, void *, int *, size_t, int);
};
static RefFns g_ref;
extern "C" int oracle_load_ref(const char *path) {
    if (g_ref.lib) return 0;
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    const char *pre[2] = {"blst_p1", "blst_p2"};
    for (int g = 0; g < 2; g++) {
        std::string p = pre[g];
        *(void **)&g_ref.tile_ches[g] = dlsym(h, (p + "_tile_pippenger_d_CHES").c_str());
        *(void **)&g_ref.tile_bgmw[g] = dlsym(h, (p + "_tile_pippenger_BGMW95").c_str());
        *(void **)&g_ref.mult_pip[g] = dlsym(h, (p + "s_mult_pippenger").c_str());
        *(void **)&g_ref.pip_scratch[g] = dlsym(h, (p + "s_mult_pippenger_scratch_sizeof").c_str());
        *(void **)&g_ref.to_affine[g] = dlsym(h, (p + "_to_affine").c_str());
        *(void **)&g_ref.add_or_double[g] = dlsym(h, (p + "_add_or_double").c_str());
        *(void **)&g_ref.integrate_ches[g] = dlsym(h, (p + "_integrate_buckets_accumulation_d_CHES").c_str());
        if (!g_ref.tile_ches[g] || !g_ref.tile_bgmw[g] || !g_ref.mult_pip[g] || !g_ref.pip_scratch[g] || !g_ref.to_affine[g] ||
            !g_ref.add_or_double[g] || !g_ref.integrate_ches[g]) { dlclose(h); return -2; }
    }
    g_ref.lib = h;
    return 0;
}
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// phases_ms: [0] glue (digit conversion + array fill), [1] reference tile calls (max over threads),
//            [2] bucket-reduction share of [1] (re-timed integrate on thread 0's buckets; CHES only), [3] combine + to_affine
template <class F> static int ref_msm_impl(const Ctx<F> &c, int gi, int method, const U256 *sc, Aff<F> *out, int threads, double *phases) {
    if (!g_ref.lib) return -10;
    int T = threads > 1 ? threads : 1;
    if ((size_t)T > c.n / 2) T = (int)std::max<size_t>(1, c.n / 2);
    std::vector<Jac<F>> part(T);
    std::vector<double> t_glue(T, 0), t_tile(T, 0);
    double t_reduce = 0;
    bool trick = (c.cfg.n_exp == 13 || c.cfg.n_exp == 14 || c.cfg.n_exp == 16 || c.cfg.n_exp == 17);
    auto work = [&](int t) {
        size_t lo = c.n * t / T, hi = c.n * (t + 1) / T, cnt = hi - lo;
        double t0 = now_ms();
        if (method == 1 || method == 2) {
            int h = c.cfg.h;
            size_t np = cnt * h;
            std::vector<int> scalars(np + 2, 0);
            std::vector<unsigned char> signs(np);
            std::vector<const void *> ptr(np);
            if (method == 1) {
                std::vector<int> em(h), eb(h);
                for (size_t i = 0; i < cnt; i++) {
                    mb_radixq(c, em.data(), eb.data(), sc[lo + i]);
                    for (int j = 0; j < h; j++) {
                        size_t idx = i * h + j, gidx = (lo + i) * h + j;
                        int m = em[j];
                        scalars[idx] = eb[j];
                        ptr[idx] = &c.table_3nh[3 * gidx + (m > 0 ? m : -m) - 1];
                        signs[idx] = m > 0 ? 0 : 1;
                    }
                }
            } else {
                for (size_t i = 0; i < cnt; i++) std_q_ary(&scalars[i * h], sc[lo + i], c.cfg.e, h);
                for (size_t i = 0; i < np; i++) {
                    Digit d = c.hash[scalars[i]];
                    scalars[i] = d.b;
                    signs[i] = (unsigned char)d.alpha;
                    if (d.alpha && i + 1 < np) ++scalars[i + 1];
                    ptr[i] = &c.table_3nh[3 * (lo * h + i) + d.m - 1];
                }
            }
            std::vector<Xyzz<F>> buckets(c.bucket_set.size());
            double t1 = now_ms();
            g_ref.tile_ches[gi](&part[t], ptr.data(), np, scalars.data(), signs.data(), buckets.data(),
                                const_cast<int *>(c.bucket_set.data()), const_cast<int *>(c.value_to_index.data()), c.bucket_set.size(), c.cfg.d);
            double t2 = now_ms();
            t_glue[t] = t1 - t0; t_tile[t] = t2 - t1;
            if (t == 0) {
                Jac<F> dummy;
                double r0 = now_ms();
                g_ref.integrate_ches[gi](&dummy, buckets.data(), const_cast<int *>(c.bucket_set.data()), c.bucket_set.size(), c.cfg.d);
                t_reduce = now_ms() - r0;
            }
        } else if (method == 3) {
            int h = c.cfg.h_bgmw, e = c.cfg.e_bgmw;
            size_t np = cnt * h;
            std::vector<int> scalars(np);
            std::vector<unsigned char> signs(np);
            std::vector<const void *> ptr(np);
            std::vector<int> ex(h);
            for (size_t i = 0; i < cnt; i++) {
                U256 aa = sc[lo + i];
                bool cond = trick && (aa.d[3] > (1ull << 62));
                if (cond) aa = u256_sub(R_ORDER, aa);
                qhalf_expr(ex.data(), aa, e, h);
                for (int j = 0; j < h; j++) {
                    size_t idx = i * h + j;
                    int v = ex[j];
                    ptr[idx] = &c.table_bgmw[(lo + i) * h + j];
                    if (v > 0) { scalars[idx] = v; signs[idx] = cond ? 1 : 0; }
                    else { scalars[idx] = -v; signs[idx] = cond ? 0 : 1; }
                }
            }
            std::vector<Xyzz<F>> buckets(((size_t)1 << (e - 1)) + 1);
            double t1 = now_ms();
            g_ref.tile_bgmw[gi](&part[t], ptr.data(), np, scalars.data(), signs.data(), buckets.data(), (size_t)e);
            double t2 = now_ms();
            t_glue[t] = t1 - t0; t_tile[t] = t2 - t1;
        } else {
            std::vector<const void *> pp(cnt);
            std::vector<const unsigned char *> sp(cnt);
            for (size_t i = 0; i < cnt; i++) { pp[i] = &c.fix_points[lo + i]; sp[i] = (const unsigned char *)&sc[lo + i]; }
            std::vector<uint64_t> scratch(g_ref.pip_scratch[gi](cnt) / 8 + 1);
            double t1 = now_ms();
            g_ref.mult_pip[gi](&part[t], pp.data(), cnt, sp.data(), 255, scratch.data());
            double t2 = now_ms();
            t_glue[t] = t1 - t0; t_tile[t] = t2 - t1;
        }
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    double t3 = now_ms();
    Jac<F> acc = part[0];
    for (int t = 1; t < T; t++) g_ref.add_or_double[gi](&acc, &acc, &part[t]);
    g_ref.to_affine[gi](out, &acc);
    double t4 = now_ms();
    if (phases) {
        phases[0] = *std::max_element(t_glue.begin(), t_glue.end());
        phases[1] = *std::max_element(t_tile.begin(), t_tile.end());
        phases[2] = t_reduce;
        phases[3] = t4 - t3;
    }
    return 0;
}
extern "C" int oracle_ctx_msm_ref(void *hv, int method, const uint64_t *scalars, void *out_affine, int threads, double *phases_ms) {
    OracleHandle *h = (OracleHandle *)hv;
    if (method < 1 || method > 4) return -1;
    if (h->g1) return ref_msm_impl(*h->g1, 0, method, (const U256 *)scalars, (Aff<Fp> *)out_affine, threads, phases_ms);
    return ref_msm_impl(*h->g2, 1, method, (const U256 *)scalars, (Aff<Fp2> *)out_affine, threads, phases_ms);
}
