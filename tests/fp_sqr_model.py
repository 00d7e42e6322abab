"""Limb-level model (PTX carry-flag semantics: add.cc / addc / mad.lo.cc / madc.hi.cc) of the dedicated Montgomery squaring
fp_sqr_inline in msm_blst_b200/csrc/fp.cuh (replaces sqr_fp -> sqrx_mont_384, reference src/fields.h:40). Every place
where the CUDA code drops a carry (addc / madc.hi without .cc) asserts here that the carry is zero. Run by
tests/test_fp_sqr_model.py; `python tests/fp_sqr_model.py` runs the long version."""
import random
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
M32 = 0xffffffff
PL = [(P >> (32 * i)) & M32 for i in range(12)]
INV32 = 0xfffcfffd
assert (PL[0] * INV32 + 1) & M32 == 0 or ((-pow(P, -1, 1 << 32)) & M32) == INV32

class CC:
    c = 0
cc = CC()
def add_cc(a, b):
    s = a + b; cc.c = s >> 32; return s & M32
def addc_cc(a, b):
    s = a + b + cc.c; cc.c = s >> 32; return s & M32
def addc(a, b):
    s = a + b + cc.c; return s & M32   # carry out dropped (must be provably zero)
def addc_checked(a, b):
    s = a + b + cc.c; assert s >> 32 == 0, "carry lost"; return s & M32
def mad_lo_cc(a, b, c):
    s = ((a * b) & M32) + c; cc.c = s >> 32; return s & M32
def madc_lo_cc(a, b, c):
    s = ((a * b) & M32) + c + cc.c; cc.c = s >> 32; return s & M32
def madc_hi_cc(a, b, c):
    s = ((a * b) >> 32) + c + cc.c; cc.c = s >> 32; return s & M32
def madc_hi(a, b, c):
    s = ((a * b) >> 32) + c + cc.c; assert s >> 32 == 0, "carry lost in madc_hi"; return s & M32

def fp_row_reduce(ev, od):
    m = (ev[0] * INV32) & M32
    od[0] = mad_lo_cc(PL[1], m, od[0])
    od[1] = madc_hi_cc(PL[1], m, od[1])
    for k in range(2, 10, 2):
        od[k] = madc_lo_cc(PL[k + 1], m, od[k])
        od[k + 1] = madc_hi_cc(PL[k + 1], m, od[k + 1])
    od[10] = madc_lo_cc(PL[11], m, od[10])
    od[11] = madc_hi(PL[11], m, od[11])
    ev[0] = mad_lo_cc(PL[0], m, ev[0])
    ev[1] = madc_hi_cc(PL[0], m, ev[1])
    for k in range(2, 12, 2):
        ev[k] = madc_lo_cc(PL[k], m, ev[k])
        ev[k + 1] = madc_hi_cc(PL[k], m, ev[k + 1])
    od[11] = addc_checked(od[11], 0)
    assert ev[0] == 0

def sqr(a):
    # --- off-diagonal products: X[par] holds products a_i a_j (i < j) with (i + j) % 2 == par at words (i+j, i+j+1) ---
    X = [[0] * 24, [0] * 24]
    for i in range(11):
        # chain ending at j = 11 (fresh top word), then chain ending at j = 10 (one ripple word)
        for jend in (11, 10):
            js = [j for j in range(i + 1, jend + 1) if (jend - j) % 2 == 0]
            if not js:
                continue
            par = (i + jend) % 2
            x = X[par]
            first = True
            for j in js:
                p = i + j
                if first:
                    x[p] = mad_lo_cc(a[i], a[j], x[p]); first = False
                else:
                    x[p] = madc_lo_cc(a[i], a[j], x[p])
                if j == jend and jend == 11:
                    assert x[p + 1] == 0
                    x[p + 1] = madc_hi(a[i], a[j], 0)
                else:
                    x[p + 1] = madc_hi_cc(a[i], a[j], x[p + 1])
            if jend == 10:
                assert x[i + 12] == 0
                x[i + 12] = addc_checked(0, 0)
    # --- t = X[0] + X[1] ---
    t = [0] * 24
    t[0] = add_cc(X[0][0], X[1][0])
    for k in range(1, 23):
        t[k] = addc_cc(X[0][k], X[1][k])
    t[23] = addc_checked(X[0][23], X[1][23])
    ref_off = sum(a[i] * a[j] << (32 * (i + j)) for i in range(12) for j in range(i + 1, 12))
    assert sum(v << (32 * k) for k, v in enumerate(t)) == ref_off
    # --- double ---
    d = [0] * 24
    for k in range(23, 0, -1):
        d[k] = ((t[k] << 1) | (t[k - 1] >> 31)) & M32
    d[0] = (t[0] << 1) & M32
    assert t[23] >> 31 == 0
    # --- add the squares ---
    d[0] = mad_lo_cc(a[0], a[0], d[0])
    d[1] = madc_hi_cc(a[0], a[0], d[1])
    for i in range(1, 12):
        d[2 * i] = madc_lo_cc(a[i], a[i], d[2 * i])
        if i < 11:
            d[2 * i + 1] = madc_hi_cc(a[i], a[i], d[2 * i + 1])
        else:
            d[23] = madc_hi(a[i], a[i], d[23])
    A = sum(v << (32 * k) for k, v in enumerate(a))
    assert sum(v << (32 * k) for k, v in enumerate(d)) == A * A
    # --- Montgomery reduction: window (ev even-aligned, od odd-aligned), roles swap every round ---
    ev = d[:12]
    od = [0] * 12
    fp_row_reduce(ev, od)
    for s in range(1, 12):
        # shift the window down one limb: new even-aligned = old od, new odd-aligned = old ev shifted, inject d[11 + s]
        nev, nod = od, ev
        nev[0] = add_cc(nev[0], nod[1])
        for k in range(0, 10):
            nod[k] = addc_cc(nod[k + 2], 0)
        nod[10] = addc_cc(d[11 + s], 0)
        nod[11] = addc_checked(0, 0)
        fp_row_reduce(nev, nod)
        ev, od = nev, nod
    # epilogue: last shift, inject d[23]
    r = [0] * 12
    r[0] = add_cc(od[0], ev[1])
    for k in range(1, 11):
        r[k] = addc_cc(od[k], ev[k + 1])
    r[11] = addc_cc(od[11], d[23])
    assert cc.c == 0
    val = sum(v << (32 * k) for k, v in enumerate(r))
    if val >= P:
        val -= P
    return val

R = 1 << 384
Rinv = pow(R, -1, P)


def check(ncases, seed=1):
    rnd = random.Random(seed)
    cases = [0, 1, P - 1, P - 2, (1 << 380) - 1, R % P, (R * R) % P]
    cases += [rnd.randrange(P) for _ in range(ncases)]
    cases += [int("f" * 8 * k, 16) % P for k in range(1, 12)]
    cases += [((1 << 32 * k) - 1) << (32 * j) for k in range(1, 6) for j in range(0, 7)]
    for x in cases:
        x %= P
        a = [(x >> (32 * i)) & M32 for i in range(12)]
        assert sqr(a) == (x * x * Rinv) % P, hex(x)
    return len(cases)


if __name__ == "__main__":
    print("ok", check(3000))
