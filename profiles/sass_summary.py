"""Opcode histogram + listing of the out-of-line field routines as ptxas emitted them for sm_100a.
usage: python profiles/sass_summary.py   (needs msm_blst_b200/csrc/build/engine_g1.o; writes profiles/r2_sass_*.txt)
The carrier kernels are the microbenchmark peak_fp_mul_kernel (calls fp_mul_fn) and ba_round_kernel<fp_t, true> (the hot loop)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "msm_blst_b200", "csrc", "build", "engine_g1.o")

def sass(fun):
    return subprocess.run(["cuobjdump", "-sass", "-fun", fun, OBJ], capture_output=True, text=True, check=True).stdout.splitlines()

def instrs(lines):
    out = []
    for l in lines:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip()))
    return out

def opcode(txt):
    t = txt.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0]

def histogram(ins):
    h = collections.Counter(opcode(t) for _, t in ins)
    reuse = sum(1 for _, t in ins if ".reuse" in t and opcode(t).startswith("IMAD.WIDE"))
    return h, reuse

def write(name, title, ins, listing=True):
    h, reuse = histogram(ins)
    wide = sum(v for k, v in h.items() if k.startswith("IMAD.WIDE"))
    with open(os.path.join(ROOT, "profiles", name), "w") as f:
        f.write("%s\n%d instructions; IMAD.WIDE* %d (with a .reuse operand flag: %d), IMAD.HI* %d, other IMAD* %d, IADD3* %d\n\n" % (
            title, len(ins), wide, reuse, sum(v for k, v in h.items() if k.startswith("IMAD.HI")),
            sum(v for k, v in h.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE") and not k.startswith("IMAD.HI")),
            sum(v for k, v in h.items() if k.startswith("IADD3"))))
        for k, v in sorted(h.items(), key=lambda kv: -kv[1]):
            f.write("  %-28s %5d\n" % (k, v))
        if listing:
            f.write("\n")
            for a, t in ins:
                f.write("/*%04x*/  %s\n" % (a, t))

rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
# fp_mul_fn: the body the two CALLs of peak_fp_mul_kernel point at, up to its RET
ins = instrs(sass("_ZN7msmb20018peak_fp_mul_kernelEPNS_4fp_tEi"))
calls = [int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) for _, t in ins if opcode(t).startswith("CALL")]
start = calls[0]
body = [(a, t) for a, t in ins if a >= start]
end = next(i for i, (_, t) in enumerate(body) if opcode(t).startswith("RET"))
write("r2_sass_fp_mul_fn.txt", "fp_mul_fn (Montgomery multiplication, csrc/fp.cuh) inside peak_fp_mul_kernel, sm_100a, commit %s" % rev, body[: end + 1])
# the batch-affine hot kernel: histogram of the whole kernel including its out-of-line routines
for fun, out in (("_ZN7msmb20015ba_round_kernelINS_4fp_tELb0EEEvNS_5ba_ioIT_EEPK5uint4jPK5uint2jPS5_mNS_7BaSchedEj", "r2_sass_ba_round_kernel_fp.txt"),):
    try:
        ins = instrs(sass(fun))
        write(out, "ba_round_kernel<fp_t, false> (rounds >= 1) incl. fp_mul_fn / fp_sqr_fn / fp_inv_warp_fn copies, sm_100a, commit %s" % rev, ins, listing=False)
    except Exception as ex:
        print("skipped", fun, ex)
print(open(os.path.join(ROOT, "profiles", "r2_sass_fp_mul_fn.txt")).read()[:1500])
