"""ctypes access to the ORACLE (test infrastructure): the CPU restatement `oracle/libmsm_oracle.so`
and, when present, the compiled reference under `oracle/_ref/` (libblst_ref.so, refdrv_p{1,2}.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

FP_BYTES = 48
AFF_BYTES = {1: 96, 2: 192}
JAC_BYTES = {1: 144, 2: 288}
XYZZ_BYTES = {1: 192, 2: 384}
SER_BYTES = {1: 96, 2: 192}
R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
P_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB


def build_oracle():
    """(Re)build the restatement and, if /root/reference is present, the compiled reference."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all", "ref"])


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libmsm_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.oracle_ctx_create.restype = C.c_void_p
        lib.oracle_ctx_create.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
        lib.oracle_ctx_destroy.argtypes = [C.c_void_p]
        lib.oracle_ctx_set_threads.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_ctx_init_fix_points.argtypes = [C.c_void_p, C.c_size_t]
        lib.oracle_ctx_set_points.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_ctx_points.restype = C.c_void_p
        lib.oracle_ctx_points.argtypes = [C.c_void_p]
        lib.oracle_ctx_build_table.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_ctx_table_len.restype = C.c_size_t
        lib.oracle_ctx_table_len.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_ctx_table.restype = C.c_void_p
        lib.oracle_ctx_table.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_ctx_load_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        lib.oracle_ctx_bucket_set.restype = C.c_long
        lib.oracle_ctx_bucket_set.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        lib.oracle_ctx_hash.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_ctx_digits.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_ctx_msm.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.oracle_ctx_hits_last_element_bug.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_bucket_set.restype = C.c_long
        lib.oracle_bucket_set.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_long]
        lib.oracle_bucket_set_check.argtypes = [C.c_int, C.c_int]
        lib.oracle_fp_op.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_fp2_op.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_point_op.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_fp_to_mont.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_fp_from_mont.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.oracle_gen_scalars.argtypes = [C.c_uint64, C.c_size_t, C.c_void_p]
        lib.oracle_config.argtypes = [C.c_char_p, C.c_void_p]
        lib.oracle_affine_serialize.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        lib.oracle_closed_form.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.oracle_naive_msm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.oracle_sum_partials.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.oracle_load_ref.argtypes = [C.c_char_p]
        lib.oracle_ctx_msm_ref.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _oracle = lib
    return _oracle


def has_ref():
    return os.path.exists(os.path.join(REF_DIR, "libblst_ref.so"))


_blst = None


def blst_ref():
    """The compiled reference library (src/server.c + build/assembly.S)."""
    global _blst
    if _blst is None:
        adx = "adx" in open("/proc/cpuinfo").read()
        name = "libblst_ref.so" if adx else "libblst_ref_noadx.so"
        _blst = C.CDLL(os.path.join(REF_DIR, name))
    return _blst


def ref_lib_path():
    adx = "adx" in open("/proc/cpuinfo").read()
    return os.path.join(REF_DIR, "libblst_ref.so" if adx else "libblst_ref_noadx.so")


def load_ref_into_oracle():
    rc = oracle().oracle_load_ref(ref_lib_path().encode())
    if rc != 0:
        raise RuntimeError("cannot load compiled reference %s (rc=%d)" % (ref_lib_path(), rc))


_refdrv = {}


def refdrv(group):
    """The reference driver main_p{group}.cpp (config 10) as a library; runs its init on first use."""
    if group not in _refdrv:
        lib = C.CDLL(os.path.join(REF_DIR, "refdrv_p%d.so" % group))
        lib.refdrv_fix_points.restype = C.c_void_p
        lib.refdrv_table.restype = C.c_void_p
        lib.refdrv_table.argtypes = [C.c_int]
        lib.refdrv_bucket_set.restype = C.c_void_p
        lib.refdrv_hash.restype = C.c_void_p
        lib.refdrv_msm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.refdrv_digits.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        # silence the driver's std::cout chatter during init
        fd = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            lib.refdrv_init()
        finally:
            os.dup2(fd, 1)
            os.close(fd)
            os.close(devnull)
        _refdrv[group] = lib
    return _refdrv[group]


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def gen_scalars(seed, n):
    out = np.empty((n, 4), dtype=np.uint64)
    oracle().oracle_gen_scalars(seed, n, ptr(out))
    return out


def scalars_to_ints(sc):
    return [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in sc]


def serialize(group, aff):
    aff = np.ascontiguousarray(aff)
    out = np.empty(SER_BYTES[group], dtype=np.uint8)
    oracle().oracle_affine_serialize(group, ptr(aff), ptr(out))
    return out.tobytes()


def config(name):
    v = (C.c_int * 9)()
    if oracle().oracle_config(name.encode(), v) != 0:
        raise KeyError(name)
    keys = ["n_exp", "e", "h", "a", "d", "bsize", "e_bgmw", "h_bgmw", "window"]
    return dict(zip(keys, list(v)))


class OracleCtx:
    """Reference driver state (main_p1.cpp globals) restated: fixed points, bucket set, tables, 4 methods."""

    def __init__(self, group, cfgname, n=None, first=0, threads=1):
        self.group = group
        self.cfg = config(cfgname)
        self.n = n if n is not None else 1 << self.cfg["n_exp"]
        self.h = oracle().oracle_ctx_create(group, cfgname.encode(), self.n)
        assert self.h
        oracle().oracle_ctx_set_threads(self.h, threads)
        self.first = first

    def init_fix_points(self):
        oracle().oracle_ctx_init_fix_points(self.h, self.first)

    def set_points(self, pts):
        pts = np.ascontiguousarray(pts)
        assert pts.nbytes == self.n * AFF_BYTES[self.group]
        oracle().oracle_ctx_set_points(self.h, ptr(pts))

    def points(self):
        p = oracle().oracle_ctx_points(self.h)
        nb = self.n * AFF_BYTES[self.group]
        return np.frombuffer((C.c_ubyte * nb).from_address(p), dtype=np.uint8).copy()

    def build_table(self, which):
        oracle().oracle_ctx_build_table(self.h, which)

    def table(self, which):
        cnt = oracle().oracle_ctx_table_len(self.h, which)
        p = oracle().oracle_ctx_table(self.h, which)
        nb = cnt * AFF_BYTES[self.group]
        return np.frombuffer((C.c_ubyte * nb).from_address(p), dtype=np.uint8).copy()

    def load_table(self, which, data):
        data = np.ascontiguousarray(data)
        oracle().oracle_ctx_load_table(self.h, which, ptr(data), data.nbytes // AFF_BYTES[self.group])

    def bucket_set(self):
        n = oracle().oracle_ctx_bucket_set(self.h, None, 0)
        out = np.empty(n, dtype=np.int32)
        oracle().oracle_ctx_bucket_set(self.h, ptr(out), n)
        return out

    def hash_table(self):
        q = 1 << self.cfg["e"]
        out = np.empty((q + 1, 3), dtype=np.int32)
        oracle().oracle_ctx_hash(self.h, ptr(out))
        return out

    def digits(self, kind, scalar):
        h = self.cfg["h"] if kind == 0 else self.cfg["h_bgmw"]
        m = np.zeros(h + 1, dtype=np.int32)
        b = np.zeros(h + 1, dtype=np.int32)
        scalar = np.ascontiguousarray(scalar, dtype=np.uint64)
        oracle().oracle_ctx_digits(self.h, kind, ptr(scalar), ptr(m), ptr(b))
        return m[:h], b[:h] if kind == 0 else b[:1]

    def msm(self, method, scalars, faithful_bug=False):
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        assert scalars.shape == (self.n, 4)
        out = np.zeros(AFF_BYTES[self.group], dtype=np.uint8)
        rc = oracle().oracle_ctx_msm(self.h, method, ptr(scalars), ptr(out), int(faithful_bug))
        assert rc == 0
        return out

    def msm_ref(self, method, scalars, threads=1):
        """Same glue, hot loops run by the compiled reference (oracle/_ref/libblst_ref.so). Returns (affine, phases_ms)."""
        load_ref_into_oracle()
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        assert scalars.shape == (self.n, 4)
        out = np.zeros(AFF_BYTES[self.group], dtype=np.uint8)
        ph = np.zeros(4, dtype=np.float64)
        rc = oracle().oracle_ctx_msm_ref(self.h, method, ptr(scalars), ptr(out), threads, ptr(ph))
        assert rc == 0, rc
        return out, dict(glue=float(ph[0]), tile=float(ph[1]), reduce=float(ph[2]), finish=float(ph[3]))

    def hits_bug(self, scalars):
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        return bool(oracle().oracle_ctx_hits_last_element_bug(self.h, ptr(scalars)))

    def close(self):
        if self.h:
            oracle().oracle_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def closed_form(group, scalars, first=0):
    """(sum s_i 2^(first+i+1) mod r) * G as affine Montgomery struct bytes + k."""
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    out = np.zeros(AFF_BYTES[group], dtype=np.uint8)
    k = np.zeros(4, dtype=np.uint64)
    oracle().oracle_closed_form(group, ptr(scalars), scalars.shape[0], first, ptr(out), ptr(k))
    kint = int(k[0]) | int(k[1]) << 64 | int(k[2]) << 128 | int(k[3]) << 192
    return out, kint
