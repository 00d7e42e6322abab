"""ctypes binding of libmsm_b200.so and a host-side mirror of the reference driver (main_p1.cpp / main_p2.cpp).

Method names follow the reference one-for-one so parity tests read like the reference's own driver:
    init_fix_point_list()                       main_p1.cpp:52
    init_pippenger_CHES_q_over_5()              main_p1.cpp:128
    init_pippenger_BGMW95()                     main_p1.cpp:94
    pippenger_variant_q_over_5_CHES(scalars)    main_p1.cpp:192
    pippenger_variant_q_over_5_CHES_integral_scalar_conversion(scalars)   main_p1.cpp:249
    pippenger_variant_BGMW95(scalars)           main_p1.cpp:294
    pippenger_blst_built_in(scalars)            main_p1.cpp:400
All computation happens in the CUDA library; this file moves pointers.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmsm_b200.so")

AFF_BYTES = {1: 96, 2: 192}
JAC_BYTES = {1: 144, 2: 288}
XYZZ_BYTES = {1: 192, 2: 384}

CHES, CHES_INTEGRAL, BGMW95, PIPPENGER = 1, 2, 3, 4


class MsmB200Error(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("n_exp", "e", "h", "a", "d", "bsize", "e_bgmw", "h_bgmw")]


def build_library(force=False):
    """Compile csrc/ for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-s", "-j4", "-C", os.path.join(HERE, "csrc"), "all"])
    return LIB_PATH


_lib = None


def lib():
    """Load the CUDA library. Fails loudly if it is missing: there is no other implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MsmB200Error(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(msm_blst_b200 has no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, sz, ci = C.c_void_p, C.c_size_t, C.c_int
        L.msmb200_config_lookup.argtypes = [C.c_char_p, C.POINTER(_Config)]
        L.msmb200_host_bucket_set.argtypes = [ci, ci, vp, C.c_long]
        L.msmb200_host_bucket_set.restype = C.c_long
        L.msmb200_host_digit_table.argtypes = [ci, ci, vp]
        L.msmb200_host_bucket_set_check.argtypes = [C.c_long, C.c_long, vp]
        L.msmb200_pippenger_window_size.argtypes = [sz]
        L.msmb200_pippenger_window_size.restype = sz
        L.msmb200_ctx_create.argtypes = [C.POINTER(vp), ci, C.POINTER(_Config), sz, ci]
        L.msmb200_ctx_destroy.argtypes = [vp]
        L.msmb200_ctx_destroy.restype = None
        L.msmb200_last_error.argtypes = [vp]
        L.msmb200_last_error.restype = C.c_char_p
        L.msmb200_set_stream.argtypes = [vp, vp]
        L.msmb200_set_points.argtypes = [vp, vp]
        L.msmb200_set_accumulator.argtypes = [vp, ci]
        L.msmb200_set_tuning.argtypes = [vp, C.c_char_p, ci]
        L.msmb200_set_reducer.argtypes = [vp, ci]
        L.msmb200_table_save.argtypes = [vp, ci, C.c_char_p, ci]
        L.msmb200_table_load.argtypes = [vp, ci, C.c_char_p]
        L.msmb200_set_bucket_shard.argtypes = [vp, ci, ci]
        L.msmb200_generate_fix_points.argtypes = [vp, sz]
        L.msmb200_table_build_ches.argtypes = [vp]
        L.msmb200_table_build_bgmw95.argtypes = [vp]
        L.msmb200_download.argtypes = [vp, ci, sz, sz, vp]
        L.msmb200_bucket_set.argtypes = [vp, vp, C.c_long]
        L.msmb200_bucket_set.restype = C.c_long
        L.msmb200_msm.argtypes = [vp, ci, vp, vp]
        L.msmb200_msm_device.argtypes = [vp, ci, vp, vp]
        L.msmb200_msm_partial_device.argtypes = [vp, ci, vp, vp]
        L.msmb200_sum_partials_device.argtypes = [vp, vp, ci, vp]
        L.msmb200_msm_bits_layout.argtypes = [vp, ci, vp]
        L.msmb200_msm_bits_device.argtypes = [vp, ci, vp, vp]
        L.msmb200_combine_bits_device.argtypes = [vp, vp, ci, vp, vp]
        L.msmb200_affine_serialize.argtypes = [ci, vp, vp]
        L.msmb200_last_timings.argtypes = [vp, vp]
        L.msmb200_last_launches.argtypes = [vp]
        L.msmb200_last_accumulator.argtypes = [vp]
        L.msmb200_measure_peaks.argtypes = [ci, vp, vp]
        L.msmb200_measure_peaks_ex.argtypes = [ci, vp]
        L.msmb200_test_field_op.argtypes = [ci, ci, ci, vp, vp, vp, sz]
        L.msmb200_test_point_op.argtypes = [ci, ci, ci, vp, vp, vp, vp, sz]
        L.msmb200_test_digits.argtypes = [vp, ci, vp, sz, vp, vp]
        L.msmb200_blst_p1s_mult_pippenger_scratch_sizeof.argtypes = [sz]
        L.msmb200_blst_p1s_mult_pippenger_scratch_sizeof.restype = sz
        L.msmb200_blst_p2s_mult_pippenger_scratch_sizeof.argtypes = [sz]
        L.msmb200_blst_p2s_mult_pippenger_scratch_sizeof.restype = sz
        for g in (1, 2):
            f = getattr(L, "msmb200_blst_p%ds_mult_pippenger" % g)
            f.argtypes = [vp, vp, sz, vp, sz, vp]
            f.restype = None
            f = getattr(L, "msmb200_blst_p%ds_add" % g)
            f.argtypes = [vp, vp, sz]
            f.restype = None
            f = getattr(L, "msmb200_blst_p%d_tile_pippenger_d_CHES" % g)
            f.argtypes = [vp, vp, sz, vp, vp, vp, vp, vp, sz, ci]
            f.restype = None
            f = getattr(L, "msmb200_blst_p%d_tile_pippenger_BGMW95" % g)
            f.argtypes = [vp, vp, sz, vp, vp, vp, sz]
            f.restype = None
            f = getattr(L, "msmb200_blst_p%d_construct_nh_scalars_nh_points" % g)
            f.argtypes = [vp, vp, vp, sz, vp, vp]
            f.restype = None
        L.msmb200_blst_register_table.argtypes = [ci, vp, sz]
        L.msmb200_blst_last_call_ms.argtypes = [ci]
        L.msmb200_blst_last_call_ms.restype = C.c_double
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def config_lookup(name):
    cfg = _Config()
    if lib().msmb200_config_lookup(str(name).encode(), C.byref(cfg)) != 0:
        raise KeyError("unknown configuration %r" % (name,))
    return cfg


def measure_peaks(device=0):
    """(IMAD.WIDE multiply-accumulates/s, dependent fp_mul/s) measured on the device right now."""
    a, b = C.c_double(), C.c_double()
    rc = lib().msmb200_measure_peaks(device, C.byref(a), C.byref(b))
    if rc:
        raise MsmB200Error("measure_peaks failed (%d)" % rc)
    return a.value, b.value


def measure_peaks_ex(device=0):
    """dict: IMAD.WIDE multiply-accumulates/s, dependent fp_mul/s, multiply-accumulates per clock per SM, SM clock (MHz)."""
    out = (C.c_double * 4)()
    rc = lib().msmb200_measure_peaks_ex(device, out)
    if rc:
        raise MsmB200Error("measure_peaks_ex failed (%d)" % rc)
    return {"imad_wide_macs_per_s": out[0], "dependent_fp_mul_per_s": out[1], "macs_per_clk_per_sm": out[2], "sm_mhz_during_microbench": out[3]}


def host_bucket_set(e, a):
    n = lib().msmb200_host_bucket_set(e, a, None, 0)
    if n < 0:
        raise MsmB200Error("host_bucket_set(%d, %d) failed" % (e, a))
    out = np.empty(n, dtype=np.int32)
    lib().msmb200_host_bucket_set(e, a, _ptr(out), n)
    return out


def host_bucket_set_check(q, a):
    """check_bucket_set_validity + max_gap of the reference's parameter tool for radix q and leading term a (host only)."""
    out = (C.c_long * 5)()
    if lib().msmb200_host_bucket_set_check(int(q), int(a), out) != 0:
        raise MsmB200Error("host_bucket_set_check(%d, %d): arguments out of range" % (q, a))
    return {"valid": bool(out[0]), "size": int(out[1]), "max_gap": int(out[2]), "leading_ok": bool(out[3]), "first_uncovered": int(out[4])}


def host_digit_table(e, a):
    out = np.empty(((1 << e) + 1, 3), dtype=np.int32)
    if lib().msmb200_host_digit_table(e, a, _ptr(out)) != 0:
        raise MsmB200Error("host_digit_table(%d, %d) failed" % (e, a))
    return out


def affine_serialize(group, aff):
    aff = np.ascontiguousarray(aff, dtype=np.uint8)
    out = np.empty(AFF_BYTES[group], dtype=np.uint8)
    rc = lib().msmb200_affine_serialize(group, _ptr(aff), _ptr(out))
    if rc:
        raise MsmB200Error("affine_serialize failed (%d)" % rc)
    return out.tobytes()


def test_field_op(field, op, a, b=None, device=0):
    a = np.ascontiguousarray(a)
    n = a.nbytes // (48 * field)
    out = np.empty_like(a)
    bb = np.ascontiguousarray(b) if b is not None else None
    rc = lib().msmb200_test_field_op(device, field, op, _ptr(a), _ptr(bb) if bb is not None else None, _ptr(out), n)
    if rc:
        raise MsmB200Error("test_field_op failed (%d): no usable CUDA device?" % rc)
    return out


def test_point_op(group, op, a, b=None, flags=None, device=0):
    A, J, X = AFF_BYTES[group], JAC_BYTES[group], XYZZ_BYTES[group]
    in_b = [J, J, X, X, X, J, X, X, X, J][op]
    out_b = [J, J, X, X, J, A, X, X, X, A][op]
    a = np.ascontiguousarray(a, dtype=np.uint8)
    n = a.nbytes // in_b
    out = np.empty(n * out_b, dtype=np.uint8)
    bb = np.ascontiguousarray(b, dtype=np.uint8) if b is not None else None
    ff = np.ascontiguousarray(flags, dtype=np.uint8) if flags is not None else None
    rc = lib().msmb200_test_point_op(device, group, op, _ptr(a), _ptr(bb) if bb is not None else None,
                                     _ptr(ff) if ff is not None else None, _ptr(out), n)
    if rc:
        raise MsmB200Error("test_point_op failed (%d)" % rc)
    return out


class MsmContext:
    """Device-resident state of the reference driver for one group and one point shard."""

    def __init__(self, group, config, npoints=None, device=0, first=0):
        self.group = group
        self.cfg = config_lookup(config) if not isinstance(config, _Config) else config
        self.n = int(npoints) if npoints is not None else 1 << self.cfg.n_exp
        self.first = int(first)
        self.device = device
        self._h = C.c_void_p()
        rc = lib().msmb200_ctx_create(C.byref(self._h), group, C.byref(self.cfg), self.n, device)
        if rc:
            raise MsmB200Error("ctx_create failed (%d): %s" % (rc, lib().msmb200_last_error(None).decode()))

    # -- plumbing --
    def _ck(self, rc):
        if rc:
            raise MsmB200Error("msm_b200 error %d: %s" % (rc, lib().msmb200_last_error(self._h).decode()))

    def close(self):
        if getattr(self, "_h", None):
            lib().msmb200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self._ck(lib().msmb200_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_accumulator(self, mode):
        """0 default, 1 XYZZ work items, 2 batch-affine rounds (identical results)."""
        self._ck(lib().msmb200_set_accumulator(self._h, int(mode)))

    def set_tuning(self, key, value):
        """Performance knobs ("ba_batch_max", "ba_batch", "item_len"); results never change."""
        self._ck(lib().msmb200_set_tuning(self._h, str(key).encode(), int(value)))

    def table_save(self, which, path, fmt=1):
        """which: 0 fixed points, 1 CHES 3nh table, 2 BGMW95 table; fmt 0 raw Montgomery layout, 1 blst_pN_affine_serialize."""
        self._ck(lib().msmb200_table_save(self._h, int(which), str(path).encode(), int(fmt)))

    def table_load(self, which, path):
        """Load (and validate entry by entry on the device) what table_save wrote; raises on a mismatching header or an invalid entry."""
        self._ck(lib().msmb200_table_load(self._h, int(which), str(path).encode()))

    def set_reducer(self, mode):
        """0 default, 1 chunked running sums (reference's tmp_d[] form), 2 digit splitting (identical results)."""
        self._ck(lib().msmb200_set_reducer(self._h, int(mode)))

    def set_bucket_shard(self, rank, world):
        """Bucket-range sharding: this context (holding ALL points) handles slice `rank` of `world` of the buckets."""
        self._ck(lib().msmb200_set_bucket_shard(self._h, int(rank), int(world)))

    # -- reference driver mirror --
    def init_fix_point_list(self):
        self._ck(lib().msmb200_generate_fix_points(self._h, self.first))

    def set_points(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.uint8)
        assert pts.nbytes == self.n * AFF_BYTES[self.group]
        self._ck(lib().msmb200_set_points(self._h, _ptr(pts)))

    def init_pippenger_CHES_q_over_5(self):
        self._ck(lib().msmb200_table_build_ches(self._h))

    def init_pippenger_BGMW95(self):
        self._ck(lib().msmb200_table_build_bgmw95(self._h))

    def _msm(self, method, scalars):
        scalars = np.ascontiguousarray(scalars)
        assert scalars.nbytes == self.n * 32, "scalars must be npoints x 32 bytes"
        out = np.zeros(AFF_BYTES[self.group], dtype=np.uint8)
        self._ck(lib().msmb200_msm(self._h, method, _ptr(scalars), _ptr(out)))
        return out

    def pippenger_variant_q_over_5_CHES(self, scalars):
        return self._msm(CHES, scalars)

    def pippenger_variant_q_over_5_CHES_integral_scalar_conversion(self, scalars):
        return self._msm(CHES_INTEGRAL, scalars)

    def pippenger_variant_BGMW95(self, scalars):
        return self._msm(BGMW95, scalars)

    def pippenger_blst_built_in(self, scalars):
        return self._msm(PIPPENGER, scalars)

    def msm(self, method, scalars):
        return self._msm(method, scalars)

    # -- device-resident / multi-GPU legs (raw device pointers, e.g. torch tensor .data_ptr()) --
    def msm_device(self, method, scalars_dev_ptr):
        out = np.zeros(AFF_BYTES[self.group], dtype=np.uint8)
        self._ck(lib().msmb200_msm_device(self._h, method, C.c_void_p(scalars_dev_ptr), _ptr(out)))
        return out

    def msm_partial_device(self, method, scalars_dev_ptr, out_jac_dev_ptr):
        self._ck(lib().msmb200_msm_partial_device(self._h, method, C.c_void_p(scalars_dev_ptr), C.c_void_p(out_jac_dev_ptr)))

    def sum_partials_device(self, partials_dev_ptr, count):
        out = np.zeros(AFF_BYTES[self.group], dtype=np.uint8)
        self._ck(lib().msmb200_sum_partials_device(self._h, C.c_void_p(partials_dev_ptr), count, _ptr(out)))
        return out

    def msm_bits_layout(self, method):
        """(windows, bit positions per window, doublings between windows) of the per-bit sums, or None when this context
        cannot hand them out (chunked reducer, bucket-range sharding, window layouts the digit splitting does not cover)."""
        out = (C.c_uint32 * 3)()
        if lib().msmb200_msm_bits_layout(self._h, method, out) != 0:
            return None
        return tuple(int(x) for x in out)

    def msm_bits_device(self, method, scalars_dev_ptr, out_bits_dev_ptr):
        self._ck(lib().msmb200_msm_bits_device(self._h, method, C.c_void_p(scalars_dev_ptr), C.c_void_p(out_bits_dev_ptr)))

    def combine_bits_device(self, gathered_dev_ptr, world, layout):
        out = np.zeros(AFF_BYTES[self.group], dtype=np.uint8)
        lay = (C.c_uint32 * 3)(*layout)
        self._ck(lib().msmb200_combine_bits_device(self._h, C.c_void_p(gathered_dev_ptr), world, lay, _ptr(out)))
        return out

    # -- introspection --
    def download(self, which, first=0, count=None):
        total = [self.n, self.n * self.cfg.h * 3, self.n * self.cfg.h_bgmw][which]
        count = total - first if count is None else count
        out = np.empty(count * AFF_BYTES[self.group], dtype=np.uint8)
        self._ck(lib().msmb200_download(self._h, which, first, count, _ptr(out)))
        return out

    def bucket_set(self):
        n = lib().msmb200_bucket_set(self._h, None, 0)
        out = np.empty(n, dtype=np.int32)
        lib().msmb200_bucket_set(self._h, _ptr(out), n)
        return out

    def digits(self, kind, scalars):
        scalars = np.ascontiguousarray(scalars)
        n = scalars.nbytes // 32
        per = self.cfg.h if kind == 0 else self.cfg.h_bgmw if kind == 1 else self.pippenger_tiles()
        keys = np.empty(n * per, dtype=np.uint32)
        vals = np.empty(n * per, dtype=np.uint32)
        self._ck(lib().msmb200_test_digits(self._h, kind, _ptr(scalars), n, _ptr(keys), _ptr(vals)))
        return keys.reshape(n, per), vals.reshape(n, per)

    def pippenger_window(self):
        wbits = self.n.bit_length() - 1
        return wbits - 3 if wbits > 12 else (wbits - 2 if wbits > 4 else (2 if wbits else 1))

    def pippenger_tiles(self):
        return 255 // self.pippenger_window() + 1

    def last_timings(self):
        out = np.zeros(6, dtype=np.float32)
        self._ck(lib().msmb200_last_timings(self._h, _ptr(out)))
        keys = ["digits", "sort", "accumulate", "reduce", "finalize", "total"]
        return dict(zip(keys, [float(x) for x in out]))

    def last_launches(self):
        return int(lib().msmb200_last_launches(self._h))

    def last_accumulator(self):
        """1 = XYZZ work items, 2 = batch-affine rounds (what the last MSM used)."""
        return int(lib().msmb200_last_accumulator(self._h))
