"""Developer probe: the drop-in reference driver (oracle/_ref/refdrv_p1_dropin_c<N>.so) with MSMB200_SHIM_TRACE=1.
usage: gpu_dropin_dev.py <config>"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle_lib as O
import msm_blst_b200 as M
cfg = sys.argv[1]
lib = C.CDLL(os.path.join(O.REF_DIR, "refdrv_p1_dropin_c%s.so" % cfg))
lib.refdrv_msm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
lib.refdrv_last_ms.restype = C.c_double
lib.refdrv_init()
n = 1 << int(cfg)
for seed in (1, 2, 3):
    sc = O.gen_scalars(seed, n)
    exp = O.serialize(1, O.closed_form(1, sc)[0])
    for method in (1, 2, 3, 4):
        ob = np.zeros(96, dtype=np.uint8)
        lib.refdrv_msm(method, O.ptr(sc), None, O.ptr(ob))
        print("seed %d method %d ok=%s driver %.2f ms, last shim call %.2f ms" % (seed, method, ob.tobytes() == exp, lib.refdrv_last_ms(), M.lib().msmb200_blst_last_call_ms(1)), flush=True)
