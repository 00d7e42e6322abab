#!/usr/bin/env python
"""bench.py — BLS12-381 fixed-base MSM latency (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload g1_n21|g2_n18|g1_n16] [--method 1..4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...       # the reference's CPU path on the host cores (oracle/_ref)

A step = one MSM (method body of main_p1.cpp incl. to_affine) over one fresh set of seeded synthetic scalars.
`value` is device-resident latency (scalars already in HBM), `e2e` the same call through the C ABI with HOST
scalars (pinned H2D copy of the step's scalars + D2H of the result inside the timed region). Rank 0 prints ONE
JSON line. Only the `cpu_baseline` leg and `--impl reference` execute anything under oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (group, reference config, description)
    "g1_n21": (1, "21", "G1 fixed-base MSM n=2^21, CHES construction (config 21: q=2^22, h=12, |B|=874437)"),
    "g2_n18": (2, "18", "G2 fixed-base MSM n=2^18, CHES construction (config 18: q=2^20, h=13, |B|=220931)"),
    "g1_n16": (1, "16", "G1 fixed-base MSM n=2^16, CHES construction (config 16: q=2^19, h=14, |B|=109244)"),
    "g1_n10": (1, "10", "G1 fixed-base MSM n=2^10, CHES construction (config 10)"),
}
METHOD_NAMES = {1: "CHES", 2: "CHES-integral", 3: "BGMW95", 4: "blst-Pippenger"}
R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


# ------------------------------------------------------------------------------------------------ inputs
def gen_scalars(seed, n):
    """Seeded splitmix64 scalar stream of SURVEY App. C (numpy restatement; vectorised, rejection of >= r)."""
    def splitmix(state_start, count):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            z = np.uint64(state_start) + idx * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))
    # a draw of 4 words is rejected as a whole, so accepted rows keep their order: over-draw rows, filter, take n
    r_limbs = [(R_ORDER >> (64 * i)) & (2**64 - 1) for i in range(4)]
    rows = int(n * 1.2) + 64
    while True:
        raw = splitmix(seed, 4 * rows).reshape(-1, 4).copy()
        raw[:, 3] >>= np.uint64(1)
        lt = np.zeros(len(raw), dtype=bool)
        eq = np.ones(len(raw), dtype=bool)
        for k in (3, 2, 1, 0):
            lt |= eq & (raw[:, k] < np.uint64(r_limbs[k]))
            eq &= raw[:, k] == np.uint64(r_limbs[k])
        good = raw[lt]
        if len(good) >= n:
            return np.ascontiguousarray(good[:n])
        rows *= 2


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.dev = device_index
        self.proc = None
        self.t_start = self.t_stop = None

    def launch(self):
        """Start the sampler process early (it needs ~0.5 s to print its first line)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=timestamp," + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def start(self):
        self.t_start = time.time()

    def stop(self):
        self.t_stop = time.time()
        time.sleep(0.05)
        lines = []
        if self.proc:
            self.proc.terminate()
            try:
                lines = self.proc.communicate(timeout=10)[0].splitlines()
            except Exception:
                lines = []
        import datetime
        sm, reasons, smax, power = [], set(), None, []
        for ln in lines:
            s = [x.strip() for x in ln.split(",")]
            try:
                ts = datetime.datetime.strptime(s[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if ts < self.t_start - 0.02 or ts > self.t_stop + 0.02:
                    continue
                sm.append(float(s[2]))
                smax = float(s[3])
                power.append(float(s[4]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[6:10]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------ cost model
def cost_model(group, cfg, n, method):
    """Algorithmic work per MSM from the REFERENCE's formulas."""
    c_add, c_dadd = (10, 14) if group == 1 else (28, 40)
    aff = 96 if group == 1 else 192
    if method in (1, 2):
        adds, dadds = n * cfg.h, 2 * (cfg.bsize - 1) + 2 * cfg.d
    elif method == 3:
        adds, dadds = n * cfg.h_bgmw, 2 * (1 << (cfg.e_bgmw - 1))
    else:
        wbits = n.bit_length() - 1
        w = wbits - 3 if wbits > 12 else (wbits - 2 if wbits > 4 else (2 if wbits else 1))
        tiles = 255 // w + 1
        adds, dadds = n * tiles, 2 * (1 << (w - 1)) * tiles
    w_fp_acc = adds * c_add
    w_fp = w_fp_acc + dadds * c_dadd
    # What the batch-affine accumulator EXECUTES: 5 multiplications + 1 squaring per addition; an Fp
    # multiplication is 300 wide multiply-accumulates, a squaring 234; Fp2: mul = 3 Fp mul, sqr = 2 Fp mul. One inversion
    # (about 12 multiplication equivalents on the multiplier pipe) per lane batch of up to 110 additions is left out.
    nonempty = cfg.bsize if method in (1, 2) else 0
    ba_adds = max(0, adds - nonempty)     # a bucket of k entries takes k - 1 additions
    ba_mac_per_add = (5 * 300 + 234) if group == 1 else (5 * 3 * 300 + 2 * 300)
    return {"adds": adds, "dadds": dadds, "w_fp": w_fp, "w_mac": 300 * w_fp, "w_mac_accumulate": 300 * w_fp_acc,
            "w_mac_accumulate_batch_affine": ba_adds * ba_mac_per_add,
            "gather_bytes": adds * aff, "w_bytes": adds * aff + 32 * n}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, secondary):
    """The reference's CPU path (compiled reference in oracle/_ref driven by the restated driver glue) on the
    host cores. Rank 0 only; other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    O.oracle()
    if not O.has_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (compiled reference) missing and /root/reference absent"}))
        return 0
    line = reference_workload(args, O, args.workload, args.method)
    if secondary:
        line["secondary"] = reference_workload(args, O, secondary, 1)
    print(json.dumps(line))
    return 0


def reference_workload(args, O, workload, method):
    group, cfgname, desc = WORKLOADS[workload]
    cfg = O.config(cfgname)
    n = 1 << cfg["n_exp"]
    threads = os.cpu_count() or 1
    steps, warmup = args.steps, args.warmup
    h = cfg["h"]
    # bounded sample: per-step estimate (1.2 us per bucket add / dadd on one core, G2 x3) -> shrink n_s until the run fits ~200 s
    unit = 1.2e-6 * (1 if group == 1 else 2.9)
    def est(ns):
        return ns * h * unit / threads + 2 * cfg["bsize"] * unit * 1.2
    n_s = n
    while n_s > 1024 and (steps + warmup) * est(n_s) > 200:
        n_s //= 2
    # the table: entries are canonical affine points, so any correct builder gives the reference's bytes. With a GPU
    # present it is built there and spot-checked against the reference chain; else on the CPU.
    oc = O.OracleCtx(group, cfgname, n=n_s, threads=threads)
    oc.init_fix_points() if n_s <= 4096 else None
    t0 = time.time()
    table_src = "cpu (restated single_scalar_multiplication chain, %d threads)" % threads
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    which = 1 if method == 3 else 0
    if method == 4 or n_s <= 4096 or not have_gpu:
        if n_s > 4096:
            oc.init_fix_points()
        if method != 4:
            oc.build_table(which)
    else:
        import msm_blst_b200 as M
        g = M.MsmContext(group, cfgname, npoints=n_s)
        g.init_fix_point_list()
        oc.set_points(g.download(0))
        if which == 0:
            g.init_pippenger_CHES_q_over_5()
        else:
            g.init_pippenger_BGMW95()
        oc.load_table(which, g.download(1 + which))
        # spot-check sampled rows against the reference's own construction
        rng = np.random.default_rng(0)
        hh = h if which == 0 else cfg["h_bgmw"]
        mult = 3 if which == 0 else 1
        for i in [0, n_s - 1] + [int(x) for x in rng.integers(0, n_s, size=6)]:
            chk = O.OracleCtx(group, cfgname, n=1, first=i)
            chk.init_fix_points()
            chk.build_table(which)
            row = oc.table(which).reshape(-1, O.AFF_BYTES[group])[mult * i * hh: mult * (i + 1) * hh].reshape(-1)
            assert (row == chk.table(which)).all(), "GPU-built table row %d differs from the reference chain" % i
        g.close()
        table_src = "gpu-built, 8 rows spot-checked against the reference chain"
    t_table = time.time() - t0
    sets = [O.gen_scalars(1 + s, n_s) for s in range(min(3, steps + warmup))]
    times, phases = [], None
    ok = True
    for it in range(warmup + steps):
        sc = sets[it % len(sets)]
        t0 = time.perf_counter()
        r, ph = oc.msm_ref(method, sc, threads=threads)
        dt = (time.perf_counter() - t0) * 1e3
        if it >= warmup:
            times.append(dt)
            phases = ph
        if it == 0:
            cf, _ = O.closed_form(group, sc)
            ok = bool((r == cf).all())
    ms = float(np.mean(times))
    scale = n / n_s
    if scale > 1:  # extrapolate the sample: glue and bucket accumulation scale with n, the bucket reduction does not
        ms_full = (phases["glue"] + (phases["tile"] - phases["reduce"])) * scale + phases["reduce"] + phases["finish"]
    else:
        ms_full = ms
    sample = "%s: n_sample=2^%d of 2^%d points (%s), %d host threads each running the reference tile on its slice; table %s" % (
        METHOD_NAMES[method], int(np.log2(n_s)), cfg["n_exp"], "full workload" if scale == 1 else "extrapolated x%g on the n-proportional phases" % scale,
        threads, table_src)
    line = {
        "impl": "reference", "metric": "BLS12-381 %s fixed-base MSM latency, n=2^%d, %s" % ("G1" if group == 1 else "G2", cfg["n_exp"], METHOD_NAMES[method]),
        "value": ms_full, "unit": "ms", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms_full,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64x6 (384-bit Montgomery, x86-64 ADX asm)",
        "data": "synthetic: P_i=2^(i+1)G, seeded splitmix64 scalars < r", "result_ok": ok,
        "config": {"workload": desc, "method": METHOD_NAMES[method], "n": n},
        "details": {"n_sample": n_s, "host_threads": threads,
                    "note": "NOT the stock single-threaded path: the compiled reference's tile functions, sharded over %d host threads the way the upstream "
                            "Rust / Go bindings do (as shipped the drivers use one core: see cpu_baseline of the b200 arm)" % threads},
        "cpu_baseline": {"value": ms_full, "unit": "ms", "cores": threads, "kind": "reference", "sample": sample,
                         "measured_ms_on_sample": ms, "phases_ms_on_sample": phases, "table_setup_s": t_table},
        "e2e": {"value": ms_full, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


def cpu_baseline_leg(group, cfgname, method, n, sets, gpu_ctx, budget_s=30.0):
    """Reference CPU path on ONE core (as shipped) next to the GPU run: the FULL workload when one MSM fits the
    budget (G1 n=2^21: ~11 s, G2 n=2^18: ~4.5 s), else a bounded sample with the n-proportional phases extrapolated."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import oracle_lib as O
        O.oracle()
        if not O.has_ref():
            return {"value": None, "unit": "ms", "cores": 1, "kind": "reference", "sample": "oracle/_ref missing"}
        cfg = O.config(cfgname)
        unit = 0.45e-6 * (1 if group == 1 else 2.9)   # measured: ~0.43 us per G1 bucket addition on one core of the GPU box
        n_s = n
        while n_s > 1024 and n_s * cfg["h"] * unit + 2 * cfg["bsize"] * unit * 1.4 > budget_s:
            n_s //= 2
        oc = O.OracleCtx(group, cfgname, n=n_s)
        which = 1 if method == 3 else 0
        oc.set_points(gpu_ctx.download(0, 0, n_s))
        if method != 4:
            per = 3 * cfg["h"] if which == 0 else cfg["h_bgmw"]
            oc.load_table(which, gpu_ctx.download(1 + which, 0, n_s * per))
        sc = np.ascontiguousarray(sets[0][:n_s])
        t0 = time.perf_counter()
        r, ph = oc.msm_ref(method, sc, threads=1)
        ms = (time.perf_counter() - t0) * 1e3
        cf, _ = O.closed_form(group, sc)
        scale = n / n_s
        ms_full = ms if scale == 1 else (ph["glue"] + ph["tile"] - ph["reduce"]) * scale + ph["reduce"] + ph["finish"]
        return {"value": ms_full, "unit": "ms", "cores": 1, "kind": "reference", "result_ok": bool((r == cf).all()),
                "sample": "%s, compiled reference (x86-64 ADX asm) tile functions on 1 core, n_sample=2^%d of 2^%d (%s), one run; table = GPU-built, "
                          "byte-identical to the reference's (tests)" % (METHOD_NAMES[method], int(np.log2(n_s)), cfg["n_exp"],
                                                                          "FULL workload, no extrapolation" if scale == 1 else "n-proportional phases x%g" % scale),
                "measured_ms_on_sample": ms, "phases_ms_on_sample": ph}
    except Exception as ex:  # the baseline must never break the bench line
        return {"value": None, "unit": "ms", "cores": 1, "kind": "reference", "sample": "failed: %r" % (ex,)}


# ------------------------------------------------------------------------------------------------ main arm
def ncu_constants(workload):
    """Per-launch DRAM traffic and heavy-FMA-pipe share of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r2_ncu_constants.json records the commit it was taken at); None when no capture exists for the workload."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_constants.json")))
        e = dict(d.get(workload) or {})
        if e:
            e["captured_at_commit"] = d.get("commit")
        return e or None
    except Exception:
        return None


def run_workload(args, workload, method, env, with_sampler, with_cpu_baseline, peaks):
    """One workload on this rank's GPU (sharded over the ranks when N > 1): returns the fields of a bench line."""
    torch, dist, M, D = env["torch"], env["dist"], env["M"], env["D"]
    N, rank, local_rank, stream = env["N"], env["rank"], env["local_rank"], env["stream"]
    group, cfgname, desc = WORKLOADS[workload]
    full_cfg = M.config_lookup(cfgname)
    n = 1 << full_cfg.n_exp
    by_buckets = N > 1 and args.shard == "buckets"
    lo, hi = (0, n) if (by_buckets or N == 1) else D.shard_range(n, rank, N)
    shard_cfgname = cfgname if (N == 1 or by_buckets) else (D.shard_config_name(hi - lo, group) if args.shard_config == "auto" else args.shard_config)
    ctx = M.MsmContext(group, shard_cfgname, npoints=hi - lo, device=local_rank, first=lo)
    if by_buckets:
        ctx.set_bucket_shard(rank, N)
    ctx.set_stream(stream.cuda_stream)
    t0 = time.time()
    ctx.init_fix_point_list()
    if method in (1, 2):
        ctx.init_pippenger_CHES_q_over_5()
    elif method == 3:
        ctx.init_pippenger_BGMW95()
    t_setup = time.time() - t0

    # several scalar sets, rotated so that no step re-reads the previous step's inputs
    nsets = 3
    sets = [gen_scalars(1 + s, n) for s in range(nsets)]
    # host side always holds this rank's 1/N of the scalars; with bucket sharding the device needs all of them
    # (device-resident for `value`; for `e2e` the shards are uploaded and all-gathered over NVLink inside the timed region)
    slo, shi = D.shard_range(n, rank, N)
    host_sets = [torch.from_numpy(s[slo:shi].view(np.uint8).copy()).pin_memory() for s in sets]
    if by_buckets:
        dev_sets = [torch.from_numpy(s.view(np.uint8).copy()).cuda() for s in sets]
        dev_shards = [torch.empty((shi - slo) * 32, dtype=torch.uint8, device="cuda") for _ in sets]
    else:
        dev_sets = [h.cuda(non_blocking=True) for h in host_sets]
    torch.cuda.synchronize()

    def step_device(i):
        if N == 1:
            return ctx.msm_device(method, dev_sets[i % nsets].data_ptr())
        return D.msm_sharded(ctx, method, dev_sets[i % nsets])

    def step_e2e(i):
        if N == 1:
            return ctx.msm(method, host_sets[i % nsets].numpy())
        d = dev_sets[i % nsets]
        if by_buckets:
            sh = dev_shards[i % nsets]
            sh.copy_(host_sets[i % nsets].view(-1), non_blocking=True)
            D.all_gather_scalars(sh, d.view(-1))
        else:
            d.copy_(host_sets[i % nsets], non_blocking=True)
        return D.msm_sharded(ctx, method, d)

    def timed(fn, steps, sampler=None):
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = None
        phase_acc = np.zeros(6)
        launches = 0
        for i in range(steps):
            res = fn(i)
            t = ctx.last_timings()
            phase_acc += np.array([t[k] for k in ("digits", "sort", "accumulate", "reduce", "finalize", "total")])
            launches += ctx.last_launches() + (3 if N > 1 else 0)   # N > 1: NCCL all-gather + the two combine kernels
        e1.record(stream)
        torch.cuda.synchronize()
        if N > 1:
            dist.barrier()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if N > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res, phase_acc / max(1, steps), launches, clocks

    for i in range(args.warmup):
        step_device(i)
        step_e2e(i)
    sampler = ClockSampler(local_rank) if (rank == 0 and with_sampler) else None
    if sampler:
        sampler.launch()
    # every rank runs the same steps (collectives inside): keep the GPUs busy while the sampler process starts
    for i in range(args.warmup):
        step_device(i)
    if with_sampler:
        time.sleep(0.6)
    step_device(0)
    ms_total, res, phases, launches, clocks = timed(step_device, args.steps, sampler)
    ms_e2e_total, res_e2e, _, _, _ = timed(step_e2e, args.steps)
    ms_step = ms_total / args.steps
    ms_e2e = ms_e2e_total / args.steps
    phases_all = None
    if N > 1:   # per-rank phase times of the device-resident steps, so that the scaling record names the limiter
        phases_all = [None] * N
        dist.all_gather_object(phases_all, [float(x) for x in phases])

    out = None
    if rank == 0:
        last_set = (args.steps - 1) % nsets
        cm = cost_model(group, full_cfg, n, method)
        names = ("digits", "sort", "accumulate", "reduce", "finalize", "device_total")
        out = {
            "metric": "BLS12-381 %s fixed-base MSM latency, n=2^%d, %s" % ("G1" if group == 1 else "G2", full_cfg.n_exp, METHOD_NAMES[method]),
            "value": ms_step, "unit": "ms", "ms_per_step": ms_step,
            "config": {"workload": desc, "method": METHOD_NAMES[method], "n": n},
            "details": {"shard_config": shard_cfgname,
                        "parallelism": ("bucket ranges x%d, tables replicated" % N) if by_buckets else ("points sharded x%d" % N),
                        "l2_policy": "inputs larger than L2: %.2f GB precomputation table gathered at random per step, %d rotating scalar sets" % (
                            (3 * n * full_cfg.h if method in (1, 2) else n * full_cfg.h_bgmw if method == 3 else n) * (96 if group == 1 else 192) / 1e9, nsets),
                        "table_setup_s": t_setup},
            "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": int(n * 32), "d2h_bytes_per_step": (96 if group == 1 else 192)},
            "gpu_launches": launches,
            "result_hex": M.affine_serialize(group, res).hex(), "result_consistent": bool((res == res_e2e).all()), "last_scalar_seed": 1 + last_set,
            "clocks": clocks,
            "fp_mul_per_s": cm["w_fp"] / (ms_step * 1e-3),
        }
        if N == 1:
            out["phases_ms"] = dict(zip(names, [float(x) for x in phases]))
        else:
            out["phases_ms_per_rank"] = [dict(zip(names, p)) for p in phases_all]
        if N == 1:
            # roofline of the dominant phase: bucket accumulation. Algorithmic MACs (reference formulas) over CUDA-event time,
            # against the IMAD.WIDE rate measured live (MACs per clock per SM x SMs x the SM clock sampled during the steps).
            acc_ms = float(phases[2])
            sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
            peak = peaks["macs_per_clk_per_sm"] * env["sms"] * sm_mhz * 1e6 / 1e12
            batch_affine = ctx.last_accumulator() == 2
            ref_formula = cm["w_mac_accumulate"] / (acc_ms * 1e-3) / 1e12
            achieved = (cm["w_mac_accumulate_batch_affine"] if batch_affine else cm["w_mac_accumulate"]) / (acc_ms * 1e-3) / 1e12
            nc = ncu_constants(workload) if method == 1 else None
            accum = ("batch-affine rounds (ba_round_kernel x ceil(log2 max bucket) launches; 5M + 1S per addition)" if batch_affine
                     else "XYZZ work items (accumulate_kernel; 8M + 2S per addition)")
            out["roofline"] = {
                "bound": "imad", "kernel": "bucket accumulation: " + accum,
                "achieved": achieved, "peak": peak, "unit": "TMAC/s (32x32+64-bit)", "frac": achieved / peak,
                "achieved_reference_formula": ref_formula, "frac_reference_formula": ref_formula / peak,
                "traffic": (nc or {}).get("dram_bytes_accumulate_phase"),
                "fmaheavy_pipe_active_pct_ncu": (nc or {}).get("fmaheavy_pct_accumulate_phase"),
                "ncu_capture": nc,
                "peak_source": "msmb200_measure_peaks_ex in this run: %.2f MAC/clk/SM (IMAD.WIDE.U32, both multiplicands changing every iteration; %.0f MHz during the "
                               "microbenchmark) x %d SMs x %.0f MHz sampled during the timed steps" % (peaks["macs_per_clk_per_sm"], peaks["sm_mhz_during_microbench"], env["sms"], sm_mhz),
                "note": "achieved = multiply-accumulates the accumulator EXECUTES by its own operation count (DESIGN.md §5: batch-affine (n*h - |B|) x "
                        "(5 x 300 + 234) for G1, x 5100 for G2; XYZZ n*h x C_add x 300) / CUDA-event time of the accumulate phase (all its launches). "
                        "achieved_reference_formula = the reference algorithm's count (n*h x C_add x 300, SURVEY §8d) over the same time: it exceeds the "
                        "pipe peak when batch-affine runs, because 5M + 1S replace the reference's 8M + 2S per addition.",
            }
            out["roofline_path"] = {"bound": "imad", "achieved": cm["w_mac"] / (ms_step * 1e-3) / 1e12, "peak": peak,
                                    "unit": "TMAC/s", "frac": cm["w_mac"] / (ms_step * 1e-3) / 1e12 / peak, "note": "whole MSM, W_MAC = 300*W_Fp"}
            try:
                hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                hbm_src = "measured (MEASURED_PEAKS.json)"
            except Exception:
                hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
            gbs = cm["gather_bytes"] / (acc_ms * 1e-3) / 1e9
            out["roofline_hbm"] = {"bound": "hbm", "kernel": "bucket accumulation (table gather)", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": gbs / hbm_peak, "traffic": (nc or {}).get("dram_bytes_accumulate_phase"),
                                   "algorithmic_bytes": cm["gather_bytes"], "peak_source": hbm_src}
            if batch_affine:
                # what the pairwise rounds move by construction, per addition: forward pass x1, x2 in + prefix product out; backward pass
                # prefix, x1, x2, y1, y2 in + the affine result out; two 16-byte descriptor reads
                fe = 48 if group == 1 else 96
                ba_bytes = (cm["adds"] - (full_cfg.bsize if method in (1, 2) else 0)) * (2 * fe + fe + fe + 4 * fe + 2 * fe + 32)
                out["roofline_hbm"]["algorithmic_bytes_batch_affine"] = ba_bytes
                out["roofline_hbm"]["achieved_batch_affine"] = ba_bytes / (acc_ms * 1e-3) / 1e9
                out["roofline_hbm"]["note"] = ("algorithmic_bytes = the reference algorithm's table gather (n*h entries, read once). The batch-affine rounds trade "
                                               "multiplications for memory traffic: every round streams its operands twice (forward / backward pass) and writes "
                                               "its results for the next round - algorithmic_bytes_batch_affine; the measured traffic exceeds it by the 128-byte "
                                               "line granularity of the round-0 table reads (48 / 96 useful bytes per line).")
            if with_cpu_baseline:
                out["cpu_baseline"] = cpu_baseline_leg(group, cfgname, method, n, sets, ctx)
    ctx.close()
    del dev_sets, host_sets
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: g1_n21 as the headline plus g2_n18 as `secondary` (the two halves of the BASELINE metric)")
    ap.add_argument("--method", type=int, default=1, choices=[1, 2, 3, 4])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--shard-config", default="auto", help="reference config used per shard when --gpus > 1 (auto: tuned for n/G)")
    ap.add_argument("--shard", default="points", choices=["buckets", "points"],
                    help="multi-GPU decomposition: point shards (default, faster at N<=8) or bucket ranges with replicated tables")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    secondary = None
    if args.workload is None:
        args.workload = "g1_n21"
        secondary = None if args.no_secondary else "g2_n18"
    if args.impl == "reference":
        return run_reference(args, secondary)

    import torch
    import torch.distributed as dist

    import msm_blst_b200 as M
    from msm_blst_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: msm_blst_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    env = {"torch": torch, "dist": dist, "M": M, "D": D, "N": world, "rank": rank, "local_rank": local_rank,
           "stream": torch.cuda.current_stream(), "sms": torch.cuda.get_device_properties(local_rank).multi_processor_count}
    peaks = M.api.measure_peaks_ex(local_rank) if rank == 0 else None

    first = run_workload(args, args.workload, args.method, env, True, not args.no_cpu_baseline, peaks)
    second = run_workload(args, secondary, 1, env, False, not args.no_cpu_baseline, peaks) if secondary else None
    if rank == 0:
        line = {
            "metric": first["metric"], "value": first["value"], "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": first["ms_per_step"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32x12 (384-bit Montgomery integers)", "data": "synthetic: P_i=2^(i+1)G, seeded splitmix64 scalars < r (SURVEY App. C)",
        }
        for k, v in first.items():
            if k not in line:
                line[k] = v
        line["peaks_measured_live"] = peaks
        if second:
            line["secondary"] = second   # the other half of the BASELINE metric ("G2 ms at 2^18"), same steps / warm-up / timing rules
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
