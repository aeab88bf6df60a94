#!/usr/bin/env python3
"""bench.py -- headline benchmark of the vcfgl simulate-and-score hot path on B200.

Metric (BASELINE.json): simulate+GL throughput in site x sample cells/s.
Workload (configs[1]): synthetic msprime-shaped genotypes, 100 diploid samples x 1M biallelic
sites, Poisson depth 10, error 0.01, GL model 1, tags GL+PL+AD(+DP).  One STEP = one pass of the
hot path over one batch of 131072 sites (13.1 M cells, ~1.9 GB of tag planes -- larger than the
126 MB L2, so no L2 flush is needed between steps); 8 steps = 1,048,576 sites.

  value  kernel-side throughput: genotypes resident in HBM, results left in HBM, CUDA events
         on the launching stream (torch's current stream, handed to libvgl).
  e2e    the same metric through the C ABI with HOST buffers: pinned H2D of the packed
         genotypes and D2H of every tag plane inside the timed region, two slots in flight.
         Planes cross PCIe in the ABI's VGL_HOST_NARROW form (GL float32; PL / AD / DP as the
         8-bit values BCF stores, narrowed on the device); `e2e.i32_planes` is the same loop
         with the int32 planes of VGL_HOST_I32 (add_tags() layout) for comparison.
  roofline      algorithmic bytes / device time of the kernels, against the
                measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference binary itself (oracle/_ref/vcfgl_ref, built from /root/reference)
                on a bounded sample of the same workload on this box's host, 1 thread.

  input_path    (N=1) SURVEY.md 8(f) row 1: one step's worth of msprime-shaped VCF text through k_vcf_* -- kernel-side
                throughput with its own roofline, end to end from pinned host text (to narrowed arrays and to finished BCF
                records), and the oracle's single-threaded C parser as CPU baseline.
  gvcf_merge    (workloads with -doGVCF) SURVEY.md 8(f) row 3: the block merger on the batch resident in HBM.

`--impl reference` times the reference CPU binary with one process per host core on contiguous
site shards (the reference cannot thread its simulation).

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 via torch.distributed.run.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sim+GL site x sample cells/s"
UNIT = "cells/s"
N_SAMPLES = 100
BATCH_SITES = 131072
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "vcfgl_ref")
VCFGL_ARGS = "-d 10 -e 0.01 -GL 1 -addGL 1 -addPL 1 -addFormatAD 1".split()
WORKLOAD = ("cfg2: 100 samples x 1M sites (steps of %d sites), Poisson depth 10, e=0.01, GL model 1, "
            "tags GL+PL+AD+DP, seed 42" % BATCH_SITES)


QS_BINS = None          # --qs-bins ranges of the workload (start, end, value)
INVARIANT_SHARE = 0.0   # share of all-hom-ref sites in the synthetic genotypes (gVCF / -explode runs)


def set_workload(name):
    """cfg2 is the bench line (BASELINE.json configs[1]).  The others are the remaining BASELINE.json configs at a
    per-step size that fits one GPU; they are measured with `--workload` for DESIGN.md and never replace the cfg2 line."""
    global N_SAMPLES, BATCH_SITES, VCFGL_ARGS, WORKLOAD, QS_BINS, INVARIANT_SHARE
    if name == "cfg5":
        N_SAMPLES, BATCH_SITES = 10000, 4440   # 5 x (148 SMs x 6 resident CTAs) tiles
        VCFGL_ARGS = "-d 30 -e 0.01 -GL 1 -addGL 1 -addPL 1 -addFormatAD 1".split()
        WORKLOAD = ("cfg5 shape: 10000 samples, steps of %d sites, Poisson depth 30, e=0.01, GL model 1, "
                    "tags GL+PL+AD+DP, seed 42" % BATCH_SITES)
    elif name == "cfg3a":
        N_SAMPLES, BATCH_SITES = 1000, 8192
        VCFGL_ARGS = "-d 10 -e 0.01 -GL 2 -eq 1 -bv 1e-5 -addGL 1 -addPL 1".split()
        WORKLOAD = ("cfg3(i): 1000 samples, steps of %d sites, Poisson depth 10, GL model 2, per-site beta error "
                    "(mean 0.01, var 1e-5), tags GL+PL+DP, seed 42" % BATCH_SITES)
    elif name == "cfg3b":
        N_SAMPLES, BATCH_SITES = 1000, 8192
        VCFGL_ARGS = "-d 10 -e 0.01 -GL 2 -eq 2 -bv 1e-5 -addGL 1 -addPL 1".split()
        QS_BINS = [(0, 2, 2), (3, 14, 12), (15, 30, 23), (31, 63, 37)]   # RTA3 bins, last range opened to 63
        WORKLOAD = ("cfg3(ii): 1000 samples, steps of %d sites, Poisson depth 10, GL model 2, per-read beta error "
                    "(mean 0.01, var 1e-5) + RTA3 qs bins, tags GL+PL+DP, seed 42" % BATCH_SITES)
    elif name == "cfg4":
        N_SAMPLES, BATCH_SITES = 100, 131072
        VCFGL_ARGS = "-d 10 -e 0.001 -GL 1 -doUnobserved 1 -doGVCF 1 --gvcf-dps 1,5,10 -addGL 1 -addPL 1 -addI16 1 -addQS 1".split()
        INVARIANT_SHARE = 0.99
        WORKLOAD = ("cfg4: 100 samples, steps of %d sites of which 99%% invariant (-explode), Poisson depth 10, e=0.001, "
                    "GL model 1, <*> allele, tags GL+PL+DP+I16+QS, seed 42" % BATCH_SITES)


def measured_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def sim_args():
    from vcfgl_b200 import args as vargs
    return vargs.parse_args(["--seed", "42"] + VCFGL_ARGS, qs_bins=QS_BINS)


# --------------------------------------------------------------------------- reference CPU arm
def run_reference_shards(n_proc, sites_per_proc, seed, tmp):
    """P processes of the unmodified reference on P contiguous site shards; returns (cells, wall_s)"""
    from vcfgl_b200 import synth
    paths = []
    for r in range(n_proc):
        hap = synth.sfs_genotypes(sites_per_proc, N_SAMPLES, seed + r)
        pos = synth.positions(sites_per_proc, sites_per_proc * 10, seed + r)
        path = os.path.join(tmp, "shard%d.vcf" % r)
        synth.write_vcf(path, hap, pos, sites_per_proc * 10)
        paths.append(path)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([REF_BIN, "-i", path, "-o", path + ".out", "-O", "u", "--seed", "42"] + VCFGL_ARGS,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for path in paths]
    rcs = [p.wait() for p in procs]
    wall = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError("reference binary failed: %s" % rcs)
    for path in paths:
        for ext in (".out.bcf", ".out.arg"):
            try:
                os.remove(path + ext)
            except OSError:
                pass
    return n_proc * sites_per_proc * N_SAMPLES, wall


def reference_arm(opt):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/vcfgl_ref was not built (needs /root/reference at build time)"}))
        return
    cores = os.cpu_count() or 1
    sites_per_proc = 8192
    tmp = tempfile.mkdtemp(prefix="vgl_refbench_")
    try:
        for w in range(opt.warmup):
            run_reference_shards(cores, 1024, 1000 + w, tmp)
        cells = 0
        wall = 0.0
        for k in range(opt.steps):
            c, t = run_reference_shards(cores, sites_per_proc, 2000 + 97 * k, tmp)
            cells += c
            wall += t
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    value = cells / wall
    sample = "%d steps x %d processes x %d sites x %d samples, -O u, one process per host core" % (
        opt.steps, cores, sites_per_proc, N_SAMPLES)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": opt.gpus, "steps": opt.steps,
        "warmup": opt.warmup, "ms_per_step": 1e3 * wall / opt.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU binary (vcfgl v1.3.0 da6a334), whole program: VCF parse + simulate + BCF -O u"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# --------------------------------------------------------------------------- GPU arm
def gpu_arm(opt):
    import numpy as np
    import torch
    from vcfgl_b200 import capi, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libvgl has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    a = sim_args()
    B, S, K, W = BATCH_SITES, N_SAMPLES, opt.steps, opt.warmup
    cells_per_step = B * S
    # contiguous site range of this rank (weak scaling: every rank simulates K + W batches of its own)
    site0 = rank * (K + W + 4) * B
    hap = synth.sfs_genotypes(B, S, 20260002 + rank)
    if INVARIANT_SHARE > 0:      # positions absent from the input VCF: hom-ref records synthesised by -explode
        inv = np.random.default_rng(7 + rank).random(B) < INVARIANT_SHARE
        hap[inv] = 0
    gt = synth.pack_gt(hap)      # binary source: REF=0 -> A, ALT=1 -> C (vcfgl.cpp:103-128)
    stream = torch.cuda.Stream()   # an explicit stream: libvgl treats a NULL stream as "use the slot's own"

    # ---------------- value: genotypes resident in HBM, results stay in HBM
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=B, n_slots=2, device_id=local, host_output=False))
    kernels = ctx.native_kernels()
    for s in (0, 1):
        ctx.set_stream(s, stream.cuda_stream)
        ctx.input_buffer(s)[:] = gt
        ctx.submit(s, site0, B)          # uploads the genotypes once (untimed)
        ctx.wait(s)
    step = [0]

    def run_steps(n, flags):
        # two slots in flight on one stream keeps the GPU queue non-empty
        pend = []
        for _ in range(n):
            s = step[0] & 1
            if len(pend) == 2:
                ctx.wait(pend.pop(0))
            ctx.submit(s, site0 + step[0] * B, B, flags=flags)
            pend.append(s)
            step[0] += 1
        for s in pend:
            ctx.wait(s)

    run_steps(W, capi.SUBMIT_GT_ON_DEVICE)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = np.zeros(capi.T_COUNT)
    ev0.record(stream)
    pend = []
    for _ in range(K):
        s = step[0] & 1
        if len(pend) == 2:
            d = pend.pop(0)
            ctx.wait(d)
            kern_ms += ctx.timing(d)
        ctx.submit(s, site0 + step[0] * B, B, flags=capi.SUBMIT_GT_ON_DEVICE)
        pend.append(s)
        step[0] += 1
    for d in pend:
        last = ctx.wait(d)
        kern_ms += ctx.timing(d)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    ctx.copy_sites(d, last)
    alg_bytes = ctx.algorithmic_bytes(last)           # of one step (last batch)
    g_share = float((last.sites["n_genotypes"] == 15).mean())
    gvcf_leg = None
    if a.do_gvcf and a.do_unobserved in (1, 2) and rank == 0:
        # the gVCF block merger (bcf_utils.cpp:662-942) over the batch still resident in HBM
        dps = [int(x) for x in a.gvcf_dps.split(",")]
        rid0, pos0 = np.zeros(B, np.int32), np.arange(B, dtype=np.int32)
        gm = [ctx.gvcf_merge(d, rid0, pos0, dps) for _ in range(4)]
        g_ms = min(x["ms_kernels"] for x in gm[1:])
        members = int(gm[-1]["recs"]["n_members"].sum())
        g_alg = 4 * B * S + 16 * members * S + 16 * gm[-1]["n_blocks"] * S      # DP plane in; DP + 3 PL per member cell in; per block out
        pk, pk_src = measured_peak()
        gvcf_leg = {"what": "k_gvcf_key / plan_local / plan_global / fin / reduce on one step (sites at consecutive positions of one contig)",
                    "kernel_ms": g_ms, "gpu_launches": 5, "records": int(len(gm[-1]["recs"])), "blocks": int(gm[-1]["n_blocks"]),
                    "member_sites": members, "sites_per_s": B / (g_ms * 1e-3),
                    "roofline": {"bound": "hbm", "achieved": g_alg / (g_ms * 1e-3) / 1e9, "peak": pk, "unit": "GB/s",
                                 "frac": g_alg / (g_ms * 1e-3) / 1e9 / pk, "peak_source": pk_src, "traffic": None,
                                 "algorithmic_bytes_per_step": g_alg}}
    ctx.close()
    t_max = ms
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_max = float(t.item())
    value = world * K * cells_per_step / (t_max * 1e-3)

    # ---------------- e2e: host buffers, H2D + kernels + D2H inside the timed region
    if opt.skip_e2e:     # profiling runs only (ncu); such a line is not a bench result
        if rank == 0:
            print(json.dumps({"profiling_only": True, "value": value, "kernel_ms": (kern_ms / K).tolist()}))
        return
    Ke = max(2, min(K, 4))
    s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()

    def e2e_steps(ctx, bufs, n, first, bcf_in=None):
        pend = []
        d2h = 0
        for i in range(n):
            s = i & 1
            if len(pend) == 2:
                b = ctx.wait(pend.pop(0))
                d2h = out_bytes(b)
            bufs[s][:] = gt                       # the caller packs this step's genotypes into pinned memory
            if bcf_in is not None:                # ... and the fields its input records pass through (CHROM, POS; ID ".", FILTER ".")
                sin = bcf_in[s]
                sin["pos"][:B] = np.arange(first + i * B, first + (i + 1) * B, dtype=np.int64) % (1 << 30)
                sin["qual_bits"][:B] = capi.F32_MISSING_BITS
            ctx.submit(s, first + i * B, B)
            pend.append(s)
        for s in pend:
            b = ctx.wait(s)
            d2h = out_bytes(b)
        return d2h

    def out_bytes(b):
        r = b.raw
        if b.bcf_off is not None:   # VGL_HOST_BCF: the record stream, its offsets, the per-site records
            return int(b.bcf_bytes + 8 * (b.n_sites + 1) + b.n_sites * capi.SITE_DTYPE.itemsize)
        w = (b.narrow_bits // 8) if b.narrow_bits else 4          # DP / AD element width
        n = b.n_sites * S * w + b.n_sites * capi.SITE_DTYPE.itemsize
        n += 4 * b.g_elems * sum(bool(x) for x in (r.gl, r.gp))
        n += b.g_elems * (4 * bool(r.pl) + bool(r.pl_u8))
        n += w * b.r_elems * sum(bool(x) for x in ((r.ad_n, r.adf_n, r.adr_n) if b.narrow_bits else (r.ad, r.adf, r.adr)))
        return int(n)

    def e2e_run(mode):
        ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=B, n_slots=2, device_id=local, host_output=mode,
                                                 bcf_dict=dict(DP=1, GL=2, PL=3, GP=4, AD=5, ADF=6, ADR=7, QS=8, I16=9)))
        ctx.set_stream(0, s_a.cuda_stream)
        ctx.set_stream(1, s_b.cuda_stream)
        bufs = [ctx.input_buffer(0), ctx.input_buffer(1)]
        bcf_in = [ctx.bcf_input(0)[0], ctx.bcf_input(1)[0]] if mode == capi.HOST_BCF else None
        e2e_steps(ctx, bufs, 2, site0 + (K + W) * B, bcf_in)
        barrier()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_a)
        d2h = e2e_steps(ctx, bufs, Ke, site0 + (K + W) * B, bcf_in)
        s_a.wait_stream(s_b)
        e1.record(s_a)
        barrier()
        ms = e0.elapsed_time(e1)
        n_launch = ctx.launch_count() - l0
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        ctx.close()
        return world * Ke * cells_per_step / (ms * 1e-3), d2h, n_launch

    e2e_i32, d2h_i32, _ = e2e_run(capi.HOST_I32)
    e2e_value, d2h_bytes, e2e_launches = e2e_run(capi.HOST_NARROW)
    e2e_bcf = None
    if not a.do_gvcf:   # serialised BCF records (the gVCF block merger consumes arrays)
        v, nb, nl = e2e_run(capi.HOST_BCF)
        e2e_bcf = {"value": v, "d2h_bytes_per_step": nb, "gpu_launches": int(nl),
                   "planes": "VGL_HOST_BCF: complete BCF records serialised on the device (k_bcf_plan/scan/emit), byte-identical to "
                             "the reference's -O u stream; the host only appends the buffer to the output"}

    # ---------------- input path: VCF text -> packed genotypes on the device -> the same kernels
    input_path = None
    if world == 1 and not opt.no_input_path:
        input_path = input_path_leg(a, gt, hap, B, S, site0 + (K + W + 2) * B, local, peak_of=measured_peak)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the kernels (device time inside the timed region)
    peak, peak_src = measured_peak()
    kern_ms /= K
    fused = "+" not in kernels       # one kernel does the whole path; libvgl reports its time in the T_EMIT interval
    dev_ms = float(kern_ms[capi.T_EMIT]) if fused else \
        float(kern_ms[capi.T_SIM] + kern_ms[capi.T_SITE] + kern_ms[capi.T_SCAN] + kern_ms[capi.T_EMIT])
    achieved = alg_bytes / (dev_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:    # DRAM bytes of one launch from the last `ncu --set full` capture of this workload (profiles/README.md)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[opt.workload]
        traffic, traffic_src = tr["dram_bytes_per_launch"], tr["source"]
    except Exception:
        pass
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "kernel": "%s (%s per step; algorithmic bytes of the whole path / its average launch duration, CUDA events "
                      "on the launching stream)" % (kernels, "one launch" if fused else "four launches"),
            "algorithmic_bytes_per_step": alg_bytes, "bytes_per_cell": alg_bytes / cells_per_step,
            "kernel_ms": {kernels: dev_ms} if fused else
                         {"k_sim": float(kern_ms[capi.T_SIM]), "k_site": float(kern_ms[capi.T_SITE]),
                          "k_scan": float(kern_ms[capi.T_SCAN]), "k_emit": float(kern_ms[capi.T_EMIT])}}

    # ---------------- CPU baseline: the reference binary, 1 thread, bounded sample
    cpu = None
    if world == 1 and not opt.no_cpu_baseline:
        if os.path.exists(REF_BIN):
            tmp = tempfile.mkdtemp(prefix="vgl_cpubase_")
            try:
                n_sites = 196608
                cells, wall = run_reference_shards(1, n_sites, 4242, tmp)
                cpu = {"value": cells / wall, "unit": UNIT, "cores": 1, "kind": "reference",
                       "sample": "%d sites x %d samples of the same workload, reference binary -O u, wall %.1f s" % (n_sites, S, wall)}
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref/vcfgl_ref missing"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": t_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_sites": B, "cells_per_step": cells_per_step,
                   "l2": "no flush: each step writes %.2f GB of tag planes (> 126 MB L2)" % (alg_bytes / 1e9),
                   "n_samples": S,
                   "sites_with_15_genotypes": g_share, "sharding": "contiguous site ranges per GPU, no collective"},
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * S), "d2h_bytes_per_step": d2h_bytes,
                "steps": Ke, "gpu_launches": int(e2e_launches),
                "planes": "VGL_HOST_NARROW: GL float32, PL/AD/DP narrowed on the device to the 8-bit values BCF stores",
                "i32_planes": {"value": e2e_i32, "d2h_bytes_per_step": d2h_i32,
                               "planes": "VGL_HOST_I32: every plane int32/float32 as add_tags() hands them to htslib"},
                "bcf_records": e2e_bcf},
        "input_path": input_path, "gvcf_merge": gvcf_leg,
        "gpu_launches": int(launches), "clocks": clk}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def input_path_leg(a, gt, hap, B, S, first_site, local, peak_of):
    """k_vcf_lines + k_vcf_gt on one step's worth of msprime-shaped VCF text (B records x S samples):
    `value`  text resident in HBM, kernels only (CUDA events on the parser's stream);
    `e2e`    host text in pinned memory -> H2D -> parse -> place -> simulate -> D2H of the narrowed planes;
    `cpu_baseline`  the oracle's single-threaded C restatement of the same parse on a bounded sample."""
    import numpy as np
    import torch
    from vcfgl_b200 import capi, synth
    pos = np.arange(1, B + 1, dtype=np.int64) * 7
    body = synth.vcf_body(hap, pos)
    n_text = len(body)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=B, n_slots=2, device_id=local, host_output=capi.HOST_NARROW))
    ps = ctx.parser(n_text + 64, B)
    res = ps.parse(body, capi.SOURCE_BINARY, 0)
    assert res.n_records == B and res.n_errors == 0 and res.bytes_consumed == n_text, (res.n_records, res.n_errors)
    assert np.array_equal(ps.rows(0, 64), gt[:64]) and np.array_equal(ps.rows(B - 64, 64), gt[B - 64:])
    reps, ms = 6, []
    for _ in range(2):
        ps.parse(None, capi.SOURCE_BINARY, capi.PARSE_TEXT_ON_DEVICE, n_bytes=n_text)
    for _ in range(reps):
        ms.append(ps.parse(None, capi.SOURCE_BINARY, capi.PARSE_TEXT_ON_DEVICE, n_bytes=n_text).ms_kernels)
    k_ms = float(np.mean(ms))
    alg = n_text + B * S + B * capi.IN_SITE_DTYPE.itemsize
    peak, peak_src = peak_of()
    # end to end from text
    def step(i, slot):
        r = ps.parse(None, capi.SOURCE_BINARY, 0, n_bytes=n_text)          # H2D of the pinned text + kernels
        ctx.place_rows(slot, ps, B)
        ctx.submit(slot, first_site + i * B, B, flags=capi.SUBMIT_GT_ON_DEVICE)
        return r
    for i in range(2):
        step(i, i & 1)
    ctx.wait(0), ctx.wait(1)
    torch.cuda.synchronize()
    n_e2e = 4
    import time as _t
    t0 = _t.perf_counter()
    pend = []
    for i in range(n_e2e):
        slot = i & 1
        if len(pend) == 2:
            ctx.wait(pend.pop(0))
        step(i, slot)
        pend.append(slot)
    for slot in pend:
        ctx.wait(slot)
    torch.cuda.synchronize()
    wall = _t.perf_counter() - t0
    ps.close()
    ctx.close()
    # VCF text in -> finished BCF records out: input path + simulation + device BCF serialisation (VGL_HOST_BCF); the host only
    # copies POS from the parser's site records into the pass-through records and would append the returned bytes to the file
    text_to_bcf = None
    if not a.do_gvcf:
        ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=B, n_slots=2, device_id=local, host_output=capi.HOST_BCF,
                                                 bcf_dict=dict(DP=1, GL=2, PL=3, GP=4, AD=5, ADF=6, ADR=7, QS=8, I16=9)))
        ps = ctx.parser(n_text + 64, B)
        ps.parse(body, capi.SOURCE_BINARY, 0)
        bcf_in = [ctx.bcf_input(0)[0], ctx.bcf_input(1)[0]]

        def step_bcf(i, slot):
            r = ps.parse(None, capi.SOURCE_BINARY, 0, n_bytes=n_text)
            sin = bcf_in[slot]
            sin["pos"][:B] = r.sites["pos"]
            sin["qual_bits"][:B] = capi.F32_MISSING_BITS
            ctx.place_rows(slot, ps, B)
            ctx.submit(slot, first_site + i * B, B, flags=capi.SUBMIT_GT_ON_DEVICE)
        for i in range(2):
            step_bcf(i, i & 1)
        ctx.wait(0), ctx.wait(1)
        torch.cuda.synchronize()
        t0 = _t.perf_counter()
        pend, nbytes = [], 0
        for i in range(n_e2e):
            slot = i & 1
            if len(pend) == 2:
                nbytes = ctx.wait(pend.pop(0)).bcf_bytes
            step_bcf(i, slot)
            pend.append(slot)
        for slot in pend:
            nbytes = ctx.wait(slot).bcf_bytes
        torch.cuda.synchronize()
        wall_bcf = _t.perf_counter() - t0
        ps.close()
        ctx.close()
        text_to_bcf = {"value": n_e2e * B * S / wall_bcf, "unit": UNIT, "h2d_bytes_per_step": n_text, "d2h_bytes_per_step": int(nbytes), "steps": n_e2e,
                       "note": "VCF text in pinned host memory -> BCF records in pinned host memory: k_vcf_* + k_place_rows + simulation + k_bcf_*; host wall clock"}
    # CPU baseline: oracle restatement, one thread, bounded sample
    cpu = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import vcfin_oracle as vo
        n_cpu = min(B, 32768)
        cut = 0
        for _ in range(n_cpu):
            cut = body.index(b"\n", cut) + 1
        t0 = _t.perf_counter()
        sites, rows, used = vo.parse(body[:cut], S, 0)
        dt = _t.perf_counter() - t0
        assert len(sites) == n_cpu and np.array_equal(rows, gt[:n_cpu])
        cpu = {"value": n_cpu * S / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "oracle/vcf_in_oracle.c on the first %d records (%.1f MB of text), %.3f s" % (n_cpu, cut / 1e6, dt)}
    except Exception as e:      # the oracle is test infrastructure; its absence must not break the bench line
        cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "unavailable: %s" % e}
    traffic, traffic_src = None, None
    try:    # DRAM bytes of the path's kernels from their ncu --set full captures (profiles/traffic.json), when this shape was profiled
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["input_path_%dx%d" % (B, S)]
        traffic, traffic_src = tr["dram_bytes_per_step"], tr["source"]
    except Exception:
        pass
    return {"what": "VCF text -> packed genotypes on the device (k_vcf_count, k_vcf_index, k_vcf_hdr, k_vcf_cells, k_vcf_gt), %d records x %d samples, %.1f MB of text per step" % (B, S, n_text / 1e6),
            "value": B * S / (k_ms * 1e-3), "unit": UNIT, "kernel_ms": k_ms, "gpu_launches_per_step": 5,
            "roofline": {"bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (k_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_step": alg, "bytes_per_cell": alg / (B * S)},
            "e2e": {"value": n_e2e * B * S / wall, "unit": UNIT, "h2d_bytes_per_step": n_text, "steps": n_e2e,
                    "note": "pinned host text -> H2D -> k_vcf_* -> k_place_rows -> simulate -> D2H (VGL_HOST_NARROW); host wall clock, two slots"},
            "text_to_bcf": text_to_bcf, "cpu_baseline": cpu}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-input-path", action="store_true", help="skip the VCF-text input-path leg")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling only: stop after the kernel-side loop")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3a", "cfg3b", "cfg4", "cfg5"])
    opt = ap.parse_args()
    set_workload(opt.workload)
    opt.warmup = max(opt.warmup, 3) if opt.impl == "b200" else opt.warmup
    if opt.impl == "reference":
        reference_arm(opt)
    else:
        gpu_arm(opt)


if __name__ == "__main__":
    main()
