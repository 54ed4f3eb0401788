#!/usr/bin/env python3
"""bench.py -- headline benchmark of the grad stencil path (BASELINE.json metric: Gcells/s, all AMR levels).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config4|config2_small] [--impl reference]

One "step" = one pass of the hot path (ghost fill + c-f fill + stencil) over the synthetic workload:
  value  device-resident inputs -> device-resident outputs, CUDA-event timed, max over ranks
  e2e    the same step through the C ABI with HOST buffers: pinned H2D of the inputs, hot path, D2H of the results
  roofline  the stencil kernel alone (CUDA events around its launch inside the timed loop) against the measured
            HBM copy bandwidth of MEASURED_PEAKS.json; bytes = SURVEY 8(d) algorithmic bytes
  cpu_baseline  the compiled reference (oracle/_ref/grad3d.timed.ex, all host cores) on a bounded sample
N > 1 (torchrun, one process per GPU): the boxes of the same workload are SFC-distributed over the ranks (strong
scaling).  --transport peer (default): every rank maps its peers' slabs through CUDA IPC and the stencil kernel's TMA
producer reads cross-rank neighbour planes / rows / columns in place over NVLink -- no pack, no exchange, no halo
kernel.  --transport slab: the cross-rank ghost cells move as NCCL send/recv of packed slabs between a pack and a fill
kernel (what an MPI code does).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gcells/s (all AMR levels) for grad"
UNIT = "Gcells/s"


def workload_spec(name):
    from peleanalysis_b200 import synth
    if name == "config2":
        return dict(n=512, mgs=128, names=list(synth.FIELD_NAMES), desc="grad, uniform 512^3, 64 boxes of 128^3, 5 components (BASELINE configs[1])")
    if name == "config4":
        return dict(n=1024, mgs=128, names=["temp"], desc="grad, uniform 1024^3, 512 boxes of 128^3, 1 component (BASELINE configs[3])")
    if name == "config2_small":
        return dict(n=256, mgs=128, names=list(synth.FIELD_NAMES), desc="grad, uniform 256^3, 8 boxes of 128^3, 5 components (smoke-size)")
    raise SystemExit("unknown workload " + name)


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the compiled, unmodified reference on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------------------------
def reference_sample(spec):
    """Bounded sample: same structure (uniform periodic level, 128^3 boxes, differentiate `temp`), 256^3 cells."""
    from peleanalysis_b200 import synth
    n = min(spec["n"], 256)
    return synth.make_hierarchy(n, [], [], spec["mgs"], ("temp",)), "grad of temp on a uniform periodic %d^3 level in %d^3 boxes (1/%d of the workload's cells, 1 of its %d variables)" % (
        n, spec["mgs"], (spec["n"] // n) ** 3, len(spec["names"]))


def run_reference(spec, steps, warmup):
    from oracle import oracle as O
    from peleanalysis_b200 import plotfile
    if not O.have_ref():
        raise RuntimeError("oracle/_ref is missing (built by __graft_entry__.build() where /root/reference exists)")
    pf, sample = reference_sample(spec)
    cells = sum(l.ncells for l in pf.levels)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    tmp = tempfile.mkdtemp(prefix="pa_ref_", dir=base)
    cores = os.cpu_count() or 1
    try:
        d = os.path.join(tmp, "plt")
        plotfile.write_plotfile(d, pf, clean="remove")
        times = []
        for it in range(warmup + steps):
            _, hot = O.run_ref("grad", d, d + "_gt", threads=cores, timed=True, gradVar="temp")
            shutil.rmtree(d + "_gt", ignore_errors=True)
            if it >= warmup:
                times.append(hot)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    t = float(np.mean(times))
    return cells / t / 1e9, t, cores, sample


def reference_arm(args, spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        val, t, cores, sample = run_reference(spec, args.steps, args.warmup)
    except Exception as e:  # the oracle always exists; this only triggers on a broken checkout
        print(json.dumps({"impl": "reference", "unavailable": str(e).splitlines()[0][:200]}))
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["desc"], "timed_region": "grad.cpp:151-236 minus FillVar (hot path, host memory to host memory)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples if len(s) > 2 + i)]
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


class DevArray:
    """Wrap a raw device pointer as a torch tensor (via __cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def gen_fields(torch, levels, local_boxes, names, prob_hi=(1.0, 1.0, 1.0)):
    """Analytic fields (same formulas as peleanalysis_b200.synth) generated on the GPU, returned as pinned host
    tensors in host-concat order: one tensor per component per level."""
    import math
    out = []
    tp = 2.0 * math.pi
    for l, lv in enumerate(levels):
        comps = [[] for _ in names]
        for b in local_boxes[l]:
            lo, hi = lv.boxes[b]
            ax = [(torch.arange(lo[d], hi[d] + 1, device="cuda", dtype=torch.float64) - lv.domain_lo[d] + 0.5) * lv.dx[d] / prob_hi[d] for d in range(3)]
            X, Y, Z = ax[0][None, None, :], ax[1][None, :, None], ax[2][:, None, None]
            r = torch.sqrt((X - 0.5) ** 2 + (Y - 0.5) ** 2 + (Z - 0.5) ** 2)
            for c, n in enumerate(names):
                if n == "temp":
                    v = 300.0 + 750.0 * (1.0 + torch.tanh((0.25 - r) / 0.05)) + 5.0 * torch.sin(tp * X) * torch.cos(2 * tp * Y) + 0.0 * Z
                elif n == "x_velocity":
                    v = torch.sin(tp * X) * torch.cos(tp * Y) * torch.cos(tp * Z)
                elif n == "y_velocity":
                    v = -torch.cos(tp * X) * torch.sin(tp * Y) * torch.cos(tp * Z)
                elif n == "z_velocity":
                    v = 0.3 * torch.sin(2 * tp * Z) * torch.cos(tp * X) + 0.0 * Y
                else:
                    v = 0.05 * (1.0 - torch.tanh((0.25 - r) / 0.05)) + 0.001 * torch.sin(tp * (X + Y + Z))
                comps[c].append(v.reshape(-1))
        lvl = []
        for c in range(len(names)):
            n = sum(int(t.numel()) for t in comps[c])
            host = torch.empty(max(n, 1), dtype=torch.float64, pin_memory=True)[:n]
            if n:
                host.copy_(torch.cat(comps[c]))
            lvl.append(host)
        out.append(lvl)
        torch.cuda.synchronize()
    return out


def bench_extra(torch, capi, synth, stream, peak, kind, steps, warmup):
    """Secondary measurements reported next to the headline line (same timing rules, N=1 only):
    curvature3 : curvature tool (default options) on BASELINE configs[2] -- 3 levels, 256^3 base, ratio 2, 64^3 boxes
    target_grad / target_curv : grad / curvature of temp on the north-star target hierarchy (3 levels, 512^3 base, 128^3 boxes)
    grad5      : grad of 12 components on a configs[4]-like 4-level hierarchy (ratios 2/4/2, 16^3 boxes)."""
    if kind == "curvature3":
        pf = synth.config3(256, 64, fill=False)
        names, desc = ["temp"], "curvature (default options), 3 levels, 256^3 base, ratio 2, 64^3 boxes (BASELINE configs[2])"
    elif kind in ("target_grad", "target_curv"):
        pf = synth.config3(512, 128, fill=False)
        names = ["temp"]
        desc = "%s, 3 levels, 512^3 base, ratio 2, 128^3 boxes (the north-star target hierarchy)" % ("grad of temp" if kind == "target_grad" else "curvature (default options)")
    else:
        pf = synth.config5(128, 16, 12, fill=False)
        names, desc = list(pf.names), "grad of 12 components, 4 levels (ratios 2/4/2), 128^3 base, 16^3 boxes (BASELINE configs[4])"
    H = capi.Hierarchy(pf.levels)
    host = gen_fields(torch, pf.levels, H.local_boxes, names)
    nv = len(names)
    fin = capi.Field(H, nv, 1)
    for l in range(H.nlev):
        for c in range(nv):
            capi.check(capi.lib().pa_field_upload_level(fin.f, l, c, host[l][c].data_ptr()))
    capi.sync()
    if kind in ("curvature3", "target_curv"):
        o = capi.CurvOpts()
        o.prog_min = min(float(h[0].min()) for h in host)
        o.prog_max = max(float(h[0].max()) for h in host)
        out = capi.Field(H, 5, 1)
        run = lambda: capi.curvature(fin, 0, 0, o, out, 0)
        alg, units = H.algorithmic_bytes(5), H.num_cells
    else:
        out = capi.Field(H, 4 * nv, 0)
        run = lambda: capi.grad(fin, 0, nv, out, 0)
        alg, units = H.algorithmic_bytes(4) * nv, H.num_cells * nv
    for _ in range(warmup):
        run()
    torch.cuda.synchronize()
    l0 = capi.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"workload": desc, "value": units / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "cells": H.num_cells,
            "boxes": [len(l.boxes) for l in pf.levels], "launches_per_step": (capi.kernel_launches() - l0) // steps,
            "algorithmic_bytes": alg, "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak, "hier_build_s": H.build_seconds}


def bench_curvature_ranks(torch, dist, capi, multigpu, synth, stream, peak, kind, steps, warmup, rank, world, transport):
    """--only-extra curvature3|target_curv under torchrun: the curvature tool (default options) with the boxes SFC-distributed
    over the ranks -- multigpu.Curvature (peer links or slab exchange in front of each pass); max over ranks, whole-job cells."""
    pf = synth.config3(256, 64, fill=False) if kind == "curvature3" else synth.config3(512, 128, fill=False)
    flags = capi.PEER_LINKS if transport == "peer" else 0
    H = capi.Hierarchy(pf.levels, (1, 1, 1), (0, 0, 0), rank, world, flags=flags)
    host = gen_fields(torch, pf.levels, H.local_boxes, ["temp"])
    state = capi.Field(H, 1, 1)
    for l in range(H.nlev):
        if H.local_cells[l]:
            capi.check(capi.lib().pa_field_upload_level(state.f, l, 0, host[l][0].data_ptr()))
    capi.sync()
    lo = torch.tensor([min([float(h[0].min()) for h in host if h[0].numel()] or [1e300])], device="cuda", dtype=torch.float64)
    hi = torch.tensor([max([float(h[0].max()) for h in host if h[0].numel()] or [-1e300])], device="cuda", dtype=torch.float64)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    o = capi.CurvOpts()
    o.prog_min, o.prog_max = float(lo.item()), float(hi.item())
    out = capi.Field(H, 5, 1)
    op = multigpu.Curvature(state, 0, o, out, 0)
    dist.barrier()
    for _ in range(warmup):
        op.run()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    l0 = capi.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        op.run()
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    alg = H.algorithmic_bytes(5)                      # this rank's boxes
    return {"workload": "curvature (default options), 3 levels, %d^3 base, boxes SFC-distributed over %d ranks (%s)" % (
                256 if kind == "curvature3" else 512, world, "peer links" if flags else "slab exchange"),
            "value": H.num_cells / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "ms_per_step": ms, "cells": H.num_cells,
            "launches_per_step": (capi.kernel_launches() - l0) // steps, "roofline_frac_rank0": alg / (ms * 1e-3) / 1e9 / peak}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--transport", default="peer", choices=["peer", "slab"])
    ap.add_argument("--no-links", action="store_true", help="materialise every ghost cell (reference data flow)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary curvature / small-box measurements")
    ap.add_argument("--only-extra", default=None, choices=["curvature3", "grad5", "target_grad", "target_curv"], help="run just one secondary measurement (profiling aid)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    spec = workload_spec(args.workload)
    if args.impl == "reference":
        return reference_arm(args, spec)

    import torch
    import torch.distributed as dist
    from peleanalysis_b200 import build, capi, multigpu, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    build.build()                       # no-op when the in-tree .so is current
    torch.cuda.set_device(local_rank)
    capi.init(local_rank)               # raises if there is no B200: no CPU fallback
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()
    capi.set_stream(stream.cuda_stream)

    if args.only_extra:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        if world > 1:
            if args.only_extra not in ("curvature3", "target_curv"):
                raise SystemExit("--only-extra under torchrun: curvature3 or target_curv")
            res = bench_curvature_ranks(torch, dist, capi, multigpu, synth, stream, peak, args.only_extra, args.steps, args.warmup, rank, world, args.transport)
            if rank == 0:
                print(json.dumps(res))
            dist.destroy_process_group()
            return
        print(json.dumps(bench_extra(torch, capi, synth, stream, peak, args.only_extra, args.steps, args.warmup)))
        return
    names = spec["names"]
    nvar = len(names)
    pf = synth.make_hierarchy(spec["n"], [], [], spec["mgs"], names, fill=False)
    t0 = time.perf_counter()
    flags = capi.NO_LINKS if args.no_links else (capi.PEER_LINKS if (world > 1 and args.transport == "peer") else 0)
    H = capi.Hierarchy(pf.levels, (1, 1, 1), (0, 0, 0), rank, world, flags=flags)
    hier_s = time.perf_counter() - t0
    cells_global = H.num_cells
    host_in = gen_fields(torch, pf.levels, H.local_boxes, names)
    fin = capi.Field(H, nvar, 1)
    fout = capi.Field(H, 4 * nvar, 0)

    def upload():
        for l in range(H.nlev):
            if H.local_cells[l]:
                for c in range(nvar):
                    capi.check(capi.lib().pa_field_upload_level(fin.f, l, c, host_in[l][c].data_ptr()))

    if flags & capi.PEER_LINKS:
        multigpu.map_peers(fin)
    upload()
    capi.sync()
    if world > 1:
        dist.barrier()                 # every rank's inputs are resident before anyone reads them in place

    # whatever the neighbour links do not cover moves as NCCL send/recv of packed slabs (nothing, for peer transport
    # on a uniform grid)
    X = multigpu.SlabExchange(fin, nvar)

    def exchange():
        X.run(0)

    def step(ev=None):
        exchange()
        capi.grad(fin, 0, nvar, fout, 0, phases=1)
        if ev is not None:
            ev[0].record(stream)
        capi.grad(fin, 0, nvar, fout, 0, phases=2)
        if ev is not None:
            ev[1].record(stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = capi.kernel_launches()
    e0.record(stream)
    for i in range(args.steps):
        step(kev[i])
    e1.record(stream)
    barrier()
    launches = capi.kernel_launches() - l0
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = cells_global * nvar / (ms_step * 1e-3) / 1e9
    stencil_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))

    # ---- e2e: host buffers through the C ABI (pinned H2D of inputs, hot path, D2H of all outputs) -------------
    host_out = [[torch.empty(max(H.local_cells[l], 1), dtype=torch.float64, pin_memory=True)[:H.local_cells[l]] for _ in range(4 * nvar)]
                for l in range(H.nlev)]

    # Pipelined by variable over three streams: H2D of variable v+1 (copy engine), hot path of v (SMs) and D2H of v's
    # four outputs (the other copy engine) overlap; PCIe is full duplex, so the step costs ~max(H2D, D2H), not the sum.
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    X1 = multigpu.SlabExchange(fin, 1) if not X.empty else None
    ev_up = [torch.cuda.Event() for _ in range(nvar)]
    ev_g = [torch.cuda.Event() for _ in range(nvar)]
    ev_end = torch.cuda.Event()
    peer = bool(flags & capi.PEER_LINKS)

    def e2e_step():
        s_up.wait_event(ev_end)             # the previous step's readers (this rank's and the peers') are done
        capi.set_stream(s_up.cuda_stream)
        for v in range(nvar):
            for l in range(H.nlev):
                if H.local_cells[l]:
                    capi.check(capi.lib().pa_field_upload_level(fin.f, l, v, host_in[l][v].data_ptr()))
            ev_up[v].record(s_up)
        for v in range(nvar):
            stream.wait_event(ev_up[v])
            capi.set_stream(stream.cuda_stream)
            if peer:
                multigpu.stream_barrier()   # peers' uploads of v landed before this rank's kernel reads them over NVLink
            if X1 is not None:
                X1.run(v)
            capi.grad(fin, v, 1, fout, 4 * v)
            ev_g[v].record(stream)
            s_dn.wait_event(ev_g[v])
            capi.set_stream(s_dn.cuda_stream)
            for l in range(H.nlev):
                if H.local_cells[l]:
                    for c in range(4 * v, 4 * v + 4):
                        capi.check(capi.lib().pa_field_download_level(fout.f, l, c, host_out[l][c].data_ptr()))
        capi.set_stream(stream.cuda_stream)
        if peer:
            multigpu.stream_barrier()       # nobody re-uploads while a peer still reads the old data
        ev_end.record(stream)

    ev_end.record(stream)
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    e2e_val = cells_global * nvar / e2e_s / 1e9
    local = H.num_local_cells
    h2d = local * nvar * 8
    d2h = local * 4 * nvar * 8

    # ---- roofline of the dominant kernel (stencil) -----------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    alg_bytes = H.algorithmic_bytes(4) * nvar          # this rank's boxes, all variables: one stencil launch
    achieved = alg_bytes / (stencil_ms * 1e-3) / 1e9
    traffic = None
    try:
        # ncu capture of one launch at N=1; a rank's launch covers its share of the boxes
        traffic = json.load(open(os.path.join(ROOT, "profiles", "stencil_traffic.json"))).get(args.workload)
        if traffic is not None and world > 1:
            traffic = int(traffic * H.num_local_cells / max(cells_global, 1))
    except Exception:
        pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    extras = None
    if world == 1 and not args.no_extras:
        del fin, fout, host_out, host_in
        extras = {}
        for kind in ("curvature3", "grad5", "target_grad", "target_curv"):
            try:
                extras[kind] = bench_extra(torch, capi, synth, stream, peak, kind, max(5, args.steps // 2), 3)
            except Exception as e:
                extras[kind] = {"error": str(e).splitlines()[0][:200]}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            v, t, cores, sample = run_reference(spec, steps=3, warmup=1)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample, "hot_path_seconds": t}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "unavailable: " + str(e).splitlines()[0][:160]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": spec["desc"], "cells": cells_global, "variables": nvar, "boxes": len(pf.levels[0].boxes),
                   "parallelism": ("1 GPU" if world == 1 else "boxes SFC-distributed over %d ranks; cross-rank ghosts: %s" % (
                       world, "read in place over NVLink (CUDA-IPC peer links, no exchange step)" if flags & capi.PEER_LINKS
                       else "NCCL send/recv of packed slabs (%d cells/step recv on rank 0)" % (X.roff[-1] // nvar))),
                   "ghosts": "materialised (PA_HIER_NO_LINKS)" if args.no_links else "same-level neighbours read in place by the stencil (neighbour links)",
                   "cache": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2; no flush needed" % ((alg_bytes) / 1e9),
                   "stencil": os.environ.get("PA_STENCIL", "tma"), "hier_build_s": hier_s},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "s_per_step": e2e_s},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "k_stencil_tma<MODE_GRAD>" if os.environ.get("PA_STENCIL", "tma") != "simple" else "k_stencil_simple<MODE_GRAD>",
                     "kernel_ms": stencil_ms, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                     "step_frac": value / world * (alg_bytes / max(H.num_local_cells * nvar, 1)) / peak},
        "cpu_baseline": cpu,
        "extras": extras,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
