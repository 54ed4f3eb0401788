#!/usr/bin/env python3
"""Run under torchrun (one process per GPU): distributed grad on SFC-partitioned boxes, cross-rank ghost cells moved
as NCCL send/recv of the packed slabs, every rank's local boxes compared bit-for-bit with the oracle.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import CASES  # noqa: E402
from oracle import oracle as O  # noqa: E402
from peleanalysis_b200 import capi, multigpu, synth  # noqa: E402


def run_case(name, pf, is_per, sym, rank, world, nvar=1, mode="slab"):
    H = capi.Hierarchy(pf.levels, is_per, sym, rank, world, flags=capi.PEER_LINKS if mode == "peer" else 0)
    fin, fout = capi.Field(H, nvar, 1), capi.Field(H, 4 * nvar, 0)
    if mode == "peer":
        multigpu.map_peers(fin)
    for v in range(nvar):
        fin.upload_fabs(v, [[f[v] for f in l.fabs] for l in pf.levels])
    capi.sync()
    dist.barrier()                      # peers' uploads are complete before anyone reads them in place
    X = multigpu.SlabExchange(fin, nvar)
    ro = X.roff
    X.run(0)
    capi.grad(fin, 0, nvar, fout, 0)
    capi.sync()
    dist.barrier()                      # nobody frees / overwrites a slab a peer may still be reading
    OH = O.OracleHier(pf, is_per, sym)
    bad = 0
    for v in range(nvar):
        want = OH.unflatten_all(OH.grad(OH.flatten(v))) if hasattr(OH, "unflatten_all") else [OH.unflatten(g) for g in OH.grad(OH.flatten(v))]
        for c in range(4):
            got = fout.download_fabs(4 * v + c)
            for l in range(len(pf.levels)):
                for b in H.local_boxes[l]:
                    if not np.array_equal(got[l][b], want[c][l][b]):
                        bad += 1
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("dist_check %-18s %-4s ranks=%d local_boxes=%s slab_cells(recv)=%d mismatching boxes=%d" % (
            name, mode, world, [len(x) for x in H.local_boxes], ro[-1] // nvar, int(t.item())), flush=True)
    return int(t.item())


def run_curv(name, pf, is_per, sym, rank, world, mode="slab", velnormal=False):
    """Default-option curvature (optionally + VelFlameNormal) through multigpu.Curvature, two steps back to back (the
    second one checks the cross-step ordering), every local box against the oracle bit for bit."""
    H = capi.Hierarchy(pf.levels, is_per, sym, rank, world, flags=capi.PEER_LINKS if mode == "peer" else 0)
    names = list(pf.names)
    cS = names.index("temp")
    state = capi.Field(H, len(names), 1)
    for v in range(len(names)):
        state.upload_fabs(v, [[f[v] for f in l.fabs] for l in pf.levels])
    OH = O.OracleHier(pf, is_per, sym)
    s = OH.flatten(cS)
    o = capi.CurvOpts()
    o.prog_min, o.prog_max = float(s.min()), float(s.max())
    o.do_velnormal = 1 if velnormal else 0
    nout = capi.curvature_num_outputs(o)
    out = capi.Field(H, nout, 1)
    cv = names.index("x_velocity") if velnormal else 0
    op = multigpu.Curvature(state, cS, o, out, 0, comp_vel=cv)
    capi.sync()
    dist.barrier()
    op.run()
    op.run()
    capi.sync()
    dist.barrier()
    if velnormal:
        r = OH.curvature_ex(s, np.stack([OH.flatten(cv + d) for d in range(3)]), o.prog_min, o.prog_max)
        want = list(r["core"]) + [r["veln"]]
    else:
        want = list(OH.curvature(s, o.prog_min, o.prog_max))
    bad = 0
    for c in range(nout):
        w = OH.unflatten(want[c])
        got = out.download_fabs(c)
        for l in range(len(pf.levels)):
            for b in H.local_boxes[l]:
                if not np.array_equal(got[l][b], w[l][b]):
                    bad += 1
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("dist_check curvature %-14s %-4s ranks=%d local_boxes=%s mismatching boxes=%d" % (
            name, mode, world, [len(x) for x in H.local_boxes], int(t.item())), flush=True)
    return int(t.item())


def run_curv_options(rank, world, mode):
    """Every curvature option at once (threshold_prog, do_gaussCurv, do_strain + ROST, do_velnormal) through
    multigpu.Curvature against the golden vectors of the compiled reference."""
    from helpers import bit_equal, fabs_from_flat, load_golden, max_rel
    pf, z = load_golden("c1_options")
    kw = dict(x.split("=") for x in z["curv_opts"])
    is_per, sym = tuple(int(v) for v in z["is_per"]), tuple(int(v) for v in z["sym_dir"])
    H = capi.Hierarchy(pf.levels, is_per, sym, rank, world, flags=capi.PEER_LINKS if mode == "peer" else 0)
    state = capi.Field(H, 4, 1)
    for v, n in enumerate(["temp", "x_velocity", "y_velocity", "z_velocity"]):
        state.upload_fabs(v, [[f[pf.comp(n)] for f in l.fabs] for l in pf.levels])
    o = capi.CurvOpts()
    o.prog_min, o.prog_max = float(z["prog_min"]), float(z["prog_max"])
    o.do_threshold, o.threshold = int(kw["threshold_prog"]), float(kw["threshold_value"])
    o.do_gauss, o.do_strain, o.get_strain_tensor, o.do_velnormal = 1, 1, 1, 1
    out = capi.Field(H, capi.curvature_num_outputs(o), 1)
    op = multigpu.Curvature(state, 0, o, out, 0, comp_vel=1)
    capi.sync()
    dist.barrier()
    op.run()
    op.run()
    capi.sync()
    dist.barrier()
    order = ["Progress", "MeanCurvature_temp", "FlameNormalX_temp", "FlameNormalY_temp", "FlameNormalZ_temp", "GaussianCurvature_temp",
             "StrainRate_temp"] + ["ROST_dU%sd%s" % (a, b) for a in "xyz" for b in "xyz"] + ["VelFlameNormal"]
    bad = 0
    for c, n in enumerate(order):
        want = fabs_from_flat(pf, z["curv_" + n])
        got = out.download_fabs(c)
        for l in range(len(pf.levels)):
            for b in H.local_boxes[l]:
                ok = max_rel(got[l][b], want[l][b]) <= 1e-12 if n.startswith("Gaussian") else bit_equal(got[l][b], want[l][b])
                bad += 0 if ok else 1
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("dist_check curvature all-options   %-4s ranks=%d mismatching boxes=%d" % (mode, world, int(t.item())), flush=True)
    return int(t.item())


def run_curv_threshold_off_centre(rank, world, mode):
    """threshold_prog on a field that is NOT mirror-symmetric about the rank boundaries (tests/test_emu_peer_links.py has the
    reasoning): an early clip of the flame normal -- before every rank's divergence has read the unclipped values of its
    peers -- shows up as a mismatch against the single-rank oracle here."""
    pf = synth.config3(64, 16)
    for lv in pf.levels:
        for (lo, hi), f in zip(lv.boxes, lv.fabs):
            ax = [(np.arange(lo[d], hi[d] + 1) - lv.domain_lo[d] + 0.5) * lv.dx[d] for d in range(3)]
            X, Y, Z = ax[0][None, None, :], ax[1][None, :, None], ax[2][:, None, None]
            rr = np.sqrt((X - 0.5) ** 2 + (Y - 0.47) ** 2 + (Z - 0.41) ** 2)
            f[0] = 300.0 + 750.0 * (1.0 + np.tanh((0.27 - rr) / 0.09)) + 9.0 * np.sin(2 * np.pi * (Z + 0.13)) * np.cos(2 * np.pi * X)
    is_per, sym = (1, 1, 1), (0, 0, 0)
    H = capi.Hierarchy(pf.levels, is_per, sym, rank, world, flags=capi.PEER_LINKS if mode == "peer" else 0)
    state = capi.Field(H, 1, 1)
    state.upload_fabs(0, [[f[0] for f in l.fabs] for l in pf.levels])
    OH = O.OracleHier(pf, is_per, sym)
    s = OH.flatten(0)
    o = capi.CurvOpts()
    o.prog_min, o.prog_max = float(s.min()), float(s.max())
    o.do_threshold, o.threshold = 1, 0.2
    out = capi.Field(H, 5, 1)
    op = multigpu.Curvature(state, 0, o, out, 0)
    capi.sync()
    dist.barrier()
    op.run()
    op.run()
    capi.sync()
    dist.barrier()
    want = OH.curvature(s, o.prog_min, o.prog_max, do_threshold=True, threshold=0.2)
    bad = 0
    for c in range(5):
        w = OH.unflatten(want[c])
        got = out.download_fabs(c)
        for l in range(len(pf.levels)):
            for b in H.local_boxes[l]:
                bad += 0 if np.array_equal(got[l][b], w[l][b]) else 1
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("dist_check curvature threshold, off-centre field %-4s ranks=%d mismatching boxes=%d" % (mode, world, int(t.item())), flush=True)
    return int(t.item())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    capi.init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    capi.set_stream(torch.cuda.current_stream().cuda_stream)
    bad = 0
    for mode in ("slab", "peer"):
        for name in ["c1_periodic", "c1_walls", "lshape", "edge_periodic", "ratio4", "c3_three_levels"]:
            builder, is_per, sym, _, _ = CASES[name]
            bad += run_case(name, builder(), is_per, sym, rank, world, mode=mode)
        bad += run_case("config3_64", synth.config3(64, 16), (1, 1, 1), (0, 0, 0), rank, world, mode=mode)
        bad += run_case("config1_5vars", synth.config1(32, 16, names=synth.FIELD_NAMES), (1, 1, 1), (0, 0, 0), rank, world, nvar=5, mode=mode)
        bad += run_case("config5_small", synth.config5(base=32, mgs=8, ncomp=2), (1, 1, 1), (0, 0, 0), rank, world, nvar=2, mode=mode)
        bad += run_case("uniform_64", synth.make_hierarchy(64, [], [], 16, ("temp",)), (1, 1, 1), (0, 0, 0), rank, world, mode=mode)
        for fused in ("0", "1"):                   # the separate NORMAL_S / DIV kernels (default), then the fused kernel + shell pass
            os.environ["PA_CURV_FUSED"] = fused
            tag = "+fused" if fused == "1" else ""
            bad += run_curv("config1" + tag, synth.config1(32, 16), (1, 1, 1), (0, 0, 0), rank, world, mode=mode)
            bad += run_curv("config3_64" + tag, synth.config3(64, 16), (1, 1, 1), (0, 0, 0), rank, world, mode=mode)
            bad += run_curv("c1_walls_vn" + tag, synth.config1(32, 16, names=synth.FIELD_NAMES, corner=True), (0, 0, 0), (1, 0, 0), rank, world, mode=mode, velnormal=True)
            bad += run_curv("uniform_64" + tag, synth.make_hierarchy(64, [], [], 16, ("temp",)), (1, 1, 1), (0, 0, 0), rank, world, mode=mode)
            bad += run_curv_options(rank, world, mode)
            bad += run_curv_threshold_off_centre(rank, world, mode)
        os.environ["PA_CURV_FUSED"] = "0"
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "OK" if bad == 0 else "FAILED (%d)" % bad, flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
