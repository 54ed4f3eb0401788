"""CPU tier, world_size 2 over gloo: the multi-rank SLAB transport end to end -- pack kernel, send/recv of the packed ghost
slabs, halo / coarse-fine fill from the recv slab, stencil -- with the kernel sources running under the CUDA
execution-model emulator of tests/emu (test infrastructure, see tests/test_emu_parity.py), every rank's boxes compared
bit for bit with the oracle.  grad and the two-pass curvature driver (multigpu.Curvature), the same checks
tests/dist_check.py makes on real GPUs over NCCL.  Peer links (CUDA IPC) are a GPU-only matter."""
import ctypes as C
import importlib.util
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_emulated():
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from peleanalysis_b200 import capi as product_capi, multigpu as product_multigpu

    def private(mod, name):
        spec = importlib.util.spec_from_file_location("peleanalysis_b200." + name, mod.__file__)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    capi = private(product_capi, "_capi_emulated")
    capi.LIB_PATH = build_emu.build()
    mg = private(product_multigpu, "_multigpu_emulated")
    mg.capi = capi
    os.environ["PA_NORMAL_MATH"] = "fast"
    return capi, mg


def _wrap_host(ptr, n):
    return torch.from_numpy(np.ctypeslib.as_array((C.c_double * int(n)).from_address(int(ptr))))


def _worker(rank, world, port, ok):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CUEMU_SEED"] = str(1 + rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cases import CASES
        from oracle import oracle as O
        from peleanalysis_b200 import synth
        capi, mg = _load_emulated()
        capi.init(0)
        bad = []
        exchanged = 0

        def check(name, out, want_flat, OH, pf, H, comps):
            for c in comps:
                w = OH.unflatten(want_flat[c])
                got = out.download_fabs(c)
                for l in range(len(pf.levels)):
                    for b in H.local_boxes[l]:
                        if not np.array_equal(got[l][b], w[l][b]):
                            bad.append((name, c, l, b))

        grad_cases = [(n,) + tuple(CASES[n][:3]) for n in ("c1_periodic", "c1_walls", "lshape", "ratio4", "c3_three_levels")]
        grad_cases = [(n, b(), per, sym) for n, b, per, sym in grad_cases]
        grad_cases.append(("config5_small", synth.config5(base=16, mgs=8, ncomp=2), (1, 1, 1), (0, 0, 0)))
        for name, pf, is_per, sym in grad_cases:
            H = capi.Hierarchy(pf.levels, is_per, sym, rank, world)
            fin, fout = capi.Field(H, 1, 1), capi.Field(H, 4, 0)
            fin.upload_fabs(0, [[f[pf.comp(pf.names[0])] for f in l.fabs] for l in pf.levels])
            X = mg.SlabExchange(fin, 1, _wrap_host)
            exchanged += 0 if X.empty else 1      # (ratio4: one box per level, everything on rank 0)
            X.run(0)
            capi.grad(fin, 0, 1, fout, 0)
            capi.sync()
            OH = O.OracleHier(pf, is_per, sym)
            check("grad " + name, fout, OH.grad(OH.flatten(pf.comp(pf.names[0]))), OH, pf, H, range(4))

        curv_cases = [("config1", synth.config1(32, 16), (1, 1, 1), (0, 0, 0), False),
                      ("config3", synth.config3(32, 16), (1, 1, 1), (0, 0, 0), False),
                      ("c1_walls_vn", synth.config1(32, 16, names=synth.FIELD_NAMES, corner=True), (0, 0, 0), (1, 0, 0), True)]
        for name, pf, is_per, sym, veln in curv_cases:
            H = capi.Hierarchy(pf.levels, is_per, sym, rank, world)
            names = list(pf.names)
            cS = names.index("temp")
            state = capi.Field(H, len(names), 1)
            for v in range(len(names)):
                state.upload_fabs(v, [[f[v] for f in l.fabs] for l in pf.levels])
            OH = O.OracleHier(pf, is_per, sym)
            s = OH.flatten(cS)
            o = capi.CurvOpts()
            o.prog_min, o.prog_max = float(s.min()), float(s.max())
            o.do_velnormal = 1 if veln else 0
            nout = capi.curvature_num_outputs(o)
            out = capi.Field(H, nout, 1)
            cv = names.index("x_velocity") if veln else 0
            op = mg.Curvature(state, cS, o, out, 0, comp_vel=cv, wrap=_wrap_host)
            op.run()
            op.run()                      # twice: the second step checks nothing stale survives in the slabs
            capi.sync()
            if veln:
                r = OH.curvature_ex(s, np.stack([OH.flatten(cv + d) for d in range(3)]), o.prog_min, o.prog_max)
                want = list(r["core"]) + [r["veln"]]
            else:
                want = list(OH.curvature(s, o.prog_min, o.prog_max))
            check("curvature " + name, out, want, OH, pf, H, range(nout))
        # every option at once through multigpu.Curvature: threshold_prog (normal exchanged per level), do_gaussCurv (the
        # internal gradient field exchanged), do_strain + ROST (velocities exchanged), do_velnormal; golden vectors of the
        # compiled reference
        from helpers import bit_equal, fabs_from_flat, load_golden, max_rel
        pf, z = load_golden("c1_options")
        kw = dict(x.split("=") for x in z["curv_opts"])
        is_per, sym = tuple(int(v) for v in z["is_per"]), tuple(int(v) for v in z["sym_dir"])
        H = capi.Hierarchy(pf.levels, is_per, sym, rank, world)
        names = ["temp", "x_velocity", "y_velocity", "z_velocity"]
        state = capi.Field(H, 4, 1)
        for v, n in enumerate(names):
            state.upload_fabs(v, [[f[pf.comp(n)] for f in l.fabs] for l in pf.levels])
        o = capi.CurvOpts()
        o.prog_min, o.prog_max = float(z["prog_min"]), float(z["prog_max"])
        o.do_threshold, o.threshold = int(kw["threshold_prog"]), float(kw["threshold_value"])
        o.do_gauss, o.do_strain, o.get_strain_tensor, o.do_velnormal = 1, 1, 1, 1
        out = capi.Field(H, capi.curvature_num_outputs(o), 1)
        op = mg.Curvature(state, 0, o, out, 0, comp_vel=1, wrap=_wrap_host)
        op.run()
        op.run()
        capi.sync()
        order = ["Progress", "MeanCurvature_temp", "FlameNormalX_temp", "FlameNormalY_temp", "FlameNormalZ_temp", "GaussianCurvature_temp",
                 "StrainRate_temp"] + ["ROST_dU%sd%s" % (a, b) for a in "xyz" for b in "xyz"] + ["VelFlameNormal"]
        for c, n in enumerate(order):
            want = fabs_from_flat(pf, z["curv_" + n])
            got = out.download_fabs(c)
            for l in range(len(pf.levels)):
                for b in H.local_boxes[l]:
                    same = max_rel(got[l][b], want[l][b]) <= 1e-12 if n.startswith("Gaussian") else bit_equal(got[l][b], want[l][b])
                    if not same:
                        bad.append(("options", n, l, b))
        assert not bad, bad[:10]
        assert exchanged >= 4
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def test_emulated_slab_transport_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0, 0])
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ok)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    assert list(ok) == [1, 1], [p.exitcode for p in procs]


def _tool_worker(rank, world, port, tmp, ok):
    """peleanalysis_b200.mgtools (the multi-process plotfile tools) end to end: plotfile in -> one Cell_D file per rank +
    Header / Cell_H from rank 0 -> read back and compared with the golden vectors of the compiled reference."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), CUEMU_SEED=str(1 + rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import bit_equal, load_golden, max_rel
        from peleanalysis_b200 import mgtools, plotfile
        capi, mg = _load_emulated()

        def flat(r, n):
            c = r.comp(n)
            return np.concatenate([f[c].ravel() for l in r.levels for f in l.fabs])
        for name in ("c3_three_levels", "mixed_boxes", "c1_corner_sym"):
            pf, z = load_golden(name)
            d = os.path.join(tmp, "plt_" + name)
            if rank == 0:
                plotfile.write_plotfile(d, pf)
            dist.barrier()
            per = " ".join(str(int(v)) for v in z["is_per"])
            sym = " ".join(str(int(v)) for v in z["sym_dir"])
            out = mgtools.run("grad", ["infile=" + d, "gradVar=temp", "is_per=" + per, "sym_dir=" + sym, "transport=slab",
                                       "outfile=" + os.path.join(tmp, "gt_" + name)], capi=capi, multigpu=mg, wrap=_wrap_host, backend="gloo", device=0)
            dist.barrier()
            if rank == 0:
                r = plotfile.read_plotfile(out)
                assert r.names == ["temp", "temp_gx", "temp_gy", "temp_gz", "||gradtemp||"]
                files = {f for l in range(len(pf.levels)) for f in os.listdir(os.path.join(out, "Level_%d" % l))}
                assert files == {"Cell_D_00000", "Cell_D_00001", "Cell_H"}, files          # every rank wrote its own boxes
                assert bit_equal(flat(r, "temp"), z["in_temp"])
                for k, n in zip(["gx", "gy", "gz", "mag"], r.names[1:]):
                    assert bit_equal(flat(r, n), z["grad_" + k]), (name, n)
                from oracle import oracle as O
                if O.have_ref():                                  # one file per rank: AMReX's fcompare reads it like any plotfile
                    import subprocess
                    ref = os.path.join(tmp, "ref_" + name)
                    O.run_ref("grad", d, ref, gradVar="temp", is_per=list(z["is_per"]), sym_dir=list(z["sym_dir"]))
                    q = subprocess.run([O.ref_exe("fcompare.ref.ex"), out, ref], capture_output=True, text=True)
                    assert "PLOTFILE AGREE" in q.stdout, q.stdout[-1500:]
        # curvature with every option, inputs file instead of key=value arguments
        pf, z = load_golden("c1_options")
        d = os.path.join(tmp, "plt_opts")
        inp = os.path.join(tmp, "inputs.curv")
        if rank == 0:
            plotfile.write_plotfile(d, pf)
            with open(inp, "w") as f:
                f.write("infile = %s\noutfile = %s  # all options\nprogressName = temp\nis_per = %s\ntransport = slab\n" % (
                    d, os.path.join(tmp, "K_opts"), " ".join(str(int(v)) for v in z["is_per"])))
                for kv in z["curv_opts"]:
                    f.write(str(kv).replace("=", " = ") + "\n")
        dist.barrier()
        out = mgtools.run("curvature", [inp], capi=capi, multigpu=mg, wrap=_wrap_host, backend="gloo", device=0)
        dist.barrier()
        if rank == 0:
            r = plotfile.read_plotfile(out)
            for key in z.files:
                if key.startswith("curv_") and key != "curv_opts":
                    n = key[5:]
                    got = flat(r, n)
                    if n.startswith("GaussianCurvature"):
                        assert max_rel(got, z[key]) <= 1e-12
                    else:
                        assert bit_equal(got, z[key]), n
            assert "SmoothedProgress" in r.names
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def test_emulated_plotfile_tools_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0, 0])
    procs = [ctx.Process(target=_tool_worker, args=(r, 2, port, str(tmp_path), ok)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    assert list(ok) == [1, 1], [p.exitcode for p in procs]


def test_emulated_plotfile_tool_single_process(tmp_path):
    """The same tool without torchrun (one rank), against the golden vectors."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import bit_equal, load_golden
    from peleanalysis_b200 import mgtools, plotfile
    capi, mg = _load_emulated()
    old = {k: os.environ.pop(k, None) for k in ("RANK", "WORLD_SIZE")}
    try:
        pf, z = load_golden("c1_periodic")
        d = str(tmp_path / "plt")
        plotfile.write_plotfile(d, pf)
        cwd = os.getcwd()
        os.chdir(tmp_path)
        try:
            out = mgtools.run("grad", ["infile=" + d], capi=capi, device=0)
        finally:
            os.chdir(cwd)
        assert out == "plt_gt"                                    # default outfile = <root>_gt in the cwd, like the reference
        r = plotfile.read_plotfile(str(tmp_path / "plt_gt"))
        got = np.concatenate([f[r.comp("temp_gx")].ravel() for l in r.levels for f in l.fabs])
        assert bit_equal(got, z["grad_gx"])
        from oracle import oracle as O
        if O.have_ref():                                          # AMReX's own fcompare reads it and agrees with the reference tool
            import subprocess
            O.run_ref("grad", d, str(tmp_path / "ref_gt"), gradVar="temp")
            p = subprocess.run([O.ref_exe("fcompare.ref.ex"), str(tmp_path / "plt_gt"), str(tmp_path / "ref_gt")], capture_output=True, text=True)
            assert "PLOTFILE AGREE" in p.stdout, p.stdout[-1500:]
    finally:
        for k, v in old.items():
            if v is not None:
                os.environ[k] = v
