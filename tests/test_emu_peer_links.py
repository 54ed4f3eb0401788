"""CPU tier: PEER LINKS under the emulator.  Two (or four) ranks of one hierarchy are stood up inside a single process --
the emulator's "IPC handle" simply carries the slab pointer -- so the stencil producer's in-place reads of another rank's
slab (planes, rows, x-ghost columns), the halo gather from peer-owned boxes and the slab exchange of whatever links do
not cover all run on the CPU, for grad and for the two-pass curvature sequence of multigpu.Curvature.  Every rank's boxes
are compared bit for bit with the oracle.  (On GPUs the same is checked by tests/dist_check.py over NVLink / NCCL.)"""
import ctypes as C
import os

import numpy as np
import pytest

from cases import CASES
from oracle import oracle as O
from peleanalysis_b200 import synth
from test_emu_parity import emu  # noqa: F401  (fixture)


class Ranks:
    """nranks hierarchies of the same levels in one process, with the cross-rank steps done by hand."""

    def __init__(self, capi, pf, is_per, sym, nranks, flags):
        self.capi, self.pf, self.n = capi, pf, nranks
        self.H = [capi.Hierarchy(pf.levels, is_per, sym, r, nranks, flags=flags) for r in range(nranks)]

    def fields(self, ncomp, ng):
        F = [self.capi.Field(h, ncomp, ng) for h in self.H]
        if self.H[0].flags & self.capi.PEER_LINKS:
            handles = [f.ipc_handles() for f in F]
            for r, f in enumerate(F):
                for q in range(self.n):
                    if q != r:
                        f.map_peer(q, handles[q])
        return F

    def exchange(self, F, comp, ncomp):
        """What multigpu.SlabExchange does over NCCL: pack everywhere, move send slab segments into the peers' recv slabs."""
        lib = self.capi.lib()
        bufs = []
        for f in F:
            sp, rp = C.c_void_p(), C.c_void_p()
            so = (C.c_int64 * (self.n + 1))()
            ro = (C.c_int64 * (self.n + 1))()
            self.capi.check(lib.pa_exchange_buffers(f.f, ncomp, C.byref(sp), C.byref(rp), so, ro))
            bufs.append((sp.value, rp.value, list(so), list(ro)))
        for f in F:
            self.capi.check(lib.pa_exchange_pack(f.f, comp, ncomp))
        self.capi.sync()
        moved = 0
        for a in range(self.n):
            for b in range(self.n):
                if a == b:
                    continue
                n = bufs[a][2][b + 1] - bufs[a][2][b]                      # a sends to b ...
                assert n == bufs[b][3][a + 1] - bufs[b][3][a]              # ... what b expects from a
                if n:
                    C.memmove(bufs[b][1] + 8 * bufs[b][3][a], bufs[a][0] + 8 * bufs[a][2][b], 8 * n)
                    moved += n
        for f in F:
            self.capi.check(lib.pa_exchange_mark_received(f.f, comp, ncomp))
        return moved

    def check(self, F, comps, want, tag):
        OHu = self.OH
        for c, w in zip(comps, want):
            wl = OHu.unflatten(w)
            for r, f in enumerate(F):
                got = f.download_fabs(c)
                for l in range(len(self.pf.levels)):
                    for b in self.H[r].local_boxes[l]:
                        assert np.array_equal(got[l][b].view(np.int64), wl[l][b].view(np.int64)), (tag, c, r, l, b)


CASE_LIST = [("c1_periodic", 2), ("c1_walls", 2), ("lshape", 2), ("edge_periodic", 2), ("c3_three_levels", 2), ("mixed_boxes", 2),
             ("uniform", 2), ("uniform", 4), ("uniform", 8), ("config3", 4)]


def _case(name):
    if name == "uniform":
        return synth.make_hierarchy(32, [], [], 8, ("temp",)), (1, 1, 1), (0, 0, 0)
    if name == "config3":
        return synth.config3(32, 8), (1, 1, 1), (0, 0, 0)
    b, per, sym, _, _ = CASES[name]
    return b(), per, sym


@pytest.mark.parametrize("transport", ["peer", "slab"])
@pytest.mark.parametrize("name,nranks", CASE_LIST)
def test_emulated_multi_rank_in_one_process(emu, name, nranks, transport):  # noqa: F811
    if transport == "slab" and (name, nranks) not in (("c3_three_levels", 2), ("mixed_boxes", 2), ("config3", 4)):
        pytest.skip("slab transport: three representative cases (tests/test_emu_multirank.py runs it over gloo as well)")
    capi = emu
    os.environ["CUEMU_SEED"] = "3"
    os.environ["PA_STENCIL"] = "tma"
    os.environ["PA_TMA_SMALL"] = "1"
    try:
        pf, is_per, sym = _case(name)
        R = Ranks(capi, pf, is_per, sym, nranks, capi.PEER_LINKS if transport == "peer" else 0)
        R.OH = O.OracleHier(pf, is_per, sym)
        s = R.OH.flatten(0)
        # ---- grad
        fin, fout = R.fields(1, 1), R.fields(4, 0)
        for f in fin:
            f.upload_fabs(0, [[x[0] for x in l.fabs] for l in pf.levels])
        moved = R.exchange(fin, 0, 1)
        if transport == "peer" and name == "uniform":
            assert moved == 0                       # a uniform grid needs no exchange at all with peer links
        for r in range(nranks):
            capi.grad(fin[r], 0, 1, fout[r], 0)
        capi.sync()
        R.check(fout, range(4), R.OH.grad(s), "grad")
        # ---- curvature, the sequence of multigpu.Curvature.run, twice
        o = capi.CurvOpts()
        o.prog_min, o.prog_max = float(s.min()), float(s.max())
        out = R.fields(5, 1)
        for _ in range(2):
            R.exchange(fin, 0, 1)
            for r in range(nranks):
                capi.curvature_phases(fin[r], 0, 0, o, out[r], 0, 1)
            capi.sync()
            R.exchange(out, 2, 3)
            for r in range(nranks):
                capi.curvature_phases(fin[r], 0, 0, o, out[r], 0, 2)
            capi.sync()
        if all(x == 2 for x in R.OH.ratios):
            R.check(out, range(5), R.OH.curvature(s, o.prog_min, o.prog_max), "curvature")
    finally:
        os.environ["CUEMU_SEED"] = "0"


@pytest.mark.parametrize("transport", ["peer", "slab"])
@pytest.mark.parametrize("name,nranks", [("c3_threshold", 2), ("c1_options", 2), ("c1_options", 3)])
def test_emulated_multi_rank_curvature_options(emu, name, nranks, transport):  # noqa: F811
    """threshold_prog (the flame normal exchanged before every level), do_gaussCurv (the internal gradient field exchanged),
    do_strain (velocities exchanged), do_velnormal: the step sequence of multigpu.Curvature.run against the golden vectors
    of the compiled reference."""
    from helpers import bit_equal, load_golden, max_rel
    capi = emu
    os.environ["CUEMU_SEED"] = "5"
    os.environ["PA_STENCIL"] = "tma"
    os.environ["PA_TMA_SMALL"] = "1"
    try:
        pf, z = load_golden(name)
        kw = dict(s.split("=") for s in z["curv_opts"])
        is_per, sym = tuple(int(v) for v in z["is_per"]), tuple(int(v) for v in z["sym_dir"])
        R = Ranks(capi, pf, is_per, sym, nranks, capi.PEER_LINKS if transport == "peer" else 0)
        o = capi.CurvOpts()
        o.prog_min, o.prog_max = float(z["prog_min"]), float(z["prog_max"])
        o.do_threshold = int(kw.get("threshold_prog", 0))
        o.threshold = float(kw.get("threshold_value", 1e-4))
        o.do_gauss, o.do_strain = int(kw.get("do_gaussCurv", 0)), int(kw.get("do_strain", 0))
        o.get_strain_tensor, o.do_velnormal = int(kw.get("getStrainTensor", 0)), int(kw.get("do_velnormal", 0))
        need_vel = o.do_strain or o.do_velnormal
        names = ["temp"] + (["x_velocity", "y_velocity", "z_velocity"] if need_vel else [])
        state = R.fields(len(names), 1)
        for f in state:
            for v, n in enumerate(names):
                f.upload_fabs(v, [[x[pf.comp(n)] for x in l.fabs] for l in pf.levels])
        nout = capi.curvature_num_outputs(o)
        out = R.fields(nout, 1)
        scratch = None
        if o.do_gauss:
            scratch = [capi.ScratchField(h) for h in R.H]
            if transport == "peer":
                handles = [f.ipc_handles() for f in scratch]
                for r, f in enumerate(scratch):
                    for q in range(nranks):
                        if q != r:
                            f.map_peer(q, handles[q])
        nlev = len(pf.levels)

        def step(steps, lo=-1, hi=-1):
            for r in rank_order:
                capi.curvature_steps(state[r], 0, 1, o, out[r], 0, steps, lo, hi)
            capi.sync()
        for rep in range(2):
            # the ranks run one after the other here: both orders must give the same bits (an in-place update that a peer
            # still has to read -- the threshold clip of n -- shows up as an order dependence)
            rank_order = list(range(nranks)) if rep == 0 else list(reversed(range(nranks)))
            R.exchange(state, 0, 1)
            step(capi.CURV_PASS1)
            if o.do_threshold:
                for l in range(nlev):
                    R.exchange(out, 2, 3)
                    step(capi.CURV_DIV, l, l)
                    step(capi.CURV_CLIP, l, l)
            else:
                R.exchange(out, 2, 3)
                step(capi.CURV_DIV)
            if o.do_gauss:
                R.exchange(scratch, 0, 3)
                step(capi.CURV_GAUSS)
            if o.do_strain:
                R.exchange(state, 1, 3)
                step(capi.CURV_STRAIN)
            if o.do_velnormal:
                step(capi.CURV_VELN)
        # compare every rank's boxes with the reference's golden vectors
        order = ["Progress", "MeanCurvature_temp", "FlameNormalX_temp", "FlameNormalY_temp", "FlameNormalZ_temp"]
        if o.do_gauss:
            order.append("GaussianCurvature_temp")
        if o.do_strain:
            order.append("StrainRate_temp")
            if o.get_strain_tensor:
                order += ["ROST_dU%sd%s" % (a, b) for a in "xyz" for b in "xyz"]
        if o.do_velnormal:
            order.append("VelFlameNormal")
        from helpers import fabs_from_flat
        for c, n in enumerate(order):
            want = fabs_from_flat(pf, z["curv_" + n])
            for r in range(nranks):
                got = out[r].download_fabs(c)
                for l in range(nlev):
                    for b in R.H[r].local_boxes[l]:
                        if n.startswith("Gaussian"):
                            assert max_rel(got[l][b], want[l][b]) <= 1e-12 or np.allclose(got[l][b], want[l][b], rtol=1e-12, atol=0), (n, r, l, b)
                        else:
                            assert bit_equal(got[l][b], want[l][b]), (name, transport, n, r, l, b)
    finally:
        os.environ["CUEMU_SEED"] = "0"


@pytest.mark.parametrize("fused", ["0", "1", "3", "n3", "nw"])
@pytest.mark.parametrize("transport", ["peer", "slab"])
@pytest.mark.parametrize("nranks", [2, 3])
def test_emulated_multi_rank_threshold_off_centre_field(emu, nranks, transport, fused):  # noqa: F811
    """threshold_prog on a field that is NOT mirror-symmetric about the rank boundaries (the golden `temp` is: the cells either
    side of every cross-rank face hold equal progress values, so a clip of the flame normal that runs too early -- before a
    peer's divergence has read the unclipped values, curvature.cpp:487-567 -- could not be seen).  Ranks run one after the
    other in both orders; every order must reproduce the single-rank oracle bit for bit."""
    from oracle import oracle as O
    from peleanalysis_b200 import synth
    capi = emu
    os.environ["CUEMU_SEED"] = "5"
    os.environ["PA_STENCIL"] = "tma"
    os.environ["PA_TMA_SMALL"] = "1"
    # "1" / "3": a fused kernel + shell pass on every rank; "n3" / "nw": the plane-staged / barrier-free flame-normal kernels
    os.environ["PA_CURV_FUSED"] = fused if fused in ("0", "1", "3") else "0"
    os.environ["PA_NORMAL_F3"] = "1" if fused == "n3" else "0"
    os.environ["PA_NORMAL_W"] = "1" if fused == "nw" else "0"
    if nranks == 3 and fused in ("3", "n3", "nw"):
        pytest.skip("combination not in the thinned-out matrix")
    try:
        pf = synth.config3(16, 8)
        for lv in pf.levels:                                    # off-centre blob + a z-dependent ripple
            for (lo, hi), f in zip(lv.boxes, lv.fabs):
                ax = [(np.arange(lo[d], hi[d] + 1) - lv.domain_lo[d] + 0.5) * lv.dx[d] for d in range(3)]
                X, Y, Z = ax[0][None, None, :], ax[1][None, :, None], ax[2][:, None, None]
                rr = np.sqrt((X - 0.5) ** 2 + (Y - 0.47) ** 2 + (Z - 0.41) ** 2)
                f[0] = 300.0 + 750.0 * (1.0 + np.tanh((0.27 - rr) / 0.09)) + 9.0 * np.sin(2 * np.pi * (Z + 0.13)) * np.cos(2 * np.pi * X)
        is_per, sym = (1, 1, 1), (0, 0, 0)
        OH = O.OracleHier(pf, is_per, sym)
        s = OH.flatten(0)
        o = capi.CurvOpts()
        o.prog_min, o.prog_max = float(s.min()), float(s.max())
        o.do_threshold, o.threshold = 1, 0.2
        want = OH.curvature(s, o.prog_min, o.prog_max, do_threshold=True, threshold=0.2)
        assert np.count_nonzero(want[1] == 0.0) > 100 and np.count_nonzero(want[1]) > 100   # the clip is active, and not everywhere
        for order in (list(range(nranks)), list(reversed(range(nranks)))):
            R = Ranks(capi, pf, is_per, sym, nranks, capi.PEER_LINKS if transport == "peer" else 0)
            R.OH = OH
            state, out = R.fields(1, 1), R.fields(5, 1)
            for f in state:
                f.upload_fabs(0, [[x[0] for x in l.fabs] for l in pf.levels])

            def step(steps, lo=-1, hi=-1):
                for r in order:
                    capi.curvature_steps(state[r], 0, 0, o, out[r], 0, steps, lo, hi)
                capi.sync()
            R.exchange(state, 0, 1)
            step(capi.CURV_PASS1)
            for l in range(len(pf.levels)):
                R.exchange(out, 2, 3)
                step(capi.CURV_DIV, l, l)
                step(capi.CURV_CLIP, l, l)
            R.check(out, range(5), want, "curvature, threshold_prog, rank order %s" % order)
    finally:
        os.environ["CUEMU_SEED"] = "0"
        os.environ["PA_CURV_FUSED"] = "0"
        os.environ["PA_NORMAL_F3"] = "0"
        os.environ["PA_NORMAL_W"] = "0"
