"""CPU, world_size 2 over gloo: the multi-rank exchange plan (hier.cpp: build_exchange).  Each rank builds its own
hierarchy tables, ships the identity of the cell in every send-slab slot to its peer with torch.distributed
send/recv, and checks that every recv-slab slot holds exactly the cell its halo rows / coarse gather index expect.
Also checks that single-rank tables are the union of the two ranks' local tables."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, cases, ok, flags=0):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for case in cases:                      # one process group for all cases: spawning and importing torch dominate otherwise
            _check_case(rank, world, case, flags)
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def _check_case(rank, world, case, flags):
    if True:
        from cases import CASES
        from peleanalysis_b200 import capi
        builder, is_per, sym, _, _ = CASES[case]
        pf = builder()
        H = capi.Hierarchy(pf.levels, is_per, sym, rank, world, flags=flags)
        send_ids, want = H.exchange_ids(0), H.exchange_ids(1)
        sc, rc = H.exchange_prefix(1)
        assert sc[rank] == 0 and rc[rank] == 0
        assert sc.sum() == send_ids.size and rc.sum() == want.size
        assert (want >= 0).all() and (send_ids >= 0).all()
        so = np.concatenate([[0], np.cumsum(sc)])
        ro = np.concatenate([[0], np.cumsum(rc)])
        got = np.full(want.size, -1, dtype=np.int64)
        reqs, bufs = [], []
        for p in range(world):
            if p == rank:
                continue
            if rc[p]:
                t = torch.empty(int(rc[p]), dtype=torch.int64)
                bufs.append((p, t))
                reqs.append(dist.irecv(t, p))
            if sc[p]:
                reqs.append(dist.isend(torch.from_numpy(send_ids[so[p]:so[p + 1]].copy()), p))
        for r in reqs:
            r.wait()
        for p, t in bufs:
            got[ro[p]:ro[p + 1]] = t.numpy()
        assert np.array_equal(got, want), int((got != want).sum())
        # every box is owned by exactly one rank and the local cell counts add up
        tot = torch.tensor([H.num_local_cells], dtype=torch.int64)
        dist.all_reduce(tot)
        assert int(tot.item()) == H.num_cells
        # remote ghost sources are flagged (-2) exactly where the single-rank table names a box owned by the peer
        H1 = capi.Hierarchy(pf.levels, is_per, sym)
        for l in range(len(pf.levels)):
            full = H1.fb_source_map(l, 1, cross=True)
            mine = H.fb_source_map(l, 1, cross=True)
            for i, b in enumerate(H.local_boxes[l]):
                f, m = full[b], mine[i]
                src_box = np.where(f >= 0, f >> 40, -1)
                remote = (src_box >= 0) & (H.owners[l][np.maximum(src_box, 0)] != rank)
                if flags & capi.PEER_LINKS:
                    # peer-linked faces name the peer's box directly (read in place); only unlinked remote cells use the slab
                    lk = H.links(l, b)
                    g = 1
                    linked = np.zeros(f.shape, dtype=bool)
                    for face in range(6):
                        if lk[face][0] >= 0:
                            sl = [slice(g, -g)] * 3
                            sl[2 - face % 3] = 0 if face < 3 else -1
                            linked[tuple(sl)] = True
                    assert np.array_equal(m == -2, remote & ~linked)
                    assert np.array_equal(m[m != -2], f[m != -2])
                    assert (H.owners[l][lk[lk[:, 0] >= 0, 0]] == lk[lk[:, 0] >= 0, 1]).all()
                else:
                    assert np.array_equal(m == -2, remote)
                    assert np.array_equal(m[~remote], f[~remote])


def _spawn2(cases, flags):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0, 0])
    procs = [ctx.Process(target=_worker, args=(r, 2, port, list(cases), ok, flags)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert list(ok) == [1, 1], [p.exitcode for p in procs]


def test_exchange_plan_world2_peer_links(palib):
    """PA_HIER_PEER_LINKS: faces linked to a peer's box leave the exchange plan; what remains is still consistent."""
    _spawn2(["c1_periodic", "c3_three_levels", "edge_walls", "mixed_boxes"], 1)


def test_uniform_grid_with_peer_links_needs_no_exchange(palib):
    from peleanalysis_b200 import capi, synth
    pf = synth.make_hierarchy(32, [], [], 16, fill=False)
    for rank in range(2):
        H = capi.Hierarchy(pf.levels, (1, 1, 1), (0, 0, 0), rank, 2, flags=capi.PEER_LINKS)
        sc, rc = H.exchange_prefix(1)
        assert sc.sum() == 0 and rc.sum() == 0
        remote = 0
        for b in range(len(pf.levels[0].boxes)):
            lk = H.links(0, b)
            assert (lk[:, 0] >= 0).all()
            remote += int((lk[:, 1] != H.owners[0][b]).sum())
        assert remote > 0
        H0 = capi.Hierarchy(pf.levels, (1, 1, 1), (0, 0, 0), rank, 2)
        sc, rc = H0.exchange_prefix(1)
        assert sc.sum() > 0 and rc.sum() > 0


def test_exchange_plan_world2(palib):
    _spawn2(["c1_periodic", "c3_three_levels", "lshape", "edge_walls", "ratio4", "mixed_boxes"], 0)
