"""GPU helper: curvature on a full-size hierarchy through the fused and the separate kernels; the two routes must give the
same output fingerprint (pa_field_hash).  Usage: python scripts/gpu_hash_check.py [base mgs]  (default 512 128)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from peleanalysis_b200 import capi, synth  # noqa: E402

base, mgs = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 128)
capi.init(0)
pf = synth.config3(base, mgs, fill=False)
H = capi.Hierarchy(pf.levels)
host = bench.gen_fields(torch, pf.levels, H.local_boxes, ["temp"])
fin = capi.Field(H, 1, 1)
for l in range(H.nlev):
    capi.check(capi.lib().pa_field_upload_level(fin.f, l, 0, host[l][0].data_ptr()))
o = capi.CurvOpts()
o.prog_min = min(float(h[0].min()) for h in host)
o.prog_max = max(float(h[0].max()) for h in host)
res = {}
for thr in (0, 1):
    o.do_threshold, o.threshold = thr, 0.05
    for fused in ("nw", "n3", "3", "1", "0"):
        os.environ["PA_CURV_FUSED"] = "0" if fused in ("n3", "nw") else fused
        os.environ["PA_NORMAL_F3"] = "1" if fused == "n3" else "0"
        os.environ["PA_NORMAL_W"] = "1" if fused == "nw" else "0"
        out = capi.Field(H, 5, 1)
        out.set_val(-3.0)
        f0 = capi.curv_fused_launches()
        capi.curvature(fin, 0, 0, o, out, 0)
        capi.sync()
        res[(thr, fused)] = [out.hash(c, 1) for c in range(5)]
        print("threshold", thr, "fused", fused, "fused launches", capi.curv_fused_launches() - f0, ["%016x" % h for h in res[(thr, fused)]], flush=True)
        del out
    assert res[(thr, "1")] == res[(thr, "0")], "fused and unfused curvature differ (threshold %d)" % thr
    assert res[(thr, "3")] == res[(thr, "0")], "third fused kernel and unfused curvature differ (threshold %d)" % thr
    assert res[(thr, "n3")] == res[(thr, "0")], "plane-staged flame-normal kernel and MODE_NORMAL_S differ (threshold %d)" % thr
    assert res[(thr, "nw")] == res[(thr, "0")], "barrier-free flame-normal kernel and MODE_NORMAL_S differ (threshold %d)" % thr
print("HASH CHECK OK: fused == unfused on config3(%d, %d), with and without threshold_prog" % (base, mgs))
