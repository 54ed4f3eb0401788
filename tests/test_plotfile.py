"""CPU: plotfile reader/writer round trip and header layout."""
import numpy as np

from peleanalysis_b200 import plotfile, synth


def test_roundtrip(tmp_path):
    pf = synth.config1(16, 8, names=("temp", "Y_CH4"))
    d = str(tmp_path / "plt")
    plotfile.write_plotfile(d, pf, nfiles_per_level=3)
    r = plotfile.read_plotfile(d)
    assert r.names == ["temp", "Y_CH4"] and r.ref_ratio == [2]
    for a, b in zip(pf.levels, r.levels):
        assert a.boxes == b.boxes and a.dx == b.dx
        for fa, fb in zip(a.fabs, b.fabs):
            assert np.array_equal(fa, fb)
    only = plotfile.read_plotfile(d, comps=["Y_CH4"], finest_level=0)
    assert len(only.levels) == 1 and np.array_equal(only.levels[0].fabs[3][0], pf.levels[0].fabs[3][1])
    lo, hi = plotfile.file_min_max(d, "temp", 2)
    assert lo == min(f[0].min() for l in pf.levels for f in l.fabs)
    assert hi == max(f[0].max() for l in pf.levels for f in l.fabs)


def test_existing_directory_is_renamed_like_the_reference(tmp_path):
    pf = synth.config1(16, 8)
    d = str(tmp_path / "plt")
    plotfile.write_plotfile(d, pf)
    plotfile.write_plotfile(d, pf)
    olds = [p.name for p in tmp_path.iterdir() if p.name.startswith("plt.old.")]
    assert len(olds) == 1
