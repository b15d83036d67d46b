"""The class-level drop-in boundary, end to end on the host: the reference's unmodified ``main.py`` loop
(``Net(args).run()``: Configurator, Dataset, PairwiseSamplerV2, torch Adam, ``loss.backward(retain_graph=True)``,
``evaluate()`` / ``test()`` every epoch) is run twice on the same tiny dataset - once with the reference's own ``models``
package and once with ``dropin/`` shadowing it - and must train to the same weights and log the same metrics.

Needs /root/reference (build container only); the kernels are stood in for by tests/sim_ops.py, see the driver."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout")


def test_reference_main_loop_drives_the_dropin(tmp_path):
    sys.path.insert(0, os.path.dirname(HERE))
    from elimrec_b200 import synth
    data_dir = str(tmp_path / "data")
    inter = synth.make_interactions(40, 70, 600, seed=7)
    synth.write_reference_files(data_dir, "synthg", inter, synth.make_features(70, (16, 8, 24), seed=7))
    outs = {}
    for which in ("reference", "dropin"):
        outs[which] = str(tmp_path / f"{which}.npz")
        r = subprocess.run([sys.executable, os.path.join(HERE, "dropin_main_driver.py"), which, data_dir, outs[which]],
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-3000:]
    ref, got = np.load(outs["reference"]), np.load(outs["dropin"])
    assert str(ref["class"]) == "models.EliMRec.EliMRec" and str(got["class"]) == "elimrec_b200.model.EliMRec"
    keys = [k for k in ref.files if k.startswith("sd/")]
    assert keys == [k for k in got.files if k.startswith("sd/")]
    for k in keys:      # 3 epochs x 4 Adam steps from the same init, same triples
        d = np.abs(got[k].astype(np.float64) - ref[k]) / np.abs(ref[k]).max()
        assert d.max() < 2e-4 and (d > 1e-5).mean() < 2e-3, (k, d.max())
    # the logged validation / test metrics (Recall / NDCG / Precision @20 of every epoch, TE and TIE) agree line by line
    import re
    num = re.compile(r"(?<![@\w])\d+\.\d+(?:e-?\d+)?")
    rl, gl = str(ref["log"]).split("\n"), str(got["log"]).split("\n")
    assert len(rl) == len(gl) >= 7
    for a, b in zip(rl, gl):
        assert num.sub("#", a) == num.sub("#", b)
        np.testing.assert_allclose([float(x) for x in num.findall(b)], [float(x) for x in num.findall(a)], atol=2e-6)
