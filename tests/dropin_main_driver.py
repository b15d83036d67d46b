"""Runs the reference's UNMODIFIED ``main.py`` training loop (``Net(args).run()``, main.py:27-155) in this process.

    python tests/dropin_main_driver.py reference|dropin <data_dir> <out.npz>

``reference``: the reference's own ``models`` / ``evaluator`` packages (CPU torch).
``dropin``   : ``dropin/`` is put ahead of the reference checkout on ``sys.path``, so ``from models import *`` (main.py:12)
               resolves to ``elimrec_b200.model.EliMRec``; the reference's Configurator, Dataset, sampler, Logger, Meter and
               the loop itself are untouched.  There is no GPU in the build container, so the C-ABI wrappers are replaced by
               the torch-CPU schedule simulator (tests/sim_ops.py) - what is exercised is the class-level drop-in boundary.
Only runs where /root/reference exists (tests/test_dropin_main.py skips otherwise)."""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
import make_golden as mg  # noqa: E402


def main(which, data_dir, out_file):
    import torch
    mg.prepare_reference()
    mg.install_shims()
    sys.path.insert(0, mg.REF)
    if which == "dropin":
        sys.path.insert(0, os.path.join(REPO, "dropin"))
        import sim_ops
        import elimrec_b200.evaluator as ev
        import elimrec_b200.linear as ln
        import elimrec_b200.model as md
        import elimrec_b200.optim as op
        for mod in (md, ev, op, ln):
            mod.ops = sim_ops
        md._require_cuda = lambda dev: None

        class _Ev:
            def __init__(self, *a, **k):
                pass

            def record(self, *a):
                pass

        class _St:
            def wait_event(self, *a):
                pass
        torch.cuda.Event = _Ev
        torch.cuda.current_stream = lambda *a: _St()
    os.chdir(mg.REF)
    sys.argv = ["main.py", f"--data.input.path={data_dir}", "--data.input.dataset=synthg", "--loss=bpr_loss", "--alpha=0.5",
                "--topks=[20]", "--batch_size=128", "--test_batch_size=16", "--no_cuda=TRUE", "--verbose=0", "--num_epoch=3",
                "--test_step=1", "--save_flag=False", "--create_log_file=False", "--proj_precision=fp32", "--rank_backend=fp32"]
    import ctypes
    import main as ref_main                       # the reference's main.py, unmodified
    import tqdm
    ref_main.tqdm = tqdm.tqdm                     # main.py:92 uses tqdm without importing it
    from util import set_seed
    from util.configurator import Configurator
    from util.logger import Logger
    args = Configurator("./NeuRec.properties", default_section="hyperparameters")
    set_seed(args["seed"])
    ctypes.CDLL("libc.so.6").srand(1)
    net = ref_main.Net(args)
    lines = []
    info = Logger.info
    Logger.info = staticmethod(lambda msg: (lines.append(str(msg)), info(msg))[1])
    net.run()
    cls = type(net.recommender)
    out = {"class": np.asarray(f"{cls.__module__}.{cls.__name__}"),
           "log": np.asarray(re.sub(r" time:[0-9.e+-]+", "", "\n".join(l for l in lines if l.startswith("[TIE]") or
                                                                        l.startswith("  [T"))))}
    for k, v in net.recommender.state_dict().items():
        out["sd/" + k] = v.detach().cpu().numpy()
    np.savez_compressed(out_file, **out)
    print(out["class"], "\n", out["log"])


if __name__ == "__main__":
    main(*sys.argv[1:4])
