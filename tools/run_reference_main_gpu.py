#!/usr/bin/env python
"""GPU BOX: the reference's UNMODIFIED main.py (baseline/_ref/main.py, staged by tools/stage_reference.py) driving this repo's
drop-in packages on a B200 - `from models import *` (main.py:12) resolves to dropin/models -> elimrec_b200.model.EliMRec, the
evaluator backend to dropin/evaluator; Configurator, Dataset, PairwiseSamplerV2, Logger, Meter, the epoch loop, torch's Adam and
the early-stopping logic are the reference's own.  Synthetic files in the reference's on-disk format.

    python tools/run_reference_main_gpu.py [shape=tiktok] [epochs=3] > profiles/<log>
"""
import collections
import collections.abc
import importlib.machinery
import os
import sys
import time
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "tiktok"
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    if not os.path.isdir(REF):
        sys.exit("baseline/_ref is missing: run tools/stage_reference.py in the build container first")
    sys.path.insert(0, REPO)
    import torch
    from elimrec_b200 import synth
    assert torch.cuda.is_available()
    # the three import shims of SURVEY.md section 8c (no reference file is edited)
    for name in ("tensorflow", "torch_scatter"):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m

    def scatter(src, index, dim=0, reduce="mean"):
        raise NotImplementedError("torch_scatter is only reached on the literal 'tiktok' branch")
    sys.modules["torch_scatter"].scatter = scatter
    collections.Iterable = collections.abc.Iterable
    name = "synthshape"
    data_dir = f"/tmp/elimrec_main_{shape}"
    t0 = time.time()
    inter, feats = synth.make_shape(shape)
    synth.write_reference_files(data_dir, name, inter, feats)
    print(f"[driver] {shape}-shape files written to {data_dir} ({time.time() - t0:.1f}s)", flush=True)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REPO, "dropin"))          # ahead of the reference: models / evaluator are the drop-ins
    os.chdir(REF)
    sys.argv = ["main.py", f"--data.input.path={data_dir}", f"--data.input.dataset={name}", "--loss=bpr_loss", "--alpha=0.5",
                "--topks=[20]", f"--num_epoch={epochs}", "--test_step=1", "--save_flag=False", "--create_log_file=False", "--verbose=1"]
    import main as ref_main                       # the reference's main.py, unmodified
    import tqdm
    ref_main.tqdm = tqdm.tqdm                     # main.py:92 uses tqdm without importing it
    from util import set_seed
    from util.configurator import Configurator
    args = Configurator("./NeuRec.properties", default_section="hyperparameters")
    set_seed(args["seed"])
    t0 = time.time()
    net = ref_main.Net(args)
    cls = type(net.recommender)
    print(f"[driver] recommender class: {cls.__module__}.{cls.__name__} on {next(net.recommender.parameters()).device} "
          f"(schedule: {'linear' if net.recommender.linear else 'slab'}, proj_precision={net.recommender.proj_precision}); "
          f"evaluator backend: {type(net.recommender.test_evaluator.evaluator).__module__}; built in {time.time() - t0:.1f}s", flush=True)
    t0 = time.time()
    net.run()
    torch.cuda.synchronize()
    from elimrec_b200._lib import CALLS
    print(f"[driver] main.py Net.run(): {epochs} epochs in {time.time() - t0:.1f}s; C-ABI calls {CALLS['n']}, kernel launches {CALLS['launches']}")


if __name__ == "__main__":
    main()
