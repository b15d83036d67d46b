#!/usr/bin/env python
"""Generate golden vectors by RUNNING THE UNMODIFIED REFERENCE (CPU) on tiny synthetic inputs.

Only runs in the build container (needs /root/reference); its outputs ``tests/golden/*.npz`` are
committed and are what travels to the GPU box.  Nothing under tests/ or bench.py imports this.

What it does:
  1. copies /root/reference to /tmp/ref_build and rebuilds its Cython (setup.py build_ext --inplace);
  2. installs the import shims (tensorflow stub with a ModuleSpec, collections.Iterable alias,
     a torch_scatter stub) - no reference file is edited;
  3. writes a tiny dataset in the reference's on-disk format (elimrec_b200.synth), builds the
     reference's Configurator / Dataset / EliMRec / PairwiseSamplerV2 / ProxyEvaluator on CPU;
  4. records: id remap + CSR splits, norm_adj COO, initial state_dict, one epoch of libc-rand
     triples (srand(1)), DataIterator batches (np.random.seed(2022)), loss + every gradient of one
     bpr_loss step, cached tables, predict() scores for TIE/TE/normal, evaluate()/test() results,
     and the parameters after 3 Adam steps.

Usage:  python tests/golden/make_golden.py            (writes tests/golden/{generic,kwai}.npz)
"""
import collections
import collections.abc
import ctypes
import importlib.machinery
import os
import shutil
import subprocess
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference"
REF = "/tmp/ref_build"
sys.path.insert(0, REPO)


def prepare_reference():
    if not os.path.exists(os.path.join(REF, "util/cython")) or not any(
            f.endswith("cpython-312-x86_64-linux-gnu.so") for f in os.listdir(os.path.join(REF, "util/cython"))):
        shutil.rmtree(REF, ignore_errors=True)
        shutil.copytree(REF_SRC, REF, ignore=shutil.ignore_patterns(".ipynb_checkpoints"))
        subprocess.run(["chmod", "-R", "u+w", REF], check=True)
        subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=REF, check=True,
                       stdout=subprocess.DEVNULL)


def install_shims():
    for name in ("tensorflow", "torch_scatter"):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m

    def scatter(src, index, dim=0, reduce="mean"):
        raise NotImplementedError("torch_scatter is only reached on the literal 'tiktok' branch")
    sys.modules["torch_scatter"].scatter = scatter
    collections.Iterable = collections.abc.Iterable


def run(name, shape, out_file, alpha=0.5):
    import torch
    from elimrec_b200 import synth

    data_dir = f"/tmp/golden_data_{name}"
    shutil.rmtree(data_dir, ignore_errors=True)
    U, I, n, dims = shape
    inter = synth.make_interactions(U, I, n, seed=7)
    feats = synth.make_features(I, dims, seed=7)
    synth.write_reference_files(data_dir, name, inter, feats)

    os.chdir(REF)
    sys.argv = ["main.py", f"--data.input.path={data_dir}", f"--data.input.dataset={name}",
                "--loss=bpr_loss", f"--alpha={alpha}", "--topks=[20]", "--batch_size=128",
                "--test_batch_size=16", "--no_cuda=TRUE", "--verbose=0"]
    from util.configurator import Configurator
    from util import set_seed
    from util.logger import Logger
    from data.dataset import Dataset
    from data import PairwiseSamplerV2
    from data.sampler import _pairwise_sampling_v2
    from models import EliMRec

    conf = Configurator("./NeuRec.properties", default_section="hyperparameters")
    set_seed(conf["seed"])
    Logger.logger = Logger(name="golden", show_in_console=False, is_creat_log_file=False, path="./log")
    conf.device = torch.device("cpu")
    ds = Dataset(conf)
    model = EliMRec(conf, ds)
    out = {}
    out["raw_train"], out["raw_valid"], out["raw_test"] = inter.train, inter.valid, inter.test
    for k, f in zip("vat", feats):
        if f is not None:
            out[f"raw_feat_{k}"] = f
    out["num_users"], out["num_items"] = np.int64(ds.num_users), np.int64(ds.num_items)
    for split in ("train", "valid", "test"):
        m = getattr(ds, f"{split}_matrix")
        out[f"{split}_indptr"], out[f"{split}_indices"] = m.indptr.astype(np.int64), m.indices.astype(np.int64)
    out["itemids_raw_in_order"] = np.array(list(ds.itemids.keys()), dtype=np.int64)
    out["userids_raw_in_order"] = np.array(list(ds.userids.keys()), dtype=np.int64)
    adj = model.norm_adj  # uncoalesced COO, row-major
    out["adj_row"], out["adj_col"] = adj._indices()[0].numpy(), adj._indices()[1].numpy()
    out["adj_val"] = adj._values().numpy()
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k, v in sd0.items():
        out["sd0/" + k] = v.numpy()

    # --- sampler: libc stream from its default state (srand(1)), one epoch -------------------
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    sampler = PairwiseSamplerV2(ds, neg_num=1, batch_size=conf.batch_size, shuffle=True)
    u, p, ng = _pairwise_sampling_v2(sampler.user_pos_dict, sampler.num_trainings, sampler.item_num)
    out["epoch_users"], out["epoch_pos"], out["epoch_neg"] = (np.asarray(u, dtype=np.int64),
                                                              np.asarray(p, dtype=np.int64),
                                                              np.asarray(ng, dtype=np.int64))
    # second epoch continues the stream
    u2, p2, n2 = _pairwise_sampling_v2(sampler.user_pos_dict, sampler.num_trainings, sampler.item_num)
    out["epoch2_users"], out["epoch2_pos"], out["epoch2_neg"] = (np.asarray(u2, dtype=np.int64),
                                                                 np.asarray(p2, dtype=np.int64),
                                                                 np.asarray(n2, dtype=np.int64))
    # --- full iterator: same libc stream + numpy permutation --------------------------------
    libc.srand(1)
    np.random.seed(conf["seed"])
    batches = [(np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64), np.asarray(c, dtype=np.int64))
               for a, b, c in sampler]
    out["n_batches"] = np.int64(len(batches))
    for i, (a, b, c) in enumerate(batches[:4]):
        out[f"batch{i}_users"], out[f"batch{i}_pos"], out[f"batch{i}_neg"] = a, b, c

    # --- one loss step: loss, grads, cached tables ----------------------------------------------
    opt = torch.optim.Adam(model.parameters(), lr=conf.lr, weight_decay=conf.weight_decay)
    model.train()
    tb = [torch.tensor(x) for x in batches[0]]
    loss = model.bpr_loss(*tb)
    opt.zero_grad()
    loss.backward(retain_graph=True)
    out["loss0"] = loss.detach().numpy()
    for k, prm in model.named_parameters():
        if prm.grad is not None:  # kwai: s_dense_a / s_dense_t are never used (EliMRec.py:88-90)
            out["grad0/" + k] = prm.grad.detach().numpy().copy()
    out["all_users"], out["all_items"] = model.all_users.detach().numpy(), model.all_items.detach().numpy()
    mods = "v" if name == "kwai" else "vat"
    for m in mods:
        out[f"s_user_{m}"] = model.all_s_embs[f"pre_fusion_user_{m}"].detach().numpy()
        out[f"s_item_{m}"] = model.all_s_embs[f"pre_fusion_item_{m}"].detach().numpy()
    out["light_i"] = model.i_emb.detach().numpy()
    out["light_v"] = model.v_emb.detach().numpy()
    if name != "kwai":
        out["light_a"], out["light_t"] = model.a_emb.detach().numpy(), model.t_emb.detach().numpy()

    # --- predict / evaluate with the tables cached by that forward ------------------------------
    model.eval()
    some_users = list(ds.get_user_valid_dict().keys())[:16]
    out["predict_users"] = np.asarray(some_users, dtype=np.int64)
    for pt in ("TIE", "TE", "normal"):
        model.predict_type = pt
        out[f"predict_{pt}"] = model.predict(some_users, None).numpy()
        res, buf = model.evaluate()
        out[f"evaluate_{pt}"] = np.asarray(res)
        res, buf = model.test()
        out[f"test_{pt}"] = np.asarray(res)
    # per-user metric rows of the C++ evaluator for the valid split (TIE), first 16 users
    model.predict_type = "TIE"
    ev = model.valid_evaluator.evaluator
    scores = np.array(model.predict(some_users, None), dtype=np.float32)
    for idx, uu in enumerate(some_users):
        scores[idx][ev.user_pos_train.get(uu, [])] = -np.inf
    out["masked_scores_TIE"] = scores.copy()
    test_items = [ev.user_pos_test[uu] for uu in some_users]
    out["metric_rows_TIE"] = ev.eval_score_matrix(scores, test_items, ev.metrics, top_k=ev.max_top,
                                                  thread_num=ev.num_thread)

    # --- 3 Adam steps (main.py:92-102) -------------------------------------------------------------
    model.train()
    opt.step()
    losses = [float(loss)]
    for b in batches[1:3]:
        tb = [torch.tensor(x) for x in b]
        loss = model.bpr_loss(*tb)
        opt.zero_grad()
        loss.backward(retain_graph=True)
        opt.step()
        losses.append(float(loss))
    out["losses"] = np.asarray(losses, dtype=np.float64)
    for k, v in model.state_dict().items():
        out["sd3/" + k] = v.detach().numpy()
    np.savez_compressed(out_file, **out)
    print(name, "->", out_file, f"{os.path.getsize(out_file) / 1e3:.0f} kB; U={ds.num_users} I={ds.num_items} "
          f"E_train={ds.train_matrix.nnz} batches={len(batches)} losses={losses}")


if __name__ == "__main__":
    prepare_reference()
    install_shims()
    sys.path.insert(0, REF)
    which = sys.argv[1] if len(sys.argv) > 1 else "generic"
    if which == "generic":
        run("synthg", (40, 70, 600, (16, 8, 24)), os.path.join(HERE, "generic.npz"))
    else:
        run("kwai", (30, 60, 500, (32, 0, 0)), os.path.join(HERE, "kwai.npz"))
