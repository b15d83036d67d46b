#!/usr/bin/env python
"""Golden vectors for the SURVEY.md section 8(f) "next" rows, again by RUNNING THE UNMODIFIED REFERENCE on CPU.

Same recipe as ``make_golden.py`` (copy + Cython rebuild under /tmp/ref_build, import shims, no reference file
edited), same tiny dataset (40 users x 70 items, seed 7), non-default knobs:

  f2  ``--adj_type=plain|norm|gcmc|mean``      adjacency COO, loss, every gradient, parameters after 3 Adam steps
  f4  ``--mm_fusion_mode=mean``                loss, gradients, tables, TIE scores, parameters after 3 steps
  f4  ``--s_fusion_mode=hm|sum``               predict() scores for TIE / TE, evaluate() results
  f3  candidate-negatives ``ProxyEvaluator`` + all five metrics at top_k=[5, 20]; ``arg_topk``
  f1  the literal ``tiktok`` dataset branch (word-id text feature, ``word_embedding`` parameter)

Only runs in the build container; writes ``tests/golden/next.npz`` and ``tests/golden/tiktok.npz`` (committed).

Usage:  python tests/golden/make_golden_next.py [next|tiktok]
"""
import ctypes
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (prepare_reference / install_shims / REF)

SHAPE = (40, 70, 600, (16, 8, 24))


def real_scatter_mean():
    """torch_scatter.scatter(src, index, dim=0, reduce='mean') restated with index_add_ (un-vendored dependency,
    readme.md:18; call site models/EliMRec.py:377-378)."""
    import torch

    def scatter(src, index, dim=0, reduce="mean"):
        assert dim == 0 and reduce == "mean"
        n = int(index.max()) + 1
        out = torch.zeros(n, src.shape[1], dtype=src.dtype).index_add(0, index, src)
        cnt = torch.zeros(n, dtype=src.dtype).index_add(0, index, torch.ones_like(index, dtype=src.dtype))
        return out / cnt.clamp(min=1).unsqueeze(1)
    sys.modules["torch_scatter"].scatter = scatter


def build(name, data_dir, extra):
    import torch
    os.chdir(mg.REF)
    sys.argv = ["main.py", f"--data.input.path={data_dir}", f"--data.input.dataset={name}", "--loss=bpr_loss",
                "--alpha=0.5", "--topks=[20]", "--batch_size=128", "--test_batch_size=16", "--no_cuda=TRUE",
                "--verbose=0"] + list(extra)
    from util.configurator import Configurator
    from util import set_seed
    from util.logger import Logger
    from data.dataset import Dataset
    from data import PairwiseSamplerV2
    from models import EliMRec
    conf = Configurator("./NeuRec.properties", default_section="hyperparameters")
    set_seed(conf["seed"])
    Logger.logger = Logger(name="golden", show_in_console=False, is_creat_log_file=False, path="./log")
    conf.device = torch.device("cpu")
    ds = Dataset(conf)
    model = EliMRec(conf, ds)
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    np.random.seed(conf["seed"])
    sampler = PairwiseSamplerV2(ds, neg_num=1, batch_size=conf.batch_size, shuffle=True)
    batches = [tuple(np.asarray(x, dtype=np.int64) for x in b) for b in sampler][:3]
    return conf, ds, model, batches


def train3(model, conf, batches, out, pre, keep_tables=False, predict_types=(), light=False):
    import torch
    for k, v in model.state_dict().items():
        out[f"{pre}/sd0/{k}"] = v.detach().numpy().copy()
    opt = torch.optim.Adam(model.parameters(), lr=conf.lr, weight_decay=conf.weight_decay)
    model.train()
    losses = []
    for s, b in enumerate(batches):
        loss = model.bpr_loss(*[torch.tensor(x) for x in b])
        opt.zero_grad()
        loss.backward(retain_graph=True)
        if s == 0:
            out[f"{pre}/loss0"] = loss.detach().numpy()
            for k, prm in model.named_parameters():
                if prm.grad is not None and not light:
                    out[f"{pre}/grad0/{k}"] = prm.grad.detach().numpy().copy()
            if keep_tables:
                out[f"{pre}/all_users"] = model.all_users.detach().numpy()
                out[f"{pre}/all_items"] = model.all_items.detach().numpy()
            users = list(model.dataset.get_user_valid_dict().keys())[:16]
            out[f"{pre}/predict_users"] = np.asarray(users, dtype=np.int64)
            model.eval()
            for pt in predict_types:
                model.predict_type = pt
                out[f"{pre}/predict_{pt}"] = model.predict(users, None).numpy()
                out[f"{pre}/evaluate_{pt}"] = np.asarray(model.evaluate()[0])
            model.predict_type = "TIE"
            model.train()
        opt.step()
        losses.append(float(loss))
    out[f"{pre}/losses"] = np.asarray(losses, dtype=np.float64)
    for k, v in model.state_dict().items():
        if not light:
            out[f"{pre}/sd3/{k}"] = v.detach().numpy().copy()
    for i, b in enumerate(batches):
        for nm, x in zip(("users", "pos", "neg"), b):
            out[f"{pre}/batch{i}_{nm}"] = x


def run_next(out_file):
    from elimrec_b200 import synth
    data_dir = "/tmp/golden_data_next"
    shutil.rmtree(data_dir, ignore_errors=True)
    U, I, n, dims = SHAPE
    inter = synth.make_interactions(U, I, n, seed=7)
    feats = synth.make_features(I, dims, seed=7)
    synth.write_reference_files(data_dir, "synthg", inter, feats)
    out = {}
    # ---- f2: adjacency types ----------------------------------------------------------------------------------
    for adj in ("plain", "norm", "gcmc", "mean"):
        conf, ds, model, batches = build("synthg", data_dir, [f"--adj_type={adj}"])
        a = model.norm_adj
        out[f"adj_{adj}/row"], out[f"adj_{adj}/col"] = a._indices()[0].numpy(), a._indices()[1].numpy()
        out[f"adj_{adj}/val"] = a._values().numpy()
        train3(model, conf, batches, out, f"adj_{adj}", predict_types=("TIE",))
        print("adj_type", adj, out[f"adj_{adj}/losses"])
    # ---- f4: mean fusion of the modality blocks -------------------------------------------------------------------
    conf, ds, model, batches = build("synthg", data_dir, ["--mm_fusion_mode=mean"])
    train3(model, conf, batches, out, "mm_mean", keep_tables=True, predict_types=("TIE", "TE"))
    print("mm_fusion_mode=mean", out["mm_mean/losses"])
    # ---- f4: score fusion modes -----------------------------------------------------------------------------------
    for fm in ("hm", "sum"):
        conf, ds, model, batches = build("synthg", data_dir, [f"--s_fusion_mode={fm}"])
        train3(model, conf, batches[:1], out, f"s_{fm}", predict_types=("TIE", "TE", "normal"), light=True)
        print("s_fusion_mode", fm, out[f"s_{fm}/evaluate_TIE"])
    # ---- f4: modality ablation under hm (hm / sum ignore `modality` in the score, not in the loss) -----------------
    conf, ds, model, batches = build("synthg", data_dir, ["--s_fusion_mode=hm", "--modality=va"])
    train3(model, conf, batches[:1], out, "s_hm_va", predict_types=("TIE",), light=True)
    # ---- f3: candidate negatives, all five metrics, two cut-offs; standalone arg_topk -----------------------------
    import torch
    from evaluator import ProxyEvaluator
    from util.cython.arg_topk import arg_topk
    conf, ds, model, batches = build("synthg", data_dir, [])
    model.bpr_loss(*[torch.tensor(x) for x in batches[0]])
    model.eval()
    model.predict_type = "TIE"
    test = ds.get_user_test_dict()
    rng = np.random.default_rng(3)
    neg = {u: rng.choice(ds.num_items, size=7, replace=False).tolist() for u in test}
    allm = ["Precision", "Recall", "MAP", "NDCG", "MRR"]
    ev = ProxyEvaluator(ds, ds.get_user_train_dict(), test, neg, metric=allm, group_view=None, top_k=[5, 20],
                        batch_size=16, num_thread=4)
    out["cand/result"] = np.asarray(ev.evaluate(model)[0])
    out["cand/users"] = np.asarray(list(test.keys()), dtype=np.int64)
    out["cand/neg"] = np.asarray([neg[u] for u in test], dtype=np.int64)
    ev2 = ProxyEvaluator(ds, ds.get_user_train_dict(), test, None, metric=allm, group_view=None, top_k=[5, 20],
                         batch_size=16, num_thread=4)
    out["allmetrics/result"] = np.asarray(ev2.evaluate(model)[0])
    ev3 = ProxyEvaluator(ds, ds.get_user_train_dict(), test, None, metric=["MAP", "MRR"], group_view=None, top_k=7,
                         batch_size=16, num_thread=4)
    out["topk_int/result"] = np.asarray(ev3.evaluate(model)[0])
    out["topk_int/info"] = np.asarray(ev3.metrics_info())
    users = list(test.keys())[:16]
    sc = np.array(model.predict(users, None), dtype=np.float32)
    out["arg_topk/scores"] = sc
    out["arg_topk/idx"] = np.asarray(arg_topk(sc, 10, 4))
    for k, v in model.state_dict().items():
        out[f"cand/sd0/{k}"] = v.detach().numpy().copy()
    for nm, x in zip(("users", "pos", "neg"), batches[0]):
        out[f"cand/batch0_{nm}"] = x
    np.savez_compressed(out_file, **out)
    print("->", out_file, f"{os.path.getsize(out_file) / 1e3:.0f} kB")


def run_tiktok(out_file):
    """dataset name literally 'tiktok': <path>/tiktok_{visual,audio,textual}_feat.pt (data/dataset.py:164-174)."""
    import torch
    from elimrec_b200 import synth
    data_dir = "/tmp/golden_data_tiktok"
    shutil.rmtree(data_dir, ignore_errors=True)
    U, I, n, dims = SHAPE
    inter = synth.make_interactions(U, I, n, seed=7)
    v, a, _ = synth.make_features(I, dims, seed=7)
    words = synth.make_words(I, seed=7)
    synth.write_tiktok_files(data_dir, inter, v, a, words)
    out = {"raw_words": words, "raw_feat_v": v, "raw_feat_a": a}
    out["raw_train"], out["raw_valid"], out["raw_test"] = inter.train, inter.valid, inter.test
    conf, ds, model, batches = build("tiktok", data_dir, [])
    out["words_tensor"] = ds.words_tensor.numpy()
    out["t_feat"] = model.t_feat.detach().numpy()
    out["num_users"], out["num_items"] = np.int64(ds.num_users), np.int64(ds.num_items)
    train3(model, conf, batches, out, "tiktok", keep_tables=True, predict_types=("TIE",))
    # word_embedding is [11574 x 128]: keep only the rows some item uses plus 16 unused ones (their gradient is zero and
    # their update is pure weight decay), so that the fixture stays small
    used = np.unique(out["words_tensor"][1])
    rest = np.setdiff1d(np.arange(11574), used)[:: max(1, (11574 - used.size) // 16)][:16]
    out["word_rows"] = np.concatenate([used, rest]).astype(np.int64)
    out["n_word_rows_used"] = np.int64(used.size)
    assert not out["tiktok/grad0/word_embedding.weight"][rest].any()
    for k in [k for k in out if k.endswith("word_embedding.weight")]:
        out[k] = out[k][out["word_rows"]]
    print("tiktok", out["tiktok/losses"], [k for k in out if k.startswith("tiktok/grad0")])
    np.savez_compressed(out_file, **out)
    print("->", out_file, f"{os.path.getsize(out_file) / 1e3:.0f} kB")


if __name__ == "__main__":
    mg.prepare_reference()
    mg.install_shims()
    real_scatter_mean()
    sys.path.insert(0, mg.REF)
    which = sys.argv[1] if len(sys.argv) > 1 else "next"
    if which == "next":
        run_next(os.path.join(HERE, "next.npz"))
    else:
        run_tiktok(os.path.join(HERE, "tiktok.npz"))
