"""Synthetic datasets of the shapes BASELINE.json names.

Generator contract (seed 2022, ``numpy.random.default_rng``): users uniform, items Zipf-like
(exponent 0.8), unique (user, item) pairs, per-user 80/10/10 train/valid/test split, every user
id and item id occurs at least once and every user has at least one train item.

Two consumers:

* :func:`make_interactions` / :func:`make_features` return in-memory arrays (bench, tests);
* :func:`write_reference_files` writes them in the exact on-disk format the reference's
  ``Dataset`` reads (``data/dataset.py:207-212`` CSV splits, ``:181-185`` generic ``.npy``
  feature branch, ``:178-180`` ``kwai_feat_v.pt``), so the same inputs can be fed to the
  reference in this container when generating golden vectors.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

# name -> (users, items, interactions, (Dv, Da, Dt))   Da = Dt = 0 means the v-only (kwai) model
SHAPES = {
    "tiktok": (36_656, 76_085, 726_000, (128, 128, 768)),
    "kwai": (7_010, 86_483, 1_300_000, (2048, 0, 0)),
    "movielens": (55_485, 5_986, 1_200_000, (2048, 128, 100)),
    "tiktok10x": (366_560, 760_850, 7_260_000, (128, 128, 768)),
    # small shapes for tests / smoke
    "tiny": (40, 70, 600, (16, 8, 24)),
    "small": (600, 900, 12_000, (32, 16, 48)),
    "medium": (4_000, 9_000, 90_000, (128, 128, 256)),
}


@dataclass
class Interactions:
    num_users: int
    num_items: int
    train: np.ndarray  # [E_train, 2] int64 (user, item)
    valid: np.ndarray
    test: np.ndarray


def make_interactions(num_users: int, num_items: int, num_inter: int, seed: int = 2022,
                      zipf: float = 0.8) -> Interactions:
    rng = np.random.default_rng(seed)
    U, I = int(num_users), int(num_items)
    num_inter = max(int(num_inter), U + I)
    # item popularity: Zipf-like over a random permutation of item ids
    p = (np.arange(1, I + 1, dtype=np.float64)) ** (-zipf)
    p /= p.sum()
    cdf = np.cumsum(p)
    item_of_rank = rng.permutation(I)

    def draw_items(n):
        r = np.searchsorted(cdf, rng.random(n), side="right")
        return item_of_rank[np.minimum(r, I - 1)]

    # coverage: every user once, every item once
    keys = [np.arange(U, dtype=np.int64) * I + draw_items(U),
            rng.integers(0, U, size=I, dtype=np.int64) * I + np.arange(I, dtype=np.int64)]
    key = np.unique(np.concatenate(keys))
    while key.size < num_inter:
        need = num_inter - key.size
        n = int(need * 1.15) + 16
        new = rng.integers(0, U, size=n, dtype=np.int64) * I + draw_items(n)
        key = np.unique(np.concatenate([key, new]))
    if key.size > num_inter:
        # drop random surplus pairs but never the coverage pairs
        cover = np.unique(np.concatenate(keys))
        extra = np.setdiff1d(key, cover, assume_unique=True)
        drop = rng.choice(extra.size, size=key.size - num_inter, replace=False)
        extra = np.delete(extra, drop)
        key = np.union1d(cover, extra)
    users = key // I
    items = key % I
    # per-user random order, then 80/10/10 split by position inside the user
    order = np.lexsort((rng.random(key.size), users))
    users, items = users[order], items[order]
    counts = np.bincount(users, minlength=U)
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    pos = np.arange(key.size) - np.repeat(start, counts)
    n_hold = np.repeat(counts // 10, counts)
    is_test = pos < n_hold
    is_valid = (pos >= n_hold) & (pos < 2 * n_hold)
    is_train = ~(is_test | is_valid)
    pairs = np.stack([users, items], axis=1)
    # shuffle the train file rows so that id remapping by first appearance is non-trivial
    tr = pairs[is_train]
    tr = tr[rng.permutation(tr.shape[0])]
    return Interactions(U, I, tr, pairs[is_valid], pairs[is_test])


def make_features(num_items: int, dims, seed: int = 2022):
    """fp32 N(0,1) item features ``[I x D]`` per modality, indexed by raw item id."""
    rng = np.random.default_rng(seed + 1)
    out = []
    for d in dims:
        if d:
            out.append(rng.standard_normal((num_items, d), dtype=np.float32))
        else:
            out.append(None)
    return out


def make_shape(name: str, seed: int = 2022):
    U, I, n, dims = SHAPES[name]
    inter = make_interactions(U, I, n, seed)
    feats = make_features(I, dims, seed)
    return inter, feats


def write_reference_files(path: str, name: str, inter: Interactions, feats) -> None:
    """Write ``<path>/<name>.{train,valid,test}`` and the feature files as the reference reads them."""
    os.makedirs(path, exist_ok=True)
    for split in ("train", "valid", "test"):
        np.savetxt(os.path.join(path, f"{name}.{split}"), getattr(inter, split), fmt="%d", delimiter=",")
    v, a, t = feats
    if name == "kwai":
        import torch
        torch.save(torch.from_numpy(v), os.path.join(path, "kwai_feat_v.pt"))
    else:
        np.save(os.path.join(path, f"{name}_FeatureVideo_normal.npy"), v)
        np.save(os.path.join(path, f"{name}_FeatureAudio_avg_normal.npy"), a)
        np.save(os.path.join(path, f"{name}_FeatureText_stl_normal.npy"), t)


def make_words(num_items: int, seed: int = 2022, vocab: int = 11574, max_words: int = 6) -> np.ndarray:
    """``[2 x n_pairs]`` int64 (raw item id, word id) - the layout of ``tiktok_textual_feat.pt``
    (``data/dataset.py:166-173``); every item has 1..max_words words, in shuffled pair order."""
    rng = np.random.default_rng(seed + 2)
    cnt = rng.integers(1, max_words + 1, size=num_items)
    item = np.repeat(np.arange(num_items, dtype=np.int64), cnt)
    word = rng.integers(0, vocab, size=item.size, dtype=np.int64)
    perm = rng.permutation(item.size)
    return np.stack([item[perm], word[perm]])


def write_tiktok_files(path: str, inter: Interactions, v, a, words) -> None:
    """The literal ``tiktok`` layout: CSV splits + ``tiktok_{visual,audio,textual}_feat.pt`` (``dataset.py:164-166``)."""
    import torch
    os.makedirs(path, exist_ok=True)
    for split in ("train", "valid", "test"):
        np.savetxt(os.path.join(path, f"tiktok.{split}"), getattr(inter, split), fmt="%d", delimiter=",")
    torch.save(torch.from_numpy(v), os.path.join(path, "tiktok_visual_feat.pt"))
    torch.save(torch.from_numpy(a), os.path.join(path, "tiktok_audio_feat.pt"))
    torch.save(torch.from_numpy(words), os.path.join(path, "tiktok_textual_feat.pt"))
