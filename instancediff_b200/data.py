"""I/O edges of the sampling path: the `.raw` float32 loader with per-modality
normalisation of ``data/MedSpeckle.py:12-88`` and the result writer / metrics of ``testUM.py:151-173``.

Host-side numpy/torch code, same names, argument meaning and arithmetic order as the reference so that a
``testUM.py``-style driver can switch imports; parity is pinned bit-for-bit by ``tests/test_data_io.py`` against
fixtures produced by importing the reference's own ``SpeckleMedDataset`` (``oracle/gen_golden_io.py``).
"""
from __future__ import annotations

import json
import logging
import math
import os
import platform
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.utils.data as data

RAW_SIDE = 224                     # data/MedSpeckle.py:44-45: every image is a 1 x 224 x 224 float32 `.raw` file


def normalize_modality(A: torch.Tensor, B: torch.Tensor, name: str):
    """Per-modality range normalisation then [-1, 1] (data/MedSpeckle.py:54-70); returns new tensors."""
    A, B = A.clone(), B.clone()
    if name == "scatter artifact in CT":                     # :54-60  clamp to [0, 1800], / 1800
        A[A < 0] = 0
        A[A > 1800] = 1800
        B[B < 0] = 0
        B[B > 1800] = 1800
        A = A / 1800.0
        B = B / 1800.0
    if name == "noise in cryo-EM image":                     # :62-66  clip to [0, 255], / 255
        A = torch.from_numpy(np.clip(A.numpy(), 0.0, 255.0))
        B = torch.from_numpy(np.clip(B.numpy(), 0.0, 255.0))
        A = A / 255.0
        B = B / 255.0
    return A * 2.0 - 1.0, B * 2.0 - 1.0                      # :68-69


class SpeckleMedDataset(data.Dataset):
    """Mirror of ``data/MedSpeckle.py::SpeckleMedDataset``: a JSON file list ``{phase: [{A, B, A_emb, name}, ...]}``
    filtered by ``use_artifact_type`` and truncated to ``max_dataset_size``; items are
    ``{'LQ', 'GT', 'LQ_path', 'GT_path', 'name', 'A_emb'}`` with LQ / GT fp32 ``[1,224,224]`` in [-1, 1] and ``A_emb``
    fp32 ``[1, D]`` (the pre-computed BiomedCLIP image embedding)."""

    def __init__(self, data_flist, phase="train", max_dataset_size=1000000, opt=None, use_artifact_type: Sequence[str] = ()):
        self.use_artifact_type = list(use_artifact_type)
        self.opt = opt
        with open(data_flist, "r") as f:
            df = json.load(f)[phase]
        self.df: List[dict] = [item for item in df if item["name"] in self.use_artifact_type]
        if max_dataset_size < len(self.df):
            self.df = self.df[:max_dataset_size]

    def __len__(self):
        return len(self.df)

    def __getitem__(self, index) -> Dict[str, object]:
        item = self.df[index]
        a_img = np.fromfile(item["A"], dtype=np.float32).reshape(1, RAW_SIDE, RAW_SIDE)
        b_img = np.fromfile(item["B"], dtype=np.float32).reshape(1, RAW_SIDE, RAW_SIDE)
        A_emb = np.fromfile(item["A_emb"], dtype=np.float32).reshape(1, -1)
        A, B = normalize_modality(torch.from_numpy(a_img), torch.from_numpy(b_img), item["name"])
        return {"LQ": A, "GT": B, "LQ_path": item["A"], "GT_path": item["B"], "name": item["name"],
                "A_emb": torch.from_numpy(A_emb)}


def create_SpeckleMedDataset(params=None):
    """``data/MedSpeckle.py:76-88``: phase = prefix of ``params['name']`` before the first underscore."""
    dataset_file = params["dataset_file_win"] if platform.system() == "Windows" else params["dataset_file"]
    return SpeckleMedDataset(dataset_file, phase=params["name"].split("_")[0], max_dataset_size=params["max_dataset_size"],
                             opt=params, use_artifact_type=params["use_artifact_type"])


def create_dataset(dataset_opt):
    """``data/__init__.py:37-52``: factory keyed by ``dataset_opt['mode']``; only ``SpeckleMed`` exists upstream."""
    mode = dataset_opt["mode"]
    if mode != "SpeckleMed":
        raise NotImplementedError("Dataset [{:s}] is not recognized.".format(mode))
    dataset = create_SpeckleMedDataset(dataset_opt)
    logging.getLogger("base").info("Dataset [{:s} - {:s}] is created.".format(dataset.__class__.__name__, dataset_opt["name"]))
    return dataset


def create_dataloader(dataset, dataset_opt, opt=None, sampler=None):
    """``data/__init__.py:8-34``.  train: per-rank batch ``batch_size // world_size`` without shuffling when
    ``opt['dist']`` (the sampler shuffles), else the full batch shuffled with ``n_workers * len(gpu_ids)`` workers;
    ``drop_last`` and pinned memory.  Any other phase: batch 1, in order, no workers, pinned only for ``val``."""
    phase = dataset_opt["phase"]
    if phase == "train":
        if opt["dist"]:
            world_size = torch.distributed.get_world_size()
            num_workers = dataset_opt["n_workers"]
            assert dataset_opt["batch_size"] % world_size == 0
            batch_size, shuffle = dataset_opt["batch_size"] // world_size, False
        else:
            num_workers = dataset_opt["n_workers"] * len(opt["gpu_ids"])
            batch_size, shuffle = dataset_opt["batch_size"], True
        return data.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, sampler=sampler,
                               drop_last=True, pin_memory=True)
    return data.DataLoader(dataset, batch_size=1, shuffle=False, num_workers=0, pin_memory=(phase == "val"))


class DistIterSampler(data.Sampler):
    """``data/data_sampler.py:13-67``: the data set enlarged ``ratio`` times for iteration-oriented training.  Each
    epoch draws ONE permutation of ``num_samples * num_replicas`` virtual indices from a generator seeded with the
    epoch number, folds it onto the data set with a modulo and gives rank r every ``num_replicas``-th entry from r."""

    def __init__(self, dataset, num_replicas=None, rank=None, ratio=100):
        if num_replicas is None or rank is None:
            if not torch.distributed.is_available():
                raise RuntimeError("Requires distributed package to be available")
            num_replicas = torch.distributed.get_world_size() if num_replicas is None else num_replicas
            rank = torch.distributed.get_rank() if rank is None else rank
        self.dataset, self.num_replicas, self.rank, self.epoch = dataset, num_replicas, rank, 0
        self.num_samples = int(math.ceil(len(dataset) * ratio / num_replicas))
        self.total_size = self.num_samples * num_replicas

    def __iter__(self):
        gen = torch.Generator()
        gen.manual_seed(self.epoch)
        n = len(self.dataset)
        mine = (torch.randperm(self.total_size, generator=gen)[self.rank:self.total_size:self.num_replicas] % n).tolist()
        assert len(mine) == self.num_samples
        return iter(mine)

    def __len__(self):
        return self.num_samples

    def set_epoch(self, epoch):
        self.epoch = epoch


# ---- result side (testUM.py:144-173) -----------------------------------------------------------------------
def to_unit_range(x: np.ndarray) -> np.ndarray:
    return x / 2 + 0.5                                       # testUM.py:151-152


def rmse_psnr(pred: np.ndarray, target: np.ndarray, data_range: float = 1.0):
    """RMSE = sqrt(mean squared error) and PSNR = 10 log10(range^2 / mse) of two arrays already in [0, 1]
    (``testUM.py:160-161``: skimage's mean_squared_error / peak_signal_noise_ratio, both computed in float64)."""
    err = np.mean((np.asarray(pred, dtype=np.float64) - np.asarray(target, dtype=np.float64)) ** 2, dtype=np.float64)
    return float(np.sqrt(err)), float(10 * np.log10((data_range ** 2) / err))


def save_triptych(LQ, pred, GT, result_root: str, name: str, index: int) -> str:
    """``testUM.py:168-171``: [LQ | restored | GT] concatenated along the width, written as raw float32 to
    ``{result_root}/{name}/{i}_{W}x{H}x1.raw``; returns the path."""
    parts = [np.asarray(t.detach().cpu().numpy() if torch.is_tensor(t) else t, dtype=np.float32).squeeze() for t in (LQ, pred, GT)]
    to_save = np.concatenate(parts, axis=-1)
    out_dir = os.path.join(result_root, name)
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"{index}_{to_save.shape[-1]}x{to_save.shape[-2]}x1.raw")
    to_save.tofile(path)
    return path


# ---- synthetic data set (no data ships with the reference snapshot) ---------------------------------------------
MODALITY_NAMES = ["speckle in OCT", "scatter artifact in CT", "noise in cryo-EM image", "speckle in ultra sound"]


def make_synthetic_dataset(root: str, seed: int = 7) -> str:
    """Writes six seeded `.raw` items (A, B, A_emb) whose value ranges exercise every clamp of
    ``normalize_modality`` plus a JSON file list under ``root``; returns the file-list path.  Used by the parity
    fixtures (``oracle/gen_golden_io.py``), the tests and ``tools/test_um.py``."""
    rng = np.random.default_rng(seed)
    items = []
    for i, name in enumerate(MODALITY_NAMES + MODALITY_NAMES[:2]):
        scale = {"scatter artifact in CT": 2400.0, "noise in cryo-EM image": 320.0}.get(name, 1.0)
        a = (rng.standard_normal((RAW_SIDE, RAW_SIDE)) * 0.35 + 0.5).astype(np.float32) * np.float32(scale)
        b = (rng.random((RAW_SIDE, RAW_SIDE)) * 1.1 - 0.05).astype(np.float32) * np.float32(scale)
        emb = rng.standard_normal(512).astype(np.float32)
        paths = {}
        for key, arr in (("A", a), ("B", b), ("A_emb", emb)):
            p = os.path.join(root, f"item{i}_{key}.raw")
            arr.tofile(p)
            paths[key] = p
        items.append(dict(paths, name=name))
    flist = os.path.join(root, "flist.json")
    with open(flist, "w") as f:
        json.dump({"test": items, "train": items[:2]}, f)
    return flist
