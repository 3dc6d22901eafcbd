"""Checkpoint file naming and key clean-up of the reference's model wrapper (host side, torch-free logic on
plain dict keys).

Reference behaviour restated here:
  * ``CLIPDriftModel.save_network``  ``models/drift_noise_model.py:670-681`` -- file ``{iter}_{label}.pth``, the
    ``state_dict`` of the unwrapped (``.module``) network with CPU tensors;
  * ``CLIPDriftModel.load_network``  ``:706-731`` -- ``torch.load`` + key clean-up in two flavours (``dist`` /
    single process) + ``load_state_dict(strict)``;
  * ``CLIPDriftModel.save`` / ``load``  ``:683-692, 733-755`` -- labels ``DN`` / ``NN`` (drift / noise net at
    ``iter``) and ``DN_ema`` / ``NN_ema`` (always under the literal iteration label ``'lastest'``, sic).

``ema_pytorch.EMA`` (``:7,151-152``, not in the snapshot) stores the averaged copy under ``ema_model.*`` next to
``online_model.*`` and the scalar buffers ``initted`` / ``step``; ``ema_weights`` picks the averaged copy so an
``*_ema.pth`` file can feed the flat-key ``ConditionalUNet``.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, Mapping

_WRAPPED = ("CLIP_ScoreMapModule", "online_model", "ema_model")
NET_LABELS = {"drift_net": "DN", "noise_net": "NN"}
EMA_ITER_LABEL = "lastest"                      # spelling of models/drift_noise_model.py:688-691, 747


def clean_key(key: str, dist: bool = False) -> str:
    """One key through ``load_network``'s clean-up (``models/drift_noise_model.py:713-729``)."""
    if not dist:
        return key.replace("module.", "")       # :726-728 -- every occurrence, not only a prefix
    k = key[7:] if key.startswith("module.") else key
    for name in _WRAPPED:                       # :718-723 -- re-insert the DDP wrapper level of sub-modules
        if name in k and name + ".module" not in k:
            k = k.replace(name, name + ".module")
    return k


def clean_state_dict_keys(state: Mapping[str, object], dist: bool = False) -> "OrderedDict[str, object]":
    out: "OrderedDict[str, object]" = OrderedDict()
    for k, v in state.items():                  # later duplicates overwrite earlier ones, as in the reference
        out[clean_key(k, dist)] = v
    return out


def ema_weights(state: Mapping[str, object]) -> "OrderedDict[str, object]":
    """The averaged network of an ``ema_pytorch.EMA`` state dict (keys ``ema_model.*``); a plain state dict is
    returned unchanged."""
    if not any(k.startswith("ema_model.") for k in state):
        return OrderedDict(state)
    out: "OrderedDict[str, object]" = OrderedDict()
    for k, v in state.items():
        if k.startswith("ema_model."):
            k = k[len("ema_model."):]
            out[k[7:] if k.startswith("module.") else k] = v      # the dist clean-up re-inserts this level
    return out


def network_path(save_dir: str, iter_label, network_label: str) -> str:
    return os.path.join(save_dir, "{}_{}.pth".format(iter_label, network_label))     # :671,674


def save_network(network, network_label: str, iter_label, save_dir: str) -> str:
    import torch
    network = getattr(network, "module", network)                                     # :675-678
    state = {k: v.detach().cpu() for k, v in network.state_dict().items()}            # :679-681
    path = network_path(save_dir, iter_label, network_label)
    torch.save(state, path)
    return path


def load_network(load_path: str, network, strict: bool = True, dist: bool = False, use_ema: bool = False, key_map=None):
    """``key_map``: optional ``f(name) -> name`` (or None to drop the tensor) applied after the reference's clean-up --
    the hook for checkpoints whose module tree is not the App. A restatement's (the upstream network source is not in the
    snapshot, so the real key names are unknown here)."""
    import torch
    network = getattr(network, "module", network)
    state: Dict[str, object] = torch.load(load_path, map_location="cpu", weights_only=True)   # tensors only: no pickle code
    state = clean_state_dict_keys(state, dist=dist)
    if use_ema:
        state = ema_weights(state)
    if key_map is not None:
        state = {nk: v for nk, v in ((key_map(k), v) for k, v in state.items()) if nk is not None}
    network.load_state_dict(state, strict=strict)
    return network
