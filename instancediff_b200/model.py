"""Driver-facing model object: the slice of ``CLIPDriftModel`` (``models/drift_noise_model.py:27``) that
``testUM.py`` touches, on the B200 path.

``testUM.py`` drives the reference in this order (SURVEY.md section 8b "Driver contract"):
``create_model(train_opt, model_opt, phase='test')`` ``:74`` -> ``.load(iter, dir)`` ``:76`` ->
``.get_nets(use_ema)`` ``:90`` -> ``create_sde(nets, sde_opt)`` ``:91`` -> ``.set_sde(sde)`` ``:92`` ->
``.set_gpu(device)`` ``:95`` -> per item ``.feed_data({'input','target','names','A_emb'})`` ``:128-139`` ->
``.test()`` ``:142`` -> ``.get_visuals()`` ``:146``.  The object below keeps those names and argument meanings.

``test()`` in the reference calls ``driftSDE.reverse_ddpm`` (``models/drift_noise_model.py:650``), whose source is
not in the snapshot; here it runs the hot path this repository implements,
``IRSDE.reverse_sde`` from the LQ image's noise state with the noise net, and keeps the result convention
``out.detach().cpu().numpy()`` (``:652``).  Training-side members (optimizers, losses, EMA updates, text
encoders) are out of scope.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import checkpoint as ckpt
from .unet import ConditionalUNet


class RestorationModel:
    def __init__(self, nnet_settings: Optional[dict] = None, dnet_settings: Optional[dict] = None, dist: bool = False,
                 use_image_context: bool = True, with_drift_net: bool = False, device="cuda", seed: int = 1):
        self.device = torch.device(device)
        self.dist = dist
        if not use_image_context:
            # the network's cross-attention layers are built with context_dim = 512 (config.yml:113,132); a forward
            # without the embedding has no defined meaning in the App. A architecture, so it is refused up front
            # instead of failing inside test()
            raise ValueError("RestorationModel: use_image_context=False is not supported -- the B200 network "
                             "conditions every SpatialTransformer on the image embedding (config.yml:132)")
        self.use_image_context = use_image_context                       # models/drift_noise_model.py:186-189
        self.noise_net = ConditionalUNet(device=self.device, seed=seed, **_net_kwargs(nnet_settings))
        # the reverse-SDE path only consumes the noise net; the drift net is built on request so that both
        # checkpoint files of a reference run can be loaded and handed out by get_nets()
        self.drift_net = (ConditionalUNet(device=self.device, seed=seed + 1, **_net_kwargs(dnet_settings))
                          if with_drift_net else None)
        self.nn_ema: Optional[ConditionalUNet] = None
        self.dn_ema: Optional[ConditionalUNet] = None
        self.sde = None
        self.input = self.target = self.A_emb = None
        self.names = None
        self.visuals = None
        self.noise_source = "philox"
        self.seed = seed
        self.test_T = -1                  # -1: the SDE's full schedule (IRSDE.reverse_sde default, :245)
        self._items = 0

    # ---- checkpoints (models/drift_noise_model.py:670-755) ------------------------------------------------
    def load_network(self, load_path, network, strict=True, use_ema=False):
        return ckpt.load_network(load_path, network, strict=strict, dist=self.dist, use_ema=use_ema)

    def save_network(self, network, network_label, iter_label, save_dir):
        return ckpt.save_network(network, network_label, iter_label, save_dir)

    def save(self, iter_label, save_dir):
        os.makedirs(save_dir, exist_ok=True)
        if self.drift_net is not None:
            self.save_network(self.drift_net, "DN", iter_label, save_dir)
        self.save_network(self.noise_net, "NN", iter_label, save_dir)

    def load(self, iter_label, save_dir):
        """``{iter}_NN.pth`` (and ``{iter}_DN.pth`` when a drift net was requested); ``lastest_*_ema.pth`` files
        are loaded when present (the reference requires them, ``:746-755``; a sampling-only directory may not)."""
        self.load_network(ckpt.network_path(save_dir, iter_label, "NN"), self.noise_net)
        if self.drift_net is not None:
            self.load_network(ckpt.network_path(save_dir, iter_label, "DN"), self.drift_net)
        for label, attr, src in (("NN_ema", "nn_ema", self.noise_net), ("DN_ema", "dn_ema", self.drift_net)):
            path = ckpt.network_path(save_dir, ckpt.EMA_ITER_LABEL, label)
            if src is not None and os.path.exists(path):
                net = ConditionalUNet(device=self.device, **{k: v for k, v in src.cfg.items()})
                setattr(self, attr, self.load_network(path, net, use_ema=True))

    def get_nets(self, use_ema=False) -> Dict[str, object]:                # :657-668
        if use_ema:
            if self.nn_ema is None:
                raise RuntimeError("get_nets(use_ema=True): no lastest_NN_ema.pth was loaded")
            return {"noise_net": self.nn_ema, "drift_net": self.dn_ema}
        return {"noise_net": self.noise_net, "drift_net": self.drift_net}

    # ---- device / mode ----------------------------------------------------------------------------------
    def set_sde(self, sde):                                                # :179-180
        self.sde = sde

    def set_gpu(self, device):                                             # :635-642
        self.device = torch.device(device)
        for net in (self.noise_net, self.drift_net, self.nn_ema, self.dn_ema):
            if net is not None:
                net.to(self.device)

    def set_eval(self):                                                    # :631-633
        return self

    # ---- per item ---------------------------------------------------------------------------------------
    def feed_data(self, data):                                             # :182-189 (test-time part)
        self.input = data["input"].to(self.device, torch.float32).contiguous()
        self.target = data["target"].to(self.device) if data.get("target") is not None else None
        self.names = data.get("names")
        self.A_emb = data["A_emb"].to(self.device, torch.float32).contiguous() if self.use_image_context else None

    def test(self):                                                        # :648-652
        if self.sde is None:
            raise RuntimeError("test(): call set_sde() first (testUM.py:91-92)")
        sde = self.sde
        sde.set_mu(self.input)
        if self.noise_source == "philox":
            sde.noise_source = "philox"
            sde.philox_seed, sde.philox_offset = self.seed, self._items * self.input[0].numel()
        kwargs = {} if self.A_emb is None else {"image_context": self.A_emb}
        out = sde.reverse_sde(sde.noise_state(self.input), T=self.test_T, **kwargs)
        self._items += self.input.shape[0]
        self.visuals = out.detach().cpu().numpy()

    def get_visuals(self):                                                 # :654-655
        return self.visuals


def _net_kwargs(settings: Optional[dict]) -> dict:
    """Network hyper-parameters of ``Configurations/config.yml:106-136`` that the B200 kernels are built for.
    ``out_nc: 5`` / ``text_module: scoremap`` (``config.yml:110,114``) describe the Score-Map variant whose source
    is absent (the IRSDE path needs out == in == 1 channel); for such a block only the shared
    backbone parameters are taken."""
    if not settings:
        return {}
    keys = ["nf", "ch_mult", "context_dim", "down_kernel"]
    if settings.get("text_module") != "scoremap":
        keys += ["in_nc", "out_nc"]
    return {k: settings[k] for k in keys if k in settings}


def create_model(train_opt: Optional[dict] = None, model_opt: Optional[dict] = None, phase: str = "test",
                 device="cuda", seed: int = 1) -> RestorationModel:
    """``create_model(train_opt, model_opt, phase)`` of ``testUM.py:74`` (factory:
    ``models/drift_noise_model.py:758-815``).  Only ``phase='test'`` exists on this path."""
    if phase != "test":
        raise NotImplementedError("instancediff_b200 implements the sampling path; phase must be 'test'")
    train_opt, model_opt = train_opt or {}, model_opt or {}
    return RestorationModel(nnet_settings=model_opt.get("nnet_settings"), dnet_settings=model_opt.get("dnet_settings"),
                            dist=bool(train_opt.get("dist", False)),
                            use_image_context=bool(model_opt.get("use_image_context", True)),
                            with_drift_net=bool(model_opt.get("with_drift_net", False)), device=device, seed=seed)
