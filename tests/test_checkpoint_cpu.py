"""Checkpoint naming / key clean-up (models/drift_noise_model.py:670-755) and the driver protocol surface
(testUM.py:74-146).  Host logic only: runs without a GPU."""
import inspect
import os
from collections import OrderedDict

import torch

from instancediff_b200 import checkpoint as C


def test_single_process_cleanup_removes_every_module_level():
    # :724-728 -- `k.replace('module.', '')`, i.e. ALL occurrences
    sd = OrderedDict([("module.init_conv.weight", 1), ("ups.0.module.block1.weight", 2), ("plain.bias", 3),
                      ("module.online_model.module.x", 4)])
    out = C.clean_state_dict_keys(sd, dist=False)
    assert list(out) == ["init_conv.weight", "ups.0.block1.weight", "plain.bias", "online_model.x"]
    assert list(out.values()) == [1, 2, 3, 4]


def test_dist_cleanup_strips_prefix_and_reinserts_wrapped_submodules():
    # :714-723 -- leading 'module.' (7 chars) dropped once; the three wrapped sub-modules get '.module' back
    cases = {
        "module.init_conv.weight": "init_conv.weight",
        "init_conv.weight": "init_conv.weight",
        "module.CLIP_ScoreMapModule.proj.weight": "CLIP_ScoreMapModule.module.proj.weight",
        "CLIP_ScoreMapModule.module.proj.weight": "CLIP_ScoreMapModule.module.proj.weight",
        "online_model.mid.weight": "online_model.module.mid.weight",
        "module.ema_model.mid.weight": "ema_model.module.mid.weight",
        "ema_model.module.mid.weight": "ema_model.module.mid.weight",
        "a.module.b": "a.module.b",                       # inner levels are kept in dist mode
    }
    for k, want in cases.items():
        assert C.clean_key(k, dist=True) == want, k


def test_ema_weights_selects_averaged_copy():
    sd = OrderedDict([("initted", torch.tensor(True)), ("step", torch.tensor(7)), ("online_model.w", 1),
                      ("ema_model.w", 2), ("ema_model.module.v", 3)])
    assert dict(C.ema_weights(sd)) == {"w": 2, "v": 3}
    plain = OrderedDict([("w", 5)])
    assert dict(C.ema_weights(plain)) == {"w": 5}


def test_file_names_follow_the_reference(tmp_path):
    assert C.network_path("/x", 400000, "NN") == os.path.join("/x", "400000_NN.pth")       # :671
    assert C.network_path("/x", C.EMA_ITER_LABEL, "NN_ema") == os.path.join("/x", "lastest_NN_ema.pth")   # :691,753
    assert C.NET_LABELS == {"drift_net": "DN", "noise_net": "NN"}                             # :688-689

    class Wrapped(torch.nn.Module):                       # stands for DataParallel / DDP: `.module` is unwrapped (:675-678)
        def __init__(self):
            super().__init__()
            self.module = torch.nn.Linear(3, 2)

    w = Wrapped()
    path = C.save_network(w, "NN", 12, str(tmp_path))
    raw = torch.load(path)
    assert sorted(raw) == ["bias", "weight"] and all(not v.is_cuda for v in raw.values())
    # a DDP-saved file ('module.' prefixes) loads into a bare network in both clean-up modes
    torch.save({"module." + k: v + 1 for k, v in raw.items()}, path)
    for dist in (False, True):
        tgt = torch.nn.Linear(3, 2)
        C.load_network(path, tgt, strict=True, dist=dist)
        assert torch.equal(tgt.weight, raw["weight"] + 1) and torch.equal(tgt.bias, raw["bias"] + 1)
    # EMA container -> averaged copy
    torch.save({"initted": torch.tensor(True), "step": torch.tensor(3), **{"online_model." + k: v for k, v in raw.items()},
                **{"ema_model." + k: v * 2 for k, v in raw.items()}}, path)
    tgt = torch.nn.Linear(3, 2)
    C.load_network(path, tgt, strict=True, use_ema=True)
    assert torch.equal(tgt.weight, raw["weight"] * 2)


def test_driver_protocol_surface():
    """Names and argument lists testUM.py relies on (testUM.py:74-146; models/drift_noise_model.py:179-189,631-755)."""
    from instancediff_b200 import RestorationModel, create_model
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(create_model)[:3] == ["train_opt", "model_opt", "phase"]
    assert sig(RestorationModel.load) == ["self", "iter_label", "save_dir"]
    assert sig(RestorationModel.save) == ["self", "iter_label", "save_dir"]
    assert sig(RestorationModel.get_nets) == ["self", "use_ema"]
    assert sig(RestorationModel.set_sde) == ["self", "sde"]
    assert sig(RestorationModel.set_gpu) == ["self", "device"]
    assert sig(RestorationModel.feed_data) == ["self", "data"]
    assert sig(RestorationModel.test) == ["self"] and sig(RestorationModel.get_visuals) == ["self"]
    assert sig(RestorationModel.load_network)[:4] == ["self", "load_path", "network", "strict"]
    assert sig(RestorationModel.save_network) == ["self", "network", "network_label", "iter_label", "save_dir"]
    import pytest
    with pytest.raises(NotImplementedError):
        create_model({}, {}, phase="train")


def test_scoremap_config_block_maps_to_backbone_kwargs():
    from instancediff_b200.model import _net_kwargs
    cfg = dict(module_name="MSM_degEmb_Unet", in_nc=2, out_nc=5, nf=64, ch_mult=[1, 2, 4, 4], context_dim=512,
               text_module="scoremap", score_map_chan=16)                                      # config.yml:105-117
    assert _net_kwargs(cfg) == dict(nf=64, ch_mult=[1, 2, 4, 4], context_dim=512)
    assert _net_kwargs(dict(in_nc=2, out_nc=1, nf=64)) == dict(nf=64, in_nc=2, out_nc=1)
    assert _net_kwargs(None) == {}


def test_cleanup_matches_the_reference_executed_on_the_same_keys(golden_dir):
    """tests/golden/ckpt_keys.json = the reference's own `load_network` body run on these keys
    (oracle/gen_golden_ckpt.py); both flavours, including key order and duplicate resolution."""
    import json
    g = json.load(open(os.path.join(golden_dir, "ckpt_keys.json")))
    state = OrderedDict((k, i) for i, k in enumerate(g["keys"]))
    for mode, dist in (("single", False), ("dist", True)):
        got = [[k, v] for k, v in C.clean_state_dict_keys(state, dist=dist).items()]
        assert got == g[mode], mode


def test_cleanup_properties():
    """Single-process clean-up is idempotent and leaves no 'module.' behind; the dist flavour is idempotent too and
    never changes a key that has no wrapper level to fix."""
    from hypothesis import given, settings, strategies as st
    part = st.sampled_from(["module", "ema_model", "online_model", "CLIP_ScoreMapModule", "conv1", "weight", "0", "submodule"])
    keys = st.lists(part, min_size=1, max_size=6).map(".".join)

    @settings(max_examples=300, deadline=None)
    @given(keys)
    def check(k):
        s = C.clean_key(k, dist=False)
        assert "module." not in s and C.clean_key(s, dist=False) == s
        d = C.clean_key(k, dist=True)
        assert C.clean_key(d, dist=True) == d or d.startswith("module.")     # 'module.module.x' loses one level per pass
        if not any(w in k for w in ("module", "ema_model", "online_model", "CLIP_ScoreMapModule")):
            assert d == k == s
    check()
