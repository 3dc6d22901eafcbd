"""TEST INFRASTRUCTURE (never imported by the product).  Generates tests/golden/ckpt_keys.json by EXECUTING the
reference's own `CLIPDriftModel.load_network` (models/drift_noise_model.py:706-731) on a list of key names.

`models.drift_noise_model` cannot be imported here (`import clip`, `ema_pytorch`, `models.modules` are missing,
SURVEY.md section 8c), so the method's source is cut out of the file with `ast`, compiled on its own and run with
a stub `self` / `torch.load` / network that records the cleaned keys.  Run in the dev container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_ckpt.py
"""
import ast
import json
import os
from collections import OrderedDict

REF = "/root/reference/models/drift_noise_model.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ckpt_keys.json")

KEYS = [
    "init_conv.weight", "module.init_conv.weight", "module.module.init_conv.bias", "downs.0.0.block1.proj.weight",
    "ups.1.module.block2.weight", "CLIP_ScoreMapModule.proj.weight", "module.CLIP_ScoreMapModule.proj.weight",
    "CLIP_ScoreMapModule.module.proj.weight", "online_model.mid_block1.weight", "module.online_model.mid_block1.weight",
    "online_model.module.mid_block1.weight", "ema_model.final_conv.bias", "module.ema_model.final_conv.bias",
    "ema_model.module.final_conv.bias", "ema_model.CLIP_ScoreMapModule.w", "initted", "step", "a.module.b.module.c",
    "submodule.weight", "module.submodule.weight",
]


def reference_load_network():
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CLIPDriftModel")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "load_network")
    mod = ast.Module(body=[fn], type_ignores=[])
    captured = {}

    class _Torch:
        @staticmethod
        def load(path):
            return OrderedDict((k, i) for i, k in enumerate(KEYS))

    class _NN:
        class DataParallel:
            pass

    class _DDP:
        pass

    env = {"torch": _Torch, "nn": _NN, "DistributedDataParallel": _DDP, "OrderedDict": OrderedDict}
    exec(compile(mod, REF, "exec"), env)

    class _Net:
        def load_state_dict(self, sd, strict=True):
            captured["sd"] = sd

    def run(dist):
        self = type("S", (), {"dist": dist})()
        env["load_network"](self, "unused.pth", _Net(), strict=True)
        return list(captured["sd"].items())
    return run


def main():
    run = reference_load_network()
    out = {"keys": KEYS, "single": run(False), "dist": run(True)}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT, len(KEYS), "keys")


if __name__ == "__main__":
    main()
