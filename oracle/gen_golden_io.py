"""Generates tests/golden/io_medspeckle.npz by running the REFERENCE's own data/MedSpeckle.py::SpeckleMedDataset
(imported from /root/reference; dev container only) on seeded synthetic `.raw` files.  The fixture stores the
seeds, the file list and, per item, the sha256 of the reference's LQ / GT / A_emb bytes plus a few sampled values,
so tests/test_data_io.py can rebuild the same inputs and require bit-identical outputs from
instancediff_b200.data.SpeckleMedDataset.  Test infrastructure only."""
import hashlib
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from instancediff_b200.data import MODALITY_NAMES as NAMES, make_synthetic_dataset as make_inputs  # noqa: E402,F401


def digest(t):
    return hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()


def main():
    spec = importlib.util.spec_from_file_location("ref_medspeckle", "/root/reference/data/MedSpeckle.py")
    ref = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(ref)
    out = {}
    with tempfile.TemporaryDirectory() as root:
        flist = make_inputs(root)
        use = NAMES[:3]
        ds = ref.SpeckleMedDataset(flist, phase="test", max_dataset_size=4, opt={"name": "test_x"}, use_artifact_type=use)
        out["len"] = np.int64(len(ds))
        for i in range(len(ds)):
            it = ds[i]
            for key in ("LQ", "GT", "A_emb"):
                arr = it[key].numpy()
                out[f"{i}_{key}_sha"] = np.array(digest(arr))
                out[f"{i}_{key}_probe"] = arr.reshape(-1)[:: max(1, arr.size // 64)][:64].copy()
                out[f"{i}_{key}_shape"] = np.array(arr.shape)
            out[f"{i}_name"] = np.array(it["name"])
            out[f"{i}_base"] = np.array(os.path.basename(it["GT_path"]))
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "io_medspeckle.npz"), **out)
    print("wrote io_medspeckle.npz with", int(out["len"]), "items")


if __name__ == "__main__":
    main()
