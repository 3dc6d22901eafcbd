"""TEST INFRASTRUCTURE (never imported by the product).  tests/golden/loader.json = what the reference's own
`data.create_dataloader` (data/__init__.py:8-34) and `data.data_sampler.DistIterSampler` (data/data_sampler.py) produce
for a list of settings.  Dev container only:   PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_loader.py
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "loader.json")
REF = os.environ.get("IDIFF_REFERENCE", "/root/reference")

SAMPLER_CASES = [dict(n=7, num_replicas=2, rank=0, ratio=3, epoch=0), dict(n=7, num_replicas=2, rank=1, ratio=3, epoch=0),
                 dict(n=5, num_replicas=3, rank=2, ratio=100, epoch=4), dict(n=11, num_replicas=1, rank=0, ratio=1, epoch=9)]
LOADER_CASES = [
    dict(dataset_opt=dict(phase="train", n_workers=0, batch_size=8), opt=dict(dist=False, gpu_ids=[0, 1])),
    dict(dataset_opt=dict(phase="val"), opt=None),
    dict(dataset_opt=dict(phase="test"), opt=None),
]


class Toy(torch.utils.data.Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i


def describe(dl):
    return dict(batch_size=dl.batch_size, drop_last=dl.drop_last, pin_memory=dl.pin_memory, num_workers=dl.num_workers,
                sampler=type(dl.sampler).__name__)


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import data as R                                    # the reference's package
    from data.data_sampler import DistIterSampler
    sys.path.pop(0)
    out = {"sampler": [], "loader": []}
    for c in SAMPLER_CASES:
        s = DistIterSampler(Toy(c["n"]), c["num_replicas"], c["rank"], c["ratio"])
        s.set_epoch(c["epoch"])
        out["sampler"].append(dict(case=c, length=len(s), indices=list(iter(s))))
    for c in LOADER_CASES:
        out["loader"].append(dict(case=c, loader=describe(R.create_dataloader(Toy(20), c["dataset_opt"], c["opt"]))))
    json.dump(out, open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
