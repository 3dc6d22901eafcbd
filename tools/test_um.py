#!/usr/bin/env python
"""testUM.py-style driver on the B200 path (sequence of testUM.py:43-175): seed 1, dataset of `.raw` files ->
per item {input, target, names, A_emb} -> reverse SDE from the LQ image's noise state -> metrics on x/2+0.5 ->
[LQ | restored | GT] triptych `.raw`.

    python tools/test_um.py --flist data.json --result-root out [--weights ckpt.pt] [--artifact-type "speckle in OCT" ...]

Without --flist a small synthetic data set is generated (no checkpoints or data ship with the reference snapshot).
"""
import argparse
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from instancediff_b200 import ConditionalUNet, create_sde  # noqa: E402
from instancediff_b200 import data as D  # noqa: E402


def run(flist, result_root, artifact_types, weights=None, max_items=1000000, T=-1, seed=1, device="cuda:0"):
    torch.manual_seed(seed)                                           # testUM.py:35-43
    np.random.seed(seed)
    dev = torch.device(device)
    net = ConditionalUNet(device=dev, seed=seed)
    if weights:
        net.load_state_dict(torch.load(weights, map_location="cpu"))
    sde = create_sde({"noise_net": net}, dict(max_sigma=0.4, T=100, schedule="cosine", eps=0.01), device=dev)   # :91
    sde.noise_source = "philox"
    ds = D.SpeckleMedDataset(flist, phase="test", max_dataset_size=max_items, opt={"name": "test_b200"},
                             use_artifact_type=artifact_types)
    results = {}
    with torch.no_grad():                                             # :109
        for i in range(len(ds)):
            item = ds[i]
            LQ, GT = item["LQ"][None].to(dev), item["GT"][None]       # [1,1,224,224], inputs in [-1, 1]
            emb = item["A_emb"][None].to(dev)                         # [1,1,D]
            sde.set_mu(LQ)
            sde.philox_seed, sde.philox_offset = seed, i * LQ.numel()
            tic = time.time()
            x0 = sde.reverse_sde(sde.noise_state(LQ), T=T, image_context=emb)
            torch.cuda.synchronize(dev)
            toc = time.time()
            pred = D.to_unit_range(x0.detach().cpu().numpy())         # :151-152
            target = D.to_unit_range(GT.numpy())
            rmse, psnr = D.rmse_psnr(pred, target)
            path = D.save_triptych(item["LQ"], x0, item["GT"], result_root, item["name"], i)
            r = results.setdefault(item["name"], dict(num=0, RMSE=[], PSNR=[], time=[]))
            r["num"] += 1
            r["RMSE"].append(rmse)
            r["PSNR"].append(psnr)
            r["time"].append(toc - tic)
            print(f" Testing {i}, {item['GT_path']}: RMSE={rmse:.5f}, PSNR={psnr:.3f}, {toc - tic:.2f} s -> {path}")
    for name, r in results.items():
        print(f"{name}: n={r['num']} RMSE={np.mean(r['RMSE']):.5f} PSNR={np.mean(r['PSNR']):.3f} "
              f"mean time {np.mean(r['time']):.2f} s")
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--flist", default=None)
    ap.add_argument("--result-root", default=None)
    ap.add_argument("--weights", default=None)
    ap.add_argument("--artifact-type", nargs="*", default=None)
    ap.add_argument("--max-items", type=int, default=1000000)
    ap.add_argument("--T", type=int, default=-1)
    args = ap.parse_args()
    tmp = None
    if args.flist is None:
        tmp = tempfile.TemporaryDirectory()
        args.flist = D.make_synthetic_dataset(tmp.name)
        args.artifact_type = args.artifact_type or D.MODALITY_NAMES[:2]
        args.max_items = min(args.max_items, 2)
    root = args.result_root or os.path.join(tempfile.gettempdir(), "idiff_results")
    run(args.flist, root, args.artifact_type or [], args.weights, args.max_items, args.T)


if __name__ == "__main__":
    main()
