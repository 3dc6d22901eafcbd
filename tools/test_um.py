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

from instancediff_b200 import create_model, create_sde  # noqa: E402
from instancediff_b200 import data as D  # noqa: E402
from instancediff_b200.metrics import ssim_unit_range  # noqa: E402


def run(flist, result_root, artifact_types, weights=None, max_items=1000000, T=-1, seed=1, device="cuda:0",
        pth_dir=None, iter_label=None, use_ema=False, batch=1):
    torch.manual_seed(seed)                                           # testUM.py:35-43
    np.random.seed(seed)
    dev = torch.device(device)
    model = create_model({"dist": False}, {"use_image_context": True}, phase="test", device=dev, seed=seed)   # :74
    if pth_dir is not None:
        model.load(iter_label, pth_dir)                               # :76  {iter}_NN.pth (+ lastest_NN_ema.pth)
    elif weights:
        model.load_network(weights, model.noise_net)
    sde_opt = dict(max_sigma=0.4, T=100, schedule="cosine", eps=0.01)
    nets = model.get_nets(use_ema=use_ema)                            # :90
    sde = create_sde(nets, sde_opt, device=dev)                       # :91
    model.set_sde(sde)                                                # :92
    model.set_gpu(dev)                                                # :95
    model.test_T = T                                                  # debugging aid: run only the last T steps
    ds = D.SpeckleMedDataset(flist, phase="test", max_dataset_size=max_items, opt={"name": "test_b200"},
                             use_artifact_type=artifact_types)
    results = {}

    def flush(group):
        """One reverse process for a group of items (testUM.py runs batch_size 1; a group of N items is the same
        computation per item -- every sample draws from its own Philox stream, indexed by its position in the data
        set, so the restored images do not depend on the grouping)."""
        LQ = torch.stack([it["LQ"] for _, it in group])               # [N,1,224,224], inputs in [-1, 1]
        GT = torch.stack([it["GT"] for _, it in group])
        model.feed_data({"input": LQ.to(dev), "target": GT.to(dev), "names": [it["name"] for _, it in group],
                         "A_emb": torch.stack([it["A_emb"] for _, it in group]).to(dev)})   # :128-139
        tic = time.time()
        model.test()                                                  # :142 (result lands on the host: synchronises)
        toc = time.time()
        visuals = model.get_visuals()                                 # :146
        for k, (i, item) in enumerate(group):
            pred = D.to_unit_range(visuals[k:k + 1])                  # :151-152
            target = D.to_unit_range(GT[k:k + 1].numpy())
            rmse, psnr = D.rmse_psnr(pred, target)
            ssim = ssim_unit_range(pred, target)                      # :158-161
            path = D.save_triptych(item["LQ"], torch.from_numpy(visuals[k]), item["GT"], result_root, item["name"], i)
            r = results.setdefault(item["name"], dict(num=0, RMSE=[], SSIM=[], PSNR=[], time=[]))
            r["num"] += 1
            r["RMSE"].append(rmse)
            r["PSNR"].append(psnr)
            r["SSIM"].append(ssim)
            r["time"].append((toc - tic) / len(group))
            print(f" Testing {i}, {item['GT_path']}: RMSE={rmse:.5f}, SSIM={ssim:.4f}, PSNR={psnr:.3f}, {(toc - tic) / len(group):.3f} s/image "
                  f"(group of {len(group)}) -> {path}")

    with torch.no_grad():                                             # :109
        group = []
        for i in range(len(ds)):
            group.append((i, ds[i]))
            if len(group) == max(1, batch):
                flush(group)
                group = []
        if group:
            flush(group)
    for name, r in results.items():
        print(f"{name}: n={r['num']} RMSE={np.mean(r['RMSE']):.5f} SSIM={np.mean(r['SSIM']):.4f} PSNR={np.mean(r['PSNR']):.3f} "
              f"mean time {np.mean(r['time']):.2f} s")
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--flist", default=None)
    ap.add_argument("--result-root", default=None)
    ap.add_argument("--weights", default=None, help="a single noise-net state dict (.pth)")
    ap.add_argument("--pth-dir", default=None, help="checkpoint directory holding {iter}_NN.pth (test.pth_dir)")
    ap.add_argument("--iter", default=None, help="iteration label of the checkpoint (test.iter)")
    ap.add_argument("--use-ema", action="store_true", help="sample with lastest_NN_ema.pth (test.use_ema)")
    ap.add_argument("--artifact-type", nargs="*", default=None)
    ap.add_argument("--max-items", type=int, default=1000000)
    ap.add_argument("--T", type=int, default=-1)
    ap.add_argument("--batch", type=int, default=1, help="items restored per reverse process (testUM.py: 1); the "
                    "results do not depend on it, the throughput does")
    args = ap.parse_args()
    tmp = None
    if args.flist is None:
        tmp = tempfile.TemporaryDirectory()
        args.flist = D.make_synthetic_dataset(tmp.name)
        args.artifact_type = args.artifact_type or D.MODALITY_NAMES[:2]
        args.max_items = min(args.max_items, 6)
    root = args.result_root or os.path.join(tempfile.gettempdir(), "idiff_results")
    run(args.flist, root, args.artifact_type or [], args.weights, args.max_items, args.T, pth_dir=args.pth_dir,
        iter_label=args.iter, use_ema=args.use_ema, batch=args.batch)


if __name__ == "__main__":
    main()
