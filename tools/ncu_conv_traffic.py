"""DRAM bytes of the convolution launches of ONE forward -> profiles/conv_dram_traffic.json (bench.py's
`roofline.traffic`).

    IDIFF_PROFILE_STEPS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:'conv_gemm_kernel|conv3_rowpair_kernel' --csv --log-file gpurun_out/conv_dram.csv python tools/profile_forward.py
    python tools/ncu_conv_traffic.py gpurun_out/conv_dram.csv "<what the capture was taken on>"

The JSON carries the launch list (layer labels, in order) of the plan the capture was taken with; bench.py only reports
the figure while its own launch list is identical, so a changed kernel set cannot silently keep a stale number."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    path, source = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ii, ki, mi, vi, ui = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_launch = {}
    for r in rows[1:]:
        if r[mi].startswith("dram__bytes") and ("conv_gemm_kernel" in r[ki] or "conv3_rowpair_kernel" in r[ki]):
            per_launch.setdefault(int(r[ii]), [r[ki], 0.0])[1] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    import torch
    from instancediff_b200 import ConditionalUNet
    B, RES = int(os.environ.get("IDIFF_PROFILE_B", "32")), int(os.environ.get("IDIFF_PROFILE_RES", "256"))
    net = ConditionalUNet(device=torch.device("cuda:0"), seed=1)
    plan = net._plan(B, RES, RES, True)
    labels = [label for kind, _, label in plan.op_info if kind == "conv_gemm"]
    steps = int(os.environ.get("IDIFF_PROFILE_STEPS", "1"))
    launches = [per_launch[k] for k in sorted(per_launch)]
    assert len(launches) == steps * len(labels), (len(launches), steps, len(labels))
    first = launches[:len(labels)]                     # the first forward of the capture
    out = {"labels": labels, "bytes_total": sum(b for _, b in first), "bytes_per_launch": [b for _, b in first],
           "kernels": [k[:60] for k, _ in first], "source": source, "batch": B, "res": RES,
           "metric": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu, cold-cache serialised replay)"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "conv_dram_traffic.json"), "w"), indent=1)
    print(f"{len(labels)} conv launches, {out['bytes_total'] / 1e9:.2f} GB DRAM traffic per forward")


if __name__ == "__main__":
    main()
