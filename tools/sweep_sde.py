"""Times the fused SDE step the way the sampling loop runs it: a captured graph of launches rotating over enough
buffer sets that the inputs come from HBM, not from the 126 MB L2.  (The launch variants compared with it were
selected by an environment variable read in csrc/sde_step.cu; the table of results is in that file.)"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from instancediff_b200 import IRSDE, _lib  # noqa: E402

dev = torch.device("cuda:0")
sde = IRSDE(0.4, T=100, schedule="cosine", eps=0.01, device=dev)
row = sde._coef_table(dev)[50]
L = _lib.lib()
out = []
for B in (32, 64, 256):
    n = B * 256 * 256
    nsets = max(2, -(-(2 * 126 * 1024 * 1024) // (16 * n)))           # one cycle over the sets exceeds 2 x L2
    g = torch.Generator(device="cuda").manual_seed(B)
    sets = [tuple(torch.randn(n, device=dev, generator=g) for _ in range(3)) for _ in range(nsets)]
    s = torch.cuda.current_stream(dev).cuda_stream

    def launch(stream, i):
        x, e, mu = sets[i % nsets]
        _lib.check(L.idiff_sde_step(x.data_ptr(), x.data_ptr(), e.data_ptr(), mu.data_ptr(), None, row.data_ptr(), 0, 1, 7, 0, n,
                                    stream), "sde_step")
    launch(s, 0)
    torch.cuda.synchronize()
    digest = hashlib.sha1(sets[0][0].cpu().numpy().tobytes()).hexdigest()[:10]
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    gr = torch.cuda.CUDAGraph()
    nl = 2 * nsets
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            for i in range(nl):
                launch(side.cuda_stream, i)
    torch.cuda.current_stream(dev).wait_stream(side)
    for _ in range(3):
        gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    us_graph = e0.elapsed_time(e1) / (10 * nl) * 1e3
    out.append(f"B={B} ({nsets} sets): graph {us_graph:.2f} us = {16.0 * n / us_graph / 1e3:.0f} GB/s, sha {digest}")
    del gr, sets
print(" | ".join(out))
