import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from instancediff_b200 import ConditionalUNet, IRSDE, sample_sharded
from oracle.unet_oracle import make_oracle_unet
oracle = make_oracle_unet(seed=1).cuda()
net = ConditionalUNet(device="cuda"); net.load_state_dict(oracle.state_dict())
B,H,W,T = 4,32,32,12
g = torch.Generator().manual_seed(9)
mu = (torch.rand(B, 1, H, W, generator=g) * 2 - 1).cuda()
x = mu + 0.4 * torch.randn(B, 1, H, W, generator=g).cuda()
ctx = torch.nn.functional.normalize(torch.randn(B, 1, 512, generator=g), dim=-1).cuda()
def run(world, use_graph, T=T):
    outs=[]
    for rank in range(world):
        sde = IRSDE(0.4, T=100, schedule="cosine", eps=0.01, device=torch.device("cuda"))
        sde.set_model(net); sde.use_cuda_graph = use_graph
        x0,(lo,hi) = sample_sharded(sde, mu.cpu(), ctx.cpu(), rank, world, seed=5, T=T)
        outs.append(x0)
    return torch.cat(outs)
for TT in (1, 2, 12):
    fe = run(1, False, TT)
    for world, gr in [(1,True),(2,False),(2,True),(4,False)]:
        o = run(world, gr, TT)
        d = (fe-o).abs().amax(dim=(1,2,3))
        print(f"T={TT} world={world} graph={gr} equal={torch.equal(fe,o)} per-sample maxdiff={[f'{v:.1e}' for v in d.tolist()]}")
# layerwise: B=4 vs the two halves
o4 = net(x, mu, 37.0, image_context=ctx).clone()
p4 = net._plans[(4,H,W,True)]
acts4 = {k: v.t.clone() for k, v in p4.named.items()}
for half in (0, 1):
    sl = slice(2*half, 2*half+2)
    o2 = net(x[sl].contiguous(), mu[sl].contiguous(), 37.0, image_context=ctx[sl].contiguous()).clone()
    p2 = net._plans[(2,H,W,True)]
    torch.cuda.synchronize()
    bad = [(k, (acts4[k][sl].float()-v.t.float()).abs().max().item()) for k, v in p2.named.items() if not torch.equal(acts4[k][sl], v.t)]
    print("half", half, "first mismatches:", bad[:4], "out equal", torch.equal(o4[sl], o2))
