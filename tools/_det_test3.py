import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from instancediff_b200 import ConditionalUNet, IRSDE, sample_sharded
net = ConditionalUNet(device="cuda", seed=1)
B,H,W,T = 4,32,32,12
g = torch.Generator().manual_seed(9)
mu=(torch.rand(B,1,H,W,generator=g)*2-1); ctx = torch.nn.functional.normalize(torch.randn(B,1,512,generator=g),dim=-1)
def run(world, use_graph, T=T):
    outs=[]
    for rank in range(world):
        sde = IRSDE(0.4, T=100, schedule="cosine", eps=0.01, device=torch.device("cuda"))
        sde.set_model(net); sde.use_cuda_graph = use_graph
        x0,(lo,hi) = sample_sharded(sde, mu, ctx, rank, world, seed=5, T=T)
        outs.append(x0)
    return torch.cat(outs)
for TT in (1, 2, 12):
    fe = run(1, False, TT)
    for world, gr in [(1,True),(2,False),(2,True),(4,False),(4,True)]:
        o = run(world, gr, TT)
        d = (fe-o).abs().amax(dim=(1,2,3))
        print(f"T={TT} world={world} graph={gr} equal={torch.equal(fe,o)} per-sample maxdiff={[f'{v:.1e}' for v in d.tolist()]}")
