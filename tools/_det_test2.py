import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from instancediff_b200 import ConditionalUNet
torch.manual_seed(0)
net = ConditionalUNet(device="cuda", seed=1)
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = torch.Generator().manual_seed(9)
x = torch.randn(4,1,H,W,generator=g).cuda(); mu=(torch.rand(4,1,H,W,generator=g)*2-1).cuda()
ctx = torch.nn.functional.normalize(torch.randn(4,1,512,generator=g),dim=-1).cuda()
o4 = net(x, mu, 37.0, image_context=ctx).clone()
p4 = net._plans[(4,H,W,True)]
acts4 = {k: v.t.clone() for k, v in p4.named.items()}
o2 = net(x[:2].contiguous(), mu[:2].contiguous(), 37.0, image_context=ctx[:2].contiguous()).clone()
p2 = net._plans[(2,H,W,True)]
torch.cuda.synchronize()
for k, v in p2.named.items():
    a, b = acts4[k][:2], v.t
    print(f"{k:16s} equal={torch.equal(a,b)} maxdiff={(a.float()-b.float()).abs().max().item():.3e}")
print("out equal", torch.equal(o4[:2], o2), (o4[:2]-o2).abs().max().item())
o4b = net(x, mu, 37.0, image_context=ctx).clone()
print("repeat B=4 equal", torch.equal(o4, o4b))
