import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from instancediff_b200 import ops
from gpu_util import rand_act
g = torch.Generator().manual_seed(3)
for (B,H,W,Cc) in [(4,32,32,64),(4,16,16,64),(4,64,64,64),(4,32,32,128)]:
    x = rand_act(B,H,W,Cc,g); xf = x.float()
    mean, var = xf.mean(-1, keepdim=True), xf.var(-1, unbiased=False, keepdim=True)
    stats = torch.cat([mean, torch.rsqrt(var+1e-5)], -1).reshape(-1,2).contiguous()
    g_pre = (torch.rand(Cc, generator=g)+0.5).cuda(); wqkv=((torch.rand(384,Cc,generator=g)*2-1)*(2.0/Cc**0.5)).cuda()
    w_out=((torch.rand(Cc,128,generator=g)*2-1)/128**0.5).cuda(); b_out=(torch.rand(Cc,generator=g)-0.5).cuda(); g_out=(torch.rand(Cc,generator=g)+0.5).cuda()
    a = ops.linattn_fused(x, stats, wqkv, g_pre, w_out, b_out, g_out)
    a2 = ops.linattn_fused(x, stats, wqkv, g_pre, w_out, b_out, g_out)
    h = ops.linattn_fused(x[:2].contiguous(), stats[:2*H*W].contiguous(), wqkv, g_pre, w_out, b_out, g_out)
    torch.cuda.synchronize()
    print("linattn", (B,H,W,Cc), "repeat equal:", torch.equal(a,a2), "shard equal:", torch.equal(a[:2],h), (a[:2].float()-h.float()).abs().max().item())
for (B,H,W) in [(4,32,32),(4,64,64)]:
    x, mu = torch.randn(B,1,H,W,generator=g).cuda(), torch.randn(B,1,H,W,generator=g).cuda()
    w=((torch.rand(64,2,7,7,generator=g)*2-1)/10).cuda(); b=(torch.rand(64,generator=g)-0.5).cuda()
    a = ops.stem_conv7_tc(x,mu,w,b); a2 = ops.stem_conv7_tc(x,mu,w,b); h = ops.stem_conv7_tc(x[:2].contiguous(),mu[:2].contiguous(),w,b)
    torch.cuda.synchronize()
    print("stem", (B,H,W), torch.equal(a,a2), torch.equal(a[:2],h))
