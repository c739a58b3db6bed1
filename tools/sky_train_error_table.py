import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import test_gpu_sky_train as T
from oracle import cases, ucnerf_oracle as O
from ucnerf_b200 import gemm
from ucnerf_b200.sky_train import sky_render_rays
heads = cases.make_heads(seed=3)
n, S = 200, 120
b = O.synthetic_rays(n, seed=11)
far = b["far"].reshape(-1, 1)
ray_batch = torch.cat([b["origins"], b["directions"], far, torch.full_like(far, float(far[0]) * 1.5), b["cam_dirs"]], -1).cuda()
target = torch.rand((n, 3), generator=torch.Generator().manual_seed(0)).cuda()
def grads(fn, net, rb, tgt):
    out = fn(net, rb)
    ((out - tgt) ** 2).sum().backward()
    return {k: p.grad.clone() for k, p in net.named_parameters()}, out
net = T._Sky(heads)
g_tc, o_tc = grads(lambda nn_, rb: sky_render_rays(rb, nn_, N_samples=S)["rgb_map"], net, ray_batch, target)
net64 = T._Sky(heads, torch.float64)
g_64, o_64 = grads(lambda nn_, rb: T._torch_render(nn_, rb, S), net64, ray_batch.double(), target.double())
net32 = T._Sky(heads)
g_32, o_32 = grads(lambda nn_, rb: T._torch_render(nn_, rb, S), net32, ray_batch, target)
# torch path but through tc_linear-free segments with fp32: same as net32. Also: fp32 torch with TF32-free matmul on segments
print("fwd err tc", float((o_tc.double()-o_64).abs().max()), "fp32", float((o_32.double()-o_64).abs().max()), "scale", float(o_64.abs().max()))
for k in g_64:
    s = float(g_64[k].abs().max())
    print(f"{k:28s} max|g| {s:10.3e}  tc {float((g_tc[k].double()-g_64[k]).abs().max())/s:9.2e}  fp32 {float((g_32[k].double()-g_64[k]).abs().max())/s:9.2e}")
# isolated TN check with N2 = 3
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn((24000, 256), device="cuda", generator=g); Bm = torch.randn((24000, 3), device="cuda", generator=g) * 5
ref = A.double().T @ Bm.double()
got = gemm.gemm_tn(A, Bm)
print("tn N2=3 rel", float((got.double()-ref).abs().max()/ref.abs().max()))
