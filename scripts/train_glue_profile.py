"""torch.profiler view of one training step (configs[2] shape): which device kernels are NOT ours (ATen glue, Adam, dropout mask)"""
import os, sys, random, contextlib
from types import SimpleNamespace
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gfs-3dseg_gws_b200"))
from gfs3d.synthetic import randomize_bn_, synthetic_blocks
from model.capl import mpti_net_Point_GeoAsWeight_v2
dev = torch.device("cuda", 0)
CLASSES, BASE, G, Bt, N = 21, 15, 180, 32, 2048
args = SimpleNamespace(edgeconv_widths=[[64, 64]] * 3, dgcnn_mlp_widths=[512, 256], pc_in_dim=9, dgcnn_k=20, base_widths=[128, 64], output_dim=64, eval_weight=1.2)
torch.manual_seed(321)
with contextlib.redirect_stdout(sys.stderr):
    m = mpti_net_Point_GeoAsWeight_v2(classes=CLASSES, criterion=torch.nn.CrossEntropyLoss(ignore_index=255), args=args, base_num=BASE,
                                      gp=torch.randn(G, 192, generator=torch.Generator().manual_seed(7)).to(dev), energy=0.9)
randomize_bn_(m, seed=6)
m = m.to(dev).train()
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
x = synthetic_blocks(Bt, N, seed=7).to(dev)
y = torch.randint(0, BASE + 1, (Bt, N), generator=torch.Generator().manual_seed(1)).to(dev)
random.seed(1)
def step():
    opt.zero_grad(set_to_none=True)
    pred, loss = m(x=x, y=y)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA] or prof.key_averages()
rows = sorted(((e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if getattr(e, "device_time_total", 0) > 0 and not e.key.startswith("aten::") and not e.key.startswith("autograd::") and "Backward" not in e.key and not e.key.startswith("Optimizer")), key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
ours = sum(r[2] for r in rows if "gfs::" in r[0])
print(f"device kernels: total {tot:.2f} ms, gfs:: {ours:.2f} ms, others {tot - ours:.2f} ms")
for k, c, t in [r for r in rows if 'gfs::' not in r[0]][:40]:
    print(f"{t:8.3f} ms {c:4d}  {k[:150]}")
