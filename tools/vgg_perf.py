"""Per-kernel timing of the VGG19 perceptual loss (forward on target + prediction, backward to the prediction) at the
training resolution: python tools/vgg_perf.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import vgg_oracle as V
import bdvs_b200 as B
from bdvs_b200.profiler import KernelProfiler
from torchvision.models import vgg19

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
vgg = vgg19(weights=None).eval()
vgg.features.load_state_dict(V.det_vgg_state_dict())
m = B.PerceptualVGG(vgg, [1.0] * 6).cuda().eval()
target = torch.rand(bs, 3, 256, 256, device="cuda") * 2 - 1
pred = (torch.rand(bs, 3, 256, 256, device="cuda") * 2 - 1).requires_grad_(True)


def step():
    l = B.vgg_loss(m, target, pred)
    torch.stack([l[k] for k in l]).sum().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
gf = 48.4 * 3  # forward twice + data gradient, GFLOP per sample (conv1_1 .. conv5_2)
print(f"vgg_loss fwd(target) + fwd(pred) + bwd, B={bs}: {ms:.3f} ms/step = {bs / ms * 1e3:.0f} samples/s, "
      f"{gf * bs / ms:.0f} TFLOP/s")
with KernelProfiler() as kp:
    step()
prof = kp.summary()
tot = sum(r["ms"] for r in prof.values())
for name, r in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    tf = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else 0
    gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else 0
    print(f"{name:64s} n={r['launches']:2d} ms={r['ms']:.4f} share={r['ms'] / tot:.3f} TF={tf:7.1f} GB/s={gb:7.1f}")
print("instrumented total", tot)
