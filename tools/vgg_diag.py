"""Diagnostic: per-term gradient error of the CUDA vgg_loss against the bf16 model of the reference graph, and how many
feature elements differ between the two forward evaluations."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import vgg_oracle as V, vunet_oracle as O
import bdvs_b200 as B
from torchvision.models import vgg19

S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
vgg = vgg19(weights=None).eval()
sd = V.det_vgg_state_dict()
vgg.features.load_state_dict(sd)
w = [1.0] * 6
m = B.PerceptualVGG(vgg, w).cuda().eval()
target = O.det_uniform("vgg_target", (2, 3, S, S))
pred0 = O.det_uniform("vgg_pred", (2, 3, S, S)) * 0.9
rel = lambda a, b: float((a.float().cpu() - b).norm() / b.norm())
feats = m(pred0.cuda())
mf = V.perceptual_features(sd, pred0, bf16_model="wn")
for k in V.TAP_NAMES[1:]:
    a = feats[k].float().cpu()
    b = mf[k]
    print(k, "feature rel", rel(a, b), "fraction of elements that differ", float((a != b).float().mean()))
for name in V.TAP_NAMES:
    p = pred0.clone().cuda().requires_grad_(True)
    l = B.vgg_loss(m, target.cuda(), p)[name].sum()
    l.backward()
    outs = {}
    for mode in ("wn", True, False):
        q = pred0.clone().requires_grad_(True)
        V.vgg_loss(sd, target, q, w, bf16_model=mode)[name].sum().backward()
        outs[mode] = q.grad
    print(f"{name}: cuda vs wn-model {rel(p.grad, outs['wn']):.3e}  vs bf16(w)-model {rel(p.grad, outs[True]):.3e}  "
          f"vs fp32 {rel(p.grad, outs[False]):.3e} | wn-model vs fp32 {rel(outs['wn'], outs[False]):.3e}")
