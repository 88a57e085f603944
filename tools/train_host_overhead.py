"""Diagnostic: host enqueue time vs total time of the D256 training step (is the step CPU-launch-bound?)."""
import sys, time, io, contextlib
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import torch
import bdvs_b200 as B
import bench
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    model = B.VunetAlter(**bench.D256)
model = model.to(dev).train()
opt = B.FusedAdam([{"params": list(getattr(model, n).parameters()), "name": n} for n in ("eu", "ed", "du", "dd")], lr=1e-4, betas=(0.5, 0.9), model=model)
Bt = 32
x = torch.rand(Bt, 30, 64, 64, device=dev) * 2 - 1
c = torch.rand(Bt, 3, 256, 256, device=dev) * 2 - 1
t = torch.rand(Bt, 3, 256, 256, device=dev) * 2 - 1
def step():
    imgs, means, logstds, _ = model(x, c)
    loss = torch.mean((imgs - t) ** 2) + 1e-3 * B.compute_kl_with_prior(means, logstds)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
for trial in range(3):
    t0 = time.perf_counter()
    for _ in range(10): step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3*(t1-t0)/10:.2f} ms/step, total {1e3*(t2-t0)/10:.2f} ms/step")
