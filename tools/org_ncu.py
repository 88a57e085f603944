"""One eager ``VunetOrg.transfer`` step (H256 geometry, 64 frames, one stream) between cudaProfilerStart/Stop for an ncu
launch list (tools/r02_ncu_org.sh)."""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
os.environ["VUB_FWD_STREAMS"] = "1"
import torch
import vunet_oracle as O
import bdvs_b200 as B
from bdvs_b200 import ops
from bench import H256

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
with contextlib.redirect_stdout(io.StringIO()):
    m = B.VunetOrg(**dict(H256))
m.load_state_dict(O.det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}))
m = m.cuda().eval()
x = torch.rand(frames, 3, 256, 256, device="cuda") * 2 - 1
c = torch.rand(frames, 3, 256, 256, device="cuda") * 2 - 1
with torch.no_grad():
    for _ in range(2):
        m.transfer(x, c)
    torch.cuda.synchronize()
    before = ops.LAUNCHES
    torch.cuda.profiler.start()
    out = m.transfer(x, c)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("launches", ops.LAUNCHES - before, "finite", bool(torch.isfinite(out).all()))
