"""Development aid: cycle trace of CTA 0 of rollout_tc2_kernel (library built with -DRNAD_TRACE)."""
import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO]
import numpy as np, torch
import bench, _b200
from nn.net import MLP

depth, a, c, batch = bench.CONFIGS["cfg2"]
tree = bench.make_tree(depth, a, c); tree.to(torch.device("cuda"))
net = MLP(a, 256, device=torch.device("cuda"))
r = bench.RolloutRunner(tree, net, batch, sys.argv[1] if len(sys.argv) > 1 else "tf32x2")
for _ in range(3):
    r.launch()
torch.cuda.synchronize()
L = _b200.lib()
buf = np.zeros((3, 16, 24), dtype=np.int64)
L.rnad_debug_trace.argtypes = [ctypes.c_void_p]
rc = L.rnad_debug_trace(buf.ctypes.data)
assert rc == 0, rc
t0 = buf[0, 0, 0]
np.set_printoptions(linewidth=250)
for role, name in enumerate(["mma", "head(w0)", "epi(w4)"]):
    print(name)
    for t in range(r.T):
        row = buf[role, t]
        print(f"  t={t}", " ".join(f"{(x - t0) if x else -1:6d}" for x in row[:20]))
