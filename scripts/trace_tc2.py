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
buf = np.zeros((4, 16, 24), dtype=np.int64)
L.rnad_debug_trace.argtypes = [ctypes.c_void_p]
rc = L.rnad_debug_trace(buf.ctypes.data)
assert rc == 0, rc
t0 = buf[0, 0, 0]
# events of a head warp: 0 waiting for D2, 1 D2 ready, 2 action / next node known, 3 next observation published, 4 bookkeeping done
np.set_printoptions(linewidth=250)
for role, name in enumerate(["head side 0", "head side 1"]):
    print(name)
    for t in range(r.T):
        row = buf[role, t]
        print(f"  t={t}", " ".join(f"{(x - t0) if x else -1:6d}" for x in row[:5]))

# stream items 16..55, MMA warp (item % 2): absolute "relu seen", then cycles after it
mma = buf[2].reshape(-1)[:320].reshape(-1, 8)
print("item warp side c | relu seen (abs) | MMA2 issued | obs barrier passed | MMA1(+3) issued")
for j in range(40):
    base = mma[j, 0]
    print(f"{j+16:4d} {(j+16)%2:4d} {((j+16)//4)%2:4d} {(j+16)%4} | {base-t0:9d} | {mma[j,1]-base:6d} | {mma[j,3]-base:6d} | {mma[j,2]-base:6d}")
