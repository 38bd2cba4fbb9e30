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
buf = np.zeros((5, 16, 24), dtype=np.int64)
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

# stream items 16..55.  MMA warp (= chunk index c): relu(item) seen, MMA2(item) + MMA1(item+3) issued.
# Epilogue warp 0: first layers of the item complete, relu written back.
mma = buf[2].reshape(-1)[:320].reshape(-1, 8)
print("item side c | MMA warp: relu seen, second layers issued, first layers of item+3 issued, those complete | "
      "epilogue warp 0: d1 seen, loads back, stores done, arrived   (absolute cycles)")
for j in range(40):
    i = j + 16
    print(f"{i:4d} {(i//4)%2:4d} {i%4} | {mma[j,0]-t0:8d} {mma[j,1]-t0:8d} {mma[j,2]-t0:8d} {mma[j,5]-t0:8d} | "
          f"{mma[j,3]-t0:8d} {mma[j,6]-t0:8d} {mma[j,7]-t0:8d} {mma[j,4]-t0:8d}")
epi = buf[3].reshape(-1)[:320].reshape(-1, 8)
print("item | relu written back by epilogue warp 0..7 (absolute cycles) | relu seen by the MMA warp")
for j in range(24):
    print(f"{j+16:4d} | " + " ".join(f"{x-t0:7d}" for x in epi[j]) + f" | {mma[j,0]-t0:8d}")
seen = buf[4].reshape(-1)[:320].reshape(-1, 8)
print("item | first layers seen by epilogue warp 0..7, then its time until the relu is written back")
for j in range(24):
    print(f"{j+16:4d} | " + " ".join(f"{x-t0:7d}" for x in seen[j]) + " | " + " ".join(f"{y-x:5d}" for x, y in zip(seen[j], epi[j])))
print("kernel entry, set-up done (weights packed), kernel end, relative to the first head event:",
      " ".join(str(int(x - t0)) for x in buf[0, 15, :3]))
