"""Development aid: cycle trace of CTA 0 of learner_bwd_f16_kernel (library built with -DRNAD_TRACE_BWD), cfg2 shape."""
import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO]
import numpy as np, torch
import _b200

from nn.net import MLP

dev = torch.device("cuda", 0)
a, T, B = 3, 8, 65536
torch.manual_seed(0)
net = MLP(a, 256, device=dev)

obs = torch.rand(T, B, 2, a, a, device=dev)
d_logit = torch.randn(T, B, a, device=dev) * 0.1
d_v = torch.randn(T, B, device=dev) * 0.1
L = _b200.lib()
ws = torch.empty(int(L.rnad_learner_mlp_workspace_bytes(a, 256)), dtype=torch.uint8, device=dev)
grads = torch.empty(2 * int(L.rnad_learner_param_count(a, 256)), device=dev)
w = _b200.mlp_weights(net, dev)
for _ in range(3):
    rc = L.rnad_learner_backward_split(_b200.ptr(obs), T, B, a, ctypes.byref(w), _b200.ptr(d_logit), _b200.ptr(d_v), _b200.ptr(grads),
                                       _b200.ptr(ws), _b200.stream())
    assert rc == 0, _b200.last_error() if hasattr(_b200, "last_error") else rc
torch.cuda.synchronize()
buf = np.zeros((5, 64, 8), dtype=np.int64)
L.rnad_debug_bwdh_trace.argtypes = [ctypes.c_void_p]
assert L.rnad_debug_bwdh_trace(buf.ctypes.data) == 0
t0 = buf[0, 0, 0]      # consumer group 0, stage 16, wait start
np.set_printoptions(linewidth=250)
print("stage (region = s & 3) | issuer of the region: loop top, hand-over passed, grads + next recompute issued | "
      "region 0 only - first consumer warp: wait start, H seen, stored, arrived at the hand-over")
for j in range(48):
    s = j + 16
    iss = " ".join(f"{x - t0:7d}" for x in buf[2, j, :3])
    con = " ".join(f"{x - t0:7d}" for x in buf[0, j, :4]) if s % 4 == 0 else ""
    print(f"{s:4d} ({s & 3}) | {iss} | {con}")
print("producer warp 0 per tile: loop top, buffer free, operands written")
for j in range(0, 64, 8):
    print(f"tile {(j + 16) // 8}: " + " ".join(f"{x - t0:7d}" for x in buf[4, j, :3]))
