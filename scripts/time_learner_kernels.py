"""Development aid: CUDA-event timings of the fused learner forward / backward kernels at cfg2 size."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO]
import torch
import learn.fused as fused
from nn.net import MLP
a, T, B = int(os.environ.get("A", 3)), 8, 65536
dev = torch.device("cuda")
nets = [MLP(a, 256, device=dev) for _ in range(4)]
obs = torch.rand(T, B, 2, a, a, device=dev)
d_logit = torch.randn(T, B, a, device=dev) / (T * B); d_v = torch.randn(T, B, device=dev) / (T * B)
fl = fused.FusedLearner(nets[0])
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / n * 1e3
print("forward  %.1f us" % timeit(lambda: fl.forward(obs, *nets)))
print("backward %.1f us" % timeit(lambda: fl.backward(obs, nets[0], d_logit, d_v)))
d_logit_u = torch.randn(T, B, a, device=dev); d_v_u = torch.randn(T, B, device=dev)
print("backward_split %.1f us" % timeit(lambda: fl.backward_split(obs, nets[0], d_logit_u, d_v_u)))
