import os, sys
REPO = "/root/repo"
sys.path[:0] = [os.path.join(REPO, "r-nad_b200"), REPO, os.path.join(REPO, "tests")]
import torch
import test_gpu_learner_step as T
tree = T.seeded_tree(ragged=True, depth=4)
batch = 4096
trials = {e: T.fresh_trial(tree, batch, f"pytest_step_{e}", e) for e in ("off", "graph")}
p_init = T.flat(trials["off"].net).clone()
for i in range(5):
    out = {}
    for e, trial in trials.items():
        torch.manual_seed(500 + i)
        ep = trial.learner_step(alpha=0.25 * i)
        out[e] = (T.flat(trial.net).clone(), trial.last_losses.clone(), ep.full("indices").clone(), ep.full("policy").clone())
    d = (out["graph"][0] - out["off"][0]).abs()
    moved = (out["off"][0] - p_init).norm()
    same_games = torch.equal(out["graph"][2], out["off"][2])
    k = int(d.argmax())
    print(i, "max diff %.3e at %d" % (d.max().item(), k), "n>3e-4:", int((d > 3e-4).sum()), "norm ratio %.4f" % ((out["graph"][0] - out["off"][0]).norm() / moved).item(),
          "same games", same_games, "policy maxdiff %.2e" % (out["graph"][3] - out["off"][3]).abs().max().item(), "losses", out["graph"][1].tolist(), out["off"][1].tolist())
