"""Development check: the rollout's recorded logits / values against the learner forward kernel's on the same rows."""
import os, sys, random
import numpy as np, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
for p in (os.path.join(REPO, "r-nad_b200"), REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
from test_gpu_learner_step import seeded_tree, fresh_trial, flat

tree = seeded_tree(ragged=True, depth=4)
trial = fresh_trial(tree, 4096, "check_reuse", "eager")
torch.manual_seed(500)
step = trial._step_engine_for()
calls = dict(step._calls(reuse=True))
import _b200
_b200.lib().rnad_step_control(step.ctrl.data_ptr(), 777, 0.5, _b200.stream())
calls["rollout"]()
full = dict(step._calls(reuse=False))
full["pack"](); full["forward"]()
torch.cuda.synchronize()
valid = step.arena["indices"] != 0
lg_r, lg_f = step.logits, step.fwd["logit"]
v_r, v_f = step.arena["values"], step.fwd["v"].squeeze(-1)
print("valid slots", int(valid.sum()), "of", valid.numel())
print("logit max |diff| on valid slots", float((lg_r - lg_f)[valid].abs().max()), " all slots", float((lg_r - lg_f).abs().max()))
print("value max |diff| on valid slots", float((v_r - v_f)[valid].abs().max()), " all slots", float((v_r - v_f).abs().max()))
pol = step.arena["policy"]
pi_f = step.fwd["pi"]
print("policy (fast softmax) vs forward pi, valid:", float((pol - pi_f)[valid].abs().max()))
# now the two target computations
full["targets"]()
torch.cuda.synchronize()
dl_full, dv_full, ls_full = step.d_logit.clone(), step.d_v.clone(), step.loss_sums.clone()
keep = {k: step.fwd[k].clone() for k in ("v_target", "log_pi_reg", "log_pi_reg_")}
for k in keep: step.fwd[k].fill_(float("nan"))
calls["pack"](); calls["forward"]()
torch.cuda.synchronize()
print("target == learner weights here, so v_target should equal v:")
print("  full   v_target vs v:", float((keep["v_target"].squeeze(-1) - v_f)[valid].abs().max()))
print("  others v_target vs v:", float((step.fwd["v_target"].squeeze(-1) - v_f)[valid].abs().max()))
print("  samples v", v_f[0, :4].tolist(), "full", keep["v_target"][0, :4, 0].tolist(), "others", step.fwd["v_target"][0, :4, 0].tolist())
for k in keep:
    d = (step.fwd[k] - keep[k])
    print(k, "others-only vs full: max |diff|", float(d[valid].abs().max()), "nan count", int(torch.isnan(step.fwd[k]).sum()))
# targets with the full io but others-only forward outputs
full["targets"]()
torch.cuda.synchronize()
print("full io after others-only forward: d_logit diff", float((step.d_logit - dl_full).abs().max()))
calls["targets"]()
torch.cuda.synchronize()
print("d_logit max |diff|", float((step.d_logit - dl_full).abs().max()), "scale", float(dl_full.abs().max()))
print("d_v max |diff|", float((step.d_v - dv_full).abs().max()), "scale", float(dv_full.abs().max()))
print("loss sums", ls_full.tolist(), step.loss_sums.tolist())
