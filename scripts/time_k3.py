"""Development aid: per-kernel times of the learner step at cfg2 for the library in $RNAD_B200_LIB."""
import json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "5", "--warmup", "3", "--cpu-budget", "0", "--fp32-steps", "0",
                      "--sustained-s", "0", "--learner-steps", "50"], capture_output=True, text=True)
d = json.loads(out.stdout.strip().splitlines()[-1])
l = d["learner"]
print(os.path.basename(os.environ.get("RNAD_B200_LIB", "default")), "free-running ms", round(l["free_running"]["ms_per_update"], 4),
      {k: round(x["ms"], 4) for k, x in l["roofline"]["kernels"].items()}, flush=True)
