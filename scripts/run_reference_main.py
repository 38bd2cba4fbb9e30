#!/usr/bin/env python
"""
Runs the reference's main.py UNMODIFIED (baseline/_ref/ref/main.py, a copy of /root/reference/main.py made by
baseline/install_reference.py) against this repository's packages: r-nad_b200/ is put first on sys.path, so the
script's `from environment.tree import Tree` / `from learn.rnad import RNaD` resolve to the API mirrors and every
rollout / learner step lands in the sm_100a kernels.  The script is executed with runpy from the file where it lies
(runpy.run_path does not put the script's own directory on sys.path, so the reference's packages next to it are not
importable).  Its schedule (4 etas x bounds [64] x delta_m [100], checkpoint every step) runs until the wall-clock
budget given on the command line ends the process from outside (`timeout`), or to completion.
"""
import os
import runpy
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "r-nad_b200"))
main_py = os.path.join(REPO, "baseline", "_ref", "ref", "main.py")
assert os.path.isfile(main_py), "baseline/_ref is not installed (python baseline/install_reference.py)"
import hashlib

print("main.py sha256", hashlib.sha256(open(main_py, "rb").read()).hexdigest(), flush=True)
runpy.run_path(main_py, run_name="__main__")
print("main.py finished", flush=True)
