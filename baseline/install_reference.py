"""
Puts the UNMODIFIED reference where `bench.py --impl reference` runs it from:

    baseline/_ref/ref/        a copy of /root/reference (git-ignored, not gpurun-ignored: it travels to the GPU box)
    baseline/_ref/_standin/   pygambit.py, the stand-in for the absent third-party solver (baseline/pygambit_standin.py)

The reference is not a pip package (no setup.py / pyproject.toml), so `pip install --target baseline/_ref` has nothing
to build; a plain copy is the install.  Called by `__graft_entry__.build()` when /root/reference is mounted (the build
container); on the GPU box the prebuilt directory is used as it arrived.  Nothing under baseline/_ref is tracked.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
DEST = os.path.join(HERE, "_ref")


def install(force=False) -> bool:
    """True if baseline/_ref is (now) in place."""
    ref = os.path.join(DEST, "ref")
    if os.path.isdir(REFERENCE) and (force or not os.path.isfile(os.path.join(ref, "learn", "rnad.py"))):
        shutil.rmtree(DEST, ignore_errors=True)
        shutil.copytree(REFERENCE, ref, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.png"))
        os.makedirs(os.path.join(DEST, "_standin"), exist_ok=True)
    if os.path.isdir(ref):
        os.makedirs(os.path.join(DEST, "_standin"), exist_ok=True)
        shutil.copyfile(os.path.join(HERE, "pygambit_standin.py"), os.path.join(DEST, "_standin", "pygambit.py"))
    return os.path.isfile(os.path.join(ref, "learn", "rnad.py"))


if __name__ == "__main__":
    print("baseline/_ref installed" if install(force=True) else "reference not available")
