"""Puts the UNMODIFIED reference under baseline/_ref (git-ignored; it travels to the GPU box with the repo snapshot).

hszhao/semseg is not an installable package (no setup.py / pyproject.toml; `pip install /root/reference` fails with
"neither 'setup.py' nor 'pyproject.toml' found"), it is run from its source tree with PYTHONPATH=./ (tool/train.sh:8).
"Installing" it is therefore a verbatim copy of the directories its entry points import (model/, lib/, util/, tool/,
config/). Nothing here is tracked by git and nothing of it is imported by the product: only bench.py's reference arm
(baseline/run_reference.py, in a subprocess whose PYTHONPATH is baseline/_ref alone) and tools/run_reference_trainer.py use it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
DIRS = ("model", "lib", "util", "tool", "config")


def install(src="/root/reference", force=False):
    """Returns DEST when the reference is available there (copied now or earlier), else None."""
    if os.path.isdir(os.path.join(DEST, "model")) and not force:
        return DEST
    if not os.path.isdir(src):
        return None
    os.makedirs(DEST, exist_ok=True)
    for d in DIRS:
        s, t = os.path.join(src, d), os.path.join(DEST, d)
        if os.path.isdir(s):
            if os.path.isdir(t):
                shutil.rmtree(t)
            shutil.copytree(s, t)
    for f in ("LICENSE", "README.md"):
        if os.path.exists(os.path.join(src, f)):
            shutil.copy(os.path.join(src, f), os.path.join(DEST, f))
    return DEST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
