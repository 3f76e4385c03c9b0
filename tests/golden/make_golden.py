"""Generate the golden fixtures in tests/golden/ from the REAL reference (hszhao/semseg at /root/reference).

Run once in the build container (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py            # everything (minutes: three full networks on CPU)
    python tests/golden/make_golden.py sliding    # only the sliding-window inference fixtures
It copies the reference to a scratch dir (lib/psa JIT-builds in place), then runs `_ref_worker` in a
subprocess whose PYTHONPATH contains ONLY the reference copy, so `model.pspnet` / `lib.psa.functional` are the
reference's own modules (this repository ships packages with the same names).
Fixtures are small .npz/.json files: sampled logits, losses, gradient norms, weight checksums, state_dict keys.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SEMSEG_REFERENCE", "/root/reference")
SCRATCH = "/tmp/semseg_ref_copy"


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at %s" % REF)
    if not os.path.isdir(SCRATCH):
        shutil.copytree(REF, SCRATCH, ignore=shutil.ignore_patterns(".git"))
    env = dict(os.environ)
    env["PYTHONPATH"] = SCRATCH
    env["GOLDEN_OUT"] = HERE
    subprocess.check_call([sys.executable, os.path.join(HERE, "_ref_worker.py")] + sys.argv[1:], cwd=SCRATCH, env=env)


if __name__ == "__main__":
    main()
