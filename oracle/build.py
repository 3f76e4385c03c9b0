"""Build recipes for the oracle (test infrastructure).

  build_oracle(): gcc psamask_oracle.c -> oracle/liboracle_psamask.so
  build_ref():    the reference's own lib/psa/src/cpu/{operator,psamask}.cpp compiled *where they lie* under
                  /root/reference (nothing is copied) into oracle/_ref/psamask_ref_cpu.so with
                  torch.utils.cpp_extension (the files include <torch/torch.h>, so torch's headers are needed;
                  no other external dependency, no build system of the reference is run).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("SEMSEG_REFERENCE", "/root/reference")


def build_oracle():
    out = os.path.join(HERE, "liboracle_psamask.so")
    src = os.path.join(HERE, "psamask_oracle.c")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", out, src])
    return out


def build_ref(verbose=False):
    """Returns the path of the built extension, or None when /root/reference is absent (GPU box)."""
    cpu_dir = os.path.join(REFERENCE, "lib", "psa", "src", "cpu")
    if not os.path.isdir(cpu_dir):
        return None
    ref_dir = os.path.join(HERE, "_ref")
    os.makedirs(ref_dir, exist_ok=True)
    for f in os.listdir(ref_dir):
        if f.startswith("psamask_ref_cpu") and f.endswith(".so"):
            return os.path.join(ref_dir, f)
    from torch.utils.cpp_extension import load
    load(name="psamask_ref_cpu",
         sources=[os.path.join(cpu_dir, "operator.cpp"), os.path.join(cpu_dir, "psamask.cpp")],
         build_directory=ref_dir, verbose=verbose)
    for f in os.listdir(ref_dir):
        if f.startswith("psamask_ref_cpu") and f.endswith(".so"):
            return os.path.join(ref_dir, f)
    raise RuntimeError("reference extension did not produce a .so in %s" % ref_dir)


if __name__ == "__main__":
    print(build_oracle())
    print(build_ref(verbose="-v" in sys.argv))
