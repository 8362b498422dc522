"""Build the oracle's C restatement (gcc).  Test infrastructure only.

``python -m oracle.build`` -> oracle/liboracle_coder.so (git-ignored, travels with gpurun).
There is no compilable reference source (the reference is pure Python on TF 1.13 and the
coder lives in an absent wheel), so there is no ``oracle/_ref`` target; see DESIGN.md.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "c", "oracle_coder.c")
    out = os.path.join(HERE, "liboracle_coder.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        cmd = ["gcc", "-O2", "-shared", "-fPIC", "-o", out, src, "-lm"]
        subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
