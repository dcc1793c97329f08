"""gcc recipe for the C oracle (test infrastructure).  No FMA contraction, no fast-math."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lwp_oracle.c")
LIB = os.path.join(HERE, "liblwp_oracle.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fno-unsafe-math-optimizations",
           "-mno-fma", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
