"""Build recipe for the oracle's C restatement (test infrastructure, not product).

    python oracle/build_oracle.py      ->  oracle/libeosvr_oracle.so

The reference is pure Python, so there is nothing to compile into oracle/_ref/; the
reference's own code is exercised by oracle/make_golden.py in the build container
(where /root/reference exists) and its outputs are committed under tests/golden/.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "eosvr_oracle.c")
OUT = os.path.join(HERE, "libeosvr_oracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-fopenmp",
           "-ffp-contract=off", "-fno-fast-math", "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
