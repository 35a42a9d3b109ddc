"""Build the CPU oracle (gcc).  Test infrastructure only -- see rz_oracle.c header."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "librz_oracle.so")
SRC = os.path.join(HERE, "rz_oracle.c")


def build(force=False):
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", SO, SRC])
    return SO


if __name__ == "__main__":
    print(build(force=True))
