"""Compile the plain-C env oracle (oracle/gm_env.c -> oracle/libgm_oracle.so).  Test infrastructure only."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libgm_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "gm_env.c")
    if (not force) and os.path.exists(SO_PATH) and os.path.getmtime(SO_PATH) >= os.path.getmtime(src):
        return SO_PATH
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-o", SO_PATH, src])
    return SO_PATH


if __name__ == "__main__":
    print(build(force=True))
