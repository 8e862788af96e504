"""Compile oracle/hgr_oracle.c into oracle/libhgr_oracle.so (gcc only).  TEST INFRASTRUCTURE ONLY.

The reference is pure Python: there is no C/C++ reference source to compile into
``oracle/_ref`` -- the real reference is instead imported in the build container by
``tests/golden/make_golden.py`` and its outputs are committed as fixtures.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hgr_oracle.c")
LIB = os.path.join(HERE, "libhgr_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
