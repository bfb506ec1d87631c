"""Build the plain-C oracle helpers (TEST INFRASTRUCTURE) into oracle/_build/ with gcc."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
SRC = os.path.join(HERE, "nlms_oracle.c")
LIB = os.path.join(OUT, "libnlms_oracle.so")


def build(force: bool = False) -> str:
    if not os.path.exists(SRC):
        raise FileNotFoundError(SRC)
    os.makedirs(OUT, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
