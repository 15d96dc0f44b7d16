"""ORACLE build recipe (test infrastructure): compiles the C restatement into oracle/_build/libbp_oracle.so.

The real reference cannot be compiled here (pure Rust, no cargo/rustc in the image, crates not
vendored -- SURVEY.md 8c), so there is no oracle/_ref; this C port is the CPU baseline ("port")."""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "bp_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libbp_oracle.so")
# x86-64-v3 (AVX2/BMI2/ADX-era) so that the .so built here also runs on the GPU box's host CPU
CFLAGS = ["-O3", "-march=x86-64-v3", "-fPIC", "-shared", "-std=gnu11", "-Wall", "-Wno-unused-function"]


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    h = hashlib.sha256()
    for f in ("bp_oracle.c", "curve.h", "scalar.h"):
        h.update(open(os.path.join(HERE, "c", f), "rb").read())
    h.update(" ".join(CFLAGS).encode())
    stamp = OUT + ".stamp"
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return OUT
    r = subprocess.run(["gcc"] + CFLAGS + [SRC, "-o", OUT], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n" + r.stdout + r.stderr)
    open(stamp, "w").write(h.hexdigest())
    return OUT


if __name__ == "__main__":
    print(build(force=True))
