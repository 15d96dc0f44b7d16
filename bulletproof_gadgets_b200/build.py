"""Builds libbpg.so (CUDA kernels + C ABI + host protocol driver) in-tree for sm_100a.

    python -m bulletproof_gadgets_b200.build [--force]

nvcc cross-compiles without a GPU.  Objects go to build/ (git-ignored); the .so is written next to
this file so that it travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbpg.so")
OBJDIR = os.path.join(ROOT, "build", "obj")
BINDIR = os.path.join(HERE, "bin")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-march=x86-64-v3,-Wall,-Wno-unused-function,-Wno-unknown-pragmas", "--expt-relaxed-constexpr",
    "-I", CSRC, "-I", os.path.join(ROOT, "include"),
] + os.environ.get("BPG_EXTRA_NVCC_FLAGS", "").split()


def sources():
    out = []
    for dirpath, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".cu", ".cpp")):
                out.append(os.path.join(dirpath, f))
    return sorted(out)


def _stamp():
    """Content hash of the sources and flags, independent of where the tree lives (the GPU box runs a copy)."""
    h = hashlib.sha256()
    for dirpath, dirs, files in os.walk(CSRC):
        dirs.sort()
        for f in sorted(files):
            p = os.path.join(dirpath, f)
            h.update(os.path.relpath(p, ROOT).encode())
            h.update(open(p, "rb").read())
    h.update(open(os.path.join(ROOT, "include", "bpg.h"), "rb").read())
    for name in ("prover.cpp", "verifier.cpp"):
        h.update(open(os.path.join(HERE, "cli", name), "rb").read())
    h.update(open(os.path.join(ROOT, "tools", "imad_peak.cu"), "rb").read())
    h.update(" ".join(x.replace(ROOT, ".") for x in NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OBJDIR, os.path.relpath(src, CSRC).replace(os.sep, "_") + ".o")
    if src.endswith("_native.cpp"):   # host-only translation units with SIMD intrinsics: straight to g++
        cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-march=x86-64-v3", "-Wall", "-c", src, "-o", obj]
    else:
        cmd = ["nvcc"] + NVCC_FLAGS + ["-x", "cu", "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build_lib(force=False, verbose=False):
    stamp_file = OUT + ".stamp"
    stamp = _stamp()

    def fresh():
        return os.path.exists(OUT) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp

    if not force and fresh():
        return OUT
    os.makedirs(OBJDIR, exist_ok=True)
    import fcntl
    with open(os.path.join(OBJDIR, ".lock"), "w") as lock:      # several ranks may import at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and fresh():
            return OUT
        return _build_locked(stamp_file, stamp, verbose)


def _build_locked(stamp_file, stamp, verbose):
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, err in results:
            if err.strip():
                print(err, file=sys.stderr)
    tmp = OUT + ".tmp"
    cmd = ["nvcc", "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, OUT)
    # process-level drop-ins (`prover <stem>`, `verifier <stem>`), linked against the library with an $ORIGIN rpath
    os.makedirs(BINDIR, exist_ok=True)
    for name in ("prover", "verifier"):
        cmd = ["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "cli", name + ".cpp"),
               "-o", os.path.join(BINDIR, name), "-L", HERE, "-l:libbpg.so", "-Wl,-rpath,$ORIGIN/.."]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("cli build failed:\n%s\n%s" % (r.stdout, r.stderr))
    # the issue-rate microbenchmark bench.py takes its roofline denominator from (tools/imad_peak.cu)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", os.path.join(ROOT, "tools", "imad_peak.cu"),
           "-o", os.path.join(BINDIR, "imad_peak")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("imad_peak build failed:\n%s\n%s" % (r.stdout, r.stderr))
    open(stamp_file, "w").write(stamp)
    return OUT


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
