"""In-tree build of the CUDA shared library (sm_100a only).

    python -m neural_radiance_caching_b200.build [--force]

Produces neural_radiance_caching_b200/libnrc_b200.so with explicit nvcc
(-gencode arch=compute_100a,code=sm_100a -lineinfo).  The .so is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# NRC_LIB_PATH: an alternative library file (debug / trace builds beside the product build)
LIB = os.environ.get("NRC_LIB_PATH") or os.path.join(HERE, "libnrc_b200.so")
OBJ_DIR = os.path.join(HERE, "build" if not os.environ.get("NRC_LIB_PATH") else "build_alt")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    # IEEE division / sqrt and no flush-to-zero: parity with the fp32 oracle.
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
] + os.environ.get("NRC_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DNRC_CHAIN_TRACE for tools/trace_chain.py


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    files.append(os.path.join(ROOT, "include", "nrc_b200.h"))
    files.append(os.path.join(ROOT, "include", "nrc_xla.h"))
    for f in files:
        h.update(os.path.relpath(f, ROOT).encode())   # repo-relative: the digest does not depend on the checkout path
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def built_digest():
    """Digest compiled into the existing library (the string nrc_build_digest() returns), or None.  Read from the file's
    bytes, not through dlopen: a library mapped here would stay mapped (same path) after a rebuild in this process,
    and the loader in _lib.py would then see the stale image."""
    if not os.path.exists(LIB):
        return None
    import re
    with open(LIB, "rb") as fh:
        m = re.search(rb"nrc-build-digest:([0-9a-f]{64})", fh.read())
    return m.group(1).decode() if m else None


def build(force=False, verbose=False):
    """Compile when the digest EMBEDDED in the existing .so differs from the sources' (the .so is git-ignored: a
    stamp file beside it could describe another checkout's binary)."""
    digest = _digest()
    if not force and built_digest() == digest:
        return LIB
    objs = []
    procs = []
    os.makedirs(OBJ_DIR, exist_ok=True)
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc_path(), *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", src, "-o", obj]
        if os.path.basename(src) == "encode.cu":   # exports nrc_build_digest()
            cmd.insert(1, f'-DNRC_BUILD_DIGEST="nrc-build-digest:{digest}"')
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [nvcc_path(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
