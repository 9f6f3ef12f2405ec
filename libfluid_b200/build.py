"""Builds libfluid_b200/_lib/liblfk.so (the C-ABI library of include/lfk.h) with nvcc for sm_100a.

In-tree build: the .so travels to the GPU box with the repository snapshot.  `python -m libfluid_b200.build`.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "liblfk.so")
SOURCES = ["lfk_api.cu", "particles.cu", "g2p.cu", "p2g.cu", "p2g_march.cu", "pressure.cu", "mg.cu", "exchange.cu", "transfer.cu", "aux.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# --fmad=false: the reference's CPU build does no FMA contraction; particle motion / classification must be
# bit-exact and every kernel there is bandwidth bound; FMAD_ON lists the exceptions.
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"),
          "-I" + CSRC]
# fused multiply-add only where the summation order already differs from the reference (tolerance-checked)
FMAD_ON = {"p2g_march.cu", "mg.cu", "g2p.cu"}
PER_FILE = {src: (["--fmad=true"] if src in FMAD_ON else ["--fmad=false"]) for src in SOURCES}


def _stamp():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ["../../include/lfk.h"]:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    h.update(" ".join(COMMON + ARCH).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp_file = os.path.join(OUT_DIR, "liblfk.stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [NVCC] + ARCH + COMMON + PER_FILE.get(src, []) + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s ----\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
