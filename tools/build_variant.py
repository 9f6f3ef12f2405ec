#!/usr/bin/env python
"""Builds a compile-time variant of liblfk.so for A/B timing (tools/variant_sweep.py --lib ...).

  python tools/build_variant.py <name> <file.cu> -DCT_THREADS=384 ...

Recompiles ONE source of libfluid_b200/csrc with the extra nvcc flags, links it with the production objects of the
other sources (libfluid_b200/_lib/*.o, built by libfluid_b200.build) and writes libfluid_b200/_lib/variants/liblfk_<name>.so.
The production library is not touched.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libfluid_b200 import build as B  # noqa: E402


def main():
    name, src, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
    B.build()
    vdir = os.path.join(B.OUT_DIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    obj = os.path.join(vdir, "%s_%s" % (name, src.replace(".cu", ".o")))
    cmd = [B.NVCC] + B.ARCH + B.COMMON + B.PER_FILE.get(src, []) + extra + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, src), "-o", obj]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        sys.stderr.write(out.stdout)
        raise SystemExit(1)
    lines = out.stdout.splitlines()
    for i, l in enumerate(lines):  # resource usage of the kernels named in LFK_SHOW (comma separated substrings)
        if "Compiling entry function" in l and any(k in l for k in os.environ.get("LFK_SHOW", "k_correct_tile,k_p2g_march").split(",")):
            print(l.split("'")[1][:60], "|", " ".join(x.strip() for x in lines[i + 1:i + 4] if "registers" in x or "spill" in x))
    objs = [obj if s == src else os.path.join(B.OUT_DIR, s.replace(".cu", ".o")) for s in B.SOURCES]
    lib = os.path.join(vdir, "liblfk_%s.so" % name)
    subprocess.check_call([B.NVCC] + B.ARCH + ["-shared", "-o", lib] + objs + ["-lcudart", "-ldl"])
    print(lib)


if __name__ == "__main__":
    main()
