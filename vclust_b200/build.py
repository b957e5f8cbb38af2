"""Builds vclust_b200/libvclust_b200.so (sm_100a, in-tree) with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libvclust_b200.so"
SOURCES = ["vb_api.cu", "dev_genomes.cu", "prefilter.cu", "align.cu", "shard.cu", "host_fasta.cpp", "host_format.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--fmad=false",
]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + [HERE.parent / "include" / "vclust_b200.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj_dir = HERE / "build"
    obj_dir.mkdir(exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = obj_dir / (src + ".o")
        objs.append(str(obj))
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("VB_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-shared", "-o", str(LIB)] + objs + ["-lz", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
