"""In-tree build of the CUDA extension (``libcsf_b200.so``) for sm_100a.

    python -m cyclistsocialforce_b200.build

nvcc cross-compiles without a GPU.  The .so stays next to this file (git-ignored)
so that it travels with the source tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcsf_b200.so")
SOURCES = ["csf_pair.cu", "csf_pair_tiled.cu", "csf_agent.cu", "csf_peer.cu", "csf_order.cu", "csf_traj.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "csf_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines: str | None = None,
          sources=None) -> str:
    """``out`` / ``defines`` / ``sources``: a variant build next to the product library (tuning experiments):
    the listed sources are compiled with the extra defines into objects of their own, the rest is linked
    from the product build's objects."""
    if out is not None:
        return _build_variant(out, defines or "", sources or ["csf_pair_tiled.cu"])
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "csrc", src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("CSF_BUILD_DEFINES", "").split(), "-I", os.path.join(ROOT, "include"),
               "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + (out or ""))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB + ".tmp", *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    os.replace(LIB + ".tmp", LIB)
    return LIB


def _build_variant(out, defines, sources):
    build()                                           # the product objects must exist
    nvcc = _nvcc()
    tag = os.path.splitext(os.path.basename(out))[0]
    objs = []
    for src in SOURCES:
        if src in sources:
            obj = os.path.join(CSRC, src.replace(".cu", f".{tag}.o"))
            cmd = [nvcc, *NVCC_FLAGS, *defines.split(), "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c",
                   os.path.join(CSRC, src), "-o", obj]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout)
            objs.append(obj)
        else:
            objs.append(os.path.join(CSRC, src.replace(".cu", ".o")))
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    for o in objs:
        if f".{tag}.o" in o:
            os.remove(o)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
