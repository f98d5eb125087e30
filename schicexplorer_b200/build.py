"""In-tree build of the CUDA library (sm_100a only) with plain nvcc -- no JIT cache, so the built
``.so`` travels with the repository snapshot to the GPU box."""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libschic_mh.so")
SOURCES = ["capi.cu", "k1_signatures.cu", "k1_table.cu", "k2_buckets.cu", "k2_combine.cu", "k3_collisions.cu", "k3_select.cu", "k4_rerank.cu", "k4c_rerank.cu", "k5_graph.cu", "k6_longrows.cu", "k7_relabel.cu",
           "int_peak.cu", "pca_dense.cu", "hostconv.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]
# SCHIC_K1_ALL_VARIANTS=1 also compiles the 20 instruction-selection variants of K1 that
# tools/k1_variants.py times (experiment only; the product uses K1_VARIANT in k1_signatures.cu)
if os.environ.get("SCHIC_K1_ALL_VARIANTS") == "1":
    NVCC_FLAGS.append("-DK1_ALL_VARIANTS")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False, variant: str = "", extra_flags=()) -> str:
    """``variant`` / ``extra_flags``: an experiment build of the same library with other compile-time constants
    (objects and the .so get the variant name as a suffix; loaded through SCHIC_CUDA_LIB, see tools/)."""
    nvcc = os.environ.get("NVCC", "nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(HERE, "..", "include", f) for f in ("schic_mh.h", "schic_mh_cuda.h", "schic_pca.h")]
    headers.append(os.path.abspath(__file__))
    objdir = os.path.join(CSRC, "build" + ("_" + variant if variant else ""))
    os.makedirs(objdir, exist_ok=True)
    out = OUT if not variant else os.path.join(objdir, f"libschic_mh_{variant}.so")

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return o

    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(out, objs):
        cmd = [nvcc, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose=True))
