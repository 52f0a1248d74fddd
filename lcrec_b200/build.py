"""Build recipe for liblcrec_b200.so (nvcc, sm_100a only, in-tree so it ships with gpurun)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblcrec_b200.so")
SOURCES = ["abi.cu", "linear_tf32x3.cu", "linear_pair.cu", "rq_fused.cu", "sinkhorn.cu", "collide.cu", "ema.cu", "pool.cu", "kmeans.cu", "rq_train.cu", "optim.cu", "kmeanspp.cu", "train_extra.cu", "json_emit.cu", "sinkhorn_cluster.cu", "exchange.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "lcrec_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- nvcc {src} (exit {p.returncode})\n{out}", file=sys.stderr)
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    tmp = LIB + ".tmp"       # link aside, then rename: a snapshot taken meanwhile never sees a half-written library
    subprocess.check_call([nvcc, "-shared", "-o", tmp, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
