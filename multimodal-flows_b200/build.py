"""Build libmmf_b200.so in-tree with nvcc for sm_100a (no torch extension machinery involved)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "mmf_b200", "libmmf_b200.so")
SOURCES = ["kernels_tc.cu", "kernels_simt.cu", "kernels_train.cu", "kernels_traingemm.cu", "kernels_trainops.cu", "kernels_trainattn.cu", "train_api.cu", "kernels_epic.cu", "epic_model.cu", "kernels_tftile.cu", "kernels_tftile.cu@trace",
           "tftile_model.cu", "model.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [*os.environ.get("MMF_EXTRA_NVCC", "").split(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "--expt-relaxed-constexpr"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mmf_b200.h"),
                                                             os.path.join(HERE, "..", "include", "mmf_b200_train.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src: str) -> str:
        extra = []
        if src.endswith("@trace"):                    # second build of the tile kernel with clock stamps (debugging aid)
            src = src[: -len("@trace")]
            obj = os.path.join(objdir, src.replace(".cu", "_trace.o"))
            extra = ["-DMMF_TILE_TRACE=1"]
        else:
            obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    res = subprocess.run([NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
