"""Builds libadmm_b200.so in-tree with nvcc for sm_100a (called by __graft_entry__.build())."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "admm_b200.cu")
# translation units: the host library + three thirds of the persistent-kernel template variants
UNITS = ["admm_b200.cu", "iter_smem.cu", "iter_gshared.cu", "iter_pp.cu", "iter_pptma.cu", "iter_res.cu", "iter_wg.cu", "iter_wgpp.cu", "iter_pint.cu"]
OUT_DIR = os.path.join(HERE, os.environ.get("ADMMB_BUILD_DIR", "lib"))   # developer builds may go elsewhere
OUT = os.path.join(OUT_DIR, "libadmm_b200.so")

NVCC_FLAGS = ["-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-std=c++17",
              # explicit fma() only: the operation order written in the kernels is the one executed
              "-fmad=false"]
# developer builds only (e.g. ADMMB_EXTRA_NVCC_FLAGS=-DWG_TIMING prints a per-warp cycle timeline of one iteration)
NVCC_FLAGS += os.environ.get("ADMMB_EXTRA_NVCC_FLAGS", "").split()


def sources() -> list[str]:
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh"))] + \
        [os.path.join(os.path.dirname(HERE), "include", "admm_b200.h")]


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(s) <= t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OUT_DIR, exist_ok=True)
    obj_dir = os.path.join(OUT_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)

    def compile_unit(name: str) -> str:
        obj = os.path.join(obj_dir, name.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", obj, os.path.join(HERE, "csrc", name)]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(compile_unit, UNITS))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
