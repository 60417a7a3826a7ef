"""Build the CUDA extension in-tree: alphadia_b200/libalphadia_b200.so (sm_100a only).

    python -m alphadia_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO_PATH = os.environ.get("ADB_LIB_PATH") or os.path.join(HERE, "libalphadia_b200.so")
SOURCES = ["adb_api.cu", "adb_select.cu", "adb_score.cu", "adb_score_dp.cu", "adb_misc.cu", "adb_select4d.cu", "adb_score4d.cu", "adb_classifier.cu"]
# per-file flags: the data-parallel scoring passes are written in plain arithmetic and must not be FMA-contracted
EXTRA_FLAGS = {"adb_score_dp.cu": ["--fmad=false"]}
HEADERS = [os.path.join(CSRC, "adb_common.cuh"), os.path.join(CSRC, "adb_score_dp.cuh"), os.path.join(os.path.dirname(HERE), "include", "alphadia_b200.h")]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found (needed to build alphadia_b200/libalphadia_b200.so)")
    return p


def is_stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return SO_PATH
    nvcc = nvcc_path()
    objs = []
    procs = []
    build_dir = os.path.join(HERE, "build" + (("_" + os.path.basename(SO_PATH)) if os.environ.get("ADB_LIB_PATH") else ""))
    os.makedirs(build_dir, exist_ok=True)
    common = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
    common += os.environ.get("ADB_NVCC_DEFS", "").split()  # tuning experiments only
    if verbose:
        common += ["-Xptxas", "-v"]
    for s in SOURCES:
        obj = os.path.join(build_dir, s.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((s, subprocess.Popen(common + EXTRA_FLAGS.get(s, []) + ["-c", os.path.join(CSRC, s), "-o", obj],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc {s} failed ---\n{out}\n")
        elif verbose and out:
            sys.stderr.write(f"--- nvcc {s} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    subprocess.check_call([nvcc, "-shared", "-o", SO_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                    "-cudart", "static"])
    return SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
