"""In-tree build of libvisgeom_b200.so (nvcc, sm_100a only).

Used by __graft_entry__.build() and on demand by the package loader.  nvcc
cross-compiles without a GPU; the resulting .so travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# developer variants (VG_VARIANT=phase: per-phase clock64 counters in the evaluation kernel) build next to
# the product library and never replace it
VARIANT = os.environ.get("VG_VARIANT", "")
LIB = os.path.join(HERE, "libvisgeom_b200%s.so" % ("_" + VARIANT if VARIANT else ""))
OBJ = os.path.join(HERE, "_obj" + ("_" + VARIANT if VARIANT else ""))

SOURCES = ["vg_eval_eucm.cu", "vg_eval_ucm.cu", "vg_eval_mei.cu", "vg_eval.cu", "vg_api.cu", "vg_solver_kernels.cu",
           "vg_problem.cu", "vg_priors.cu", "vg_solver_fast.cu", "vg_project.cu", "vg_host.cu", "vg_corner.cu", "vg_refine.cu", "vg_detector.cu"]
# every header of csrc/ is a dependency of every object (a stale object with a mismatched cross-rank protocol or
# argument struct would load silently)
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))) + ["../../include/visgeom_b200.h"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]
if VARIANT.startswith("phase"):
    FLAGS.append("-DVG_PHASE_CLOCKS")
FLAGS += os.environ.get("VG_EXTRA_FLAGS", "").split()
# the detector's refinement rounds operation by operation like the reference's host code (see vg_detector.cu)
FILE_FLAGS = {"vg_detector.cu": ["-fmad=false"]}


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    procs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + FILE_FLAGS.get(s, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        subprocess.check_call(cmd)
    return LIB


CLI = os.path.join(HERE, "bin", "vg_calib")
HOST = os.path.join(HERE, "host")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")


def build_cli(force: bool = False) -> str:
    """g++ -> visgeom_b200/bin/vg_calib: the reference's `calib` tool (JSON front end, host C++) over the C ABI."""
    srcs = [os.path.join(HOST, f) for f in ("calibration.cpp", "calib_main.cpp")]
    deps = srcs + [os.path.join(HOST, f) for f in ("json.hpp", "image_io.hpp")] + [
        os.path.join(INCLUDE, "visgeom_b200", f) for f in ("calibration.hpp", "camera.hpp", "geometry.hpp", "corner_detector.hpp")] + [LIB]
    if force or _stale(CLI, deps):
        os.makedirs(os.path.dirname(CLI), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I" + INCLUDE, "-I" + HOST] + srcs +
                              ["-o", CLI, "-L" + HERE, "-lvisgeom_b200", "-lz", "-lpthread", "-Wl,-rpath,$ORIGIN/.."])
    return CLI


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if not VARIANT:
        print(build_cli(force="--force" in sys.argv))
