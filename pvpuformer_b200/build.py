"""In-tree build of libvpuformer_b200.so (hand-written sm_100a CUDA + the C ABI).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
repo snapshot.  `python -m pvpuformer_b200.build` or build(force=True).
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvpuformer_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")
SOURCES = ["gemm.cu", "gemm_b2b.cu", "gemm_res.cu", "gemm_gn.cu", "gemm_ln.cu", "head_tail.cu", "attention.cu", "attention_tc.cu", "attention_dma.cu", "prompt.cu", "elementwise.cu", "noc.cu", "session.cu", "raster.cu", "api.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def is_current():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _digest()


def build(force=False, verbose=False, debug=False):
    """Compile every .cu for sm_100a and link the shared library.  Returns the library path.

    debug=True builds libvpuformer_b200_debug.so with -DVPU_DEBUG -DVPU_ATTN_DEBUG next to the shipped library: only that build
    reads the development knobs (VPU_GEMM_IMPL, VPU_ATTN_TC, VPU_ATTN_NARROW, VPU_PDL, VPU_GEMM_*, VPU_ATTN_ABLATE) and compiles
    the clock64 trace points; load it with VPU_LIB_PATH (tools/ab_bench.sh)."""
    if debug:
        return _build_to(os.path.join(HERE, "libvpuformer_b200_debug.so"), os.path.join(HERE, "build_debug"),
                         ["-DVPU_DEBUG", "-DVPU_ATTN_DEBUG"], verbose)
    if not force and is_current():
        return LIB
    _build_to(LIB, os.path.join(HERE, "build"), [], verbose)
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB


def _build_to(lib_path, objdir, extra, verbose):
    nvcc = _nvcc()
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + [f for f in FLAGS if f != "--use_fast_math=false"] + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose:
            print(out)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", lib_path] + objs + ["-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
