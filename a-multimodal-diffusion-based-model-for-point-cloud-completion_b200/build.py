"""In-tree build of libpcd_b200.so (nvcc, sm_100a only).

    python -m <package>.build        # or: python build.py

The shared library is written next to this file so that it travels with the
repository snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpcd_b200.so")
SOURCES = ["pcd_api.cu", "elementwise.cu", "sampler.cu", "gemm_simt.cu", "attn_simt.cu",
           "gemm_tc.cu", "attn_tc_launch.cu", "attn_tc5.cu", "attn_tc8.cu", "pointops.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return nvcc


STAMP = LIB + ".srchash"


def source_hash() -> str:
    """Content hash of everything the library is built from (sources, header, flags).  Content, not mtimes: the
    snapshot that travels to the GPU box does not preserve them."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS + SOURCES).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "pcd_b200.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build_variant(out_path: str, extra_flags, objdir_name: str) -> str:
    """A side build of the library with extra compiler flags (profiling / tracing builds under tools/; never the
    product library).  Load it with PCD_B200_LIB=<out_path>."""
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build", objdir_name)
    os.makedirs(objdir, exist_ok=True)
    procs = [(src, os.path.join(objdir, src.replace(".cu", ".o"))) for src in SOURCES]
    running = [(src, obj, subprocess.Popen([nvcc, *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj],
                                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for src, obj in procs]
    for src, obj, p in running:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    r = subprocess.run([nvcc, "-shared", "-o", out_path, *[o for _, o in procs], "-Xcompiler", "-fPIC"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return out_path


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    import fcntl
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lock:  # several ranks may import at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():
            return LIB
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
        objs.append(obj)
    tmp = LIB + ".tmp"
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB)
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
