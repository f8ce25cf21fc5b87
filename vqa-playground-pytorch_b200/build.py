"""Build libvqacore_sm100a.so in-tree with nvcc (sm_100a only; cross-compiles without a GPU).

    python vqa-playground-pytorch_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.  No torch headers are
involved: the library is plain CUDA behind the C ABI of include/vqacore.h.
"""
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_NAME = "libvqacore_sm100a.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "--std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-DVQA_BUILDING_LIB",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libvqacore_sm100a.so cannot be built (there is no CPU fallback)")
    return exe


def _newest_header_mtime():
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))
    hdrs.append(os.path.join(os.path.dirname(PKG_DIR), "include", "vqacore.h"))
    hdrs.append(os.path.abspath(__file__))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link the shared library. Returns its path."""
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_m = _newest_header_mtime()
    nvcc = _nvcc()
    todo, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            todo.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for s, r in ex.map(compile_one, todo):
                if verbose or r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed on %s" % s)
    need_link = force or bool(todo) or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs)
    if need_link:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
