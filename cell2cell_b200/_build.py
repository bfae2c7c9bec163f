"""In-tree build of ``libc2c_b200.so`` (hand-written sm_100a kernels + the C ABI of ``include/c2c_b200.h``).

``nvcc`` cross-compiles without a GPU; the translation units are independent, so they are compiled in parallel and
linked into ``cell2cell_b200/libc2c_b200.so`` (git-ignored, shipped to the GPU box with the tree).  Objects are cached
under ``build/`` keyed by a hash of the sources and flags.
"""
import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libc2c_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Wno-deprecated-gpu-targets"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libc2c_b200.so")
    return exe


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(extra=()):
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/c2c_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            with open(p, "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(list(NVCC_FLAGS) + list(extra)).encode())
    return h.hexdigest()[:16]


def build(force=False, verbose=False, jobs=None, extra_flags=(), out_path=None):
    """Compile every ``csrc/*.cu`` for sm_100a and link the shared library.  Returns its path.
    ``extra_flags`` / ``out_path`` build a tuning variant (e.g. ``-DC2C_PC_UNROLL=8``) beside the product library."""
    extra_flags = list(extra_flags)
    lib_path = out_path or LIB_PATH
    tag = _digest(extra_flags)
    obj_dir = os.path.join(ROOT, "build", tag)
    # the tag of the sources the library at `lib_path` was linked from is kept beside it: a stamp inside the object directory
    # alone would call a library current after the sources went back to an earlier state (the objects of that state are still
    # cached, but the library on disk is the later build)
    stamp = lib_path + ".tag"
    if not force and os.path.exists(lib_path) and os.path.exists(stamp) and open(stamp).read() == tag:
        return lib_path
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(obj_dir, src[:-3] + ".o")
        if os.path.exists(obj) and not force:
            return obj
        cmd = [nvcc] + NVCC_FLAGS + extra_flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj + ".tmp"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            with open(obj[:-2] + ".ptxas.log", "w") as fh:
                fh.write(res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, res.stderr[-4000:]))
        os.replace(obj + ".tmp", obj)
        return obj

    with ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", lib_path + ".tmp"] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s" % res.stderr[-4000:])
    os.replace(lib_path + ".tmp", lib_path)
    with open(stamp, "w") as fh:
        fh.write(tag)
    return lib_path


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
