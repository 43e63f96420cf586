"""Build the CUDA shared library in-tree with nvcc for sm_100a (no JIT cache, no torch types).

    python -m redback_b200._build          # build if stale
    python -m redback_b200._build --force
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libredback_b200.so")

SOURCES = ["capi.cu"]
DEPS = ["capi.cu", "sn_kernel.cu", "rb_common.cuh", "rb_tables.inc"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _dep_files():
    return sorted(os.path.join(CSRC, d) for d in os.listdir(CSRC) if not d.startswith(".")) + \
        [os.path.join(ROOT, "include", "redback_b200.h")]


def source_hash():
    """Digest of every source the library is built from; compiled in (-DRB_SOURCE_HASH) and compared at load time,
    so a library left over from older sources is rebuilt instead of being bound with newer ctypes structures."""
    import hashlib
    h = hashlib.sha1()
    for path in _dep_files():
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _dep_files())


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB_PATH) and _built_hash() == source_hash():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    import fcntl
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)     # several ranks may find the library stale at the same time
        digest = source_hash()
        if not force and os.path.exists(LIB_PATH) and _built_hash() == digest:
            return LIB_PATH
        tmp = LIB_PATH + f".tmp{os.getpid()}"
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "--fmad=true", "-Xcompiler", "-fPIC", "-shared", f'-DRB_SOURCE_HASH="{digest}"',
               "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        os.replace(tmp, LIB_PATH)
    return LIB_PATH


def _built_hash():
    """The digest compiled into the library (rb_source_hash() returns "RBHASH:<digest>"), read from the file."""
    try:
        with open(LIB_PATH, "rb") as fh:
            blob = fh.read()
    except OSError:
        return None
    i = blob.find(b"RBHASH:")
    return blob[i + 7:i + 23].decode("ascii", "replace") if i >= 0 else None


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
