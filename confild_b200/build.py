"""Build libconfild_cnf.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m confild_b200.build [--force] [--trace]

The library is compiled next to this file so that it travels with the repository snapshot to the
GPU box; there is no JIT cache and no fallback when it is missing.  The translation units under
``csrc/`` are compiled in parallel and linked against the SHARED CUDA runtime (``-cudart shared``:
the runtime is already in the process through PyTorch, and no runtime code is embedded in the .so).
Concurrent builders (several ranks of one torchrun) serialise on a file lock and the finished
library is moved into place atomically, so a reader never maps a half-written file.
"""
from __future__ import annotations

import fcntl
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libconfild_cnf.so")
OBJ_DIR = os.path.join(PKG_DIR, "_obj")
LOCK_PATH = os.path.join(PKG_DIR, ".build.lock")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    return ([os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inl", ".h"))]
            + [os.path.join(PKG_DIR, "..", "include", "confild_cnf.h")])


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libconfild_cnf.so")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    lib_m = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > lib_m for d in _deps() if os.path.exists(d))


TRACE_LIB_PATH = os.path.join(PKG_DIR, "libconfild_cnf_trace.so")  # -DCNF_TRACE build for tests/tools/trace_*.py


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    lib_path = TRACE_LIB_PATH if trace else LIB_PATH
    obj_dir = OBJ_DIR + ("_trace" if trace else "")
    if trace:
        force = True
    if not force and not is_stale():
        return lib_path
    nvcc = _nvcc()
    os.makedirs(obj_dir, exist_ok=True)
    with open(LOCK_PATH, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():  # another process built it while we waited for the lock
                return lib_path
            flags = NVCC_FLAGS + (["-DCNF_TRACE"] if trace else [])
            srcs = _sources()

            def compile_one(src):
                obj = os.path.join(obj_dir, src[:-3] + ".o")
                cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
                proc = subprocess.run(cmd, capture_output=True, text=True)
                return src, obj, " ".join(cmd), proc.returncode, proc.stdout + proc.stderr

            with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as pool:
                results = list(pool.map(compile_one, srcs))
            log = "".join(f"$ {cmd}\n{out}\n" for _, _, cmd, _, out in results)
            failed = [(src, out) for src, _, _, rc, out in results if rc != 0]
            if not failed:
                fd, tmp = tempfile.mkstemp(prefix=".libconfild_cnf.", suffix=".so", dir=PKG_DIR)
                os.close(fd)
                cmd = [nvcc, "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a",
                       "-Xlinker", "-rpath,/usr/local/cuda/lib64"] + [obj for _, obj, _, _, _ in results] + ["-o", tmp]
                proc = subprocess.run(cmd, capture_output=True, text=True)
                log += "$ " + " ".join(cmd) + "\n" + proc.stdout + proc.stderr
                if proc.returncode != 0:
                    failed = [("link", proc.stdout + proc.stderr)]
                    os.unlink(tmp)
                else:
                    os.chmod(tmp, 0o755)
                    os.replace(tmp, lib_path)  # atomic: a concurrent dlopen sees the old or the new file, never a part
            with open(os.path.join(PKG_DIR, "build.log"), "w") as f:
                f.write(log)
            if failed:
                raise RuntimeError("nvcc failed:\n" + "\n".join(f"[{s}]\n{o[-4000:]}" for s, o in failed))
            if verbose:
                print(log)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return lib_path


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, trace="--trace" in sys.argv)
    print("built", path)
