"""Builds libmopa_scn.so (sm_100a only) in-tree with nvcc. No torch dependency: the library is plain CUDA + C ABI."""
import fcntl
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.environ.get("MOPA_SCN_LIB") or os.path.join(HERE, "libmopa_scn.so")  # MOPA_SCN_LIB: a debug build (e.g. trace)
SOURCES = ["geometry.cu", "conv.cu", "conv_tc.cu", "conv_dw_tc.cu", "bn_io.cu", "program.cu", "xm_ops.cu", "vgi.cu"]
HEADERS = ["common.cuh", "geometry.cuh", "ptx.cuh"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmopa_scn.so cannot be built")


def stale():
    if os.environ.get("MOPA_SCN_LIB"):
        return False  # an explicitly chosen library is used as it is
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, h) for h in ("mopa_scn.h", "mopa_xm.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile + link under an exclusive file lock (torchrun starts one process per GPU on the same checkout: only one
    of them builds, the others wait and find a fresh library). Objects and the library are written under temporary names
    and moved into place with os.replace, so a concurrent dlopen never sees a half-written file."""
    if not force and not stale():
        return LIB
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale():  # another process built it while we waited
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
    tmpdir = tempfile.mkdtemp(prefix=".build-", dir=HERE)
    try:
        tmp_objs, procs = [], []
        for src in SOURCES:
            obj = os.path.join(tmpdir, src.replace(".cu", ".o"))
            cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                   "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            for d in os.environ.get("MOPA_BUILD_DEFS", "").split():  # e.g. MOPA_TC_TRACE (debug timeline in conv_tc.cu)
                cmd.insert(1, "-D" + d)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            tmp_objs.append(obj)
        for src, p in procs:
            out, _ = p.communicate()
            if verbose or p.returncode:
                print(out)
            if p.returncode:
                raise RuntimeError("nvcc failed on %s" % src)
        tmp_lib = os.path.join(tmpdir, "libmopa_scn.so")
        subprocess.check_call([_nvcc(), "-shared", "-o", tmp_lib] + tmp_objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
        for obj in tmp_objs:  # keep the objects beside the sources (cuobjdump / -res-usage inspection)
            os.replace(obj, os.path.join(CSRC, os.path.basename(obj)))
        os.replace(tmp_lib, LIB)
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
