"""Build libnele_score.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every ``csrc/*.cu`` is compiled to an object file under ``csrc/build/`` (in parallel, only when the source or a
header is newer than the object) and the objects are linked into ``nele_gan_b200/libnele_score.so``."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libnele_score.so")
SOURCES = ["engine.cu", "haspi.cu", "haspi_v1.cu", "estoi.cu", "siib.cu", "siib_knn.cu", "siib_eig.cu", "siib_klt.cu",
           "features.cu", "resyn.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers_mtime():
    t = os.path.getmtime(os.path.join(HERE, "..", "include", "nele_score.h"))
    for f in os.listdir(CSRC):
        if f.endswith((".h", ".hpp", ".cuh")):
            t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    return t


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return _headers_mtime() > t or any(os.path.getmtime(s) > t for s in sources())


def _compile(src, obj, verbose):
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    r = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, r.returncode, r.stdout


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers_mtime()
    jobs, objs = [], []
    for s in sources():
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or verbose or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr):
            jobs.append((s, o))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, rc, out in ex.map(lambda so: _compile(so[0], so[1], verbose), jobs):
            if rc != 0:
                raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
            if verbose:
                print(out)
    r = subprocess.run([_nvcc(), "--shared", "-o", LIB] + objs, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
