"""Build libpdeopt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Each csrc/*.cu is compiled to an object file in parallel (one nvcc process per file, objects under
csrc/_obj/, rebuilt only when the source or any header is newer) and the objects are linked into the
shared library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libpdeopt_b200.so")

NVCC_FLAGS = [
    "-std=c++17",
    "--expt-relaxed-constexpr",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-diag-suppress", "128",
]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _headers():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "pdeopt_b200.h"))
    deps.append(os.path.abspath(__file__))
    return deps


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    # the library is current when it is newer than every source and header (the object files are a
    # local cache: they do not travel to the GPU box)
    return _stale(LIB, _sources() + _headers())


def build(force=False, verbose=False, extra_flags=()):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers()
    todo = [s for s in _sources() if force or _stale(_obj(s), [s] + hdr)]

    def compile_one(src):
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", src, "-o", _obj(src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        return src, r

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        results = list(ex.map(compile_one, todo))
    for src, r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise subprocess.CalledProcessError(r.returncode, f"nvcc -c {src}")
    cmd = [nvcc, "--shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
           *[_obj(s) for s in _sources()], "-o", LIB, "-lcudart"]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
