"""Build recipe for libroar_sup.so (sm_100a only, in-tree so it travels with the snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libroar_sup.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-shared", "-Xcompiler", "-fPIC"]


def sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    out.append(os.path.join(HERE, "..", "include", "roar_sup.h"))
    return out


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False, out=None, defines=()):
    """``out`` / ``defines``: a variant build (``-DNAME[=V]`` flags) written elsewhere, selected at run time with
    ``ROAR_SUP_LIB=<out>`` -- kernel A/B measurements and the instrumented Viterbi (``ROAR_VIT_STATS``)."""
    if out is None and not force and not needs_build():
        return LIB_PATH
    target = out or LIB_PATH
    cmd = [NVCC] + FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", target, os.path.join(CSRC, "roar_sup.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libroar_sup.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else None, defines=defs))
