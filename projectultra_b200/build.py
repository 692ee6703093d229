"""Builds projectultra_b200/libpu_b200.so IN-TREE with nvcc for sm_100a (cross-compiles without a GPU).

    python -m projectultra_b200.build [--force] [--verbose]
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpu_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# -fmad=false: parity-critical arithmetic (LDPC messages, FFT butterflies, equaliser) must round like the
# reference's unfused x86-64 build; kernels that want an FMA ask for it explicitly with fmaf().
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++20", "-fmad=false",
         "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3,-Wall", "-Xptxas", "-v", "-shared",
         "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(os.path.join(HERE, "pu_sweep")):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "pu", "*.h")) + [os.path.abspath(__file__), os.path.join(ROOT, "tools", "pu_sweep.cpp")]
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    return glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "pu", "*.h")) + [os.path.abspath(__file__)]


def build(force=False, verbose=False):
    """One nvcc process per translation unit (in parallel; objects cached under projectultra_b200/build/), then one link."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(d) for d in _deps())
    cflags = [f for f in FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, 0, ""
        r = subprocess.run([NVCC] + cflags + ["-c", "-o", obj, src], capture_output=True, text=True)
        return obj, r.returncode, "== " + os.path.basename(src) + "\n" + r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, sources()))
    logtxt = "".join(t for _, _, t in results)
    if verbose or any(rc for _, rc, _ in results):
        sys.stderr.write(logtxt)
    if any(rc for _, rc, _ in results):
        raise RuntimeError("nvcc failed building libpu_b200.so")
    r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [o for o, _, _ in results],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libpu_b200.so failed")
    if logtxt:
        with open(os.path.join(HERE, "build_ptxas.log"), "a" if not force else "w") as f:
            f.write(logtxt)
    build_tools()
    return LIB


SWEEP_BIN = os.path.join(HERE, "pu_sweep")


def build_tools():
    """The C++20 host program of the config-5 sweep (tools/pu_sweep.cpp), linked against the in-tree library."""
    src = os.path.join(ROOT, "tools", "pu_sweep.cpp")
    cmd = [os.environ.get("CXX", "g++"), "-std=c++20", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", SWEEP_BIN,
           "-L", HERE, "-lpu_b200", "-ldl", "-pthread", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("g++ failed building pu_sweep")
    return SWEEP_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
