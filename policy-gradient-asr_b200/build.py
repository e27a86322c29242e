"""In-tree nvcc build of libpgasr_b200.so (sm_100a only).

    python policy-gradient-asr_b200/build.py [--force] [--verbose]

The library is written next to the sources (policy-gradient-asr_b200/lib/), so it travels to the GPU box
with the repo snapshot; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpgasr_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
    "--shared", "-cudart", "static", "--threads", "0",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "pgasr.h")]
    return any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)


def build_timing(extra=(), tag=""):
    """Instrumented variant (-DPGASR_TIMING) for tools/phase_timing.py; not used by the product."""
    out = os.path.join(LIBDIR, f"libpgasr_b200_timing{tag}.so")
    os.makedirs(LIBDIR, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    r = subprocess.run([nvcc] + NVCC_FLAGS + ["-DPGASR_TIMING"] + list(extra) + sources() + ["-o", out], env=env,
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building the timing variant")
    return out


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", LIB]
    env = dict(os.environ)
    env.pop("CC", None)      # the image's CC wrapper is not what nvcc should use as host compiler
    env.pop("CXX", None)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libpgasr_b200.so")
    return LIB


if __name__ == "__main__":
    if "--timing" in sys.argv:
        extra = [a for a in sys.argv[1:] if a.startswith("-D")]
        tag = "".join("_" + a[2:].lower() for a in extra)
        print(build_timing(extra, tag))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
