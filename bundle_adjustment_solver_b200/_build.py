"""In-tree build of the CUDA engine (libba_b200.so) for sm_100a.  nvcc cross-compiles without a GPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# BA_B200_LIB: alternative build of the same sources (A/B kernel experiments); product default is in-tree
LIB = os.environ.get("BA_B200_LIB") or os.path.join(HERE, "libba_b200.so")
SOURCES = ["ba_engine.cu", "ba_poseonly.cu", "ba_geometry.cu"]
HEADERS = [os.path.join("..", "..", "include", "ba_b200.h"),
           os.path.join("..", "..", "include", "ba_b200", "utility", "geometry_math.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]
# BA_B200_NVCC_EXTRA: extra compiler flags for A/B builds (e.g. -DBA_ND_CONS=11), together with BA_B200_LIB
NVCC_FLAGS += os.environ.get("BA_B200_NVCC_EXTRA", "").split()


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps += [os.path.join(CSRC, f) for f in HEADERS]
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source into one shared library next to the package."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, f) for f in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd)
    return LIB
