"""Build libwavegpu.so (sm_100a) and the host executables in-tree with nvcc / g++.

    python nmpde-wave-equation_b200/build.py [--force]

Outputs: nmpde-wave-equation_b200/lib/libwavegpu.so, bin/main-newmark, bin/main-theta, bin/wave-mpirun.
nvcc cross-compiles without a GPU; the built files travel to the GPU box with the snapshot."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
HOST = HERE / "host"
OBJ = HERE / "build"
LIB = HERE / "lib" / "libwavegpu.so"
BIN = HERE / "bin"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVFLAGS = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", *ARCH]
CXXFLAGS = ["-std=c++17", "-O2", "-fPIC", "-Wall"]

LIB_SOURCES = ["kernels.cu", "ctx.cu", "cg_fused.cu", "expr.cpp", "quadrature.cpp"]
HOST_SOURCES = ["ParameterReader.cpp", "WaveEquationBase.cpp", "WaveNewmark.cpp", "WaveTheta.cpp", "cli.cpp",
                "launch_env.cpp", "vtu_writer.cpp"]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def _run(cmd):
    subprocess.check_call([str(c) for c in cmd])


EXECUTABLES = ("main-newmark", "main-theta", "host_selftest", "wave-mpirun")


def build_executables(force=False, only_missing=False):
    """Host executables (same names as the reference's CMake targets, plus the CPU self-test and the
    launcher).  only_missing: build what is absent and leave existing files alone, whatever their age
    (a snapshot copied to another machine does not keep modification times)."""
    BIN.mkdir(exist_ok=True)
    hdeps = list(HOST.glob("*.hpp")) + list(HOST.glob("*.cpp")) + [HERE.parent / "include" / "wavegpu.h"]
    for exe in EXECUTABLES:
        target = BIN / exe
        if only_missing:
            if target.exists():
                continue
        elif not (force or _newer(target, hdeps + [LIB])):
            continue
        # the launcher is a single translation unit; the others share the host classes
        common = [] if exe == "wave-mpirun" else [HOST / s for s in HOST_SOURCES]
        _run(["g++", *CXXFLAGS, "-I", HERE.parent / "include", "-o", target, HOST / (exe + ".cpp"),
              *common, "-L", LIB.parent, "-lwavegpu", "-Wl,-rpath,$ORIGIN/../lib"])


def build(force=False, verbose=False):
    OBJ.mkdir(exist_ok=True)
    LIB.parent.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.h")) + list(CSRC.glob("*.hpp")) + list(CSRC.glob("*.cuh")) + \
        [HERE.parent / "include" / "wavegpu.h"]
    jobs = []
    objs = []
    for src in LIB_SOURCES:
        s = CSRC / src
        o = OBJ / (src + ".o")
        objs.append(o)
        if force or _newer(o, [s, *headers]):
            if src.endswith(".cu"):
                extra = ["-Xptxas", "-v"] if verbose else []
                jobs.append([NVCC, *NVFLAGS, *extra, "-c", s, "-o", o])
            else:
                jobs.append(["g++", *CXXFLAGS, "-c", s, "-o", o])
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(_run, jobs))
    if force or jobs or not LIB.exists():
        _run([NVCC, "-shared", *ARCH, "-o", LIB, *objs, "-ldl"])
    build_executables(force=force, only_missing=False)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
