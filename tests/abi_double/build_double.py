"""Build the host executables against the TEST DOUBLE of the C ABI (wave_abi_on_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- see the header of wave_abi_on_oracle.cpp.  Outputs go to tests/_build/
(git-ignored): main-newmark, main-theta (the product's host sources from nmpde-wave-equation_b200/host/,
unchanged, linked with the double and the oracle instead of libwavegpu.so) and wave-mpirun (the product's
launcher, whose only ABI call is wave_device_count).  Nothing under nmpde-wave-equation_b200/ uses them.

    python tests/abi_double/build_double.py [--force]
"""
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
PKG = ROOT / "nmpde-wave-equation_b200"
OUT = ROOT / "tests" / "_build"
HOST_SOURCES = ["ParameterReader.cpp", "WaveEquationBase.cpp", "WaveNewmark.cpp", "WaveTheta.cpp", "cli.cpp",
                "launch_env.cpp", "vtu_writer.cpp"]
CXX = ["g++", "-std=c++17", "-O2", "-Wall"]


def _stale(target, deps):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force=False):
    """Returns the directory holding main-newmark, main-theta and wave-mpirun linked with the double."""
    sys.path.insert(0, str(ROOT))
    from oracle import oracle as O

    oracle_so = O.build()
    OUT.mkdir(exist_ok=True)
    inc = ["-I", str(ROOT / "include"), "-I", str(PKG / "csrc"), "-I", str(PKG / "host")]
    double_src = [HERE / "wave_abi_on_oracle.cpp", PKG / "csrc" / "expr.cpp", PKG / "csrc" / "quadrature.cpp"]
    host_src = [PKG / "host" / s for s in HOST_SOURCES]
    deps = double_src + host_src + list((PKG / "host").glob("*.hpp")) + list((PKG / "csrc").glob("*.h")) + \
        list((PKG / "csrc").glob("*.hpp")) + [ROOT / "include" / "wavegpu.h", oracle_so, Path(__file__)]
    link = [str(oracle_so), f"-Wl,-rpath,{oracle_so.parent}", "-fopenmp"]
    mains = [PKG / "host" / (exe + ".cpp") for exe in ("main-newmark", "main-theta", "wave-mpirun")]
    obj_dir = OUT / "obj"
    obj_dir.mkdir(exist_ok=True)
    objs = {src: obj_dir / (src.name + ".o") for src in [*double_src, *host_src, *mains]}
    jobs = [[*CXX, *inc, "-c", str(src), "-o", str(o)] for src, o in objs.items() if force or _stale(o, deps + [src])]
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(subprocess.check_call, jobs))
    for main in mains:
        target = OUT / main.stem
        if force or jobs or not target.exists():
            common = double_src if main.stem == "wave-mpirun" else [*host_src, *double_src]
            subprocess.check_call([*CXX, "-o", str(target), str(objs[main]), *(str(objs[s]) for s in common), *link])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
