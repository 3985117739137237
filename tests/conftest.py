import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "nmpde-wave-equation_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# oracle / reference parity first, API semantics and GPU-vs-GPU consistency afterwards: with `-x` a
# failure in a later file can no longer hide the parity files
ORDER = ["test_oracle_golden", "test_oracle_structure", "test_oracle_crosscheck", "test_gpu_parity",
         "test_gpu_golden_sweep", "test_gpu_stencil", "test_gpu_multigrid", "test_gpu_multi", "test_gpu_cli"]


def pytest_collection_modifyitems(config, items):
    def rank(item):
        stem = Path(str(item.fspath)).stem
        return ORDER.index(stem) if stem in ORDER else len(ORDER)

    items.sort(key=rank)  # stable: the order inside a file is kept
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def host_executables():
    """bin/main-newmark, main-theta, host_selftest, wave-mpirun are build products (git-ignored): make
    sure they exist, building only what is missing (g++ against the in-tree library)."""
    import importlib.util

    pkg = ROOT / "nmpde-wave-equation_b200"
    spec = importlib.util.spec_from_file_location("wave_build", pkg / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if any(not (mod.BIN / exe).exists() for exe in mod.EXECUTABLES):
        from wavegpu import api

        api.lib()  # builds the library first when it is missing
        mod.build_executables(only_missing=True)
