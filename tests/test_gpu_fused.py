"""Parity of the fused CG kernel K6f (csrc/cg_fused.cu: a whole Jacobi-PCG solve as one cooperative
kernel) against the oracle and against the three-kernel iteration.  K6f is the default on one GPU whenever
the rows fit on chip (WAVE_CG_FUSED=0 switches it off, =1 also selects it for several ranks)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from wavegpu import WaveSolver, api, problem

pytestmark = pytest.mark.gpu


@pytest.fixture
def fused_env(monkeypatch):
    monkeypatch.setenv("WAVE_CG_FUSED", "1")


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.parametrize("name,scheme,over", [
    ("standing-mode-wsol", "newmark", dict(Nel="24", R=1, Dt="0.05")),
    ("standing-mode-wsol", "theta", dict(Nel="16", R=2, Dt="0.05", Theta="0.5")),
    ("sine-membrane", "theta", dict(Nel="45, 15")),
    ("standing-mode-wsol", "newmark", dict(Nel="300, 40", R=1, Dt="0.01")),
    ("ricker-wavelet", "newmark", dict(Nel="150", R=2)),
])
def test_fused_matches_oracle(fused_env, name, scheme, over):
    p = problem(name, **over)
    o = O.Oracle.from_params(p)
    g = WaveSolver(p, scheme)
    assert g.cg_fused_active()
    if scheme == "newmark":
        o.newmark_init(float(p["Dt"]), float(p["Beta"]), float(p["Gamma"]))
    else:
        o.theta_init(float(p["Dt"]), float(p["Theta"]))
    g.init()
    for _ in range(6):
        (o.newmark_step if scheme == "newmark" else o.theta_step)()
        its, _ = g.step()
        assert its == o.iterations()
    assert rel(g.vector(api.VEC_U), o.vector(O.Oracle.U)) < 1e-10
    assert rel(g.vector(api.VEC_V), o.vector(O.Oracle.V)) < 1e-10
    g.close()


def test_fused_matches_three_kernel_path_at_bench_size(monkeypatch):
    p = problem("standing-mode-wsol", Nel="1024", R=1, Dt="0.01")
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("WAVE_CG_FUSED", mode)
        g = WaveSolver(p, "newmark")
        assert g.cg_fused_active() == (mode == "1")
        g.init()
        its = [g.step()[0] for _ in range(5)]
        out[mode] = (its, g.vector(api.VEC_U), g.energy())
        g.close()
    assert out["0"][0] == out["1"][0]
    assert rel(out["1"][1], out["0"][1]) < 1e-10
    assert abs(out["1"][2] - out["0"][2]) < 1e-10 * abs(out["0"][2])


def _rank_worker(rank, world, nccl_id, name, scheme, over, nsteps, q):
    import sys
    from pathlib import Path

    import torch

    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    sys.path.insert(0, str(root / "nmpde-wave-equation_b200"))
    os.environ["WAVE_CG_FUSED"] = "1"
    from wavegpu import WaveSolver, api, problem

    torch.cuda.set_device(rank)
    g = WaveSolver(problem(name, **over), scheme, rank=rank, nranks=world, nccl_id=nccl_id, device=rank)
    fused = g.cg_fused_active()
    g.init()
    its = [g.step()[0] for _ in range(nsteps)]
    u, e = g.vector(api.VEC_U), g.energy()
    if rank == 0:
        q.put((fused, u, e, its))
    g.close()


@pytest.mark.parametrize("name,scheme,over", [
    ("standing-mode-wsol", "newmark", dict(Nel="40, 64", R=1, Dt="0.01")),
    ("sine-membrane", "theta", dict(Nel="30, 12", R=2)),
])
def test_fused_ranks_match_one_rank(monkeypatch, name, scheme, over):
    """K6f over two (or four) GPUs: sums over the NVLink mailboxes and halo stores inside the cooperative
    kernel; the one-GPU three-kernel path is the yardstick."""
    import torch
    import torch.multiprocessing as mp

    world = 4 if torch.cuda.device_count() >= 4 else 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("WAVE_CG_FUSED", "0")
    nsteps = 8
    g = WaveSolver(problem(name, **over), scheme)
    g.init()
    its1 = [g.step()[0] for _ in range(nsteps)]
    u1, e1 = g.vector(api.VEC_U), g.energy()
    g.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    nccl_id = api.comm_unique_id()
    procs = [ctx.Process(target=_rank_worker, args=(rk, world, nccl_id, name, scheme, over, nsteps, q))
             for rk in range(world)]
    for p in procs:
        p.start()
    fused, u2, e2, its2 = q.get(timeout=300)
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs) and fused
    assert its2 == its1
    assert rel(u2, u1) < 1e-10 and abs(e2 - e1) <= 1e-10 * abs(e1)
