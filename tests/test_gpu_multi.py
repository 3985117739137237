"""Multi-GPU parity (needs >= 2 GPUs, run with `gpurun --gpus 2`): the same run on 2 ranks (strip
partition, NCCL halo exchange + all-reduces) must give the 1-rank answer to 1e-10 in the canonical
numbering."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, nccl_id, name, scheme, over, nsteps, q, cg=None, fused=None, no_p2p=False):
    if fused is not None:
        os.environ["WAVE_CG_FUSED"] = fused
    if no_p2p:
        os.environ["WAVE_NO_P2P"] = "1"
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "nmpde-wave-equation_b200"))
    from wavegpu import WaveSolver, api, problem

    torch.cuda.set_device(rank)
    g = WaveSolver(problem(name, **over), scheme, rank=rank, nranks=world, nccl_id=nccl_id, device=rank, cg=cg)
    g.init()
    its = []
    for _ in range(nsteps):
        i, nrm = g.step()
        its.append(i)
    u, v = g.vector(api.VEC_U), g.vector(api.VEC_V)
    e = g.energy()
    err = g.errors() if g.has_solution else None
    if rank == 0:
        q.put((u, v, e, err, its, nrm))
    g.close()


CASES = [
    ("standing-mode-wsol", "newmark", dict(Nel="24", R=2, Dt="0.01")),
    ("standing-mode-wsol", "theta", dict(Nel="33, 17", R=1, Dt="0.02", Theta="0.5")),
    ("sine-membrane", "newmark", dict(Nel="30, 10")),
    ("ricker-wavelet", "theta", dict(Nel="16", R=2, Theta="1.0")),
]


def _world_sizes():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [w for w in (2, 4, 8) if w <= n] or [2]


@pytest.mark.parametrize("world", _world_sizes())
@pytest.mark.parametrize("fused", ["0", "1"])
@pytest.mark.parametrize("name,scheme,over", CASES)
def test_ranks_match_one_rank(name, scheme, over, fused, world):
    """2 ranks exercise the end strips; 4 and 8 ranks also the strips with two neighbours.  fused = "0": the
    three-kernel iteration with its sums over the NVLink mailboxes; "1": the cooperative kernel K6f (the
    default at these sizes)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    from wavegpu import WaveSolver, api, problem

    nsteps = 12
    g = WaveSolver(problem(name, **over), scheme)
    g.init()
    its1 = []
    for _ in range(nsteps):
        i, nrm1 = g.step()
        its1.append(i)
    u1, v1, e1 = g.vector(api.VEC_U), g.vector(api.VEC_V), g.energy()
    err1 = g.errors() if g.has_solution else None
    g.close()

    nccl_id = api.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, world, nccl_id, name, scheme, over, nsteps, q, None, fused))
             for rk in range(world)]
    for p in procs:
        p.start()
    u2, v2, e2, err2, its2, nrm2 = q.get(timeout=300)
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    scale = max(np.abs(u1).max(), 1e-300)
    assert np.abs(u2 - u1).max() <= 1e-10 * scale
    assert np.abs(v2 - v1).max() <= 1e-10 * max(np.abs(v1).max(), 1e-300)
    assert abs(e2 - e1) <= 1e-10 * max(abs(e1), 1e-300)
    assert its2 == its1
    assert np.allclose(nrm2, nrm1, rtol=1e-10)
    if err1 is not None:
        assert np.allclose([err2[0], err2[2]], [err1[0], err1[2]], rtol=1e-8)


def _mg_cases(world):
    """Meshes whose strips coarsen exactly like the single-rank hierarchy: every strip keeps whole coarse quad
    rows on every level (8 * world fine quad rows per rank, at most three mesh coarsenings at these Dt)."""
    ny = 8 * world * 2
    sq = f"[0.0, 1.0] x [0.0, {ny / 64.0!r}]"   # dy = dx = 1/64
    return [
        ("standing-mode-wsol", "newmark", dict(Nel=f"64, {ny}", Geometry=sq, R=1, Dt="0.1")),
        ("standing-mode-wsol", "newmark", dict(Nel=f"32, {ny}", Geometry=f"[0.0, 1.0] x [0.0, {ny / 32.0!r}]", R=2, Dt="0.05")),
        ("ricker-wavelet", "theta", dict(Nel=f"64, {ny}", Geometry=sq, R=2, Dt="0.02", Theta="1.0")),
        ("traveling-square-bump", "newmark",
         dict(Nel=f"48, {ny}", Geometry=f"[0.0, 3.0] x [0.0, {3.0 * ny / 48.0!r}]", R=1, Dt="0.05",
              C={"Function constants": "", "Variable names": "x, y, t",
                 "Function expression": "1.0 + 0.25*sin(2*pi*x/3)*sin(2*pi*y/3)"})),
    ]


@pytest.mark.parametrize("world", _world_sizes())
@pytest.mark.parametrize("case", range(4))
def test_multigrid_over_strips_matches_oracle(case, world):
    """The V-cycle preconditioner over strips (coarse strips owned by the fine strip's owner, halo exchange
    before every smoothing sweep and transfer): the oracle's multigrid-PCG iteration counts and solution."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    from oracle import oracle as O
    from wavegpu import api, problem

    name, scheme, over = _mg_cases(world)[case]
    mg = dict(precond=2)
    nsteps = 6
    p = problem(name, **over)
    o = O.Oracle.from_params(p)
    o.set_cg(**mg)
    dt = float(p["Dt"])
    if scheme == "newmark":
        o.newmark_init(dt, float(p["Beta"]), float(p["Gamma"]))
    else:
        o.theta_init(dt, float(p["Theta"]))
    its1 = []
    for _ in range(nsteps):
        (o.newmark_step if scheme == "newmark" else o.theta_step)()
        its1.append(tuple(o.iterations()))
    nccl_id = api.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, world, nccl_id, name, scheme, over, nsteps, q, mg))
             for rk in range(world)]
    for pr in procs:
        pr.start()
    u2, v2, e2, err2, its2, nrm2 = q.get(timeout=300)
    for pr in procs:
        pr.join(120)
    assert all(pr.exitcode == 0 for pr in procs)
    assert [tuple(i) for i in its2] == its1
    uo, vo = o.vector(O.Oracle.U), o.vector(O.Oracle.V)
    assert np.abs(u2 - uo).max() <= 1e-10 * max(np.abs(uo).max(), 1e-300)
    assert np.abs(v2 - vo).max() <= 1e-10 * max(np.abs(vo).max(), 1e-300)


@pytest.mark.parametrize("cg", [None, dict(precond=2)])
def test_nccl_fallback_matches_one_rank(cg):
    """WAVE_NO_P2P=1 (also the path for more than 8 ranks): no NVLink mailboxes, the sums of an iteration are
    NCCL all-reduces of device scalars and the halo of the search direction an NCCL exchange per iteration --
    the same iterates as one rank, with Jacobi and with the multigrid V-cycle."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from wavegpu import WaveSolver, api, problem

    name, scheme, over = "standing-mode-wsol", "newmark", dict(Nel="32, 64", R=2, Dt="0.05")
    nsteps = 6
    g = WaveSolver(problem(name, **over), scheme, cg=cg)
    g.init()
    its1 = [g.step()[0] for _ in range(nsteps)]
    u1, e1 = g.vector(api.VEC_U), g.energy()
    g.close()
    nccl_id = api.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, 2, nccl_id, name, scheme, over, nsteps, q, cg, None, True))
             for rk in range(2)]
    for p in procs:
        p.start()
    u2, v2, e2, err2, its2, nrm2 = q.get(timeout=300)
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert [tuple(i) for i in its2] == [tuple(i) for i in its1]
    assert np.abs(u2 - u1).max() <= 1e-10 * max(np.abs(u1).max(), 1e-300)
    assert abs(e2 - e1) <= 1e-10 * max(abs(e1), 1e-300)
