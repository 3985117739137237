"""world_size-2 (and 3) gloo test of the multi-GPU host logic on CPU: every rank takes its strip
from wave_partition_plan, exchanges the contiguous halo blocks the plan prescribes with
torch.distributed send/recv, multiplies its owned rows (oracle CSR rows) with its local ghosted
vector and all-reduces a dot product; the result must equal the single-rank product exactly as
SURVEY section 4 asks (N ranks give the 1-rank answer)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nx, ny, r, out):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "nmpde-wave-equation_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from wavegpu import partition_plan, problem

    o = O.Oracle.from_params(problem("standing-mode-wsol", Nel=f"{nx}, {ny}", R=r))
    rowptr, col = o.csr()
    K = o.values(O.Oracle.K)
    rng = np.random.default_rng(5)
    x_global = rng.standard_normal(o.n)
    plan = partition_plan(nx, ny, r, rank, world)
    plans = [partition_plan(nx, ny, r, q, world) for q in range(world)]
    lo, hi = plan.ghost_lo_begin, plan.ghost_hi_end
    own0, own1 = plan.row_begin, plan.row_end
    local = np.full(hi - lo, np.nan)
    local[own0 - lo:own1 - lo] = x_global[own0:own1]  # owned values only; ghosts arrive by exchange
    t = torch.from_numpy(local)
    reqs = []
    if rank > 0:  # lower neighbour: receive its last block, send my first block
        nb = plans[rank - 1]
        reqs.append(dist.irecv(t[0:own0 - lo], src=rank - 1))
        reqs.append(dist.isend(t[own0 - lo:own0 - lo + (nb.ghost_hi_end - nb.row_end)].clone(), dst=rank - 1))
    if rank < world - 1:  # upper neighbour: receive its first block, send my last block
        nb = plans[rank + 1]
        reqs.append(dist.irecv(t[own1 - lo:], src=rank + 1))
        reqs.append(dist.isend(t[own1 - lo - (nb.row_begin - nb.ghost_lo_begin):own1 - lo].clone(), dst=rank + 1))
    for q in reqs:
        q.wait()
    assert not np.isnan(local).any()
    assert np.array_equal(local, x_global[lo:hi])
    y = np.empty(own1 - own0)
    for i in range(own0, own1):
        s = 0.0
        for e in range(rowptr[i], rowptr[i + 1]):
            s += K[e] * local[col[e] - lo]
        y[i - own0] = s
    y_ref = o.spmv(O.Oracle.K, x_global)[own0:own1]
    assert np.array_equal(y, y_ref)
    part = torch.tensor([float(y @ local[own0 - lo:own1 - lo])], dtype=torch.float64)
    dist.all_reduce(part)
    full = float(x_global @ o.spmv(O.Oracle.K, x_global))
    assert abs(part.item() - full) <= 1e-12 * abs(full)
    out[rank] = 1
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny,r", [(2, 6, 8, 1), (2, 5, 6, 2), (3, 4, 7, 2)])
def test_strip_partition_halo_exchange_gloo(world, nx, ny, r):
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", world)
    procs = [ctx.Process(target=_worker, args=(rk, world, port, nx, ny, r, out)) for rk in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs)
    assert list(out) == [1] * world
