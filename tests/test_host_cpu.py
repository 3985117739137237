"""CPU-side tests (no GPU): the C ABI loads and exports every declared symbol, the expression
compiler agrees with the oracle's independent evaluator, the closed-form DoF numbering equals the
oracle's first-touch numbering, the partition plan is consistent, and the host executables keep
the reference's error behaviour (src/main-newmark.cpp:92-102,155-166: message + exit code 1)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from wavegpu import api, cell_dofs, partition_plan, problem
from wavegpu.problems import NAMES, write_json
from wavegpu.vtu import read_vtu

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "nmpde-wave-equation_b200" / "bin"


def test_abi_exports_every_declared_symbol():
    header = (ROOT / "include" / "wavegpu.h").read_text()
    names = set(re.findall(r"\b(wave_[a-z0-9_]+)\s*\(", header))
    names -= {"wave_status"}
    L = api.lib()
    assert len(names) > 35
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/wavegpu.h but not exported"


def test_no_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(api.WaveError) as ei:
        api.WaveSolver(problem("standing-mode-wsol", Nel=4), "newmark")
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)


def test_product_never_links_or_loads_the_oracle_or_the_test_double():
    """The oracle (oracle/) and the test double of the ABI (tests/abi_double/) are test infrastructure: the
    product library and executables neither contain nor depend on them, no product source refers to their
    files, and the executables fail with the library's "no CPU fallback" message where there is no GPU."""
    pkg = ROOT / "nmpde-wave-equation_b200"
    lib = pkg / "lib" / "libwavegpu.so"
    syms = subprocess.run(["nm", "-D", "--defined-only", str(lib)], capture_output=True, text=True, check=True).stdout
    assert "wave_create" in syms and "oracle_" not in syms
    for exe in ("main-newmark", "main-theta", "wave-mpirun"):
        needed = subprocess.run(["readelf", "-d", str(BIN / exe)], capture_output=True, text=True, check=True).stdout
        assert "libwavegpu.so" in needed and "oracle" not in needed and "tests/_build" not in needed, exe
    needed = subprocess.run(["readelf", "-d", str(lib)], capture_output=True, text=True, check=True).stdout
    assert "oracle" not in needed
    for src in list(pkg.rglob("*.py")) + list((pkg / "csrc").iterdir()) + list((pkg / "host").iterdir()):
        if src.is_file():
            text = src.read_text(errors="ignore")
            for banned in ("wave_oracle", "libwaveoracle", "abi_double", "from oracle", "import oracle", "oracle/"):
                assert banned not in text, f"{src.name} refers to {banned}"


@pytest.mark.parametrize("nx,ny,r", [(1, 1, 1), (1, 1, 2), (2, 1, 2), (1, 3, 2), (3, 2, 1), (5, 4, 2), (7, 3, 1),
                                     (4, 6, 2), (16, 9, 2), (9, 16, 1)])
def test_closed_form_numbering_is_first_touch(nx, ny, r):
    """DoFHandler::distribute_dofs numbering: closed form == oracle's cell walk."""
    o = O.Oracle.from_params(problem("standing-mode-wsol", Nel=f"{nx}, {ny}", R=r))
    assert np.array_equal(cell_dofs(nx, ny, r), o.cell_dofs())


def _expr(block):
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    L = api.lib()
    L.wave_expr_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    L.wave_expr_value.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    L.wave_expr_value.restype = C.c_double
    L.wave_expr_destroy.argtypes = [C.c_void_p]
    L.wave_expr_destroy.restype = None
    rc = L.wave_expr_create(block["Function expression"].encode(), block["Variable names"].encode(),
                            block["Function constants"].encode(), C.byref(h), err, 512)
    return rc, h, err.value.decode()


@pytest.mark.parametrize("name", NAMES)
def test_expressions_match_oracle_evaluator(name):
    p = problem(name, Nel=2)
    o = O.Oracle.from_params(p)
    rng = np.random.default_rng(11)
    L = api.lib()
    for i, blk in enumerate(O.EXPR_NAMES):
        if blk not in p:
            continue
        rc, h, err = _expr(p[blk])
        assert rc == 0, err
        for _ in range(40):
            x, y, t = rng.uniform(-1, 3), rng.uniform(-1, 3), rng.uniform(0, 2)
            a = L.wave_expr_value(h, x, y, t)
            b = o.eval(i, x, y, t)
            assert a == pytest.approx(b, rel=1e-13, abs=1e-300), (blk, x, y, t)
        L.wave_expr_destroy(h)


@pytest.mark.parametrize("text,expect", [
    ("2^3^2", 512.0), ("-2^2", -4.0), ("2^-1", 0.5), ("1 - 2 - 3", -4.0), ("8/4/2", 1.0),
    ("if(1<2 && 3>=3, 10, 20)", 10.0), ("if(0 || 0, 1, 2)", 2.0), ("(1<2) + (2<=2) + (3>4)", 2.0),
    ("max(min(3, 5), 4)", 4.0), ("1 ? 7 : 9", 7.0), ("abs(-3)*sign(-2)", -3.0), ("1e-3*1E3", 1.0),
    ("sqrt(16)+exp(0)+cos(0)+tanh(0)+cosh(0)", 7.0), ("pow(2, 10)", 1024.0), ("pi", np.pi), ("k*pi", 4 * np.pi),
])
def test_expression_semantics(text, expect):
    rc, h, err = _expr({"Function expression": text, "Variable names": "x, y, t", "Function constants": "k=4.0"})
    assert rc == 0, err
    assert api.lib().wave_expr_value(h, 0.3, 0.4, 0.5) == pytest.approx(expect, rel=1e-15)
    api.lib().wave_expr_destroy(h)


@pytest.mark.parametrize("text", ["", "sin(pi*x", "foo(x)", "x +* y", "1 +", "if(x, 1)", "q + 1", "x y"])
def test_bad_expressions_are_rejected(text):
    rc, h, err = _expr({"Function expression": text, "Variable names": "x, y", "Function constants": ""})
    assert rc == -2 and err


def test_constants_with_pi():
    rc, h, err = _expr({"Function expression": "a+b+c", "Variable names": "x, y",
                        "Function constants": "a=pi, b= 2.5*pi , c=1e-1"})
    assert rc == 0, err
    assert api.lib().wave_expr_value(h, 0, 0, 0) == pytest.approx(3.5 * np.pi + 0.1, rel=1e-15)


@pytest.mark.parametrize("text,expect", [
    ("2*pi", 2 * np.pi), (".5*pi", 0.5 * np.pi), ("2.5 * PI", 2.5 * np.pi), ("pi", np.pi),
    # outside the reference's pattern [0-9]*\\.?[0-9]+ (src/ParameterReader.cpp:253): std::stod's prefix parse
    ("-2*pi", -2.0), ("1e3*pi", 1000.0), ("2.*pi", 2.0), ("+3*pi", 3.0), ("4", 4.0),
])
def test_constant_multiples_of_pi_follow_the_reference_pattern(text, expect):
    """One 'Function constants' string must mean the same in the library (device functions), the oracle and
    the reference's ParameterReader (the host FunctionParser shares the pattern: host_selftest)."""
    blk = {"Function expression": "k", "Variable names": "x, y", "Function constants": f"k={text}"}
    rc, h, err = _expr(blk)
    assert rc == 0, err
    assert api.lib().wave_expr_value(h, 0, 0, 0) == pytest.approx(expect, rel=1e-15)
    api.lib().wave_expr_destroy(h)
    p = problem("standing-mode-wsol", Nel=2)
    p["U0"] = dict(blk)
    o = O.Oracle.from_params(p)
    assert o.eval(O.EXPR_NAMES.index("U0"), 0.1, 0.2, 0.0) == pytest.approx(expect, rel=1e-15)


@pytest.mark.parametrize("nx,ny,r,nranks", [(8, 8, 1, 2), (8, 8, 2, 3), (5, 9, 2, 4), (16, 8, 1, 8), (7, 2, 2, 2)])
def test_partition_plan_is_consistent(nx, ny, r, nranks):
    o = O.Oracle.from_params(problem("standing-mode-wsol", Nel=f"{nx}, {ny}", R=r))
    rowptr, col = o.csr()
    prev_end = 0
    for rank in range(nranks):
        p = partition_plan(nx, ny, r, rank, nranks)
        assert p.row_begin == prev_end and p.row_end > p.row_begin
        prev_end = p.row_end
        assert p.ghost_lo_begin <= p.row_begin and p.ghost_hi_end >= p.row_end
        cols = col[rowptr[p.row_begin]:rowptr[p.row_end]]
        # every column an owned row touches lies inside the local (ghosted) range
        assert cols.min() >= p.ghost_lo_begin and cols.max() < p.ghost_hi_end
        # halos are whole blocks owned by the direct neighbours
        if rank > 0:
            lo = partition_plan(nx, ny, r, rank - 1, nranks)
            assert lo.row_begin <= p.ghost_lo_begin < lo.row_end == p.row_begin
        if rank < nranks - 1:
            hi = partition_plan(nx, ny, r, rank + 1, nranks)
            assert hi.row_begin == p.row_end < p.ghost_hi_end <= hi.row_end
    assert prev_end == o.n


def test_multigrid_plan_over_strips():
    """Level planning of the V-cycle (host logic of the multi-GPU multigrid): the hierarchy follows the
    stiffness criterion s c^2 / (dx dy) > 1/4, a P2 problem starts with P1 on the same mesh, and with several
    ranks coarsening stops where a strip would lose whole coarse quad rows -- the same plan on every rank."""
    s = 0.25 * 0.1 ** 2                              # Newmark beta dt^2 at dt = 0.1
    assert api.mg_plan(64, 64, 1, s) == [(32, 32), (16, 16), (8, 8)]
    assert api.mg_plan(64, 64, 2, s) == [(64, 64), (32, 32), (16, 16), (8, 8)]
    assert api.mg_plan(64, 64, 1, 1e-9) == []        # mass dominated: Jacobi sweeps on the fine level only
    assert api.mg_plan(48, 36, 1, 1.0) == [(24, 18), (12, 9)]   # odd ny stops the halving
    # the oracle builds its hierarchy independently (oracle/wave_oracle.c mg_setup): same number of levels
    for nel, r, dt in (("64", 1, 0.1), ("32", 2, 0.05), ("48, 24", 2, 0.05), ("40", 1, 0.05), ("36, 12", 1, 0.05)):
        p = problem("standing-mode-wsol", Nel=nel, R=r, Dt=str(dt))
        o = O.Oracle.from_params(p)
        o.set_cg(precond=2)
        o.newmark_init(dt, 0.25, 0.5)
        nx, ny = api.parse_nel(nel)
        assert o.mg_levels() == 1 + len(api.mg_plan(nx, ny, r, 0.25 * dt * dt)), (nel, r, dt)
    # several ranks: 64 quad rows over 8 ranks = 8 per rank -> 4 -> 2, then a strip would drop to one coarse row
    assert api.mg_plan(64, 64, 1, s, nranks=8) == [(32, 32), (16, 16)]
    assert api.mg_plan(64, 64, 1, s, nranks=2) == [(32, 32), (16, 16), (8, 8)]
    assert api.mg_plan(64, 128, 1, s, nranks=8, box=(0.0, 1.0, 0.0, 2.0)) == [(32, 64), (16, 32), (8, 16)]
    # strips that do not start at even quad rows cannot be coarsened at all: 30 rows over 4 ranks = 7, 8, 7, 8
    assert api.mg_plan(64, 30, 1, 1.0, nranks=4) == []
    for nranks in (2, 3, 4, 8):
        plans = {tuple(api.mg_plan(96, 96, 2, 1.0, nranks=nranks, rank=rk)) for rk in range(nranks)}
        assert len(plans) == 1
        for nx, ny in next(iter(plans))[1:]:
            for rk in range(nranks):  # every strip of every coarse level holds whole quad rows, at least two
                pl = partition_plan(nx, ny, 1, rk, nranks)
                assert pl.quad_row_end - pl.quad_row_begin >= 2


def _run(exe, param_path, cwd):
    return subprocess.run([str(BIN / exe), str(param_path)], cwd=cwd, capture_output=True, text=True, timeout=120)


@pytest.mark.parametrize("exe", ["main-newmark", "main-theta"])
def test_cli_error_behaviour(exe, tmp_path):
    (tmp_path / "build").mkdir()
    (tmp_path / "parameters").mkdir()
    bad_num = {**problem("standing-mode-wsol"), "R": "abc"}
    write_json(tmp_path / "parameters" / "bad-num.json", bad_num)
    r = _run(exe, "../parameters/bad-num.json", tmp_path / "build")
    assert r.returncode == 1 and "Error while parsing parameters/functions" in r.stdout
    bad_expr = problem("standing-mode-wsol")
    bad_expr["G"]["Function expression"] = ""
    write_json(tmp_path / "parameters" / "bad-expr.json", bad_expr)
    r = _run(exe, "../parameters/bad-expr.json", tmp_path / "build")
    assert r.returncode == 1 and "Function expression for 'G' must be specified" in r.stdout
    no_sol = problem("gaussian-pulse")  # Solution block absent: allowed
    no_sol["Theta"] = "1.5"  # outside Patterns::Double(0, 1)
    write_json(tmp_path / "parameters" / "range.json", no_sol)
    r = _run(exe, "../parameters/range.json", tmp_path / "build")
    assert r.returncode == 1
    r = _run(exe, "../parameters/missing.json", tmp_path / "build")
    assert r.returncode == 1 and "Unexpected error while parsing parameters" in r.stdout


def test_host_selftest_binary(tmp_path):
    """clean_double, JSON reader, get_nel/get_geometry, constants, FunctionParser: C++ checks, CPU only."""
    r = subprocess.run([str(BIN / "host_selftest"), str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL OK" in r.stdout and "FAIL " not in r.stdout


def test_mpirun_shim_drops_mpi_options():
    """No GPU here, so -np 16 becomes one process: the program runs once with its own arguments."""
    shim = ROOT / "tools" / "mpirun-shim"
    r = subprocess.run([str(shim), "-np", "16", "--bind-to", "core", "--map-by", "socket", "--hostfile", "/tmp/x",
                        "--mca", "btl", "self,vader", "echo", "binary", "params.json"],
                       capture_output=True, text=True, timeout=30)
    assert r.returncode == 0 and r.stdout.split() == ["binary", "params.json"]
    assert "starting 1" in r.stderr
    assert subprocess.run([str(shim), "-np", "2"], capture_output=True, timeout=30).returncode == 2  # no program


def _launch(args, **env):
    import os

    return subprocess.run([str(BIN / "wave-mpirun"), *args], capture_output=True, text=True, timeout=60,
                          env=dict(os.environ, **env))


def test_launcher_starts_ranks_and_shares_the_communicator_id():
    """`mpirun -np P` of the reference (README.md:113-114) -> bin/wave-mpirun: P processes with
    WAVE_RANK / WAVE_NRANKS / WAVE_LOCAL_RANK, the 128-byte id published by rank 0 read by all."""
    r = _launch(["-np", "3", "--bind-to", "core", "-x", "FOO=bar", str(BIN / "host_selftest"), "--rendezvous"],
                WAVE_LAUNCH_MAX_RANKS="3")
    assert r.returncode == 0, r.stdout + r.stderr
    lines = sorted(r.stdout.strip().splitlines())
    expect = sum((7 * k + 3) % 256 * (k + 1) for k in range(128))
    assert lines == [f"rank {k} of 3 local {k} via WAVE_* id {expect}" for k in range(3)]
    # more ranks than the cap: clamped, with a note
    r = _launch(["-np", "8", str(BIN / "host_selftest"), "--rendezvous"], WAVE_LAUNCH_MAX_RANKS="2")
    assert r.returncode == 0 and len(r.stdout.strip().splitlines()) == 2 and "starting 2" in r.stderr
    # -x VAR=value reaches the ranks
    r = _launch(["-np", "2", "-x", "FOO=bar", "sh", "-c", "echo $WAVE_RANK:$FOO"], WAVE_LAUNCH_MAX_RANKS="2")
    assert sorted(r.stdout.split()) == ["0:bar", "1:bar"]


def test_launcher_returns_the_first_failure_and_stops_the_other_ranks():
    import time

    t0 = time.time()
    r = _launch(["-np", "2", "sh", "-c", "if [ $WAVE_RANK = 1 ]; then exit 3; else exec sleep 30; fi"],
                WAVE_LAUNCH_MAX_RANKS="2")
    assert r.returncode == 3 and time.time() - t0 < 20


def test_rank_variables_of_other_launchers(tmp_path):
    """Open MPI / MPICH / Slurm / torchrun rank variables are understood without an MPI library."""
    import os

    exe, rdv = str(BIN / "host_selftest"), str(tmp_path / "rdv")
    for fam, (rk, sz, loc) in {"OMPI_COMM_WORLD_*": ("OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE",
                                                     "OMPI_COMM_WORLD_LOCAL_RANK"),
                               "PMI_*": ("PMI_RANK", "PMI_SIZE", "MPI_LOCALRANKID"),
                               "SLURM_*": ("SLURM_PROCID", "SLURM_NTASKS", "SLURM_LOCALID"),
                               "RANK/WORLD_SIZE": ("RANK", "WORLD_SIZE", "LOCAL_RANK")}.items():
        env = {k: v for k, v in os.environ.items() if not k.startswith(("WAVE_", "OMPI_", "PMI_", "SLURM_"))}
        env.pop("RANK", None), env.pop("WORLD_SIZE", None), env.pop("LOCAL_RANK", None)
        env["WAVE_RENDEZVOUS"] = rdv
        if fam in ("SLURM_*", "RANK/WORLD_SIZE"):  # not launcher-specific: only on request
            alone = subprocess.run([exe, "--rendezvous"], env={**env, rk: "1", sz: "2", loc: "1"},
                                   capture_output=True, text=True, timeout=30)
            assert alone.returncode == 0 and alone.stdout.startswith("rank 0 of 1 ")
            env["WAVE_LAUNCHER"] = "slurm" if fam == "SLURM_*" else "torchrun"
        p1 = subprocess.Popen([exe, "--rendezvous"], env={**env, rk: "1", sz: "2", loc: "1"},
                              stdout=subprocess.PIPE, text=True)
        r0 = subprocess.run([exe, "--rendezvous"], env={**env, rk: "0", sz: "2", loc: "0"}, capture_output=True,
                            text=True, timeout=30)
        out1, _ = p1.communicate(timeout=30)
        assert r0.returncode == 0 and p1.returncode == 0
        assert r0.stdout.startswith(f"rank 0 of 2 local 0 via {fam} id ")
        assert out1.startswith(f"rank 1 of 2 local 1 via {fam} id ") and out1.split()[-1] == r0.stdout.split()[-1]
        os.remove(rdv)
    env = {k: v for k, v in os.environ.items() if not k.startswith("WAVE_")}
    r = subprocess.run([exe, "--rendezvous"], env={**env, "RANK": "5", "WORLD_SIZE": "2", "WAVE_LAUNCHER": "torchrun"},
                       capture_output=True, text=True, timeout=30)
    assert r.returncode == 1 and "outside WORLD_SIZE=2" in r.stdout


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times): one JSON line with the contract keys."""
    import json
    import os
    import sys

    env = dict(os.environ, WAVE_BENCH_NEL="48", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "dof_steps_per_sec" and line["unit"] == "DoF-steps/s"
    assert line["value"] > 0 and line["steps"] == 2 and line["higher_is_better"] is True
    assert line["e2e"] == {"value": line["value"], "unit": "DoF-steps/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    # the default workload is the north star's P2 Newmark run (WAVE_BENCH_NEL shrinks the mesh for this test)
    assert line["config"]["workload"] == "newmark-4096-p2" and line["config"]["n_dofs"] == 97 * 97
    assert line["warmup"] == 3 and line["run"]["steps_timed"] == 2
    # ranks other than 0 print nothing and exit 0 (torchrun launch of the reference arm)
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1"],
                       capture_output=True, text=True, timeout=60, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.parametrize("nx,ny,r", [(1, 1, 2), (3, 2, 2), (5, 7, 2), (8, 4, 1), (16, 5, 2)])
def test_storage_numbering_is_a_blockwise_permutation(nx, ny, r):
    """The internal (kind-major) numbering permutes DoFs only inside a block [block_start(j),
    block_start(j+1)): strips and halos are the same index ranges in both numberings."""
    canon = cell_dofs(nx, ny, r)
    stor = api.cell_dofs_storage(nx, ny, r)
    n = (nx + 1) * (ny + 1) if r == 1 else (2 * nx + 1) * (2 * ny + 1)
    c2s = np.full(n, -1, dtype=np.int64)
    for c, s in zip(canon.ravel(), stor.ravel()):
        assert c2s[c] in (-1, s)  # one storage index per canonical DoF, the same in every cell
        c2s[c] = s
    assert sorted(c2s.tolist()) == list(range(n))  # a permutation of all DoFs
    if r == 1:
        assert np.array_equal(c2s, np.arange(n))
    starts = [partition_plan(nx, ny, r, k, ny).row_begin for k in range(ny)] + [n]  # one block per quad row
    for b0, b1 in zip(starts[:-1], starts[1:]):
        assert sorted(c2s[b0:b1].tolist()) == list(range(b0, b1))


@pytest.mark.parametrize("n1d,nq,degree", [(2, 4, 3), (3, 7, 5), (4, 15, 7)])
def test_library_quadrature_tables(n1d, nq, degree):
    """The library integrates with the same QGaussSimplex<2>(n) tables as the oracle (deal.II >= 9.4:
    Hillion 4-point, Hammer-Marlowe-Stroud 7-point, Witherden-Vincent 15-point), in the same point
    order, each exact to its degree (src/WaveEquationBase.cpp:82,371)."""
    from math import factorial as f

    xi, eta, w = api.quadrature(n1d)
    assert len(w) == nq and np.all(w > 0) and np.all(xi > 0) and np.all(eta > 0) and np.all(xi + eta < 1)
    oxi, oeta, ow = np.zeros(16), np.zeros(16), np.zeros(16)
    dp = C.POINTER(C.c_double)
    assert O.lib().oracle_get_quadrature(n1d, oxi.ctypes.data_as(dp), oeta.ctypes.data_as(dp),
                                         ow.ctypes.data_as(dp)) == nq
    assert np.abs(xi - oxi[:nq]).max() < 1e-16 and np.abs(eta - oeta[:nq]).max() < 1e-16
    assert np.abs(w - ow[:nq]).max() < 1e-17
    for a in range(degree + 1):
        for b in range(degree + 1 - a):
            assert (w * xi ** a * eta ** b).sum() == pytest.approx(f(a) * f(b) / f(a + b + 2), abs=1e-15)
    with pytest.raises(api.WaveError):
        api.quadrature(5)


def test_hillion_rule_is_the_collapsed_gauss_product():
    """n = 2 (assembly and forcing at R = 1): eta = Gauss-Jacobi(1,0) nodes (4 -+ sqrt 6)/10 with weights
    (9 +- sqrt 6)/36, xi = (1 - eta)(1 -+ 1/sqrt 3)/2, in deal.II's point order."""
    xi, eta, w = api.quadrature(2)
    s6, s3 = 6.0 ** 0.5, 3.0 ** 0.5
    e = [(4 - s6) / 10, (4 + s6) / 10]
    ww = [(9 + s6) / 72, (9 - s6) / 72]
    expect = [((1 - e[0]) * (1 - 1 / s3) / 2, e[0], ww[0]), ((1 - e[1]) * (1 - 1 / s3) / 2, e[1], ww[1]),
              ((1 - e[0]) * (1 + 1 / s3) / 2, e[0], ww[0]), ((1 - e[1]) * (1 + 1 / s3) / 2, e[1], ww[1])]
    got = np.stack([xi, eta, w], axis=1)
    assert np.abs(got - np.array(expect)).max() < 2e-16


def test_vtu_writer_round_trip(tmp_path):
    """solution_NNNN.0.vtu / solution_NNNN.pvtu as DataOut::write_vtu_with_pvtu_record names them
    (src/WaveEquationBase.cpp:363-364): unshared corner points, triangles, point data."""
    import xml.etree.ElementTree as ET

    r = subprocess.run([str(BIN / "host_selftest"), "--vtu", str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout
    assert r.stdout.split() == ["solution_0007.0.vtu", "solution_12345.pvtu"]
    pts, conn, offs, types, data = read_vtu(tmp_path / "solution_0007.0.vtu")
    assert pts.tolist() == [[0, 0, 0], [1, 0, 0], [0, 2, 0], [1, 2, 0], [0, 2, 0], [1, 0, 0]]
    assert conn.tolist() == list(range(6)) and offs.tolist() == [3, 6] and types.tolist() == [5, 5]
    assert data["u"].tolist() == [0.5, 1.5, 2.5, 3.5, 2.5, 1.5] and data["partitioning"].tolist() == [0, 0, 0, 1, 1, 1]
    rec = ET.parse(tmp_path / "solution_0007.pvtu").getroot()
    assert rec.attrib["type"] == "PUnstructuredGrid"
    assert [p.attrib["Source"] for p in rec.findall("PUnstructuredGrid/Piece")] == ["solution_0007.0.vtu"]
    assert [a.attrib["Name"] for a in rec.findall("PUnstructuredGrid/PPointData/PDataArray")] == ["u", "partitioning"]
    assert not (tmp_path / "bad.vtu").exists() or (tmp_path / "bad.vtu").stat().st_size == 0


def test_bench_cpu_sample_is_a_squeezed_strip_of_the_workload():
    """bench.py's bounded CPU sample: same dx, dy and Dt on a strip of the mesh, functions squeezed in y so
    the data stay compatible with the strip's boundary (only the stand-alone variable y is rewritten)."""
    import sys

    sys.path.insert(0, str(ROOT))
    import bench

    p, scheme, scaling = bench.make_params("newmark-4096-p2", 1)
    q, what = bench.cpu_sample_params(p)
    assert scaling == "strong" and q["Nel"] == "4096, 256" and q["Dt"] == p["Dt"]
    x0, x1, y0, y1 = api.parse_geometry(q["Geometry"])
    assert (x0, x1, y0) == (0.0, 1.0, 0.0) and y1 == pytest.approx(1.0 / 16.0)
    assert q["U0"]["Function expression"] == "sin(pi*x)*sin(pi*(0.0 + 16.0*(y - 0.0)))"
    assert p["U0"]["Function expression"] == "sin(pi*x)*sin(pi*y)"          # the workload itself is untouched
    p4, _, _ = bench.make_params("c4-ricker-be-4096-p2", 1)
    q4, _ = bench.cpu_sample_params(p4)
    assert "ys" in q4["F"]["Function expression"] and "(0.0 + 16.0*(y - 0.0))-ys" in q4["F"]["Function expression"]
    small, what_small = bench.cpu_sample_params(bench.make_params("c2-standing-newmark-1024-p1", 1)[0])
    assert what_small == "the whole workload" and small["Nel"] == "1024"
    assert bench.full_n_dofs(p) == 67125249
