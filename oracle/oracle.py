"""ctypes front-end of the CPU ORACLE (oracle/wave_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under nmpde-wave-equation_b200/ imports it.

The driver loops mirror the reference's run() loops:
  Newmark  src/WaveNewmark.cpp:407-456   (while (time < T) { time += dt; ... })
  theta    src/WaveTheta.cpp:372-411
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libwaveoracle.so"

EXPR_NAMES = ("C", "F", "U0", "V0", "G", "DGDT", "Solution")


def build(force: bool = False) -> Path:
    """Compile the C restatement (gcc, -O3, OpenMP)."""
    src = _HERE / "wave_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        _LIB_PATH.parent.mkdir(exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", str(src), "-lm", "-o", str(_LIB_PATH)]
        )
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        # idle OpenMP threads sleep instead of spinning: thousands of tiny parallel regions (small test
        # meshes) otherwise crawl when the cores are shared with other processes
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        L = C.CDLL(str(_LIB_PATH))
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        lp = C.POINTER(C.c_int64)
        vp = C.c_void_p
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_last_error.argtypes = [vp]
        L.oracle_set_expr.argtypes = [vp, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p]
        L.oracle_eval.restype = C.c_double
        L.oracle_eval.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double]
        L.oracle_setup.argtypes = [vp]
        L.oracle_assemble.argtypes = [vp]
        L.oracle_set_cg.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_int]
        L.oracle_set_forcing_every_step.argtypes = [vp, C.c_int]
        L.oracle_newmark_init.argtypes = [vp, C.c_double, C.c_double, C.c_double]
        L.oracle_newmark_step.argtypes = [vp]
        L.oracle_theta_init.argtypes = [vp, C.c_double, C.c_double]
        L.oracle_theta_step.argtypes = [vp]
        L.oracle_energy.restype = C.c_double
        L.oracle_energy.argtypes = [vp]
        L.oracle_errors.argtypes = [vp, C.c_double, dp]
        L.oracle_probe.restype = C.c_double
        L.oracle_probe.argtypes = [vp]
        L.oracle_mg_levels.argtypes = [vp]
        for name in ("oracle_n", "oracle_nnz", "oracle_nb", "oracle_ncells"):
            getattr(L, name).restype = C.c_int64
            getattr(L, name).argtypes = [vp]
        L.oracle_time.restype = C.c_double
        L.oracle_time.argtypes = [vp]
        L.oracle_last_iterations.argtypes = [vp, ip]
        L.oracle_get_csr.argtypes = [vp, lp, ip]
        L.oracle_get_values.argtypes = [vp, C.c_int, dp]
        L.oracle_get_vector.argtypes = [vp, C.c_int, dp]
        L.oracle_set_vector.argtypes = [vp, C.c_int, dp]
        L.oracle_get_cell_dofs.argtypes = [vp, ip]
        L.oracle_get_support_points.argtypes = [vp, dp, dp]
        L.oracle_get_boundary_dofs.argtypes = [vp, ip]
        L.oracle_spmv.argtypes = [vp, C.c_int, dp, dp]
        L.oracle_cg.argtypes = [vp, C.c_int, dp, dp]
        L.oracle_norm.restype = C.c_double
        L.oracle_norm.argtypes = [vp, C.c_int]
        L.oracle_get_quadrature.argtypes = [C.c_int, dp, dp, dp]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_destroy.argtypes = [vp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OracleError(RuntimeError):
    pass


class Oracle:
    """One problem instance.  `params` is the reference's JSON dictionary
    (parameters/*.json schema, src/ParameterReader.cpp:39-126)."""

    M, K, A, A2, S = 0, 1, 2, 3, 4
    U, V, ACC, RHS = 0, 1, 2, 3

    def __init__(self, nel, geometry, r, exprs):
        L = lib()
        self.L = L
        nx, ny = nel
        (x0, x1), (y0, y1) = geometry
        self.h = L.oracle_create(nx, ny, x0, x1, y0, y1, r)
        if not self.h:
            raise OracleError("invalid mesh / degree")
        self.r = r
        for i, name in enumerate(EXPR_NAMES):
            blk = exprs.get(name)
            if not blk or not blk.get("Function expression"):
                if name == "Solution":
                    continue
                raise OracleError(f"Function expression for '{name}' must be specified")
            rc = L.oracle_set_expr(
                self.h, i, blk["Function expression"].encode(),
                blk.get("Variable names", "").encode(), blk.get("Function constants", "").encode())
            if rc:
                raise OracleError(f"{name}: {L.oracle_last_error(self.h).decode()}")
        self.has_solution = bool(exprs.get("Solution", {}).get("Function expression"))
        L.oracle_setup(self.h)
        if L.oracle_assemble(self.h):
            raise OracleError(L.oracle_last_error(self.h).decode())
        self.n = L.oracle_n(self.h)
        self.nnz = L.oracle_nnz(self.h)
        self.nb = L.oracle_nb(self.h)
        self.ncells = L.oracle_ncells(self.h)

    @classmethod
    def from_params(cls, params):
        from .params import parse_geometry, parse_nel  # noqa: WPS433

        return cls(parse_nel(params["Nel"]), parse_geometry(params["Geometry"]), int(params["R"]), params)

    def __del__(self):
        try:
            if self.h:
                self.L.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- configuration ----------------------------------------------------
    def set_cg(self, maxit=10000, tol=1e-12, reduce=1e-6, precond=0):
        self.L.oracle_set_cg(self.h, maxit, tol, reduce, precond)

    def set_forcing_every_step(self, flag: bool):
        self.L.oracle_set_forcing_every_step(self.h, int(flag))

    # -- structure --------------------------------------------------------
    def csr(self):
        rowptr = np.empty(self.n + 1, dtype=np.int64)
        col = np.empty(self.nnz, dtype=np.int32)
        self.L.oracle_get_csr(self.h, rowptr.ctypes.data_as(C.POINTER(C.c_int64)), _ip(col))
        return rowptr, col

    def values(self, which):
        v = np.empty(self.nnz)
        self.L.oracle_get_values(self.h, which, _dp(v))
        return v

    def vector(self, which):
        v = np.empty(self.n)
        self.L.oracle_get_vector(self.h, which, _dp(v))
        return v

    def set_vector(self, which, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        self.L.oracle_set_vector(self.h, which, _dp(arr))

    def cell_dofs(self):
        dpc = 3 if self.r == 1 else 6
        a = np.empty((self.ncells, dpc), dtype=np.int32)
        self.L.oracle_get_cell_dofs(self.h, _ip(a))
        return a

    def support_points(self):
        sx, sy = np.empty(self.n), np.empty(self.n)
        self.L.oracle_get_support_points(self.h, _dp(sx), _dp(sy))
        return sx, sy

    def boundary_dofs(self):
        a = np.empty(self.nb, dtype=np.int32)
        self.L.oracle_get_boundary_dofs(self.h, _ip(a))
        return a

    def spmv(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.n)
        self.L.oracle_spmv(self.h, which, _dp(x), _dp(y))
        return y

    def cg(self, which, x0, b):
        x = np.array(x0, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        its = self.L.oracle_cg(self.h, which, _dp(x), _dp(b))
        return x, its

    def eval(self, which, x, y, t=0.0):
        return self.L.oracle_eval(self.h, which, x, y, t)

    # -- time stepping ----------------------------------------------------
    def newmark_init(self, dt, beta, gamma):
        if self.L.oracle_newmark_init(self.h, dt, beta, gamma):
            raise OracleError("CG failed in a0 solve")

    def theta_init(self, dt, theta):
        self.L.oracle_theta_init(self.h, dt, theta)

    def newmark_step(self):
        if self.L.oracle_newmark_step(self.h):
            raise OracleError("CG failed")

    def theta_step(self):
        if self.L.oracle_theta_step(self.h):
            raise OracleError("CG failed")

    @property
    def time(self):
        return self.L.oracle_time(self.h)

    def iterations(self):
        its = (C.c_int * 2)()
        self.L.oracle_last_iterations(self.h, its)
        return its[0], its[1]

    def mg_levels(self):
        """Levels of the multigrid hierarchy built by the last *_init with precond=2 (0: none)."""
        return int(self.L.oracle_mg_levels(self.h))

    def energy(self):
        return self.L.oracle_energy(self.h)

    def errors(self, t=None):
        out = (C.c_double * 4)()
        if self.L.oracle_errors(self.h, self.time if t is None else t, out):
            raise OracleError("no exact solution")
        return tuple(out)

    def probe(self):
        return self.L.oracle_probe(self.h)

    def norm(self, which):
        return self.L.oracle_norm(self.h, which)


def run(params, scheme, log_every=0, cg=None, max_steps=None):
    """Mirror of WaveNewmark::run / WaveTheta::run (time loop + logging).

    Returns dict(steps, time, energy=[(step,t,E)], error=[(step,t,L2,H1,rL2,rH1)],
    final_errors, iterations=[...], oracle=<Oracle>)."""
    o = Oracle.from_params(params)
    if cg:
        o.set_cg(**cg)
    dt = float(params["Dt"])
    T = float(params["T"])
    if scheme == "newmark":
        o.newmark_init(dt, float(params["Beta"]), float(params["Gamma"]))
        step_fn = o.newmark_step
    elif scheme == "theta":
        o.theta_init(dt, float(params["Theta"]))
        step_fn = o.theta_step
    else:
        raise ValueError(scheme)
    out = {"energy": [], "error": [], "iterations": [], "probe": []}
    time, step = 0.0, 0
    while time < T:  # src/WaveNewmark.cpp:407 -- float accumulation decides the step count
        time += dt
        step += 1
        step_fn()
        nu, nv = o.norm(Oracle.U), o.norm(Oracle.V)
        if not (np.isfinite(nu) and np.isfinite(nv)) or nu > 1e130 or nv > 1e130:
            out["diverged"] = step
            break
        if log_every > 0 and step % log_every == 0:
            out["energy"].append((step, time, o.energy()))
            if o.has_solution:
                out["error"].append((step, time) + o.errors(time))
            out["probe"].append((step, time, o.probe()))
            out["iterations"].append((step, time) + o.iterations())
        if max_steps and step >= max_steps:
            break
    out["steps"] = step
    out["time"] = time
    out["final_errors"] = o.errors(time) if o.has_solution else None
    out["oracle"] = o
    return out
