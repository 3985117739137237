from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent.parent
_LIB = _PKG / "lib" / "libwavegpu.so"

EXPR_NAMES = ("C", "F", "U0", "V0", "G", "DGDT", "Solution")
VEC_U, VEC_V, VEC_A, VEC_RHS = 0, 1, 2, 3
MAT_M, MAT_K, MAT_SYS1, MAT_SYS2 = 0, 1, 2, 3
SCHEME_NEWMARK, SCHEME_THETA = 0, 1
FLAG_FORCING_EVERY_STEP = 1
FLAG_NO_STENCIL = 2
PRECOND_JACOBI, PRECOND_NONE, PRECOND_MG = 0, 1, 2

STATUS = {0: "WAVE_OK", -1: "WAVE_ERR_ARG", -2: "WAVE_ERR_EXPR", -3: "WAVE_ERR_CUDA", -4: "WAVE_ERR_STATE",
          -5: "WAVE_ERR_NOCONV", -6: "WAVE_ERR_DIVERGED", -7: "WAVE_ERR_UNSUPPORTED"}


class WaveConfig(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("x0", C.c_double), ("x1", C.c_double),
                ("y0", C.c_double), ("y1", C.c_double), ("r", C.c_int32), ("scheme", C.c_int32),
                ("dt", C.c_double), ("theta", C.c_double), ("beta", C.c_double), ("gamma", C.c_double),
                ("cg_maxit", C.c_int32), ("cg_tol", C.c_double), ("cg_reduce", C.c_double),
                ("precond", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("device", C.c_int32),
                ("nccl_unique_id", C.c_void_p), ("flags", C.c_uint32), ("stream", C.c_void_p)]


class WavePartition(C.Structure):
    _fields_ = [("quad_row_begin", C.c_int32), ("quad_row_end", C.c_int32), ("row_begin", C.c_int64),
                ("row_end", C.c_int64), ("ghost_lo_begin", C.c_int64), ("ghost_hi_end", C.c_int64)]


class WaveError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


def library_path() -> Path:
    return _LIB


_lib = None


def lib():
    """Load libwavegpu.so.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not _LIB.exists():  # build in-tree (nvcc cross-compiles sm_100a without a GPU); never falls back to CPU code
            import importlib.util

            spec = importlib.util.spec_from_file_location("wave_build", _PKG / "build.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        if not _LIB.exists():
            raise FileNotFoundError(f"{_LIB} missing: run python nmpde-wave-equation_b200/build.py")
        L = C.CDLL(str(_LIB))
        dp, ip, lp, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_void_p
        L.wave_default_config.argtypes = [C.POINTER(WaveConfig)]
        L.wave_default_config.restype = None
        L.wave_create.argtypes = [C.POINTER(WaveConfig), C.POINTER(vp)]
        L.wave_destroy.argtypes = [vp]
        L.wave_destroy.restype = None
        L.wave_last_error.argtypes = [vp]
        L.wave_last_error.restype = C.c_char_p
        L.wave_comm_unique_id.argtypes = [vp]
        L.wave_set_expr.argtypes = [vp, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p]
        L.wave_eval_expr.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double, dp]
        L.wave_setup.argtypes = [vp]
        L.wave_init.argtypes = [vp]
        L.wave_step.argtypes = [vp, C.c_double, ip, dp]
        L.wave_run.argtypes = [vp, C.c_double, C.c_int32, dp, ip, ip, dp, lp]
        L.wave_step_host.argtypes = [vp, C.c_double, dp, dp, dp, ip, dp]
        L.wave_norms.argtypes = [vp, dp]
        L.wave_energy.argtypes = [vp, dp]
        L.wave_errors.argtypes = [vp, C.c_double, dp]
        L.wave_probe.argtypes = [vp, C.c_double, C.c_double, dp]
        for name in ("wave_n_dofs", "wave_nnz", "wave_n_cells", "wave_local_nnz", "wave_n_boundary_dofs",
                     "wave_launch_count"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = C.c_int64
        L.wave_local_rows.argtypes = [vp, lp]
        L.wave_local_rows.restype = C.c_int64
        L.wave_get_vector.argtypes = [vp, C.c_int, dp, C.c_size_t]
        L.wave_set_vector.argtypes = [vp, C.c_int, dp, C.c_size_t]
        L.wave_get_csr.argtypes = [vp, C.c_int, lp, ip, dp]
        L.wave_get_support_points.argtypes = [vp, dp, dp, C.c_size_t]
        L.wave_get_boundary_dofs.argtypes = [vp, ip, C.c_size_t]
        L.wave_cell_dofs.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, ip]
        L.wave_cell_dofs_storage.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, ip]
        L.wave_quadrature.argtypes = [C.c_int32, dp, dp, dp]
        L.wave_device_count.argtypes = []
        L.wave_cg_fused_active.argtypes = [vp]
        L.wave_operator_info.argtypes = [vp, lp]
        L.wave_spmv.argtypes = [vp, C.c_int, dp, dp, C.c_size_t]
        L.wave_cg.argtypes = [vp, C.c_int, dp, dp, C.c_size_t, ip]
        L.wave_bench_spmv.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp]
        L.wave_bench_cg_iter.argtypes = [vp, C.c_int, C.c_int, dp, dp]
        L.wave_timers_enable.argtypes = [vp, C.c_int]
        L.wave_timers.argtypes = [vp, dp, C.c_int]
        L.wave_cg_stats.argtypes = [vp, dp, C.c_int]
        L.wave_spmv_timing.argtypes = [vp, C.c_int, dp, dp]
        L.wave_kernel_timing.argtypes = [vp, C.c_int, dp, dp]
        L.wave_partition_plan.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.POINTER(WavePartition)]
        L.wave_mg_plan.argtypes = [C.POINTER(WaveConfig), C.c_double, C.c_double, ip, ip, C.c_int32]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def parse_geometry(s):
    m = re.fullmatch(r"\[\s*([-\d\.]+)\s*,\s*([-\d\.]+)\s*\]\s*x\s*\[\s*([-\d\.]+)\s*,\s*([-\d\.]+)\s*\]", s.strip())
    if not m:
        raise ValueError("Invalid Geometry format in parameters.")
    return tuple(float(g) for g in m.groups())


def parse_nel(s):
    toks = [t.strip() for t in str(s).split(",") if t.strip()]
    if len(toks) == 1:
        return int(toks[0]), int(toks[0])
    if len(toks) == 2:
        return int(toks[0]), int(toks[1])
    raise ValueError("Invalid Nel format.")


def partition_plan(nx, ny, r, rank, nranks):
    out = WavePartition()
    rc = lib().wave_partition_plan(nx, ny, r, rank, nranks, C.byref(out))
    if rc:
        raise WaveError(rc, "wave_partition_plan")
    return out


def mg_plan(nx, ny, r, s, c0=1.0, nranks=1, rank=0, box=(0.0, 1.0, 0.0, 1.0)):
    """(nx, ny) of the V-cycle's coarse levels for M + s K on an nx x ny mesh of `box` (host only)."""
    cfg = WaveConfig()
    lib().wave_default_config(C.byref(cfg))
    cfg.nx, cfg.ny, cfg.r = nx, ny, r
    cfg.x0, cfg.x1, cfg.y0, cfg.y1 = box
    cfg.rank, cfg.nranks = rank, nranks
    ox, oy = (C.c_int32 * 16)(), (C.c_int32 * 16)()
    n = lib().wave_mg_plan(C.byref(cfg), s, c0, ox, oy, 16)
    if n < 0:
        raise WaveError(n, "wave_mg_plan")
    return [(ox[k], oy[k]) for k in range(n)]


def cell_dofs(nx, ny, r):
    """Closed-form cell->DoF table through the C ABI (host only)."""
    dpc = 3 if r == 1 else 6
    ncells = 2 * nx * ny
    out = np.empty((ncells, dpc), dtype=np.int32)
    L = lib()
    for c in range(ncells):
        L.wave_cell_dofs(nx, ny, r, c, _ip(out[c]))
    return out


def cell_dofs_storage(nx, ny, r):
    """Cell->DoF table in the internal storage numbering (host only)."""
    dpc = 3 if r == 1 else 6
    ncells = 2 * nx * ny
    out = np.empty((ncells, dpc), dtype=np.int32)
    L = lib()
    for c in range(ncells):
        L.wave_cell_dofs_storage(nx, ny, r, c, _ip(out[c]))
    return out


def quadrature(n1d):
    """(xi, eta, w) of the library's QGaussSimplex<2>(n1d) table (host only)."""
    xi, eta, w = np.zeros(16), np.zeros(16), np.zeros(16)
    nq = lib().wave_quadrature(n1d, _dp(xi), _dp(eta), _dp(w))
    if nq < 0:
        raise WaveError(nq, "unsupported quadrature order")
    return xi[:nq].copy(), eta[:nq].copy(), w[:nq].copy()


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().wave_comm_unique_id(buf)
    if rc:
        raise WaveError(rc, lib().wave_last_error(None).decode())
    return buf.raw


class WaveSolver:
    """Thin handle over wave_ctx, constructed from a reference-schema parameter dictionary."""

    def __init__(self, params, scheme, rank=0, nranks=1, nccl_id: bytes | None = None, device=-1, cg=None,
                 flags=0, stream=None):
        L = lib()
        self.L = L
        cfg = WaveConfig()
        L.wave_default_config(C.byref(cfg))
        cfg.nx, cfg.ny = parse_nel(params["Nel"])
        cfg.x0, cfg.x1, cfg.y0, cfg.y1 = parse_geometry(params["Geometry"])
        cfg.r = int(params["R"])
        cfg.scheme = SCHEME_NEWMARK if scheme == "newmark" else SCHEME_THETA
        cfg.dt = float(params["Dt"])
        cfg.theta = float(params.get("Theta", 0.5))
        cfg.beta = float(params.get("Beta", 0.25))
        cfg.gamma = float(params.get("Gamma", 0.5))
        if cg:
            cfg.cg_maxit = int(cg.get("maxit", 10000))
            cfg.cg_tol = float(cg.get("tol", 1e-12))
            cfg.cg_reduce = float(cg.get("reduce", 1e-6))
            cfg.precond = int(cg.get("precond", 0))
        cfg.rank, cfg.nranks, cfg.device, cfg.flags = rank, nranks, device, flags
        cfg.stream = stream
        self._idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id else None
        cfg.nccl_unique_id = C.cast(self._idbuf, C.c_void_p) if self._idbuf else None
        self.scheme = scheme
        self.dt = cfg.dt
        self.T = float(params.get("T", 1.0))
        h = C.c_void_p()
        rc = L.wave_create(C.byref(cfg), C.byref(h))
        if rc:
            raise WaveError(rc, L.wave_last_error(None).decode())
        self.h = h
        for i, name in enumerate(EXPR_NAMES):
            blk = params.get(name)
            if not blk or not blk.get("Function expression"):
                continue
            self._ck(L.wave_set_expr(self.h, i, blk["Function expression"].encode(),
                                     blk.get("Variable names", "").encode(),
                                     blk.get("Function constants", "").encode()))
        self.has_solution = bool(params.get("Solution", {}).get("Function expression"))
        self._ck(L.wave_setup(self.h))
        self.n = L.wave_n_dofs(self.h)
        self.nnz_local = L.wave_local_nnz(self.h)
        first = C.c_int64()
        self.nown = L.wave_local_rows(self.h, C.byref(first))
        self.row0 = first.value
        self.time = 0.0
        self.step_no = 0

    def _ck(self, rc):
        if rc:
            raise WaveError(rc, self.L.wave_last_error(self.h).decode())

    def cg_fused_active(self):
        return bool(self.L.wave_cg_fused_active(self.h))

    def operator_info(self):
        """{'stencil_rows', 'sell_rows', 'sell_nnz', 'spmv_bytes'} of this rank's operator."""
        out = (C.c_int64 * 4)()
        self._ck(self.L.wave_operator_info(self.h, out))
        return dict(stencil_rows=out[0], sell_rows=out[1], sell_nnz=out[2], spmv_bytes=out[3])

    def close(self):
        if getattr(self, "h", None):
            self.L.wave_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference operations ------------------------------------------------------------
    def init(self):
        self._ck(self.L.wave_init(self.h))
        self.time, self.step_no = 0.0, 0

    def step(self):
        self.time += self.dt
        self.step_no += 1
        its = (C.c_int32 * 2)()
        nrm = (C.c_double * 2)()
        self._ck(self.L.wave_step(self.h, self.time, its, nrm))
        return (its[0], its[1]), (nrm[0], nrm[1])

    def run(self, n_steps):
        t_end = C.c_double()
        done = C.c_int32()
        its = (C.c_int32 * 2)()
        nrm = (C.c_double * 2)()
        tot = C.c_int64()
        rc = self.L.wave_run(self.h, self.time, n_steps, C.byref(t_end), C.byref(done), its, nrm, C.byref(tot))
        self.time = t_end.value
        self.step_no += done.value
        self._ck(rc)
        return done.value, (its[0], its[1]), (nrm[0], nrm[1]), tot.value

    def step_host(self, u, v, a=None):
        self.time += self.dt
        self.step_no += 1
        its = (C.c_int32 * 2)()
        nrm = (C.c_double * 2)()
        self._ck(self.L.wave_step_host(self.h, self.time, _dp(u), _dp(v), _dp(a) if a is not None else None,
                                       its, nrm))
        return (its[0], its[1]), (nrm[0], nrm[1])

    def norms(self):
        out = (C.c_double * 2)()
        self._ck(self.L.wave_norms(self.h, out))
        return out[0], out[1]

    def energy(self):
        out = C.c_double()
        self._ck(self.L.wave_energy(self.h, C.byref(out)))
        return out.value

    def errors(self, t=None):
        out = (C.c_double * 4)()
        self._ck(self.L.wave_errors(self.h, self.time if t is None else t, out))
        return tuple(out)

    def probe(self, x, y):
        out = C.c_double()
        self._ck(self.L.wave_probe(self.h, x, y, C.byref(out)))
        return out.value

    # ---- data ----------------------------------------------------------------------------
    def vector(self, which):
        v = np.empty(self.n)
        self._ck(self.L.wave_get_vector(self.h, which, _dp(v), v.size))
        return v

    def vector_owned(self, which):
        """This rank's owned rows only (canonical order), for multi-rank contexts."""
        v = np.empty(self.nown)
        self._ck(self.L.wave_get_vector(self.h, which, _dp(v), v.size))
        return v

    def set_vector(self, which, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        self._ck(self.L.wave_set_vector(self.h, which, _dp(arr), arr.size))

    def csr(self, which=None):
        rowptr = np.empty(self.nown + 1, dtype=np.int64)
        col = np.empty(self.nnz_local, dtype=np.int32)
        val = np.empty(self.nnz_local) if which is not None else None
        self._ck(self.L.wave_get_csr(self.h, 0 if which is None else which,
                                     rowptr.ctypes.data_as(C.POINTER(C.c_int64)), _ip(col),
                                     _dp(val) if val is not None else None))
        return (rowptr, col, val) if which is not None else (rowptr, col)

    def support_points(self):
        x, y = np.empty(self.n), np.empty(self.n)
        self._ck(self.L.wave_get_support_points(self.h, _dp(x), _dp(y), self.n))
        return x, y

    def boundary_dofs(self):
        nb = self.L.wave_n_boundary_dofs(self.h)
        out = np.empty(nb, dtype=np.int32)
        self._ck(self.L.wave_get_boundary_dofs(self.h, _ip(out), nb))
        return out

    def eval_expr(self, which, x, y, t=0.0):
        out = C.c_double()
        self._ck(self.L.wave_eval_expr(self.h, which, x, y, t, C.byref(out)))
        return out.value

    # ---- kernel-level --------------------------------------------------------------------
    def spmv(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.n)
        self._ck(self.L.wave_spmv(self.h, which, _dp(x), _dp(y), x.size))
        return y

    def cg(self, which, x0, b):
        x = np.array(x0, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        its = C.c_int32()
        self._ck(self.L.wave_cg(self.h, which, _dp(x), _dp(b), x.size, C.byref(its)))
        return x, its.value

    def bench_spmv(self, which, reps=20, flush_l2=False):
        ms, nbytes = C.c_double(), C.c_double()
        self._ck(self.L.wave_bench_spmv(self.h, which, reps, int(flush_l2), C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value

    def bench_cg_iter(self, which, reps=3):
        ms, nbytes = C.c_double(), C.c_double()
        self._ck(self.L.wave_bench_cg_iter(self.h, which, reps, C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value

    def launch_count(self):
        return self.L.wave_launch_count(self.h)

    def timers_enable(self, on=True):
        self._ck(self.L.wave_timers_enable(self.h, int(on)))

    def timers(self, reset=False):
        out = (C.c_double * 6)()
        self._ck(self.L.wave_timers(self.h, out, int(reset)))
        return dict(zip(("rhs", "bc", "cg", "update", "energy", "other"), out))

    def spmv_timing(self, on):
        """Switch live SpMV bracketing on/off; returns (launches, ms_total) accumulated so far."""
        cnt, ms = C.c_double(), C.c_double()
        self._ck(self.L.wave_spmv_timing(self.h, int(on), C.byref(cnt), C.byref(ms)))
        return cnt.value, ms.value

    def kernel_timing(self, on):
        """Live bracketing of the CG kernels; returns ([launches] * 3, [ms] * 3) for SpMV, update, direction."""
        cnt, ms = (C.c_double * 3)(), (C.c_double * 3)()
        self._ck(self.L.wave_kernel_timing(self.h, int(on), cnt, ms))
        return list(cnt), list(ms)

    def cg_stats(self, reset=False):
        out = (C.c_double * 4)()
        self._ck(self.L.wave_cg_stats(self.h, out, int(reset)))
        return dict(zip(("solves", "iterations", "spmv_launches", "ms_total"), out))
