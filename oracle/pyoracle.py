"""Python access to the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under citcomcu_b200/ does.

* `run_harness` executes oracle/_ref/ref_harness (the unmodified reference behind our
  driver) on an input file and returns the binary dumps it wrote.
* `Restate` wraps oracle/_ref/libccu_restate.so (oracle/restate.c, the plain-C
  restatement + the 8-colour smoother model).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFDIR = HERE / "_ref"
_DT = {"f64": np.float64, "f32": np.float32, "i32": np.int32, "u32": np.uint32}


def build(quiet: bool = True) -> None:
    """Compile the restatement, and the reference-backed oracle when /root/reference exists."""
    subprocess.run(["make", "-C", str(HERE), "-j8", "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have_ref() -> bool:
    return (REFDIR / "ref_harness").exists() and (REFDIR / "libcitcom_ref.so").exists()


def have_restate() -> bool:
    return (REFDIR / "libccu_restate.so").exists()


class Dump:
    """Lazy reader of a ref_harness dump directory (one rank)."""

    def __init__(self, path, rank: int = 0):
        self.path = Path(path)
        self.rank = rank
        self.entries = {}
        for line in (self.path / f"manifest.r{rank}.txt").read_text().split("\n"):
            if line.strip():
                name, dt, cnt = line.split()
                self.entries[name] = (dt, int(cnt))
        self._cache = {}

    def __contains__(self, name):
        return name in self.entries

    def __getitem__(self, name):
        if name not in self._cache:
            dt, cnt = self.entries[name]
            a = np.fromfile(self.path / f"{name}.r{self.rank}.bin", dtype=_DT[dt])
            assert a.size == cnt, (name, a.size, cnt)
            self._cache[name] = a
        return self._cache[name]

    @property
    def levmin(self):
        return int(self["meta"][0])

    @property
    def levmax(self):
        return int(self["meta"][1])

    def dims(self, lev):
        d = self[f"L{lev}_dims"]
        return dict(nox=int(d[0]), noy=int(d[1]), noz=int(d[2]), elx=int(d[3]), ely=int(d[4]), elz=int(d[5]),
                    nno=int(d[6]), nel=int(d[7]), neq=int(d[8]), npno=int(d[9]))

    def control(self):
        m, c = self["meta"], self["ctl"]
        return dict(levmin=int(m[0]), levmax=int(m[1]), nproc=(int(m[2]), int(m[3]), int(m[4])),
                    me_loc=(int(m[5]), int(m[6]), int(m[7])), v_steps_low=int(m[8]), v_steps_high=int(m[9]),
                    down_heavy=int(m[10]), up_heavy=int(m[11]), mg_cycle=int(m[12]), p_iterations=int(m[13]),
                    precondition=int(m[14]), augmented_Lagr=int(m[15]), accuracy=float(c[0]), augmented=float(c[1]))


def run_harness(input_text: str, workdir, nsteps: int = 0, kat: bool = False, nproc: int = 1,
                preload: str | None = None, timeout: float = 600.0, quiet: bool = True, setup_only: bool = False,
                marker_kat: bool = False):
    """Run the reference through ref_harness in `workdir`; returns (list of Dump per rank, stderr text)."""
    workdir = Path(workdir)
    (workdir / "out").mkdir(parents=True, exist_ok=True)
    (workdir / "dump").mkdir(parents=True, exist_ok=True)
    (workdir / "in.input").write_text(input_text)
    env = dict(os.environ)
    env["CCU_MPI_NP"] = str(nproc)
    if preload:
        env["LD_PRELOAD"] = preload
    if marker_kat:
        env["CCU_MARKER_KAT"] = "1"          # only the marker known answers (the operator KATs are single-rank)
    if setup_only:
        env["CCU_SETUP_ONLY"] = "1"
    cmd = [str(REFDIR / "ref_harness"), "dump", "in.input", "dump", str(nsteps)] + (["kat"] if kat else [])
    r = subprocess.run(cmd, cwd=workdir, env=env, capture_output=True, text=True, timeout=timeout)
    if r.returncode not in (0, 8):
        raise RuntimeError(f"ref_harness failed rc={r.returncode}\n{r.stderr[-4000:]}")
    return [Dump(workdir / "dump", rank=k) for k in range(nproc)], r.stderr


def run_timing(input_text: str, workdir, nsteps: int, nproc: int = 1, timeout: float = 3600.0):
    """`ref_harness time`: returns list of dicts {step, energy_s, stokes_s} measured by the reference run."""
    workdir = Path(workdir)
    (workdir / "out").mkdir(parents=True, exist_ok=True)
    (workdir / "in.input").write_text(input_text)
    env = dict(os.environ)
    env["CCU_MPI_NP"] = str(nproc)
    r = subprocess.run([str(REFDIR / "ref_harness"), "time", "in.input", str(nsteps)], cwd=workdir, env=env,
                       capture_output=True, text=True, timeout=timeout)
    if r.returncode not in (0, 8):
        raise RuntimeError(f"ref_harness time failed rc={r.returncode}\n{r.stderr[-4000:]}")
    out = []
    for line in r.stdout.split("\n"):
        if line.startswith("CCU_TIME"):
            tok = line.split()
            rec = {"step": int(tok[2])}
            for k, v in zip(tok[3::2], tok[4::2]):
                rec[k] = float(v)
            out.append(rec)
    return out


def run_timezero(input_text: str, workdir, repeats: int, nproc: int = 1, timeout: float = 3600.0, dump_up: bool = False):
    """`ref_harness timezero`: seconds of each of `repeats` step-0 general_stokes_solver calls (zero guess).
    dump_up: every rank also writes its final U, P (tz_U, tz_P, tz_meta) into <workdir>/dump (read them with Dump)."""
    workdir = Path(workdir)
    (workdir / "out").mkdir(parents=True, exist_ok=True)
    (workdir / "in.input").write_text(input_text)
    env = dict(os.environ)
    env["CCU_MPI_NP"] = str(nproc)
    if dump_up:
        (workdir / "dump").mkdir(parents=True, exist_ok=True)
        env["CCU_TZ_DUMP"] = "dump"
    r = subprocess.run([str(REFDIR / "ref_harness"), "timezero", "in.input", str(repeats)], cwd=workdir, env=env,
                       capture_output=True, text=True, timeout=timeout)
    if r.returncode not in (0, 8):
        raise RuntimeError(f"ref_harness timezero failed rc={r.returncode}\n{r.stderr[-4000:]}")
    import re
    global last_pressure_loops        # the reference's own "after (NNN) pressure loops" lines of this run (rank 0, see_convergence)
    last_pressure_loops = [int(m) for m in re.findall(r"after \((\d+)\) pressure loops", r.stderr)]
    return [float(line.split()[4]) for line in r.stdout.split("\n") if line.startswith("CCU_TIME")]


last_pressure_loops = []


# ---------------------------------------------------------------- restatement (ctypes)
class _Level(C.Structure):
    _fields_ = [("nox", C.c_int), ("noy", C.c_int), ("noz", C.c_int),
                ("nno", C.c_int), ("nel", C.c_int), ("neq", C.c_int), ("npno", C.c_int),
                ("node", C.c_void_p), ("k1", C.c_void_p), ("k2", C.c_void_p), ("k3", C.c_void_p),
                ("BI", C.c_void_p), ("BPI", C.c_void_p), ("elt_del", C.c_void_p), ("TWW", C.c_void_p),
                ("MASS", C.c_void_p), ("eco_size", C.c_void_p)]


class _MG(C.Structure):
    _fields_ = [("levmin", C.c_int), ("levmax", C.c_int), ("v_steps_low", C.c_int), ("v_steps_high", C.c_int),
                ("down_heavy", C.c_int), ("up_heavy", C.c_int), ("mg_cycle", C.c_int), ("smoother", C.c_int),
                ("accuracy", C.c_double), ("lev", _Level * 12)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Restate:
    """ctypes view of oracle/restate.c bound to the per-level arrays of a Dump."""

    def __init__(self, dump: Dump, smoother: int = 0, accuracy: float | None = None):
        self.lib = C.CDLL(str(REFDIR / "libccu_restate.so"))
        self.d = dump
        ctl = dump.control()
        self.mg = _MG()
        self.mg.levmin, self.mg.levmax = ctl["levmin"], ctl["levmax"]
        self.mg.v_steps_low, self.mg.v_steps_high = ctl["v_steps_low"], ctl["v_steps_high"]
        self.mg.down_heavy, self.mg.up_heavy, self.mg.mg_cycle = ctl["down_heavy"], ctl["up_heavy"], ctl["mg_cycle"]
        self.mg.smoother = smoother
        self.mg.accuracy = ctl["accuracy"] if accuracy is None else accuracy
        self._keep = []
        for lev in range(ctl["levmin"], ctl["levmax"] + 1):
            dm = dump.dims(lev)
            L = self.mg.lev[lev]
            for k in ("nox", "noy", "noz", "nno", "nel", "neq", "npno"):
                setattr(L, k, dm[k])
            for fld, name in (("node", "NODE"), ("k1", "Eqn_k1"), ("k2", "Eqn_k2"), ("k3", "Eqn_k3"), ("BI", "BI"),
                              ("BPI", "BPI"), ("elt_del", "elt_del"), ("TWW", "TWW"), ("MASS", "MASS"),
                              ("eco_size", "eco_size")):
                a = np.ascontiguousarray(dump[f"L{lev}_{name}"])
                self._keep.append(a)
                setattr(L, fld, a.ctypes.data)
        self.lib.ccu_r_vdot.restype = C.c_double
        self.lib.ccu_r_pdot.restype = C.c_double
        self.lib.ccu_r_multi_grid.restype = C.c_double
        self.lib.ccu_r_solve_Ahat_p_fhat.restype = C.c_float

    def L(self, lev):
        return C.byref(self.mg.lev[lev])

    def neq(self, lev):
        return self.mg.lev[lev].neq

    def matvec(self, lev, u, strip=1, mc=False):
        Au = np.zeros(self.neq(lev) + 2)
        u = np.ascontiguousarray(u, dtype=np.float64)
        (self.lib.ccu_r_mc_matvec if mc else self.lib.ccu_r_matvec)(self.L(lev), _p(u), _p(Au), C.c_int(strip))
        return Au[:self.neq(lev)]

    def set_tile(self, ti, tj, tk):
        """Tile shape (colour cells) of the tile-ordered smoother model (ccu_r_ordered_gs mode 9; smoother=19)."""
        t = (C.c_int * 3).in_dll(self.lib, "g_ccu_r_tile")
        t[0], t[1], t[2] = ti, tj, tk

    def set_col(self, ti, tj):
        """Column shape (nodes in y, x) of the column-ordered smoother model (ccu_r_ordered_gs mode 10; smoother=20)."""
        t = (C.c_int * 2).in_dll(self.lib, "g_ccu_r_col")
        t[0], t[1] = ti, tj

    def gauss_seidel(self, lev, F, cycles, guess, d0=None, mc=False, tile=None, col=None):
        if col is not None:
            self.set_col(*col)
            n = self.neq(lev)
            d = np.zeros(n + 2)
            if d0 is not None:
                d[:n] = d0
            Ad = np.zeros(n + 2)
            F = np.ascontiguousarray(F, dtype=np.float64)
            self.lib.ccu_r_ordered_gs(self.L(lev), _p(d), _p(F), _p(Ad), C.c_int(cycles), C.c_int(guess), C.c_int(10))
            return d[:n], Ad[:n]
        if tile is not None:
            self.set_tile(*tile)
            n = self.neq(lev)
            d = np.zeros(n + 2)
            if d0 is not None:
                d[:n] = d0
            Ad = np.zeros(n + 2)
            F = np.ascontiguousarray(F, dtype=np.float64)
            self.lib.ccu_r_ordered_gs(self.L(lev), _p(d), _p(F), _p(Ad), C.c_int(cycles), C.c_int(guess), C.c_int(9))
            return d[:n], Ad[:n]
        n = self.neq(lev)
        d = np.zeros(n + 2)
        if d0 is not None:
            d[:n] = d0
        Ad = np.zeros(n + 2)
        F = np.ascontiguousarray(F, dtype=np.float64)
        fn = self.lib.ccu_r_mc_gauss_seidel if mc else self.lib.ccu_r_gauss_seidel
        fn(self.L(lev), _p(d), _p(F), _p(Ad), C.c_int(cycles), C.c_int(guess))
        return d[:n], Ad[:n]

    def project_vector(self, lev, AU):
        out = np.zeros(self.neq(lev - 1) + 2)
        AU = np.ascontiguousarray(AU, dtype=np.float64)
        self.lib.ccu_r_project_vector(self.L(lev), self.L(lev - 1), _p(AU), _p(out))
        return out[:self.neq(lev - 1)]

    def interp_vector(self, lev, AD):
        out = np.zeros(self.neq(lev + 1) + 2)
        AD = np.ascontiguousarray(AD, dtype=np.float64)
        self.lib.ccu_r_interp_vector(self.L(lev), self.L(lev + 1), _p(AD), _p(out))
        return out[:self.neq(lev + 1)]

    def div_u(self, lev, U):
        out = np.zeros(self.mg.lev[lev].npno)
        U = np.ascontiguousarray(U, dtype=np.float64)
        self.lib.ccu_r_div_u(self.L(lev), _p(U), _p(out))
        return out

    def grad_p(self, lev, P):
        out = np.zeros(self.neq(lev) + 2)
        P = np.ascontiguousarray(P, dtype=np.float64)
        self.lib.ccu_r_grad_p(self.L(lev), _p(P), _p(out))
        return out[:self.neq(lev)]

    def vdot(self, lev, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        return self.lib.ccu_r_vdot(self.L(lev), _p(a), _p(b))

    def conj_grad(self, lev, F, acc, cycles):
        n = self.neq(lev)
        d0 = np.zeros(n + 2)
        F = np.ascontiguousarray(F, dtype=np.float64)
        cyc = C.c_int(cycles)
        self.lib.ccu_r_conj_grad.restype = C.c_double
        res = self.lib.ccu_r_conj_grad(self.L(lev), _p(d0), _p(F), C.c_double(acc), C.byref(cyc))
        return d0[:n], res, cyc.value

    def multi_grid(self, F):
        n = self.neq(self.mg.levmax)
        d1 = np.zeros(n + 2)
        Fw = np.zeros(n + 2)
        Fw[:n] = F
        res = self.lib.ccu_r_multi_grid(C.byref(self.mg), _p(d1), _p(Fw), C.c_double(1e-30))
        return d1[:n], Fw[:n], res

    def solve_del2_u(self, F, acc=1e-30):
        n = self.neq(self.mg.levmax)
        d0 = np.zeros(n + 2)
        F = np.ascontiguousarray(F, dtype=np.float64)
        cyc = C.c_int(0)
        valid = self.lib.ccu_r_solve_del2_u(C.byref(self.mg), _p(d0), _p(F), C.c_double(acc), C.byref(cyc))
        return d0[:n], valid, cyc.value

    def solve_Ahat_p_fhat(self, V, P, F, imp, steps_max):
        n = self.neq(self.mg.levmax)
        Vw = np.zeros(n + 2)
        Vw[:n] = V
        Pw = np.array(P, dtype=np.float64)
        F = np.ascontiguousarray(F, dtype=np.float64)
        steps = C.c_int(steps_max)
        hist = np.zeros((steps_max + 1, 4))
        self.lib.ccu_r_solve_Ahat_p_fhat(C.byref(self.mg), _p(Vw), _p(Pw), _p(F), C.c_double(imp), C.byref(steps), _p(hist))
        return Vw[:n], Pw, steps.value, hist[:steps.value]
