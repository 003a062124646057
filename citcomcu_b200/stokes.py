"""Host-side mirror of the reference's Stokes hot-path interface, over the C ABI.

Method names, argument meaning and return conventions follow the reference functions they
stand in for (src/prototypes.h): `n_assemble_del2_u`, `gauss_seidel`, `project_vector`,
`interp_vector`, `assemble_div_u`, `assemble_grad_p`, `global_vdot`, `multi_grid`,
`solve_del2_u`, `solve_Ahat_p_fhat`.  Vectors are numpy float64 arrays in the reference's
equation numbering; PyTorch is not involved in the data path (ctypes -> libcitcomcu_b200.so).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ccu_config

VEC = dict(VEL=0, RES=1, RHS=2, FL=3, DEL_VEL=4, AU=5, U=6, F=7, T0=8, T1=9, T2=10)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.c_void_p)


class StokesContext:
    """Device context for one subdomain: `struct All_variables` members the hot path reads."""

    def __init__(self, levmin, levmax, nox, noy, noz, *, v_steps_low=20, v_steps_high=3, down_heavy=3,
                 up_heavy=3, mg_cycle=1, p_iterations=375, accuracy=1e-3, device=0):
        self.lib = _lib.lib()
        cfg = ccu_config()
        cfg.levmin, cfg.levmax = levmin, levmax
        for lev in range(levmin, levmax + 1):
            cfg.nox[lev], cfg.noy[lev], cfg.noz[lev] = nox[lev], noy[lev], noz[lev]
        cfg.v_steps_low, cfg.v_steps_high = v_steps_low, v_steps_high
        cfg.down_heavy, cfg.up_heavy, cfg.mg_cycle = down_heavy, up_heavy, mg_cycle
        cfg.p_iterations, cfg.accuracy, cfg.device = p_iterations, accuracy, device
        self.cfg = cfg
        self.levmin, self.levmax = levmin, levmax
        self.dims = {lev: (nox[lev], noy[lev], noz[lev]) for lev in range(levmin, levmax + 1)}
        self._ctx = C.c_void_p()
        check(self.lib.ccu_create(C.byref(cfg), C.byref(self._ctx)))

    # -- geometry helpers
    def nno(self, lev):
        x, y, z = self.dims[lev]
        return x * y * z

    def neq(self, lev):
        return 3 * self.nno(lev)

    def nel(self, lev):
        x, y, z = self.dims[lev]
        return (x - 1) * (y - 1) * (z - 1)

    def close(self):
        if self._ctx and getattr(self, "_owned", True):
            self.lib.ccu_destroy(self._ctx)
        self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        check(self.lib.ccu_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    OPTIONS = dict(graphs=0, small_nodes=1, warp_nodes=2, quad_nodes=3, lanes_large=4, matvec_tab=5, relax_tab=6, smem_nodes=7,
                   col_nodes=9, relax_col=10, matvec_col=11, col_shape=13, col_wf=14, bottom_cluster=15, full_nodes=18, relax_full=19, matvec_full=20, p2p_halo=21, halo_overlap=22)

    def set_option(self, name, value):
        check(self.lib.ccu_set_option(self._ctx, self.OPTIONS[name], int(value)))

    def get_option(self, name, lev):
        v = C.c_int(0)
        check(self.lib.ccu_get_option(self._ctx, self.OPTIONS[name], int(lev), C.byref(v)))
        return int(v.value)

    def synchronize(self):
        check(self.lib.ccu_synchronize(self._ctx))

    # -- subdomain-per-GPU runs
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.lib().ccu_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nproc, me_loc, unique_id: bytes | None):
        """Collective over all ranks: duplicated-node tables + NCCL communicator (parallel_domain_decomp1 /
        parallel_communication_routs1, Parallel_related.c:80,477).  Must precede the operator construction."""
        uid = None if unique_id is None else C.create_string_buffer(bytes(unique_id), 128)
        check(self.lib.ccu_comm_init(self._ctx, int(nproc[0]), int(nproc[1]), int(nproc[2]), int(me_loc[0]), int(me_loc[1]),
                                     int(me_loc[2]), uid))
        self.nproc, self.me_loc = tuple(nproc), tuple(me_loc)

    def agglomerate(self, agg_lev, global_dims):
        """Replicate the multigrid levels <= agg_lev on every rank (ccu_agglomerate).  Returns a non-owning
        StokesContext for the GLOBAL mesh of those levels: give it the global node flags and coordinates, then
        build_geometry().  `global_dims[lev]` = (nox, noy, noz) of the global mesh."""
        out = C.c_void_p()
        check(self.lib.ccu_agglomerate(self._ctx, int(agg_lev), C.byref(out)))
        g = StokesContext.__new__(StokesContext)
        g.lib, g._ctx, g._owned = self.lib, out, False
        g.levmin, g.levmax = self.levmin, agg_lev
        g.dims = {lev: tuple(global_dims[lev]) for lev in range(self.levmin, agg_lev + 1)}
        self.coarse = g
        return g

    @property
    def launch_count(self) -> int:
        return int(self.lib.ccu_launch_count(self._ctx))

    # -- operator upload (what construct_stiffness_B_matrix leaves in E)
    def set_node_flags(self, lev, node):
        a = np.ascontiguousarray(node, dtype=np.uint32)
        assert a.size == self.nno(lev)
        check(self.lib.ccu_set_node_flags(self._ctx, lev, a.ctypes.data_as(C.c_void_p)))

    def set_stiffness(self, lev, k1, k2, k3, BI):
        ks = [np.ascontiguousarray(k, dtype=np.float32) for k in (k1, k2, k3)]
        for k in ks:
            assert k.size == self.nno(lev) * 42
        b, bp = _f64(BI)
        assert b.size >= self.neq(lev)
        check(self.lib.ccu_set_stiffness(self._ctx, lev, *[k.ctypes.data_as(C.c_void_p) for k in ks], bp))

    def set_pressure_ops(self, lev, elt_del, BPI):
        g = np.ascontiguousarray(elt_del, dtype=np.float32)
        assert g.size == self.nel(lev) * 24
        b, bp = _f64(BPI)
        check(self.lib.ccu_set_pressure_ops(self._ctx, lev, g.ctypes.data_as(C.c_void_p), bp))

    def set_transfer_weights(self, lev, TWW, MASS, eco_size):
        t = np.ascontiguousarray(TWW, dtype=np.float32)
        m = np.ascontiguousarray(MASS, dtype=np.float32)
        e = np.ascontiguousarray(eco_size, dtype=np.float32)
        assert t.size == self.nel(lev) * 8 and m.size == self.nno(lev) and e.size == self.nel(lev) * 3
        check(self.lib.ccu_set_transfer_weights(self._ctx, lev, t.ctypes.data_as(C.c_void_p),
                                                m.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p)))

    def load_operator_from(self, src):
        """Upload every level from an object with mapping access to the reference's arrays
        (`L{lev}_NODE`, `L{lev}_Eqn_k1`, ...), e.g. an oracle dump."""
        for lev in range(self.levmin, self.levmax + 1):
            self.set_node_flags(lev, src[f"L{lev}_NODE"])
            self.set_stiffness(lev, src[f"L{lev}_Eqn_k1"], src[f"L{lev}_Eqn_k2"], src[f"L{lev}_Eqn_k3"], src[f"L{lev}_BI"])
            self.set_transfer_weights(lev, src[f"L{lev}_TWW"], src[f"L{lev}_MASS"], src[f"L{lev}_eco_size"])
        lm = self.levmax
        self.set_pressure_ops(lm, src[f"L{lm}_elt_del"], src[f"L{lm}_BPI"])

    # -- operator construction on the device
    def set_coordinates(self, lev, X1, X2, X3):
        xs = [np.ascontiguousarray(x, dtype=np.float32) for x in (X1, X2, X3)]
        for x in xs:
            assert x.size == self.nno(lev)
        check(self.lib.ccu_set_coordinates(self._ctx, lev, *[x.ctypes.data_as(C.c_void_p) for x in xs]))

    def set_spherical_coordinates(self, lev, theta, phi, r):
        """E->SXX[lev]: switches the context to the regional-spherical element routines; set_coordinates then carries E->XX (Cartesian)."""
        S = [np.ascontiguousarray(a, dtype=np.float32) for a in (theta, phi, r)]
        assert all(a.size == self.nno(lev) for a in S)
        check(self.lib.ccu_set_spherical_coordinates(self._ctx, lev, *[a.ctypes.data_as(C.c_void_p) for a in S]))

    def build_geometry(self):
        check(self.lib.ccu_build_geometry(self._ctx))

    def set_viscosity_law(self, tdepv, rheol, N0, E, T, Z, vmin=0, min_value=0.0, vmax=0, max_value=0.0, smooth_cycles=1):
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (N0, E, T, Z)]
        check(self.lib.ccu_set_viscosity_law(self._ctx, int(tdepv), int(rheol), len(arrs[0]),
                                             *[a.ctypes.data_as(C.c_void_p) for a in arrs], int(vmin), C.c_float(min_value),
                                             int(vmax), C.c_float(max_value), int(smooth_cycles)))

    def set_sdepv(self, on, rheology, expt, trns, misfit=0.001, iter_damp=1.0, max_iter=50, start_from_newtonian=0, trns_T=0.0, trns_c=0.0):
        """Stress-dependent viscosity (visc_from_S, sdepv_rheology 1 / 2) and its outer iteration (Drive_solvers.c:120-159)."""
        e = np.ascontiguousarray(expt, dtype=np.float32); t = np.ascontiguousarray(trns, dtype=np.float32)
        check(self.lib.ccu_set_sdepv(self._ctx, int(on), int(rheology), e.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), C.c_float(misfit),
                                     C.c_float(iter_damp), int(max_iter), int(start_from_newtonian), C.c_float(trns_T), C.c_float(trns_c)))

    def set_cdepv(self, on, pre_comp, layer_pre_comp=0, absolute=0, check_c_irange=0):
        """visc_from_C (prefactor mode): pre_comp = (background, second material) viscosity factors, per layer with layer_pre_comp."""
        pc = np.ascontiguousarray(pre_comp, dtype=np.float32)
        check(self.lib.ccu_set_cdepv(self._ctx, int(on), int(layer_pre_comp), pc.ctypes.data_as(C.c_void_p), int(absolute), int(check_c_irange)))

    def set_composition(self, Cn):
        Cn = np.ascontiguousarray(Cn, dtype=np.float32)
        assert Cn.size == self.nno(self.levmax)
        check(self.lib.ccu_set_composition(self._ctx, Cn.ctypes.data_as(C.c_void_p)))

    def set_bdepv(self, on, abyerlee, bbyerlee, lbyerlee, dimensional=0, length_scale=1.0, tau_scale=1.0, plasticity_trans=1, viscosity_offset=0.0):
        """visc_from_B, regular branch: yield stress a depth + b (capped at l), or (a z[m] + b) l / tau_scale when dimensional."""
        a, b, l = [np.ascontiguousarray(v, dtype=np.float32) for v in (abyerlee, bbyerlee, lbyerlee)]
        check(self.lib.ccu_set_bdepv(self._ctx, int(on), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p),
                                     int(dimensional), C.c_float(length_scale), C.c_float(tau_scale), int(plasticity_trans), C.c_float(viscosity_offset)))

    def sdepv_iterations(self):
        n, m = C.c_int(0), C.c_double(0.0)
        check(self.lib.ccu_get_sdepv_iterations(self._ctx, C.byref(n), C.byref(m)))
        return int(n.value), float(m.value)

    def set_material(self, mat):
        m = np.ascontiguousarray(mat, dtype=np.int32)
        assert m.size == self.nel(self.levmax)
        check(self.lib.ccu_set_material(self._ctx, m.ctypes.data_as(C.c_void_p)))

    def set_temperature(self, T):
        t = np.ascontiguousarray(T, dtype=np.float32)
        assert t.size == self.nno(self.levmax)
        check(self.lib.ccu_set_temperature(self._ctx, t.ctypes.data_as(C.c_void_p)))

    def set_element_viscosity(self, lev, EVI):
        e = np.ascontiguousarray(EVI, dtype=np.float32)
        assert e.size == 8 * self.nel(lev)
        check(self.lib.ccu_set_element_viscosity(self._ctx, lev, e.ctypes.data_as(C.c_void_p)))

    def get_system_viscosity(self):
        check(self.lib.ccu_get_system_viscosity(self._ctx))

    def construct_stiffness_B_matrix(self, augmented_Lagr=1, augmented=1.0e3, precondition=1):
        check(self.lib.ccu_construct_stiffness_B_matrix(self._ctx, int(augmented_Lagr), C.c_double(augmented), int(precondition)))

    def assemble_forces(self, buoyancy=None, want_host=True):
        b = None if buoyancy is None else np.ascontiguousarray(buoyancy, dtype=np.float32)
        out = np.empty(self.neq(self.levmax)) if want_host else None
        check(self.lib.ccu_assemble_forces(self._ctx, None if b is None else b.ctypes.data_as(C.c_void_p),
                                           None if out is None else out.ctypes.data_as(C.c_void_p)))
        return out

    def set_velocity_bcs(self, VB1, VB2, VB3):
        """E->VB: imposed boundary velocities per direction, [nno] in the reference's node order; None, None, None clears them."""
        if VB1 is None and VB2 is None and VB3 is None:
            check(self.lib.ccu_set_velocity_bcs(self._ctx, None, None, None))
            return
        v = [np.ascontiguousarray(a, dtype=np.float32) for a in (VB1, VB2, VB3)]
        assert all(a.size == self.nno(self.levmax) for a in v)
        check(self.lib.ccu_set_velocity_bcs(self._ctx, *[a.ctypes.data_as(C.c_void_p) for a in v]))

    def conform_velocity_bcs(self):
        check(self.lib.ccu_conform_velocity_bcs(self._ctx))

    def get_stiffness(self, lev):
        n = self.nno(lev) * 42
        ks = [np.empty(n, dtype=np.float32) for _ in range(3)]
        BI = np.empty(self.neq(lev))
        check(self.lib.ccu_get_stiffness(self._ctx, lev, *[k.ctypes.data_as(C.c_void_p) for k in ks], BI.ctypes.data_as(C.c_void_p)))
        return ks[0], ks[1], ks[2], BI

    _ARR = dict(TWW=(0, np.float32, lambda s, l: 8 * s.nel(l)), MASS=(1, np.float32, lambda s, l: s.nno(l)),
                eco_size=(2, np.float32, lambda s, l: 3 * s.nel(l)), elt_del=(3, np.float32, lambda s, l: 24 * s.nel(l)),
                BPI=(4, np.float64, lambda s, l: s.nel(l)), EVI=(5, np.float32, lambda s, l: 8 * s.nel(l)))

    def get_level_array(self, lev, name):
        idx, dt, size = self._ARR[name]
        out = np.empty(size(self, lev), dtype=dt)
        check(self.lib.ccu_get_level_array(self._ctx, lev, idx, out.ctypes.data_as(C.c_void_p)))
        return out

    # -- reference-named operators (host vectors in/out)
    def n_assemble_del2_u(self, u, level, strip_bcs=1):
        u, up = _f64(u)
        Au = np.empty(self.neq(level))
        check(self.lib.ccu_n_assemble_del2_u(self._ctx, level, up, Au.ctypes.data_as(C.c_void_p), int(strip_bcs)))
        return Au

    assemble_del2_u = n_assemble_del2_u

    def e_assemble_del2_u(self, u, level, strip_bcs=1):
        u, up = _f64(u)
        Au = np.empty(self.neq(level))
        check(self.lib.ccu_e_assemble_del2_u(self._ctx, level, up, Au.ctypes.data_as(C.c_void_p), int(strip_bcs)))
        return Au

    def conj_grad(self, F, acc, cycles, level):
        """conj_grad (General_matrix_functions.c:661): returns (d0, residual, cycles done)."""
        n = self.neq(level)
        d = np.zeros(n)
        F = np.ascontiguousarray(F, dtype=np.float64)
        cyc = C.c_int(int(cycles))
        res = C.c_double(0.0)
        check(self.lib.ccu_conj_grad(self._ctx, int(level), d.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p),
                                     C.c_double(acc), C.byref(cyc), C.byref(res)))
        return d, res.value, cyc.value

    def gauss_seidel(self, F, cycles, level, guess, d0=None):
        n = self.neq(level)
        d = np.zeros(n) if d0 is None else np.array(d0, dtype=np.float64)
        F, Fp = _f64(F)
        Ad = np.empty(n)
        check(self.lib.ccu_gauss_seidel(self._ctx, level, d.ctypes.data_as(C.c_void_p), Fp,
                                        Ad.ctypes.data_as(C.c_void_p), int(cycles), int(guess)))
        return d, Ad

    def project_vector(self, start_lev, AU):
        AU, p = _f64(AU)
        AD = np.empty(self.neq(start_lev - 1))
        check(self.lib.ccu_project_vector(self._ctx, start_lev, p, AD.ctypes.data_as(C.c_void_p)))
        return AD

    def interp_vector(self, start_lev, AD):
        AD, p = _f64(AD)
        AU = np.empty(self.neq(start_lev + 1))
        check(self.lib.ccu_interp_vector(self._ctx, start_lev, p, AU.ctypes.data_as(C.c_void_p)))
        return AU

    def strip_bcs_from_residual(self, Res, level):
        r = np.array(Res, dtype=np.float64)
        check(self.lib.ccu_strip_bcs_from_residual(self._ctx, level, r.ctypes.data_as(C.c_void_p)))
        return r

    def assemble_div_u(self, U, level):
        U, p = _f64(U)
        out = np.empty(self.nel(level))
        check(self.lib.ccu_assemble_div_u(self._ctx, level, p, out.ctypes.data_as(C.c_void_p)))
        return out

    def assemble_grad_p(self, P, lev):
        P, p = _f64(P)
        out = np.empty(self.neq(lev))
        check(self.lib.ccu_assemble_grad_p(self._ctx, lev, p, out.ctypes.data_as(C.c_void_p)))
        return out

    def global_vdot(self, A, B, lev):
        A, ap = _f64(A)
        B, bp = _f64(B)
        out = C.c_double()
        check(self.lib.ccu_global_vdot(self._ctx, lev, ap, bp, C.byref(out)))
        return out.value

    def global_pdot(self, A, B, lev):
        A, ap = _f64(A)
        B, bp = _f64(B)
        out = C.c_double()
        check(self.lib.ccu_global_pdot(self._ctx, lev, ap, bp, C.byref(out)))
        return out.value

    def multi_grid(self, F):
        """Returns (d1, residual vector, residual norm) like multi_grid's in/out arguments."""
        Fw = np.array(F, dtype=np.float64)
        d1 = np.empty_like(Fw)
        res = C.c_double()
        check(self.lib.ccu_multi_grid(self._ctx, d1.ctypes.data_as(C.c_void_p), Fw.ctypes.data_as(C.c_void_p), C.byref(res)))
        return d1, Fw, res.value

    def solve_del2_u(self, F, acc):
        F, p = _f64(F)
        d0 = np.empty_like(F)
        valid, cyc = C.c_int(), C.c_int()
        check(self.lib.ccu_solve_del2_u(self._ctx, d0.ctypes.data_as(C.c_void_p), p, C.c_double(acc), C.byref(valid), C.byref(cyc)))
        return d0, valid.value, cyc.value

    def solve_Ahat_p_fhat(self, V, P, F, imp, steps_max):
        """Returns (V, P, iterations, residual, hist[iterations,5])."""
        Vw = np.array(V, dtype=np.float64)
        Pw = np.array(P, dtype=np.float64)
        F, fp = _f64(F)
        steps = C.c_int(steps_max)
        res = C.c_float()
        hist = np.zeros((max(steps_max, 1), 5))
        check(self.lib.ccu_solve_Ahat_p_fhat(self._ctx, Vw.ctypes.data_as(C.c_void_p), Pw.ctypes.data_as(C.c_void_p), fp,
                                             C.c_double(imp), C.byref(steps), C.byref(res), hist.ctypes.data_as(C.c_void_p)))
        return Vw, Pw, steps.value, res.value, hist[:steps.value]

    def general_stokes_solver(self, T=None, buoyancy=None, *, rebuild=1, augmented_Lagr=1, augmented=1.0e3, precondition=1,
                              guess=0, U=None, P=None, want_host=True):
        """general_stokes_solver (Drive_solvers.c:45): returns (U, P, iterations, residual); U, P are None when
        want_host is False (results stay in HBM)."""
        t = None if T is None else np.ascontiguousarray(T, dtype=np.float32)
        b = None if buoyancy is None else np.ascontiguousarray(buoyancy, dtype=np.float32)
        lm = self.levmax
        if want_host:
            U = np.zeros(self.neq(lm)) if U is None else np.ascontiguousarray(U, dtype=np.float64)
            P = np.zeros(self.nel(lm)) if P is None else np.ascontiguousarray(P, dtype=np.float64)
        else:
            U = P = None
        it, res = C.c_int(), C.c_float()
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        check(self.lib.ccu_general_stokes_solver(self._ctx, ptr(t), ptr(b), int(rebuild), int(augmented_Lagr), C.c_double(augmented),
                                                 int(precondition), int(guess), ptr(U), ptr(P), C.byref(it), C.byref(res)))
        return U, P, it.value, res.value

    # -- energy step (PG_timestep and its parts, Advection_diffusion.c)
    def set_energy_params(self, fine_tune_dt, fixed_timestep, gamma, temp_iterations, diffusivity, expansivity, Q0=0.0):
        noz = self.dims[self.levmax][2]
        dz = np.ascontiguousarray(diffusivity, dtype=np.float32)
        ex = np.ascontiguousarray(expansivity, dtype=np.float32)
        assert dz.size == noz and ex.size == noz
        check(self.lib.ccu_set_energy_params(self._ctx, C.c_float(fine_tune_dt), C.c_float(fixed_timestep), C.c_float(gamma),
                                             int(temp_iterations), dz.ctypes.data_as(C.c_void_p), ex.ctypes.data_as(C.c_void_p), C.c_float(Q0)))

    def set_heating_params(self, adi_heating, visc_heating, disptn_number, surf_temp, Atemp):
        check(self.lib.ccu_set_heating_params(self._ctx, int(adi_heating), int(visc_heating), C.c_float(disptn_number),
                                              C.c_float(surf_temp), C.c_float(Atemp)))
        self._heating = bool(adi_heating or visc_heating)

    def set_phase_params(self, zlm, z410, Ra_670, clapeyron670, width670, Ra_410, clapeyron410, width410):
        """Phase-change parameters as the reference holds them after its first phase_change call (Phase_change.c:51-67)."""
        f = C.c_float
        check(self.lib.ccu_set_phase_params(self._ctx, f(zlm), f(z410), f(Ra_670), f(clapeyron670), f(width670), f(Ra_410),
                                            f(clapeyron410), f(width410)))
        self._heating = True

    def set_step(self, solution_cycles):
        self._step = int(solution_cycles)
        check(self.lib.ccu_set_step(self._ctx, self._step))

    def phase_change(self, update_transT=True):
        """phase_change (Phase_change.c:43): returns (Fas670, Fas410, (transT670, transT410))."""
        n = self.nno(self.levmax)
        a, b, t = np.empty(n, dtype=np.float32), np.empty(n, dtype=np.float32), np.empty(2, dtype=np.float32)
        check(self.lib.ccu_phase_change(self._ctx, int(update_transT), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                        t.ctypes.data_as(C.c_void_p)))
        return a, b, t

    def get_heating_latent(self):
        out = np.empty(self.nel(self.levmax), dtype=np.float32)
        check(self.lib.ccu_get_heating_latent(self._ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def process_heating(self, want_host=True):
        """process_heating (Advection_diffusion.c:813): returns (heating_adi, heating_visc) float32[nel] or None."""
        nel = self.nel(self.levmax)
        a = np.empty(nel, dtype=np.float32) if want_host else None
        v = np.empty(nel, dtype=np.float32) if want_host else None
        ptr = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)  # noqa: E731
        check(self.lib.ccu_process_heating(self._ctx, ptr(a), ptr(v)))
        return (a, v) if want_host else None

    def set_tdot(self, Tdot=None):
        t = None if Tdot is None else np.ascontiguousarray(Tdot, dtype=np.float32)
        check(self.lib.ccu_set_tdot(self._ctx, None if t is None else t.ctypes.data_as(C.c_void_p)))

    def set_velocity(self, V1, V2, V3):
        vs = [np.ascontiguousarray(v, dtype=np.float32) for v in (V1, V2, V3)]
        for v in vs:
            assert v.size == self.nno(self.levmax)
        check(self.lib.ccu_set_velocity(self._ctx, *[v.ctypes.data_as(C.c_void_p) for v in vs]))

    def v_from_vector(self, want_host=True):
        """v_from_vector (Stokes_flow_Incomp.c:530) on the resident U; returns (V1, V2, V3) float32 or None."""
        out = np.empty(3 * self.nno(self.levmax), dtype=np.float32) if want_host else None
        check(self.lib.ccu_v_from_vector(self._ctx, None if out is None else out.ctypes.data_as(C.c_void_p)))
        return None if out is None else tuple(out.reshape(3, -1))

    def std_timestep(self):
        dt = C.c_float()
        check(self.lib.ccu_std_timestep(self._ctx, C.byref(dt)))
        return np.float32(dt.value)

    def pg_solver(self):
        out = np.empty(self.nno(self.levmax), dtype=np.float32)
        check(self.lib.ccu_pg_solver(self._ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def PG_timestep(self, T=None, Tdot=None):
        """PG_timestep (Advection_diffusion.c:251): returns (T, Tdot, dt, T_interior); T, Tdot None = resident fields
        (then the returned fields are None as well)."""
        t = None if T is None else np.array(T, dtype=np.float32)
        td = None if Tdot is None else np.array(Tdot, dtype=np.float32)
        dt, ti = C.c_float(), C.c_float()
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        check(self.lib.ccu_PG_timestep(self._ctx, ptr(t), ptr(td), C.byref(dt), C.byref(ti)))
        return t, td, np.float32(dt.value), np.float32(ti.value)

    def thermal_buoyancy(self, Atemp, want_host=True):
        out = np.empty(self.nno(self.levmax), dtype=np.float32) if want_host else None
        check(self.lib.ccu_thermal_buoyancy(self._ctx, C.c_float(Atemp), None if out is None else out.ctypes.data_as(C.c_void_p)))
        return out

    def heat_flux(self):
        """heat_flux (Process_buoyancy.c:63): returns (Nut, Nub)."""
        a, b = C.c_float(), C.c_float()
        check(self.lib.ccu_heat_flux(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_temperature(self, want_tdot=False):
        T = np.empty(self.nno(self.levmax), dtype=np.float32)
        Td = np.empty_like(T) if want_tdot else None
        check(self.lib.ccu_get_temperature(self._ctx, T.ctypes.data_as(C.c_void_p), None if Td is None else Td.ctypes.data_as(C.c_void_p)))
        return (T, Td) if want_tdot else T

    # -- markers of the compositional field (Composition_adv.c)
    def markers_setup(self, capacity, markers_per_ele, rnoz, XP1, XP2, XP3, RG3, XG1, XG2, Element, Acomp=0.0):
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        xs = [f64(XP1), f64(XP2), f64(XP3)]
        rg = np.ascontiguousarray(RG3, dtype=np.int32)
        g1, g2 = f64(XG1), f64(XG2)
        el = np.ascontiguousarray(Element, dtype=np.uint32)
        assert rg.size == rnoz + 1 and el.size == self.nel(self.levmax)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        check(self.lib.ccu_markers_setup(self._ctx, int(capacity), int(markers_per_ele), int(rnoz), p(xs[0]), p(xs[1]), p(xs[2]), p(rg),
                                         p(g1), p(g2), p(el), C.c_float(Acomp)))

    def markers_upload(self, XMC1, XMC2, XMC3, C12, CElement, CE=None):
        xs = [np.ascontiguousarray(a, dtype=np.float64) for a in (XMC1, XMC2, XMC3)]
        c12 = np.ascontiguousarray(C12, dtype=np.int32)
        ce = np.ascontiguousarray(CElement, dtype=np.int32)
        cef = None if CE is None else np.ascontiguousarray(CE, dtype=np.float32)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        self._nmarkers = xs[0].size
        check(self.lib.ccu_markers_upload(self._ctx, int(xs[0].size), p(xs[0]), p(xs[1]), p(xs[2]), p(c12), p(ce), p(cef)))

    def markers_download(self):
        """dict with XMC, XMCpred (float64 [3, n]), VO, Vpred (float32 [3, n]), CElement, C (nodal), CE (elemental)."""
        n, lm = self.markers_count(), self.levmax
        out = dict(XMC=np.empty((3, n)), XMCpred=np.empty((3, n)), VO=np.empty((3, n), np.float32), Vpred=np.empty((3, n), np.float32),
                   CElement=np.empty(n, np.int32), C=np.empty(self.nno(lm), np.float32), CE=np.empty(self.nel(lm), np.float32))
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        check(self.lib.ccu_markers_download(self._ctx, p(out["XMC"]), p(out["XMCpred"]), p(out["VO"]), p(out["Vpred"]), p(out["CElement"]),
                                            p(out["C"]), p(out["CE"])))
        return out

    def markers_download_C(self, out):
        """nodal composition E->C only, into a caller-owned (pinned) float32[nno] array."""
        assert out.dtype == np.float32 and out.size == self.nno(self.levmax)
        check(self.lib.ccu_markers_download(self._ctx, None, None, None, None, None, out.ctypes.data_as(C.c_void_p), None))

    def Euler(self, timestep):
        check(self.lib.ccu_Euler(self._ctx, C.c_float(timestep)))

    def Runge_Kutta(self, timestep):
        check(self.lib.ccu_Runge_Kutta(self._ctx, C.c_float(timestep)))

    # -- one pass of main()'s time loop (Citcom.c:111-161), everything resident on the device
    # -- markers changing subdomain, host hand-over variant (two subdomains in one process; see citcomcu_b200.h)
    def markers_set_decomp(self, nproc, me):
        a, b = (C.c_int * 3)(*nproc), (C.c_int * 3)(*me)
        check(self.lib.ccu_markers_set_decomp(self._ctx, a, b))

    def markers_step_export(self, timestep, corrector, max_records=1 << 20):
        cnt = (C.c_int * 27)()
        rec = np.empty(8 * max_records, dtype=np.float64)
        check(self.lib.ccu_markers_step_export(self._ctx, C.c_float(timestep), int(corrector), cnt, rec.ctypes.data_as(C.c_void_p), int(max_records)))
        cnt = np.array(cnt[:], dtype=np.int64)
        return cnt, rec[:8 * int(cnt.sum())].reshape(-1, 8).copy()

    def markers_import_finish(self, corrector, records):
        r = np.ascontiguousarray(records, dtype=np.float64).reshape(-1, 8)
        check(self.lib.ccu_markers_import_finish(self._ctx, int(corrector), int(r.shape[0]), r.ctypes.data_as(C.c_void_p)))

    def markers_count(self):
        return int(self.lib.ccu_markers_count(self._ctx))

    def output_stage(self):
        check(self.lib.ccu_output_stage(self._ctx))

    def output_write(self, prefix, me, file_number, timesteps, elapsed_time, composition=0):
        check(self.lib.ccu_output_write(self._ctx, str(prefix).encode(), int(me), int(file_number), int(timesteps), C.c_double(elapsed_time), int(composition)))

    def output_wait(self):
        check(self.lib.ccu_output_wait(self._ctx))

    def get_stress_topo(self):
        """get_stress + get_STD_topo (Topo_gravity.c:352,307): (S[6, nno] = SXX, SXY, SXZ, SYY, SZY, SZZ; tpg[nsf]; tpgb[nsf])."""
        lm = self.levmax
        nox, noy, _ = self.dims[lm]
        S = np.empty((6, self.nno(lm)), np.float32); tpg = np.empty(nox * noy, np.float32); tpgb = np.empty(nox * noy, np.float32)
        check(self.lib.ccu_get_stress_topo(self._ctx, S.ctypes.data_as(C.c_void_p), tpg.ctypes.data_as(C.c_void_p), tpgb.ctypes.data_as(C.c_void_p)))
        return S, tpg, tpgb

    def averages(self, composition=False):
        """averages (Process_velocity.c:179): layer vrms, layer mean viscosity (and composition): float[noz] arrays."""
        noz = self.dims[self.levmax][2]
        vr, vi = np.zeros(noz, np.float32), np.zeros(noz, np.float32)
        cc = np.zeros(noz, np.float32) if composition else None
        check(self.lib.ccu_averages(self._ctx, vr.ctypes.data_as(C.c_void_p), vi.ctypes.data_as(C.c_void_p),
                                    cc.ctypes.data_as(C.c_void_p) if composition else None))
        return (vr, vi, cc) if composition else (vr, vi)

    @staticmethod
    def volume_vrms(layer_vrms, z):
        """Volume-weighted Vrms of a Cartesian box from the layer values: sqrt of the trapezoid z-integral of vrms(z)^2 over the
        height (the reference keeps only the layers, SURVEY.md 8c: both arms of a parity check use this same definition)."""
        v2 = np.asarray(layer_vrms, dtype=np.float64) ** 2
        z = np.asarray(z, dtype=np.float64)
        return float(np.sqrt(np.trapezoid(v2, z) / (z[-1] - z[0])))

    def PG_timestep_particle(self, Atemp):
        """PG_timestep_particle (Advection_diffusion.c:128): alternates, like the reference's static `on_off`, between
        (0) std_timestep + thermal step + Euler marker predictor and (1) the Runge_Kutta marker corrector with the new
        velocity; both end with thermal_buoyancy."""
        on_off = getattr(self, "_on_off", 0)
        if on_off == 0:
            _, _, self._dt, _ = self.PG_timestep()
            self.Euler(self._dt)
        else:
            self.Runge_Kutta(self._dt)
        self.thermal_buoyancy(Atemp, want_host=False)
        self._on_off = 1 - on_off
        return self._dt

    def advance(self, Atemp, *, composition=False, rebuild=1, **stokes_kw):
        """One timestep as main() runs it: next_buoyancy_field (PG_timestep or PG_timestep_particle), general_stokes_solver
        from the previous solution, v_from_vector, and with markers the second next_buoyancy_field call.  Returns dt."""
        if getattr(self, "_heating", False):
            self.process_heating(want_host=False)           # Citcom.c:116, before next_buoyancy_field
        self.set_step(getattr(self, "_step", 0) + 1)        # E->monitor.solution_cycles++ (Citcom.c:120)
        if composition:
            dt = self.PG_timestep_particle(Atemp)
        else:
            _, _, dt, _ = self.PG_timestep()
            self.thermal_buoyancy(Atemp, want_host=False)
        _, _, its, _ = self.general_stokes_solver(None, None, rebuild=rebuild, guess=2, want_host=False, **stokes_kw)
        self.v_from_vector(want_host=False)
        if composition:
            self.PG_timestep_particle(Atemp)
        return dt, its

    # -- CUDA-event profile of the finest-level kernels
    PROF = dict(relax_fine=0, matvec_fine=1, build=2, coarse=3, transfer_fine=4, faces_fine=5 + _lib.MAX_LEVELS)

    def profile_enable(self, on=True):
        check(self.lib.ccu_profile_enable(self._ctx, int(on)))

    def profile_reset(self):
        check(self.lib.ccu_profile_reset(self._ctx))

    def profile_read(self, cls):
        """cls: a class name of PROF or an int multigrid level (per-level totals; recorded with graphs off)."""
        ms, n = C.c_double(), C.c_longlong()
        cid = 5 + cls if isinstance(cls, int) else self.PROF[cls]
        check(self.lib.ccu_profile_read(self._ctx, cid, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- device-resident forms
    def vec_upload(self, lev, vec, host):
        h, p = _f64(host)
        check(self.lib.ccu_vec_upload(self._ctx, lev, VEC[vec], p))

    def vec_download(self, lev, vec):
        out = np.empty(self.neq(lev))
        check(self.lib.ccu_vec_download(self._ctx, lev, VEC[vec], out.ctypes.data_as(C.c_void_p)))
        return out

    def pvec_upload(self, host):
        h, p = _f64(host)
        check(self.lib.ccu_pvec_upload(self._ctx, p))

    def pvec_download(self):
        out = np.empty(self.nel(self.levmax))
        check(self.lib.ccu_pvec_download(self._ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def dev_matvec(self, lev, u, Au, strip_bcs=1):
        check(self.lib.ccu_dev_matvec(self._ctx, lev, VEC[u], VEC[Au], int(strip_bcs)))

    def dev_gauss_seidel(self, lev, d0, F, Ad, cycles, guess):
        check(self.lib.ccu_dev_gauss_seidel(self._ctx, lev, VEC[d0], VEC[F], VEC[Ad], int(cycles), int(guess)))

    def dev_relax_sweeps(self, lev, d0, F, cycles):
        check(self.lib.ccu_dev_relax_sweeps(self._ctx, lev, VEC[d0], VEC[F], int(cycles)))

    def dev_multi_grid(self, d1, F):
        res = C.c_double()
        check(self.lib.ccu_dev_multi_grid(self._ctx, VEC[d1], VEC[F], C.byref(res)))
        return res.value

    def dev_solve_Ahat_p_fhat(self, imp, steps_max):
        steps = C.c_int(steps_max)
        res = C.c_float()
        hist = np.zeros((max(steps_max, 1), 5))
        check(self.lib.ccu_dev_solve_Ahat_p_fhat(self._ctx, C.c_double(imp), C.byref(steps), C.byref(res), hist.ctypes.data_as(C.c_void_p)))
        return steps.value, res.value, hist[:steps.value]


def context_from_dump(dump, **overrides) -> StokesContext:
    """Build a context from the control block of an oracle dump (tests / smoke only pass the
    dump object in; this module does not import the oracle)."""
    ctl = dump.control()
    nox, noy, noz = {}, {}, {}
    for lev in range(ctl["levmin"], ctl["levmax"] + 1):
        d = dump.dims(lev)
        nox[lev], noy[lev], noz[lev] = d["nox"], d["noy"], d["noz"]
    kw = dict(v_steps_low=ctl["v_steps_low"], v_steps_high=ctl["v_steps_high"], down_heavy=ctl["down_heavy"],
              up_heavy=ctl["up_heavy"], mg_cycle=ctl["mg_cycle"], p_iterations=ctl["p_iterations"], accuracy=ctl["accuracy"])
    kw.update(overrides)
    ctx = StokesContext(ctl["levmin"], ctl["levmax"], nox, noy, noz, **kw)
    ctx.load_operator_from(dump)
    return ctx


AGG_NODES = 30000      # multigrid levels whose GLOBAL mesh has at most this many nodes are solved replicated on every rank


def context_from_problem(prob, device=0, unique_id=None, agglomerate=True, communicator=True, **overrides) -> StokesContext:
    """Build a context for one subdomain of a `citcomcu_b200.problem.CartesianProblem`: mesh, flags and
    coordinates go up, every operator array is then constructed on the device (ccu_build_geometry here,
    viscosity / stiffness at the first general_stokes_solver call)."""
    kw = {k: prob.control[k] for k in ("v_steps_low", "v_steps_high", "down_heavy", "up_heavy", "mg_cycle", "p_iterations", "accuracy")}
    kw.update(overrides)
    ctx = StokesContext(prob.levmin, prob.levmax, prob.nox, prob.noy, prob.noz, device=device, **kw)
    if prob.nproc != (1, 1, 1) and communicator:      # communicator=False: a lone subdomain (tests hand data over on the host)
        ctx.comm_init(prob.nproc, prob.me_loc, unique_id)
    for lev in range(prob.levmin, prob.levmax + 1):
        ctx.set_node_flags(lev, prob.node_flags(lev))
        ctx.set_coordinates(lev, *prob.coordinates(lev))
        if hasattr(prob, "spherical_coordinates"):                 # regional-spherical block: E->SXX next to the Cartesian E->XX
            ctx.set_spherical_coordinates(lev, *prob.spherical_coordinates(lev))
    ctx.build_geometry()
    vb = prob.velocity_bcs()
    if any(np.any(v != 0) for v in vb) or getattr(prob, "force_velocity_bcs", False):
        ctx.set_velocity_bcs(*vb)
    if prob.nproc != (1, 1, 1) and agglomerate and communicator:
        gp = prob.global_problem()
        levs = [lev for lev in range(prob.levmin, prob.levmax) if gp.nno(lev) <= AGG_NODES]
        if levs:
            g = ctx.agglomerate(max(levs), {lev: gp.dims(lev) for lev in range(gp.levmin, gp.levmax + 1)})
            for lev in range(g.levmin, g.levmax + 1):
                g.set_node_flags(lev, gp.node_flags(lev))
                g.set_coordinates(lev, *gp.coordinates(lev))
                if hasattr(gp, "spherical_coordinates"):
                    g.set_spherical_coordinates(lev, *gp.spherical_coordinates(lev))
            g.build_geometry()
    v = prob.visc
    ctx.set_viscosity_law(v["tdepv"], v["rheol"], v["N0"], v["E"], v["T"], v["Z"], v["vmin"], v["min_value"], v["vmax"],
                          v["max_value"], v["smooth_cycles"])
    ctx.set_material(prob.material())
    if v.get("sdepv"):
        ctx.set_sdepv(1, v["sdepv_rheology"], v["sdepv_expt"], v["sdepv_trns"], v["sdepv_misfit"], v["sdepv_iter_damp"], v["sdepv_max_iter"],
                      v["sdepv_start_from_newtonian"], v["sdepv_trns_T"], v["sdepv_trns_c"])
    return ctx
