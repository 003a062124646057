"""One rank of a multi-GPU parity run (spawned by tests/test_gpu_multi.py, one process per GPU).

Every rank builds its subdomain of the same input file, joins the library's NCCL communicator, and returns
its part of: a matvec on a seeded global vector, a smoother call, and the converged Stokes solve."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def seeded_global_vector(gp, lev, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, 3 * gp.nno(lev))


def run_rank(rank, world, text, uid_q, out_q, accuracy, agglomerate=True, device_of_rank=None):
    try:
        from citcomcu_b200 import decomp
        from citcomcu_b200.problem import CartesianProblem
        from citcomcu_b200.stokes import StokesContext, context_from_problem
        nproc = CartesianProblem(text).nproc
        me = decomp.me_loc_of(rank, nproc)
        prob = CartesianProblem(text, me_loc=me)
        gp = prob.global_problem()
        if rank == 0:
            uid = StokesContext.comm_unique_id()
            for _ in range(world - 1):
                uid_q.put(uid)
        else:
            uid = uid_q.get(timeout=120)
        dev = rank if device_of_rank is None else device_of_rank[rank]
        ctx = context_from_problem(prob, device=dev, unique_id=uid, accuracy=accuracy, agglomerate=agglomerate)
        lm = prob.levmax
        Tg = gp.initial_temperature()
        bg = gp.buoyancy(Tg)
        T, b = prob.local_slice(Tg), prob.local_slice(bg)
        ctl = prob.control
        res = {"rank": rank, "me": me}
        ctx.set_temperature(T)
        res["F"] = ctx.assemble_forces(b)
        ctx.get_system_viscosity()
        ctx.construct_stiffness_B_matrix(ctl["augmented_Lagr"], ctl["augmented"], ctl["precondition"])
        for lev in range(prob.levmin, prob.levmax + 1):
            ug = seeded_global_vector(gp, lev, 100 + lev)
            u = prob.local_slice(ug, lev, 3)
            res[f"Au{lev}"] = ctx.n_assemble_del2_u(u, lev, 1)
            res[f"dot{lev}"] = ctx.global_vdot(u, u, lev)
            if lev > prob.levmin:
                res[f"proj{lev}"] = ctx.project_vector(lev, u)
            if lev < prob.levmax:
                res[f"interp{lev}"] = ctx.interp_vector(lev, u)
        pg = np.random.default_rng(7).uniform(-1, 1, gp.nel(lm))
        res["gradp"] = ctx.assemble_grad_p(prob.local_slice_elements(pg), lm)
        res["MASS"] = ctx.get_level_array(lm, "MASS")
        res["BPI"] = ctx.get_level_array(lm, "BPI")
        fg = seeded_global_vector(gp, lm, 55)
        f = ctx.strip_bcs_from_residual(prob.local_slice(fg, lm, 3), lm)
        d0, Ad = ctx.gauss_seidel(f, 2, lm, 0)
        res["gs_d0"], res["gs_Ad"] = d0, Ad
        res["gs_KD"] = ctx.n_assemble_del2_u(d0, lm, 1)
        # overlapped sweeps (duplicated-node exchange on a second stream, colour passes split by distance from the faces) give the
        # serial order's result bit for bit: smoother calls on the two finest levels, then the whole solve
        same = True
        for lev in ((lm, lm - 1) if world == 2 else ()):        # checked on hardware for the three 2-subdomain splits
            fl = ctx.strip_bcs_from_residual(prob.local_slice(seeded_global_vector(gp, lev, 77 + lev), lev, 3), lev)
            out = {}
            for mode in (0, 2):
                ctx.set_option("halo_overlap", mode)
                out[mode] = ctx.gauss_seidel(fl, 3, lev, 0)
            same = same and np.array_equal(out[0][0], out[2][0]) and np.array_equal(out[0][1], out[2][1])
        res["overlap_sweeps_bitwise"] = bool(same)
        ctx.set_option("halo_overlap", 0)
        U0, P0, its0, _ = ctx.general_stokes_solver(T, b, rebuild=1, augmented_Lagr=ctl["augmented_Lagr"], augmented=ctl["augmented"],
                                                    precondition=ctl["precondition"], guess=0)
        if world == 2:
            ctx.set_option("halo_overlap", 2)
            U, P, its, r = ctx.general_stokes_solver(T, b, rebuild=1, augmented_Lagr=ctl["augmented_Lagr"], augmented=ctl["augmented"],
                                                     precondition=ctl["precondition"], guess=0)
            ctx.set_option("halo_overlap", 0)
        else:
            U, P, its = U0, P0, its0
        res["overlap_solve_bitwise"] = bool(np.array_equal(U, U0) and np.array_equal(P, P0) and its == its0)
        res["U"], res["P"], res["its"] = U, P, its
        # two coupled timesteps on the subdomains: energy step, buoyancy with cross-rank layer averages, Stokes solve
        noz = prob.dims(lm)[2]
        ctx.set_energy_params(0.75, 0.0, 0.5, 2, np.ones(noz, np.float32), np.ones(noz, np.float32), 0.0)
        ctx.set_tdot(None)
        ctx.v_from_vector(want_host=False)
        kw = dict(augmented_Lagr=ctl["augmented_Lagr"], augmented=ctl["augmented"], precondition=ctl["precondition"])
        res["dt"] = []
        for _ in range(2):
            dt, _its = ctx.advance(float(prob.rayleigh), rebuild=1, **kw)
            res["dt"].append(float(dt))
        res["T2"] = ctx.get_temperature()
        res["b2"] = ctx.thermal_buoyancy(float(prob.rayleigh))
        # phase changes across the subdomains: transition temperatures from the layer averages (summed over the ranks of a plane)
        # of whichever z subdomain holds the phase depth (sum_across_depth), then the phase functions of the local nodes
        ctx.set_phase_params(0.7665505, 0.857143, -50.0, -0.05, 30.0, 80.0, 0.07, 30.0)
        F6, F4, tT = ctx.phase_change(update_transT=True)
        res["Fas670"], res["Fas410"], res["transT"] = F6, F4, tT
        res["launches"] = ctx.launch_count
        ctx.close()
        out_q.put(res)
    except Exception as e:  # surface the failure in the parent instead of hanging it
        import traceback
        out_q.put({"rank": rank, "error": f"{e}\n{traceback.format_exc()}"})


def run_rank_markers(rank, world, text, uid_q, out_q, setup):
    """Marker Euler / Runge_Kutta across subdomains over NCCL: `setup` = this rank's arrays of the reference dump."""
    try:
        from citcomcu_b200 import decomp
        from citcomcu_b200.problem import CartesianProblem
        from citcomcu_b200.stokes import StokesContext, context_from_problem
        nproc = CartesianProblem(text).nproc
        me = decomp.me_loc_of(rank, nproc)
        prob = CartesianProblem(text, me_loc=me)
        if rank == 0:
            uid = StokesContext.comm_unique_id()
            for _ in range(world - 1):
                uid_q.put(uid)
        else:
            uid = uid_q.get(timeout=120)
        ctx = context_from_problem(prob, device=rank, unique_id=uid, agglomerate=False)
        d = setup
        ip, dp = d["mk_ints"], d["mk_doubles"]
        ctx.markers_setup(int(ip[3]), int(ip[1]), int(ip[0]), d["mk_XP1"], d["mk_XP2"], d["mk_XP3"], d["mk_RG3"], dp[0:3], dp[3:6],
                          d["mk_Element"], Acomp=float(dp[7]))
        ctx.markers_upload(d["mk_in_XMC1"], d["mk_in_XMC2"], d["mk_in_XMC3"], d["mk_in_C12"], d["mk_in_CElement"], d["mk_in_CE"])
        ctx.set_velocity(d["mk_in_V1"], d["mk_in_V2"], d["mk_in_V3"])
        dt = np.float32(dp[6])
        res = {"rank": rank}
        ctx.Euler(dt)
        res["euler"] = ctx.markers_download()
        ctx.Runge_Kutta(dt)
        res["rk"] = ctx.markers_download()
        ctx.close()
        out_q.put(res)
    except Exception as e:
        import traceback
        out_q.put({"rank": rank, "error": f"{e}\n{traceback.format_exc()}"})
