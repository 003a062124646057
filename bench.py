#!/usr/bin/env python
"""bench.py -- Stokes-solve seconds per timestep and MG smoother / matvec GB/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--mesh 256x256x128 --levels 6]

One "step" = one `general_stokes_solver` (Drive_solvers.c:45) of BASELINE config 3 (3-D Cartesian
256x256x128 elements, Ra=1e7, temperature-dependent viscosity contrast 1e5, free slip): viscosity from
T, stiffness / BI / BPI rebuild on all multigrid levels, body force, Uzawa pressure iteration with the
full-multigrid velocity solves, from a zero initial guess (the reference's step-0 solve; every step
does identical work).  Synthetic data: the reference's analytic initial temperature field.

  value   seconds per step with T and the buoyancy already resident in HBM (CUDA events on the
          library's stream, max over ranks)
  e2e     the same call through the C ABI with pinned HOST buffers: T and buoyancy go host->device,
          U and P come back device->host inside the timed region
  roofline  the finest-level Gauss-Seidel smoother kernel (the kernel with the largest share of the step):
          algorithmic bytes 600 B/node/sweep (SURVEY.md 8d without Ad, which the smoother never touches:
          K 504 + F 24 + BI 24 + d0 read/write 48) / launches per sweep, over its mean launch duration
          measured with CUDA events in a separate profiled pass of the same steps (`value` itself is timed
          with the per-kernel events off); `traffic` = DRAM bytes per launch from the committed
          ncu --set full capture (profiles/traffic.json)
  parity  every run also solves a 64x64x32 sample of the same configuration to accuracy 1e-8 on the same
          N GPUs and compares U, P with the unmodified reference's solution of that sample (relative L2)
  cpu_baseline / --impl reference  the UNMODIFIED reference (oracle/_ref, built from /root/reference
          by oracle/Makefile over oracle/mpi_shim, a process-based shared-memory MPI) running the SAME
          step on the SAME mesh on the box's host cores; repeats are capped (one warm-up + two timed
          steps, about 30 s each on 16 cores) and the line reports the counts it actually ran.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

RELAX_BYTES_PER_NODE_SWEEP = 600.0     # SURVEY.md 8(d) less Ad (never touched by the sweep): K 504 + F 24 + BI 24 + d0 r/w 48
PARITY_MESH, PARITY_LEVELS, PARITY_ACC = (64, 64, 32), 4, 1e-8
REF_MAX_WARMUP, REF_MAX_STEPS = 1, 2   # cap of the CPU reference arm: a full-size step takes ~30 s on 16 host cores
MATVEC_BYTES_PER_NODE = 552.0          # K 504 + u 24 + Au 24


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh", default="256x256x128")
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--ref-mesh", default=None, help="mesh for the CPU reference (default: the benchmark mesh itself)")
    ap.add_argument("--ref-levels", type=int, default=None)
    ap.add_argument("--no-parity", action="store_true", help="skip the 64x64x32 parity sample")
    ap.add_argument("--config", type=int, default=3, choices=[3, 5],
                    help="BASELINE config: 3 = Stokes solve at 256x256x128 (the headline), 5 = thermochemical marker advection, "
                         "128x128x64 elements x 20 markers per element")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="diagnostic: launch coarse levels kernel by kernel and report per-level times")
    ap.add_argument("--opt", action="append", default=[], help="diagnostic: library option name=value (ccu_set_option)")
    return ap.parse_args()


def measured_peak_gbs():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.proc = None
        self.path = Path(tempfile.mkdtemp(prefix="ccu_clk_")) / "clocks.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.path.read_text().split("\n"):
            tok = [t.strip() for t in line.split(",")]
            if len(tok) < 9:
                continue
            try:
                sm.append(float(tok[1])); smax.append(float(tok[2]))
            except ValueError:
                continue
            for nm, v in zip(names, tok[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


def mesh_tuple(s):
    a = [int(v) for v in s.lower().split("x")]
    assert len(a) == 3
    return tuple(a)


def nproc_for_cores(cores, mg):
    """Largest processor grid (x, y, z) with <= cores ranks that divides mgunit (README:149-151)."""
    best = (1, 1, 1)
    for px in (1, 2, 4, 8):
        for py in (1, 2, 4, 8):
            for pz in (1, 2, 4):
                if mg[0] % px or mg[1] % py or mg[2] % pz:
                    continue
                n, nb = px * py * pz, best[0] * best[1] * best[2]
                if n <= cores and (n > nb or (n == nb and max(px, py, pz) < max(best))):
                    best = (px, py, pz)
    return best


def reference_times(ref_mesh, ref_levels, repeats, cores, accuracy=None, dump_up=False):
    """Run the unmodified reference's general_stokes_solver (zero guess) `repeats` times on `ref_mesh`."""
    from oracle import pyoracle as po
    from citcomcu_b200 import inputfile
    if not po.have_ref():
        raise RuntimeError("oracle/_ref is not built (python -c 'import __graft_entry__ as g; g.build()' where /root/reference exists)")
    f = 2 ** (ref_levels - 1)
    mg = (ref_mesh[0] // f, ref_mesh[1] // f, ref_mesh[2] // f)
    nproc = nproc_for_cores(cores, mg)
    nranks = nproc[0] * nproc[1] * nproc[2]
    kw = {} if accuracy is None else {"accuracy": accuracy}
    txt = inputfile.tdepv_box(*ref_mesh, ref_levels, nproc=nproc, maxstep=1, **kw)
    wd = Path(tempfile.mkdtemp(prefix="ccu_refbench_"))
    times = po.run_timezero(txt, wd, repeats, nproc=nranks, dump_up=dump_up)
    if dump_up:
        return times, nproc, nranks, [po.Dump(wd / "dump", rank=k) for k in range(nranks)]
    return times, nproc, nranks


def assemble_global(parts, gmesh):
    """Global U (3 per node) and P (per element) from per-subdomain pieces [(nproc, me_loc, (nox, noy, noz), U, P)]
    in the reference's numbering n = k + noz*(j + nox*i), e = ez + elz*(ex + elx*ey) (duplicated face nodes agree)."""
    from citcomcu_b200 import decomp
    gx, gy, gz = gmesh[0] + 1, gmesh[1] + 1, gmesh[2] + 1
    U = np.zeros((gx * gy * gz, 3)); P = np.zeros(gmesh[0] * gmesh[1] * gmesh[2])
    for nproc, me, (nox, noy, noz), u, pp in parts:
        U[decomp.global_node_ids(nproc, me, nox, noy, noz)] = np.asarray(u).reshape(-1, 3)
        ex, ey, ez = nox - 1, noy - 1, noz - 1
        i = np.arange(ey)[:, None, None] + me[1] * ey
        j = np.arange(ex)[None, :, None] + me[0] * ex
        k = np.arange(ez)[None, None, :] + me[2] * ez
        P[(k + gmesh[2] * (j + gmesh[0] * i)).reshape(-1)] = np.asarray(pp)
    return U.reshape(-1), P


def reference_parity_solution(cores):
    """The unmodified reference's converged U, P of the parity sample (global arrays) and its seconds."""
    times, nproc, nranks, dumps = reference_times(PARITY_MESH, PARITY_LEVELS, 1, cores, accuracy=PARITY_ACC, dump_up=True)
    parts = []
    for d in dumps:
        m = [int(v) for v in d["tz_meta"]]
        parts.append((tuple(m[0:3]), tuple(m[3:6]), tuple(m[6:9]), d["tz_U"], d["tz_P"]))
    U, P = assemble_global(parts, PARITY_MESH)
    from oracle import pyoracle as po
    ref_its = po.last_pressure_loops[-1] if po.last_pressure_loops else None
    return U, P, float(times[-1]), nranks, ref_its


def run_reference(args):
    mesh = mesh_tuple(args.mesh)
    ref_mesh = mesh_tuple(args.ref_mesh) if args.ref_mesh else mesh
    ref_levels = args.ref_levels or args.levels
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    warm, steps = min(args.warmup, REF_MAX_WARMUP), max(1, min(args.steps, REF_MAX_STEPS))
    times, nproc, nranks = reference_times(ref_mesh, ref_levels, warm + steps, cores)
    t = times[warm:]
    scale = (mesh[0] * mesh[1] * mesh[2]) / (ref_mesh[0] * ref_mesh[1] * ref_mesh[2])
    v = float(np.mean(t)) * scale
    sample = (f"unmodified reference general_stokes_solver on the {ref_mesh[0]}x{ref_mesh[1]}x{ref_mesh[2]} mesh, {nranks} ranks "
              f"({nproc[0]}x{nproc[1]}x{nproc[2]}) on {cores} host cores over oracle/mpi_shim (process-based shared-memory MPI: spin + "
              f"sched_yield mailboxes, collectives through point-to-point); {warm} warm-up + {steps} timed steps "
              f"(capped from --warmup {args.warmup} --steps {args.steps}); {float(np.mean(t)):.3f} s per step measured"
              + ("" if scale == 1 else f", scaled x{scale:g} by element count to {args.mesh} (an ESTIMATE)"))
    cfg = workload_config(args, mesh)
    cfg["reference_ran"] = {"mesh": list(ref_mesh), "levels": ref_levels, "nproc": list(nproc), "host_cores": cores,
                            "steps_timed": steps, "warmup_run": warm, "extrapolated": scale != 1}
    line = {"impl": "reference", "metric": "stokes_solve_s_per_timestep", "value": v, "unit": "s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": "s", "cores": nranks, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(args, mesh, nproc=(1, 1, 1)):
    return {"workload": f"BASELINE config 3: 3-D Cartesian {mesh[0]}x{mesh[1]}x{mesh[2]} elements, Ra=1e7, TDEPV contrast 1e5, "
                        f"free slip, {args.levels} multigrid levels; step = general_stokes_solver from a zero guess "
                        "(viscosity + stiffness rebuild + forces + Uzawa/FMG solve to accuracy 1e-3)",
            "mesh": list(mesh), "levels": args.levels, "nproc": list(nproc),
            "partition": "one subdomain per GPU, the reference's nprocx x nprocy x nprocz block decomposition; halo sums as grouped NCCL "
                         "send/recv (or through peer memory, option p2p_halo), allreduce over NCCL",
            "l2": "inputs larger than L2 (finest-level stiffness alone is > 4 GB)"}


def parity_gate(args, world, rank, local, nproc):
    """Solve the parity sample on the same processor grid as the benchmark and compare with the unmodified reference."""
    import torch.distributed as dist
    from citcomcu_b200 import decomp, inputfile
    from citcomcu_b200.problem import CartesianProblem
    from citcomcu_b200.stokes import StokesContext, context_from_problem
    uid = None
    if world > 1:
        box = [StokesContext.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    text = inputfile.tdepv_box(*PARITY_MESH, PARITY_LEVELS, nproc=nproc, maxstep=1, accuracy=PARITY_ACC)
    me = decomp.me_loc_of(rank, nproc)
    prob = CartesianProblem(text, me_loc=me)
    ctx = context_from_problem(prob, device=local, unique_id=uid, accuracy=PARITY_ACC)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    gp = prob.global_problem()
    Tg = gp.initial_temperature()
    ctl = prob.control
    U, P, its, _ = ctx.general_stokes_solver(prob.local_slice(Tg), prob.local_slice(gp.buoyancy(Tg)), rebuild=1, augmented_Lagr=ctl["augmented_Lagr"],
                                             augmented=ctl["augmented"], precondition=ctl["precondition"], guess=0)
    lm = prob.levmax
    col = bool(ctx.get_option("relax_col", lm))
    ctx.close()
    piece = (tuple(nproc), tuple(me), tuple(prob.dims(lm)), U, P)
    if world > 1:
        pieces = [None] * world
        dist.all_gather_object(pieces, piece)
    else:
        pieces = [piece]
    if rank != 0:
        return None
    Ug, Pg = assemble_global(pieces, PARITY_MESH)
    Ur, Pr, ref_s, ref_ranks, ref_its = reference_parity_solution(os.cpu_count() or 1)
    return {"sample": f"{PARITY_MESH[0]}x{PARITY_MESH[1]}x{PARITY_MESH[2]} elements, {PARITY_LEVELS} levels, same configuration, accuracy {PARITY_ACC:g}; "
                      f"GPU nproc {nproc[0]}x{nproc[1]}x{nproc[2]} vs the unmodified reference on {ref_ranks} ranks",
            "u_rel_l2": float(np.linalg.norm(Ug - Ur) / np.linalg.norm(Ur)), "p_rel_l2": float(np.linalg.norm(Pg - Pr) / np.linalg.norm(Pr)),
            "uzawa_its": int(its), "ref_its": ref_its, "tolerance": 1e-6, "column_smoother_on_sample_finest_level": col, "reference_s": ref_s}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from citcomcu_b200 import inputfile
    from citcomcu_b200.problem import CartesianProblem
    from citcomcu_b200.stokes import context_from_problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; citcomcu_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    mesh = mesh_tuple(args.mesh)
    t_setup = time.time()
    from citcomcu_b200 import decomp
    from citcomcu_b200.stokes import StokesContext
    f = 2 ** (args.levels - 1)
    nproc = decomp.nproc_for(world, (mesh[0] // f, mesh[1] // f, mesh[2] // f))
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [StokesContext.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    text = inputfile.tdepv_box(*mesh, args.levels, nproc=nproc, maxstep=1)
    prob = CartesianProblem(text, me_loc=decomp.me_loc_of(rank, nproc))
    ctx = context_from_problem(prob, device=local, unique_id=uid)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    if args.no_graphs:
        ctx.set_option("graphs", 0)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    lm = prob.levmax
    nno, neq, npno = prob.nno(lm), 3 * prob.nno(lm), prob.nel(lm)
    ctl = prob.control

    def pinned(a):
        t = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype)), pin_memory=True)
        n = t.numpy()
        n[...] = a
        return t, n

    gp = prob.global_problem()                      # T and the layer-averaged buoyancy are global fields
    Tg = gp.initial_temperature()
    T_t, T_h = pinned(prob.local_slice(Tg))
    b_t, b_h = pinned(prob.local_slice(gp.buoyancy(Tg)))
    del Tg
    U_t, U_h = pinned(np.zeros(neq))
    P_t, P_h = pinned(np.zeros(npno))
    kw = dict(rebuild=1, augmented_Lagr=ctl["augmented_Lagr"], augmented=ctl["augmented"], precondition=ctl["precondition"], guess=0)
    # make T and the buoyancy resident
    ctx.set_temperature(T_h)
    ctx.assemble_forces(b_h, want_host=False)
    setup_s = time.time() - t_setup

    def step_resident():
        return ctx.general_stokes_solver(None, None, want_host=False, **kw)

    def step_e2e():
        return ctx.general_stokes_solver(T_h, b_h, U=U_h, P=P_h, want_host=True, **kw)

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            e0.record(stream)
            its = [fn()[2] for _ in range(k)]
            e1.record(stream)
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3
        if world > 1:
            dist.barrier()
            t = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)            # max over ranks
            sec = float(t.item())
        return sec, its

    for _ in range(args.warmup):
        step_resident()
    ctx.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    # `value`: the K steps with the per-kernel CUDA-event profiling OFF
    ctx.profile_enable(False)
    l0 = ctx.launch_count
    total_s, its = timed(step_resident, args.steps)
    launches = ctx.launch_count - l0
    # e2e: host buffers through the C ABI
    step_e2e()
    e2e_s, _ = timed(step_e2e, args.steps)
    # class timings (smoother / matvec launch durations, rebuild, coarse levels): a separate pass of the same K steps with the events on
    ctx.profile_enable(True)
    ctx.profile_reset()
    prof_s, _ = timed(step_resident, args.steps)
    relax_ms, relax_n = ctx.profile_read("relax_fine")
    mv_ms, mv_n = ctx.profile_read("matvec_fine")
    build_ms, _ = ctx.profile_read("build")
    coarse_ms, coarse_n = ctx.profile_read("coarse")
    transfer_ms, _ = ctx.profile_read("transfer_fine")
    faces_ms, faces_n = ctx.profile_read("faces_fine")
    level_ms = {lev: ctx.profile_read(lev) for lev in range(prob.levmin, prob.levmax + 1)} if args.no_graphs else None
    ctx.profile_enable(False)
    clk = clocks.stop()
    s_per_step = total_s / args.steps
    peak, peak_src = measured_peak_gbs()
    relax_col, matvec_col, col_wf = ctx.get_option("relax_col", lm), ctx.get_option("matvec_col", lm), ctx.get_option("col_wf", lm)
    launches_per_sweep = (1 if col_wf else 4) if relax_col else 8
    if relax_col:
        relax_kernel = ("ccu_k_col<.., MODE 0> (finest-level column-resident Gauss-Seidel smoother: stiffness streamed once per sweep through a "
                        "bulk-copy ring in shared memory; " + ("one launch per sweep" if col_wf else "one launch per column colour") + ")")
        tkey = "ccu_k_col_relax"
    elif ctx.get_option("relax_full", lm):
        relax_kernel = ("ccu_k_relax_full<2> (one colour pass of the finest-level 8-colour Gauss-Seidel smoother, rows streamed from "
                        "the full-row copy of the stiffness)")
        tkey = "ccu_k_relax_full"
    else:
        relax_kernel = "ccu_k_relax_tab<2> (one colour pass of the finest-level 8-colour Gauss-Seidel smoother)"
        tkey = "ccu_k_relax_tab"
    relax_bytes_per_launch = RELAX_BYTES_PER_NODE_SWEEP * nno / launches_per_sweep
    relax_gbs = relax_bytes_per_launch * relax_n / (relax_ms * 1e-3) / 1e9 if relax_ms > 0 else 0.0
    mv_gbs = MATVEC_BYTES_PER_NODE * nno * mv_n / (mv_ms * 1e-3) / 1e9 if mv_ms > 0 else 0.0
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            t = json.loads(tf.read_text()).get(tkey, {}).get(args.mesh) if world == 1 else None
            traffic = None if t is None else t["dram_bytes_read_per_launch"] + t["dram_bytes_write_per_launch"]
        except Exception:
            traffic = None
    prof_total_ms = prof_s * 1e3
    line = {"metric": "stokes_solve_s_per_timestep", "value": s_per_step, "unit": "s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, mesh, nproc),
            "roofline": {"bound": "hbm", "kernel": relax_kernel,
                         "achieved": relax_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": relax_gbs / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": relax_bytes_per_launch,
                         "algorithmic_bytes_per_node_sweep": RELAX_BYTES_PER_NODE_SWEEP, "launches_per_sweep": launches_per_sweep,
                         "launches": relax_n, "avg_launch_ms": relax_ms / max(relax_n, 1),
                         "share_of_step": relax_ms / prof_total_ms},
            "smoother_gbs": relax_gbs, "matvec_gbs": mv_gbs,
            "matvec": {"kernel": "ccu_k_matvec_full" if ctx.get_option("matvec_full", lm) else ("ccu_k_col<.., MODE 1/2>" if matvec_col else "ccu_k_matvec_tab"), "achieved": mv_gbs, "frac": mv_gbs / peak,
                       "algorithmic_bytes_per_launch": MATVEC_BYTES_PER_NODE * nno, "launches": mv_n, "avg_launch_ms": mv_ms / max(mv_n, 1),
                       "share_of_step": mv_ms / prof_total_ms},
            "operator_rebuild_ms_per_step": build_ms / args.steps,
            "step_breakdown_ms": {"relax_fine": relax_ms / args.steps, "matvec_fine": mv_ms / args.steps, "build": build_ms / args.steps,
                                  "coarse_levels": coarse_ms / args.steps, "transfer_fine": transfer_ms / args.steps,
                                  "faces_fine_within_relax": faces_ms / args.steps,
                                  "other_fine_vector_ops_and_sync": (prof_total_ms - relax_ms - mv_ms - build_ms - coarse_ms - transfer_ms) / args.steps,
                                  "profiled_pass_ms_per_step": prof_total_ms / args.steps},
            "halo_exchange": ("peer-memory" if (world > 1 and ctx.get_option("p2p_halo", lm)) else ("nccl send/recv" if world > 1 else "none")),
            "uzawa_iterations": its, "gpu_launches": launches * world, "clocks": clk, "setup_s": setup_s,
            "e2e": {"value": e2e_s / args.steps, "unit": "s", "h2d_bytes_per_step": int(T_h.nbytes + b_h.nbytes) * world,
                    "d2h_bytes_per_step": int(U_h.nbytes + P_h.nbytes) * world}}
    if level_ms:
        line["level_ms_per_step"] = {str(k): {"ms": v[0] / args.steps, "sweeps": v[1] / args.steps} for k, v in level_ms.items()}
    ctx.close()
    # ---- parity gate: the 64x64x32 sample of the same configuration on the same N GPUs against the unmodified reference
    if not args.no_parity:
        try:
            line["parity"] = parity_gate(args, world, rank, local, nproc)
        except Exception as e:
            line["parity"] = {"error": str(e)[:400]}
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        try:
            cores = os.cpu_count() or 1
            times, rnproc, nranks = reference_times(mesh, args.levels, 1, cores)
            line["cpu_baseline"] = {"value": float(times[-1]), "unit": "s", "cores": nranks, "kind": "reference",
                                    "sample": f"one general_stokes_solver step of the unmodified reference (oracle/_ref) on the same {args.mesh} mesh, "
                                              f"{nranks} ranks ({rnproc[0]}x{rnproc[1]}x{rnproc[2]}) on {cores} host cores over oracle/mpi_shim; no warm-up"}
        except Exception as e:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


MARKER_BYTES_PER_SUBSTEP = 170.0     # XMC / XMCpred / VO / Vpred / C12 / CElement read + written, per marker per substep


def marker_setup_arrays(prob, rng, markers_per_ele, comp_depth=0.605):
    """Synthetic marker state for one subdomain, laid out as the reference keeps it (Convection.c:650-676 seeds uniformly in
    the box; C12 by depth, :846; pre_interpolation's z lookup table, Nodal_mesh.c:284; SIDEE flags, Parallel_related.c:1005)."""
    lm = prob.levmax
    nox, noy, noz = prob.dims(lm)
    elx, ely, elz = nox - 1, noy - 1, noz - 1
    X1, X2, X3 = prob.coordinates(lm)
    XP1 = X1.reshape(noy, nox, noz)[0, :, 0].astype(np.float64)
    XP2 = X2.reshape(noy, nox, noz)[:, 0, 0].astype(np.float64)
    XP3 = X3.reshape(noy, nox, noz)[0, 0, :].astype(np.float64)
    rnoz = 50 * elz + 1
    XRG = XP3[0] + np.arange(rnoz) * (XP3[-1] - XP3[0]) / (rnoz - 1)
    XRG[-1], XRG[0] = XP3[-1], XP3[0]
    RG3 = np.zeros(rnoz + 1, dtype=np.int32)
    for j in range(1, rnoz):                     # E->RG[3][j]: first element e with XRG[j+1] <= XP[e+1] and XRG[j] >= XP[e]
        ok = np.nonzero((XRG[j] <= XP3[1:]) & (XRG[j - 1] >= XP3[:-1]))[0]
        RG3[j] = ok[0] + 1 if ok.size else 0
    gp = prob.global_problem()
    g1, g2, g3 = gp.coordinates(lm)
    XG1 = np.array([g1.min(), g2.min(), g3.min()], dtype=np.float64) + 1e-6     # the reference keeps the markers 1e-6 inside the box
    XG2 = np.array([g1.max(), g2.max(), g3.max()], dtype=np.float64) - 1e-6
    El = np.zeros((ely, elx, elz), dtype=np.uint32)
    for sl in ((0, slice(None), slice(None)), (-1, slice(None), slice(None)), (slice(None), 0, slice(None)), (slice(None), -1, slice(None)),
               (slice(None), slice(None), 0), (slice(None), slice(None), -1)):
        El[sl] |= np.uint32(0x800000)            # SIDEE
    n = markers_per_ele * elx * ely * elz
    x = rng.uniform(XP1[0], XP1[-1], n); y = rng.uniform(XP2[0], XP2[-1], n); z = rng.uniform(XP3[0], XP3[-1], n)
    ex = np.clip(np.searchsorted(XP1, x, side="right") - 1, 0, elx - 1)
    ey = np.clip(np.searchsorted(XP2, y, side="right") - 1, 0, ely - 1)
    ez = np.clip(np.searchsorted(XP3, z, side="right") - 1, 0, elz - 1)
    CElement = (ez + elz * (ex + elx * ey) + 1).astype(np.int32)
    C12 = (z <= 1.0 - comp_depth).astype(np.int32)
    return dict(n=n, rnoz=rnoz, XP=(XP1, XP2, XP3), RG3=RG3, XG1=XG1, XG2=XG2, Element=El.reshape(-1), X=(x, y, z), C12=C12, CElement=CElement,
                nodes=(X1, X2, X3))


def run_markers(args):
    """BASELINE config 5: thermochemical convection, composition=1, markers_per_ele=20, 128x128x64 elements.  One step = the two
    marker substeps of a timestep (Euler predictor + Runge_Kutta corrector, Composition_adv.c:108,61: velocity at the markers,
    position update, element lookup, markers changing subdomain over NCCL, nodal composition) in a steady cellular flow."""
    import torch
    import torch.distributed as dist
    from citcomcu_b200 import decomp, inputfile
    from citcomcu_b200.problem import CartesianProblem
    from citcomcu_b200.stokes import StokesContext, context_from_problem
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; citcomcu_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    mesh, levels, mpe = (128, 128, 64), 5, 20
    f = 2 ** (levels - 1)
    nproc = decomp.nproc_for(world, (mesh[0] // f, mesh[1] // f, mesh[2] // f))
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [StokesContext.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    text = inputfile.tdepv_box(*mesh, levels, nproc=nproc, maxstep=1, composition=1, rayleigh_comp=1e6, markers_per_ele=mpe, comp_depth=0.605)
    prob = CartesianProblem(text, me_loc=decomp.me_loc_of(rank, nproc))
    ctx = context_from_problem(prob, device=local, unique_id=uid, agglomerate=False)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    M = marker_setup_arrays(prob, np.random.default_rng(1234 + rank), mpe)
    cap = int(1.3 * M["n"]) + 1024
    ctx.markers_setup(cap, mpe, M["rnoz"], *M["XP"], M["RG3"], M["XG1"], M["XG2"], M["Element"], Acomp=1.0)
    ctx.markers_upload(*M["X"], M["C12"], M["CElement"])
    X1, X2, X3 = M["nodes"]
    Lx, Lz = float(M["XG2"][0] - M["XG1"][0]), float(M["XG2"][2] - M["XG1"][2])
    # a steady two-cell flow (divergence free, no flow through the walls), speed ~ 1
    V1 = (np.sin(np.pi * X1 / Lx * 2) * np.cos(np.pi * X3 / Lz)).astype(np.float32)
    V2 = np.zeros_like(V1)
    V3 = (-2.0 * Lz / Lx * np.cos(np.pi * X1 / Lx * 2) * np.sin(np.pi * X3 / Lz)).astype(np.float32)
    dx = Lx / mesh[0]
    dt = np.float32(0.5 * dx / max(1e-30, float(np.abs(V3).max()), float(np.abs(V1).max())))       # half an element per substep
    V_t = [torch.from_numpy(v).pin_memory() for v in (V1, V2, V3)]
    V_h = [t.numpy() for t in V_t]
    C_t = torch.empty(prob.nno(prob.levmax), dtype=torch.float32, pin_memory=True)
    ctx.set_velocity(*V_h)
    n_total = M["n"] * world

    def step_resident():
        ctx.Euler(dt); ctx.Runge_Kutta(dt)

    def step_e2e():
        ctx.set_velocity(*V_h)                     # E->V in
        ctx.Euler(dt); ctx.Runge_Kutta(dt)
        ctx.markers_download_C(C_t.numpy())        # nodal composition E->C out (what thermal_buoyancy and the output read)

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(k):
                fn()
            e1.record(stream)
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3
        if world > 1:
            dist.barrier()
            t = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec

    for _ in range(args.warmup):
        step_resident()
    ctx.synchronize()
    clocks = ClockSampler(local); clocks.start()
    l0 = ctx.launch_count
    sec = timed(step_resident, args.steps)
    launches = ctx.launch_count - l0
    step_e2e()
    e2e = timed(step_e2e, args.steps)
    clk = clocks.stop()
    # size-independent checks: no marker lost or duplicated, all inside the box, composition in [0, 1]
    out = ctx.markers_download()
    cnt = torch.tensor([ctx.markers_count()], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(cnt)
    inside = bool(all((out["XMC"][d] >= M["XG1"][d] - 1e-12).all() and (out["XMC"][d] <= M["XG2"][d] + 1e-12).all() for d in range(3)))
    ok = torch.tensor([1 if (inside and float(out["C"].min()) >= -1e-6 and float(out["C"].max()) <= 1 + 1e-6) else 0], device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    peak, peak_src = measured_peak_gbs()
    substeps = 2 * args.steps
    rate = n_total * substeps / sec
    gbs = MARKER_BYTES_PER_SUBSTEP * (n_total / world) * substeps / sec / 1e9          # per GPU
    line = {"metric": "marker_substeps_per_s", "value": rate, "unit": "markers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "BASELINE config 5: thermochemical convection, composition=1, markers_per_ele=20, 128x128x64 elements "
                                   f"({n_total} markers); step = Euler + Runge_Kutta marker substeps of one timestep in a steady two-cell flow",
                       "mesh": list(mesh), "levels": levels, "nproc": list(nproc), "markers": n_total,
                       "l2": "marker arrays (1.8 GB at one GPU) larger than L2"},
            "roofline": {"bound": "hbm", "kernel": "one marker substep (mk_velocity + mk_advance + clamp/split + element assignment + nodal composition)",
                         "achieved": gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                         "algorithmic_bytes_per_marker_substep": MARKER_BYTES_PER_SUBSTEP, "per": "GPU"},
            "checks": {"markers_conserved": int(cnt.item()) == n_total, "inside_box_and_C_in_0_1": bool(ok.item())},
            "gpu_launches": launches * world, "clocks": clk,
            "e2e": {"value": n_total * substeps / e2e, "unit": "markers/s", "h2d_bytes_per_step": int(sum(v.nbytes for v in V_h)) * world,
                    "d2h_bytes_per_step": int(C_t.numel() * 4) * world},
            "cpu_baseline": {"value": None, "unit": "markers/s", "cores": 0, "kind": "reference",
                             "sample": "not timed for this config (the headline and its CPU baseline are config 3)"}}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.config == 5:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "config 5 has no reference arm; the headline (config 3) has"}))
            return 0
        return run_markers(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
