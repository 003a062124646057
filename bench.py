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
  roofline  finest-level 8-colour Gauss-Seidel relaxation kernel (ccu_k_relax_tab): algorithmic bytes
          648 B/node/sweep / 8 colour passes per launch, over its mean launch
          duration measured live with CUDA events inside the timed region; `traffic` = DRAM bytes
          per launch from the committed ncu --set full capture (profiles/traffic.json)
  cpu_baseline / --impl reference  the UNMODIFIED reference (oracle/_ref, built from
          /root/reference by oracle/Makefile over a process-based MPI shim) running the same step on
          the host cores on a bounded sample (a 1/64-size mesh of the same configuration), scaled
          to the full mesh by element count.
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

RELAX_BYTES_PER_NODE_SWEEP = 648.0     # K 504 + F 24 + BI 24 + d0 r/w 48 + Ad r/w 48
MATVEC_BYTES_PER_NODE = 552.0          # K 504 + u 24 + Au 24


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh", default="256x256x128")
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--ref-mesh", default="64x64x32", help="bounded sample mesh for the CPU reference")
    ap.add_argument("--ref-levels", type=int, default=4)
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


def reference_times(ref_mesh, ref_levels, repeats, cores):
    """Run the unmodified reference's general_stokes_solver (zero guess) `repeats` times on the sample mesh."""
    from oracle import pyoracle as po
    from citcomcu_b200 import inputfile
    if not po.have_ref():
        raise RuntimeError("oracle/_ref is not built (python -c 'import __graft_entry__ as g; g.build()' where /root/reference exists)")
    f = 2 ** (ref_levels - 1)
    mg = (ref_mesh[0] // f, ref_mesh[1] // f, ref_mesh[2] // f)
    nproc = nproc_for_cores(cores, mg)
    nranks = nproc[0] * nproc[1] * nproc[2]
    txt = inputfile.tdepv_box(*ref_mesh, ref_levels, nproc=nproc, maxstep=1)
    wd = Path(tempfile.mkdtemp(prefix="ccu_refbench_"))
    times = po.run_timezero(txt, wd, repeats, nproc=nranks)
    return times, nproc, nranks


def run_reference(args):
    mesh, ref_mesh = mesh_tuple(args.mesh), mesh_tuple(args.ref_mesh)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    times, nproc, nranks = reference_times(ref_mesh, args.ref_levels, args.warmup + args.steps, cores)
    t = times[args.warmup:]
    scale = (mesh[0] * mesh[1] * mesh[2]) / (ref_mesh[0] * ref_mesh[1] * ref_mesh[2])
    v = float(np.mean(t)) * scale
    sample = (f"unmodified reference general_stokes_solver on a {args.ref_mesh} mesh of the same configuration, {nranks} ranks "
              f"({nproc[0]}x{nproc[1]}x{nproc[2]}) over a shared-memory MPI shim; {float(np.mean(t)):.3f} s per step measured, "
              f"scaled x{scale:g} by element count to {args.mesh}")
    line = {"impl": "reference", "metric": "stokes_solve_s_per_timestep", "value": v, "unit": "s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, mesh),
            "cpu_baseline": {"value": v, "unit": "s", "cores": nranks, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(args, mesh, nproc=(1, 1, 1)):
    return {"workload": f"BASELINE config 3: 3-D Cartesian {mesh[0]}x{mesh[1]}x{mesh[2]} elements, Ra=1e7, TDEPV contrast 1e5, "
                        f"free slip, {args.levels} multigrid levels; step = general_stokes_solver from a zero guess "
                        "(viscosity + stiffness rebuild + forces + Uzawa/FMG solve to accuracy 1e-3)",
            "mesh": list(mesh), "levels": args.levels, "nproc": list(nproc),
            "partition": "one subdomain per GPU, the reference's nprocx x nprocy x nprocz block decomposition; halo sums + allreduce over NCCL",
            "l2": "inputs larger than L2 (finest-level stiffness alone is > 4 GB)"}


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
    ctx.profile_enable(True)
    ctx.profile_reset()
    l0 = ctx.launch_count
    total_s, its = timed(step_resident, args.steps)
    launches = ctx.launch_count - l0
    relax_ms, relax_n = ctx.profile_read("relax_fine")
    mv_ms, mv_n = ctx.profile_read("matvec_fine")
    build_ms, _ = ctx.profile_read("build")
    coarse_ms, coarse_n = ctx.profile_read("coarse")
    transfer_ms, _ = ctx.profile_read("transfer_fine")
    level_ms = {lev: ctx.profile_read(lev) for lev in range(prob.levmin, prob.levmax + 1)} if args.no_graphs else None
    ctx.profile_enable(False)
    clk = clocks.stop()
    # e2e: host buffers through the C ABI
    step_e2e()
    e2e_s, _ = timed(step_e2e, args.steps)
    s_per_step = total_s / args.steps
    peak, peak_src = measured_peak_gbs()
    relax_bytes_per_launch = RELAX_BYTES_PER_NODE_SWEEP * nno / 8.0
    relax_gbs = relax_bytes_per_launch * relax_n / (relax_ms * 1e-3) / 1e9 if relax_ms > 0 else 0.0
    mv_gbs = MATVEC_BYTES_PER_NODE * nno * mv_n / (mv_ms * 1e-3) / 1e9 if mv_ms > 0 else 0.0
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            t = json.loads(tf.read_text()).get("ccu_k_relax_tab", {}).get(args.mesh) if world == 1 else None
            traffic = None if t is None else t["dram_bytes_read_per_launch"] + t["dram_bytes_write_per_launch"]
        except Exception:
            traffic = None
    line = {"metric": "stokes_solve_s_per_timestep", "value": s_per_step, "unit": "s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, mesh, nproc),
            "roofline": {"bound": "hbm", "kernel": "ccu_k_relax_tab<2> (one colour pass of the finest-level 8-colour Gauss-Seidel smoother)",
                         "achieved": relax_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": relax_gbs / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": relax_bytes_per_launch,
                         "launches": relax_n, "avg_launch_ms": relax_ms / max(relax_n, 1),
                         "share_of_step": relax_ms / (total_s * 1e3)},
            "smoother_gbs": relax_gbs, "matvec_gbs": mv_gbs,
            "matvec": {"achieved": mv_gbs, "frac": mv_gbs / peak, "launches": mv_n, "avg_launch_ms": mv_ms / max(mv_n, 1),
                       "share_of_step": mv_ms / (total_s * 1e3)},
            "operator_rebuild_ms_per_step": build_ms / args.steps,
            "step_breakdown_ms": {"relax_fine": relax_ms / args.steps, "matvec_fine": mv_ms / args.steps, "build": build_ms / args.steps,
                                  "coarse_levels": coarse_ms / args.steps, "transfer_fine": transfer_ms / args.steps,
                                  "other_fine_vector_ops_and_sync": (total_s * 1e3 - relax_ms - mv_ms - build_ms - coarse_ms - transfer_ms) / args.steps},
            "uzawa_iterations": its, "gpu_launches": launches * world, "clocks": clk, "setup_s": setup_s,
            "e2e": {"value": e2e_s / args.steps, "unit": "s", "h2d_bytes_per_step": int(T_h.nbytes + b_h.nbytes) * world,
                    "d2h_bytes_per_step": int(U_h.nbytes + P_h.nbytes) * world}}
    if level_ms:
        line["level_ms_per_step"] = {str(k): {"ms": v[0] / args.steps, "sweeps": v[1] / args.steps} for k, v in level_ms.items()}
    if not args.no_cpu_baseline and rank == 0:
        try:
            ref_mesh = mesh_tuple(args.ref_mesh)
            cores = os.cpu_count() or 1
            times, nproc, nranks = reference_times(ref_mesh, args.ref_levels, 2, cores)
            scale = (mesh[0] * mesh[1] * mesh[2]) / (ref_mesh[0] * ref_mesh[1] * ref_mesh[2])
            line["cpu_baseline"] = {"value": float(times[-1]) * scale, "unit": "s", "cores": nranks, "kind": "reference",
                                    "sample": f"unmodified reference (oracle/_ref) general_stokes_solver on a {args.ref_mesh} mesh of the same "
                                              f"configuration, {nranks} ranks; {times[-1]:.3f} s measured, scaled x{scale:g} by element count"}
        except Exception as e:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
