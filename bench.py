#!/usr/bin/env python
"""bench.py -- headline benchmark of the smoothMesh hot path on B200.

Metric (BASELINE.json): mesh point-updates/s per smoothing iteration.
  step      = one smoothing iteration (the loop body of src/smoothMesh.C:2257-2437) over the mesh
  workload  = BASELINE config 3 by default: blockMesh-numbered 200^3 hex block on the unit cube per GPU,
              interior points jittered U(-0.25h, 0.25h) (seed 12345), all constraints on, default options,
              -relTol 0 so that every one of the K iterations runs.  --size 368 is the 50 M-cell headline,
              --size 271 config 5 (weak scaling, generated per rank), --workload kelvin --size 246 config 4
              (polyhedral, 2 x 246^3 cells over the GPUs, generated per rank)
  value     = nPoints(all ranks) * K / device time of the K iterations, inputs resident in HBM
  e2e       = the same job through the host-buffer C ABI: upload points, K iterations, download points +
              per-iteration log, host<->device copies inside the timed region; e2e_cold adds mesh flattening,
              connectivity build and the one-time upload (smgpu_create)
  parity    = before anything is timed, a small case of the same kind (<= 24^3 cells per rank, tight angle
              limits so the freeze paths are busy) runs through the same library path -- NCCL exchanges
              included when N > 1 -- and is compared bit for bit with the CPU oracle's rank emulation

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size CELLS_PER_SIDE]
                    [--workload hex|kelvin]

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 12345
JITTER = 0.25
METRIC = "mesh point-updates/s per smoothing iteration"
SM_COUNT, FP64_OP_PER_CLK_SM = 148, 62.3  # profiles/microbench/r1_fp64_rates_b200.txt (DADD / DMUL / DFMA issue rate)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_name(args, world):
    if args.workload == "hex":
        return (f"hex {args.n}^3 jittered blockMesh block per GPU (U(-0.25h,0.25h), seed {SEED}), "
                f"edge/face angle constraints on, relTol 0")
    return (f"Kelvin-cell polyhedral mesh, 2x{args.n}^3 cells in total, jittered 0.2 x shortest edge, lattice bricks, "
            f"edge/face angle constraints on, relTol 0")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_used_gb(device):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        return pynvml.nvmlDeviceGetMemoryInfo(h).used / 1e9
    except Exception:
        return None


def algorithmic_bytes(P, C, F, Fi, E, FV, PC, EC):
    """Per-iteration algorithmic HBM bytes, SURVEY.md 8(d) (every resident array touched once per
    phase that needs it; int32 labels, FP64 coordinates), split per kernel in DESIGN.md."""
    state = 24 * (2 * P + 2 * P + 2 * C) + 2 * P
    conn = 4 * ((FV + F + 1) + (F + Fi) + (PC + P + 1) + (2 * E + P + 1) + (FV + P + 1) + 2 * E
                + (FV + E + 1) + (EC + E + 1) + (F + Fi + C + 1) + (2 * E + P + 1))
    return state + conn


def kernel_bytes(P, C, F, Fi, E, FV, PC, EC, tiles):
    """Algorithmic bytes per launch of each kernel."""
    Fl, Pl = tiles.get("listed_faces", 0), tiles.get("listed_points", 0)
    return {
        # points gathered once (24P), faces CSR read, face records written (centre, area, vertex mean)
        "k_face_geom": 24 * P + 4 * (FV + F + 1) + 72 * F,
        # face centre + area read once per side (48 B x (F + Fi)), cell->faces list, cellCtr written
        "k_cell_centres": 48 * (F + Fi) + 4 * (F + Fi + C + 1) + 24 * C,
        # cellCtr + points read, pointCells + pointPoints CSR, newPoints written, per-point state reset
        "k_predict": 24 * C + 24 * P + 4 * (PC + P + 1) + 4 * (2 * E + P + 1) + 24 * P + 18 * P,
        # points + newPoints read, pointPoints CSR + corner table (2 words per point-face), mask written
        "k_edge_constraints": 48 * P + 4 * (2 * E + P + 1) + 4 * (2 * FV + P + 1) + P,
        # with the fused filter only the suspect flags are read (one byte per point); without it: points +
        # cellCtr + face means, edges, edgeFaces, edgeCells(+pairs)
        "k_face_current": P if tiles.get("fused") else 24 * P + 24 * C + 24 * F + 4 * (2 * E) + 4 * (FV + E + 1) + 4 * (2 * EC + E + 1),
        # fused face + cell geometry + face-angle certificates: points gathered once, the tile lists (point
        # labels, face words, 2-byte face vertex references), per cell its label, six 2-byte face references and
        # the 32-byte canonical hexahedron record, cell centres written; no face records, no mirrors
        "k_geom_tiles": 24 * P + 4 * Pl + (4 + 8) * Fl + (4 + 12 + 32) * C + 24 * C,
        "k_layer": 0,
        "k_active_compact": 2 * P,
        "k_face_tests": 0,
        "k_face_resolve": 0,
        "halo_exchange": 0, "x_pack": 0, "x_merge": 0, "x_frozen": 0, "x_finish": 0,
        # points + newPoints + mask read, points written
        "k_commit": 24 * P + 24 * P + P + 24 * P,
    }


# thread-level FP64-pipe instructions (DADD/DMUL/DFMA/MUFU.64/F2F) per unit, from the ncu source counters of the
# kernels at 200^3 (profiles/r2_ncu_geom_tiles_f_n200.txt): what the FP64 pipe must issue, at 62.3 per clock and SM
FP64_SLOTS = {"k_geom_tiles": ("cell", 950.0)}


def make_mesh(args, world, rank, n=None, small=False):
    """The rank's mesh of the workload (n cells per side per GPU / lattice cells in total)."""
    import smoothmesh_b200 as sm
    from smoothmesh_b200 import multi
    n = n or args.n
    if args.workload == "kelvin":
        amp = 0.2 * 2 ** 0.5 / 4.0
        if world == 1:
            return sm.Mesh.kelvin(n, 1.0).jitter(amp, SEED)
        px, py, pz = multi.brick_dims(world)
        return sm.Mesh.kelvin_part(n, 1.0, px, py, pz, rank).jitter(amp, SEED)
    if world == 1:
        return sm.Mesh.hex_block(n, n, n).jitter(JITTER / n, SEED)
    return multi.weak_scaling_part(n, world, rank, JITTER, SEED)


def parity_check(args, world, rank, local_rank, dist):
    """A small case of the same kind through the same library path (NCCL exchanges when world > 1), compared bit
    for bit with the CPU oracle's rank emulation on rank 0.  Returns the dict printed as `parity`."""
    import smoothmesh_b200 as sm
    from smoothmesh_b200 import multi
    from oracle import Oracle
    n = 16 if args.workload == "hex" else max(4, 2 * max(multi.brick_dims(world)))
    kw = dict(rel_tol=0.0, min_angle_deg=70.0, max_angle_deg=110.0)
    iters = 8
    mine = make_mesh(args, world, rank, n=n)
    g = sm.Smoother(mine, device=local_rank, **kw)
    if world > 1:
        multi.init_comm(g, rank, world, dist)
    log = g.iterate(iters)
    res = dict(n=log.iterations, nf=log.n_frozen, res=log.residual, pts=g.points(), fz=g.frozen(), mesh=mine.desc_arrays())
    if world > 1:
        multi.shutdown_comm(g, dist)
    else:
        g.close()
    allres = [res]
    if world > 1:
        allres = [None] * world
        dist.all_gather_object(allres, res)
    out = None
    if rank == 0:
        o = Oracle([a["mesh"] for a in allres] if world > 1 else allres[0]["mesh"], **kw)
        on, onf, ores = o.iterate(iters)
        ok = True
        for r, a in enumerate(allres):
            ok = ok and a["n"] == on and np.array_equal(a["nf"], onf) and np.array_equal(a["res"], ores)
            ok = ok and np.array_equal(a["fz"], o.get("frozen", r) if world > 1 else o.get("frozen"))
            ok = ok and np.array_equal(a["pts"], o.get("points", r) if world > 1 else o.get("points"))
        out = {"checked": True, "ok": bool(ok), "world": world, "against": "CPU oracle, rank emulation on the same parts",
               "case": f"{args.workload} {n} per {'GPU' if args.workload == 'hex' else 'lattice side'}, minAngle 70 maxAngle 110, {iters} iterations",
               "points": int(sum(len(a["pts"]) for a in allres)), "frozen_internal_last": int(onf[-1]),
               "bitwise": ["iterations", "nFrozenPoints", "residual", "freeze mask", "points"]}
    return out


def cpu_baseline(n_side, iters, threads):
    """Times the CPU oracle (a port of the reference; the reference itself needs OpenFOAM) on a bounded
    sample of the same workload: a jittered n_side^3 block with the same jitter rule and options."""
    import smoothmesh_b200 as sm
    from oracle import Oracle
    mesh = sm.Mesh.hex_block(n_side, n_side, n_side).jitter(JITTER / n_side, SEED)
    if threads > 1:
        dims = pow2_dims(threads)
        parts = mesh.decompose(*dims)
        used = dims[0] * dims[1] * dims[2]
        o = Oracle([p.desc_arrays() for p in parts], rel_tol=0.0, threads=used)
    else:
        o = Oracle(mesh.desc_arrays(), rel_tol=0.0)
        used = 1
    t0 = time.perf_counter()
    n, _, _ = o.iterate(iters)
    dt = time.perf_counter() - t0
    return mesh.n_points * n / dt, used, dt


def pow2_dims(threads):
    """Brick counts per axis: the largest power of two <= threads, split as evenly as possible."""
    p2 = 1
    while p2 * 2 <= threads:
        p2 *= 2
    dims = [1, 1, 1]
    i = 0
    while p2 > 1:
        dims[i % 3] *= 2
        p2 //= 2
        i += 1
    return dims


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "smoothMesh_ref")


def reference_tu_run(case_dir, parallel, warm, steps, budget_s):
    """Runs oracle/_ref/smoothMesh_ref (the reference's own translation unit compiled against the OpenFOAM facade;
    prebuilt, nothing under /root/reference is read at run time) on a prepared case and times its iteration loop
    from the 'Smoothing iteration=' lines it prints (src/smoothMesh.C:2396), one per iteration as they arrive:
    the first `warm` lines are warm-up, the following ones are timed.  The process is stopped once warm + steps
    lines have arrived or the wall budget is spent.  Returns (timed iterations, seconds, set-up seconds) or None."""
    cmd = [REF_BIN, "-case", case_dir] + (["-parallel"] if parallel else []) + [
        "-centroidalIters", str(warm + steps), "-relTol", "0", "-smoothingPatches", "()"]
    t0 = time.perf_counter()
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    stamps = []
    killer = threading.Timer(budget_s, p.kill)
    killer.start()
    try:
        for line in p.stdout:
            if line.startswith("Smoothing iteration="):
                stamps.append(time.perf_counter())
                if len(stamps) >= warm + steps:
                    break
    finally:
        killer.cancel()
        p.kill()
        p.wait()
    if len(stamps) < warm + 1:
        return None
    n = len(stamps) - warm
    return n, stamps[-1] - stamps[warm - 1], stamps[0] - t0


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores, on the SAME
    workload as the GPU arm (the per-GPU-sized mesh; a CPU's rate per point does not depend on how many such
    meshes a multi-GPU job holds).  A real build of the reference needs OpenFOAM + wmake; its translation unit
    compiled against the OpenFOAM facade (oracle/_ref) is what runs here, the way the reference scales on a CPU:
    `-parallel`, one process per host core (power of two), plus, on the same mesh, the same binary serially
    (time-boxed) and the oracle port in rank emulation on the same parts.  `value` is the fastest of them."""
    if rank != 0:
        return
    import shutil
    import tempfile
    import smoothmesh_b200 as sm
    from oracle import Oracle
    threads = os.cpu_count() or 1
    dims = pow2_dims(min(threads, 64))  # the facade forks at most 64 rank processes
    procs = dims[0] * dims[1] * dims[2]
    n = args.ref_n if args.ref_n > 0 else args.n
    warm, steps = max(args.warmup, 1), args.steps
    if args.workload == "kelvin":
        mesh = sm.Mesh.kelvin(min(n, 64), 1.0).jitter(0.2 * 2 ** 0.5 / 4.0, SEED)
        parts = mesh.decompose(procs, method="rcb") if procs > 1 else [mesh]
    else:
        mesh = sm.Mesh.hex_block(n, n, n).jitter(JITTER / n, SEED)
        parts = mesh.decompose(*dims) if procs > 1 else [mesh]
    P = mesh.n_points
    rates = {}
    tmp = tempfile.mkdtemp(prefix="smref_")
    try:
        if os.path.exists(REF_BIN):
            mesh.write(os.path.join(tmp, "constant", "polyMesh"), binary=True)
            os.makedirs(os.path.join(tmp, "system"))
            with open(os.path.join(tmp, "system", "controlDict"), "w") as f:
                f.write("startFrom startTime;\nstartTime 0;\ndeltaT 1;\nwriteFormat binary;\n")
            if procs > 1:
                sm.Mesh.write_decomposed(parts, tmp, binary=True)
                r = reference_tu_run(tmp, True, warm, steps, args.ref_budget)
                if r:
                    rates["reference_tu_parallel"] = dict(value=P * r[0] / r[1], cores=procs, kind="reference", iterations=r[0],
                                                          seconds=r[1], setup_s=r[2],
                                                          how=f"-parallel, {procs} rank processes ({'x'.join(map(str, dims))} bricks)")
            r = reference_tu_run(tmp, False, 1, max(2, min(steps, 3)), args.ref_serial_budget)
            if r:
                rates["reference_tu_serial"] = dict(value=P * r[0] / r[1], cores=1, kind="reference", iterations=r[0],
                                                    seconds=r[1], setup_s=r[2], how="serial")
            else:
                rates["reference_tu_serial"] = dict(unavailable=f"no timed iteration within {args.ref_serial_budget:.0f} s on this mesh "
                                                                f"(its serial set-up alone takes longer); see reference_tu_serial_sample")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if os.path.exists(REF_BIN) and "value" not in rates.get("reference_tu_serial", {}) and args.workload == "hex":
        # the serial rate on a 64^3 sample of the same workload instead (rate per point is size independent)
        tmp = tempfile.mkdtemp(prefix="smref_")
        try:
            small = sm.Mesh.hex_block(64, 64, 64).jitter(JITTER / 64, SEED)
            small.write(os.path.join(tmp, "constant", "polyMesh"), binary=True)
            os.makedirs(os.path.join(tmp, "system"))
            with open(os.path.join(tmp, "system", "controlDict"), "w") as f:
                f.write("startFrom startTime;\nstartTime 0;\ndeltaT 1;\nwriteFormat binary;\n")
            r = reference_tu_run(tmp, False, 1, 3, 120.0)
            if r:
                rates["reference_tu_serial_sample"] = dict(value=small.n_points * r[0] / r[1], cores=1, kind="reference",
                                                           iterations=r[0], seconds=r[1], setup_s=r[2],
                                                           how="serial, 64^3 sample of the workload")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    if not args.ref_no_port:
        o = Oracle([p.desc_arrays() for p in parts] if procs > 1 else mesh.desc_arrays(), rel_tol=0.0, threads=procs)
        o.iterate(warm)
        t0 = time.perf_counter()
        k, _, _ = o.iterate(steps)
        dt = time.perf_counter() - t0
        rates["port"] = dict(value=P * k / dt, cores=procs, kind="port", iterations=int(k), seconds=dt,
                             how=f"CPU oracle (bit-identical to the translation unit), rank emulation, {procs} threads")
    best = max((k for k in rates if "value" in rates[k]), key=lambda k: rates[k]["value"])
    b = rates[best]
    short = f"hex {n}^3 jittered block" if args.workload == "hex" else f"Kelvin 2x{min(n, 64)}^3 jittered mesh"
    sample = (f"{b['iterations']} iterations of the {short} ({P} points) after "
              f"{warm} warm-up iterations, {b['how']}, {b['seconds']:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": b["value"], "unit": "point-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * P / b["value"], "higher_is_better": True,
        "scaling": "weak" if args.workload == "hex" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, world),
                   "note": "the per-GPU-sized mesh of the GPU arm, whole, on the host cores; value = the fastest of the three "
                           "CPU arms in cpu_baseline.arms (all on this one mesh)"},
        "cpu_baseline": {"value": b["value"], "unit": "point-updates/s", "cores": b["cores"], "kind": b["kind"],
                         "sample": sample, "fastest": best, "arms": rates},
        "e2e": {"value": b["value"], "unit": "point-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", "--n", dest="n", type=int, default=200,
                    help="cells per side of the (per-GPU) hex block / of the Kelvin lattice (use --size under torchrun)")
    ap.add_argument("--ref-n", type=int, default=0, help="cells per side of the reference arm's mesh (0 = --size)")
    ap.add_argument("--ref-budget", type=float, default=600.0, help="wall budget [s] of the parallel reference run")
    ap.add_argument("--ref-serial-budget", type=float, default=90.0, help="wall budget [s] of the serial reference run")
    ap.add_argument("--ref-no-port", action="store_true", help="skip the oracle-port arm of --impl reference")
    ap.add_argument("--cpu-n", type=int, default=96, help="cells per side of the cpu_baseline sample mesh")
    ap.add_argument("--cpu-iters", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check before the timed region")
    ap.add_argument("--renumber", type=int, default=0, help="1 = Morton storage order (smgpu_params.renumber)")
    ap.add_argument("--min-angle", type=float, default=35.0, help="-minAngle (default = reference default)")
    ap.add_argument("--max-angle", type=float, default=160.0, help="-maxAngle (default = reference default)")
    ap.add_argument("--workload", default="hex", choices=["hex", "kelvin"],
                    help="hex: n^3 jittered blockMesh block per GPU (weak scaling, BASELINE configs 3/5); kelvin: "
                         "2 n^3 Kelvin-cell polyhedral mesh in total, lattice bricks over the GPUs, generated per rank "
                         "(BASELINE config 4, strong scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # torchrun pins OMP_NUM_THREADS=1; the one-time host-side connectivity build is OpenMP-parallel
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import smoothmesh_b200 as sm

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    parity = None if args.no_parity else parity_check(args, world, rank, local_rank, dist)

    n = args.n
    t0 = time.perf_counter()
    mesh = make_mesh(args, world, rank)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = sm.Smoother(mesh, rel_tol=0.0, device=local_rank, renumber=args.renumber, min_angle_deg=args.min_angle,
                    max_angle_deg=args.max_angle)
    exchange = None
    if world > 1:
        from smoothmesh_b200 import multi
        exchange = "peer-memory (NVLink stores from the producer kernels)" if multi.init_comm(g, rank, world, dist) else "nccl"
    t_setup = time.perf_counter() - t0
    stats = g.mesh_stats()
    tiles = g.tile_stats()
    P, C, F, Fi = mesh.n_points, mesh.n_cells, mesh.n_faces, mesh.n_internal_faces
    E, FV = stats["n_edges"], int(mesh.face_offsets[-1])
    PC = int(g.csr("pointCells")[0][-1])
    EC = int(g.csr("edgeCells")[0][-1])
    init_pts = np.array(mesh.points, dtype=np.float64)
    hbm_gb = hbm_used_gb(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()

    # warm-up, then restart from the initial jittered mesh so the timed region is iterations 1..K
    barrier()
    g.iterate(args.warmup)
    g.set_points(init_pts)
    g.profile(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.4)  # let nvidia-smi start sampling before the timed region opens
    barrier()
    log = g.iterate(args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    prof = g.profile_get()
    g.profile(False)
    fstats = g.filter_stats()
    tiles["fused"] = fstats["fused"]
    ms = log.ms
    total_points = P
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tp = torch.tensor([P], dtype=torch.float64, device="cuda")
        dist.all_reduce(tp, op=dist.ReduceOp.SUM)
        total_points = int(tp.item())
    K = log.iterations
    value = total_points * K / (ms * 1e-3)

    # end-to-end through the host-buffer C ABI: H2D points, K iterations, D2H points + log
    out_pts = np.empty((P, 3), dtype=np.float64)  # the caller's result buffer (a real host keeps its pointField)
    out_pts.fill(0.0)
    barrier()
    t0 = time.perf_counter()
    g.set_points(init_pts)
    t_up = time.perf_counter() - t0
    log2 = g.iterate(args.steps)
    t1 = time.perf_counter()
    g.points(out=out_pts)
    t_e2e = time.perf_counter() - t0
    t_down = time.perf_counter() - t1
    barrier()
    t_cold = t_setup + t_e2e
    if dist is not None:
        import torch
        t = torch.tensor([t_e2e, t_cold, t_setup, t_gen], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e, t_cold, t_setup_max, t_gen_max = (float(x) for x in t.tolist())
    else:
        t_setup_max, t_gen_max = t_setup, t_gen
    e2e_value = total_points * log2.iterations / t_e2e
    assert np.array_equal(log2.n_frozen, log.n_frozen)
    assert np.isfinite(out_pts).all()

    if dist is not None:
        from smoothmesh_b200 import multi
        multi.shutdown_comm(g, dist)
    if rank != 0:
        dist.barrier()
        dist.destroy_process_group()
        return

    hbm, hbm_src = peaks()
    kb = kernel_bytes(P, C, F, Fi, E, FV, PC, EC, tiles)
    top = max(prof, key=lambda k: prof[k]["ms"])
    top_ms = prof[top]["ms"] / max(prof[top]["launches"], 1)
    if top == "k_active_compact":
        top_ms = prof[top]["ms"] / max(prof[top]["launches"] // 4, 1)
    achieved = kb[top] / (top_ms * 1e-3) / 1e9
    iter_bytes = algorithmic_bytes(P, C, F, Fi, E, FV, PC, EC)
    traffic = None  # measured DRAM bytes per launch of the dominant kernel (one ncu --set full capture)
    tpath = os.path.join(ROOT, "profiles", "r2_traffic_n200.json")
    if args.workload == "hex" and n == 200 and os.path.exists(tpath):
        traffic = json.load(open(tpath))["bytes_per_launch"].get(top)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp64 = None
    if top in FP64_SLOTS:
        unit, per = FP64_SLOTS[top]
        slots = per * (C if unit == "cell" else P)
        fp64 = {"issue_slots": slots, "peak_op_per_clk_sm": FP64_OP_PER_CLK_SM, "sm_mhz": sm_mhz,
                "frac": slots / (top_ms * 1e-3 * sm_mhz * 1e6 * SM_COUNT * FP64_OP_PER_CLK_SM),
                "note": "thread-level FP64-pipe instructions of one launch (ncu source counters at 200^3, scaled by cells) "
                        "over what the FP64 pipes of 148 SMs can issue in the launch's duration"}
    hbm_frac = achieved / hbm
    line = {
        "metric": METRIC, "value": value, "unit": "point-updates/s",
        "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": ms / K, "higher_is_better": True,
        "scaling": "weak" if args.workload == "hex" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, world),
                   "iterations": K, "points_per_gpu": P, "cells_per_gpu": C, "points_total": total_points,
                   "renumber": args.renumber, "min_angle": args.min_angle, "max_angle": args.max_angle,
                   "l2": "working set (>= 4 GB per iteration) exceeds the 126 MB L2",
                   "setup_s": {"mesh_generation": t_gen_max, "create_upload": t_setup_max},
                   "hbm_resident_gb": hbm_gb, "exchange": exchange,
                   "filter": fstats},
        "clocks": clocks,
        "parity": parity,
        "e2e": {"value": e2e_value, "unit": "point-updates/s", "h2d_bytes_per_step": 24 * P / K,
                "d2h_bytes_per_step": (24 * P + 16 * K) / K,
                "upload_s": t_up, "download_s": t_down,
                "note": "points uploaded once, K iterations, points + log downloaded once; bytes amortised over K steps"},
        "e2e_cold": {"value": total_points * K / t_cold, "unit": "point-updates/s", "seconds": t_cold,
                     "note": "smgpu_create (mesh flattening, connectivity build, tiles, one-time upload) + upload of the points + "
                             "K iterations + download; mesh generation excluded"},
        "gpu_launches": int(log.launches),
        "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": hbm_frac, "traffic": traffic, "peak_source": hbm_src,
                     "algorithmic_bytes_per_launch": kb[top], "avg_launch_ms": top_ms,
                     "fp64": fp64,
                     "binding": ("fp64 pipe / issue" if fp64 and fp64["frac"] > hbm_frac else "hbm"),
                     "whole_iteration": {"algorithmic_bytes": iter_bytes,
                                         "achieved": iter_bytes / (ms / K * 1e-3) / 1e9,
                                         "frac": iter_bytes / (ms / K * 1e-3) / 1e9 / hbm}},
        "kernel_ms_per_step": {k: v["ms"] / K for k, v in prof.items()},
    }
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        v1, c1, dt1 = cpu_baseline(args.cpu_n, args.cpu_iters, 1)
        vN, cN, dtN = cpu_baseline(args.cpu_n, args.cpu_iters, threads)
        line["cpu_baseline"] = {"value": vN, "unit": "point-updates/s", "cores": cN, "kind": "port",
                                "sample": f"{args.cpu_iters} iterations of a jittered {args.cpu_n}^3 hex block "
                                          f"(same jitter rule/options), rank-emulation on {cN} threads ({dtN:.1f} s); "
                                          f"serial: {v1:.4g} point-updates/s ({dt1:.1f} s); the whole workload on the CPU: "
                                          f"bench.py --impl reference",
                                "serial_value": v1}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
