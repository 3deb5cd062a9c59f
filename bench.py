#!/usr/bin/env python
"""bench.py -- headline benchmark of the smoothMesh hot path on B200.

Metric (BASELINE.json): mesh point-updates/s per smoothing iteration.
  step      = one smoothing iteration (the loop body of src/smoothMesh.C:2257-2437) over the mesh
  workload  = BASELINE config 3: blockMesh-numbered 200^3 hex block on the unit cube, interior
              points jittered U(-0.25h, 0.25h) (seed 12345), all constraints on, default options,
              -relTol 0 so that every one of the K iterations runs
  value     = nPoints(all ranks) * K / device time of the K iterations, inputs resident in HBM
  e2e       = the same job through the host-buffer C ABI: upload points, K iterations, download
              points + per-iteration log, host<->device copies inside the timed region

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size CELLS_PER_SIDE]

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


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def algorithmic_bytes(P, C, F, Fi, E, FV, PC, EC):
    """Per-iteration algorithmic HBM bytes, SURVEY.md 8(d) (every resident array touched once per
    phase that needs it; int32 labels, FP64 coordinates), split per kernel in DESIGN.md."""
    state = 24 * (2 * P + 2 * P + 2 * C) + 2 * P
    conn = 4 * ((FV + F + 1) + (F + Fi) + (PC + P + 1) + (2 * E + P + 1) + (FV + P + 1) + 2 * E
                + (FV + E + 1) + (EC + E + 1) + (F + Fi + C + 1) + (2 * E + P + 1))
    return state + conn


def kernel_bytes(P, C, F, Fi, E, FV, PC, EC):
    """Algorithmic bytes per launch of each kernel."""
    return {
        # points gathered once (24P), faces CSR read, face records written (centre, area, vertex mean)
        "k_face_geom": 24 * P + 4 * (FV + F + 1) + 72 * F,
        # face centre + area read once per side (48 B x (F + Fi)), cell->faces list, cellCtr written
        "k_cell_centres": 48 * (F + Fi) + 4 * (F + Fi + C + 1) + 24 * C,
        # cellCtr + points read, pointCells + pointPoints CSR, newPoints written, per-point state reset
        "k_predict": 24 * C + 24 * P + 4 * (PC + P + 1) + 4 * (2 * E + P + 1) + 24 * P + 18 * P,
        # points + newPoints read, pointPoints CSR + corner table (2 words per point-face), mask written
        "k_edge_constraints": 48 * P + 4 * (2 * E + P + 1) + 4 * (2 * FV + P + 1) + P,
        # points + cellCtr + face means read, edges, edgeFaces, edgeCells(+pairs)
        "k_face_current": 24 * P + 24 * C + 24 * F + 4 * (2 * E) + 4 * (FV + E + 1) + 4 * (2 * EC + E + 1),
        # fused face + cell geometry: points gathered once, faces CSR, tile lists (cells, faces, 2-byte face
        # references), vertex means (fp32 mirror) and cell centres (fp64 + fp32 mirror) written; no face records
        "k_geom_tiles": 24 * P + 4 * (FV + F + 1) + 4 * (C + F) + 2 * (F + Fi) + 16 * F + (24 + 16) * C,
        "k_layer": 0,
        "k_active_compact": 2 * P,
        "k_face_tests": 0,
        "k_face_resolve": 0,
        "halo_exchange": 0,
        # points + newPoints + mask read, points written
        "k_commit": 24 * P + 24 * P + P + 24 * P,
    }


def cpu_baseline(n_side, iters, threads):
    """Times the CPU oracle (a port of the reference; the reference itself needs OpenFOAM) on a bounded
    sample of the same workload: a jittered n_side^3 block with the same jitter rule and options."""
    import smoothmesh_b200 as sm
    from oracle import Oracle
    mesh = sm.Mesh.hex_block(n_side, n_side, n_side).jitter(JITTER / n_side, SEED)
    if threads > 1:
        # rank emulation: decomposePar-style bricks, one thread per part (the reference's mpirun mode)
        px = py = pz = 1
        t = threads
        while t % 2 == 0 and t > 1:
            if px <= py and px <= pz:
                px *= 2
            elif py <= pz:
                py *= 2
            else:
                pz *= 2
            t //= 2
        parts = mesh.decompose(px, py, pz)
        o = Oracle([p.desc_arrays() for p in parts], rel_tol=0.0, threads=px * py * pz)
        used = px * py * pz
    else:
        o = Oracle(mesh.desc_arrays(), rel_tol=0.0)
        used = 1
    t0 = time.perf_counter()
    n, _, _ = o.iterate(iters)
    dt = time.perf_counter() - t0
    return mesh.n_points * n / dt, used, dt


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "smoothMesh_ref")


def reference_binary_rate(n_side=40, iters=(1, 5), procs=1):
    """Rate of oracle/_ref/smoothMesh_ref, the reference's own translation unit compiled against the OpenFOAM
    facade (prebuilt; nothing under /root/reference is read at run time): two runs with different iteration counts
    on the same jittered block, the difference of their wall times is the loop alone.  procs > 1 runs it the way
    the reference scales on a CPU -- `-parallel`, one process per processor directory of a brick decomposition
    (the facade forks them and combines interface points through shared memory).  Returns None when the binary is
    absent."""
    if not os.path.exists(REF_BIN):
        return None
    import shutil
    import subprocess
    import tempfile
    import smoothmesh_b200 as sm
    mesh = sm.Mesh.hex_block(n_side, n_side, n_side).jitter(JITTER / n_side, SEED)
    tmp = tempfile.mkdtemp(prefix="smref_")
    try:
        mesh.write(os.path.join(tmp, "constant", "polyMesh"))
        os.makedirs(os.path.join(tmp, "system"))
        with open(os.path.join(tmp, "system", "controlDict"), "w") as f:
            f.write("startFrom startTime;\nstartTime 0;\ndeltaT 1;\nwriteFormat binary;\n")
        extra = []
        dims = [1, 1, 1]
        if procs > 1:
            t, i = procs, 0
            while t % 2 == 0 and t > 1:
                dims[i % 3] *= 2
                t //= 2
                i += 1
            sm.Mesh.write_decomposed(mesh.decompose(*dims), tmp, binary=True)
            extra = ["-parallel"]
        times = []
        for k in iters:
            t0 = time.perf_counter()
            subprocess.run([REF_BIN, "-case", tmp] + extra + ["-centroidalIters", str(k), "-relTol", "0", "-smoothingPatches", "()"],
                           check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
            times.append(time.perf_counter() - t0)
        dt = times[1] - times[0]
        if dt <= 0:
            return None
        used = dims[0] * dims[1] * dims[2]
        how = "serial" if used == 1 else f"-parallel, {used} processes ({'x'.join(map(str, dims))} bricks)"
        return dict(value=mesh.n_points * (iters[1] - iters[0]) / dt, cores=used,
                    sample=f"{iters[1] - iters[0]} iterations of a jittered {n_side}^3 hex block ({dt:.1f} s), {how}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  A real build of the reference needs
    OpenFOAM + wmake; its translation unit compiled against the OpenFOAM facade (oracle/_ref) is bit-identical to
    the oracle port.  When that binary exists it is what this arm times: `-parallel` with one process per host core
    (the reference's `mpirun` strategy), loop time = difference of two runs; the (faster) oracle port's rate is
    reported next to it in cpu_baseline.port.  Without the binary the oracle port is timed in rank-emulation mode
    with all host threads on a bounded sample."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import smoothmesh_b200 as sm
    from oracle import Oracle
    p2 = 1
    while p2 * 2 <= threads:
        p2 *= 2
    dims = [1, 1, 1]
    t = p2
    i = 0
    while t > 1:
        dims[i % 3] *= 2
        t //= 2
        i += 1

    def build(n_side):
        mesh = sm.Mesh.hex_block(n_side, n_side, n_side).jitter(JITTER / n_side, SEED)
        parts = mesh.decompose(*dims) if p2 > 1 else [mesh]
        return mesh, Oracle([p.desc_arrays() for p in parts], rel_tol=0.0, threads=p2)

    # the reference's own translation unit (oracle/_ref) as one process per host core, when it was built
    if os.path.exists(REF_BIN) and not os.environ.get("SMBENCH_REFERENCE_PORT"):
        n_ref = args.ref_n if args.ref_n > 0 else 64
        warm = max(args.warmup, 1)
        try:  # the facade forks at most 64 rank processes; any failure falls back to the port below
            ref = reference_binary_rate(n_side=n_ref, iters=(warm, warm + args.steps), procs=min(p2, 64))
        except Exception:
            ref = None
        if ref is not None:
            mesh, o = build(32)
            t0 = time.perf_counter()
            o.iterate(4)
            port_rate = mesh.n_points * 4 / (time.perf_counter() - t0)
            value = ref["value"]
            pts_ref = (n_ref + 1) ** 3
            line = {
                "impl": "reference", "metric": "mesh point-updates/s per smoothing iteration", "value": value,
                "unit": "point-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
                "ms_per_step": 1e3 * pts_ref / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"hex {n_ref}^3 jittered block (bounded sample of the 200^3 config), all constraints on",
                           "decomposition": ref["sample"].split("(")[-1].rstrip(")") if "bricks" in ref["sample"] else "1x1x1",
                           "note": "the reference's own translation unit (src/smoothMesh.C compiled in place against the OpenFOAM "
                                   "facade, oracle/_ref) run as `-parallel` rank processes; the loop time is the difference of two "
                                   "runs (warm-up only / warm-up + steps), which cancels mesh reading and set-up"},
                "cpu_baseline": {"value": value, "unit": "point-updates/s", "cores": ref["cores"], "kind": "reference",
                                 "sample": ref["sample"],
                                 "port": {"value": port_rate, "unit": "point-updates/s", "cores": p2, "kind": "port",
                                          "sample": "4 iterations of a jittered 32^3 hex block, oracle port in rank emulation"}},
                "e2e": {"value": value, "unit": "point-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            }
            print(json.dumps(line), flush=True)
            return

    # bounded sample: size the block so that warmup + steps iterations take about args.ref_budget
    # seconds at the rate calibrated on a 32^3 block (per-point cost is size independent)
    n_side = args.ref_n
    if n_side <= 0:
        mesh, o = build(32)
        t0 = time.perf_counter()
        o.iterate(2)
        rate = mesh.n_points * 2 / (time.perf_counter() - t0)
        pts = args.ref_budget * rate / max(args.steps + args.warmup, 1)
        n_side = int(min(128, max(24, round(pts ** (1.0 / 3.0)) - 1)))
    mesh, o = build(n_side)
    o.iterate(args.warmup)
    t0 = time.perf_counter()
    n, _, _ = o.iterate(args.steps)
    dt = time.perf_counter() - t0
    value = mesh.n_points * n / dt
    line = {
        "impl": "reference", "metric": "mesh point-updates/s per smoothing iteration", "value": value,
        "unit": "point-updates/s", "n_gpus": args.gpus, "steps": n, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(n, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"hex {n_side}^3 jittered block (bounded sample of the 200^3 config), all constraints on",
                   "decomposition": "x".join(map(str, dims)), "note": "CPU oracle port in rank-emulation mode on all host threads (bit-identical to "
                   "the reference's translation unit, oracle/_ref, which is single-threaded without MPI and therefore the "
                   "slower arm; its serial rate is in cpu_baseline.reference_tu)"},
        "cpu_baseline": {"value": value, "unit": "point-updates/s", "cores": p2, "kind": "port",
                         "sample": f"{n} iterations of a jittered {n_side}^3 hex block, {p2} threads",
                         "reference_tu": reference_binary_rate(),
                         "reference_tu_parallel": reference_binary_rate(n_side=64, iters=(1, 4), procs=min(p2, 64))},
        "e2e": {"value": value, "unit": "point-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    ap.add_argument("--ref-n", type=int, default=0, help="cells per side of the CPU sample mesh (0 = sized from --ref-budget)")
    ap.add_argument("--ref-budget", type=float, default=100.0, help="seconds of CPU work for --impl reference")
    ap.add_argument("--cpu-n", type=int, default=96, help="cells per side of the cpu_baseline sample mesh")
    ap.add_argument("--cpu-iters", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--renumber", type=int, default=0, help="1 = Morton storage order (smgpu_params.renumber)")
    ap.add_argument("--min-angle", type=float, default=35.0, help="-minAngle (default = reference default)")
    ap.add_argument("--max-angle", type=float, default=160.0, help="-maxAngle (default = reference default)")
    ap.add_argument("--workload", default="hex", choices=["hex", "kelvin"],
                    help="hex: n^3 jittered blockMesh block per GPU (weak scaling, BASELINE configs 3/5); kelvin: "
                         "2 n^3 Kelvin-cell polyhedral mesh in total, RCB-decomposed over the GPUs (BASELINE config 4, "
                         "strong scaling)")
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

    n = args.n
    t0 = time.perf_counter()
    if args.workload == "kelvin":
        # every rank builds the same global mesh and keeps its RCB part (decomposePar stand-in)
        whole = sm.Mesh.kelvin(n, 1.0).jitter(0.2 * 2 ** 0.5 / 4.0, SEED)
        mesh = whole.decompose(world, method="rcb")[rank] if world > 1 else whole
        del whole
    elif world == 1:
        mesh = sm.Mesh.hex_block(n, n, n).jitter(JITTER / n, SEED)
    else:
        from smoothmesh_b200 import multi
        mesh = multi.weak_scaling_part(n, world, rank, JITTER, SEED)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = sm.Smoother(mesh, rel_tol=0.0, device=local_rank, renumber=args.renumber, min_angle_deg=args.min_angle,
                    max_angle_deg=args.max_angle)
    if world > 1:
        from smoothmesh_b200 import multi
        multi.init_comm(g, rank, world, dist)
    t_setup = time.perf_counter() - t0
    stats = g.mesh_stats()
    P, C, F, Fi = mesh.n_points, mesh.n_cells, mesh.n_faces, mesh.n_internal_faces
    E, FV = stats["n_edges"], int(mesh.face_offsets[-1])
    PC = int(g.csr("pointCells")[0][-1])
    EC = int(g.csr("edgeCells")[0][-1])
    init_pts = np.array(mesh.points, dtype=np.float64)

    def barrier():
        if dist is not None:
            dist.barrier()

    # warm-up, then restart from the initial jittered mesh so the timed region is iterations 1..K
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
    barrier()
    t0 = time.perf_counter()
    g.set_points(init_pts)
    log2 = g.iterate(args.steps)
    out_pts = g.points()
    t_e2e = time.perf_counter() - t0
    barrier()
    if dist is not None:
        import torch
        t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = total_points * log2.iterations / t_e2e
    assert np.array_equal(log2.n_frozen, log.n_frozen)
    assert np.isfinite(out_pts).all()

    if rank != 0:
        dist.barrier()
        dist.destroy_process_group()
        return

    hbm, hbm_src = peaks()
    kb = kernel_bytes(P, C, F, Fi, E, FV, PC, EC)
    top = max(prof, key=lambda k: prof[k]["ms"])
    top_ms = prof[top]["ms"] / max(prof[top]["launches"], 1)
    if top == "k_active_compact":
        top_ms = prof[top]["ms"] / max(prof[top]["launches"] // 4, 1)
    achieved = kb[top] / (top_ms * 1e-3) / 1e9
    iter_bytes = algorithmic_bytes(P, C, F, Fi, E, FV, PC, EC)
    traffic = None  # measured DRAM bytes per launch of the dominant kernel (one ncu --set full capture)
    tpath = os.path.join(ROOT, "profiles", "r1_traffic_n200.json")
    if args.workload == "hex" and n == 200 and os.path.exists(tpath):
        traffic = json.load(open(tpath))["bytes_per_launch"].get(top)
    line = {
        "metric": "mesh point-updates/s per smoothing iteration", "value": value, "unit": "point-updates/s",
        "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": ms / K, "higher_is_better": True,
        "scaling": "weak" if args.workload == "hex" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"hex {n}^3 jittered blockMesh block per GPU (U(-0.25h,0.25h), seed {SEED}), "
                                if args.workload == "hex" else
                                f"Kelvin-cell polyhedral mesh, 2x{n}^3 cells in total, jittered 0.2 x shortest edge, RCB parts, ")
                               + f"{K} iterations, edge/face angle constraints on, relTol 0",
                   "points_per_gpu": P, "cells_per_gpu": C, "renumber": args.renumber, "min_angle": args.min_angle, "max_angle": args.max_angle, "l2": "working set (>= 4 GB per iteration) exceeds the 126 MB L2",
                   "setup_s": {"mesh_generation": t_gen, "create_upload": t_setup}},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "point-updates/s", "h2d_bytes_per_step": 24 * P / K,
                "d2h_bytes_per_step": (24 * P + 16 * K) / K,
                "note": "points uploaded once, K iterations, points + log downloaded once; bytes amortised over K steps"},
        "gpu_launches": int(log.launches),
        "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": achieved / hbm, "traffic": traffic, "peak_source": hbm_src,
                     "algorithmic_bytes_per_launch": kb[top], "avg_launch_ms": top_ms,
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
                                          f"serial: {v1:.4g} point-updates/s ({dt1:.1f} s)",
                                "serial_value": v1}
        try:
            ref = reference_binary_rate()
        except Exception:  # a failing baseline sample must not take the GPU line down
            ref = None
        if ref is not None:
            # the reference's own translation unit (oracle/_ref, compiled against the OpenFOAM facade), serial and
            # as one process per host core (power of two)
            line["cpu_baseline"]["reference_tu"] = {"value": ref["value"], "unit": "point-updates/s", "cores": 1,
                                                    "kind": "reference", "sample": ref["sample"]}
            try:
                refp = reference_binary_rate(n_side=64, iters=(1, 4), procs=min(cN, 64))
                if refp is not None:
                    line["cpu_baseline"]["reference_tu_parallel"] = {"value": refp["value"], "unit": "point-updates/s",
                                                                     "cores": refp["cores"], "kind": "reference",
                                                                     "sample": refp["sample"]}
            except Exception as e:  # a failing baseline sample must not take the GPU line down
                line["cpu_baseline"]["reference_tu_parallel"] = {"error": str(e)[:200]}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
