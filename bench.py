#!/usr/bin/env python
"""bench.py -- viscosity-solve throughput (DOF*iters/s) and CG-SpMV roofline on B200.

A "step" is one pass of the hot path over one synthetic sphere-drop scene: labelling + assembly + CG to
the configured tolerance (everything HDK_AdaptiveViscosity::solveGasSubclass does after validation,
HDK_AdaptiveViscosity.cpp:233-707).

  value   N * CG iterations / device time of a step with the fields already resident in HBM
  e2e     the same through the host-buffer C-ABI call (avs_solve): pinned host inputs -> H2D -> solve ->
          D2H of the velocity, all inside the timed region
  roofline  the dominant kernel, k_cg_persistent (the whole Jacobi-PCG loop in one cooperative launch): algorithmic
          bytes = SpMV phases * (nnz*(s+4)+(N+1)*4+2*N*s) + x,r phases * 7*N*s + p updates * 4*N*s, divided by the launch
          duration measured with CUDA events on the library's stream during the timed steps; the SpMV phase alone
          (BASELINE.json's "CG SpMV achieved HBM GB/s") is reported beside it, timed inside the kernel and cross-checked
          with CUDA events around stand-alone launches of the same SpMV code
  cpu_baseline  the oracle's Eigen-equivalent Jacobi-PCG (OpenMP, all host cores) on the same matrix,
          a bounded number of iterations

`--impl reference` times the CPU restatement of the reference path (oracle/) on the host cores: the
reference itself needs Houdini + Eigen and cannot be built.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":   # NCCL prints its version banner on STDOUT: keep the JSON line alone
    os.environ["NCCL_DEBUG"] = "WARN"

WORKLOADS = {
    # name: (n, radius_cells, octree_levels, tolerance, description)  -- BASELINE.md section 4
    "c1": dict(n=32, R=10, L=1, tol=1e-3, desc="C1 32^3 uniform (octree depth 1) sphere drop"),
    "c2": dict(n=256, R=82, L=5, tol=1e-6, desc="C2 (DOF-matched) 256^3 R=82 octree depth 5, ~1.1 M DOF"),
    "c2lit": dict(n=128, R=56, L=5, tol=1e-6, desc="C2 (literal) 128^3 R=56 octree depth 5, ~0.5 M DOF"),
    "c3": dict(n=512, R=246, L=7, tol=1e-6, desc="C3 512^3 R=246 octree depth 7, ~10 M DOF"),
    "c3lit": dict(n=256, R=120, L=7, tol=1e-6, desc="C3 (literal) 256^3 R=120 octree depth 7, ~2.4 M DOF"),
    # BASELINE.json configs[3] (~40 M DOF, the config of the >= 6x multi-GPU target): fields generated on the device (scenes_torch),
    # never resident in host memory; fits one B200 (N = 40,270,056, 462 iterations; profiles/r2_scaling.md has 1 / 2 / 4 / 8 GPUs).
    "c4": dict(n=1024, R=492, L=8, tol=1e-6, device_gen=True, desc="C4 1024^3 R=492 octree depth 8, ~40 M DOF (device-generated fields)"),
    # BASELINE.json configs[4]: 10 prescribed-geometry frames of the buckling sheet (scenes.buckling_sheet); frames are
    # independent solves, so they are dealt round-robin to the ranks with no data-path collective
    "c5": dict(frames=10, dx=0.001, L=4, tol=1e-6, dt=1.0 / 120.0,
               desc="C5 buckling sheet (viscousBuckling.hip geometry, dx=0.001, variable viscosity, ground plane), 10 frames"),
}
METRIC = "viscosity_solve_dof_iters_per_s"
UNIT = "DOF*iters/s"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.device = device

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "500"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def make_scene(w, frame=None):
    from adaptiveviscositysolver_b200.scenes import buckling_sheet, sphere_drop
    if "frames" in w:
        return buckling_sheet(w["frames"] - 1 if frame is None else frame, dx=w["dx"])
    return sphere_drop(w["n"], w["R"])


def to_device_scene(scene, torch, dev):
    """Same Scene, every dense field as a CUDA tensor (inputs resident in HBM before the timed region)."""
    from adaptiveviscositysolver_b200.scenes import SampledField, Scene

    def mv(f):
        if f.data is None:
            return f
        return SampledField(torch.from_numpy(f.data).to(dev), f.org, f.dx, f.constant)

    return Scene(scene.res, scene.origin, scene.dx, mv(scene.surface), [mv(v) for v in scene.vel],
                 [mv(v) for v in scene.face_weights], mv(scene.viscosity), mv(scene.density), mv(scene.collision),
                 [mv(v) for v in scene.collision_vel])


def to_pinned_scene(scene, torch):
    from adaptiveviscositysolver_b200.scenes import SampledField, Scene

    def mv(f):
        if f.data is None:
            return f
        t = torch.from_numpy(f.data).pin_memory()
        return SampledField(t, f.org, f.dx, f.constant)

    return Scene(scene.res, scene.origin, scene.dx, mv(scene.surface), [mv(v) for v in scene.vel],
                 [mv(v) for v in scene.face_weights], mv(scene.viscosity), mv(scene.density), mv(scene.collision),
                 [mv(v) for v in scene.collision_vel])


def dense_bytes(scene):
    n = 0
    for f in [scene.surface, scene.viscosity, scene.density, scene.collision, *scene.vel, *scene.face_weights, *scene.collision_vel]:
        if f.data is not None:
            n += int(np.prod(f.data.shape)) * 4
    return n


def compiled_reference_check(orc, cores):
    """The reference's OWN sources (oracle/_ref/libavs_ref.so: /root/reference/Source/*.cpp compiled unchanged, with the threading
    its UT_ThreadedAlgorithm / UTparallelFor calls ask for on `cores` threads; its Eigen CG is serial, as upstream Eigen's is for a
    column-major matrix) timed beside the OpenMP port on ONE WHOLE SOLVE of a bounded workload (C2 literal, ~0.5 M DOF): shows which
    way the `kind: port` arm errs -- the port is the FASTER of the two, so every GPU/CPU ratio formed with it is conservative."""
    try:
        from oracle import avs_ref as ref
        if not ref.available():
            return {"unavailable": "oracle/_ref/libavs_ref.so not in this tree"}
        w2 = WORKLOADS["c2lit"]
        scene = make_scene(w2)
        prm = orc.OracleParams(octree_levels=w2["L"], tolerance=w2["tol"], dt=1.0 / 24.0)
        ref.set_threads(cores)
        t = time.time(); r = ref.RefRun(scene, prm); t_ref = time.time() - t
        n, it_ref, err_ref = int(r.n_face), int(r.iterations), float(r.error)
        del r
        ref.set_threads(1)
        t = time.time(); o = orc.OracleRun(scene, prm); t_port = time.time() - t
        it_port = int(o.iterations)
        return {"workload": w2["desc"], "N": n, "threads": cores,
                "reference": {"whole_solve_s": round(t_ref, 2), "iterations": it_ref, "rel_error": err_ref, "dof_iters_per_s": n * it_ref / t_ref},
                "port": {"whole_solve_s": round(t_port, 2), "iterations": it_port, "dof_iters_per_s": n * it_port / t_port},
                "port_over_reference": round(t_ref / t_port, 2)}
    except Exception as e:  # the cross-check must never take the arm's line with it
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args, w):
    """CPU arm.  The reference's own sources DO compile here against Houdini / Eigen stand-ins (oracle/_ref, used by the parity
    tests), but single-threaded stand-ins at 512^3 would take minutes per solve, so the timed arm is the restated oracle --
    proven bit-identical to the compiled reference on labels, numbering, matrix and rhs (tests/test_reference_pin.py) -- with
    OpenMP on all host cores (`kind: port`).  A step is a BOUNDED SAMPLE OF ONE WHOLE SOLVE: m Eigen-equivalent Jacobi-PCG
    iterations (m >= 100, timed every step) plus the m / iterations share of the assembly time (stages 1-9, timed once), so
    `value` compares like with like against the GPU arm's whole-step `value`; `cg_only` is reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import avs_oracle as orc
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the reference arm runs on rank 0 alone and takes all host cores
    orc.set_num_threads(os.cpu_count() or 1)
    cores = orc.num_threads()
    scene = make_scene(w)
    t0 = time.time()
    run = orc.OracleRun(scene, orc.OracleParams(octree_levels=w["L"], tolerance=w["tol"], dt=w.get("dt", 1.0 / 24.0)), stop_after_stage=9)
    t_asm = time.time() - t0          # stages 1-9 of the restated reference path (weights ... linear system)
    ptr, col, val = run.csr()
    b, x0 = run.rhs(), run.x0()
    n = run.n_face
    orc.cg(ptr, col, val, b, x0, 0.0, 3)                                         # thread pool + first touch
    t = time.time(); orc.cg(ptr, col, val, b, x0, 0.0, 10); per_it = (time.time() - t) / 10
    # one whole CG solve to the workload's tolerance (untimed steps aside): iteration count for the assembly share
    it_full, err_full, t_cg_full = None, None, None
    if per_it * 600 < 120.0:
        t = time.time(); _, it_full, err_full = orc.cg(ptr, col, val, b, x0, w["tol"], 2500); t_cg_full = time.time() - t
    m = int(max(100, min(400, 3.0 / max(per_it, 1e-6))))
    if it_full:
        m = min(m, int(it_full))
    for _ in range(args.warmup):
        orc.cg(ptr, col, val, b, x0, 0.0, m)
    t = time.time()
    for _ in range(args.steps):
        orc.cg(ptr, col, val, b, x0, 0.0, m)
    dt_cg = (time.time() - t) / args.steps
    cg_only = n * m / dt_cg
    asm_share = t_asm * m / it_full if it_full else 0.0
    dt = dt_cg + asm_share
    value = n * m / dt
    whole = None
    if it_full:
        whole = {"assembly_s": round(t_asm, 2), "cg_s": round(t_cg_full, 2), "iterations": int(it_full), "rel_error": float(err_full),
                 "dof_iters_per_s": n * it_full / (t_asm + t_cg_full)}
    sample = (f"{m} Eigen-equivalent Jacobi-PCG iterations per step (OpenMP, {cores} threads, {dt_cg:.2f} s) on the oracle-assembled system "
              f"(N={n}) + {m}/{it_full} of the {t_asm:.1f} s assembly = a bounded sample of one whole solve")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "N": int(n), "nnz": int(run.nnz), "levels": int(run.levels), "tolerance": w["tol"],
                   "oracle_assembly_s": round(t_asm, 2), "cg_only": cg_only, "whole_solve": whole, "sample": sample,
                   "compiled_reference": compiled_reference_check(orc, cores)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_frames(args, w):
    """C5: the frames of the sequence are independent solves (prescribed geometry), dealt round-robin to the ranks;
    every rank runs its own single-GPU context -- no collective on the data path.  A step = all frames once."""
    import torch
    import torch.distributed as dist
    from adaptiveviscositysolver_b200.solver import Params, Solver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = peaks()
    params = Params(octree_levels=w["L"], tolerance=w["tol"], dt=w["dt"], single_precision=args.fp32)
    from adaptiveviscositysolver_b200.dist_plan import deal_frames
    mine = deal_frames(w["frames"], rank, world)
    scenes = [make_scene(w, f) for f in mine]
    dscenes = [to_device_scene(sc, torch, dev) for sc in scenes]
    douts = [[v.data.clone() for v in ds.vel] for ds in dscenes]
    solver = Solver(device=local, time_spmv=not args.no_spmv_events)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(scs, outs):
        infos = [solver.solve(sc, params, o) for sc, o in zip(scs, outs)]
        return infos

    for _ in range(args.warmup):
        one_pass(dscenes, douts)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    dev_ms = spmv_ms = 0.0
    spmv_n = launches = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        infos = one_pass(dscenes, douts)
        dev_ms += sum(i.stage_ms["total"] for i in infos)
        spmv_ms += sum(i.spmv_ms for i in infos)
        spmv_n += sum(i.spmv_launches for i in infos)
        launches += sum(i.kernel_launches for i in infos)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clk = clocks.stop()
    dev_ms /= args.steps
    work = sum(i.octree_dofs * i.iterations for i in infos)          # DOF*iters of this rank's frames, one pass
    s = 4 if args.fp32 else 8
    last = infos[-1] if infos else None
    alg_bytes = (last.nnz * (s + 4) + (last.octree_dofs + 1) * 4 + 2 * last.octree_dofs * s) if last else 0

    pscenes = [to_pinned_scene(sc, torch) for sc in scenes]
    pouts = [[torch.from_numpy(v.data.copy()).pin_memory() for v in sc.vel] for sc in scenes]
    one_pass(pscenes, pouts)
    barrier()
    t0 = time.perf_counter()
    einfos = one_pass(pscenes, pouts)
    torch.cuda.synchronize()
    e_ms = (time.perf_counter() - t0) * 1e3
    # face weights are pinned here and read in place, one value per level-0 row, unless AVS_FACEW_UPLOAD=bulk (see main()); rows of
    # level 0 are not read back per frame here, so the gather is counted at its upper bound of one 32-byte sector per ROW
    if os.environ.get("AVS_FACEW_UPLOAD") == "bulk":
        h2d = sum(dense_bytes(sc) for sc in scenes)
    else:
        h2d = sum(dense_bytes(sc) - sum(int(np.prod(f.data.shape)) * 4 for f in sc.face_weights if f.data is not None) for sc in scenes)
        h2d += sum(32 * int(i.octree_dofs) for i in einfos)
    d2h = sum(sum(int(np.prod(v.data.shape)) * 4 for v in sc.vel) for sc in scenes)

    tot = torch.tensor([float(work), float(h2d), float(d2h), float(launches)], device=dev, dtype=torch.float64)
    mx = torch.tensor([dev_ms, e_ms, wall_ms], device=dev, dtype=torch.float64)
    box = [None] * world
    rec = {"rank": rank, "frames": mine, "N": [int(i.octree_dofs) for i in infos], "iterations": [int(i.iterations) for i in infos],
           "levels": [int(i.levels) for i in infos], "ms": [round(i.stage_ms["total"], 3) for i in infos]}
    if world > 1:
        dist.all_reduce(tot)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_gather_object(box, rec)
    else:
        box = [rec]
    work_all, h2d_all, d2h_all, launches_all = (float(v) for v in tot.tolist())
    dev_ms_max, e_ms_max, wall_ms_max = (float(v) for v in mx.tolist())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and last is not None:
        from oracle import avs_oracle as orc
        ptr, col, val, rhs, x0 = solver.system()
        cores = orc.num_threads()
        orc.cg(ptr, col, val, rhs, x0, 0.0, 5)
        t = time.time(); orc.cg(ptr, col, val, rhs, x0, 0.0, 20); per_it = (time.time() - t) / 20
        m = int(max(20, min(20000, 12.0 / max(per_it, 1e-7))))
        t = time.time(); orc.cg(ptr, col, val, rhs, x0, 0.0, m); dt = time.time() - t
        cpu = {"value": last.octree_dofs * m / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{m} Eigen-equivalent Jacobi-PCG iterations of the oracle (OpenMP, {cores} threads) on the GPU-assembled "
                         f"matrix of frame {mine[-1]} (N={last.octree_dofs}, nnz={last.nnz}); {dt:.1f} s"}
    if rank == 0:
        spmv_avg_ms = spmv_ms / max(spmv_n, 1)
        achieved = alg_bytes / (spmv_avg_ms * 1e-3) / 1e9 if spmv_avg_ms > 0 else None
        line = {
            "metric": METRIC, "value": work_all / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "frames": w["frames"], "sharding": "frames round-robin over ranks, no collective",
                       "dt": w["dt"], "tolerance": w["tol"], "octree_levels": w["L"], "per_rank": box, "wall_ms_per_step": wall_ms_max,
                       "l2": "per-frame matrices (~60 MB) fit the 126 MB L2: the SpMV figure below is NOT an HBM roofline number"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": None,
                         "kernel": "SpMV phase of k_cg_persistent (in-kernel timer), last frame of rank 0; L2-resident at this size",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": spmv_avg_ms},
            "cpu_baseline": cpu,
            "e2e": {"value": work_all / (e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_all),
                    "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e_ms_max},
            "gpu_launches": int(launches_all),
            "clocks": clk,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--fp32", action="store_true", help="USESINGLEPRECISION solve (C3's ncu capture config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather-output", action="store_true", help="N > 1: all-gather the output z-slabs so that every rank holds the whole velocity field")
    ap.add_argument("--no-spmv-events", action="store_true", help="do not bracket SpMV launches with CUDA events (roofline from back-to-back timing)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        if w.get("device_gen"):
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the CPU restatement at 1024^3 exceeds the time and memory budget of a bench run"}), flush=True)
            return
        run_reference(args, w)
        return
    if "frames" in w:
        run_frames(args, w)
        return

    import torch
    import torch.distributed as dist
    from adaptiveviscositysolver_b200.solver import Params, Solver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    peak, peak_src = peaks()
    device_gen = bool(w.get("device_gen"))
    scene = None if device_gen else make_scene(w)
    params = Params(octree_levels=w["L"], tolerance=w["tol"], single_precision=args.fp32)
    uid = None
    if world > 1:
        from adaptiveviscositysolver_b200.solver import nccl_unique_id
        box = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    # N > 1: every rank keeps (device-resident steps) / downloads (e2e steps) the z-slab of the output it computed -- the row-partitioned
    # solve leaves a slab-partitioned velocity field; --gather-output restores "whole field on every rank" (NVLink all-gather of 1.6 GB)
    solver = Solver(device=local, rank=rank, nranks=world, time_spmv=not args.no_spmv_events, nccl_unique_id=uid,
                    distributed_output=(world > 1 and not args.gather_output))
    if device_gen:
        from adaptiveviscositysolver_b200.scenes_torch import sphere_drop_device
        dscene = sphere_drop_device(w["n"], w["R"], dev)
        torch.cuda.synchronize()
    else:
        dscene = to_device_scene(scene, torch, dev)
    dout = [v.data.clone() for v in dscene.vel]
    stream = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps -------------------------------------------------------------------
    for _ in range(args.warmup):
        info = solver.solve(dscene, params, dout)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    step_ms, spmv_ms, spmv_n, launches = [], 0.0, 0, 0
    cg_ms, cg_launches, cg_iters, xr_ms, pu_ms = 0.0, 0, 0, 0.0, 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        info = solver.solve(dscene, params, dout)       # returns after the stream has drained (stage events)
        step_ms.append(info.stage_ms["total"])
        spmv_ms += info.spmv_ms
        spmv_n += info.spmv_launches
        launches += info.kernel_launches
        cg_ms += info.cg_kernel_ms
        cg_launches += info.cg_kernel_launches
        cg_iters += info.iterations
        xr_ms += info.cg_update_xr_ms
        pu_ms += info.cg_update_p_ms
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clk = clocks.stop()
    dev_ms = float(np.mean(step_ms))
    t = torch.tensor([dev_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    N, iters, nnz = info.octree_dofs, info.iterations, info.nnz
    n_local = info.local_rows
    if world > 1:
        tn = torch.tensor([nnz], device=dev, dtype=torch.int64)
        dist.all_reduce(tn)
        nnz_total = int(tn.item())
    else:
        nnz_total = nnz
    value = N * iters / (dev_ms * 1e-3)

    s = 4 if args.fp32 else 8
    alg_bytes = nnz * (s + 4) + (n_local + 1) * 4 + 2 * n_local * s   # this rank's SpMV
    per_rank = None
    if world > 1:
        box = [None] * world
        dist.all_gather_object(box, {"rank": rank, "rows": int(n_local), "nnz": int(nnz), "spmv_ms": spmv_ms / max(spmv_n, 1),
                                     "halo": int(info.halo_columns), "solve_ms": info.stage_ms["solve"]})
        per_rank = box
    # the same kernel launched back to back (matrix >> L2), for reference / when events are off
    iso_ms, _ = solver.time_spmv_resident(20)
    spmv_avg_ms = spmv_ms / max(spmv_n, 1) if spmv_ms > 0 else iso_ms
    achieved = alg_bytes / (spmv_avg_ms * 1e-3) / 1e9

    persistent = cg_launches > 0
    # DRAM traffic of one launch of the dominant kernel from the committed ncu --set full capture (same matrix only)
    traffic = None
    tp = ROOT / "profiles" / (("r2_cg_persistent_fp32_traffic.json" if args.fp32 else "r1_cg_persistent_traffic.json") if persistent else "r1_spmv_traffic.json")
    if tp.exists() and world == 1 and (persistent or not args.fp32):
        tj = json.loads(tp.read_text())
        if tj.get("N") == int(N) and tj.get("nnz") == int(nnz):
            traffic = tj["traffic_bytes_per_launch"]
    if persistent:
        # dominant kernel = k_cg_persistent: one cooperative launch runs the whole Jacobi-PCG loop.  Algorithmic bytes of a
        # launch: every SpMV phase B_spmv, every x,r phase 7 N s, every p update 4 N s.
        cg_alg = spmv_n * (alg_bytes + 7 * n_local * s) + cg_iters * 4 * n_local * s
        cg_achieved = cg_alg / (cg_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": cg_achieved, "peak": peak, "unit": "GB/s", "frac": cg_achieved / peak, "traffic": traffic,
                    "kernel": "k_cg_persistent (whole Jacobi-PCG loop in one cooperative launch: SpMV + p.Ap | x,r update + r.r, r.z | "
                              "p update; 3 grid barriers per iteration)",
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": cg_alg / max(cg_launches, 1),
                    "avg_launch_ms": cg_ms / max(cg_launches, 1), "launches_timed": int(cg_launches),
                    "timer": "CUDA events on the library's stream around every cooperative launch of the timed steps",
                    "iterations_per_launch": cg_iters / max(cg_launches, 1),
                    "spmv_phase": {"achieved": achieved, "frac": achieved / peak, "unit": "GB/s", "algorithmic_bytes": alg_bytes,
                                   "avg_ms": spmv_avg_ms, "phases_timed": int(spmv_n),
                                   "timer": "inside the kernel: %globaltimer of CTA 0, grid barrier to grid barrier"},
                    "spmv_standalone": {"achieved": alg_bytes / (iso_ms * 1e-3) / 1e9, "frac": alg_bytes / (iso_ms * 1e-3) / 1e9 / peak,
                                        "avg_ms": iso_ms, "timer": "CUDA events around 20 back-to-back launches of k_spmv_sjds (same slice code)"},
                    "xr_phase_ms_per_iter": xr_ms / max(spmv_n, 1), "p_phase_ms_per_iter": pu_ms / max(cg_iters, 1)}
    else:
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "k_spmv_sjds (CG SpMV + fused p.Ap), AVS_CG_MODE=launch", "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": spmv_avg_ms, "launches_timed": int(spmv_n),
                    "back_to_back_ms": iso_ms, "cg_update_xr_ms_per_iter": info.cg_update_xr_ms / max(info.spmv_launches - 1, 1),
                    "cg_update_p_ms_per_iter": info.cg_update_p_ms / max(info.spmv_launches - 1, 1),
                    "cg_iteration_frac": (iters * (alg_bytes + 11 * n_local * s)) / (info.stage_ms["solve"] * 1e-3) / 1e9 / peak}

    # ---- end to end through the host-buffer call --------------------------------------------------
    e2e = None
    if device_gen and not args.no_e2e:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
               "skipped": "device-generated workload: its fields (30 GB) are never materialised in host memory"}
    elif not args.no_e2e:
        pscene = to_pinned_scene(scene, torch)
        pout = [torch.from_numpy(v.data.copy()).pin_memory() for v in scene.vel]
        # Host -> device traffic of one step: every dense field is copied whole EXCEPT the three face-weight arrays -- they are
        # pinned here, so the library reads them through mapped host pointers, one value per level-0 row of the rank
        # (k_gather_face_weights): counted as one 32-byte PCIe sector per such row, an upper bound (neighbouring rows share sectors).
        # AVS_FACEW_UPLOAD=bulk restores the whole-array upload (then all 7 dense fields are counted).
        bulk_facew = os.environ.get("AVS_FACEW_UPLOAD") == "bulk"
        facew_bytes = sum(int(np.prod(f.data.shape)) * 4 for f in scene.face_weights if f.data is not None)
        level0_rows = int((solver.keys()[:, 0] == 0).sum()) if not bulk_facew else 0
        h2d = dense_bytes(scene) if bulk_facew else dense_bytes(scene) - facew_bytes
        h2d_gather = 0 if bulk_facew else 32 * level0_rows   # whole job (each rank gathers only its own rows)
        d2h = sum(int(np.prod(v.data.shape)) * 4 for v in scene.vel)
        solver.solve(pscene, params, pout)
        barrier()
        t0 = time.perf_counter()
        k = max(1, min(args.steps, 2))
        for _ in range(k):
            einfo = solver.solve(pscene, params, pout)
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - t0) * 1e3 / k
        if world > 1:
            te = torch.tensor([e_ms], device=dev)
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e_ms = float(te.item())
        if world > 1 and not args.gather_output:   # every rank downloads only the z-slab it computed: count what this rank moved
            plane = [int(np.prod(v.data.shape[1:])) for v in scene.vel]
            d2h = sum(plane[a] * max(0, solver.output_slab(a)[1] - solver.output_slab(a)[0]) * 4 for a in range(3))
        if world > 1:   # whole-job bytes (every rank uploads the full surface / velocity fields: labelling is replicated)
            tb = torch.tensor([float(h2d), float(d2h)], device=dev, dtype=torch.float64)
            dist.all_reduce(tb)
            h2d, d2h = (float(v) for v in tb.tolist())
        h2d += h2d_gather
        e2e = {"value": N * einfo.iterations / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e_ms,
               "face_weights": ("whole arrays uploaded" if bulk_facew else
                                f"read in place from pinned host memory: one value per level-0 row ({level0_rows} rows, counted at 32 B each)"),
               "stage_ms": {k2: round(v, 3) for k2, v in einfo.stage_ms.items()}}
        del pscene, pout

    # ---- CPU baseline: oracle CG on the same matrix ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not device_gen:
        from oracle import avs_oracle as orc
        ptr, col, val, rhs, x0 = solver.system()
        cores = orc.num_threads()
        orc.cg(ptr, col, val, rhs, x0, 0.0, 2)                                       # first touch / thread pool warm-up
        t = time.time(); orc.cg(ptr, col, val, rhs, x0, 0.0, 4); per_it = (time.time() - t) / 4
        m = int(max(3, min(5000, 15.0 / max(per_it, 1e-6))))
        t = time.time(); orc.cg(ptr, col, val, rhs, x0, 0.0, m); dt = time.time() - t
        cpu = {"value": N * m / dt, "unit": UNIT, "cores": cores, "kind": "port", "scope": "cg_only",
               "sample": f"{m} Eigen-equivalent Jacobi-PCG iterations of the oracle (OpenMP, {cores} threads) on the GPU-assembled "
                         f"matrix of this workload (N={N}, nnz={nnz}); {dt:.1f} s",
               "spmv_gbs": None}
        del ptr, col, val

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "N": int(N), "nnz": int(nnz_total), "rows_per_rank": int(n_local), "dist_mode": {0: "single", 1: "nccl", 2: "peer-memory"}[info.dist_mode], "output": ("z-slab per rank" if (world > 1 and not args.gather_output) else "whole field"), "halo_columns": int(info.halo_columns), "levels": int(info.levels), "iterations": int(iters),
                       "rel_error": info.error, "tolerance": w["tol"], "regular_dofs": int(info.regular_dofs),
                       # `value` is the whole step (labelling + assembly + CG + write-back); the CG stage alone, for comparison
                       # with the reference arm's `config.cg_only` and this arm's `cpu_baseline` (both CG only):
                       "cg_only": N * iters / (info.stage_ms["solve"] * 1e-3),
                       "l2": "matrix+vectors per SpMV = %.0f MB vs 126 MB L2 (inputs larger than L2, no flush)" % (alg_bytes / 1e6),
                       "wall_ms_per_step": wall_ms, "per_rank": per_rank,
                       "stage_ms": {k2: round(v, 3) for k2, v in info.stage_ms.items()}},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
