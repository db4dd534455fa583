#!/usr/bin/env python
"""bench.py -- body-steps/sec of the Edyn per-step hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b2d|reference] [--workload NAME]

A "step" is one fixed simulation step (broadphase -> narrowphase -> islands -> solve -> integrate) over the
whole scene.  N = 1 runs BASELINE.json's 262 144-body mixed pile (config 4, the one the >=10x / >=40 % targets
are quoted on); N > 1 is weak scaling: every rank owns an independent island group of the same size (islands
shard with no data-path collective) and `value` is total dynamic bodies x steps / max-over-ranks time.

Keys beyond the base contract:
  roofline      dominant kernel (k_solve): algorithmic bytes (388 B per contact point per velocity iteration,
                544 B per hinge per iteration, + 0.75 pass for the warm start) / mean CUDA-event
                duration of that kernel over the timed steps, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  the oracle port timed on this box's host cores on a bounded sample (rank 0, N = 1 only).
  e2e           same metric through the C ABI with HOST buffers: every step uploads the body state from pinned
                host memory (b2d_upload_state), steps, and downloads it again (b2d_download_state).
--impl reference times the reference's CPU path (the oracle port: the reference stepper itself needs EnTT,
which this image lacks -- DESIGN.md section 6) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SETTLE_STEPS = 150            # untimed: lets the dropped pile come to rest so contact counts are stationary
BYTES_PER_POINT_ITER = 388
BYTES_PER_HINGE_ITER = 544
BYTES_PER_BODY_INTEGRATE = 276


def make_scene(name, scale=1.0):
    import edyn_b200 as E
    if name == "mixed_262144":
        side = max(4, int(round(64 * scale ** (1 / 3))))
        return E.scenes.mixed_pile(side)
    if name == "boxes_4096":
        return E.scenes.boxes_on_plane(max(2, int(round(16 * scale ** (1 / 3)))))
    if name == "spheres_65536":
        return E.scenes.spheres_in_box(max(2, int(round(64 * scale ** 0.5))), 16, max(2, int(round(64 * scale ** 0.5))))
    if name == "chains_1048576":
        k = max(2, int(round(512 * scale ** 0.5)))
        return E.scenes.hinge_chains(k, k)
    raise SystemExit(f"unknown workload {name}")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.t = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.15)

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def oracle_from_device(scene, w, threads):
    """Oracle world holding the device's current (settled) state, contacts included."""
    from oracle import oracle as O
    o = O.OracleWorld(vel_iters=scene["settings"]["velocity_iterations"], pos_iters=scene["settings"]["position_iterations"], threads=threads)
    o.add_bodies(scene["bodies"])
    if scene["hinges"]:
        h = scene["hinges"]
        o.add_hinges(h["a"], h["b"], h["pivot_a"], h["pivot_b"], h["axis_a"], h["axis_b"])
    if scene["exclusions"] is not None:
        o.add_exclusions(*scene["exclusions"])
    st = w.download_state(aabb=False)
    o.set_state(st["pos"], st["orn"], st["linvel"], st["angvel"])
    c = w.contacts()
    o.set_contacts(c["pairs"], c["num"], c["pts"], c["att"], c["lifetime"])
    return o


def run_reference(args):
    """Reference arm: the CPU path (oracle port, all host threads) on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    scale = args.ref_scale
    scene = make_scene(args.workload, scale)
    o = O.OracleWorld(vel_iters=scene["settings"]["velocity_iterations"], pos_iters=scene["settings"]["position_iterations"], threads=cores)
    o.add_bodies(scene["bodies"])
    if scene["hinges"]:
        h = scene["hinges"]
        o.add_hinges(h["a"], h["b"], h["pivot_a"], h["pivot_b"], h["axis_a"], h["axis_b"])
    if scene["exclusions"] is not None:
        o.add_exclusions(*scene["exclusions"])
    o.step(args.ref_settle)                      # untimed settle, like the device arm
    for _ in range(args.warmup):
        o.step(1)
    t0 = time.perf_counter()
    o.step(args.steps)
    dt = time.perf_counter() - t0
    value = scene["dynamic"] * args.steps / dt
    sample = f"{scene['name']} ({scene['dynamic']} dynamic bodies, same generator and settings as the device arm, " \
             f"{args.ref_settle} settle steps), {args.steps} timed steps"
    line = {"impl": "reference", "metric": "body-steps/sec", "value": value, "unit": "body-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_device(args):
    import torch
    import edyn_b200 as E

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world_size > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    scene = make_scene(args.workload, args.scale)
    n_dyn = scene["dynamic"]
    if world_size > 1:
        # weak scaling: rank r owns an island group of the same size, placed side by side along x
        b = scene["bodies"]
        p = b["pos"][b["kind"] == 0]
        stride_x = float(p[:, 0].max() - p[:, 0].min()) + 20.0
        b["pos"] = b["pos"].copy()
        b["pos"][:, 0] += np.float32(rank * stride_x)
    w = E.scenes.build_world(scene, device=local_rank)
    w.step(SETTLE_STEPS)
    w.sync()
    stream = torch.cuda.ExternalStream(w.stream, device=local_rank)
    # cross-GPU AABB exchange: per step, bounds of the owned islands reduced on the device, all-gathered
    # over NCCL, overlap-tested on the device; the hit counter is read once after the timed region
    bounds = torch.zeros(6, dtype=torch.float32, device="cuda")
    gathered = torch.zeros(world_size, 6, dtype=torch.float32, device="cuda")
    hits = torch.zeros((), dtype=torch.int64, device="cuda")
    margin = 0.026

    def exchange():
        if dist is None:
            return
        w.device_bounds(bounds.data_ptr())
        torch.cuda.current_stream().wait_stream(stream)
        dist.all_gather_into_tensor(gathered, bounds)
        lo, hi = gathered[:, None, :3], gathered[:, None, 3:]
        ov = ((lo - margin <= hi.transpose(0, 1)) & (hi + margin >= lo.transpose(0, 1))).all(dim=2)
        hits.add_(ov.sum() - world_size)               # minus the diagonal
        stream.wait_stream(torch.cuda.current_stream())
    st0 = w.stats()
    if st0["error_flags"]:
        raise SystemExit(f"device error flags {st0['error_flags']} after settling")

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---------------- device-resident throughput ("value")
    for _ in range(args.warmup):
        w.step(1)
        exchange()
    barrier()
    w.sync()
    w.reset_timers()
    launches0 = w.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        w.step(1)
        exchange()
    e1.record(stream)
    w.sync()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    st = w.stats()
    launches = st["kernel_launches"] - launches0 - 1        # minus the stats kernel itself
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_bodies = n_dyn * world_size
    value = total_bodies * args.steps / (ms_max * 1e-3)

    # ---------------- end to end through the C ABI with host buffers
    n_all = w.num_bodies
    pinned = {k: torch.empty((n_all, d), dtype=torch.float32).pin_memory() for k, d in (("pos", 3), ("orn", 4), ("linvel", 3), ("angvel", 3))}
    host = {k: v.numpy() for k, v in pinned.items()}
    w.download_state(aabb=False, out=host)
    e2e_steps = max(3, min(args.steps, 50))
    for _ in range(3):
        w.upload_state(host["pos"], host["orn"], host["linvel"], host["angvel"])
        w.step(1)
        w.download_state(aabb=False, out=host)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
        w.upload_state(host["pos"], host["orn"], host["linvel"], host["angvel"])
        w.step(1)
        w.download_state(aabb=False, out=host)         # blocks until the step's results are on the host
    e1.record(stream)
    w.sync()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_bodies * e2e_steps / (float(t.item()) * 1e-3)
    bytes_state = n_all * 13 * 4

    # ---------------- roofline of the dominant kernel
    peak, peak_src = peaks()
    iters = scene["settings"]["velocity_iterations"]
    algo_bytes = (iters + 0.75) * (BYTES_PER_POINT_ITER * st["contact_points"] + BYTES_PER_HINGE_ITER * st["hinges"])
    achieved = algo_bytes / (st["solve_ms"] * 1e-3) / 1e9 if st["solve_ms"] > 0 else 0.0
    integ_gbs = BYTES_PER_BODY_INTEGRATE * n_dyn / (st["integrate_ms"] * 1e-3) / 1e9 if st["integrate_ms"] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath) and args.scale == 1.0:
        tj = json.load(open(tpath))
        if tj.get("workload") == args.workload:
            traffic = tj["k_solve"]["dram_bytes_read"] + tj["k_solve"]["dram_bytes_write"]     # ncu --set full capture, per launch
    roofline = {"bound": "hbm", "kernel": "k_solve_df", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_ms": st["solve_ms"], "kernel_share_of_step": st["solve_ms"] / (ms / args.steps),
                "integrate": {"kernel": "k_integrate", "achieved": integ_gbs, "frac": integ_gbs / peak, "kernel_ms": st["integrate_ms"]}}

    # ---------------- CPU baseline on a bounded sample (rank 0, N = 1)
    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        o = oracle_from_device(scene, w, cores)
        o.step(1)                                   # untimed: first touch
        nsteps, t0 = 0, time.perf_counter()
        while nsteps < 2 or (time.perf_counter() - t0 < args.cpu_seconds and nsteps < 50):
            o.step(1)
            nsteps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": n_dyn * nsteps / dt, "unit": "body-steps/s", "cores": cores, "kind": "port",
               "sample": f"{nsteps} steps of the same settled {scene['name']} state downloaded from the device "
                         f"(oracle/ CPU restatement of stepper_sequential, narrowphase + per-island solve on {cores} threads)"}

    if rank == 0:
        line = {"metric": "body-steps/sec", "value": value, "unit": "body-steps/s", "n_gpus": world_size, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload if args.scale == 1.0 else f"{args.workload} x{args.scale}",
                           "scene": scene["name"], "dynamic_bodies_per_gpu": n_dyn,
                           "velocity_iterations": iters, "position_iterations": scene["settings"]["position_iterations"],
                           "settle_steps": SETTLE_STEPS, "manifolds": st["manifolds"], "contact_points": st["contact_points"],
                           "hinges": st["hinges"], "contact_colors": st["contact_colors"], "islands": st["islands"],
                           "parallelism": f"islands sharded over {world_size} GPU(s); per-step NCCL all-gather of island-group bounds "
                                          f"(24 B/rank), {int(hits.item())} cross-rank overlaps seen",
                           "l2": "working set (rows + bodies) exceeds the 126 MB L2; no explicit flush"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": "body-steps/s", "h2d_bytes_per_step": bytes_state, "d2h_bytes_per_step": bytes_state,
                        "steps": e2e_steps},
                "roofline": roofline}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b2d", choices=["b2d", "reference"])
    ap.add_argument("--workload", default="mixed_262144")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (development only; invalid as a bench value)")
    ap.add_argument("--ref-scale", type=float, default=1.0 / 16, help="reference arm: fraction of the workload simulated on the CPU")
    ap.add_argument("--ref-settle", type=int, default=SETTLE_STEPS)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_device(args)


if __name__ == "__main__":
    main()
