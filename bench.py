#!/usr/bin/env python
"""bench.py — rays/s of the volumetric-rendering hot path (BASELINE.json: 800x800 rays, 192 samples/ray).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--field hull|fog]

A "step" is one pass of the hot path over one batch of synthetic rays:
  N = 1   one 800x800 TriPlane frame (640 000 rays, 192 samples/ray, gauge on, alpha mask) = BASELINE configs[1];
  N > 1   a batch of N such frames whose rays are dealt to the ranks in interleaved 2000-ray blocks (each rank renders
          640 000 rays per step = weak scaling) followed by ONE NCCL all-gather of the rendered frames (configs[4]).
Inputs rotate over 16 different camera poses (16 x 15.4 MB of rays per rank > 126 MB L2), so every step reads its rays
from HBM.  The field (25 MB of packed planes) is model state and stays wherever the hardware keeps it.

The JSON line (rank 0) follows the driver contract; see DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

N_POSES = 16
BLOCK = 2000
H = W = 800
S = 192
RAYS_PER_FRAME = H * W
HBM_BYTES_PER_RAY = 40            # 24 B ray in + 16 B (rgb, depth) out: SURVEY.md §8(d)
MLP_FLOPS_PER_COLOUR_SAMPLE = 70400   # reference arithmetic incl. the 144x144 basis: SURVEY.md §8(a) row a9
MLP_FLOPS_PER_DENSITY_SAMPLE = 96


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--field", default="hull", choices=["hull", "fog"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-regime (tensor-bound) side measurement")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    tensor_burst=float(d["bf16_tflops"]), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------
def shard_of_batch(synth, pose0: int, world: int, rank: int):
    """Rays rank owns of the batch made of frames pose0 .. pose0+world-1 (interleaved BLOCK-ray blocks)."""
    if world == 1:
        return synth.config_rays("C2", pose0)
    g = torch.arange(world * RAYS_PER_FRAME)
    mine = g[((g // BLOCK) % world) == rank]
    frames = {}
    out = torch.empty((mine.numel(), 6))
    fidx = mine // RAYS_PER_FRAME
    for f in fidx.unique().tolist():
        frames[f] = synth.config_rays("C2", pose0 + f)
        sel = fidx == f
        out[sel] = frames[f][mine[sel] - f * RAYS_PER_FRAME]
    return out.contiguous()


def cpu_port_rate(spec, R, rays, budget_s: float, threads: int):
    """The oracle port (reference algorithm, PyTorch CPU fp32, 4096-ray chunks as TriPlane/main.py:60-71) on a
    bounded sample: chunks are rendered until `budget_s` of CPU time is used."""
    torch.set_num_threads(threads)
    R.render(spec, rays[:4096], N_samples=S)               # warm-up
    done, t0 = 0, time.perf_counter()
    while done < rays.shape[0] and time.perf_counter() - t0 < budget_s:
        R.render(spec, rays[done:done + 4096], N_samples=S)
        done += min(4096, rays.shape[0] - done)
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def run_reference(args, rank):
    """--impl reference: the reference's own algorithm on the host cores.  /root/reference does not exist on the GPU
    box and the reference is Python (nothing to compile into oracle/_ref), so this times the oracle port."""
    if rank != 0:
        return
    from oracle import cases as K
    from oracle import restate_field as R
    case = K.Case("bench", kind=args.field, config="C2", n_samples=S)
    state, kw, occ, _ = K.build_inputs(case)
    spec = R.spec_from_state("triplane", state, alpha_volume=occ, gauge_on=True, **kw)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n_sample = 16384                                        # rays per step: 4 chunks of 4096, strided over the frame
    def step(i):
        rays = K.synth.config_rays("C2", i % N_POSES)[:: RAYS_PER_FRAME // n_sample][:n_sample].contiguous()
        t0 = time.perf_counter()
        R.render(spec, rays, N_samples=S)
        return time.perf_counter() - t0
    for i in range(args.warmup):
        step(i)
    total = sum(step(i) for i in range(args.steps))
    v = n_sample * args.steps / total
    print(json.dumps({
        "impl": "reference", "metric": "rays/sec (800x800, 192 samples/ray)", "value": v, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"TriPlane {args.field} field 800x800 rays x 192 samples, gauge on, alpha mask "
                               f"(BASELINE configs[1]); each step = {n_sample}-ray strided sample of one frame"},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": f"{n_sample} rays/step strided over the frame, 4096-ray chunks"},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    import torch.distributed as dist
    import ngf_b200
    from ngf_b200 import synth
    from ngf_b200.render import frame_allgather

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- field (model state) and inputs
    kw = synth.field_kwargs("C2")
    state = synth.field_state("triplane", args.field)
    occ = synth.occupancy_volume(args.field)
    field = ngf_b200.TriPlane(kw["aabb"], kw["gridSize"], dev, near_far=kw["near_far"], step_ratio=kw["step_ratio"],
                              distance_scale=kw["distance_scale"], rayMarch_weight_thres=kw["rayMarch_weight_thres"],
                              gauge_start=0)
    synth.load_into(field, state, occ, ngf_b200.AlphaGridMask)
    assert field.nSamples == S, field.nSamples
    host = [shard_of_batch(synth, p * world, world, rank).pin_memory() for p in range(N_POSES)]
    n_local = host[0].shape[0]
    dev_rays = [h.to(dev) for h in host]
    img_w = W if world == 1 else 0
    n_batch = world * RAYS_PER_FRAME

    def step_device(rays):
        out = field(rays, white_bg=True, N_samples=S, iteration=30001, image_width=img_w)
        if world == 1:
            return out["rgb_map"], out["depth_map"]
        local_res = torch.cat([out["rgb_map"], out["depth_map"][:, None]], 1)
        return frame_allgather(local_res, n_batch, BLOCK), None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timed region (value)
    for i in range(max(args.warmup, 3)):
        step_device(dev_rays[i % N_POSES])
    barrier()
    stats = field.last_stats()
    field.kernel_timing(args.steps)
    l0 = ngf_b200._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for i in range(args.steps):
            step_device(dev_rays[i % N_POSES])
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ngf_b200._lib.launch_count() - l0
    n_k, march_ms, colour_ms = field.kernel_timing_read()
    field.kernel_timing(0)
    value = n_batch * args.steps / (ms * 1e-3)

    # ---- end-to-end timed region: pinned host rays -> H2D -> render (+ all-gather) -> D2H of the results
    if world == 1:
        rgb_h = torch.empty((n_local, 3)).pin_memory()
        dep_h = torch.empty((n_local,)).pin_memory()
        def step_e2e(i):
            field.render_host(host[i % N_POSES], rgb_h, dep_h, white_bg=True, N_samples=S, image_width=W, iteration=30001)
        d2h = n_local * 16
    else:
        stage = torch.empty_like(dev_rays[0])
        res_h = torch.empty((n_local, 4)).pin_memory()
        def step_e2e(i):
            stage.copy_(host[i % N_POSES], non_blocking=True)
            out = field(stage, white_bg=True, N_samples=S, iteration=30001)
            loc = torch.cat([out["rgb_map"], out["depth_map"][:, None]], 1)
            frame_allgather(loc, n_batch, BLOCK)
            res_h.copy_(loc, non_blocking=True)
            torch.cuda.synchronize()
        d2h = n_local * 16
    for i in range(3):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_batch * args.steps / e2e_s

    # ---- roofline of the dominant kernel, from CUDA events around each kernel of the pair (measured live above)
    pk = peaks()
    n_k = max(n_k, 1)
    march_s, colour_s = march_ms / n_k * 1e-3, colour_ms / n_k * 1e-3
    step_s = ms / args.steps * 1e-3
    nV = stats["samples_density"] / n_local
    nA = stats["samples_colour"] / n_local
    alg_bytes = n_local * HBM_BYTES_PER_RAY
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"march_kernel_{args.field}_dram_bytes_per_launch")
    achieved = alg_bytes / march_s / 1e9
    mlp_flops = n_local * nA * MLP_FLOPS_PER_COLOUR_SAMPLE
    roofline = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
                "traffic": traffic, "kernel": "ngf_march_kernel<TriPlane>", "kernel_ms": march_s * 1e3,
                "kernel_share_of_step": march_s / step_s,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": pk["src"],
                "note": "sparse regime: 40 B/ray of compulsory HBM traffic; the march is issue / L2-gather bound (DESIGN.md)",
                "colour_kernel": {"kernel": "ngf_colour_kernel<TriPlane,tcgen05>", "kernel_ms": colour_s * 1e3,
                                  "kernel_share_of_step": colour_s / step_s, "bound": "tensor",
                                  "achieved": mlp_flops / colour_s / 1e12 if colour_s > 0 else None,
                                  "peak": pk["tensor"], "unit": "TFLOP/s",
                                  "frac": mlp_flops / colour_s / 1e12 / pk["tensor"] if colour_s > 0 else None},
                "density_samples_per_ray": nV, "colour_samples_per_ray": nA,
                "mlp_flop_roofline_frac_of_step": n_local * (nV * MLP_FLOPS_PER_DENSITY_SAMPLE + nA * MLP_FLOPS_PER_COLOUR_SAMPLE)
                                                  / step_s / 1e12 / pk["tensor"]}

    line = {
        "metric": "rays/sec (800x800, 192 samples/ray)", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp16 tensor-core operands, fp32 accumulate)",
        "data": "synthetic",
        "config": {"workload": f"TriPlane {args.field} field, 800x800 rays x 192 samples/ray, gauge on, 256^3 alpha mask "
                               f"(BASELINE configs[1]" + (")" if world == 1 else f"; configs[4]: {world} frames/step ray-sharded + NCCL all-gather)"),
                   "rays_per_step": n_batch, "l2": f"inputs rotate over {N_POSES} poses ({N_POSES * n_local * 24 / 1e6:.0f} MB of rays per rank > 126 MB L2)",
                   "parallelism": "single GPU" if world == 1 else f"ray-sharded dp{world}, {BLOCK}-ray interleaved blocks"},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": n_local * 24, "d2h_bytes_per_step": d2h,
                "api": "ngf_field_render_host (C ABI, host buffers)" if world == 1 else "H2D + ngf_field_render + all-gather + D2H"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "roofline": roofline,
    }

    # ---- dense regime side measurement (tensor-bound): every in-box sample is colour-active
    if world == 1 and not args.no_dense:
        line["dense_regime"] = dense_regime(ngf_b200, synth, dev, pk, dev_rays)

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample of the same workload
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import restate_field as R
        spec = R.spec_from_state("triplane", state, alpha_volume=occ, gauge_on=True, **kw)
        threads = os.cpu_count() or 1
        sample = host[0][::2].contiguous()
        v, done, dt = cpu_port_rate(spec, R, sample, args.cpu_seconds, threads)
        line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": f"{done} rays (every 2nd ray of frame 0, 4096-ray chunks) in {dt:.1f} s"}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def dense_regime(ngf_b200, synth, dev, pk, dev_rays):
    """Same frame, fog field without alpha mask and weight threshold -1: every sample inside the box runs the colour
    MLP (SURVEY.md §8d "dense-MLP microbench").  Reports the tensor roofline of the same kernel."""
    kw = synth.field_kwargs("C2")
    f = ngf_b200.TriPlane(kw["aabb"], kw["gridSize"], dev, near_far=kw["near_far"], step_ratio=kw["step_ratio"],
                          distance_scale=kw["distance_scale"], rayMarch_weight_thres=-1.0, gauge_start=0)
    st = synth.field_state("triplane", "fog")
    st["density_decoder.bias"] = st["density_decoder.bias"] - 12.0          # thin fog: rays never saturate
    synth.load_into(f, st)
    n = dev_rays[0].shape[0]
    for i in range(2):
        f(dev_rays[i], white_bg=True, N_samples=S, iteration=30001, image_width=W)
    torch.cuda.synchronize()
    stats = f.last_stats()
    steps = 5
    f.kernel_timing(steps)
    for i in range(steps):
        f(dev_rays[i], white_bg=True, N_samples=S, iteration=30001, image_width=W)
    n_k, march_ms, colour_ms = f.kernel_timing_read()
    m_s, c_s = march_ms / n_k * 1e-3, colour_ms / n_k * 1e-3
    nV, nA = stats["samples_density"] / n, stats["samples_colour"] / n
    flops = n * nA * MLP_FLOPS_PER_COLOUR_SAMPLE
    executed = n * nA * 2 * (160 * 64 + 64 * 64 + 64 * 3)                   # basis folded into layer 1, K padded to 160
    return {"workload": "fog field, no alpha mask, weight threshold -1: all in-box samples colour-active",
            "rays_per_s": n / (m_s + c_s), "march_kernel_ms": m_s * 1e3, "colour_kernel_ms": c_s * 1e3,
            "density_samples_per_ray": nV, "colour_samples_per_ray": nA,
            "roofline": {"bound": "tensor", "kernel": "ngf_colour_kernel<TriPlane,tcgen05>", "achieved": flops / c_s / 1e12,
                         "peak": pk["tensor"], "unit": "TFLOP/s", "frac": flops / c_s / 1e12 / pk["tensor"],
                         "executed_tflops": executed / c_s / 1e12,
                         "note": "achieved counts the reference's arithmetic (70 400 FLOP/colour sample incl. the 144x144 "
                                 "basis); executed counts what the kernel runs after folding the basis into layer 1"}}


if __name__ == "__main__":
    main()
