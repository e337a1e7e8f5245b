#!/usr/bin/env python
"""bench.py — rays/s of the volumetric-rendering hot path (BASELINE.json: 800x800 rays, 192 samples/ray).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--field hull|fog]

A "step" is one pass of the hot path over one batch of synthetic rays:
  N = 1   one 800x800 TriPlane frame (640 000 rays, 192 samples/ray, gauge on, alpha mask) = BASELINE configs[1];
  N > 1   a batch of N such frames whose rays are dealt to the ranks in interleaved 3200-ray blocks (4 image rows) (each rank renders
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
BLOCK = 3200                      # 4 image rows: shards keep whole 8x4-pixel warp tiles
H = W = 800
S = 192
RAYS_PER_FRAME = H * W
HBM_BYTES_PER_RAY = 40            # 24 B ray in + 16 B (rgb, depth) out: SURVEY.md §8(d)
MLP_FLOPS_PER_COLOUR_SAMPLE = 70400   # reference arithmetic incl. the 144x144 basis: SURVEY.md §8(a) row a9
MLP_FLOPS_PER_DENSITY_SAMPLE = 96
MLP_FLOPS_EXECUTED_PER_COLOUR_SAMPLE = 2 * (160 * 64 + 64 * 64 + 64 * 3)   # basis folded into layer 1, K padded to 160


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--field", default="hull", choices=["hull", "fog"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-regime (tensor-bound) side measurement")
    ap.add_argument("--no-extra", action="store_true", help="skip the InfoInv / UV-Mapping side measurements (BASELINE configs[2], [3])")
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
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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
    """Rays rank owns of the batch made of the frames of poses (pose0 + j) % N_POSES, j < world (interleaved BLOCK-ray
    blocks).  Every N renders the same 16 poses equally often, so per-rank work is the same at every N (weak scaling)
    while each of a rank's 16 batches is a different ray set."""
    if world == 1:
        return synth.config_rays("C2", pose0)
    g = torch.arange(world * RAYS_PER_FRAME)
    mine = g[((g // BLOCK) % world) == rank]
    frames = {}
    out = torch.empty((mine.numel(), 6))
    fidx = mine // RAYS_PER_FRAME
    for f in fidx.unique().tolist():
        frames[f] = synth.config_rays("C2", (pose0 + f) % N_POSES)
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
    box and the reference is Python (nothing to compile into oracle/_ref), so this times the oracle port — bit-identical
    to the imported reference on the same torch build (tests/golden).  A step renders the whole 640 000-ray frame in
    4096-ray chunks (TriPlane/main.py:60-71) when warm-up + K steps fit ~4 minutes, else a strided sample of it."""
    if rank != 0:
        return
    from oracle import cases as K
    from oracle import restate_field as R
    case = K.Case("bench", kind=args.field, config="C2", n_samples=S)
    state, kw, occ, _ = K.build_inputs(case)
    spec = R.spec_from_state("triplane", state, alpha_volume=occ, gauge_on=True, **kw)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    def render_sample(i, n_sample):
        rays = K.synth.config_rays("C2", i % N_POSES)
        if n_sample < RAYS_PER_FRAME:
            rays = rays[:: RAYS_PER_FRAME // n_sample][:n_sample].contiguous()
        t0 = time.perf_counter()
        R.render(spec, rays, N_samples=S)
        return time.perf_counter() - t0
    render_sample(0, 4096)
    probe = 16384 / render_sample(1, 16384)
    n_steps = max(args.steps + args.warmup, 1)
    n_sample = int(probe * 240.0 / n_steps) // 1024 * 1024
    n_sample = RAYS_PER_FRAME if n_sample >= RAYS_PER_FRAME else max(1024, n_sample)
    for i in range(args.warmup):
        render_sample(i, n_sample)
    total = sum(render_sample(i, n_sample) for i in range(args.steps))
    v = n_sample * args.steps / total
    what = "the whole 640000-ray frame" if n_sample == RAYS_PER_FRAME else f"a {n_sample}-ray strided sample of one frame"
    print(json.dumps({
        "impl": "reference", "metric": "rays/sec (800x800, 192 samples/ray)", "value": v, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"TriPlane {args.field} field, 800x800 rays x 192 samples/ray, gauge on, 256^3 alpha mask "
                               f"(BASELINE configs[1]); each step = {what}", "rays_per_step": n_sample},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": f"{what} per step, 4096-ray chunks, torch CPU fp32"},
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
    host = [shard_of_batch(synth, p, world, rank).pin_memory() for p in range(N_POSES)]
    n_local = host[0].shape[0]
    dev_rays = [h.to(dev) for h in host]
    n_batch = world * RAYS_PER_FRAME

    # ---- multi-GPU: the exchange runs inside the C ABI over NVLink peer memory (copy engines by default);
    # NGF_BENCH_COMM=store|nccl selects the fused-store variant or the torch.distributed NCCL all-gather baseline
    sharded, comm_mode, comm_note = None, None, None
    if world > 1:
        comm_mode = os.environ.get("NGF_BENCH_COMM", "copy")
        try:
            sharded = ngf_b200.ShardedFrameRenderer(field, n_batch, BLOCK, mode=comm_mode)
            ok = 1
        except RuntimeError as e:
            ok, comm_note = 0, f"{comm_mode} unavailable on rank {rank}: {e}"
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if sharded is not None and sharded.comm is not None:
                sharded.comm.close()
            comm_note = comm_note or f"{comm_mode} unavailable on another rank"
            comm_mode = "nccl"
            sharded = ngf_b200.ShardedFrameRenderer(field, n_batch, BLOCK, mode="nccl")

    prev = []

    def step_device(rays):
        if world == 1:
            out = field(rays, white_bg=True, N_samples=S, iteration=30001, image_width=W)
            return out["rgb_map"], out["depth_map"]
        # render my shard; the exchange of this batch overlaps the kernels of the next one.  The consumer of a gathered
        # batch trails by one step: complete the all-gather of the previous ticket on this stream, then release it.
        t = sharded.submit(rays, N_samples=S, white_bg=True, iteration=30001, image_width=W)
        if prev:
            p0 = prev.pop(0)
            sharded.result(p0)
            sharded.release(p0)
        prev.append(t)
        return t, None

    def drain_device():
        while prev:
            p0 = prev.pop(0)
            sharded.result(p0)
            sharded.release(p0)

    def barrier():
        if world > 1:
            torch.cuda.synchronize()          # includes the side streams of the overlapped exchange
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- parity of the gathered batch (outside every timed region): batch 0 = frames 0..world-1, gathered on every
    # rank, against this rank's own single-GPU render of the same frames through ngf_field_render
    parity = None
    if world > 1:
        t = sharded.submit(dev_rays[0], N_samples=S, white_bg=True, iteration=30001, image_width=W)
        got = sharded.result(t).clone()
        sharded.release(t)
        e_rgb = e_dep = 0.0
        for fidx in range(world):
            full = synth.config_rays("C2", fidx).to(dev)
            o = field(full, white_bg=True, N_samples=S, iteration=30001, image_width=W)
            g = got[fidx * RAYS_PER_FRAME:(fidx + 1) * RAYS_PER_FRAME]
            e_rgb = max(e_rgb, float((g[:, :3] - o["rgb_map"]).abs().max()))
            e_dep = max(e_dep, float((g[:, 3] - o["depth_map"]).abs().max()))
            del full, o
        e_rgb, e_dep = max_over_ranks(e_rgb), max_over_ranks(e_dep)
        parity = {"parity_ok": bool(e_rgb < 1e-3 and e_dep < 2e-3), "rgb_max_abs": e_rgb, "depth_max_abs": e_dep,
                  "what": f"gathered batch 0 ({world} frames, every rank) vs single-GPU ngf_field_render of the same frames; "
                          "tolerance 1e-3 rgb / 2e-3 depth (the golden tolerance; differences come from the order of the "
                          "fp32 atomic adds only)"}

    # ---- device-resident timed region (value).  One GPU: the frames are independent, so step i is issued on CUDA stream
    # i % N_STREAMS (ngf_field_render keeps one workspace per caller stream) and the march of one frame runs beside the
    # colour pass of its neighbours; the region is forked from / joined to the current stream, which carries the events.
    N_STREAMS = int(os.environ.get("NGF_BENCH_STREAMS", "3"))
    if world > 1 and comm_mode == "nccl":
        N_STREAMS = 1                 # the NCCL fall-back double-buffers on its own side stream
    side = [torch.cuda.Stream(dev) for _ in range(N_STREAMS)] if N_STREAMS > 1 else []

    def run_steps(n, streams):
        if not streams:
            for i in range(n):
                step_device(dev_rays[i % N_POSES])
            if world > 1:
                drain_device()
            return
        cur = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        for st in streams:
            st.wait_event(fork)
        for i in range(n):
            with torch.cuda.stream(streams[i % len(streams)]):
                step_device(dev_rays[i % N_POSES])
        for st in streams:
            cur.wait_stream(st)
        if world > 1:
            drain_device()

    run_steps(max(args.warmup, 3), side)
    barrier()
    stats = field.last_stats() if world == 1 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    serial = None
    if True:
        # kernel durations for the roofline come from a serial pass on one stream (events around overlapping kernels would
        # time the overlap, not the kernel); its step time is reported beside `value` as value_single_stream
        n_serial = min(args.steps, 4096)
        run_steps(3, [])
        field.kernel_timing(n_serial)
        barrier()
        e0.record()
        run_steps(n_serial, [])
        e1.record()
        barrier()
        serial_ms = max_over_ranks(e0.elapsed_time(e1))
        n_k, march_ms, colour_ms = field.kernel_timing_read()
        field.kernel_timing(0)
        serial = {"steps": n_serial, "ms_per_step": serial_ms / n_serial, "value": n_batch * n_serial / (serial_ms * 1e-3)}
        run_steps(3, side)
    l0 = ngf_b200._lib.launch_count()
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        run_steps(args.steps, side)
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ngf_b200._lib.launch_count() - l0
    value = n_batch * args.steps / (ms * 1e-3)
    if stats is None:
        # the sharded path renders from the comm's own workspace; the sample statistics come from one plain render
        field(dev_rays[0], white_bg=True, N_samples=S, iteration=30001, image_width=W)
        stats = field.last_stats()

    # ---- end-to-end timed region: pinned host rays -> H2D -> render (+ exchange) -> D2H of the results, through the
    # C ABI's host-buffer entry points, 3 steps in flight
    pend = []
    if world == 1:
        outs = [(torch.empty((n_local, 3)).pin_memory(), torch.empty((n_local,)).pin_memory()) for _ in range(3)]
        def step_e2e(i):
            r, d = outs[i % 3]
            pend.append(field.render_host_async(host[i % N_POSES], r, d, white_bg=True, N_samples=S, image_width=W,
                                                iteration=30001))
            if len(pend) > 2:
                field.host_wait(pend.pop(0))
        def e2e_drain():
            while pend:
                field.host_wait(pend.pop(0))
        e2e_api = "ngf_b200.render_frames: ngf_field_render_host_async per frame (C ABI, pinned host buffers, 3 frames in flight)"
    elif sharded.comm is not None:
        # each rank uploads its 640 000 rays and downloads one whole gathered frame (rows of frame `rank` of the batch)
        res_h = [torch.empty((RAYS_PER_FRAME, 4)).pin_memory() for _ in range(3)]
        def step_e2e(i):
            pend.append(sharded.comm.submit_host(host[i % N_POSES], res_h[i % 3], first_row=rank * RAYS_PER_FRAME,
                                                 N_samples=S, white_bg=True, image_width=W, iteration=30001))
            if len(pend) > 2:
                sharded.comm.wait(pend.pop(0))
        def e2e_drain():
            while pend:
                sharded.comm.wait(pend.pop(0))
        e2e_api = ("ngf_field_render_sharded_host_async per batch (C ABI: pinned H2D of the rank's rays, render, peer-memory "
                   "all-gather, D2H of one gathered frame per rank; 3 batches in flight)")
    else:
        copy_s = torch.cuda.Stream(device=dev)
        stage = [torch.empty_like(dev_rays[0]) for _ in range(2)]
        res_h = [torch.empty((n_local, 4)).pin_memory() for _ in range(2)]
        up, free = [torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)]
        def step_e2e(i):
            b = i % 2
            cur = torch.cuda.current_stream(dev)
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(free[b])
                stage[b].copy_(host[i % N_POSES], non_blocking=True)
                up[b].record(copy_s)
            cur.wait_event(up[b])
            t = sharded.submit(stage[b], N_samples=S, white_bg=True, iteration=30001, image_width=W)
            free[b].record(cur)
            pend.append(sharded.download(t, res_h[b], n_local))
            if len(pend) > 1:
                pend.pop(0).synchronize()
        def e2e_drain():
            while pend:
                pend.pop(0).synchronize()
        e2e_api = "pinned H2D + ngf_field_render + torch.distributed NCCL all-gather + D2H, host one step behind"
    d2h = n_local * 16
    for i in range(3):
        step_e2e(i)
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    e2e_drain()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_batch * args.steps / e2e_s

    # ---- the transport floor of that region: the same H2D and D2H copies alone (no kernels), all ranks at once
    cp_in, cp_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    d_in = torch.empty_like(dev_rays[0])
    d_out = torch.empty((n_local, 4), device=dev)
    h_out = torch.empty((n_local, 4)).pin_memory()
    def copies(n):
        for i in range(n):
            with torch.cuda.stream(cp_in):
                d_in.copy_(host[i % N_POSES], non_blocking=True)
            with torch.cuda.stream(cp_out):
                h_out.copy_(d_out, non_blocking=True)
    copies(3)
    barrier()
    n_cp = max(10, min(args.steps, 100))
    t0 = time.perf_counter()
    copies(n_cp)
    barrier()
    floor_s = max_over_ranks(time.perf_counter() - t0) / n_cp
    e2e_floor = n_batch / floor_s

    # ---- second end-to-end figure: what evaluation_path needs (TriPlane/main.py:155-161) — a camera pose in, the
    # uint8 image out: rays generated on the device, uint8 conversion on the device, 1.92 MB D2H per frame
    if world == 1:
        e2e_cam = camera_e2e(ngf_b200, synth, field, dev, args.steps)
    elif sharded.comm is not None:
        e2e_cam = camera_e2e_sharded(synth, sharded.comm, rank, world, args.steps, barrier, max_over_ranks)
    else:
        e2e_cam = None

    # ---- roofline of the dominant kernel, from CUDA events around each kernel of the pair (measured live above)
    pk = peaks()
    n_k = max(n_k, 1)
    march_s, colour_s = march_ms / n_k * 1e-3, colour_ms / n_k * 1e-3
    # share of the step: against the step the kernels were timed in (one GPU: the serial single-stream pass)
    step_s = (serial["ms_per_step"] if serial else ms / args.steps) * 1e-3
    nV = stats["samples_density"] / n_local
    nA = stats["samples_colour"] / n_local
    alg_bytes = n_local * HBM_BYTES_PER_RAY
    tj = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
    mlp_flops = n_local * nA * MLP_FLOPS_PER_COLOUR_SAMPLE
    mlp_exec = n_local * nA * MLP_FLOPS_EXECUTED_PER_COLOUR_SAMPLE
    k_march = {"bound": "hbm", "kernel": "ngf_march_kernel<TriPlane>", "achieved": alg_bytes / march_s / 1e9 if march_s > 0 else None,
               "peak": pk["hbm"], "unit": "GB/s", "frac": alg_bytes / march_s / 1e9 / pk["hbm"] if march_s > 0 else None,
               "traffic": tj.get(f"march_kernel_{args.field}_dram_bytes_per_launch"), "kernel_ms": march_s * 1e3,
               "kernel_share_of_step": march_s / step_s, "algorithmic_bytes_per_launch": alg_bytes,
               "note": "sparse regime: 40 B/ray of compulsory HBM traffic; the march is issue / L2-gather bound (DESIGN.md)"}
    k_colour = {"bound": "tensor", "kernel": "ngf_colour_kernel<TriPlane,tcgen05>",
                "achieved": mlp_flops / colour_s / 1e12 if colour_s > 0 else None, "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": mlp_flops / colour_s / 1e12 / pk["tensor"] if colour_s > 0 else None,
                "frac_executed": mlp_exec / colour_s / 1e12 / pk["tensor"] if colour_s > 0 else None,
                "traffic": tj.get(f"colour_kernel_{args.field}_dram_bytes_per_launch"), "kernel_ms": colour_s * 1e3,
                "kernel_share_of_step": colour_s / step_s, "algorithmic_flops_per_launch": mlp_flops,
                "note": "frac counts the reference's arithmetic (70 400 FLOP per colour sample incl. the 144x144 basis); "
                        "frac_executed what the kernel runs after folding the basis into layer 1 (K padded to 160)"}
    dom, other = (k_colour, k_march) if colour_s >= march_s else (k_march, k_colour)
    roofline = dict(dom)
    roofline.update({"peak_source": pk["src"], "other_kernel": other,
                     "density_samples_per_ray": nV, "colour_samples_per_ray": nA,
                     "mlp_flop_roofline_frac_of_step": n_local * (nV * MLP_FLOPS_PER_DENSITY_SAMPLE + nA * MLP_FLOPS_PER_COLOUR_SAMPLE)
                                                       / step_s / 1e12 / pk["tensor"]})

    coll = "" if world == 1 else {"copy": "peer-memory all-gather on the copy engines (C ABI)",
                                  "store": "peer-memory all-gather fused into the finalize kernel (C ABI)",
                                  "nccl": "torch.distributed NCCL all-gather"}[comm_mode]
    line = {
        "metric": "rays/sec (800x800, 192 samples/ray)", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"TriPlane {args.field} field, 800x800 rays x 192 samples/ray, gauge on, 256^3 alpha mask "
                               f"(BASELINE configs[1]" + (")" if world == 1 else f"; configs[4]: {world} frames/step ray-sharded + one all-gather of the rendered batch)"),
                   "rays_per_step": n_batch,
                   "arithmetic": "fp32 march / density / compositing; colour MLP fp16 operands with fp32 accumulation (tcgen05, TMEM)", "l2": f"inputs rotate over {N_POSES} poses ({N_POSES * n_local * 24 / 1e6:.0f} MB of rays per rank > 126 MB L2)",
                   "streams": N_STREAMS,
                   "parallelism": f"single GPU, independent frames on {N_STREAMS} CUDA streams" if world == 1 else f"ray-sharded dp{world}, {BLOCK}-ray interleaved blocks, {coll}"},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": n_local * 24, "d2h_bytes_per_step": d2h,
                "api": e2e_api,
                "transport_floor": {"value": e2e_floor, "unit": "rays/s", "frac_of_floor": e2e_value / e2e_floor,
                                    "what": "the same pinned H2D (rays) and D2H (results) copies alone, no kernels, all ranks "
                                            "concurrently: what the PCIe / host-memory path of this box allows"}},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "roofline": roofline,
    }
    if serial:
        line["value_single_stream"] = {"value": serial["value"], "ms_per_step": serial["ms_per_step"], "steps": serial["steps"],
                                       "what": "the same frames issued back to back on ONE stream (no overlap between frames); "
                                               "roofline.kernel_ms and kernel_share_of_step are measured in this pass"}
    if e2e_cam is not None:
        line["e2e_camera"] = e2e_cam
    if parity is not None:
        line["parity_ok"] = parity["parity_ok"]
        line["parity"] = parity
    if comm_note:
        line["config"]["collective_note"] = comm_note

    # ---- dense regime side measurement (tensor-bound): every in-box sample is colour-active
    if world == 1 and not args.no_dense:
        line["dense_regime"] = dense_regime(ngf_b200, synth, dev, pk, dev_rays)

    # ---- the other single-GPU configurations of BASELINE.json, as side measurements
    if world == 1 and not args.no_extra:
        line["other_configs"] = {"infoinv": infoinv_config(ngf_b200, synth, dev, pk, dev_rays, host, args),
                                 "neutex": neutex_config(ngf_b200, synth, dev, pk, args)}

    # ---- baselines: the oracle port on this box's host cores and on this GPU through torch (the reference's normal
    # device, TriPlane/main.py:18), bounded samples of the same workload
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import restate_field as R
        spec = R.spec_from_state("triplane", state, alpha_volume=occ, gauge_on=True, **kw)
        threads = os.cpu_count() or 1
        sample = torch.cat([host[0], host[1]])
        v, done, dt = cpu_port_rate(spec, R, sample, args.cpu_seconds, threads)
        line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": f"first {done} rays of frames 0-1 (4096-ray chunks, torch CPU fp32, {threads} threads) in {dt:.1f} s"}
        line["torch_cuda_baseline"] = torch_cuda_rate(R, spec, dev, host)
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        if sharded.comm is not None:
            barrier()
            sharded.comm.close()
        dist.destroy_process_group()


def torch_cuda_rate(R, spec, dev, host):
    """The reference's algorithm on its normal device: the oracle port (same torch ops, same order) with every tensor on
    this GPU, one whole 640 000-ray frame in 4096-ray chunks as TriPlane/main.py:60-71 renders it (rays uploaded per chunk,
    main.py:65).  One warm-up frame, one timed frame."""
    gspec = R.spec_to_device(spec, dev)
    def frame(i):
        rays = host[i]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = []
        for s0 in range(0, rays.shape[0], 4096):
            out.append(R.render_chunk(gspec, rays[s0:s0 + 4096].to(dev, non_blocking=True), white_bg=True, N_samples=S))
        rgb = torch.cat([o[0] for o in out]).cpu()
        torch.cuda.synchronize()
        return time.perf_counter() - t0, rgb
    frame(1)
    dt, _ = frame(0)
    return {"value": RAYS_PER_FRAME / dt, "unit": "rays/s", "kind": "port on torch-CUDA (eager PyTorch ops on the same B200)",
            "sample": f"one whole frame, 4096-ray chunks, {dt * 1e3:.0f} ms", "torch": torch.__version__}


def camera_e2e(ngf_b200, synth, field, dev, steps):
    """evaluation_path's per-frame work (TriPlane/main.py:155-161,116): c2w -> rays -> render -> uint8 image, through
    ngf_field_render_camera_u8_host_async: the pose travels as 18 numbers, rays are generated in the march kernel, the
    uint8 conversion runs on the device and 1.92 MB come back per frame; 3 frames in flight."""
    steps = max(10, min(steps, 1000))
    poses = [synth.look_at_c2w(*synth.pose_angles(p)) for p in range(N_POSES)]
    u8_h = [torch.empty((RAYS_PER_FRAME, 3), dtype=torch.uint8).pin_memory() for _ in range(3)]
    pend = []
    def step(i):
        pend.append(field.render_camera_u8_host_async(poses[i % N_POSES], H, W, synth.FOCAL_800, u8_h[i % 3], white_bg=True,
                                                      N_samples=S, iteration=30001))
        if len(pend) > 2:
            field.host_wait(pend.pop(0))
    def drain():
        while pend:
            field.host_wait(pend.pop(0))
    for i in range(3):
        step(i)
    drain()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    drain()
    dt = time.perf_counter() - t0
    return {"value": RAYS_PER_FRAME * steps / dt, "unit": "rays/s", "h2d_bytes_per_step": 18 * 4,
            "d2h_bytes_per_step": RAYS_PER_FRAME * 3, "steps": steps,
            "api": "ngf_field_render_camera_u8_host_async (C ABI): pose in, rays generated in the march kernel, uint8 image "
                   "out, 3 frames in flight"}


def camera_e2e_sharded(synth, comm, rank, world, steps, barrier, max_over_ranks):
    """The same job for a ray-sharded batch of `world` frames per step (ngf_field_render_sharded_camera_u8_host_async): every
    rank gets the batch's poses (48 B per frame) from pinned host memory, renders its interleaved blocks with rays generated in
    the march kernel, the rows are exchanged over peer memory and every rank downloads ONE gathered frame as uint8."""
    steps = max(10, min(steps, 1000))
    batches = [torch.stack([synth.look_at_c2w(*synth.pose_angles((p + j) % N_POSES)) for j in range(world)]).contiguous().pin_memory()
               for p in range(N_POSES)]
    u8_h = [torch.empty((RAYS_PER_FRAME, 3), dtype=torch.uint8).pin_memory() for _ in range(3)]
    pend = []
    def step(i):
        pend.append(comm.submit_camera_host(batches[i % N_POSES], H, W, synth.FOCAL_800, u8_h[i % 3], first_row=rank * RAYS_PER_FRAME,
                                            N_samples=S, white_bg=True, iteration=30001))
        if len(pend) > 2:
            comm.wait(pend.pop(0))
    def drain():
        while pend:
            comm.wait(pend.pop(0))
    for i in range(3):
        step(i)
    drain()
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    drain()
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    return {"value": world * RAYS_PER_FRAME * steps / dt, "unit": "rays/s", "h2d_bytes_per_step": world * 48,
            "d2h_bytes_per_step": RAYS_PER_FRAME * 3, "steps": steps,
            "api": "ngf_field_render_sharded_camera_u8_host_async (C ABI): the batch's poses in, rays generated in the march "
                   "kernel, peer-memory all-gather, one gathered frame per rank out as uint8, 3 batches in flight"}


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


def infoinv_config(ngf_b200, synth, dev, pk, dev_rays, host, args):
    """BASELINE configs[2]: InfoInv (sinusoidal phase product, 72->32->32->1 density MLP, 216-wide colour input),
    800x800 rays x 192 samples, hull field + alpha mask."""
    kw = synth.field_kwargs("C2")
    f = ngf_b200.InfoInvTriPlane(kw["aabb"], kw["gridSize"], dev, near_far=kw["near_far"], step_ratio=kw["step_ratio"],
                                 distance_scale=kw["distance_scale"], rayMarch_weight_thres=kw["rayMarch_weight_thres"])
    state, occ = synth.field_state("infoinv", "hull"), synth.occupancy_volume("hull")
    synth.load_into(f, state, occ, ngf_b200.AlphaGridMask)
    n = dev_rays[0].shape[0]
    for i in range(3):
        f(dev_rays[i], white_bg=True, N_samples=S, infoinv=True, image_width=W)
    torch.cuda.synchronize()
    st = f.last_stats()
    steps = 20
    f.kernel_timing(steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        f(dev_rays[i % N_POSES], white_bg=True, N_samples=S, infoinv=True, image_width=W)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n_k, m_ms, c_ms = f.kernel_timing_read()
    f.kernel_timing(0)
    nV, nA = st["samples_density"] / n, st["samples_colour"] / n
    flops = n * (nV * 6720 + nA * 131456)                        # InfoInv MLP FLOPs per sample
    # end to end through the C ABI's host-buffer call, 3 frames in flight
    outs = [(torch.empty((n, 3)).pin_memory(), torch.empty((n,)).pin_memory()) for _ in range(3)]
    pend = []
    def step_e2e(i):
        r, d = outs[i % 3]
        pend.append(f.render_host_async(host[i % N_POSES], r, d, white_bg=True, N_samples=S, image_width=W, infoinv=True))
        if len(pend) > 2:
            f.host_wait(pend.pop(0))
    for i in range(3):
        step_e2e(i)
    while pend:
        f.host_wait(pend.pop(0))
    t0 = time.perf_counter()
    for i in range(steps):
        step_e2e(i)
    while pend:
        f.host_wait(pend.pop(0))
    e2e_s = (time.perf_counter() - t0) / steps
    march_s, colour_s = m_ms / n_k * 1e-3, c_ms / n_k * 1e-3
    dom_is_march = march_s >= colour_s
    res = {"workload": "InfoInv hull field, 800x800 rays x 192 samples/ray, infoinv=True, 256^3 alpha mask (BASELINE configs[2])",
           "rays_per_s": n / (ms * 1e-3), "ms_per_frame": ms, "march_kernel_ms": m_ms / n_k, "colour_kernel_ms": c_ms / n_k,
           "density_samples_per_ray": nV, "colour_samples_per_ray": nA,
           "mlp_flop_roofline_frac": flops / (ms * 1e-3) / 1e12 / pk["tensor"],
           "e2e": {"value": n / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 16,
                   "api": "ngf_field_render_host_async, 3 frames in flight"},
           "roofline": {"bound": "tensor", "kernel": "ngf_march_kernel<InfoInv>" if dom_is_march else "ngf_colour_kernel<InfoInv,tcgen05>",
                        "achieved": (n * nV * 6720 / march_s if dom_is_march else n * nA * 131456 / colour_s) / 1e12,
                        "peak": pk["tensor"], "unit": "TFLOP/s",
                        "frac": (n * nV * 6720 / march_s if dom_is_march else n * nA * 131456 / colour_s) / 1e12 / pk["tensor"],
                        "kernel_ms": (march_s if dom_is_march else colour_s) * 1e3,
                        "note": "reference MLP arithmetic of the kernel with the larger share (density MLP 6 720 FLOP per valid "
                                "sample in the march kernel, colour MLP 131 456 FLOP per active sample in the colour kernel)"}}
    if not args.no_cpu_baseline:
        from oracle import restate_field as R
        spec = R.spec_from_state("infoinv", state, alpha_volume=occ, infoinv=True, **kw)
        threads = os.cpu_count() or 1
        v, done, dt = cpu_port_rate(spec, R, host[0], min(args.cpu_seconds, 6.0), threads)
        res["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                               "sample": f"first {done} rays of frame 0 (4096-ray chunks, torch CPU fp32) in {dt:.1f} s"}
    return res


def neutex_config(ngf_b200, synth, dev, pk, args):
    """BASELINE configs[3]: UV-Mapping NeuTex test render of DTU scan83, camera 33 (the cameras ship with the reference and
    are kept in neural-gauge-fields_b200/data/scan83_cameras.npz), 600x800 rays x 64 samples, random-init networks of the
    reference's shapes (no checkpoint exists offline), explicit jitter noise."""
    m = ngf_b200.NeuTex(device=dev)
    nstate = synth.neutex_state(0)
    m.load_state_dict(nstate)
    campos, raydir = synth.scan83_camera(33)              # the real DTU scan83 centre camera (data/dtu.py:113-114,119-182)
    R = raydir.shape[1]
    noise = synth.neutex_noise(R)
    bg = torch.ones(1, 3)
    d_cam, d_rd, d_nz, d_bg = campos.to(dev), raydir.to(dev), noise.to(dev), bg.to(dev)
    m(d_cam, d_rd, d_bg, noise=d_nz)
    steps = 3
    m.kernel_timing(steps)
    for i in range(steps):
        m(d_cam, d_rd, d_bg, noise=d_nz)
    k, a_ms, b_ms, c_ms = m.kernel_timing_read()
    nv = m.last_valid_samples()
    per_sample = 2 * (63 * 256 + 10 * 256 * 256 + 256 + 63 * 64 + 64 * 128 + 2 * 128 * 128 + 256 + 42 * 256 + 5 * 256 * 256
                      + 768 + 295 * 256 + 3 * 256 * 256 + 768)      # geometry + gauge + texture, reference arithmetic
    total_ms = (a_ms + b_ms + c_ms) / k
    # end to end through host buffers: the jitter is drawn on the device from a seed (ngf_neutex_render_host_seeded), as the
    # reference draws it inside cube_ray_generation; the explicit-noise call (256 B more per ray over PCIe) is timed beside it
    h_rd, h_nz = raydir.pin_memory(), noise.pin_memory()
    m.render_host(campos, h_rd, bg, seed=1)
    t0 = time.perf_counter()
    m.render_host(campos, h_rd, bg, seed=2)
    e2e_s = time.perf_counter() - t0
    m.render_host(campos, h_rd, bg, h_nz)
    t0 = time.perf_counter()
    m.render_host(campos, h_rd, bg, h_nz)
    e2e_noise_s = time.perf_counter() - t0
    res = {"workload": "UV-Mapping NeuTex, DTU scan83 camera 33, 600x800 rays x 64 samples/ray, square primitive, jitter 0.05 (BASELINE configs[3])",
           "rays_per_s": R / (total_ms * 1e-3), "ms_per_frame": total_ms, "raygen_ms": a_ms / k, "mlp_kernel_ms": b_ms / k,
           "march_ms": c_ms / k, "in_cube_samples_per_ray": nv / R,
           "e2e": {"value": R / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": R * 3 * 4, "d2h_bytes_per_step": R * 16,
                   "api": "ngf_neutex_render_host_seeded",
                   "with_uploaded_noise": {"value": R / e2e_noise_s, "h2d_bytes_per_step": R * (3 + 64) * 4,
                                           "api": "ngf_neutex_render_host"}},
           "roofline": {"bound": "tensor", "kernel": "ntx_mlp_kernel", "achieved": nv * per_sample / (b_ms / k * 1e-3) / 1e12,
                        "peak": pk["tensor"], "unit": "TFLOP/s", "frac": nv * per_sample / (b_ms / k * 1e-3) / 1e12 / pk["tensor"],
                        "kernel_ms": b_ms / k,
                        "note": "reference arithmetic (2.66 MFLOP) of the in-cube samples only; the reference itself pushes all "
                                "64 samples/ray through the MLPs"}}
    if not args.no_cpu_baseline:
        from oracle import restate_neutex as RN
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        spec = RN.NeuTexSpec(state=nstate)
        n_s = 4096
        try:
            sel = slice(R // 2, R // 2 + n_s)              # rows through the image centre: rays that cross the cube
            RN.render(spec, campos, raydir[:, :1024], bg, noise[:, :1024])
            t0 = time.perf_counter()
            RN.render(spec, campos, raydir[:, sel], bg, noise[:, sel])
            dt = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": n_s / dt, "unit": "rays/s", "cores": threads, "kind": "port",
                                   "sample": f"{n_s} rays from the image centre (1024-ray chunks, torch CPU fp32) in {dt:.1f} s"}
        except Exception as e:                      # the baseline is a report, never a reason to lose the bench line
            res["cpu_baseline"] = {"unavailable": repr(e)[:200]}
    return res


if __name__ == "__main__":
    main()
