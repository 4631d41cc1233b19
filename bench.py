#!/usr/bin/env python
"""Benchmark of the ConvVAE hot path (BASELINE.json metric: spectral frames/sec, ConvVAE fwd+bwd
@ [B,513,256]; cfg2 = 64 x 256 = 16,384 frames per GPU per step, 10 speakers, fwd+bwd+Adam).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # the reference arithmetic on host cores
                                                           # (oracle port: TF 1.2.1 cannot run here)
Prints ONE JSON line (rank 0).  A "step" = one pass of the hot path over one batch of synthetic
frames: encode -> sample -> decode -> KL + Gaussian log-density -> backward -> (all-reduce) -> Adam.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_GPU = 64 * 256            # cfg2: batch 64 utterances x T 256 frames
FLOP_PER_FRAME_TRAIN = 30.35e6       # fwd 10,127,856 FLOP; fwd+bwd ~ 3x - E0 dgrad
BYTES_PER_FRAME_TRAIN = 4406.0       # compulsory HBM bytes / frame at N = 16,384
METRIC = "spectral frames/sec ConvVAE fwd+bwd+Adam @ [64,513,256] per GPU"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], source="measured")
    return dict(hbm_gbs=6650.0, tf=1400.0, tf_burst=1590.0, source="fallback")   # B200_PROFILING.md


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU side: the reference arithmetic (oracle fp32 twin) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_reference_step_fn(n_frames):
    """Returns (step, cores): one fwd+bwd+Adam of the oracle's fp32 twin on n_frames frames."""
    import numpy as np
    import torch
    from oracle import convvae_ref as R          # bench.py's cpu_baseline / --impl reference legs only
    from vae_npvc_b200 import vcc2016_vae_arch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arch = vcc2016_vae_arch()
    P = {k: v.astype(np.float32) for k, v in R.init_params(arch, 0).items()}
    x, y, eps = R.make_inputs(arch, n_frames)
    st = {"theta": R.flatten_params(arch, P), "m": 0.0, "v": 0.0, "t": 0, "P": P}

    def step():
        out = R.forward(arch, st["P"], x, y, eps, dtype=torch.float32, with_grads=True)
        g = R.flatten_params(arch, out["grads"])
        st["t"] += 1
        st["theta"], st["m"], st["v"] = R.adam_step(st["theta"], g, st["m"], st["v"], st["t"], 1e-4, 0.5, 0.999)
        st["P"] = R.unflatten_params(arch, st["theta"].astype(np.float32))
        return float(out["G"])
    return step, cores


def pick_threads():
    """The oneDNN/ATen CPU path does not scale to every hardware thread on small tensors: take the
    fastest of a few thread counts on a tiny sample so the baseline is the CPU's best, not its worst."""
    import torch
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    step, _ = cpu_reference_step_fn(128)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        step()
        t0 = time.perf_counter(); step(); dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def time_cpu(n_frames, steps, warmup):
    import torch
    threads = pick_threads()
    step, cores = cpu_reference_step_fn(n_frames)
    torch.set_num_threads(threads)
    cores = threads
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    return n_frames / statistics.median(ts), cores, sum(ts)


def run_reference(args, rank):
    """--impl reference: the reference's CPU arithmetic (TF 1.2.1 is not installable: oracle port)."""
    if rank != 0:
        return
    n = 1024                                    # bounded sample of the cfg2 workload (same frame shape)
    fps, cores, total = time_cpu(n, args.steps, max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * n / fps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the native arm's workload (same frame shape, speakers, fwd+bwd+Adam); each CPU step is a bounded sample of it
        "config": {"workload": "cfg2: ConvVAE (architecture-vae-vcc2016) 10-speaker, %d frames/GPU/step (64x256), fwd+bwd+Adam" % args.frames,
                   "frames_per_gpu_per_step": args.frames, "frames_per_step_sample": n,
                   "note": "CPU restatement of the reference graph (PyTorch/oneDNN fp32), not TF 1.2.1; a step = %d frames of the workload" % n},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d frames/step x %d steps (cfg2 shapes, N bounded for CPU time)" % (n, args.steps)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from vae_npvc_b200 import vcc2016_vae_arch
    from importlib import import_module
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the native arm")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the ONE JSON line: NCCL prints its version banner (and
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p")   # warnings) to stdout unless given a file
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    arch = vcc2016_vae_arch()
    MODEL = getattr(import_module("model.vae"), "ConvVAE")          # the plugin lookup of main.py:39-43
    TRAINER = getattr(import_module("trainer.vae"), "VAETrainer")
    n = args.frames
    machine = MODEL(arch, device=dev, seed=0)
    machine.theta.copy_(machine.engine.init_theta(0, perturb=0.1))  # biases / LN params off 0 / 1
    g = torch.Generator(device="cpu").manual_seed(1 + rank)
    NPOOL = 4
    host_x = [(torch.rand(n, 513, generator=g) * 2 - 1).pin_memory() for _ in range(NPOOL)]
    host_y = [torch.randint(0, arch["y_dim"], (n,), generator=g).pin_memory() for _ in range(NPOOL)]
    dev_x = [t.to(dev) for t in host_x]; dev_y = [t.to(dev) for t in host_y]
    loss = machine.loss(dev_x[0], dev_y[0])
    trainer = TRAINER(loss, arch, None, None)
    step_fn = trainer.opt["g"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- device-resident arm: inputs already in HBM ------------------------------------------
    for i in range(args.warmup):
        step_fn(dev_x[i % NPOOL], dev_y[i % NPOOL])
    l0 = machine.engine.launch_count()
    sampler = ClockSampler(local); sampler.start()
    ms = timed(lambda i: step_fn(dev_x[i % NPOOL], dev_y[i % NPOOL]), args.steps)
    clocks = sampler.stop()
    launches = machine.engine.launch_count() - l0
    value = world * n * args.steps / (ms / 1000.0)

    # ---- end-to-end arm: pinned host inputs, H2D inside the timed region, D2H of the losses -----
    losses_host = torch.empty(3).pin_memory()

    # every step's x / y still cross PCIe inside the timed region; like analyzer.FrameLoader the copy of
    # batch i+1 is issued on a side stream while step i computes (double-buffered device slots)
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [(torch.empty(n, 513, device=dev), torch.empty(n, dtype=torch.int64, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]; consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])                 # the step that used this slot has finished
            slots[k][0].copy_(host_x[i % NPOOL], non_blocking=True)
            slots[k][1].copy_(host_y[i % NPOOL], non_blocking=True)
            ready[k].record(copy_stream)

    def e2e_step(i):
        k = i % 2
        if i == 0:
            prefetch(0)
        prefetch(i + 1)
        torch.cuda.current_stream().wait_event(ready[k])
        lo = step_fn(slots[k][0], slots[k][1])
        consumed[k].record(torch.cuda.current_stream())
        losses_host.copy_(lo, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the user reads the step's loss
    for k in range(2):
        consumed[k].record(torch.cuda.current_stream())
    e2e_step(0); e2e_step(1); e2e_step(2)
    torch.cuda.synchronize()
    for k in range(2):
        consumed[k].record(torch.cuda.current_stream())
    ms_e2e = timed(e2e_step, args.steps)
    e2e = world * n * args.steps / (ms_e2e / 1000.0)
    h2d = host_x[0].numel() * 4 + host_y[0].numel() * 8

    # ---- per-op device timing for the roofline of the dominant kernel -------------------------
    pk = peaks()
    machine.engine.handle.profile_enable(True)
    PS = min(args.steps, 5)
    for i in range(PS):
        step_fn(dev_x[i % NPOOL], dev_y[i % NPOOL])
    torch.cuda.synchronize()
    prof = machine.engine.handle.profile()
    machine.engine.handle.profile_enable(False)
    tot_ms = sum(p["ms"] for p in prof)
    top = max(prof, key=lambda p: p["ms"])                           # the dominant kernel of the step
    gemm_like = [p for p in prof if p["kind"] in (0, 1)]
    top_ms = top["ms"] / top["calls"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))["bytes_per_launch"].get(top["name"])
    ops_ms = {p["name"]: round(p["ms"] / PS, 4) for p in sorted(prof, key=lambda p: -p["ms"])[:12]}
    common = {"kernel": top["name"], "traffic": traffic, "share_of_step": top["ms"] / tot_ms if tot_ms else None,
              "ops_ms_per_step": ops_ms, "gemm_ms_share": sum(p["ms"] for p in gemm_like) / tot_ms if tot_ms else None}
    if top.get("tensor"):
        # tcgen05 op: algorithmic FLOPs of the dense contraction (2*rows*K*N) / CUDA-event time of the op
        top_flops = 2.0 * (top["rows"] / top["calls"]) * top["K"] * top["N"]
        achieved = top_flops / (top_ms * 1e-3) / 1e12 if top_ms > 0 else 0.0
        roofline = dict(bound="tensor", achieved=achieved, peak=pk["tf"], unit="TFLOP/s", frac=achieved / pk["tf"],
                        peak_source=pk["source"] + " bf16 dense (sustained)",
                        note="fp32 path on bf16 tensor cores: every product is 3 bf16 MMAs (hi.hi + hi.lo + lo.hi), "
                             "so 1/3 of the bf16 peak is the ceiling of this fraction", **common)
    else:
        # CUDA-core / streaming op: algorithmic bytes (every operand buffer once) / CUDA-event time of the op
        achieved = top["bytes"] / top["calls"] / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
        roofline = dict(bound="hbm", achieved=achieved, peak=pk["hbm_gbs"], unit="GB/s", frac=achieved / pk["hbm_gbs"],
                        peak_source=pk["source"] + " copy bandwidth",
                        note="algorithmic bytes = each operand buffer of the op read / written once", **common)
    # the largest tensor-core op as well, for the record
    tens = [p for p in prof if p.get("tensor")]
    if tens:
        tt = max(tens, key=lambda p: p["ms"]); tms = tt["ms"] / tt["calls"]
        tf = 2.0 * (tt["rows"] / tt["calls"]) * tt["K"] * tt["N"] / (tms * 1e-3) / 1e12
        roofline["top_tensor_op"] = {"kernel": tt["name"], "achieved_tflops": tf, "frac_of_bf16_peak": tf / pk["tf"],
                                     "frac_of_bf16x3_ceiling": 3.0 * tf / pk["tf"]}
    fps_gpu = value / world
    extra = {
        "flop_per_frame": FLOP_PER_FRAME_TRAIN, "achieved_tflops_whole_step": fps_gpu * FLOP_PER_FRAME_TRAIN / 1e12,
        "frac_of_bf16_peak_whole_step": fps_gpu * FLOP_PER_FRAME_TRAIN / 1e12 / pk["tf"],
        "hbm_compulsory_gbs": fps_gpu * BYTES_PER_FRAME_TRAIN / 1e9,
        "frac_of_hbm_roofline_compulsory": fps_gpu * BYTES_PER_FRAME_TRAIN / 1e9 / pk["hbm_gbs"],
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nc = 1024
        fps, cores, total = time_cpu(nc, 3, 1)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "%d frames/step x 3 steps (+1 warm-up), fp32 PyTorch-CPU restatement of the reference graph" % nc}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: ConvVAE (architecture-vae-vcc2016) 10-speaker, %d frames/GPU/step (64x256), fwd+bwd+Adam" % n,
                       "frames_per_gpu_per_step": n, "global_frames_per_step": n * world, "parallelism": "dp%d" % world,
                       "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("NPVC_")},   # library A/B switches in effect ({} = defaults)
                       "l2": "per-step working set (%.1f GB activations) >> 126 MB L2; %d distinct input batches cycled"
                             % (machine.engine.handle.workspace_bytes(n, True) / 1e9, NPOOL)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
