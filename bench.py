#!/usr/bin/env python
"""Benchmark of the ConvVAE hot path (BASELINE.json metric: spectral frames/sec, ConvVAE fwd+bwd
@ [B,513,256]; the metric's configuration is cfg2 = 64 x 256 = 16,384 frames per GPU per step,
10 speakers, fwd+bwd+Adam).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (cfg2)
  python bench.py --impl reference --gpus N ...            # the reference arithmetic on the host cores
                                                           # (oracle port: TF 1.2.1 cannot run here)
  python bench.py --config cfg1|cfg2|cfg3|cfg5|b16 ...     # the other BASELINE configurations (1 GPU lines)

Prints ONE JSON line (rank 0).  A "step" = one pass of the hot path over one batch of synthetic frames:
training configs: encode -> sample -> decode -> KL + Gaussian log-density -> backward -> (all-reduce) -> Adam;
cfg3: encode -> mu -> decode (convert.py path).
"""
import argparse
import glob
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_FRAME_TRAIN = 30.35e6       # fwd 10,127,856 FLOP; fwd+bwd ~ 3x - E0 dgrad
FLOP_PER_FRAME_INFER = 10.13e6
BYTES_PER_FRAME_INFER = 4112.0       # x in + y + xh out


def bytes_per_frame_train(n):
    """SURVEY 8d compulsory HBM bytes per frame: x + y in (no eps buffer: drawn in-kernel) + the 8 x 3.76 MB
    parameter / gradient / Adam traffic of a step amortised over its frames."""
    return 2052.0 + 8.0 + 8.0 * 3756648.0 / n


CONFIGS = {
    # name: frames per GPU per step, speakers, kind, workload text (BASELINE.json configs[i])
    "cfg1": dict(frames=16 * 128, speakers=1, kind="train",
                 workload="cfg1: ConvVAE architecture-vae-vcc2016.json, 1 speaker, 2048 frames/step (16x128), fwd+bwd+Adam"),
    "cfg2": dict(frames=64 * 256, speakers=10, kind="train",
                 workload="cfg2: ConvVAE (architecture-vae-vcc2016) 10-speaker, 16384 frames/GPU/step (64x256), fwd+bwd+Adam"),
    "cfg3": dict(frames=256 * 512, speakers=10, kind="infer",
                 workload="cfg3: ConvVAE inference encode -> z (= mu) -> decode, 131072 frames/step (256x512), convert.py path"),
    "cfg5": dict(frames=32 * 256, speakers=10, kind="stacks",
                 workload="cfg5: VAWGAN (architecture-vawgan-vcc2016) conv stacks, 8192 frames/step (32x256): encoder + generator "
                          "stacks (== ConvVAE graph) fwd+bwd, then the discriminator stack (7/7/115 taps) fwd+bwd"),
    "b16": dict(frames=16, speakers=10, kind="train",
                workload="b16: the literal batch of architecture-vae-vcc2016.json:23 (16 frames/step), fwd+bwd+Adam: the "
                         "launch-bound regime the reference trains in"),
}
METRICS = {
    "train": "spectral frames/sec ConvVAE fwd+bwd+Adam @ [%s] per GPU",
    "infer": "spectral frames/sec ConvVAE inference encode->decode @ [%s] per GPU",
    "stacks": "spectral frames/sec VAWGAN conv stacks fwd+bwd @ [%s] per GPU",
}
SHAPES = {"cfg1": "16,513,128", "cfg2": "64,513,256", "cfg3": "256,513,512", "cfg5": "32,513,256", "b16": "16,513,1"}


def metric_name(cfg):
    return METRICS[CONFIGS[cfg]["kind"]] % SHAPES[cfg]


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
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")] + [time.perf_counter()])

    def mark(self):
        return time.perf_counter()

    def stop(self, t0=None, t1=None):
        """Samples taken while the workload ran (from the first warm-up step to the end of the load that follows the
        timed steps); `samples_timed` of them fall inside the timed region [t0, t1] itself."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.05)
        self.proc.terminate()
        timed = sum(1 for r in self.rows if t0 is not None and t0 <= r[-1] <= t1)
        self.rows = [r[:-1] for r in self.rows]
        self._timed = timed
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm), "samples_timed": self._timed}


# ----------------------------------------------------------------------------------------------
# CPU side: the reference arithmetic (oracle fp32 twin) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_threads():
    """Fixed rule (no per-run calibration): every host core, at most 32 -- the oneDNN / ATen CPU kernels of this graph
    stop scaling beyond that on the 513-bin frames."""
    return max(1, min(os.cpu_count() or 1, 32))


def cpu_reference_step_fn(cfg):
    """Returns step(): one pass of the oracle's fp32 twin over the SAME workload the native arm times
    (same frames per step, speakers, arithmetic); bench.py's cpu_baseline / --impl reference legs only."""
    import numpy as np
    import torch
    from oracle import convvae_ref as R
    from vae_npvc_b200 import vcc2016_vae_arch
    torch.set_num_threads(cpu_threads())
    c = CONFIGS[cfg]
    arch = vcc2016_vae_arch()
    n = c["frames"]
    P = {k: v.astype(np.float32) for k, v in R.init_params(arch, 0).items()}
    x, y, eps = R.make_inputs(arch, n, n_speakers=c["speakers"])
    x, eps = x.astype(np.float32), eps.astype(np.float32)
    st = {"theta": R.flatten_params(arch, P), "m": 0.0, "v": 0.0, "t": 0, "P": P}

    def train_step():
        out = R.forward(arch, st["P"], x, y, eps, dtype=torch.float32, with_grads=True)
        g = R.flatten_params(arch, out["grads"])
        st["t"] += 1
        st["theta"], st["m"], st["v"] = R.adam_step(st["theta"], g, st["m"], st["v"], st["t"], 1e-4, 0.5, 0.999)
        st["P"] = R.unflatten_params(arch, st["theta"].astype(np.float32))
        return float(out["G"])

    def infer_step(chunk=16384):
        for c0 in range(0, n, chunk):
            mu, _ = R.encode(arch, st["P"], x[c0:c0 + chunk], dtype=torch.float32)
            R.decode(arch, st["P"], mu, y[c0:c0 + chunk], dtype=torch.float32)

    def stacks_step():
        from vae_npvc_b200.arch import vawgan_d_stack_arch
        R.forward(arch, st["P"], x, y, eps, dtype=torch.float32, with_grads=True)
        d = st.setdefault("d", (vawgan_d_stack_arch(), {k: v.astype(np.float32) for k, v in R.init_params(vawgan_d_stack_arch(), 0).items()}))
        R.forward(d[0], d[1], x, y, eps, dtype=torch.float32, with_grads=True)
    return {"train": train_step, "infer": infer_step, "stacks": stacks_step}[c["kind"]]


def time_cpu(cfg, steps, warmup):
    step = cpu_reference_step_fn(cfg)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    return CONFIGS[cfg]["frames"] / statistics.median(ts), statistics.median(ts), sum(ts)


def run_reference(args, rank):
    """--impl reference: the reference's CPU arithmetic (TF 1.2.1 is not installable: oracle port) on the native
    arm's config -- the same frames per step, every step a full step of the workload, fixed thread count."""
    if rank != 0:
        return
    cfg = args.config
    n = CONFIGS[cfg]["frames"]
    fps, med, total = time_cpu(cfg, args.steps, max(1, min(args.warmup, 2)))
    cores = cpu_threads()
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * med, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CONFIGS[cfg]["workload"], "frames_per_gpu_per_step": n, "frames_per_step_sample": n,
                   "note": "CPU restatement of the reference graph (PyTorch/oneDNN fp32, oracle/convvae_ref.py), not TF 1.2.1 "
                           "(not installable); every step is a full %d-frame step of the workload on %d threads" % (n, cores)},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d frames/step x %d steps (the whole workload step, nothing sampled down)" % (n, args.steps)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def rank_roster(world, local):
    """Every rank's (rank, host, GPU uuid) gathered through the communicator itself: N distinct GPUs took part in a
    collective of this run (evidence that does not depend on NCCL's debug log)."""
    import socket
    import torch
    import torch.distributed as dist
    me = {"rank": dist.get_rank(), "host": socket.gethostname(), "gpu": str(torch.cuda.get_device_properties(local).uuid),
          "pci_bus": torch.cuda.get_device_properties(local).pci_bus_id}
    box = [None] * world
    dist.all_gather_object(box, me)
    return box


def nccl_evidence(world):
    """When the caller sent NCCL's debug output to files (NCCL_DEBUG_FILE), rank 0 echoes the communicator lines
    (nranks / algorithm) to stderr; by default NCCL_DEBUG=INFO already lands on stderr (see main)."""
    out = []
    pat = os.environ.get("NCCL_DEBUG_FILE", "")
    files = sorted(glob.glob(re.sub(r"%[hp]", "*", pat))) if pat else []
    for p in files:
        try:
            for ln in open(p, errors="replace"):
                if re.search(r"nranks \d+|nRanks \d+|NVLS|Connected all (rings|trees)", ln):
                    out.append(os.path.basename(p) + ": " + ln.strip())
        except OSError:
            pass
    for ln in out[:40]:
        print("[nccl] " + ln, file=sys.stderr)
    n = sorted({int(m.group(1)) for ln in out for m in [re.search(r"n[rR]anks (\d+)", ln)] if m})
    return {"nranks_seen": n, "lines": len(out), "log": pat or "stderr (NCCL_DEBUG=INFO, subsystem INIT)"}


def op_rooflines(prof, pk, steps):
    """Roofline entries from the per-op CUDA-event profile of `steps` eager steps.  Peaks: a kernel timed alone
    between events -> the burst bf16 figure and the measured copy bandwidth of MEASURED_PEAKS.json."""
    tot_ms = sum(p["ms"] for p in prof)
    traffic, step_dram = {}, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, step_dram = tj.get("bytes_per_launch", {}), tj.get("step_dram_bytes")

    def entry(p):
        ms = p["ms"] / p["calls"]
        if p.get("tensor"):
            fl = 2.0 * (p["rows"] / p["calls"]) * p["K"] * p["N"]
            a = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            return dict(kernel=p["name"], bound="tensor", achieved=a, peak=pk["tf_burst"], unit="TFLOP/s", frac=a / pk["tf_burst"],
                        frac_of_bf16x3_ceiling=3.0 * a / pk["tf_burst"], ms=ms, traffic=traffic.get(p["name"]),
                        share_of_step=p["ms"] / tot_ms if tot_ms else None)
        a = p["bytes"] / p["calls"] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return dict(kernel=p["name"], bound="hbm", achieved=a, peak=pk["hbm_gbs"], unit="GB/s", frac=a / pk["hbm_gbs"], ms=ms,
                    traffic=traffic.get(p["name"]), share_of_step=p["ms"] / tot_ms if tot_ms else None)
    top = max(prof, key=lambda p: p["ms"])
    roof = entry(top)
    roof["peak_source"] = pk["source"] + (" bf16 dense, burst (the kernel is timed alone between CUDA events)" if top.get("tensor") else " copy bandwidth")
    roof["note"] = ("fp32 path on bf16 tensor cores: every product is 3 bf16 MMAs (hi.hi + hi.lo + lo.hi), so 1/3 of the bf16 "
                    "peak is the ceiling of `frac`; algorithmic FLOPs = 2 rows K N" if top.get("tensor") else
                    "algorithmic bytes = each operand buffer of the op read / written once")
    roof["ops_ms_per_step"] = {p["name"]: round(p["ms"] / steps, 4) for p in sorted(prof, key=lambda p: -p["ms"])[:14]}
    roof["gemm_ms_share"] = sum(p["ms"] for p in prof if p["kind"] in (0, 1)) / tot_ms if tot_ms else None
    roof["step_dram_bytes"] = step_dram
    hb = [p for p in prof if not p.get("tensor") and p.get("bytes", 0) > 0]
    tn = [p for p in prof if p.get("tensor")]
    second = {}
    if hb:
        second["top_hbm_op"] = entry(max(hb, key=lambda p: p["ms"]))
    if tn:
        second["top_tensor_op"] = entry(max(tn, key=lambda p: p["ms"]))
    return roof, second


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="every training step eager (no CUDA-graph replay)")
    args = ap.parse_args()
    if args.frames:
        CONFIGS[args.config] = dict(CONFIGS[args.config], frames=args.frames)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.no_graph:
        os.environ["NPVC_GRAPH"] = "0"

    import torch
    import torch.distributed as dist
    from vae_npvc_b200 import vcc2016_vae_arch
    from importlib import import_module
    real_stdout = sys.stdout
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the native arm")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    if world > 1:
        # Rank evidence: NCCL_DEBUG=INFO (communicator lines: "... rank r nranks N ..."), left where NCCL writes it by default.
        # That default is STDOUT, which must carry ONE JSON line -- so for the whole run file descriptor 1 points at stderr
        # (NCCL's lines land there) and the JSON line is written to the saved, real stdout at the end.
        print("[nccl] environment before init: %s" % {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}, file=sys.stderr)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):      # (a caller's INFO / TRACE setting is left alone)
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        sys.stdout.flush()
        real_stdout = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        t = torch.ones(1, device=torch.device("cuda", local)); dist.all_reduce(t); torch.cuda.synchronize()
    dev = torch.device("cuda", local)
    cfg = CONFIGS[args.config]
    kind, n = cfg["kind"], cfg["frames"]
    arch = vcc2016_vae_arch()
    MODEL = getattr(import_module("model.vae"), "ConvVAE")          # the plugin lookup of main.py:39-43
    TRAINER = getattr(import_module("trainer.vae"), "VAETrainer")
    machine = MODEL(arch, device=dev, seed=0)
    machine.theta.copy_(machine.engine.init_theta(0, perturb=0.1))  # biases / LN params off 0 / 1
    g = torch.Generator(device="cpu").manual_seed(1 + rank)
    NPOOL = 4 if n * 513 * 4 * 4 < (2 << 30) else 2
    host_x = [(torch.rand(n, 513, generator=g) * 2 - 1).pin_memory() for _ in range(NPOOL)]
    host_y = [torch.randint(0, cfg["speakers"], (n,), generator=g).pin_memory() for _ in range(NPOOL)]
    dev_x = [t.to(dev) for t in host_x]; dev_y = [t.to(dev) for t in host_y]
    trainer = None
    d_machine = None
    if kind == "train":
        trainer = TRAINER(machine.loss(dev_x[0], dev_y[0]), arch, None, None)
        step_fn = trainer.opt["g"]
    elif kind == "infer":
        def step_fn(x, y):
            return machine.decode(machine.encode(x), y)
    else:
        from vae_npvc_b200.arch import vawgan_d_stack_arch
        d_machine = MODEL(vawgan_d_stack_arch(), device=dev, seed=1)
        grads = [torch.empty_like(machine.theta), torch.empty_like(d_machine.theta)]

        def step_fn(x, y):
            o = machine.loss_and_grad(x, y, grads[0])
            d_machine.loss_and_grad(x, y, grads[1])
            return o["losses"]
    engines = [machine.engine] + ([d_machine.engine] if d_machine else [])

    def launch_count():
        return sum(e.launch_count() for e in engines)

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

    # ---- launches of one step, counted on an eager step (a replayed CUDA graph re-issues exactly these) ----
    if trainer is not None:
        trainer.use_graph = False
    l0 = launch_count(); step_fn(dev_x[0], dev_y[0]); launches_per_step = launch_count() - l0
    if trainer is not None:
        trainer.use_graph = not args.no_graph and os.environ.get("NPVC_GRAPH", "1") != "0"

    # ---- device-resident arm: inputs already in HBM ------------------------------------------
    # clocks / throttle reasons are sampled (nvidia-smi, 20 ms period) from the first warm-up step on, through the timed
    # steps, the end-to-end arm, and the same load repeated for ~0.3 s AFTER both (a 20-step timed region lasts < 0.1 s: a
    # handful of samples; the follow-on load shows what the clocks settle to -- it does not precede the timed steps, which run
    # after exactly W warm-up steps as the contract says)
    sampler = ClockSampler(local); sampler.start()
    for i in range(args.warmup):
        step_fn(dev_x[i % NPOOL], dev_y[i % NPOOL])
    t0 = sampler.mark()
    ms = timed(lambda i: step_fn(dev_x[i % NPOOL], dev_y[i % NPOOL]), args.steps)
    t1 = sampler.mark()
    value = world * n * args.steps / (ms / 1000.0)
    graphed = bool(trainer is not None and trainer._state and any(q["graph"] is not None for q in trainer._state["graphs"].values()))

    # ---- end-to-end arm: pinned host inputs, H2D inside the timed region, D2H of the step's result -----------
    # every step's x / y cross PCIe inside the timed region; like analyzer.FrameLoader the copy of batch i+1 is issued
    # on a side stream while step i computes (double-buffered device slots); training reads the 3 loss scalars back,
    # inference the whole reconstructed batch (xh)
    out_host = torch.empty(3).pin_memory() if kind != "infer" else torch.empty(n, 513).pin_memory()
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
        res = step_fn(slots[k][0], slots[k][1])
        consumed[k].record(torch.cuda.current_stream())
        out_host.copy_(res.view(out_host.shape), non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the user reads the step's result
    for k in range(2):
        consumed[k].record(torch.cuda.current_stream())
    e2e_step(0); e2e_step(1); e2e_step(2)
    torch.cuda.synchronize()
    for k in range(2):
        consumed[k].record(torch.cuda.current_stream())
    ms_e2e = timed(e2e_step, args.steps)
    # every rank must run the SAME number of follow-on steps (a step contains a collective): derived from the agreed `ms`
    n_load = int(max(5, min(2000, 300.0 / max(ms / args.steps, 1e-3))))
    for i in range(n_load):
        step_fn(dev_x[i % NPOOL], dev_y[i % NPOOL])
    torch.cuda.synchronize()
    clocks = sampler.stop(t0, t1)
    e2e = world * n * args.steps / (ms_e2e / 1000.0)
    h2d = host_x[0].numel() * 4 + host_y[0].numel() * 8
    d2h = out_host.numel() * 4

    # ---- per-op device timing (eager steps, CUDA events around every op) for the rooflines ----------------
    pk = peaks()
    if trainer is not None:
        trainer.use_graph = False
    for e in engines:
        e.handle.profile_enable(True)
    PS = min(args.steps, 5)
    for i in range(PS):
        step_fn(dev_x[i % NPOOL], dev_y[i % NPOOL])
    torch.cuda.synchronize()
    prof = []
    for e in engines:
        prof += e.handle.profile(); e.handle.profile_enable(False)
    roofline, second = op_rooflines(prof, pk, PS)
    fps_gpu = value / world
    if kind == "infer":
        flop, byts = FLOP_PER_FRAME_INFER, BYTES_PER_FRAME_INFER
    else:
        flop, byts = FLOP_PER_FRAME_TRAIN, bytes_per_frame_train(n)
    extra = dict(second)
    extra.update({
        "flop_per_frame": flop, "achieved_tflops_whole_step": fps_gpu * flop / 1e12,
        "frac_of_bf16_peak_whole_step": fps_gpu * flop / 1e12 / pk["tf"],
        "hbm_compulsory_bytes_per_frame": byts, "hbm_compulsory_gbs": fps_gpu * byts / 1e9,
        "frac_of_hbm_roofline_compulsory": fps_gpu * byts / 1e9 / pk["hbm_gbs"],
    })
    if roofline.get("step_dram_bytes"):
        # measured DRAM traffic of one step (ncu, profiles/ncu_traffic.json) at this run's step time: the achieved HBM rate
        extra["achieved_hbm_gbs_whole_step"] = roofline["step_dram_bytes"] / (ms / args.steps * 1e-3) / 1e9
        extra["frac_of_hbm_peak_whole_step"] = extra["achieved_hbm_gbs_whole_step"] / pk["hbm_gbs"]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, med, total = time_cpu(args.config, 3, 1)
        cpu = {"value": fps, "unit": "frames/s", "cores": cpu_threads(), "kind": "port",
               "sample": "%d frames/step x 3 steps (+1 warm-up): the whole workload step, fp32 PyTorch-CPU restatement of the "
                         "reference graph (oracle/convvae_ref.py)" % n}
    roster = rank_roster(world, local) if world > 1 else None
    nccl = nccl_evidence(world) if (rank == 0 and world > 1) else None
    if nccl is not None:
        nccl["ranks"] = roster
        nccl["distinct_gpus"] = len({(r["host"], r["gpu"]) for r in roster})
        print("[nccl] %d ranks on %d distinct GPUs: %s" % (world, nccl["distinct_gpus"], roster), file=sys.stderr)

    if rank == 0:
        line = {
            "metric": metric_name(args.config), "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 storage and accumulation, bf16x3 tensor-core products (hi.hi + hi.lo + lo.hi: ~2^-17 relative per product)",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "frames_per_gpu_per_step": n, "global_frames_per_step": n * world,
                       "parallelism": "dp%d" % world, "cuda_graph": graphed,
                       "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("NPVC_")},   # library A/B switches in effect ({} = defaults)
                       "l2": "per-step working set (%.1f GB workspace) >> 126 MB L2; %d distinct input batches cycled"
                             % (machine.engine.handle.workspace_bytes(n, kind != "infer") / 1e9, NPOOL)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
        }
        if nccl:
            line["nccl"] = nccl
        real_stdout.write(json.dumps(line) + "\n"); real_stdout.flush()
    if world > 1:
        # CUDA graphs that captured NCCL kernels go before the communicator; a communicator teardown that does not
        # return (seen once with captured collectives) must not hang the run: the line is out, leave after 20 s
        if trainer is not None and trainer._state:
            trainer._state["graphs"].clear()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        th = threading.Thread(target=dist.destroy_process_group, daemon=True)
        th.start(); th.join(20.0)
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
