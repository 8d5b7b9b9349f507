"""bench.py — Palette UNet 256x256 training-step throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32] [--size 256]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.

A "step" = set_input + optimize_parameters() = noising prologue + UNet forward + eps-loss + backward
(dgrad + wgrad) + gradient all-reduce + fused AdamW + EMA, on BASELINE config 2 (Palette UNet ngf 64,
mults 1-2-4-8, 2 res blocks / level, attention at the 32x32 middle block, 256x256, per-GPU batch 32,
bf16 activations / fp32 accumulate, synthetic self-supervised box masks, random-init de-zeroed weights).

  value  : images/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e    : images/s through the public API with HOST (pinned) batches: H2D copy of A, B, mask and a D2H
           read of the loss inside every timed step
  roofline: all implicit-GEMM (tcgen05) conv launches of a step, CUDA-event timed per launch in an
           instrumented pass right after the timed region: algorithmic conv FLOPs / their summed duration
  cpu_baseline / --impl reference: the oracle port of the reference's CPU path (the Python reference
           cannot travel to the GPU box) on all host cores, bounded sample.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_GMAC_PER_IMG_256 = 206.63  # Palette UNet 256^2 forward; train step = 3x
METRIC = "train-step images/sec Palette UNet 256^2 bf16"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of CUDA-graph replay")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "hbm_gbs": d.get("hbm_gbs"),
                "source": "MEASURED_PEAKS.json (sustained cuBLAS bf16)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def usable_cpu_threads():
    """Host threads this process can really run on: scheduler affinity capped by the cgroup CPU quota (a container
    that sees 128 logical CPUs but is throttled to a few cores makes a 128-thread torch run pathologically slow)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, q // period))
            break
        except Exception:
            continue
    return n


def make_cfg(O, size):
    return O.UNetCfg(image_size=size)


def cpu_reference_throughput(size, steps, threads=None, budget_s=40.0):
    """Oracle port of the reference's CPU training step (fp32, torch CPU kernels) on the host cores.
    Bounded sample: thread pools are warmed on one 64x64 step, then up to `steps` steps of batch 1 at the
    full resolution are timed, stopping once `budget_s` seconds are spent (at least one step)."""
    from oracle import palette_oracle as O
    if threads is None:
        threads = int(os.environ.get("JG_CPU_THREADS", "0")) or min(usable_cpu_threads(), torch.get_num_threads())
    torch.set_num_threads(threads)
    oc = O.OptimCfg(lr=1e-4)
    b = 1

    def one_step(state, cfg, sz, seed):
        data = O.synthetic_batch(b, sz, 77 + seed)
        torch.manual_seed(seed)
        t, u = O.sample_t_gamma(cfg, b)
        noise = torch.randn_like(data["gt"])
        t0 = time.perf_counter()
        O.train_step(state, cfg, oc, data["gt"], data["cond"], data["mask"], noise, t, u)
        return time.perf_counter() - t0

    cfg = make_cfg(O, size)
    state = O.TrainState(params=O.init_params(cfg, 1234))
    one_step(state, make_cfg(O, 64), 64, 0)  # warm-up (same weights, small crop)
    times = []
    for s in range(steps):
        times.append(one_step(state, cfg, size, 1 + s))
        if sum(times) > budget_s:
            break
    sec = sum(times) / len(times)
    return {"value": b / sec, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": "%d step(s) of batch %d at %dx%d (fwd+loss+bwd+AdamW+EMA), fp32 torch-CPU oracle port, "
                      "%d threads of %d logical CPUs, %.1f s/step" % (len(times), b, size, size, threads,
                                                                      os.cpu_count() or 0, sec)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_throughput(args.size, max(1, min(args.steps, args.cpu_steps)))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / cb["value"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "palette_model Palette UNet %dx%d bf16 batch=%d/GPU, synthetic self-supervised masks"
                               % (args.size, args.size, args.batch),
                   "sample": "reference arithmetic (fp32 torch-CPU oracle port) on the host cores: " + cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def conv_traffic(n_launches):
    """DRAM bytes (read + write) per implicit-GEMM call, averaged over the calls of one step, from the committed
    ncu pass over one training step (tools/gpu_step_once.py + tools/ncu_traffic.py); None if the capture is missing or
    was taken on a different number of launches."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_conv_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
    except (OSError, ValueError):
        return None
    # the capture counts KERNELS (a 3x3 wgrad with Cout >= 128 is two), n_launches counts C-ABI calls
    if not (n_launches <= d.get("launches", 0) <= 1.5 * n_launches):
        return None
    return d["dram_bytes_total"] / n_launches


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from joligen_b200 import lib as L
    from joligen_b200 import nets
    from joligen_b200 import synthetic
    from joligen_b200.trainer import PaletteTrainer

    L.load()
    assert L.load().jg_check_device() == 0, L.load().jg_last_error()
    net = nets.build_palette_generator(image_size=args.size)
    synthetic.dezero_init_(net, 1234)
    tr = PaletteTrainer(net, lr=1e-4, optim="adamw", ema=True, ema_beta=0.999, device="cuda",
                        cuda_graph=not args.no_graph, graph_warmup=2)
    tr.broadcast_parameters()

    B = args.batch
    host = synthetic.synthetic_batch(B, args.size, 1234 + rank)
    host = {k: v.pin_memory() for k, v in host.items()}
    dev = {k: v.cuda() for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(steps, from_host):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        last = None
        for _ in range(steps):
            tr.set_input(host if from_host else dev)
            loss = tr.optimize_parameters()
            if from_host:
                last = float(loss)  # D2H read of the step's loss
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps, last

    run(max(args.warmup, 3), False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.launch_count[0] = 0
    ms, _ = run(args.steps, False)
    launches = L.launch_count[0]
    clocks = sampler.stop() if rank == 0 else None
    run(1, True)
    ms_e2e, last_loss = run(args.steps, True)

    # instrumented pass: CUDA events around every C-ABI call; conv launches carry their FLOPs
    recs = []
    allrecs = []

    @contextlib.contextmanager
    def hook(name, cargs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        if name in ("jg_conv2d_fwd", "jg_conv2d_wgrad", "jg_conv2d_wgrad_acc"):
            d = cargs[0]._obj
            flops = 2.0 * d.N * d.Ho * d.Wo * d.Cout * d.Cin * d.R * d.S
            recs.append((name, flops, e0, e1))
            allrecs.append((name, (d.N, d.H, d.W, d.Cin, d.Cout, d.R, d.stride), flops, e0, e1))
        elif name == "jg_groupnorm_fwd":
            allrecs.append((name, tuple(int(v) for v in cargs[4:8]), 0.0, e0, e1))
        elif name == "jg_groupnorm_bwd":
            allrecs.append((name, tuple(int(v) for v in cargs[7:11]), 0.0, e0, e1))
        else:
            allrecs.append((name, None, 0.0, e0, e1))

    L.call_hook[0] = hook
    tr.set_input(dev)
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_a.record()
    tr._forward_backward()  # eager launches (the timed region above replays the same kernels from CUDA graphs)
    tr._optimizer_step()
    ev_b.record()
    torch.cuda.synchronize()
    L.call_hook[0] = None
    if rank == 0 and os.environ.get("JG_BREAKDOWN"):
        agg = {}
        for name, shape, flops, e0, e1 in allrecs:
            key = name if shape is None else "%s %s" % (name, list(shape))
            a = agg.setdefault(key, {"calls": 0, "ms": 0.0, "flops": 0.0})
            a["calls"] += 1
            a["ms"] += e0.elapsed_time(e1)
            a["flops"] += flops
        rows = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
        with open(os.environ["JG_BREAKDOWN"], "w") as f:
            json.dump({"instrumented_step_ms": ev_a.elapsed_time(ev_b),
                       "sum_call_ms": sum(v["ms"] for _, v in rows),
                       "rows": [dict(key=k, tflops=(v["flops"] / v["ms"] / 1e9 if v["ms"] > 0 else 0), **v)
                                for k, v in rows]}, f, indent=1)
    conv_ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in recs)
    conv_flops_padded = sum(f for _, f, _, _ in recs)
    alg_flops_step = 3 * 2 * FWD_GMAC_PER_IMG_256 * 1e9 * B * (args.size / 256.0) ** 2
    peaks = measured_peaks()

    if rank == 0:
        imgs = B * world
        value = imgs / (ms / 1000.0)
        achieved = alg_flops_step / (conv_ms / 1000.0) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "palette_model Palette UNet %dx%d bf16 batch=%d/GPU, synthetic self-supervised masks"
                                   % (args.size, args.size, B),
                       "global_batch": imgs, "parallelism": "dp%d" % world,
                       "l2": "per-step activations (>20 GB) exceed the 126 MB L2; no explicit flush",
                       "optimizer": "fused AdamW+EMA", "loss_last": last_loss,
                       "launch": "eager" if args.no_graph else "CUDA graph replay (fwd+bwd graph, optimizer graph)"},
            "e2e": {"value": imgs / (ms_e2e / 1000.0), "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "clocks": clocks,
            "step_tflops_algorithmic": alg_flops_step / (ms / 1000.0) / 1e12,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops"], "traffic": conv_traffic(len(recs)),
                         "kernel": "conv_halo / conv_fwd / wgrad_halo / conv_wgrad kernels (all %d implicit-GEMM calls of one step)"
                                   % len(recs),
                         "conv_ms_per_step": conv_ms, "conv_share_of_step": conv_ms / ms,
                         "padded_flops_over_algorithmic": conv_flops_padded / alg_flops_step,
                         "peak_source": peaks["source"]},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_throughput(args.size, args.cpu_steps)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
