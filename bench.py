#!/usr/bin/env python
"""bench.py -- MultiRes-UNet 256x256 tiles/sec, full training step (fwd + weighted-BCE + bwd + Adam).

    python bench.py --gpus N --steps K --warmup W            (driver: torchrun for N>1)
    python bench.py --impl reference ...                     (the CPU arm: oracle on the host cores)

Workload = BASELINE.json configs[1]: "UNet training 256x256x1 tiles bf16 batch 32 on 1xB200"; with N GPUs every
rank trains batch 32 (weak scaling) and gradients are summed with ONE NCCL all-reduce of the flat gradient buffer.
Prints ONE JSON line on rank 0 (contract in the task statement).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_TILE_FWD_BWD = {256: 31.669e9, 512: 126.676e9}     # SURVEY.md 8d / BASELINE.md section 2 (true channel counts)
METRIC = "UNet 256x256 tiles/sec fwd+bwd"


def load_peaks():
    """Roofline denominators: MEASURED_PEAKS.json when the driver wrote it (key names are matched loosely), else the
    fallback stated in B200_PROFILING.md."""
    fb = {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if not os.path.exists(p):
        return fb
    try:
        with open(p) as fh:
            d = json.load(fh)

        def flat(obj, prefix=""):
            out = {}
            if isinstance(obj, dict):
                for k, v in obj.items():
                    out.update(flat(v, f"{prefix}{k}.".lower()))
            elif isinstance(obj, (int, float)) and not isinstance(obj, bool):
                out[prefix.rstrip(".")] = float(obj)
            return out

        f = flat(d)

        def pick(*needles, avoid=()):
            for k, v in f.items():
                if all(n in k for n in needles) and not any(a in k for a in avoid) and v > 0:
                    return v
            return None

        hbm = pick("hbm", "gb") or pick("hbm") or pick("copy", "gb")
        sus = pick("bf16", "sustain") or pick("tflop", "sustain") or pick("sustain")
        burst = pick("bf16", "tflop", avoid=("sustain",)) or pick("bf16", avoid=("sustain",)) or pick("tflop", avoid=("sustain",))
        if hbm and hbm < 100:          # TB/s
            hbm *= 1000.0
        if burst and burst > 100000:    # GFLOP/s
            burst /= 1000.0
        if sus and sus > 100000:
            sus /= 1000.0
        if not hbm or not burst:
            return dict(fb, source="fallback (B200_PROFILING.md); MEASURED_PEAKS.json present but not understood")
        return {"hbm_gbs": hbm, "tflops_burst": burst, "tflops_sustained": sus or burst, "source": "measured (MEASURED_PEAKS.json)"}
    except Exception as exc:       # never let the peaks file break the bench
        return dict(fb, source=f"fallback (B200_PROFILING.md); MEASURED_PEAKS.json unreadable: {exc}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


def finish_distributed(world, holders):
    """Orderly end of a multi-rank run: CUDA graphs that captured NCCL kernels are released BEFORE the communicator, all
    ranks leave together, and a watchdog ends the process if the teardown itself blocks (seen once with captured
    collectives): the JSON line is already out at this point."""
    if world <= 1:
        return 0
    import torch.distributed as dist
    threading.Timer(45.0, lambda: os._exit(0)).start()
    for h in holders:
        for attr in ("graph", "graph2", "_graphs"):
            if hasattr(h, attr):
                setattr(h, attr, None)
    torch.cuda.synchronize()
    try:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    finally:
        os._exit(0)


# ------------------------------------------------------------------------------------------------- CPU arm
def cpu_train_throughput(size: int, batch: int, steps: int, warmup: int):
    """Oracle (torch-CPU restatement of the Keras-torch path) train step on the host cores; tiles/s."""
    from oracle import unet as OU
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = OU.UNetSpec(16)
    x, y, wgt = OU.synthetic_batch(batch, size, size)
    tr = OU.UNetTrainer(spec, spec.init_params(0), wgt)
    for _ in range(warmup):
        tr.train_step(x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.train_step(x, y)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    batch = args.cpu_batch
    steps, warmup = min(args.steps, 4), min(args.warmup, 1)
    tps, spt, cores = cpu_train_throughput(args.size, batch, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": "tiles/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": spt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"MultiRes-UNet filters=16 {args.size}x{args.size}x1 full train step (fwd+wBCE+bwd+Adam)",
                   "sample": f"batch {batch} per step on the host cores (the GPU arm runs batch {args.batch})"},
        "cpu_baseline": {"value": tps, "unit": "tiles/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps of batch {batch}, oracle/ (torch {torch.__version__} CPU fp32, ATen/oneDNN; Keras is not installable here)"},
        "e2e": {"value": tps, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------- GPU arm
def profile_ops(inst, weighting, peaks, size, batch):
    """One eager fwd+bwd with CUDA events around every engine op (weight gradients timed separately from the data
    gradients); returns one row per (op, kernel family) with its time and ALGORITHMIC bytes / flops."""
    import sem_b200  # noqa: F401
    from sem_b200.engine import ConvOp, AffineOp
    e = inst.eng
    recs = []
    side, e.wgrad_stream = e.wgrad_stream, None        # per-op events need every kernel on the timed stream
    wg = {}

    def timed(label, op, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        recs.append((label, op, a, b))

    orig_side = e.on_wgrad_stream

    def timed_wgrad(fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        wg["last"] = (a, b)

    e.on_wgrad_stream = timed_wgrad
    e.zero_step(zero_grads=True)
    inst.stage_in()
    # The host needs ~20-40 us per op (ctypes call + two event records), more than most kernels run: without a head start the
    # device idles between ops and every event pair also measures the launch latency of its kernel (round 1/2: the eager
    # sum was 18.3 ms for a 14.4 ms serialised graph step, small kernels were overstated by 30-100 %).  A spin kernel keeps
    # the device busy until the whole forward + backward is queued, so the events see back-to-back execution.
    # (one spin per pass: the launch queue holds ~1k commands)
    torch.cuda._sleep(int(0.040 * 1.9e9))
    for op in e.ops:
        timed("fwd", op, lambda op=op: op.fwd(True))
    inst.loss(weighting, with_grad=True)
    torch.cuda._sleep(int(0.040 * 1.9e9))
    for op in reversed(e.ops):
        wg.pop("last", None)
        timed("bwd", op, lambda op=op: op.bwd())
        if "last" in wg:
            recs.append(("wgrad", op, *wg["last"]))
    torch.cuda.synchronize()
    e.on_wgrad_stream = orig_side
    e.wgrad_stream = side
    esz = 2 if e.dtype_name == "bf16" else 4
    rows = []
    last_bwd = None
    for label, op, a, b in recs:
        ms = a.elapsed_time(b)
        row = {"phase": label, "op": type(op).__name__, "ms": ms, "kernel": "other"}
        if isinstance(op, ConvOp):
            g = op.geom
            pix = g.N * g.OH * g.OW
            # physical (8-padded) channels; a merged res_path conv counts the work of its two reference layers (alg_flops)
            flops = getattr(op, "alg_flops", None) or 2.0 * pix * g.R * g.S * g.Cin * g.Cout
            nbytes = (g.N * g.H * g.W * g.Cin + pix * g.Cout) * esz             # input + output activations once
            tc = op.use_tc and e.dtype_name == "bf16"
            row.update({"geom": f"{g.H}x{g.W} {g.Cin}->{g.Cout} k{g.R} s{g.stride}{' T' if op.transposed else ''}",
                        "flops": flops, "bytes": nbytes, "w": op.w})
            if label == "fwd":
                row["kernel"] = "conv_tma_kernel (fwd+dgrad)" if tc else "conv_simt"
            elif label == "wgrad":
                row["kernel"] = ("wgrad_tma_kernel" if g.R == 3 else "wgrad_tc_kernel") if tc else "wgrad_simt"
                last_bwd["ms"] = max(last_bwd["ms"] - ms, 0.0)                  # what is left of the op is the data gradient
            else:
                row["kernel"] = "conv_tma_kernel (fwd+dgrad)" if tc else "conv_simt"
                if not op.x.requires_grad:
                    row["flops"], row["bytes"] = 0.0, 0
                last_bwd = row
        if isinstance(op, AffineOp):
            hb = 1 if op.b is not None else 0
            elems = op.n * op.hw * op.a.C
            nt = (2 + hb) if label == "fwd" else (2 + hb) + (3 + 2 * hb)       # reduce reads dy,a[,b]; apply reads the same, writes da[,db]
            row.update({"geom": f"C={op.a.C} hw={op.hw} two_operands={bool(hb)}", "bytes": elems * esz * nt,
                        "kernel": "affine_act_fwd_kernel" if label == "fwd" else "affine_act_bwd_{reduce,apply}_kernel"})
        rows.append(row)
    return rows


def kernel_table(rows, peaks):
    """Per kernel family: time share of the (eager) step, achieved algorithmic GB/s and TFLOP/s, roofline bound."""
    total = sum(r["ms"] for r in rows)
    fam = {}
    for r in rows:
        f = fam.setdefault(r["kernel"], {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "launch_groups": 0})
        f["ms"] += r["ms"]
        f["bytes"] += r.get("bytes", 0)
        f["flops"] += r.get("flops", 0.0)
        f["launch_groups"] += 1
    ridge = peaks["tflops_burst"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    out = {}
    for k, f in fam.items():
        if f["ms"] <= 0:
            continue
        gbs = f["bytes"] / (f["ms"] * 1e-3) / 1e9
        tfs = f["flops"] / (f["ms"] * 1e-3) / 1e12
        ai = f["flops"] / f["bytes"] if f["bytes"] else 0.0
        out[k] = {"ms": f["ms"], "share_of_step": f["ms"] / total, "algorithmic_GBps": gbs, "TFLOPs": tfs,
                  "arithmetic_intensity_flop_per_byte": ai, "bound": "tensor" if ai >= ridge else "hbm",
                  "frac_of_hbm_peak": gbs / peaks["hbm_gbs"], "frac_of_tensor_peak": tfs / peaks["tflops_burst"],
                  "ops": f["launch_groups"], "bytes_per_step": f["bytes"]}
    return out, total


def run_ours(args):
    import torch.distributed as dist
    import sem_b200
    from sem_b200 import UNetModel, _lib as L
    from oracle import unet as OU   # only for the synthetic inputs definition and the cpu_baseline leg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks = load_peaks()
    size, batch = args.size, args.batch

    x, y, wgt = OU.synthetic_batch(batch, size, size)
    if world > 1:   # each rank gets its own tiles
        g = torch.Generator().manual_seed(1000 + rank)
        x = torch.rand(x.shape, generator=g)
    m = UNetModel((size, size, 1), 16, dtype=args.dtype, batch_size=batch, seed=0, use_cuda_graph=not args.no_graph)
    m.compile(weighting=wgt, learning_rate=1e-3)
    if world > 1:
        m.set_distributed()
    inst = m._current
    # host inputs live in pinned memory (the contract's e2e definition): train_step uploads them without a staging memcpy
    xn, yn = x.contiguous().pin_memory(), y.contiguous().pin_memory()

    # ---- parity gate of the benched path: first-step loss on THIS batch against the oracle (same initial weights)
    first_step = None
    if rank == 0 and world == 1 and not args.no_check:
        spec = OU.UNetSpec(16)
        p0 = spec.init_params(seed=0)
        m.set_named_weights({k: v.detach().numpy() for k, v in p0.items()})
        chk = OU.UNetTrainer(spec, p0, wgt)
        torch.set_num_threads(os.cpu_count() or 1)
        ref0, _ = chk.train_step(x, y)
        got0 = dict(m.train_step(xn, yn))
        rel = abs(got0["loss"] - ref0["loss"]) / abs(ref0["loss"])
        first_step = {"loss": got0["loss"], "oracle_loss": ref0["loss"], "rel_err": rel, "acc": got0["acc"], "oracle_acc": ref0["acc"],
                      "tolerance": 3e-2 if args.dtype == "bf16" else 1e-3}
        assert rel < first_step["tolerance"], f"first-step loss {got0['loss']} differs from the oracle's {ref0['loss']} (rel {rel:.3e})"
        del chk

    # launches per step, counted on an eager step (graph replays do not pass through the host-side counter)
    saved_graph = m.use_cuda_graph
    m.use_cuda_graph = False
    c0 = L.launch_count()
    m.train_step(xn, yn)
    launches_per_step = L.launch_count() - c0
    m.use_cuda_graph = saved_graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (`value`)
    for _ in range(max(args.warmup, 3)):
        m.train_step(xn, yn)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        m._step_device(inst)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    # ---- end-to-end timing through the public API with HOST buffers (`e2e`)
    barrier()
    t0 = time.perf_counter()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    last = None
    for _ in range(args.steps):
        last = m.train_step(xn, yn)
    ev3.record()
    barrier()
    e2e_ms = max(ev2.elapsed_time(ev3), (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    tiles = batch * world * args.steps
    value = tiles / (dev_ms / 1e3)
    e2e_value = tiles / (e2e_ms / 1e3)

    if rank != 0:
        return finish_distributed(world, [inst])

    # ---- per-op profile (eager, CUDA events on the launching stream) -> roofline of the dominant kernel
    rows = profile_ops(inst, wgt, peaks, size, batch)
    table, total_ms = kernel_table(rows, peaks)
    named = {k: v for k, v in table.items() if k != "other"}
    top_name = max(named, key=lambda k: named[k]["ms"])
    top = named[top_name]
    if top["bound"] == "tensor":
        roof = {"bound": "tensor", "achieved": top["TFLOPs"], "peak": peaks["tflops_burst"], "unit": "TFLOP/s"}
    else:
        roof = {"bound": "hbm", "achieved": top["algorithmic_GBps"], "peak": peaks["hbm_gbs"], "unit": "GB/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["kernel"] = top_name
    roof["kernel_ms_per_step"] = top["ms"]
    roof["kernel_share_of_step"] = top["share_of_step"]
    roof["launches_per_step"] = top["ops"]
    roof["arithmetic_intensity_flop_per_byte"] = top["arithmetic_intensity_flop_per_byte"]
    roof["algorithmic_bytes_per_launch"] = top["bytes_per_step"] / max(top["ops"], 1)
    roof["avg_launch_ms"] = top["ms"] / max(top["ops"], 1)
    roof["peak_source"] = peaks["source"] + ", burst figure (kernels timed one by one)"
    # DRAM traffic of that kernel from the committed `ncu --set full` capture (profiles/traffic.json), per launch
    roof["traffic"] = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        ent = tj.get(top_name.split(" ")[0])
        if ent:
            roof["traffic"] = ent["dram_bytes_per_launch"]
            roof["traffic_source"] = ent.get("source")
    net_tflops = value / world * FLOP_PER_TILE_FWD_BWD.get(size, 31.669e9 * (size / 256) ** 2) / 1e12
    roof["net_conv_tflops"] = net_tflops
    roof["net_frac_of_tensor_peak_sustained"] = net_tflops / peaks["tflops_sustained"]
    # the "fused 3x3 stage" of SURVEY 8d: the three chained 3x3 convs of mres6 (32x32) and mres7 (64x64), forward
    names = {f"conv2d_{i}/kernel" for i in (42, 43, 44, 46, 47, 48)}       # creation order: 41 / 45 are the 1x1 shortcuts
    stage = [r for r in rows if r["phase"] == "fwd" and r.get("w") in names]
    if stage:
        sf, st = sum(r["flops"] for r in stage), sum(r["ms"] for r in stage)
        roof["stage_3x3_mres6_mres7_fwd"] = {"TFLOPs": sf / (st * 1e-3) / 1e12, "frac_of_tensor_peak": sf / (st * 1e-3) / 1e12 / peaks["tflops_burst"],
                                             "convs": len(stage), "ms": st}

    # ---- CPU baseline beside it (bounded sample)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        tps, spt, cores = cpu_train_throughput(size, args.cpu_batch, 2, 1)
        cpu = {"value": tps, "unit": "tiles/s", "cores": cores, "kind": "port",
               "sample": f"2 steps of batch {args.cpu_batch} after 1 warm-up, oracle/ train step (torch CPU fp32)"}

    esz = 4
    line = {
        "metric": METRIC, "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"MultiRes-UNet filters=16 {size}x{size}x1 full train step (fwd+wBCE+bwd+Adam), batch {batch} per GPU",
                   "global_batch": batch * world, "parallelism": f"dp{world}", "cuda_graph": not args.no_graph,
                   "l2": "per-step working set (activations+gradients, several GB) exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_value, "unit": "tiles/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(2 * batch * size * size * esz), "d2h_bytes_per_step": 16},
        "gpu_launches": int(launches_per_step * args.steps),
        "launches_per_step": int(launches_per_step),
        "roofline": roof,
        "roofline_by_kernel": table,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "last_step_metrics": dict(last),
        "first_step_vs_oracle": first_step,
    }
    emit(line)
    if args.profile_out:
        with open(args.profile_out, "w") as fh:
            json.dump({"total_ms_eager": total_ms, "rows": rows}, fh, indent=1)
    return finish_distributed(world, [inst])


# ------------------------------------------------------------------------------------------------- CycleGAN (configs[2])
CG_METRIC = "CycleGAN 256x256 image pairs/sec, full train step (G_A2B/G_B2A + PatchGAN-D, image pools)"
CG_GFLOP_PER_PAIR = 1967.0      # algorithmic minimum at filters=64, 256x256 (6 G fwd + 12 bwd-eq, 6 D fwd + 10 bwd-eq)


def cg_inputs(batch, size):
    g = torch.Generator().manual_seed(0)
    a = torch.rand(batch, size, size, 1, generator=g) * 2 - 1
    b = torch.rand(batch, size, size, 1, generator=torch.Generator().manual_seed(1)) * 2 - 1
    return a, b


def cg_cpu_throughput(size, filters, batch, steps, warmup):
    import random
    from oracle import cyclegan as OC
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr = OC.CycleGanTrainer(filters=filters, pool_batch=batch, seed=0)
    a, b = cg_inputs(batch, size)
    for _ in range(warmup):
        tr.train_step(a, b)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.train_step(a, b)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, cores


def run_cyclegan_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    batch = 1
    steps, warmup = min(args.steps, 2), min(args.warmup, 1)
    pps, spt, cores = cg_cpu_throughput(args.size, args.filters, batch, steps, warmup)
    emit({"impl": "reference", "metric": CG_METRIC, "value": pps, "unit": "image pairs/s", "n_gpus": args.gpus, "steps": steps,
          "warmup": warmup, "ms_per_step": spt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
          "data": "synthetic",
          "config": {"workload": f"CycleGAN filters={args.filters} {args.size}x{args.size}x1 train_step_torch", "sample": f"batch {batch} per step on the host cores (the GPU arm runs batch {args.batch})"},
          "cpu_baseline": {"value": pps, "unit": "image pairs/s", "cores": cores, "kind": "port",
                           "sample": f"{steps} steps of batch {batch}, oracle/cyclegan.py (torch {torch.__version__} CPU fp32)"},
          "e2e": {"value": pps, "unit": "image pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
    return 0


def instrument(engines):
    """Wraps fwd / bwd of every op of `engines` (and their weight-gradient launches) with CUDA events; returns the record
    list [(phase, op, ev0, ev1)] filled by the next eager step and an undo function."""
    recs, undo = [], []
    for e in engines:
        wg = {}

        def timed_wgrad(fn, e=e, wg=wg):
            if e.skip_wgrad:
                return None
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            wg["last"] = (a, b)

        undo.append((e, "on_wgrad_stream", e.__dict__.get("on_wgrad_stream")))
        e.on_wgrad_stream = timed_wgrad
        for op in e.ops:
            for phase, name in (("fwd", "fwd"), ("bwd", "bwd")):
                orig = getattr(op, name)

                def wrapped(*a, _orig=orig, _op=op, _phase=phase, _wg=wg, **k):
                    _wg.pop("last", None)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); r = _orig(*a, **k); e1.record()
                    recs.append((_phase, _op, e0, e1))
                    if "last" in _wg:
                        recs.append(("wgrad", _op, *_wg.pop("last")))
                    return r
                undo.append((op, name, op.__dict__.get(name)))
                setattr(op, name, wrapped)

    def restore():
        for obj, name, prev in undo:
            if prev is None:
                obj.__dict__.pop(name, None)
            else:
                setattr(obj, name, prev)
    return recs, restore


def cg_rows(recs, dtype_name):
    from sem_b200.engine import ConvOp, AffineOp
    esz = 2 if dtype_name == "bf16" else 4
    rows, last_bwd = [], None
    for label, op, a, b in recs:
        ms = a.elapsed_time(b)
        row = {"phase": label, "op": type(op).__name__, "ms": ms, "kernel": "other"}
        if isinstance(op, ConvOp):
            g = op.geom
            pix = g.N * g.OH * g.OW
            row.update({"geom": f"{g.H}x{g.W} {g.Cin}->{g.Cout} k{g.R} s{g.stride}{' T' if op.transposed else ''}",
                        "flops": 2.0 * pix * g.R * g.S * g.Cin * g.Cout, "bytes": (g.N * g.H * g.W * g.Cin + pix * g.Cout) * esz, "w": op.w})
            tc = (op.use_tc or op.s2d is not None or op.tapfold is not None) and dtype_name == "bf16"
            big = g.R == 3 and g.stride == 1 and g.Cin >= 256
            if label == "wgrad":
                row["kernel"] = ("wgrad_tma_kernel [3x3 s1 Cin>=256]" if big else "wgrad (other layers)") if tc else "wgrad_simt"
                if last_bwd is not None:
                    last_bwd["ms"] = max(last_bwd["ms"] - ms, 0.0)
            else:
                row["kernel"] = ("conv_tma_kernel [3x3 s1 Cin>=256, fwd+dgrad]" if big else "conv fwd+dgrad (other layers)") if tc else "conv_simt"
                if label == "bwd":
                    if not op.x.requires_grad:
                        row["flops"], row["bytes"] = 0.0, 0
                    last_bwd = row
        elif isinstance(op, AffineOp):
            row["kernel"] = "affine_act_*"
            row["geom"] = f"C={op.a.C} hw={op.hw} n={op.n} two_operands={op.b is not None}"
        rows.append(row)
    return rows


def run_cyclegan(args):
    import random
    import torch.distributed as dist
    import sem_b200
    from sem_b200 import CycleGanModel, ImagePool, _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks = load_peaks()
    size, n, F = args.size, args.batch, args.filters
    m = CycleGanModel((size, size, 1), batch_size=n, filters=F, dtype=args.dtype,
                      image_pool_a=ImagePool(n, 50, random.Random(rank)), image_pool_b=ImagePool(n, 50, random.Random(100 + rank)),
                      use_cuda_graph=not args.no_graph)
    m.compile()
    if world > 1:
        m.set_distributed()
    a, b = cg_inputs(n, size)
    if world > 1:
        a = torch.rand(a.shape, generator=torch.Generator().manual_seed(1000 + rank)) * 2 - 1
    ap_, bp_ = a.contiguous().pin_memory(), b.contiguous().pin_memory()

    c0 = L.launch_count()
    m.train_step((ap_, bp_))                 # eager first step: counts the launches, allocates gradient buffers
    launches_per_step = L.launch_count() - c0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up runs until both image pools are full (50 entries, `n` per step): the fill -> swap transition of ImagePool.query
    # changes the host work of a step and must not fall into the timed region
    warm = max(args.warmup, 3, -(-50 // n) + 1)
    for _ in range(warm):
        m.train_step((ap_, bp_))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        m.step_device()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    barrier()
    t0 = time.perf_counter()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    last = None
    for _ in range(args.steps):
        last = m.train_step((ap_, bp_))
    ev3.record()
    barrier()
    e2e_ms = max(ev2.elapsed_time(ev3), (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    pairs = n * world * args.steps
    value, e2e_value = pairs / (dev_ms / 1e3), pairs / (e2e_ms / 1e3)
    if rank != 0:
        return finish_distributed(world, [m])

    # ---- per-op profile of one eager step (CUDA events around every op of all twelve towers)
    engines = [bb.e for bb in (m.GA_ra, m.GB_rb, m.GB_fb, m.GA_fa, m.GB_ra, m.GA_rb, m.DA_fa, m.DB_fb, m.DA_real, m.DA_pool, m.DB_real, m.DB_pool)]
    recs, restore = instrument(engines)
    saved, saved_world = m.use_cuda_graph, m.world_size
    m.use_cuda_graph = False
    m.world_size = 1           # rank 0 profiles alone: the other ranks have left, a collective here would wait for them forever
    m.step_device()
    torch.cuda.synchronize()
    m.use_cuda_graph, m.world_size = saved, saved_world
    restore()
    rows = cg_rows(recs, args.dtype)
    table, total_ms = kernel_table(rows, peaks)
    named = {k: v for k, v in table.items() if k not in ("other", "affine_act_*")}
    top_name = max(named, key=lambda k: named[k]["ms"])
    top = named[top_name]
    gflop_pair = CG_GFLOP_PER_PAIR * (F / 64.0) ** 2 * (size / 256.0) ** 2
    step_tflops = gflop_pair * n / (dev_ms / args.steps)          # GFLOP / ms == TFLOP/s
    roof = {"bound": "tensor", "achieved": top["TFLOPs"], "peak": peaks["tflops_burst"], "unit": "TFLOP/s",
            "frac": top["TFLOPs"] / peaks["tflops_burst"], "kernel": top_name, "kernel_ms_per_step": top["ms"],
            "kernel_share_of_step": top["share_of_step"], "launches_per_step": top["ops"], "traffic": None,
            "arithmetic_intensity_flop_per_byte": top["arithmetic_intensity_flop_per_byte"],
            "algorithmic_flops_per_launch": (top["TFLOPs"] * 1e12 * top["ms"] * 1e-3) / max(top["ops"], 1),
            "avg_launch_ms": top["ms"] / max(top["ops"], 1),
            "peak_source": peaks["source"] + ", burst figure (kernels timed one by one)",
            "step_conv_tflops_vs_algorithmic_minimum": step_tflops,
            "step_frac_of_tensor_peak_sustained": step_tflops / peaks["tflops_sustained"]}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        pps, spt, cores = cg_cpu_throughput(size, F, 1, 1, 1)
        cpu = {"value": pps, "unit": "image pairs/s", "cores": cores, "kind": "port",
               "sample": "1 step of batch 1 after 1 warm-up, oracle/cyclegan.py train step (torch CPU fp32)"}
    emit({"metric": CG_METRIC, "value": value, "unit": "image pairs/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
          "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
          "data": "synthetic",
          "config": {"workload": f"CycleGAN filters={F} {size}x{size}x1 train_step_torch (6 G fwd, 2+4 D fwd, 2 backward phases, 4 Adam), batch {n} per GPU, image pools 50",
                     "global_batch": n * world, "parallelism": f"dp{world}", "cuda_graph": (not args.no_graph) and world == 1,
                     "l2": "per-step working set (activations of 12 towers, several GB) exceeds the 126 MB L2; no explicit flush"},
          "e2e": {"value": e2e_value, "unit": "image pairs/s", "ms_per_step": e2e_ms / args.steps,
                  "h2d_bytes_per_step": int(2 * n * size * size * 4), "d2h_bytes_per_step": 64},
          "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
          "roofline": roof, "roofline_by_kernel": table, "cpu_baseline": cpu, "clocks": clocks,
          "last_step_metrics": {k: float(v) for k, v in last.items()}})
    if args.profile_out:
        with open(args.profile_out, "w") as fh:
            json.dump({"total_ms_eager": total_ms, "rows": rows}, fh, indent=1)
    return finish_distributed(world, [m])


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print to stdout (NCCL writes "NCCL version ..." there): point fd 1 at stderr for the whole run and keep the
    # original stdout for the JSON line only.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the first-step loss check against the oracle")
    ap.add_argument("--workload", default="unet", choices=["unet", "cyclegan"],
                    help="unet = BASELINE configs[1] (and configs[3] with --size 512 --batch 8 under torchrun); cyclegan = configs[2]")
    ap.add_argument("--filters", type=int, default=64, help="CycleGAN base filters (StartProcess.py:37)")
    ap.add_argument("--profile-out", default=None)
    args = ap.parse_args()
    if args.workload == "cyclegan":
        if args.batch == 32:
            args.batch = 8              # BASELINE configs[2]: batch 8
        return run_cyclegan_reference(args) if args.impl == "reference" else run_cyclegan(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
