"""Kernel timeline of graph-replayed train steps (torch.profiler / CUPTI): where the step's time goes with the real overlap
of the weight-gradient stream, and how large the gaps between dependent kernels are.

    python scripts/trace_step.py [--workload unet|cyclegan] [--out gpurun_out/trace_summary.json]
"""
import argparse, collections, json, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from torch.profiler import profile, ProfilerActivity
import sem_b200
from _inputs import synthetic_batch          # (scripts/ does not import oracle/)

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="unet")
ap.add_argument("--out", default="gpurun_out/trace_summary.json")
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()

if args.workload == "unet":
    from sem_b200 import UNetModel
    x, y, wgt = synthetic_batch(32, 256, 256)
    m = UNetModel((256, 256, 1), 16, dtype="bf16", batch_size=32)
    m.compile(weighting=wgt)
    xp, yp = x.pin_memory(), y.pin_memory()
    step = lambda: m.train_step(xp, yp)
else:
    from sem_b200 import CycleGanModel, ImagePool
    m = CycleGanModel((256, 256, 1), batch_size=8, filters=64, dtype="bf16", image_pool_a=ImagePool(8, 50, random.Random(0)),
                      image_pool_b=ImagePool(8, 50, random.Random(1)))
    m.compile()
    g = torch.Generator().manual_seed(0)
    a = (torch.rand(8, 256, 256, 1, generator=g) * 2 - 1).pin_memory()
    b = (torch.rand(8, 256, 256, 1, generator=g) * 2 - 1).pin_memory()
    step = lambda: m.train_step((a, b))
for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev), key=lambda t: t[0])
if not ks:
    print("no CUDA events captured"); sys.exit(0)
t0, t1 = ks[0][0], max(k[1] for k in ks)
span = (t1 - t0) / args.steps
# union of busy intervals
busy, cur_s, cur_e = 0.0, ks[0][0], ks[0][1]
for s, e, _ in ks[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
by = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in ks:
    k = n.split("<")[0].split("(")[0].replace("void ", "").strip()
    by[k][0] += 1
    by[k][1] += e - s
tot = sum(v[1] for v in by.values())
rows = sorted(((k, v[0] / args.steps, v[1] / args.steps) for k, v in by.items()), key=lambda r: -r[2])
out = {"workload": args.workload, "steps": args.steps, "span_us_per_step": span, "gpu_busy_us_per_step": busy / args.steps,
       "idle_us_per_step": span - busy / args.steps, "sum_kernel_us_per_step": tot / args.steps,
       "overlap_us_per_step": (tot - busy) / args.steps, "kernels_per_step": len(ks) / args.steps,
       "by_kernel": [{"kernel": k, "launches_per_step": n, "us_per_step": round(us, 1)} for k, n, us in rows[:30]]}
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(out, open(args.out, "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "by_kernel"}))
for r in out["by_kernel"][:22]:
    print(f"  {r['us_per_step']:9.1f} us  n={r['launches_per_step']:6.1f}  {r['kernel']}")
