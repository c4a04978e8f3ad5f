"""Per-launch kernel durations of ONE graph-replayed UNet train step (torch.profiler chrome trace: name, grid, block,
start, duration), in launch order.  With SEMB_NO_WGRAD_STREAM=1 the weight gradients are serialised on the main stream,
so every duration is the kernel's own (only the programmatic-dependent-launch prologues overlap).

    [SEMB_NO_WGRAD_STREAM=1] python scripts/trace_kernels.py --out gpurun_out/kernels.json
"""
import argparse, collections, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from torch.profiler import profile, ProfilerActivity
import sem_b200
from _inputs import synthetic_batch          # (scripts/ does not import oracle/)

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/kernels.json")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=256)
args = ap.parse_args()

from sem_b200 import UNetModel
x, y, wgt = synthetic_batch(args.batch, args.size, args.size)
m = UNetModel((args.size, args.size, 1), 16, dtype="bf16", batch_size=args.batch)
m.compile(weighting=wgt)
xp, yp = x.pin_memory(), y.pin_memory()
for _ in range(6):
    m.train_step(xp, yp)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.train_step(xp, yp)
    torch.cuda.synchronize()
tmp = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(tmp)
tr = json.load(open(tmp))
ev = [e for e in tr["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
rows = []
for e in ev:
    a = e.get("args", {})
    name = e["name"].split("<")[0].split("(")[0].replace("void ", "").replace("semb::", "").strip()
    rows.append({"k": name, "full": e["name"][:160], "ts": e["ts"], "us": e["dur"], "grid": a.get("grid"), "block": a.get("block"),
                 "smem": a.get("shared memory"), "stream": a.get("stream")})
t0 = rows[0]["ts"]
for r in rows:
    r["ts"] = round(r["ts"] - t0, 2)
span = max(r["ts"] + r["us"] for r in rows)
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump({"span_us": span, "n": len(rows), "rows": rows}, open(args.out, "w"))
by = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    by[r["k"]][0] += 1
    by[r["k"]][1] += r["us"]
print(json.dumps({"span_us": span, "kernels": len(rows), "sum_us": sum(r["us"] for r in rows)}))
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {v[1]:9.1f} us  n={v[0]:4d}  {k}")
