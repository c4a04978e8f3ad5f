"""Per-layer timings of the hot kernels at the bench workload's shapes (MultiRes-UNet filters=16, 256x256, batch 32).

    python scripts/bench_layers.py [--batch 32] [--reps 10] [--out gpurun_out/layers.json]

Every kernel is called through the C ABI on bf16 tensors; each timing is the mean of `reps` launches bracketed by
CUDA events on the launching stream, rotating over enough buffer copies that the working set exceeds the 126 MB L2.
GB/s = algorithmic bytes (input + output activations once) / time.
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sem_b200  # noqa: F401
from sem_b200 import _lib as L

# (H=W, Cin, Cout, k): physical channels of the bench network's heaviest layers
CONV_SHAPES = [
    (256, 8, 8, 3), (256, 8, 16, 3), (256, 16, 16, 3), (256, 32, 16, 3), (256, 32, 8, 3), (256, 32, 32, 1), (256, 32, 16, 1),
    (128, 32, 8, 3), (128, 8, 24, 3), (128, 24, 32, 3), (128, 64, 32, 3), (128, 32, 32, 3), (128, 64, 64, 1),
    (64, 64, 24, 3), (64, 24, 40, 3), (64, 40, 56, 3), (64, 120, 64, 3), (64, 128, 120, 1),
    (32, 128, 40, 3), (32, 72, 112, 3), (32, 256, 72, 3), (32, 144, 216, 3),
    (16, 224, 72, 3), (16, 144, 216, 3), (16, 432, 432, 1),
]
AFF_SHAPES = [(256, 8, False), (256, 16, False), (256, 16, True), (256, 32, True), (256, 32, False), (128, 64, True), (128, 64, False),
              (64, 120, True), (32, 224, True)]


def timed(fn, reps, ncopies):
    for i in range(min(3, ncopies)):
        fn(i % ncopies)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i % ncopies)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="conv,affine")
    ap.add_argument("--shapes", default=None, help="override the conv shapes: 'H,Cin,Cout,k;...'")
    args = ap.parse_args()
    L.require_device()
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    n = args.batch
    rows = []

    def copies(nbytes):
        return max(2, min(8, int(300e6 // max(nbytes, 1)) + 1))

    if "conv" in args.only:
        shapes = CONV_SHAPES if not args.shapes else [tuple(int(v) for v in t.split(",")) for t in args.shapes.split(";")]
        for (hw, cin, cout, k) in shapes:
            p = k // 2
            geom = L.ConvGeom(n, hw, hw, hw, hw, cin, cout, k, k, 1, p, p, L.PAD_ZERO, L.BF16)
            gd = L.ConvGeom(n, hw, hw, hw, hw, cout, cin, k, k, 1, k - 1 - p, k - 1 - p, L.PAD_ZERO, L.BF16)
            nb = n * hw * hw * (cin + cout) * 2
            nc = copies(nb)
            xs = [torch.randn((n, hw, hw, cin), device="cuda").to(torch.bfloat16) for _ in range(nc)]
            ys = [torch.randn((n, hw, hw, cout), device="cuda").to(torch.bfloat16) for _ in range(nc)]
            w = (torch.randn((k, k, cin, cout), device="cuda") * 0.1).contiguous()
            dw = torch.zeros_like(w)
            nby = lib.semb_pack_weights_tc(None, k, k, cin, cout, 0, None, None)
            wp = torch.zeros(nby // 2, dtype=torch.bfloat16, device="cuda")
            wpf = torch.zeros(lib.semb_pack_weights_tc(None, k, k, cin, cout, 1, None, None) // 2, dtype=torch.bfloat16, device="cuda")
            lib.semb_pack_weights_tc(w.data_ptr(), k, k, cin, cout, 0, wp.data_ptr(), st)
            lib.semb_pack_weights_tc(w.data_ptr(), k, k, cin, cout, 1, wpf.data_ptr(), st)
            stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
            xv = [L.Tensor(t.data_ptr(), cin, cin, 0) for t in xs]
            yv = [L.Tensor(t.data_ptr(), cout, cout, 0) for t in ys]

            def fwd(i):
                L.check(lib.semb_conv2d_fwd_tc(C.byref(geom), C.byref(xv[i]), wp.data_ptr(), None, C.byref(yv[i]), stats.data_ptr(), 0, cout, 0, st))

            def dgrad(i):
                L.check(lib.semb_conv2d_fwd_tc(C.byref(gd), C.byref(yv[i]), wpf.data_ptr(), None, C.byref(xv[i]), None, 0, 0, 0, st))

            def wgrad(i):
                L.check(lib.semb_conv2d_wgrad_tc(C.byref(geom), C.byref(xv[i]), C.byref(yv[i]), dw.data_ptr(), st))

            flops = 2.0 * n * hw * hw * k * k * cin * cout
            for name, fn in (("fwd", fwd), ("dgrad", dgrad), ("wgrad", wgrad)):
                ms = timed(fn, args.reps, nc)
                rows.append({"kernel": f"conv_{name}", "shape": f"{hw}x{hw} {cin}->{cout} k{k}", "ms": ms, "GBps": nb / ms / 1e6,
                             "TFLOPs": flops / ms / 1e9})
                print(f"conv_{name:6s} {hw:3d}x{hw:<3d} {cin:3d}->{cout:<3d} k{k}  {ms * 1e3:8.1f} us  {nb / ms / 1e6:7.0f} GB/s  {flops / ms / 1e9:7.1f} TF/s",
                      flush=True)
            del xs, ys

    if "affine" in args.only:
        for (hw, c, has_b) in AFF_SHAPES:
            nelem = n * hw * hw * c
            nc = copies(nelem * 2 * 3)
            A = [torch.randn((n, hw, hw, c), device="cuda").to(torch.bfloat16) for _ in range(nc)]
            B = [torch.randn((n, hw, hw, c), device="cuda").to(torch.bfloat16) for _ in range(nc)]
            Y = [torch.zeros((n, hw, hw, c), device="cuda", dtype=torch.bfloat16) for _ in range(nc)]
            DA = [torch.zeros((n, hw, hw, c), device="cuda", dtype=torch.bfloat16) for _ in range(nc)]
            DB = [torch.zeros((n, hw, hw, c), device="cuda", dtype=torch.bfloat16) for _ in range(nc)]
            vec = lambda: torch.rand(c, device="cuda") + 0.5
            sa, ta, ma, ia, c1a, c2a, sb, tb, mb, ib, c1b, c2b = [vec() for _ in range(12)]
            stats = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
            sums = torch.zeros(4 * c, device="cuda")
            d = L.AffineDesc(n, hw * hw, c, L.BF16, L.ACT_RELU, L.ACT_NONE, L.AFF_BATCH, L.AFF_BATCH if has_b else L.AFF_NONE, 0)
            tv = lambda t: L.Tensor(t.data_ptr(), c, c, 0)
            av, bv, yv, dav, dbv = [[tv(t) for t in X] for X in (A, B, Y, DA, DB)]
            P = lambda t: t.data_ptr()

            def fwd(i):
                L.check(lib.semb_affine_act_fwd(C.byref(d), C.byref(av[i]), P(sa), P(ta), C.byref(bv[i]) if has_b else None,
                                                P(sb) if has_b else None, P(tb) if has_b else None, C.byref(yv[i]), P(stats), 0, c, st))

            def red(i):
                L.check(lib.semb_affine_act_bwd_reduce(C.byref(d), C.byref(yv[i]), C.byref(av[i]), C.byref(bv[i]) if has_b else None,
                                                       P(sa), P(ta), P(ma), P(ia), P(sb), P(tb), P(mb), P(ib), P(sums), 0, c, st))

            def app(i):
                L.check(lib.semb_affine_act_bwd_apply(C.byref(d), C.byref(yv[i]), C.byref(av[i]), C.byref(bv[i]) if has_b else None,
                                                      P(sa), P(ta), P(ma), P(ia), P(c1a), P(c2a), P(sb), P(tb), P(mb), P(ib), P(c1b), P(c2b),
                                                      C.byref(dav[i]), 0, C.byref(dbv[i]) if has_b else None, 0, st))

            nt = {"fwd": 2 + has_b, "bwd_reduce": 2 + has_b, "bwd_apply": 3 + 2 * has_b}
            for name, fn in (("fwd", fwd), ("bwd_reduce", red), ("bwd_apply", app)):
                ms = timed(fn, args.reps, nc)
                nb = nelem * 2 * nt[name]
                rows.append({"kernel": f"affine_{name}", "shape": f"{hw}x{hw} C={c} b={int(has_b)}", "ms": ms, "GBps": nb / ms / 1e6})
                print(f"affine_{name:10s} {hw:3d}x{hw:<3d} C={c:<3d} b={int(has_b)}  {ms * 1e3:8.1f} us  {nb / ms / 1e6:7.0f} GB/s", flush=True)
            del A, B, Y, DA, DB
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
