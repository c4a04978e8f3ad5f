"""Op-level parity of the C-ABI kernels against the oracle layer restatements (torch CPU fp32).

fp32 storage: rtol 1e-4 (normalised by max|ref|).  bf16 storage: inputs are rounded to bf16 first and the
oracle runs on the rounded values, so the only differences are the bf16 rounding of the stored result
(2^-8 relative) and fp32 summation order; tolerance 1e-2 on stored tensors, 1e-4 on fp32 outputs
(moments, weight gradients).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import layers as OL
from tests import util as U
from sem_b200 import _lib as L

pytestmark = pytest.mark.gpu

DT = ["f32", "bf16"]


def _tol(dtype):
    return 1e-4 if dtype == "f32" else 1e-2


def _prep(x, dtype):
    return U.bf16_round(x) if dtype == "bf16" else x


# (N,H,W,Cin,Cout,k,stride,padding,pad_mode)
CONV_CASES = [
    (2, 16, 16, 1, 4, 3, 1, "same", "zero"),
    (2, 24, 20, 8, 13, 3, 1, "same", "zero"),
    (1, 16, 16, 25, 51, 1, 1, "same", "zero"),
    (2, 16, 24, 51, 32, 3, 1, "same", "zero"),
    (1, 8, 8, 212, 71, 3, 1, "same", "zero"),
    (1, 12, 12, 105, 212, 1, 1, "same", "zero"),
    (2, 16, 16, 16, 32, 3, 2, "same", "zero"),      # CycleGAN generator downsample (asymmetric 0/1 pad)
    (1, 18, 18, 32, 32, 3, 1, "reflect1", "reflect"),  # residual block: reflect-pad 1 + valid
    (1, 22, 22, 1, 16, 7, 1, "reflect3", "reflect"),   # generator stem
    (2, 20, 20, 1, 24, 4, 2, "valid", "zero"),      # PatchGAN d0
    (1, 15, 15, 24, 48, 4, 2, "valid", "zero"),     # odd input
    (1, 12, 12, 48, 1, 4, 1, "valid", "zero"),      # PatchGAN output conv
    (1, 17, 19, 5, 7, 3, 1, "same", "zero"),        # ragged sizes, partial tiles
]


def _geom(n, h, w, cin, cout, k, stride, padding, pad_mode, dtype):
    if padding == "same":
        if stride == 1:
            pt = pl = (k - 1) // 2
            oh, ow = h, w
        else:
            pt, _ = OL.same_pad_amounts(h, k, stride)
            pl, _ = OL.same_pad_amounts(w, k, stride)
            oh, ow = -(-h // stride), -(-w // stride)
    elif padding == "valid":
        pt = pl = 0
        oh, ow = (h - k) // stride + 1, (w - k) // stride + 1
    else:  # reflectP: reflect pad P each side then valid
        p = int(padding[-1])
        pt = pl = p
        oh, ow = h + 2 * p - k + 1, w + 2 * p - k + 1
    pm = L.PAD_REFLECT if pad_mode == "reflect" else L.PAD_ZERO
    return L.ConvGeom(n, h, w, oh, ow, U.pad8(cin), U.pad8(cout), k, k, stride, pt, pl, pm, U.ldtype(dtype)), (oh, ow)


def _oracle_conv(x, w, b, stride, padding):
    if padding.startswith("reflect"):
        p = int(padding[-1])
        return OL.conv2d(OL.reflection_pad(x, 2 * p, 2 * p), w, b, stride, "valid")
    return OL.conv2d(x, w, b, stride, padding)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(case, dtype):
    n, h, w_, cin, cout, k, stride, padding, pad_mode = case
    lib = L.load()
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = _prep(torch.randn(n, h, w_, cin, generator=g), dtype)
    wt = torch.randn(k, k, cin, cout, generator=g) * 0.2
    bias = torch.randn(cout, generator=g)
    geom, (oh, ow) = _geom(n, h, w_, cin, cout, k, stride, padding, pad_mode, dtype)

    xr = x.clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    y_ref = _oracle_conv(xr, wr, br, stride, padding)
    assert y_ref.shape[1:3] == (oh, ow)
    dy = _prep(torch.randn(y_ref.shape, generator=g), dtype)
    y_ref.backward(dy)

    xd = U.to_dev(x, dtype, pitch=U.pad8(cin) + 8, coff=8)           # exercise pitch/coff
    yd = torch.zeros((n, oh, ow, U.pad8(cout) + 16), dtype=U.tdtype(dtype), device="cuda")
    wd, bd = U.pad_w(wt), U.pad_v(bias)
    stats = torch.zeros(2 * U.pad8(cout), device="cuda", dtype=torch.float64)
    xv, yv = U.view(xd, 8, U.pad8(cin)), U.view(yd, 8, U.pad8(cout))
    L.check(lib.semb_conv2d_fwd(C.byref(geom), C.byref(xv), wd.data_ptr(), bd.data_ptr(), C.byref(yv), stats.data_ptr(), 0,
                                U.pad8(cout), 0, U.stream()))
    torch.cuda.synchronize()
    y = yd[..., 8:8 + cout].float().cpu()
    assert U.rel_err(y, y_ref) < _tol(dtype)
    assert float(yd[..., :8].abs().max()) == 0 and float(yd[..., 8 + U.pad8(cout):].abs().max()) == 0
    # moments come from the fp32 accumulators
    s_ref = y_ref.detach().sum(dim=(0, 1, 2))
    q_ref = (y_ref.detach() ** 2).sum(dim=(0, 1, 2))
    assert U.rel_err(stats[:cout], s_ref) < 1e-4 and U.rel_err(stats[U.pad8(cout):U.pad8(cout) + cout], q_ref) < 1e-4

    # wgrad (+ dbias)
    dyd = U.to_dev(dy, dtype)
    dw = torch.zeros_like(wd)
    db = torch.zeros_like(bd)
    dyv = U.view(dyd)
    L.check(lib.semb_conv2d_wgrad(C.byref(geom), C.byref(xv), C.byref(dyv), dw.data_ptr(), db.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(dw[:, :, :cin, :cout], wr.grad) < 1e-4
    assert U.rel_err(db[:cout], br.grad) < 1e-4
    assert float(dw[:, :, cin:, :].abs().max() if U.pad8(cin) > cin else 0) == 0

    # dgrad (zero padding only; reflect is folded separately)
    if pad_mode == "zero":
        dxd = torch.zeros((n, h, w_, U.pad8(cin)), dtype=U.tdtype(dtype), device="cuda")
        dxv = U.view(dxd)
        L.check(lib.semb_conv2d_dgrad(C.byref(geom), C.byref(dyv), wd.data_ptr(), None, C.byref(dxv), None, 0, 0, 0, U.stream()))
        torch.cuda.synchronize()
        assert U.rel_err(dxd[..., :cin], xr.grad) < _tol(dtype)
        # accumulate flag adds on top
        L.check(lib.semb_conv2d_dgrad(C.byref(geom), C.byref(dyv), wd.data_ptr(), None, C.byref(dxv), None, 0, 0, 1, U.stream()))
        torch.cuda.synchronize()
        assert U.rel_err(dxd[..., :cin], 2 * xr.grad) < 2 * _tol(dtype)
    else:
        # gradient on the padded domain, then reflect-fold
        p = geom.pad_t
        g2 = L.ConvGeom(n, h + 2 * p, w_ + 2 * p, oh, ow, geom.Cin, geom.Cout, k, k, stride, 0, 0, L.PAD_ZERO, geom.dtype)
        dxp = torch.zeros((n, h + 2 * p, w_ + 2 * p, U.pad8(cin)), dtype=U.tdtype(dtype), device="cuda")
        dxd = torch.zeros((n, h, w_, U.pad8(cin)), dtype=U.tdtype(dtype), device="cuda")
        dxpv, dxv = U.view(dxp), U.view(dxd)
        L.check(lib.semb_conv2d_dgrad(C.byref(g2), C.byref(dyv), wd.data_ptr(), None, C.byref(dxpv), None, 0, 0, 0, U.stream()))
        L.check(lib.semb_pad_crop(C.byref(dxpv), C.byref(dxv), n, h + 2 * p, w_ + 2 * p, h, w_, p, p, 3, geom.dtype, 0, U.stream()))
        torch.cuda.synchronize()
        assert U.rel_err(dxd[..., :cin], xr.grad) < 2 * _tol(dtype)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("case", [(2, 8, 8, 51, 16, 2), (1, 6, 10, 426, 128, 2), (2, 8, 8, 32, 16, 3)])
def test_conv_transpose(case, dtype):
    """Conv2DTranspose 2x2 s2 (UNet up path, with bias) and 3x3 s2 (CycleGAN upsample) via the dgrad kernel."""
    n, h, w_, cin, cout, k = case
    lib = L.load()
    g = torch.Generator().manual_seed(7)
    x = _prep(torch.randn(n, h, w_, cin, generator=g), dtype)
    wt = torch.randn(k, k, cout, cin, generator=g) * 0.1        # Keras layout (kh,kw,Cout,Cin)
    bias = torch.randn(cout, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    y_ref = OL.conv2d_transpose(xr, wr, br, 2)
    assert y_ref.shape == (n, 2 * h, 2 * w_, cout)
    dy = _prep(torch.randn(y_ref.shape, generator=g), dtype)
    y_ref.backward(dy)
    p, _ = OL.conv_transpose_pads(k, 2)
    geom = L.ConvGeom(n, 2 * h, 2 * w_, h, w_, U.pad8(cout), U.pad8(cin), k, k, 2, p, p, L.PAD_ZERO, U.ldtype(dtype))
    xd, yd = U.to_dev(x, dtype), torch.zeros((n, 2 * h, 2 * w_, U.pad8(cout)), dtype=U.tdtype(dtype), device="cuda")
    wd, bd = U.pad_w(wt), U.pad_v(bias)
    xv, yv = U.view(xd), U.view(yd)
    L.check(lib.semb_conv2d_dgrad(C.byref(geom), C.byref(xv), wd.data_ptr(), bd.data_ptr(), C.byref(yv), None, 0, 0, 0, U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(yd[..., :cout], y_ref) < _tol(dtype)
    # backward: d_in = conv(d_out), dW = wgrad(d_out, in), dbias = channel_sum(d_out)
    dyd = U.to_dev(dy, dtype)
    dyv = U.view(dyd)
    dxd = torch.zeros_like(xd)
    dxv = U.view(dxd)
    dw, db = torch.zeros_like(wd), torch.zeros_like(bd)
    L.check(lib.semb_conv2d_fwd(C.byref(geom), C.byref(dyv), wd.data_ptr(), None, C.byref(dxv), None, 0, 0, 0, U.stream()))
    L.check(lib.semb_conv2d_wgrad(C.byref(geom), C.byref(dyv), C.byref(xv), dw.data_ptr(), None, U.stream()))
    L.check(lib.semb_channel_sum(C.byref(dyv), n, 4 * h * w_, db.data_ptr(), U.ldtype(dtype), U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(dxd[..., :cin], xr.grad) < _tol(dtype)
    assert U.rel_err(dw[:, :, :cout, :cin], wr.grad) < 1e-4
    assert U.rel_err(db[:cout], br.grad) < 1e-4


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("per_sample", [False, True])
@pytest.mark.parametrize("c", [24, 5, 120, 426])       # 3, 1, 15 and 54 channel groups: every block-reduction layout
def test_norm_affine_fwd_bwd(dtype, per_sample, c):
    """y = relu(BN_a(a) + relu(BN_b(b))) with batch statistics (res_path unit) and the InstanceNorm variant,
    forward + both gradient passes against torch autograd on the oracle formulas."""
    lib = L.load()
    n, h, w_ = 3, 10, 12
    cp = U.pad8(c)
    g = torch.Generator().manual_seed(3)
    a = _prep(torch.randn(n, h, w_, c, generator=g) * 2 + 0.5, dtype)
    b = _prep(torch.randn(n, h, w_, c, generator=g) - 0.3, dtype)
    gamma_b = torch.rand(c, generator=g) + 0.5
    beta_a, beta_b = torch.randn(c, generator=g) * 0.1, torch.randn(c, generator=g) * 0.1
    ar, br_ = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    gb, ba, bb = gamma_b.clone().requires_grad_(True), beta_a.clone().requires_grad_(True), beta_b.clone().requires_grad_(True)
    eps = 1e-5 if per_sample else 1e-3
    if per_sample:
        ya = OL.instance_norm(ar, torch.ones(c), ba, eps)
        yb = OL.instance_norm(br_, gb, bb, eps)
    else:
        ya, _, _ = OL.batch_norm(ar, None, ba, torch.zeros(c), torch.ones(c), True)
        yb, _, _ = OL.batch_norm(br_, gb, bb, torch.zeros(c), torch.ones(c), True)
    y_ref = torch.relu(ya + torch.relu(yb))
    dy = _prep(torch.randn(y_ref.shape, generator=g), dtype)
    y_ref.backward(dy)

    groups = n if per_sample else 1
    count = float(h * w_ if per_sample else n * h * w_)
    ad, bd, yd = U.to_dev(a, dtype), U.to_dev(b, dtype), torch.zeros((n, h, w_, cp), dtype=U.tdtype(dtype), device="cuda")
    av, bv, yv = U.view(ad), U.view(bd), U.view(yd)

    def moments(t):
        tt = t.double()
        dims = (1, 2) if per_sample else (0, 1, 2)
        st = torch.zeros(groups, 2, cp, dtype=torch.float64)
        st[:, 0, :c] = tt.sum(dim=dims).reshape(groups, c)
        st[:, 1, :c] = (tt * tt).sum(dim=dims).reshape(groups, c)
        return st.cuda().contiguous()

    nstride = 2 * cp if per_sample else 0
    arrs = {}
    keep = []       # device temporaries must outlive the asynchronous launches that read them
    for tag, t, gam, bet in (("a", a, None, beta_a), ("b", b, gamma_b, beta_b)):
        st = moments(t)
        o = {k: torch.zeros(groups * cp, device="cuda") for k in ("scale", "shift", "mean", "invstd", "c1", "c2")}
        gam_d = U.pad_v(gam) if gam is not None else None
        bet_d = U.pad_v(bet)
        keep += [st, gam_d, bet_d]
        L.check(lib.semb_norm_finalize(st.data_ptr(), groups, cp, cp, nstride, count, eps,
                                       gam_d.data_ptr() if gam_d is not None else None, bet_d.data_ptr(),
                                       o["scale"].data_ptr(), o["shift"].data_ptr(), o["mean"].data_ptr(), o["invstd"].data_ptr(),
                                       None, None, 0.99, U.stream()))
        arrs[tag] = o
    d = L.AffineDesc(n, h * w_, cp, U.ldtype(dtype), L.ACT_RELU, L.ACT_RELU, L.AFF_BATCH, L.AFF_BATCH, cp if per_sample else 0)
    ystats = torch.zeros(groups * 2 * cp, device="cuda", dtype=torch.float64)
    L.check(lib.semb_affine_act_fwd(C.byref(d), C.byref(av), arrs["a"]["scale"].data_ptr(), arrs["a"]["shift"].data_ptr(),
                                    C.byref(bv), arrs["b"]["scale"].data_ptr(), arrs["b"]["shift"].data_ptr(), C.byref(yv),
                                    ystats.data_ptr(), nstride, cp, U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(yd[..., :c], y_ref) < _tol(dtype)
    ys = ystats.view(groups, 2, cp)
    dims = (1, 2) if per_sample else (0, 1, 2)
    assert U.rel_err(ys[:, 0, :c], y_ref.detach().sum(dim=dims).reshape(groups, c)) < (1e-4 if dtype == "f32" else 1e-2)

    # backward
    dyd = U.to_dev(dy, dtype)
    dyv = U.view(dyd)
    # the saved forward output must be the one the kernel produced
    sums = torch.zeros(groups * 4 * cp, device="cuda")
    L.check(lib.semb_affine_act_bwd_reduce(C.byref(d), C.byref(dyv), C.byref(av), C.byref(bv),
                                           arrs["a"]["scale"].data_ptr(), arrs["a"]["shift"].data_ptr(),
                                           arrs["a"]["mean"].data_ptr(), arrs["a"]["invstd"].data_ptr(),
                                           arrs["b"]["scale"].data_ptr(), arrs["b"]["shift"].data_ptr(),
                                           arrs["b"]["mean"].data_ptr(), arrs["b"]["invstd"].data_ptr(),
                                           sums.data_ptr(), 4 * cp if per_sample else 0, cp, U.stream()))
    dgam_b, dbeta_a, dbeta_b = torch.zeros(cp, device="cuda"), torch.zeros(cp, device="cuda"), torch.zeros(cp, device="cuda")
    L.check(lib.semb_norm_bwd_finalize(sums.data_ptr(), 0, groups, cp, cp, 4 * cp if per_sample else 0, count, None, None,
                                       arrs["a"]["c1"].data_ptr(), arrs["a"]["c2"].data_ptr(), None, dbeta_a.data_ptr(), U.stream()))
    L.check(lib.semb_norm_bwd_finalize(sums.data_ptr(), 1, groups, cp, cp, 4 * cp if per_sample else 0, count, None, None,
                                       arrs["b"]["c1"].data_ptr(), arrs["b"]["c2"].data_ptr(), dgam_b.data_ptr(), dbeta_b.data_ptr(),
                                       U.stream()))
    dad, dbd = torch.zeros_like(ad), torch.zeros_like(bd)
    dav, dbv = U.view(dad), U.view(dbd)
    A, B = arrs["a"], arrs["b"]
    L.check(lib.semb_affine_act_bwd_apply(C.byref(d), C.byref(dyv), C.byref(av), C.byref(bv),
                                          A["scale"].data_ptr(), A["shift"].data_ptr(), A["mean"].data_ptr(), A["invstd"].data_ptr(),
                                          A["c1"].data_ptr(), A["c2"].data_ptr(),
                                          B["scale"].data_ptr(), B["shift"].data_ptr(), B["mean"].data_ptr(), B["invstd"].data_ptr(),
                                          B["c1"].data_ptr(), B["c2"].data_ptr(), C.byref(dav), 0, C.byref(dbv), 0, U.stream()))
    torch.cuda.synchronize()
    tol = 1e-3 if dtype == "f32" else 2e-2
    assert U.rel_err(dad[..., :c], ar.grad) < tol
    assert U.rel_err(dbd[..., :c], br_.grad) < tol
    assert U.rel_err(dgam_b[:c], gb.grad) < tol
    assert U.rel_err(dbeta_a[:c], ba.grad) < tol and U.rel_err(dbeta_b[:c], bb.grad) < tol

    # folded variants: forward that finalizes the norms itself, pass 2 straight from the sums (+ dgamma / dbeta)
    fins = {}
    for tag, t, gam, bet in (("a", a, None, beta_a), ("b", b, gamma_b, beta_b)):
        st = moments(t)
        o2 = {k: torch.zeros(groups * cp, device="cuda") for k in ("scale", "shift", "mean", "invstd")}
        gam_d = U.pad_v(gam) if gam is not None else None
        bet_d = U.pad_v(bet)
        keep += [st, gam_d, bet_d, o2]
        f = L.NormFin()
        f.stats, f.stats_nstride, f.cstride, f.count, f.eps = st.data_ptr(), nstride, cp, count, eps
        f.gamma, f.beta = (gam_d.data_ptr() if gam_d is not None else None), bet_d.data_ptr()
        f.scale, f.shift, f.mean, f.invstd = (o2[k].data_ptr() for k in ("scale", "shift", "mean", "invstd"))
        fins[tag] = (f, o2)
    yd2 = torch.zeros_like(yd)
    yv2 = U.view(yd2)
    L.check(lib.semb_affine_act_fwd_fin(C.byref(d), C.byref(av), C.byref(fins["a"][0]), C.byref(bv), C.byref(fins["b"][0]), C.byref(yv2),
                                        None, 0, 0, U.stream()))
    torch.cuda.synchronize()
    assert torch.equal(yd2, yd)
    for tag in ("a", "b"):
        for k in ("scale", "shift", "mean", "invstd"):
            assert torch.equal(fins[tag][1][k], arrs[tag][k]), (tag, k)
    dgam_b3, dbeta_a3, dbeta_b3 = torch.zeros(cp, device="cuda"), torch.zeros(cp, device="cuda"), torch.zeros(cp, device="cuda")
    dad3, dbd3 = torch.zeros_like(ad), torch.zeros_like(bd)
    dav3, dbv3 = U.view(dad3), U.view(dbd3)
    L.check(lib.semb_affine_act_bwd_apply_sums(
        C.byref(d), C.byref(dyv), C.byref(av), C.byref(bv),
        A["scale"].data_ptr(), A["shift"].data_ptr(), A["mean"].data_ptr(), A["invstd"].data_ptr(), count, None, dbeta_a3.data_ptr(),
        B["scale"].data_ptr(), B["shift"].data_ptr(), B["mean"].data_ptr(), B["invstd"].data_ptr(), count, dgam_b3.data_ptr(),
        dbeta_b3.data_ptr(), sums.data_ptr(), 4 * cp if per_sample else 0, cp, C.byref(dav3), 0, C.byref(dbv3), 0, U.stream()))
    torch.cuda.synchronize()
    assert torch.equal(dad3, dad) and torch.equal(dbd3, dbd)
    assert U.rel_err(dgam_b3[:c], dgam_b[:c]) < 1e-5 and U.rel_err(dbeta_a3[:c], dbeta_a[:c]) < 1e-5 and U.rel_err(dbeta_b3[:c], dbeta_b[:c]) < 1e-5

    # the fused cooperative kernel (sums -> grid barrier -> finalize -> gradients) must reproduce the three-launch path
    sums2 = torch.zeros(groups * 4 * cp, device="cuda")
    bar = torch.zeros(2, dtype=torch.int32, device="cuda")
    dgam_b2, dbeta_a2, dbeta_b2 = torch.zeros(cp, device="cuda"), torch.zeros(cp, device="cuda"), torch.zeros(cp, device="cuda")
    dad2, dbd2 = torch.ones_like(ad), torch.zeros_like(bd)          # da accumulates on top of ones
    dav2, dbv2 = U.view(dad2), U.view(dbd2)
    for rep in range(2):                                              # twice: the barrier words must reset themselves
        if rep == 1:
            sums2.zero_(); dgam_b2.zero_(); dbeta_a2.zero_(); dbeta_b2.zero_(); dad2.fill_(1.0)
        L.check(lib.semb_affine_act_bwd_fused(
            C.byref(d), C.byref(dyv), C.byref(av), C.byref(bv),
            A["scale"].data_ptr(), A["shift"].data_ptr(), A["mean"].data_ptr(), A["invstd"].data_ptr(), count, None, dbeta_a2.data_ptr(),
            B["scale"].data_ptr(), B["shift"].data_ptr(), B["mean"].data_ptr(), B["invstd"].data_ptr(), count, dgam_b2.data_ptr(),
            dbeta_b2.data_ptr(), sums2.data_ptr(), 4 * cp if per_sample else 0, cp, bar.data_ptr(),
            C.byref(dav2), 1, C.byref(dbv2), 0, U.stream()))
        torch.cuda.synchronize()
        assert int(bar.abs().max()) == 0
        assert U.rel_err(dad2[..., :c].float() - 1.0, ar.grad) < tol
        assert U.rel_err(dbd2[..., :c], br_.grad) < tol
        assert U.rel_err(dgam_b2[:c], gb.grad) < tol
        assert U.rel_err(dbeta_a2[:c], ba.grad) < tol and U.rel_err(dbeta_b2[:c], bb.grad) < tol


def test_norm_finalize_moving_stats():
    lib = L.load()
    c = 16
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 6, 6, c, generator=g) * 3 + 1
    mm, mv = torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5
    y_ref, m_ref, v_ref = OL.batch_norm(x, None, torch.zeros(c), mm, mv, True)
    st = torch.stack([x.double().sum(dim=(0, 1, 2)), (x.double() ** 2).sum(dim=(0, 1, 2))]).cuda().contiguous()
    mmd, mvd = mm.cuda(), mv.cuda()
    outs = [torch.zeros(c, device="cuda") for _ in range(4)]
    beta0 = torch.zeros(c, device="cuda")
    L.check(lib.semb_norm_finalize(st.data_ptr(), 1, c, c, 0, float(4 * 36), 1e-3, None, beta0.data_ptr(),
                                   outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(),
                                   mmd.data_ptr(), mvd.data_ptr(), 0.99, U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(mmd, m_ref) < 1e-5 and U.rel_err(mvd, v_ref) < 1e-5
    assert U.rel_err(x.cuda() * outs[0] + outs[1], y_ref) < 1e-4


@pytest.mark.parametrize("dtype", DT)
def test_maxpool(dtype):
    lib = L.load()
    g = torch.Generator().manual_seed(1)
    x = _prep(torch.randn(2, 8, 12, 13, generator=g), dtype)
    x[0, 0, 0, 0] = x[0, 0, 1, 0] = x[0, 1, 0, 0] = x[0, 1, 1, 0] = 1.5      # a 4-way tie: first position wins
    xr = x.clone().requires_grad_(True)
    y_ref = OL.max_pool_2x2(xr)
    dy = _prep(torch.randn(y_ref.shape, generator=g), dtype)
    y_ref.backward(dy)
    xd, yd = U.to_dev(x, dtype), torch.zeros((2, 4, 6, 16), dtype=U.tdtype(dtype), device="cuda")
    xv, yv = U.view(xd), U.view(yd)
    L.check(lib.semb_maxpool2x2_fwd(C.byref(xv), C.byref(yv), 2, 8, 12, U.ldtype(dtype), U.stream()))
    dyd, dxd = U.to_dev(dy, dtype), torch.zeros_like(xd)
    dyv, dxv = U.view(dyd), U.view(dxd)
    L.check(lib.semb_maxpool2x2_bwd(C.byref(xv), C.byref(dyv), C.byref(dxv), 2, 8, 12, U.ldtype(dtype), 0, U.stream()))
    torch.cuda.synchronize()
    assert torch.equal(yd[..., :13].float().cpu(), y_ref.detach())
    assert torch.equal(dxd[..., :13].float().cpu(), xr.grad)


@pytest.mark.parametrize("dtype", DT)
def test_pad_crop_modes(dtype):
    lib = L.load()
    g = torch.Generator().manual_seed(2)
    n, h, w_, c = 2, 9, 11, 8
    x = _prep(torch.randn(n, h, w_, c, generator=g), dtype)
    xr = x.clone().requires_grad_(True)
    y_ref = OL.reflection_pad(xr, 5, 7)      # totals: w 5 -> (2,3), h 7 -> (3,4)
    dy = _prep(torch.randn(y_ref.shape, generator=g), dtype)
    y_ref.backward(dy)
    oh, ow = h + 7, w_ + 5
    xd, yd = U.to_dev(x, dtype), torch.zeros((n, oh, ow, c), dtype=U.tdtype(dtype), device="cuda")
    xv, yv = U.view(xd), U.view(yd)
    L.check(lib.semb_pad_crop(C.byref(xv), C.byref(yv), n, h, w_, oh, ow, 3, 2, 0, U.ldtype(dtype), 0, U.stream()))
    dyd, dxd = U.to_dev(dy, dtype), torch.zeros_like(xd)
    dyv, dxv = U.view(dyd), U.view(dxd)
    L.check(lib.semb_pad_crop(C.byref(dyv), C.byref(dxv), n, oh, ow, h, w_, 3, 2, 3, U.ldtype(dtype), 0, U.stream()))
    torch.cuda.synchronize()
    assert torch.equal(yd.float().cpu(), y_ref.detach())
    assert U.rel_err(dxd, xr.grad) < _tol(dtype)
    # crop and its gradient (zero pad)
    cd = torch.zeros((n, h, w_, c), dtype=U.tdtype(dtype), device="cuda")
    cv = U.view(cd)
    L.check(lib.semb_pad_crop(C.byref(yv), C.byref(cv), n, oh, ow, h, w_, 3, 2, 1, U.ldtype(dtype), 0, U.stream()))
    zd = torch.ones((n, oh, ow, c), dtype=U.tdtype(dtype), device="cuda")
    zv = U.view(zd)
    L.check(lib.semb_pad_crop(C.byref(xv), C.byref(zv), n, h, w_, oh, ow, 3, 2, 2, U.ldtype(dtype), 0, U.stream()))
    torch.cuda.synchronize()
    assert torch.equal(cd.float().cpu(), x)
    ref = torch.zeros(n, oh, ow, c)
    ref[:, 3:3 + h, 2:2 + w_] = x
    assert torch.equal(zd.float().cpu(), ref)


@pytest.mark.parametrize("dtype", DT)
def test_loss_wbce(dtype):
    lib = L.load()
    g = torch.Generator().manual_seed(5)
    n, h, w_ = 2, 16, 16
    p = torch.rand(n, h, w_, 1, generator=g)
    p[0, 0, 0, 0], p[0, 0, 1, 0] = 0.0, 1.0          # outside the clip range -> zero gradient
    p = _prep(p, dtype)
    y = (torch.rand(n, h, w_, 1, generator=g) < 0.3).float()
    pr = p.clone().requires_grad_(True)
    loss = OL.weighted_bce(y, pr, 4.5)
    loss.backward()
    pd = U.to_dev(p, dtype)
    dpd = torch.zeros_like(pd)
    out = torch.zeros(4, device="cuda")
    pv, dpv = U.view(pd), U.view(dpd)
    yd = y.cuda().contiguous()
    L.check(lib.semb_loss_wbce(C.byref(pv), yd.data_ptr(), C.byref(dpv), n * h * w_, 4.5, out.data_ptr(),
                               U.ldtype(dtype), U.stream()))
    torch.cuda.synchronize()
    cnt = n * h * w_
    assert abs(float(out[0]) / cnt - float(loss)) < 1e-5 * max(1.0, abs(float(loss)))
    assert abs(float(out[1]) / cnt - float((y - p).abs().mean())) < 1e-5
    assert abs(float(out[2]) / cnt - float(((p > 0.5).float() == y).float().mean())) < 1e-6
    assert U.rel_err(dpd[..., :1], pr.grad) < (1e-5 if dtype == "f32" else 1e-2)


@pytest.mark.parametrize("dtype", DT)
def test_loss_wbce_logits_saturated_pixels_keep_their_gradient(dtype):
    """with the probability recomputed in fp32 from the stored pre-activation, a confidently wrong pixel
    (logit 8, p = 0.99966 -- exactly 1.0 after bf16 rounding) keeps the Keras gradient; the clip only acts past |u| ~ 16."""
    lib = L.load()
    g = torch.Generator().manual_seed(8)
    n, h, w_ = 2, 16, 16
    z = torch.randn(n, h, w_, 1, generator=g) * 3.0
    z[0, 0, 0, 0], z[0, 0, 1, 0], z[0, 0, 2, 0] = 8.0, -8.0, 40.0
    z = _prep(z, dtype)
    y = (torch.rand(n, h, w_, 1, generator=g) < 0.3).float()
    y[0, 0, 0, 0], y[0, 0, 1, 0], y[0, 0, 2, 0] = 0.0, 1.0, 0.0
    sc, sh = 1.25, -0.1
    zr = z.clone().requires_grad_(True)
    pr = torch.sigmoid(zr * sc + sh)
    pr.retain_grad()
    loss = OL.weighted_bce(y, pr, 4.5)
    loss.backward()
    zd = U.to_dev(z, dtype)
    dpd = torch.zeros_like(zd)
    out = torch.zeros(4, device="cuda")
    scd, shd = torch.full((8,), sc, device="cuda"), torch.full((8,), sh, device="cuda")
    zv, dpv = U.view(zd), U.view(dpd)
    yd = y.cuda().contiguous()
    L.check(lib.semb_loss_wbce_logits(C.byref(zv), scd.data_ptr(), shd.data_ptr(), yd.data_ptr(), C.byref(dpv), n * h * w_, 4.5,
                                      out.data_ptr(), U.ldtype(dtype), U.stream()))
    torch.cuda.synchronize()
    cnt = n * h * w_
    assert abs(float(out[0]) / cnt - float(loss)) < 1e-4 * max(1.0, abs(float(loss)))
    assert abs(float(out[2]) / cnt - float(((pr > 0.5).float() == y).float().mean())) < 1e-6
    got = dpd[..., :1].float().cpu()
    ref = pr.grad
    assert float(got[0, 0, 2, 0]) == 0.0 and float(ref[0, 0, 2, 0]) == 0.0            # past the fp32 clip: zero in Keras too
    for ix in ((0, 0, 0, 0), (0, 0, 1, 0)):                                            # saturated in bf16, alive in fp32
        assert abs(float(got[ix]) - float(ref[ix])) < (1e-4 if dtype == "f32" else 1e-2) * abs(float(ref[ix])) and float(ref[ix]) != 0.0
    assert U.rel_err(got, ref) < (1e-4 if dtype == "f32" else 1e-2)


@pytest.mark.parametrize("kind", [0, 1])
def test_loss_l1_l2(kind):
    lib = L.load()
    g = torch.Generator().manual_seed(6)
    a, b = torch.randn(2, 8, 8, 1, generator=g), torch.randn(2, 8, 8, 1, generator=g)
    ar = a.clone().requires_grad_(True)
    ref = (OL.mae(ar, b) if kind == 0 else OL.mse(ar, b)) * 10.0
    ref.backward()
    ad, bd = U.to_dev(a, "f32"), U.to_dev(b, "f32")
    dad = torch.zeros_like(ad)
    out = torch.zeros(1, device="cuda")
    av, bv, dav = U.view(ad), U.view(bd), U.view(dad)
    L.check(lib.semb_loss_l1_l2(C.byref(av), C.byref(bv), 0.0, kind, 128, 1, 10.0 / 128, C.byref(dav), 0, out.data_ptr(), L.F32, U.stream()))
    torch.cuda.synchronize()
    assert abs(float(out[0]) * 10.0 / 128 - float(ref)) < 1e-5 * abs(float(ref))
    assert U.rel_err(dad[..., :1], ar.grad) < 1e-5
    # constant target (LSGAN label)
    out.zero_()
    L.check(lib.semb_loss_l1_l2(C.byref(av), None, 1.0, 1, 128, 1, 1.0 / 128, None, 0, out.data_ptr(), L.F32, U.stream()))
    torch.cuda.synchronize()
    assert abs(float(out[0]) / 128 - float(OL.mse(a, torch.ones_like(a)))) < 1e-5


def test_adam_matches_keras_formula():
    lib = L.load()
    g = torch.Generator().manual_seed(8)
    n = 1003
    w0, grads = torch.randn(n, generator=g), [torch.randn(n, generator=g) for _ in range(3)]
    wr = w0.clone()
    opt = OL.KerasAdam([wr], lr=2e-4, beta_1=0.5)
    wd, m, v = w0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    state = torch.zeros(4, dtype=torch.int32, device="cuda")
    lr = torch.full((1,), 2e-4, device="cuda")
    for gr in grads:
        opt.apply([gr], [wr])
        gd = (gr * 4.0).cuda()       # gscale = 1/4 emulates the 1/world_size after an all-reduce sum
        L.check(lib.semb_adam_step(wd.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, lr.data_ptr(), 0.5, 0.999, 1e-7,
                                   0.25, state.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    assert int(state.view(torch.int64)[0]) == 3
    assert U.rel_err(wd, wr) < 1e-6


def test_error_codes():
    lib = L.load()
    t = torch.zeros((1, 4, 4, 12), device="cuda")       # 12 channels: not 8-padded
    v = L.Tensor(t.data_ptr(), 12, 12, 0)
    with pytest.raises(ValueError):
        L.check(lib.semb_maxpool2x2_fwd(C.byref(v), C.byref(v), 1, 4, 4, L.F32, U.stream()))
    assert "maxpool" in L.last_error()


# ---- tcgen05 implicit-GEMM conv -----------------------------------------------------------------------------
# (N,H,W,Cin,Cout,k,pad_mode)
TC_CASES = [
    (2, 32, 24, 8, 16, 3, "zero"),
    (1, 16, 16, 32, 8, 3, "zero"),
    (2, 16, 16, 64, 24, 3, "zero"),
    (1, 16, 16, 144, 216, 3, "zero"),      # K chunks of 16
    (1, 16, 8, 216, 432, 1, "zero"),       # two N chunks
    (1, 20, 12, 24, 40, 3, "zero"),        # partial tiles
    (1, 16, 16, 256, 72, 3, "zero"),
    (2, 16, 16, 56, 32, 1, "zero"),
    (1, 16, 16, 32, 32, 3, "reflect"),     # CycleGAN residual block conv
    (4, 64, 64, 16, 16, 3, "zero"),        # many tiles
    (1, 16, 16, 13, 26, 3, "zero"),        # channel counts that are not multiples of 8 (padded lanes)
    (1, 16, 16, 105, 51, 1, "zero"),
    # two-tiles-per-weight-chunk mode of conv_tma (streamed weights, image wider than one tile):
    (3, 40, 24, 136, 40, 3, "zero"),       # second super-tile of a row has only its left tile inside the image; partial rows
    (8, 128, 128, 72, 24, 3, "zero"),      # 512 super-tiles on <= 148 CTAs: ring wrap-around and accumulator phase flips
    (2, 32, 40, 200, 136, 1, "zero"),      # 1x1, four K chunks, two N chunks
    (2, 32, 32, 264, 200, 3, "zero"),      # N chunks of 112 (re-planned from one chunk of 208)
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_fwd_and_dgrad(case):
    """bf16 tensor-core conv vs the oracle on bf16-rounded operands (x and w): only fp32 summation order and the
    bf16 rounding of the stored output differ."""
    n, h, w_, cin, cout, k, pad_mode = case
    lib = L.load()
    g = torch.Generator().manual_seed(11)
    x = U.bf16_round(torch.randn(n, h, w_, cin, generator=g))
    wt = U.bf16_round(torch.randn(k, k, cin, cout, generator=g) * 0.1)
    bias = torch.randn(cout, generator=g)
    p = (k - 1) // 2
    xr = x.clone().requires_grad_(True)
    if pad_mode == "reflect":
        y_ref = OL.conv2d(OL.reflection_pad(xr, 2 * p, 2 * p), wt, bias, 1, "valid")
    else:
        y_ref = OL.conv2d(xr, wt, bias, 1, "same")
    dy = U.bf16_round(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    cpi, cpo = U.pad8(cin), U.pad8(cout)
    xd = U.to_dev(x, "bf16", pitch=cpi + 8, coff=8)
    geom = L.ConvGeom(n, h, w_, h, w_, cpi, cpo, k, k, 1, p, p, L.PAD_ZERO, L.BF16)
    if pad_mode == "reflect":
        # what the engine does for CycleGAN's residual convs: materialise the reflection padding, then a 'valid' conv
        xpad = torch.zeros((n, h + 2 * p, w_ + 2 * p, cpi + 8), dtype=torch.bfloat16, device="cuda")
        sv, dv = U.view(xd, 8, cpi), U.view(xpad, 8, cpi)
        L.check(lib.semb_pad_crop(C.byref(sv), C.byref(dv), n, h, w_, h + 2 * p, w_ + 2 * p, p, p, 0, L.BF16, 0, U.stream()))
        xd = xpad
        geom = L.ConvGeom(n, h + 2 * p, w_ + 2 * p, h, w_, cpi, cpo, k, k, 1, 0, 0, L.PAD_ZERO, L.BF16)
    yd = torch.zeros((n, h, w_, cpo + 16), dtype=torch.bfloat16, device="cuda")
    wd, bd = U.pad_w(wt), U.pad_v(bias)
    nbytes = lib.semb_pack_weights_tc(None, k, k, cpi, cpo, 0, None, None)
    assert nbytes > 0
    wp = torch.zeros(nbytes // 2, dtype=torch.bfloat16, device="cuda")
    assert lib.semb_pack_weights_tc(wd.data_ptr(), k, k, cpi, cpo, 0, wp.data_ptr(), U.stream()) == nbytes
    stats = torch.zeros(2 * cpo, device="cuda", dtype=torch.float64)
    xv, yv = U.view(xd, 8, cpi), U.view(yd, 8, cpo)
    L.check(lib.semb_conv2d_fwd_tc(C.byref(geom), C.byref(xv), wp.data_ptr(), bd.data_ptr(), C.byref(yv), stats.data_ptr(), 0, cpo,
                                   0, U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(yd[..., 8:8 + cout], y_ref) < 1e-2
    assert float(yd[..., :8].abs().max()) == 0 and float(yd[..., 8 + cpo:].abs().max()) == 0
    if cpo > cout:
        assert float(yd[..., 8 + cout:8 + cpo].abs().max()) == 0      # padded output channels stay exactly zero
    assert U.rel_err(stats[:cout], y_ref.detach().double().sum(dim=(0, 1, 2))) < 1e-4
    assert U.rel_err(stats[cpo:cpo + cout], (y_ref.detach().double() ** 2).sum(dim=(0, 1, 2))) < 1e-4
    if pad_mode != "zero":
        return
    # data gradient through the SAME kernel with the mirrored / transposed pack, accumulate on top of ones
    wpf = torch.zeros(lib.semb_pack_weights_tc(None, k, k, cpi, cpo, 1, None, None) // 2, dtype=torch.bfloat16, device="cuda")
    assert lib.semb_pack_weights_tc(wd.data_ptr(), k, k, cpi, cpo, 1, wpf.data_ptr(), U.stream()) > 0
    gd = L.ConvGeom(n, h, w_, h, w_, cpo, cpi, k, k, 1, k - 1 - p, k - 1 - p, L.PAD_ZERO, L.BF16)
    dyd = U.to_dev(dy, "bf16")
    dxd = torch.ones((n, h, w_, cpi), dtype=torch.bfloat16, device="cuda")
    dyv, dxv = U.view(dyd), U.view(dxd)
    L.check(lib.semb_conv2d_fwd_tc(C.byref(gd), C.byref(dyv), wpf.data_ptr(), None, C.byref(dxv), None, 0, 0, 1, U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(dxd[..., :cin].float().cpu() - 1.0, xr.grad) < 2e-2


@pytest.mark.parametrize("case", [
    (2, 16, 16, 56, 32),        # resident weights, one N chunk of 128 = the four quadrants
    (2, 16, 8, 24, 40),         # quadrant boundaries (40, 80, 120) inside one N chunk of 160
    (1, 16, 24, 216, 128),      # streamed weights (two-tiles-per-chunk mode), four N chunks = one quadrant each
    (2, 8, 16, 432, 256),       # eight N chunks, two per quadrant
    (1, 20, 12, 51, 13),        # logical channel counts that are not multiples of 8; partial tiles
])
def test_conv_transpose_2x2_fused_depth_to_space(case):
    """Conv2DTranspose(2x2, stride 2, bias) (UNet_Segmentation.py:542-551) as ONE tensor-core launch: 1x1 conv with 4*C virtual
    outputs whose epilogue scatters into a channel slice of the (N, 2H, 2W) skip-concat buffer and adds the bias; against
    oracle.layers.conv2d_transpose on bf16-rounded operands, and bit-identical to nothing else: the neighbouring channels of
    the destination buffer must stay untouched."""
    n, h, w_, cin, co = case
    lib = L.load()
    g = torch.Generator().manual_seed(41)
    x = U.bf16_round(torch.randn(n, h, w_, cin, generator=g))
    wk = U.bf16_round(torch.randn(2, 2, co, cin, generator=g) * 0.1)          # Keras Conv2DTranspose kernel (kh, kw, Cout, Cin)
    bias = torch.randn(co, generator=g)
    y_ref = OL.conv2d_transpose(x, wk, bias, 2)                               # (n, 2h, 2w, co)
    cpi, cpo = U.pad8(cin), U.pad8(co)
    wphys = torch.zeros(1, 1, cpi, 4 * cpo)
    for r in range(2):
        for s_ in range(2):
            wphys[0, 0, :cin, (2 * r + s_) * cpo:(2 * r + s_) * cpo + co] = wk[r, s_].T
    xd = U.to_dev(x, "bf16", pitch=cpi + 8, coff=8)
    geom = L.ConvGeom(n, h, w_, h, w_, cpi, 4 * cpo, 1, 1, 1, 0, 0, L.PAD_ZERO, L.BF16)
    nbytes = lib.semb_pack_weights_tc(None, 1, 1, cpi, 4 * cpo, 0, None, None)
    wp = torch.zeros(nbytes // 2, dtype=torch.bfloat16, device="cuda")
    wd = wphys.cuda().contiguous()
    assert lib.semb_pack_weights_tc(wd.data_ptr(), 1, 1, cpi, 4 * cpo, 0, wp.data_ptr(), U.stream()) == nbytes
    up = torch.full((n, 2 * h, 2 * w_, cpo + 24), 7.0, dtype=torch.bfloat16, device="cuda")     # destination = channels [16, 16 + cpo)
    bd = U.pad_v(bias)
    xv, uv = U.view(xd, 8, cpi), U.view(up, 16, cpo)
    L.check(lib.semb_conv2d_fwd_tc_d2s(C.byref(geom), C.byref(xv), wp.data_ptr(), bd.data_ptr(), C.byref(uv), 2 * h, 2 * w_, U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(up[..., 16:16 + co], y_ref) < 1e-2
    assert float((up[..., :16].float() - 7.0).abs().max()) == 0 and float((up[..., 16 + cpo:].float() - 7.0).abs().max()) == 0
    if cpo > co:
        assert float(up[..., 16 + co:16 + cpo].float().abs().max()) == 0       # padded channels: zero weights + zero bias
    # the unfused pair (1x1 conv -> semb_pixel_shuffle2 with bias) agrees to bf16 rounding of the intermediate
    y4 = torch.zeros((n, h, w_, 4 * cpo), dtype=torch.bfloat16, device="cuda")
    up2 = torch.zeros((n, 2 * h, 2 * w_, cpo), dtype=torch.bfloat16, device="cuda")
    y4v, up2v = U.view(y4), U.view(up2)
    L.check(lib.semb_conv2d_fwd_tc(C.byref(geom), C.byref(xv), wp.data_ptr(), None, C.byref(y4v), None, 0, 0, 0, U.stream()))
    L.check(lib.semb_pixel_shuffle2(C.byref(y4v), C.byref(up2v), n, h, w_, bd.data_ptr(), 0, L.BF16, U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(up[..., 16:16 + cpo], up2.float().cpu()) < 1e-2


@pytest.mark.parametrize("case", [(3, 32, 24, 8, 16, 3), (2, 16, 16, 64, 24, 3), (2, 16, 16, 144, 216, 3), (3, 16, 8, 56, 32, 1)])
def test_conv_tc_per_sample_moments(case):
    """InstanceNorm statistics (one moment pair per sample) from the conv epilogue: register-moment and generic paths,
    sample boundaries inside a CTA's tile sequence."""
    n, h, w_, cin, cout, k = case
    lib = L.load()
    g = torch.Generator().manual_seed(29)
    x = U.bf16_round(torch.randn(n, h, w_, cin, generator=g) + 0.3)
    wt = U.bf16_round(torch.randn(k, k, cin, cout, generator=g) * 0.1)
    y_ref = OL.conv2d(x, wt, None, 1, "same")
    cpi, cpo = U.pad8(cin), U.pad8(cout)
    p = (k - 1) // 2
    geom = L.ConvGeom(n, h, w_, h, w_, cpi, cpo, k, k, 1, p, p, L.PAD_ZERO, L.BF16)
    xd = U.to_dev(x, "bf16")
    yd = torch.zeros((n, h, w_, cpo), dtype=torch.bfloat16, device="cuda")
    wd = U.pad_w(wt)
    wp = torch.zeros(lib.semb_pack_weights_tc(None, k, k, cpi, cpo, 0, None, None) // 2, dtype=torch.bfloat16, device="cuda")
    assert lib.semb_pack_weights_tc(wd.data_ptr(), k, k, cpi, cpo, 0, wp.data_ptr(), U.stream()) > 0
    stats = torch.zeros(n, 2, cpo, device="cuda", dtype=torch.float64)
    xv, yv = U.view(xd), U.view(yd)
    L.check(lib.semb_conv2d_fwd_tc(C.byref(geom), C.byref(xv), wp.data_ptr(), None, C.byref(yv), stats.data_ptr(), 2 * cpo, cpo, 0, U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(yd[..., :cout], y_ref) < 1e-2
    assert U.rel_err(stats[:, 0, :cout], y_ref.double().sum(dim=(1, 2))) < 1e-4
    assert U.rel_err(stats[:, 1, :cout], (y_ref.double() ** 2).sum(dim=(1, 2))) < 1e-4


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_wgrad(case):
    """tcgen05 weight gradient (MN-major operands, split-K over pixel tiles) vs autograd on bf16-rounded x, dy."""
    n, h, w_, cin, cout, k, pad_mode = case
    lib = L.load()
    g = torch.Generator().manual_seed(13)
    x = U.bf16_round(torch.randn(n, h, w_, cin, generator=g))
    wt = torch.randn(k, k, cin, cout, generator=g) * 0.1
    p = (k - 1) // 2
    wr = wt.clone().requires_grad_(True)
    if pad_mode == "reflect":
        y_ref = OL.conv2d(OL.reflection_pad(x, 2 * p, 2 * p), wr, None, 1, "valid")
    else:
        y_ref = OL.conv2d(x, wr, None, 1, "same")
    dy = U.bf16_round(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    cpi, cpo = U.pad8(cin), U.pad8(cout)
    geom = L.ConvGeom(n, h, w_, h, w_, cpi, cpo, k, k, 1, p, p, L.PAD_ZERO, L.BF16)
    xd = U.to_dev(x, "bf16", pitch=cpi + 8, coff=8)
    if pad_mode == "reflect":
        xpad = torch.zeros((n, h + 2 * p, w_ + 2 * p, cpi + 8), dtype=torch.bfloat16, device="cuda")
        sv, dv = U.view(xd, 8, cpi), U.view(xpad, 8, cpi)
        L.check(lib.semb_pad_crop(C.byref(sv), C.byref(dv), n, h, w_, h + 2 * p, w_ + 2 * p, p, p, 0, L.BF16, 0, U.stream()))
        xd = xpad
        geom = L.ConvGeom(n, h + 2 * p, w_ + 2 * p, h, w_, cpi, cpo, k, k, 1, 0, 0, L.PAD_ZERO, L.BF16)
    dyd = U.to_dev(dy, "bf16", pitch=cpo + 8, coff=0)
    dw = torch.ones((k, k, cpi, cpo), device="cuda")          # accumulates on top of existing content
    xv, dyv = U.view(xd, 8, cpi), U.view(dyd, 0, cpo)
    L.check(lib.semb_conv2d_wgrad_tc(C.byref(geom), C.byref(xv), C.byref(dyv), dw.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    got = dw.cpu() - 1.0
    assert U.rel_err(got[:, :, :cin, :cout], wr.grad) < 1e-4
    if cpi > cin:
        assert float(got[:, :, cin:, :].abs().max()) == 0
    if cpo > cout:
        assert float(got[:, :, :, cout:].abs().max()) == 0


# (N,H,W,Cin,Cout,pad): layers with a planar-staged weight gradient (>= 128 channels on both sides)
WGRAD_PLANAR_CASES = [
    (2, 16, 16, 128, 128, 1),
    (1, 20, 12, 144, 216, 1),              # partial tiles, channel chunks that do not divide
    (2, 34, 34, 512, 512, 0),              # the CycleGAN residual conv at its real shape: 'valid' over the reflect-padded 34x34 input (K12)
    (1, 18, 26, 136, 256, 0),
]


@pytest.mark.parametrize("case", WGRAD_PLANAR_CASES)
def test_conv_tc_wgrad_planar_workspace(case):
    """semb_conv2d_wgrad_tc_ws: x / dy re-laid out as planar copies in caller scratch (128-byte TMA rows), same gradient as
    autograd on bf16-rounded operands; sliced source views, accumulation on top of existing content."""
    n, h, w_, cin, cout, p = case
    lib = L.load()
    g = torch.Generator().manual_seed(31)
    x = U.bf16_round(torch.randn(n, h, w_, cin, generator=g))
    wt = torch.randn(3, 3, cin, cout, generator=g) * 0.05
    wr = wt.clone().requires_grad_(True)
    y_ref = OL.conv2d(x, wr, None, 1, "same" if p == 1 else "valid")
    oh, ow = y_ref.shape[1], y_ref.shape[2]
    dy = U.bf16_round(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    geom = L.ConvGeom(n, h, w_, oh, ow, cin, cout, 3, 3, 1, p, p, L.PAD_ZERO, L.BF16)
    need = int(lib.semb_conv2d_wgrad_tc_workspace(C.byref(geom)))
    assert need >= 2 * (x.numel() + dy.numel())
    xd = U.to_dev(x, "bf16", pitch=cin + 8, coff=8)
    dyd = U.to_dev(dy, "bf16", pitch=cout + 16, coff=8)
    ws = torch.empty(need + 256, dtype=torch.uint8, device="cuda")
    base = (ws.data_ptr() + 127) // 128 * 128
    dw = torch.ones((3, 3, cin, cout), device="cuda")
    xv, dyv = U.view(xd, 8, cin), U.view(dyd, 8, cout)
    L.check(lib.semb_conv2d_wgrad_tc_ws(C.byref(geom), C.byref(xv), C.byref(dyv), dw.data_ptr(), base, need, U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(dw.cpu() - 1.0, wr.grad) < 1e-4
    with pytest.raises(L.SembError):
        L.check(lib.semb_conv2d_wgrad_tc_ws(C.byref(geom), C.byref(xv), C.byref(dyv), dw.data_ptr(), base, need - 1, U.stream()))
    small = L.ConvGeom(1, 16, 16, 16, 16, 32, 32, 3, 3, 1, 1, 1, L.PAD_ZERO, L.BF16)
    assert int(lib.semb_conv2d_wgrad_tc_workspace(C.byref(small))) == 0


def test_cyclegan_residual_conv_real_shape_fwd_dgrad():
    """K12 of SURVEY 2.1 at its real shape: 512 -> 512 3x3 over a reflect-padded 8 x 32 x 32 tile batch (CycleGAN.py:323-337),
    forward and data gradient of the TMA / tcgen05 kernel against the oracle's layer (bf16-rounded operands)."""
    n, h, cin, cout = 8, 32, 512, 512
    lib = L.load()
    g = torch.Generator().manual_seed(37)
    x = U.bf16_round(torch.randn(n, h, h, cin, generator=g))
    wt = U.bf16_round(torch.randn(3, 3, cin, cout, generator=g) * 0.02)
    xr = x.clone().requires_grad_(True)
    xpad = OL.reflection_pad(xr, 2, 2)
    y_ref = OL.conv2d(xpad, wt, None, 1, "valid")
    dy = U.bf16_round(torch.randn(y_ref.shape, generator=g))
    xpad.retain_grad()
    y_ref.backward(dy)
    geom = L.ConvGeom(n, h + 2, h + 2, h, h, cin, cout, 3, 3, 1, 0, 0, L.PAD_ZERO, L.BF16)
    xd = U.to_dev(xpad.detach(), "bf16")
    yd = torch.zeros((n, h, h, cout), dtype=torch.bfloat16, device="cuda")
    wdev = wt.cuda().contiguous()
    nbytes = int(lib.semb_pack_weights_tc(None, 3, 3, cin, cout, 0, None, None))
    wp = torch.zeros(nbytes // 2, dtype=torch.bfloat16, device="cuda")
    assert int(lib.semb_pack_weights_tc(wdev.data_ptr(), 3, 3, cin, cout, 0, wp.data_ptr(), U.stream())) >= 0
    stats = torch.zeros(n, 2, cout, device="cuda", dtype=torch.float64)
    xv, yv = U.view(xd), U.view(yd)
    L.check(lib.semb_conv2d_fwd_tc(C.byref(geom), C.byref(xv), wp.data_ptr(), None, C.byref(yv), stats.data_ptr(), 2 * cout, cout, 0, U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(yd.float(), y_ref) < 1e-2
    assert U.rel_err(stats[:, 0], y_ref.double().sum(dim=(1, 2))) < 1e-4          # InstanceNorm moments per (sample, channel)
    # data gradient on the padded domain = full correlation of dy with the mirrored, transposed kernel
    wpb = torch.zeros(nbytes // 2, dtype=torch.bfloat16, device="cuda")
    assert int(lib.semb_pack_weights_tc(wdev.data_ptr(), 3, 3, cin, cout, 1, wpb.data_ptr(), U.stream())) >= 0
    geom_d = L.ConvGeom(n, h, h, h + 2, h + 2, cout, cin, 3, 3, 1, 2, 2, L.PAD_ZERO, L.BF16)
    dyd = U.to_dev(dy, "bf16")
    dxd = torch.zeros((n, h + 2, h + 2, cin), dtype=torch.bfloat16, device="cuda")
    dyv, dxv = U.view(dyd), U.view(dxd)
    L.check(lib.semb_conv2d_fwd_tc(C.byref(geom_d), C.byref(dyv), wpb.data_ptr(), None, C.byref(dxv), None, 0, 0, 0, U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(dxd.float(), xpad.grad) < 1e-2


# ---- stride-2 convs on the stride-1 tensor-core kernels (space-to-depth) ---------------------------------------------
# (n, h, w, cin, cout, k, (pad_t, pad_l), transposed)
S2D_CASES = [
    (2, 32, 32, 16, 24, 3, (0, 0), False),      # generator downsample: 3x3 s2 'same' on an even size (pad 0 before, 1 after)
    (1, 30, 22, 8, 16, 4, (0, 0), False),       # PatchGAN: 4x4 s2 valid
    (2, 31, 27, 13, 8, 4, (0, 0), False),       # odd input size (127 -> 62 in the real discriminator), padded channel lanes
    (2, 16, 16, 24, 16, 3, (1, 1), True),       # generator upsample: Conv2DTranspose 3x3 s2 'same'
    (2, 32, 32, 1, 64, 5, (1, 1), False),       # WGAN critic (WassersteinGAN.py:571-580): 5x5 s2 'same' on an even size, one input channel
    (2, 16, 24, 64, 40, 5, (1, 1), False),      # 5x5 s2 'same': all nine positions of the virtual 3x3 kernel are populated
]


@pytest.mark.parametrize("case", S2D_CASES)
def test_strided_conv_space_to_depth(case):
    """ConvOp in space-to-depth mode (forward, data gradient, weight gradient folded back into the Keras-layout gradient)
    against the oracle's layers (oracle/layers.py) on bf16-rounded operands."""
    import numpy as np
    import torch.nn.functional as F
    from sem_b200.engine import ConvOp, Engine, ParamSpec
    n, h, w_, cin, cout, k, (pt, pl), transposed = case
    g = torch.Generator().manual_seed(17)
    cpi, cpo = U.pad8(cin), U.pad8(cout)
    if not transposed:
        oh, ow = (h + 1) // 2 if k != 4 else (h - 4) // 2 + 1, (w_ + 1) // 2 if k != 4 else (w_ - 4) // 2 + 1
        x = U.bf16_round(torch.randn(n, h, w_, cin, generator=g))
        wt = U.bf16_round(torch.randn(k, k, cin, cout, generator=g) * 0.1)          # Keras HWIO
        xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
        # the oracle's Keras rules (oracle/layers.py: asymmetric 'same' padding for stride 2)
        y_ref = OL.conv2d(xr, wr, None, 2, "same" if k != 4 else "valid")
        assert tuple(y_ref.shape[1:3]) == (oh, ow)
        in_hw, out_hw, lw, pw = (h, w_), (oh, ow), (k, k, cin, cout), (k, k, cpi, cpo)
        maps = {2: np.arange(cin), 3: np.arange(cout)}
    else:
        oh, ow = 2 * h, 2 * w_                                                     # x is the SMALL tensor, cin -> cout channels
        x = U.bf16_round(torch.randn(n, h, w_, cin, generator=g))
        wt = U.bf16_round(torch.randn(k, k, cout, cin, generator=g) * 0.1)          # Keras Conv2DTranspose kernel (kh,kw,Cout,Cin)
        xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
        y_ref = OL.conv2d_transpose(xr, wr, None, 2)       # Keras Conv2DTranspose 'same' 
        in_hw, out_hw, lw, pw = (h, w_), (oh, ow), (k, k, cout, cin), (k, k, cpo, cpi)
        maps = {2: np.arange(cout), 3: np.arange(cin)}
    dy = U.bf16_round(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)

    e = Engine(n, "bf16")
    xb = e.new_buf(in_hw[0], in_hw[1], cpi, "x")
    yb = e.new_buf(out_hw[0], out_hw[1], cpo, "y")
    e.add_param(ParamSpec("w/kernel", "conv_kernel", lw, pw, maps, True, "glorot", (1, 1)))
    op = e.add_op(ConvOp(e, xb.view(), yb.view(), in_hw, out_hw, "w/kernel", None, k, 2, (pt, pl), L.PAD_ZERO, transposed))
    assert op.s2d is not None
    e.finalize()
    e.set_param("w/kernel", wt.numpy())
    xb.data[..., :cin] = x.cuda().to(torch.bfloat16)
    e.zero_step(zero_grads=True)
    e.forward(True)
    torch.cuda.synchronize()
    assert U.rel_err(yb.data[..., :cout].float(), y_ref) < 1e-2
    if cpo > cout:
        assert float(yb.data[..., cout:].abs().max()) == 0
    yb.grad_tensor()[..., :cout] = dy.cuda().to(torch.bfloat16)
    xb.grad_tensor().fill_(1.0)
    op.acc_x = 1                                   # accumulate on top of ones
    e.backward()
    e.fold_virtual_grads()
    torch.cuda.synchronize()
    assert U.rel_err(xb.grad_tensor()[..., :cin].float().cpu() - 1.0, xr.grad) < 2e-2
    assert U.rel_err(torch.from_numpy(e.get_grad("w/kernel")), wr.grad) < 1e-3


# ---- 7x7 reflect-padded convs with one input / one output channel as 1x1 tensor-core convs ------------------------------
@pytest.mark.parametrize("case", [(2, 20, 16, 1, 24, False), (1, 18, 26, 16, 1, True), (2, 16, 16, 64, 1, True)])
def test_conv7x7_tap_folding(case):
    """Generator stem (1 -> F) and head (F -> 1, bias) of CycleGAN.py:372,393 through ConvOp's tap-folded tensor-core path."""
    import numpy as np
    from sem_b200.engine import ConvOp, Engine, ParamSpec
    n, h, w_, cin, cout, has_bias = case
    g = torch.Generator().manual_seed(23)
    x = U.bf16_round(torch.randn(n, h, w_, cin, generator=g))
    wt = U.bf16_round(torch.randn(7, 7, cin, cout, generator=g) * 0.1)
    bias = torch.randn(cout, generator=g) if has_bias else None
    xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True) if has_bias else None
    y_ref = OL.conv2d(OL.reflection_pad(xr, 6, 6), wr, br, 1, "valid")
    dy = U.bf16_round(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    cpi, cpo = U.pad8(cin), U.pad8(cout)
    e = Engine(n, "bf16")
    xb = e.new_buf(h, w_, cpi, "x")
    yb = e.new_buf(h, w_, cpo, "y")
    e.add_param(ParamSpec("w/kernel", "conv_kernel", (7, 7, cin, cout), (7, 7, cpi, cpo), {2: np.arange(cin), 3: np.arange(cout)}, True,
                          "glorot", (1, 1)))
    if has_bias:
        e.add_param(ParamSpec("w/bias", "vector", (cout,), (cpo,), {0: np.arange(cout)}, True, "zeros"))
    op = e.add_op(ConvOp(e, xb.view(), yb.view(), (h, w_), (h, w_), "w/kernel", "w/bias" if has_bias else None, 7, 1, (3, 3),
                         L.PAD_REFLECT, False))
    assert op.tapfold is not None
    e.finalize()
    e.set_param("w/kernel", wt.numpy())
    if has_bias:
        e.set_param("w/bias", bias.numpy())
    xb.data[..., :cin] = x.cuda().to(torch.bfloat16)
    e.zero_step(zero_grads=True)
    e.forward(True)
    torch.cuda.synchronize()
    assert U.rel_err(yb.data[..., :cout].float(), y_ref) < 1e-2
    if cpo > cout:
        assert float(yb.data[..., cout:].abs().max()) == 0
    yb.grad_tensor()[..., :cout] = dy.cuda().to(torch.bfloat16)
    xb.grad_tensor().fill_(1.0)
    op.acc_x = 1
    e.backward()
    e.fold_virtual_grads()
    torch.cuda.synchronize()
    assert U.rel_err(xb.grad_tensor()[..., :cin].float().cpu() - 1.0, xr.grad) < 2e-2
    assert U.rel_err(torch.from_numpy(e.get_grad("w/kernel")), wr.grad) < 1e-3
    if has_bias:
        assert U.rel_err(torch.from_numpy(e.get_grad("w/bias")), br.grad) < 1e-3


@pytest.mark.parametrize("case", [(2, 30, 30, 64, True), (1, 13, 17, 24, True), (2, 12, 12, 512, False)])
def test_patchgan_output_conv_tap_folding(case):
    """The PatchGAN output conv (4x4, stride 1, 'valid', one output channel, bias; CycleGAN.py:448) through ConvOp's
    tap-folded tensor-core path (per-tap 1x1 conv on the un-padded input + shift-and-add) against the oracle."""
    import numpy as np
    from sem_b200.engine import ConvOp, Engine, ParamSpec
    n, h, w_, cin, has_bias = case
    k, cout = 4, 1
    g = torch.Generator().manual_seed(29)
    x = U.bf16_round(torch.randn(n, h, w_, cin, generator=g))
    wt = U.bf16_round(torch.randn(k, k, cin, cout, generator=g) * 0.05)
    bias = torch.randn(cout, generator=g) if has_bias else None
    xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True) if has_bias else None
    y_ref = OL.conv2d(xr, wr, br, 1, "valid")
    oh, ow = h - 3, w_ - 3
    assert tuple(y_ref.shape[1:3]) == (oh, ow)
    dy = U.bf16_round(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    cpi, cpo = U.pad8(cin), 8
    e = Engine(n, "bf16")
    xb = e.new_buf(h, w_, cpi, "x")
    yb = e.new_buf(oh, ow, cpo, "y")
    e.add_param(ParamSpec("w/kernel", "conv_kernel", (k, k, cin, cout), (k, k, cpi, cpo), {2: np.arange(cin), 3: np.arange(cout)}, True,
                          "glorot", (1, 1)))
    if has_bias:
        e.add_param(ParamSpec("w/bias", "vector", (cout,), (cpo,), {0: np.arange(cout)}, True, "zeros"))
    op = e.add_op(ConvOp(e, xb.view(), yb.view(), (h, w_), (oh, ow), "w/kernel", "w/bias" if has_bias else None, k, 1, (0, 0),
                         L.PAD_ZERO, False))
    assert op.tapfold is not None and op.tf_valid
    e.finalize()
    e.set_param("w/kernel", wt.numpy())
    if has_bias:
        e.set_param("w/bias", bias.numpy())
    xb.data[..., :cin] = x.cuda().to(torch.bfloat16)
    e.zero_step(zero_grads=True)
    e.forward(True)
    torch.cuda.synchronize()
    assert U.rel_err(yb.data[..., :cout].float(), y_ref) < 1e-2
    assert float(yb.data[..., cout:].abs().max()) == 0
    yb.grad_tensor()[..., :cout] = dy.cuda().to(torch.bfloat16)
    xb.grad_tensor().fill_(1.0)
    op.acc_x = 1
    e.backward()
    e.fold_virtual_grads()
    torch.cuda.synchronize()
    assert U.rel_err(xb.grad_tensor()[..., :cin].float().cpu() - 1.0, xr.grad) < 2e-2
    assert U.rel_err(torch.from_numpy(e.get_grad("w/kernel")), wr.grad) < 2e-3
    if has_bias:
        assert U.rel_err(torch.from_numpy(e.get_grad("w/bias")), br.grad) < 1e-3


@pytest.mark.parametrize("case", [(712, 1024, 256, 256, 2), (300, 500, 128, 96, 2), (257, 255, 128, 128, 10), (100, 90, 128, 128, 2), (512, 768, 256, 256, 0)])
def test_device_tile_gather_and_stitch_match_the_host_grid(case):
    """semb_tile_gather / semb_tile_stitch against HelperFunctions.tile_image / stitch_image (reference :17-141): bit-exact,
    all three overlap modes, images smaller than a tile, exact fits (extra tile column rule)."""
    from sem_b200 import HelperFunctions as HF
    H, W, th, tw, mo = case
    lib = L.load()
    rng = np.random.default_rng(H * 7 + W)
    img = rng.random((H, W, 1), dtype=np.float32)
    tiles_ref = HF.tile_image(img, tw, th, min_overlap=mo)
    nx, xs = HF._grid(W, tw, mo)
    ny, ys = HF._grid(H, th, mo)
    nt = nx * ny
    assert tiles_ref.shape[0] == nt
    img_d = torch.from_numpy(img[:, :, 0].copy()).cuda()
    xs_d, ys_d = torch.tensor(xs, dtype=torch.int32, device="cuda"), torch.tensor(ys, dtype=torch.int32, device="cuda")
    got = torch.full((nt, th, tw), -1.0, device="cuda")
    half = max(nt // 2, 1)
    L.check(lib.semb_tile_gather(img_d.data_ptr(), H, W, got.data_ptr(), th, tw, xs_d.data_ptr(), nx, ys_d.data_ptr(), ny, 0, half, U.stream()))
    if nt > half:
        L.check(lib.semb_tile_gather(img_d.data_ptr(), H, W, got[half:].data_ptr(), th, tw, xs_d.data_ptr(), nx, ys_d.data_ptr(), ny, half,
                                     nt - half, U.stream()))
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), tiles_ref[..., 0])
    pred = rng.random((nt, th, tw, 1), dtype=np.float32)
    pred_d = torch.from_numpy(pred[..., 0].copy()).cuda()
    for mode in (0, 1, 2):
        ref = HF.stitch_image(pred, W, H, min_overlap=mo, manage_overlap_mode=mode)[:, :, 0]
        out = torch.full((H, W), -1.0, device="cuda")
        L.check(lib.semb_tile_stitch(pred_d.data_ptr(), th, tw, out.data_ptr(), H, W, xs_d.data_ptr(), nx, ys_d.data_ptr(), ny, mode, U.stream()))
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        if mode == 1:
            assert np.abs(o - ref).max() < 1e-6, mode
        else:
            assert np.array_equal(o, ref), (mode, int((o != ref).sum()))


@pytest.mark.parametrize("case", [(3, 20, 20, 16, 24), (8, 32, 32, 128, 128), (1, 16, 24, 40, 8)])
def test_reflect_padded_conv_op_bf16(case):
    """ConvOp for ReflectionPadding2D(1) + 3x3 'valid' (CycleGAN.py:326-333) in bf16: materialised padding, TMA forward,
    data gradient on the padded domain, reflect fold, weight gradient (planar-staged for >= 128 channels) -- against the
    oracle."""
    import numpy as np
    from sem_b200.engine import ConvOp, Engine, ParamSpec
    n, h, w_, cin, cout = case
    g = torch.Generator().manual_seed(41)
    x = U.bf16_round(torch.randn(n, h, w_, cin, generator=g))
    wt = U.bf16_round(torch.randn(3, 3, cin, cout, generator=g) * 0.05)
    xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    y_ref = OL.conv2d(OL.reflection_pad(xr, 2, 2), wr, None, 1, "valid")
    dy = U.bf16_round(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    e = Engine(n, "bf16")
    xb = e.new_buf(h, w_, cin, "x")
    yb = e.new_buf(h, w_, cout, "y")
    e.add_param(ParamSpec("w/kernel", "conv_kernel", (3, 3, cin, cout), (3, 3, cin, cout), {2: np.arange(cin), 3: np.arange(cout)}, True,
                          "glorot", (1, 1)))
    op = e.add_op(ConvOp(e, xb.view(), yb.view(), (h, w_), (h, w_), "w/kernel", None, 3, 1, (1, 1), L.PAD_REFLECT, False))
    e.finalize()
    assert op.x_pad is not None and op.pad_buf is not None
    e.set_param("w/kernel", wt.numpy())
    xb.data.copy_(x.cuda().to(torch.bfloat16))
    e.zero_step(zero_grads=True)
    e.forward(True)
    torch.cuda.synchronize()
    assert U.rel_err(yb.data.float(), y_ref) < 1e-2
    yb.grad_tensor().copy_(dy.cuda().to(torch.bfloat16))
    xb.grad_tensor().fill_(1.0)
    op.acc_x = 1
    e.backward()
    torch.cuda.synchronize()
    assert U.rel_err(xb.grad_tensor().float().cpu() - 1.0, xr.grad) < 2e-2
    assert U.rel_err(torch.from_numpy(e.get_grad("w/kernel")), wr.grad) < 1e-3


@pytest.mark.parametrize("case", [(2, 24, 40, 16, 16, 16), (3, 16, 16, 32, 32, 32), (1, 20, 12, 8, 8, 24)])
def test_pair_conv_op_bf16(case):
    """PairConvOp: the 3x3 conv and the 1x1 shortcut conv of a res_path unit (UNet_Segmentation.py:490-499) as ONE 3x3 conv
    over the merged virtual kernel (semb_merge_weights): both outputs, their batch moments in the shared record, the data
    gradient of the pair and the two Keras-layout weight gradients after the fold -- against the oracle's two convs."""
    import numpy as np
    from sem_b200.engine import Engine, PairConvOp, ParamSpec
    n, h, w_, cin, ca, cs = case
    g = torch.Generator().manual_seed(7)
    x = U.bf16_round(torch.randn(n, h, w_, cin, generator=g))
    wa = U.bf16_round(torch.randn(3, 3, cin, ca, generator=g) * 0.1)
    ws = U.bf16_round(torch.randn(1, 1, cin, cs, generator=g) * 0.2)
    xr, war, wsr = x.clone().requires_grad_(True), wa.clone().requires_grad_(True), ws.clone().requires_grad_(True)
    ya = OL.conv2d(xr, war, None, 1, "same")
    ys = OL.conv2d(xr, wsr, None, 1, "same")
    y_ref = torch.cat([ya, ys], dim=-1)
    dy = U.bf16_round(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    e = Engine(n, "bf16")
    xb = e.new_buf(h, w_, cin, "x")
    yb = e.new_buf(h, w_, ca + cs, "y")
    ident = lambda c: np.arange(c)
    e.add_param(ParamSpec("a/kernel", "conv_kernel", (3, 3, cin, ca), (3, 3, cin, ca), {2: ident(cin), 3: ident(ca)}, True, "glorot", (1, 1)))
    e.add_param(ParamSpec("s/kernel", "conv_kernel", (1, 1, cin, cs), (1, 1, cin, cs), {2: ident(cin), 3: ident(cs)}, True, "glorot", (1, 1)))
    mom = e.zeroed.add("pair/moments", 4 * (ca + cs))
    op = e.add_op(PairConvOp(e, xb.view(), yb.view(), (h, w_), "a/kernel", "s/kernel", ca, cs, stats=(mom, 0, 0, ca + cs)))
    e.finalize()
    e.set_param("a/kernel", wa.numpy())
    e.set_param("s/kernel", ws.numpy())
    xb.data.copy_(x.cuda().to(torch.bfloat16))
    e.zero_step(zero_grads=True)
    e.forward(True)
    torch.cuda.synchronize()
    assert U.rel_err(yb.data.float(), y_ref) < 1e-2
    o, cnt = e.zeroed.entries[mom]
    st = e.zeroed.t[o:o + cnt].view(torch.float64).cpu()
    yq = yb.data.float().cpu()           # moments are taken from the fp32 accumulators, compare with the oracle's fp32 result
    assert U.rel_err(st[:ca + cs].float(), y_ref.detach().sum(dim=(0, 1, 2))) < 2e-3, (st[:4], yq.sum(dim=(0, 1, 2))[:4])
    assert U.rel_err(st[ca + cs:].float(), (y_ref.detach() ** 2).sum(dim=(0, 1, 2))) < 2e-3
    yb.grad_tensor().copy_(dy.cuda().to(torch.bfloat16))
    e.backward()
    e.fold_virtual_grads()
    torch.cuda.synchronize()
    assert op.acc_x == 0
    assert U.rel_err(xb.grad_tensor().float().cpu(), xr.grad) < 2e-2
    assert U.rel_err(torch.from_numpy(e.get_grad("a/kernel")), war.grad) < 1e-3
    assert U.rel_err(torch.from_numpy(e.get_grad("s/kernel")), wsr.grad) < 1e-3
    # a second fold must not add anything (the virtual gradient buffer is zero again)
    e.fold_virtual_grads()
    torch.cuda.synchronize()
    assert U.rel_err(torch.from_numpy(e.get_grad("s/kernel")), wsr.grad) < 1e-3
