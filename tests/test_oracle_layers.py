"""Layer algebra of the oracle against naive numpy loops."""
import numpy as np
import torch

from oracle import layers as OL, cyclegan as OC


def naive_conv(x, w, stride, pt, pl, oh, ow):
    n, h, wd, ci = x.shape
    k = w.shape[0]
    co = w.shape[3]
    y = np.zeros((n, oh, ow, co))
    for oy in range(oh):
        for ox in range(ow):
            for r in range(k):
                for s in range(k):
                    iy, ix = oy * stride - pt + r, ox * stride - pl + s
                    if 0 <= iy < h and 0 <= ix < wd:
                        y[:, oy, ox, :] += x[:, iy, ix, :] @ w[r, s]
    return y


def test_conv_same_valid_and_strided_same():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 9, 8, 3))
    w = rng.standard_normal((3, 3, 3, 5))
    xt, wt = torch.from_numpy(x), torch.from_numpy(w)
    assert np.allclose(OL.conv2d(xt, wt, None, 1, "same").numpy(), naive_conv(x, w, 1, 1, 1, 9, 8))
    assert np.allclose(OL.conv2d(xt, wt, None, 1, "valid").numpy(), naive_conv(x, w, 1, 0, 0, 7, 6))
    # stride 2 'same': even size pads (0,1), odd size pads (1,1)  (Appendix B item 2)
    assert OL.same_pad_amounts(8, 3, 2) == (0, 1) and OL.same_pad_amounts(9, 3, 2) == (1, 1)
    assert np.allclose(OL.conv2d(xt, wt, None, 2, "same").numpy(), naive_conv(x, w, 2, 1, 0, 5, 4))
    w4 = rng.standard_normal((4, 4, 3, 2))
    assert np.allclose(OL.conv2d(xt, torch.from_numpy(w4), None, 2, "valid").numpy(), naive_conv(x, w4, 2, 0, 0, 3, 3))
    # the WGAN critic's 5x5 stride-2 'same' (WassersteinGAN.py:571-614): total pad 3 on even sizes (1 before, 2 after), 4 on odd sizes
    assert OL.same_pad_amounts(8, 5, 2) == (1, 2) and OL.same_pad_amounts(9, 5, 2) == (2, 2)
    w5 = rng.standard_normal((5, 5, 3, 2))
    assert np.allclose(OL.conv2d(xt, torch.from_numpy(w5), None, 2, "same").numpy(), naive_conv(x, w5, 2, 2, 1, 5, 4))


def test_conv_transpose_2x2_and_3x3():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((1, 4, 5, 3))
    for k in (2, 3):
        w = rng.standard_normal((k, k, 2, 3))          # (kh,kw,Cout,Cin)
        p, op = OL.conv_transpose_pads(k, 2)
        assert (p, op) == ((0, 0) if k == 2 else (1, 1))
        y = np.zeros((1, 8, 10, 2))
        for iy in range(4):
            for ix in range(5):
                for r in range(k):
                    for s in range(k):
                        oy, ox = iy * 2 - p + r, ix * 2 - p + s
                        if 0 <= oy < 8 and 0 <= ox < 10:
                            y[:, oy, ox, :] += x[:, iy, ix, :] @ w[r, s].T
        assert np.allclose(OL.conv2d_transpose(torch.from_numpy(x), torch.from_numpy(w), None, 2).numpy(), y)


def test_norms_pad_pool():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3, 6, 4, 5)) * 2 + 1
    xt = torch.from_numpy(x)
    y, m, v = OL.batch_norm(xt, None, torch.zeros(5, dtype=torch.float64), torch.zeros(5, dtype=torch.float64),
                            torch.ones(5, dtype=torch.float64), True)
    mean, var = x.mean((0, 1, 2)), x.var((0, 1, 2))
    assert np.allclose(y.numpy(), (x - mean) / np.sqrt(var + 1e-3))
    assert np.allclose(m.numpy(), 0.01 * mean) and np.allclose(v.numpy(), 0.99 + 0.01 * var)     # biased variance
    yi = OL.instance_norm(xt, torch.ones(5, dtype=torch.float64), torch.zeros(5, dtype=torch.float64))
    assert np.allclose(yi.numpy(), (x - x.mean((1, 2), keepdims=True)) / np.sqrt(x.var((1, 2), keepdims=True) + 1e-5))
    assert np.allclose(OL.reflection_pad(xt, 5, 3).numpy(), np.pad(x, ((0, 0), (1, 2), (2, 3), (0, 0)), mode="reflect"))
    assert np.allclose(OL.max_pool_2x2(xt).numpy(), x.reshape(3, 3, 2, 2, 2, 5).max(axis=(2, 4)))


def test_losses_and_adam():
    rng = np.random.default_rng(3)
    p = rng.uniform(0, 1, (2, 4, 4, 1))
    p[0, 0, 0, 0] = 0.0
    y = (rng.uniform(0, 1, (2, 4, 4, 1)) < 0.3).astype(np.float64)
    pc = np.clip(p, 1e-7, 1 - 1e-7)
    ref = (-(y * np.log(pc) + (1 - y) * np.log(1 - pc)) * (y * 2.5 + 1)).mean()
    assert np.isclose(float(OL.weighted_bce(torch.from_numpy(y), torch.from_numpy(p), 3.5)), ref)
    w = torch.tensor([1.0, -2.0], dtype=torch.float64)
    opt = OL.KerasAdam([w], lr=0.1, beta_1=0.9)
    g = torch.tensor([0.5, -0.25], dtype=torch.float64)
    opt.apply([g], [w])
    m, v = 0.1 * g, 0.001 * g * g
    alpha = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(w.numpy(), np.array([1.0, -2.0]) - alpha * m.numpy() / (np.sqrt(v.numpy()) + 1e-7))


def test_image_pool_quirk_and_determinism():
    import random
    pool = OC.ImagePool(batch_size=2, pool_size=3, rng=random.Random(0))
    out = pool.query(torch.arange(5.0).reshape(5, 1, 1, 1))
    assert out.shape[0] == 2          # only the first batch_size images are ever looked at
    pool.query(torch.arange(5.0, 10.0).reshape(5, 1, 1, 1))
    assert pool.num_imgs == 3
    a = [float(pool.query(torch.full((2, 1, 1, 1), float(i))).sum()) for i in range(20)]
    pool2 = OC.ImagePool(batch_size=2, pool_size=3, rng=random.Random(0))
    pool2.query(torch.arange(5.0).reshape(5, 1, 1, 1)); pool2.query(torch.arange(5.0, 10.0).reshape(5, 1, 1, 1))
    b = [float(pool2.query(torch.full((2, 1, 1, 1), float(i))).sum()) for i in range(20)]
    assert a == b
