"""CycleGAN generator / discriminator / train-step parity against the oracle (oracle/cyclegan.py)."""
import random

import numpy as np
import pytest
import torch

from oracle import cyclegan as OC
from tests import util as U
import sem_b200
from sem_b200 import CycleGanModel, ImagePool

pytestmark = pytest.mark.gpu


def _setup(dtype, filters=8, n_res=2, n=2, size=64, use_tc=True):
    tr = OC.CycleGanTrainer(filters=filters, n_res=n_res, pool_batch=2, seed=0)
    m = CycleGanModel((size, size, 1), batch_size=n, filters=filters, dtype=dtype, n_res=n_res,
                      image_pool_a=ImagePool(2, 50, random.Random(0)), image_pool_b=ImagePool(2, 50, random.Random(1)), use_tc=use_tc)
    for name, net in m.nets.items():
        net.set_named({k: v.detach().numpy() for k, v in tr.nets[name].items()})
    g = torch.Generator().manual_seed(5)
    a = torch.rand(n, size, size, 1, generator=g) * 2 - 1
    b = torch.rand(n, size, size, 1, generator=g) * 2 - 1
    return tr, m, a, b


def test_generator_forward_f32():
    tr, m, a, b = _setup("f32")
    with torch.no_grad():
        ref = tr.G("gen_a", a)
    out = m.generate("gen_a", a.numpy())
    assert U.rel_err(torch.from_numpy(out), ref) < 1e-3
    w = m.gen_a.get_weights()
    assert sum(x.size for x in w) == sum(v.numel() for v in tr.nets["gen_a"].values())


@pytest.mark.parametrize("dtype,tol,gtol", [("f32", 1e-3, 2e-3), ("bf16", 5e-2, None)])
def test_train_step_matches_oracle(dtype, tol, gtol):
    tr, m, a, b = _setup(dtype)
    ref, fa_ref, fb_ref = tr.train_step(a, b)
    logs = m.train_step((a.numpy(), b.numpy()))
    print(dtype, {k: (round(logs[k], 5), round(ref[k], 5)) for k in logs})
    for k in ref:
        assert abs(logs[k] - ref[k]) < tol * max(abs(ref[k]), 1e-2), (k, logs[k], ref[k])
    if gtol is None:
        return
    # fp64 evaluation of the same oracle: per-tensor sensitivity of the gradient to fp32 rounding (ReLU / LeakyReLU
    # decisions at |x| ~ 1e-7); see tests/test_unet_gpu.py for the rationale of the slack term.
    tr64 = OC.CycleGanTrainer(filters=8, n_res=2, pool_batch=2, seed=0, dtype=torch.float64)
    tr64.train_step(a.double(), b.double())
    for name, net in m.nets.items():
        gmax = max(float(v.abs().max()) for v in tr.last_grads[name].values())
        for pname in net.names:
            ref_g = tr.last_grads[name][pname]
            ref64 = tr64.last_grads[name][pname].float()
            got = torch.from_numpy(net.root.get_grad(pname))
            den = max(float(ref_g.abs().max()), 1e-3 * gmax)
            err = min(float((got - ref_g).abs().max()), float((got - ref64).abs().max())) / den
            slack = 2.0 * float((ref_g - ref64).abs().max()) / den
            assert err - slack < gtol, (name, pname, err, slack)
    # post-Adam weights
    for name, net in m.nets.items():
        for pname, wt in zip(net.names, net.get_weights()):
            ref_w = tr.nets[name][pname].detach()
            ref64 = tr64.nets[name][pname].detach().float()
            err = min(float((torch.from_numpy(wt) - ref_w).abs().max()), float((torch.from_numpy(wt) - ref64).abs().max()))
            assert err < 1e-3 * float(ref_w.abs().max()) + 2e-5 + 2.0 * float((ref_w - ref64).abs().max()), (name, pname)


def test_second_step_uses_pool_and_runs():
    tr, m, a, b = _setup("bf16")
    l1 = m.train_step((a.numpy(), b.numpy()))
    l2 = m.train_step((a.numpy(), b.numpy()))
    assert all(np.isfinite(v) for v in l2.values())
    assert m.pool_a.num_imgs == 4 and l2["g_cyc_a"] != l1["g_cyc_a"]


# ---- round 2: the options of the class defaults (CycleGAN.py:76-87), initialisation, pool swap branch --------------------
def _gen_pair(options, filters=8, n_res=1, n=2, size=32, dtype="f32"):
    """(oracle params, engine GeneratorModel) with identical weights."""
    from sem_b200.cyclegan_model import GeneratorModel
    spec = OC.generator_spec(filters, n_res=n_res, **options)
    p = OC.init_params(spec, torch.Generator().manual_seed(3))
    p["head/bias"] = torch.full((1,), 0.05)
    for k in p:
        if k.endswith("/beta"):
            p[k] = 0.1 * torch.randn(p[k].shape, generator=torch.Generator().manual_seed(len(k)))
    g = GeneratorModel((size, size, 1), n, filters=filters, dtype=dtype, n_res=n_res, **options)
    assert g.net.names == [nm for nm, _, _ in spec]
    g.set_named({k: v.numpy() for k, v in p.items()})
    return p, g


@pytest.mark.parametrize("options", [{"use_skip_connection": True}, {"use_resize_convolution": True},
                                     {"use_skip_connection": True, "use_resize_convolution": True}])
def test_generator_options_forward_and_gradients_f32(options):
    """Skip-connection branch (CycleGAN.py:396-415) and resize-convolution upsampling (:348-351): forward and every
    parameter gradient of an L1 loss against the oracle."""
    from sem_b200 import _lib as L
    import ctypes as C
    n, size = 2, 32
    p, g = _gen_pair(options, n=n, size=size)
    x = torch.rand(n, size, size, 1, generator=torch.Generator().manual_seed(9)) * 2 - 1
    t = torch.rand(n, size, size, 1, generator=torch.Generator().manual_seed(10)) * 2 - 1
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ref = OC.generator_forward(x, pr, n_res=1)
    (ref - t).abs().mean().backward()
    out = g(x.numpy())
    assert U.rel_err(torch.from_numpy(out), ref) < 1e-3
    # backward through the engine: L1 against t
    e, b = g.eng, g.b
    tbuf = e.new_buf(size, size, 8, "target", requires_grad=False)
    tdev = t.cuda().contiguous()
    L.check(e.lib.semb_cast_in(tdev.data_ptr(), 1, C.byref(tbuf.view().t), n * size * size, e.dtype, e.stream))
    sums = torch.zeros(4, device="cuda")
    e.zero_step(True)
    e.forward(True)
    npix = n * size * size
    L.check(e.lib.semb_loss_l1_l2(C.byref(b.out_buf.view().t), C.byref(tbuf.view().t), 0.0, 0, npix, 1, 1.0 / npix,
                                  C.byref(b.out_buf.view().g), 0, sums.data_ptr(), e.dtype, e.stream))
    e.backward()
    e.fold_virtual_grads()
    torch.cuda.synchronize()
    gmax = max(float(v.grad.abs().max()) for v in pr.values())
    for name in g.net.names:
        got = torch.from_numpy(e.get_grad(name))
        rg = pr[name].grad
        den = max(float(rg.abs().max()), 1e-3 * gmax)
        assert float((got - rg).abs().max()) / den < 5e-3, (name, float((got - rg).abs().max()) / den)


def test_discriminator_gaussian_noise_matches_oracle_with_the_same_deviates():
    """GaussianNoise(0.15) in front of every PatchGAN conv (CycleGAN.py:427-447): the engine's deviates are read back and
    fed to the oracle; the forward must then agree, and eval mode must be noise-free."""
    from sem_b200.engine import Engine
    from sem_b200.gan_nets import DiscriminatorBuilder
    n, size, f = 2, 64, 16
    spec = OC.discriminator_spec(f)
    p = OC.init_params(spec, torch.Generator().manual_seed(4))
    e = Engine(n, "f32")
    b = DiscriminatorBuilder(e, size, size, f, gaussian_noise=0.15)
    e.finalize()
    for k, v in p.items():
        e.set_param(k, v.numpy())
    x = torch.rand(n, size, size, 1, generator=torch.Generator().manual_seed(2)) * 2 - 1
    b.in_buf.data.zero_()
    b.in_buf.data[..., :1] = x.cuda()
    e.zero_step(False)
    e.forward(True)
    torch.cuda.synchronize()
    assert len(b.noise_ops) == 4
    noise = []
    for op, c in zip(b.noise_ops, (1, f, 2 * f, 4 * f)):
        nz = op.noise.data.float().cpu()
        assert float(nz[..., c:].abs().max() if nz.shape[-1] > c else 0.0) == 0.0      # padded lanes stay zero
        assert 0.12 < float(nz[..., :c].std()) < 0.18
        noise.append(nz[..., :c])
    ref = OC.discriminator_forward(x, p, noise=noise)
    got = b.out_buf.data[..., :1].float().cpu()
    assert U.rel_err(got, ref) < 1e-3
    for op in b.noise_ops:
        op.frozen = True
    e.zero_step(False)
    e.forward(True)
    torch.cuda.synchronize()
    assert torch.equal(b.out_buf.data[..., :1].float().cpu(), got)            # frozen deviates: reproducible
    # (InstanceNorm has no inference mode, so a discriminator is never run with training=False: the noise-free path is
    # checked by switching the deviates off)
    for op in b.noise_ops:
        op.noise.data.zero_()
    e.zero_step(False)
    e.forward(True)
    torch.cuda.synchronize()
    ref0 = OC.discriminator_forward(x, p)
    assert U.rel_err(b.out_buf.data[..., :1].float().cpu(), ref0) < 1e-3


def test_default_constructed_model_is_initialised_and_trains():
    """weights used to stay zero unless a test called set_named.  Class defaults of CycleGAN
    (use_skip_connection=True, gaussian_noise_value=0.15) must build and produce non-zero gradients."""
    m = CycleGanModel((64, 64, 1), batch_size=2, filters=8, n_res=1, dtype="bf16", use_skip_connection=True, gaussian_noise_value=0.15)
    for name, net in m.nets.items():
        for pname, w in zip(net.names, net.get_weights()):
            if pname.endswith("/kernel"):
                assert np.abs(w).max() > 0, (name, pname)
            if pname.endswith("/gamma"):
                assert np.all(w == 1.0), (name, pname)
    k0 = {name: net.root.get_param(net.names[0]).copy() for name, net in m.nets.items()}
    assert not np.array_equal(k0["gen_a"], k0["gen_b"])          # one seed per network
    a = np.random.default_rng(0).uniform(-1, 1, (2, 64, 64, 1)).astype(np.float32)
    b = np.random.default_rng(1).uniform(-1, 1, (2, 64, 64, 1)).astype(np.float32)
    logs = m.train_step((a, b))
    assert all(np.isfinite(v) for v in logs.values())
    for name, net in m.nets.items():
        assert np.abs(net.root.get_grad(net.names[0])).max() > 0, name
        assert not np.array_equal(net.root.get_param(net.names[0]), k0[name]), name


def test_cyclegan_facade_default_options_build():
    import tempfile
    from sem_b200 import CycleGAN as CG
    with tempfile.TemporaryDirectory() as d:
        g = CG.CycleGAN(root_dir=d, image_shape=(64, 64, 1))
        g.filters, g.num_residual_blocks_gen = 8, 1
        m = g.create_model()           # class defaults: skip connection + Gaussian noise
        assert any("skip_out" in n for n in m.gen_a.names)
        g.use_binary_crossentropy = True
        with pytest.raises(NotImplementedError):
            g.create_model()
