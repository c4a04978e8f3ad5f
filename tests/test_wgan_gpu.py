"""WGAN-GP (WassersteinGAN.py) on the CUDA path against oracle/wgan.py, all randomness injected:
latent vectors, interpolation factors and Dropout keep masks are the SAME tensors on both sides."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import sem_b200  # noqa: F401
from sem_b200 import _lib as L
from oracle import wgan as OW
from tests import util as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_mask_mul_and_gp_direction_kernels(dtype):
    lib = L.load()
    g = torch.Generator().manual_seed(3)
    n, h, w, c = 3, 5, 7, 24
    rnd = (lambda t: U.bf16_round(t)) if dtype == "bf16" else (lambda t: t)
    x, z = rnd(torch.randn(n, h, w, c, generator=g)), rnd(torch.randn(n, h, w, c, generator=g))
    m = (torch.rand(n, h, w, c, generator=g) >= 0.3).float() / 0.7
    m = rnd(m)
    y0 = rnd(torch.randn(n, h, w, c, generator=g))
    xd, zd, md = U.to_dev(x, dtype, pitch=c + 8, coff=8), U.to_dev(z, dtype), U.to_dev(m, dtype)
    yd = U.to_dev(y0, dtype, pitch=c + 16, coff=8)
    xv, zv, mv, yv = U.view(xd, 8, c), U.view(zd), U.view(md), U.view(yd, 8, c)
    ref = x * torch.where(z > 0, torch.ones_like(z), torch.full_like(z, 0.2)) * m
    tol = 1e-2 if dtype == "bf16" else 1e-6
    L.check(lib.semb_mask_mul(C.byref(xv), C.byref(zv), C.byref(mv), C.byref(yv), n * h * w, 0.2, 1, U.ldtype(dtype), U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(yd[..., 8:8 + c], ref + y0) < tol                      # accumulate on top of y0
    assert float(yd[..., :8].float().abs().max()) == 0 and float(yd[..., 8 + c:].float().abs().max()) == 0
    L.check(lib.semb_mask_mul(C.byref(xv), None, None, C.byref(yv), n * h * w, 0.2, 0, U.ldtype(dtype), U.stream()))
    torch.cuda.synchronize()
    assert U.rel_err(yd[..., 8:8 + c], x) < tol                            # no activation, no mask: a copy
    # gradient-penalty direction
    gx = rnd(torch.randn(n, h, w, 8, generator=g) * 0.3)
    gd, ud = U.to_dev(gx, dtype), torch.zeros((n, h, w, 8), dtype=U.tdtype(dtype), device="cuda")
    sums = torch.zeros(4, device="cuda")
    gv, uv = U.view(gd), U.view(ud)
    scale = 2.0 * 10.0 / n
    L.check(lib.semb_gp_direction(C.byref(gv), C.byref(uv), n, h * w, scale, sums.data_ptr(), U.ldtype(dtype), U.stream()))
    torch.cuda.synchronize()
    norm = gx.flatten(1).norm(dim=1)
    u_ref = (scale * (norm - 1) / norm).view(n, 1, 1, 1) * gx
    assert U.rel_err(ud, u_ref) < tol
    assert abs(float(sums[0]) - float(((norm - 1) ** 2).sum())) < 1e-4 * n and abs(float(sums[1]) - float(norm.sum())) < 1e-4 * n


def _setup(h, w, n, n_z, seed=0):
    tr = OW.WganGpTrainer(h, w, n_z=n_z, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    real = torch.rand(n, h, w, 1, generator=g) * 2 - 1
    zs = [torch.randn(n, n_z, generator=g) for _ in range(4)]
    alphas = [torch.randn(n, 1, 1, 1, generator=g) for _ in range(3)]
    masks = [{k: OW.draw_masks(n, h, w, g) for k in ("fake", "real", "hat")} for _ in range(3)] + [{"gen": OW.draw_masks(n, h, w, g)}]
    return tr, real, zs, alphas, masks


def _model(tr, h, w, n, n_z, dtype):
    from sem_b200 import WganGpModel
    m = WganGpModel((h, w, 1), batch_size=n, latent_dim=n_z, dtype=dtype)
    m.discriminator.set_named({k: v.detach().numpy() for k, v in tr.d.items()})
    gp = {k: v.detach().numpy() for k, v in tr.g.items()}
    gp.update({k: v.numpy() for k, v in tr.g_state.items()})
    m.generator.set_named(gp)
    return m


def test_generator_and_critic_forward_f32():
    h, w, n, n_z = 32, 32, 4, 16
    tr, real, zs, alphas, masks = _setup(h, w, n, n_z)
    m = _model(tr, h, w, n, n_z, "f32")
    state = {k: v.clone() for k, v in tr.g_state.items()}
    with torch.no_grad():
        img_ref = OW.generator_forward(zs[0], tr.g, state, True, h, w)
        logit_ref = OW.critic_forward(real, tr.d, masks[0]["real"])[:, 0]
        logit_inf = OW.critic_forward(real, tr.d, None)[:, 0]
    img = m(zs[0].numpy(), training=True)
    assert U.rel_err(torch.from_numpy(img), img_ref) < 1e-3
    # BatchNormalization moving statistics after one training-mode call
    for k in ("bn0/moving_mean", "bn0/moving_variance", "bn3/moving_mean", "bn3/moving_variance"):
        assert U.rel_err(torch.from_numpy(m.generator.root.get_param(k)), state[k]) < 1e-4, k
    m.real_dev.copy_(real)
    m._cast_in(m.real_dev, m.D_real.in_buf, 1, n * h * w)
    m._critic_forward(m.D_real, masks[0]["real"])
    torch.cuda.synchronize()
    assert U.rel_err(m._logits(m.D_real).cpu(), logit_ref) < 1e-3
    m.D_real.e.zero_step(False)
    m.D_real.e.forward(False)                      # inference: Dropout is the identity
    torch.cuda.synchronize()
    assert U.rel_err(m._logits(m.D_real).cpu(), logit_inf) < 1e-3


def test_train_step_matches_oracle_f32():
    """One WGAN_GP.train_step_torch (three critic updates with gradient penalty, one generator update): the five metrics,
    the critic gradient of the last critic update (loss + 10 x penalty, i.e. INCLUDING the double-backward term), the
    generator gradient and the weights after the Adam steps."""
    h, w, n, n_z = 32, 32, 4, 16
    tr, real, zs, alphas, masks = _setup(h, w, n, n_z)
    m = _model(tr, h, w, n, n_z, "f32")
    ref = tr.train_step(real, zs, alphas, masks)
    m.inject = {"z": [z.numpy() for z in zs], "alpha": [a.numpy() for a in alphas], "masks": masks}
    logs = m.train_step(real.numpy())
    for k, v in ref.items():
        assert abs(logs[k] - v) < 2e-3 * max(1.0, abs(v)), (k, logs[k], v)
    for net, key in ((m.discriminator, "critic"), (m.generator, "generator")):
        worst = 0.0
        for name, gref in tr.last_grads[key].items():
            got = torch.from_numpy(net.root.get_grad(name))
            worst = max(worst, U.rel_err(got, gref))
            assert U.rel_err(got, gref) < 5e-3, (key, name, U.rel_err(got, gref))
        print(key, "worst gradient error", worst)
    lr = 2e-4
    for net, params, steps in ((m.discriminator, tr.d, 3), (m.generator, tr.g, 1)):
        for name, pref in params.items():
            d = (torch.from_numpy(net.root.get_param(name)) - pref.detach()).abs()
            # an Adam step at t <= 3 is ~ lr * sign(g): elements whose gradient is rounding noise may step the other way
            assert float(d.max()) <= 2.2 * lr * steps and float(d.mean()) < 0.1 * lr * steps, (name, float(d.max()), float(d.mean()))


def test_penalty_gradient_is_the_double_backward_term():
    """With d_cost switched off (gp_weight huge relative to it is not enough): compare the `lin` tower's weight gradient alone
    with autograd's d(gp_weight * gp)/dW on the same x_hat and masks."""
    h, w, n, n_z = 32, 32, 4, 16
    tr, real, zs, alphas, masks = _setup(h, w, n, n_z, seed=5)
    m = _model(tr, h, w, n, n_z, "f32")
    fake = (torch.rand(n, h, w, 1, generator=torch.Generator().manual_seed(9)) * 2 - 1)
    for v in tr.d.values():
        v.grad = None
    gp, norm, interp = OW.gradient_penalty(tr.d, real, fake, alphas[0], masks[0]["hat"])
    (10.0 * gp).backward()
    m.hat_dev.copy_(interp.detach())
    m._cast_in(m.hat_dev, m.D_hat.in_buf, 1, n * h * w)
    m._critic_forward(m.D_hat, masks[0]["hat"])
    m.discriminator.root.zero_grads()
    m._seed(m.D_hat, 1.0)
    m.D_hat.e.backward()
    m.sums.zero_()
    e = m.D_hat.e
    L.check(m.lib.semb_gp_direction(C.byref(m.D_hat.in_buf.view().g), C.byref(m.u0.view().t), n, h * w, 2.0 * 10.0 / n,
                                    m.sums.data_ptr(), e.dtype, e.stream))
    m.D_lin.e.zero_step(False)
    m.D_lin.e.forward(True)
    m._seed(m.D_lin, 1.0)
    m.D_lin.e.backward()
    torch.cuda.synchronize()
    assert abs(float(m.sums[0]) / n - float(gp)) < 1e-4 * max(1.0, float(gp))
    assert abs(float(m.sums[1]) / n - float(norm.mean())) < 1e-4
    for name, v in tr.d.items():
        got = torch.from_numpy(m.discriminator.root.get_grad(name))
        if name.endswith("bias"):
            assert float(got.abs().max()) == 0.0            # the penalty does not depend on the biases (piecewise linear critic)
        else:
            assert U.rel_err(got, v.grad) < 2e-3, (name, U.rel_err(got, v.grad))


def test_train_steps_bf16_close_to_oracle():
    h, w, n, n_z = 32, 32, 8, 16
    tr, real, zs, alphas, masks = _setup(h, w, n, n_z, seed=2)
    m = _model(tr, h, w, n, n_z, "bf16")
    ref = tr.train_step(real, zs, alphas, masks)
    m.inject = {"z": [z.numpy() for z in zs], "alpha": [a.numpy() for a in alphas], "masks": masks}
    logs = m.train_step(real.numpy())
    for k, v in ref.items():
        assert np.isfinite(logs[k]) and abs(logs[k] - v) < 0.1 * max(1.0, abs(v)), (k, logs[k], v)
    m.inject = None
    logs2 = m.train_step(real.numpy())                      # device-drawn randomness
    assert all(np.isfinite(v) for v in logs2.values()), logs2


def test_wgan_facade_trains_saves_and_logs(tmp_path):
    from PIL import Image
    from sem_b200 import WassersteinGAN, keras_io
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "Input_Masks"))
    for d in ("1_WGAN/Output_Images", "1_WGAN/Models", "2_CycleGAN/data/trainB"):
        os.makedirs(os.path.join(root, d))
    yy, xx = np.mgrid[0:28, 0:30]
    for i in range(3):
        disk = ((yy - 14) ** 2 + (xx - 15) ** 2 < (6 + 2 * i) ** 2).astype(np.uint8) * 255
        Image.fromarray(disk).save(os.path.join(root, "Input_Masks", f"p{i}.tif"))
    wg = WassersteinGAN.WGAN(root_dir=root, dtype="bf16")
    assert wg.train_images.shape == (12, 32, 32, 1) and set(np.unique(wg.train_images)) <= {-1.0, 0.0, 1.0}
    wg.batch_size, wg.epochs, wg.n_z = 8, 2, 32
    model = wg.start_training()
    out = os.path.join(wg.model_dir, wg.prefix)
    rows = open(os.path.join(out, "training_log.csv")).read().strip().splitlines()
    assert rows[0].split(",") == ["epoch", "d_loss", "d_total_loss", "g_loss", "grad_penalty", "grad_norm"] and len(rows) == 3
    assert all(np.isfinite(float(v)) for v in rows[-1].split(","))
    cfg, named = keras_io.load_keras(os.path.join(out, "model.keras"), rename=lambda s: s)
    assert cfg["class_name"] == "WGAN_GP" and named["discriminator/c0/kernel"].shape == (5, 5, 1, 64)
    assert named["generator/dense/kernel"].shape == (32, 4 * 4 * 256)
    assert os.path.exists(os.path.join(wg.output_dir, wg.prefix, "Epoch_00000.png"))
    img = model(np.zeros((2, 32), dtype=np.float32))
    assert img.shape == (2, 32, 32, 1) and np.abs(img).max() <= 1.0
    # workflow step 2 on the model just trained, re-loaded from its archive (the generator is untrained: most samples vanish in the
    # morphological clean-up, so only the mechanics are checked here; the placement logic is tests/test_host_logic.py)
    wg.model = None
    wg.simulate_masks(no_of_images=1, img_width=64, img_height=64)
    out_mask = np.array(Image.open(os.path.join(wg.generate_dir, "00000.tif")))
    assert out_mask.shape == (64, 64) and set(np.unique(out_mask)) <= {0, 255}
    assert wg.model is not None and wg.model.latent_dim == 32
