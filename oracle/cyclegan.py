"""CycleGAN ResNet generator / PatchGAN discriminator / train step restated in plain torch (CPU).
TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows /root/reference/Releases/Version 1.2.0/CycleGAN.py:
  residual_block :323-337, downsample :339-345, upsample :347-358,
  get_resnet_generator :360-423, get_discriminator :425-451,
  generator_loss_fn / discriminator_loss_fn :301-308,
  CycleGanModel.train_step_torch :615-710, ImagePool.query :927-964.

Default configuration = the one StartProcess.py drives (StartProcess.py:91-102): use_skip_connection=False,
gaussian_noise_value=0.0, use_resize_convolution=False, use_binary_crossentropy=False, lambda_identity 0.5; the
skip-connection (:396-415), resize-convolution (:348-351) and GaussianNoise (:427-447) variants are restated as options.
Parity status: unpinned by the reference (no tests, Keras not installable); layer semantics per SURVEY.md
Appendix B; cross-checked against naive loops in tests/test_oracle_layers.py.
"""
from __future__ import annotations

import random
from collections import OrderedDict

import torch

from . import layers as L


# ------------------------------------------------------------------------------------------ parameter specs
def generator_spec(filters: int = 64, n_down: int = 3, n_res: int = 9, n_up: int = 3, channels: int = 1,
                   use_skip_connection: bool = False, use_resize_convolution: bool = False):
    """[(name, shape, kind)] in creation order."""
    e = []
    f = filters
    e.append(("stem/kernel", (7, 7, channels, f), "conv"))
    e += [("stem_in/gamma", (f,), "gamma"), ("stem_in/beta", (f,), "beta")]
    for i in range(n_down):
        e.append((f"down{i}/kernel", (3, 3, f, 2 * f), "conv"))
        f *= 2
        e += [(f"down{i}_in/gamma", (f,), "gamma"), (f"down{i}_in/beta", (f,), "beta")]
    for i in range(n_res):
        for j in (0, 1):
            e.append((f"res{i}_{j}/kernel", (3, 3, f, f), "conv"))
            e += [(f"res{i}_{j}_in/gamma", (f,), "gamma"), (f"res{i}_{j}_in/beta", (f,), "beta")]
    for i in range(n_up):
        if use_resize_convolution:
            e.append((f"up{i}/kernel", (3, 3, f, f // 2), "conv"))      # UpSampling2D + ReflectionPadding2D + Conv2D (:348-351)
        else:
            e.append((f"up{i}/kernel", (3, 3, f // 2, f), "convT"))       # Keras Conv2DTranspose kernel (kh,kw,Cout,Cin)
        f //= 2
        e += [(f"up{i}_in/gamma", (f,), "gamma"), (f"up{i}_in/beta", (f,), "beta")]
    e.append(("head/kernel", (7, 7, f, channels), "conv"))
    e.append(("head/bias", (channels,), "bias"))
    if use_skip_connection:         # CycleGAN.py:396-415
        e.append(("skip_short/kernel", (1, 1, channels, f), "conv"))
        e += [("skip_short_in/gamma", (f,), "gamma"), ("skip_short_in/beta", (f,), "beta")]
        e.append(("skip_conv/kernel", (3, 3, channels, f), "conv"))
        e += [("skip_conv_in/gamma", (f,), "gamma"), ("skip_conv_in/beta", (f,), "beta")]
        e += [("skip_sum_in/gamma", (f,), "gamma"), ("skip_sum_in/beta", (f,), "beta")]
        e.append(("skip_out/kernel", (1, 1, f + channels, channels), "conv"))
    return e


def discriminator_spec(filters: int = 128, n_down: int = 2, channels: int = 1):
    e = []
    f = filters
    e += [("d0/kernel", (4, 4, channels, f), "conv"), ("d0/bias", (f,), "bias")]
    for i in range(n_down):
        e.append((f"d{i + 1}/kernel", (4, 4, f, 2 * f), "conv"))
        f *= 2
        e += [(f"d{i + 1}_in/gamma", (f,), "gamma"), (f"d{i + 1}_in/beta", (f,), "beta")]
    e += [("out/kernel", (4, 4, f, 1), "conv"), ("out/bias", (1,), "bias")]
    return e


def init_params(spec, gen: torch.Generator, dtype=torch.float32):
    """One shared GlorotUniform stream for all kernels (CycleGAN.py:125), gamma ones, beta / bias zeros."""
    p = OrderedDict()
    for name, shape, kind in spec:
        if kind == "conv":
            kh, kw, ci, co = shape
            t = L.glorot_uniform(shape, gen, kh * kw * ci, kh * kw * co)
        elif kind == "convT":
            kh, kw, co, ci = shape
            t = L.glorot_uniform(shape, gen, kh * kw * co, kh * kw * ci)
        elif kind == "gamma":
            t = torch.ones(shape)
        else:
            t = torch.zeros(shape)
        p[name] = t.to(dtype)
    return p


# ------------------------------------------------------------------------------------------ networks
def _in_relu(x, p, name, act=torch.relu):
    x = L.instance_norm(x, p[name + "/gamma"], p[name + "/beta"])
    return act(x) if act is not None else x


def generator_forward(x, p, n_down: int = 3, n_res: int = 9, n_up: int = 3, taps=None):
    """get_resnet_generator(...)(x, training=True).  x NHWC in [-1,1].  The skip-connection / resize-convolution
    variants are recognised from the parameter set (generator_spec)."""
    img_input = x
    H, W = x.shape[1], x.shape[2]
    m = 2 ** n_down
    ph, pw = (m - H % m) % m, (m - W % m) % m
    x = L.reflection_pad(x, pw, ph)
    x = L.reflection_pad(x, 6, 6)
    x = L.conv2d(x, p["stem/kernel"], None, 1, "valid")
    x = _in_relu(x, p, "stem_in")
    if taps is not None:
        taps["stem"] = x
    for i in range(n_down):
        x = L.conv2d(x, p[f"down{i}/kernel"], None, 2, "same")
        x = _in_relu(x, p, f"down{i}_in")
    if taps is not None:
        taps["down"] = x
    for i in range(n_res):
        y = L.reflection_pad(x, 2, 2)
        y = L.conv2d(y, p[f"res{i}_0/kernel"], None, 1, "valid")
        y = _in_relu(y, p, f"res{i}_0_in")
        y = L.reflection_pad(y, 2, 2)
        y = L.conv2d(y, p[f"res{i}_1/kernel"], None, 1, "valid")
        y = _in_relu(y, p, f"res{i}_1_in", act=None)
        x = x + y
    if taps is not None:
        taps["res"] = x
    for i in range(n_up):
        k = p[f"up{i}/kernel"]
        if k.shape[2] > k.shape[3]:        # Conv2D kernel (3,3,f,f/2): the use_resize_convolution branch (:348-351)
            x = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)       # UpSampling2D nearest
            x = L.reflection_pad(x, 2, 2)
            x = L.conv2d(x, k, None, 1, "valid")
        else:
            x = L.conv2d_transpose(x, k, None, 2)
        x = _in_relu(x, p, f"up{i}_in")
    x = L.reflection_pad(x, 6, 6)
    x = L.conv2d(x, p["head/kernel"], p["head/bias"], 1, "valid")
    if "skip_out/kernel" in p:          # use_skip_connection (:396-415)
        shortcut = L.conv2d(img_input, p["skip_short/kernel"], None, 1, "valid")
        shortcut = _in_relu(shortcut, p, "skip_short_in")
        out = L.reflection_pad(img_input, 2, 2)
        out = L.conv2d(out, p["skip_conv/kernel"], None, 1, "valid")
        out = _in_relu(out, p, "skip_conv_in")
        out = _in_relu(shortcut + out, p, "skip_sum_in")
        x = torch.cat([out, x], dim=3)
        x = L.conv2d(x, p["skip_out/kernel"], None, 1, "valid")
    return torch.tanh(x)


def discriminator_forward(x, p, n_down: int = 2, noise=None):
    """get_discriminator(...)(x, training=True).  `noise`: the GaussianNoise deviates added in front of each conv
    (d0, d1.., out), one tensor per conv, or None (gaussian_noise_value = 0)."""
    nz = (lambda t, i: t + noise[i]) if noise is not None else (lambda t, i: t)
    x = L.conv2d(nz(x, 0), p["d0/kernel"], p["d0/bias"], 2, "valid")
    x = L.leaky_relu(x, 0.2)
    for i in range(n_down):
        x = L.conv2d(nz(x, i + 1), p[f"d{i + 1}/kernel"], None, 2, "valid")
        x = _in_relu(x, p, f"d{i + 1}_in", act=lambda t: L.leaky_relu(t, 0.2))
    return L.conv2d(nz(x, n_down + 1), p["out/kernel"], p["out/bias"], 1, "valid")


# ------------------------------------------------------------------------------------------ image pool
class ImagePool:
    """CycleGAN.py:908-964 verbatim in behaviour: only the first `batch_size` images of a batch are considered."""

    def __init__(self, batch_size, pool_size=50, rng=random):
        self.pool_size, self.batch_size, self.rng = pool_size, batch_size, rng
        self.num_imgs, self.images = 0, []

    def query(self, images):
        if self.pool_size == 0:
            return images
        out = []
        for index in range(self.batch_size):
            if index >= images.shape[0]:
                break
            image = images[index:index + 1]
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(image)
                out.append(image)
            else:
                if self.rng.uniform(0, 1) > 0.5:
                    rid = self.rng.randint(0, self.pool_size - 1)
                    tmp = self.images[rid].clone()
                    self.images[rid] = image
                    out.append(tmp)
                else:
                    out.append(image)
        return torch.cat(out, 0)


# ------------------------------------------------------------------------------------------ train step
class CycleGanTrainer:
    def __init__(self, filters=64, lr=2e-4, lambda_cycle=10.0, lambda_identity=0.5, pool_batch=2, pool_size=50, seed=0,
                 label_smoothing=0.0, dtype=torch.float32, n_res=9):
        gen = torch.Generator().manual_seed(seed)
        self.n_res = n_res
        self.gspec = generator_spec(filters, n_res=n_res)
        self.dspec = discriminator_spec(2 * filters)
        # construction order gen_a, gen_b, disc_a, disc_b shares one initializer stream (Appendix B item 10)
        self.nets = OrderedDict()
        self.nets["gen_a"] = init_params(self.gspec, gen, dtype)
        self.nets["gen_b"] = init_params(self.gspec, gen, dtype)
        self.nets["disc_a"] = init_params(self.dspec, gen, dtype)
        self.nets["disc_b"] = init_params(self.dspec, gen, dtype)
        for net in self.nets.values():
            for v in net.values():
                v.requires_grad_(True)
        self.opt = {k: L.KerasAdam(list(v.values()), lr=lr, beta_1=0.5) for k, v in self.nets.items()}
        self.lc, self.li, self.ls = lambda_cycle, lambda_identity, label_smoothing
        rng_a, rng_b = random.Random(seed), random.Random(seed + 1)
        self.pool_a = ImagePool(pool_batch, pool_size, rng_a)
        self.pool_b = ImagePool(pool_batch, pool_size, rng_b)
        self.last_grads = {}

    def _adv(self, pred, target_one: bool):
        t = (1.0 - self.ls) + self.ls / 2 if target_one else self.ls / 2
        return L.mse(torch.full_like(pred, t), pred)

    def _zero(self, *names):
        for n in names:
            for v in self.nets[n].values():
                v.grad = None

    def G(self, which, x):
        return generator_forward(x, self.nets[which], n_res=self.n_res)

    def D(self, which, x):
        return discriminator_forward(x, self.nets[which])

    def train_step(self, real_a, real_b):
        fake_b = self.G("gen_a", real_a)
        fake_a = self.G("gen_b", real_b)
        cycled_a = self.G("gen_b", fake_b)
        cycled_b = self.G("gen_a", fake_a)
        same_a = self.G("gen_b", real_a)
        same_b = self.G("gen_a", real_b)
        disc_fake_a = self.D("disc_a", fake_a)
        disc_fake_b = self.D("disc_b", fake_b)
        adv_a = self._adv(disc_fake_b, True)
        adv_b = self._adv(disc_fake_a, True)
        cyc_a = L.mae(real_b, cycled_b) * self.lc
        cyc_b = L.mae(real_a, cycled_a) * self.lc
        id_a = L.mae(real_b, same_b) * self.lc * self.li
        id_b = L.mae(real_a, same_a) * self.lc * self.li
        total_a = adv_a + cyc_a + id_a
        total_b = adv_b + cyc_b + id_b
        self._zero("gen_a", "gen_b")
        total_a.backward(retain_graph=True)
        total_b.backward(retain_graph=True)
        self.last_grads["gen_a"] = OrderedDict((k, v.grad.clone()) for k, v in self.nets["gen_a"].items())
        self.last_grads["gen_b"] = OrderedDict((k, v.grad.clone()) for k, v in self.nets["gen_b"].items())
        for n in ("gen_a", "gen_b"):
            ps = list(self.nets[n].values())
            self.opt[n].apply([p.grad for p in ps], ps)

        d_real_a = self.D("disc_a", real_a)
        d_fake_a = self.D("disc_a", self.pool_a.query(fake_a.detach().clone()))
        d_real_b = self.D("disc_b", real_b)
        d_fake_b = self.D("disc_b", self.pool_b.query(fake_b.detach().clone()))
        la_real, la_fake = self._adv(d_real_a, True), self._adv(d_fake_a, False)
        lb_real, lb_fake = self._adv(d_real_b, True), self._adv(d_fake_b, False)
        loss_da, loss_db = (la_real + la_fake) * 0.5, (lb_real + lb_fake) * 0.5
        self._zero("disc_a", "disc_b")
        loss_da.backward()
        loss_db.backward()
        self.last_grads["disc_a"] = OrderedDict((k, v.grad.clone()) for k, v in self.nets["disc_a"].items())
        self.last_grads["disc_b"] = OrderedDict((k, v.grad.clone()) for k, v in self.nets["disc_b"].items())
        for n in ("disc_a", "disc_b"):
            ps = list(self.nets[n].values())
            self.opt[n].apply([p.grad for p in ps], ps)
        f = lambda t: float(t.detach())
        return {"d_a": f(loss_da), "d_b": f(loss_db), "d_fake_a": f(la_fake), "d_fake_b": f(lb_fake), "d_real_a": f(la_real),
                "d_real_b": f(lb_real), "g_a": f(total_a), "g_b": f(total_b), "g_adv_a": f(adv_a), "g_adv_b": f(adv_b),
                "g_cyc_a": f(cyc_a), "g_cyc_b": f(cyc_b), "g_id_a": f(id_a), "g_id_b": f(id_b)}, fake_a.detach(), fake_b.detach()
