"""WGAN_GP on libsemb200: critic + generator, one train step.

Mirrors /root/reference/Releases/Version 1.2.0/WassersteinGAN.py: WGAN_GP :27-69, gradient_penalty :88-121,
train_step_torch :181-238 (discriminator_extra_steps = 3 critic updates, then one generator update), losses :690-698,
optimizers :703-704.

The gradient penalty needs d/dW of ||grad_x D(x_hat)||: a double backward through the critic.  The critic is piecewise
linear in x (Conv2D + LeakyReLU + inverted Dropout + Dense), so with the masks of the forward pass at x_hat held fixed

    <grad_x D(x_hat), u>  =  D_lin(u)        (D_lin: same weights, no biases, activations replaced by those masks)

and therefore d(penalty)/dW = backprop of D_lin at input u = d(penalty)/d(grad_x D)  -- an ordinary first-order backward of
a fourth critic tower (`lin`) that shares the `hat` tower's masks.  Per critic update the towers run:

    gen (forward) -> fake (fwd+bwd), real (fwd+bwd), hat (fwd, data-gradient-only bwd) -> semb_gp_direction -> lin (fwd+bwd) -> Adam

All arithmetic of the path runs in libsemb200 kernels; torch draws the random numbers (latent vectors, interpolation
factors, Dropout keep masks), forms x_hat = real + alpha (fake - real) on the fp32 staging copies and seeds the constant
output gradients (+-1/N)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from .engine import Engine
from .wgan_nets import WganCriticBuilder, WganGeneratorBuilder

METRICS = ["d_loss", "d_total_loss", "g_loss", "grad_penalty", "grad_norm"]


class _WNet:
    """One network (flat parameters, Adam state) plus its towers (engines sharing those parameters)."""

    def __init__(self, kind: str, h: int, w: int, n_z: int, dtype: str, use_tc: bool, seed: int):
        self.kind, self.h, self.w, self.n_z, self.dtype, self.use_tc, self.seed = kind, h, w, n_z, dtype, use_tc, seed
        self.root: Optional[Engine] = None
        self.towers: Dict[str, tuple] = {}

    def tower(self, name: str, n: int, **kw):
        eng = Engine(n, self.dtype, use_tc=self.use_tc, share=self.root)
        if self.kind == "gen":
            b = WganGeneratorBuilder(eng, self.h, self.w, self.n_z)
        else:
            b = WganCriticBuilder(eng, self.h, self.w, **kw)
        eng.finalize()
        if self.root is None:
            self.root = eng
            self.names = list(b.creation_names)
            eng.init_params(self.seed)          # GlorotUniform kernels, gamma / moving_variance ones, the rest zeros
        self.towers[name] = (eng, b)
        return eng, b

    def get_weights(self) -> List[np.ndarray]:
        return [self.root.get_param(n) for n in self.names]

    def set_weights(self, ws):
        if len(ws) != len(self.names):
            raise ValueError(f"expected {len(self.names)} weight arrays, got {len(ws)}")
        for n, w in zip(self.names, ws):
            self.root.set_param(n, np.asarray(w))

    def set_named(self, d):
        for n in self.names:
            self.root.set_param(n, np.asarray(d[n]))


class WganGpModel:
    """`WGAN_GP(discriminator, generator, latent_dim, discriminator_extra_steps=3, gp_weight=10.0)` with both networks built in."""

    def __init__(self, image_shape=(64, 64, 1), batch_size: int = 64, latent_dim: int = 128, discriminator_extra_steps: int = 3,
                 gp_weight: float = 10.0, dtype: str = "bf16", use_tc: bool = True, seed: int = 0):
        h, w = image_shape[0], image_shape[1]
        self.h, self.w, self.n, self.latent_dim = h, w, batch_size, latent_dim
        self.d_steps, self.gp_weight, self.dtype = discriminator_extra_steps, gp_weight, dtype
        n = batch_size
        self.generator = _WNet("gen", h, w, latent_dim, dtype, use_tc, seed)
        self.discriminator = _WNet("critic", h, w, latent_dim, dtype, use_tc, seed + 1)
        _, self.G = self.generator.tower("z", n)
        _, self.D_real = self.discriminator.tower("real", n)
        _, self.D_fake = self.discriminator.tower("fake", n, in_buf=self.G.out_buf)
        _, self.D_hat = self.discriminator.tower("hat", n, input_requires_grad=True)
        self.u0 = self.D_hat.e.new_buf(h, w, 8, "gp_direction", requires_grad=False)
        _, self.D_lin = self.discriminator.tower("lin", n, in_buf=self.u0, like=self.D_hat)
        self.D_hat.e.skip_wgrad = True          # this tower only yields grad_x D(x_hat); its weight gradients are not part of any loss
        self.lib, self.dev = self.G.e.lib, self.G.e.device
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.real_dev = torch.zeros((n, h, w, 1), **f32)
        self.fake_dev = torch.zeros((n, h, w, 1), **f32)
        self.hat_dev = torch.zeros((n, h, w, 1), **f32)
        self.z_dev = torch.zeros((n, latent_dim), **f32)
        self.sums = torch.zeros(4, **f32)
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = 2e-4, 0.5, 0.9, 1e-7
        # injected randomness (parity tests): lists consumed in order; None = draw on the device
        self.inject: Optional[dict] = None
        self.last_logits: Dict[str, torch.Tensor] = {}

    # ---- Keras-like surface -----------------------------------------------------------------------------------------------
    def compile(self, learning_rate: float = 2e-4, beta_1: float = 0.5, beta_2: float = 0.9, **_ignored):
        self.learning_rate, self.beta_1, self.beta_2 = learning_rate, beta_1, beta_2

    @property
    def nets(self):
        return {"generator": self.generator, "discriminator": self.discriminator}

    def __call__(self, z, training: bool = False) -> np.ndarray:
        """generator(z): (m <= batch, latent_dim) -> (m, H, W, 1) in [-1, 1] (WGAN_GP.call :68-69)."""
        z = np.asarray(z, dtype=np.float32)
        m = z.shape[0]
        self.z_dev.zero_()
        self.z_dev[:m].copy_(torch.from_numpy(np.ascontiguousarray(z)))
        self._gen_forward(training)
        self._cast_out(self.G.out_buf, self.fake_dev)
        return self.fake_dev[:m].cpu().numpy()

    # ---- plumbing -----------------------------------------------------------------------------------------------------------
    def _cast_in(self, src: torch.Tensor, buf, channels: int, npix: int):
        e = self.G.e
        L.check(self.lib.semb_cast_in(src.data_ptr(), channels, C.byref(buf.view().t), npix, e.dtype, e.stream))

    def _cast_out(self, buf, dst: torch.Tensor):
        e = self.G.e
        L.check(self.lib.semb_cast_out(C.byref(buf.view().t), dst.data_ptr(), 1, self.n * self.h * self.w, e.dtype, e.stream))

    def _gen_forward(self, training: bool):
        self._cast_in(self.z_dev, self.G.in_buf, self.latent_dim, self.n)
        self.G.e.zero_step(False)
        self.G.e.forward(training)

    def _critic_forward(self, b, masks=None):
        """One critic tower forward with training=True (fresh Dropout masks, or the injected ones)."""
        for op, key in zip((b.masks[1], b.masks[2], b.masks[4]), (1, 2, "flat")):
            op.frozen = masks is not None
            if masks is not None:
                op.drop.data.copy_(torch.as_tensor(masks[key]).to(op.drop.data.dtype).reshape(op.drop.data.shape))
        b.e.zero_step(False)
        b.e.forward(True)

    def _seed(self, b, value: float):
        g = b.out_buf.grad_tensor()
        g.zero_()
        g[..., 0] = value               # channel 0 is the logit; the padded lanes carry no gradient

    def _logits(self, b) -> torch.Tensor:
        return b.out_buf.data[0, 0, :, 0].float()

    def _take(self, key: str, i: int):
        if self.inject is None or key not in self.inject:
            return None
        return self.inject[key][i]

    def _adam(self, net: _WNet):
        e = net.root
        e.lr.fill_(self.learning_rate)
        e.adam(self.beta_1, self.beta_2, self.epsilon, 1.0)

    # ---- the step -----------------------------------------------------------------------------------------------------------
    def train_step(self, real_images) -> Dict[str, float]:
        if isinstance(real_images, tuple):
            real_images = real_images[0]
        x = torch.as_tensor(np.ascontiguousarray(np.asarray(real_images, dtype=np.float32)))
        if tuple(x.shape) != (self.n, self.h, self.w, 1):
            raise ValueError(f"expected a batch of shape {(self.n, self.h, self.w, 1)}, got {tuple(x.shape)}")
        self.real_dev.copy_(x, non_blocking=True)
        n, npix = self.n, self.n * self.h * self.w
        self._cast_in(self.real_dev, self.D_real.in_buf, 1, npix)
        d_cost = gp = gn = None
        for i in range(self.d_steps):
            # ---- critic update i (train_step_torch :187-214)
            z = self._take("z", i)
            self.z_dev.copy_(torch.as_tensor(z)) if z is not None else self.z_dev.normal_()
            self._gen_forward(True)
            m = self._take("masks", i) or {}
            self._critic_forward(self.D_fake, m.get("fake"))
            self._critic_forward(self.D_real, m.get("real"))
            # interpolated = real + alpha * (fake - real), alpha ~ N(0, 1) per sample (:97-99: a NORMAL deviate, kept)
            self._cast_out(self.G.out_buf, self.fake_dev)
            a = self._take("alpha", i)
            alpha = torch.as_tensor(a).to(self.dev).reshape(n, 1, 1, 1) if a is not None else torch.randn((n, 1, 1, 1), device=self.dev)
            torch.add(self.real_dev, alpha * (self.fake_dev - self.real_dev), out=self.hat_dev)
            self._cast_in(self.hat_dev, self.D_hat.in_buf, 1, npix)
            self._critic_forward(self.D_hat, m.get("hat"))
            self.discriminator.root.zero_grads()
            # grad_x D(x_hat): backward of the `hat` tower with d(out) = 1 (grad_outputs = ones, :114), data gradients only
            self._seed(self.D_hat, 1.0)
            self.D_hat.e.backward()
            self.sums.zero_()
            e = self.D_hat.e
            L.check(self.lib.semb_gp_direction(C.byref(self.D_hat.in_buf.view().g), C.byref(self.u0.view().t), n, self.h * self.w,
                                               2.0 * self.gp_weight / n, self.sums.data_ptr(), e.dtype, e.stream))
            # d(gp_weight * gp)/dW: first-order backward of the linearised critic at input u0
            self.D_lin.e.zero_step(False)
            self.D_lin.e.forward(True)
            self._seed(self.D_lin, 1.0)
            self.D_lin.e.backward()
            # d_cost = mean(fake_logits) - mean(real_logits) (:690-694)
            self._seed(self.D_fake, 1.0 / n)
            self._seed(self.D_real, -1.0 / n)
            self.D_fake.e.backward()
            self.D_real.e.backward()
            d_cost = self._logits(self.D_fake).mean() - self._logits(self.D_real).mean()
            gp, gn = self.sums[0] / n, self.sums[1] / n
            self._adam(self.discriminator)
        # ---- generator update (:216-232)
        z = self._take("z", self.d_steps)
        self.z_dev.copy_(torch.as_tensor(z)) if z is not None else self.z_dev.normal_()
        self._gen_forward(True)
        m = self._take("masks", self.d_steps) or {}
        self._critic_forward(self.D_fake, m.get("gen"))
        g_loss = -self._logits(self.D_fake).mean()
        self._seed(self.D_fake, -1.0 / n)
        self.D_fake.e.skip_wgrad = True           # generator_loss only needs the data gradient through the critic
        self.D_fake.e.backward()
        self.D_fake.e.skip_wgrad = False
        self.generator.root.zero_grads()
        self.G.e.backward()
        self._adam(self.generator)
        out = torch.stack([d_cost, d_cost + self.gp_weight * gp, g_loss, gp, gn]).cpu()
        return {k: float(v) for k, v in zip(METRICS, out)}
