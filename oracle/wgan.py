"""WGAN-GP critic / generator / train step restated in plain torch (CPU).  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows /root/reference/Releases/Version 1.2.0/WassersteinGAN.py:
  conv_block :547-567, get_discriminator_model :569-621, upsample_block :624-649, get_generator_model :651-684,
  discriminator_loss / generator_loss :690-698, optimizers :703-704 (Adam 2e-4, beta_1 0.5, beta_2 0.9),
  WGAN_GP.gradient_penalty :88-121 (torch branch: autograd.grad(..., create_graph=True)), train_step_torch :181-238
  (discriminator_extra_steps = 3, gp_weight = 10).

Randomness is INJECTED so that the CUDA path can be fed the same deviates: latent vectors, the interpolation factors
(keras.random.normal -- the reference draws them from a NORMAL distribution, :97, kept) and the Dropout keep masks
(already scaled by 1/(1-rate), Keras' inverted dropout).  Reference quirk kept: d_loss.backward() also back-propagates into
the generator through fake_images; those gradients are zeroed before the generator step (:221) and never used.
Parity status: unpinned by the reference (no tests, Keras not installable); layer semantics per SURVEY.md Appendix B.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import layers as L

CRITIC_FILTERS = (64, 128, 256, 512)
DROP = {1: 0.3, 2: 0.3, "flat": 0.2}          # Dropout after conv blocks 1 and 2 (:586-606) and before the Dense (:618)


def critic_spec(h: int, w: int, channels: int = 1):
    e, cin = [], channels
    for i, f in enumerate(CRITIC_FILTERS):
        e += [(f"c{i}/kernel", (5, 5, cin, f), "conv"), (f"c{i}/bias", (f,), "bias")]
        cin = f
    e += [("dense/kernel", ((h // 16) * (w // 16) * cin, 1), "dense"), ("dense/bias", (1,), "bias")]
    return e


def generator_spec(h: int, w: int, n_z: int = 128):
    feat = (h // 8) * (w // 8) * 256
    e = [("dense/kernel", (n_z, feat), "dense"), ("bn0/gamma", (feat,), "gamma"), ("bn0/beta", (feat,), "beta")]
    cin = 256
    for i, f in enumerate((128, 64, 1)):
        e += [(f"up{i}/kernel", (3, 3, cin, f), "conv"), (f"bn{i + 1}/gamma", (f,), "gamma"), (f"bn{i + 1}/beta", (f,), "beta")]
        cin = f
    return e


def generator_state(h: int, w: int):
    feat = (h // 8) * (w // 8) * 256
    s = OrderedDict()
    for i, c in enumerate((feat, 128, 64, 1)):
        s[f"bn{i}/moving_mean"] = torch.zeros(c)
        s[f"bn{i}/moving_variance"] = torch.ones(c)
    return s


def init_params(spec, gen: torch.Generator):
    p = OrderedDict()
    for name, shape, kind in spec:
        if kind == "conv":
            kh, kw, ci, co = shape
            t = L.glorot_uniform(shape, gen, kh * kw * ci, kh * kw * co)
        elif kind == "dense":
            t = L.glorot_uniform(shape, gen, shape[0], shape[1])
        elif kind == "gamma":
            t = torch.ones(shape)
        else:
            t = torch.zeros(shape)
        p[name] = t
    return p


def mask_shapes(n: int, h: int, w: int):
    """Shapes of the three Dropout masks of one critic call."""
    return {1: (n, h // 4, w // 4, 128), 2: (n, h // 8, w // 8, 256), "flat": (n, h // 16, w // 16, 512)}


def draw_masks(n: int, h: int, w: int, gen: torch.Generator):
    return {k: (torch.rand(s, generator=gen) >= DROP[k]).float() / (1.0 - DROP[k]) for k, s in mask_shapes(n, h, w).items()}


def critic_forward(x, p, masks=None):
    """x (N,H,W,1) -> (N,1).  masks: {1, 2, "flat"} keep masks (training=True) or None (inference: Dropout is the identity)."""
    for i in range(4):
        x = L.leaky_relu(L.conv2d(x, p[f"c{i}/kernel"], p[f"c{i}/bias"], 2, "same"), 0.2)
        if masks is not None and i in (1, 2):
            x = x * masks[i]
    if masks is not None:
        x = x * masks["flat"]
    flat = x.reshape(x.shape[0], -1)                      # keras Flatten of NHWC: (h, w, c) order
    return flat @ p["dense/kernel"] + p["dense/bias"]


def generator_forward(z, p, state, training: bool, h: int, w: int):
    """z (N, n_z) -> (N,H,W,1) in [-1,1]; BatchNormalization with its moving statistics updated in `state` when training."""
    def bn(x, i):
        y, m, v = L.batch_norm(x, p[f"bn{i}/gamma"], p[f"bn{i}/beta"], state[f"bn{i}/moving_mean"], state[f"bn{i}/moving_variance"], training)
        if training:
            state[f"bn{i}/moving_mean"], state[f"bn{i}/moving_variance"] = m, v
        return y

    x = (z @ p["dense/kernel"]).reshape(z.shape[0], 1, 1, -1)
    x = L.leaky_relu(bn(x, 0), 0.2).reshape(z.shape[0], h // 8, w // 8, 256)
    for i in range(3):
        x = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)          # UpSampling2D((2,2)), nearest
        x = bn(L.conv2d(x, p[f"up{i}/kernel"], None, 1, "same"), i + 1)
        x = torch.tanh(x) if i == 2 else L.leaky_relu(x, 0.2)
    return x


def gradient_penalty(p, real, fake, alpha, masks):
    """WGAN_GP.gradient_penalty :88-121.  Returns (gp, norm (N,), interpolated)."""
    interp = (real + alpha * (fake - real)).detach().requires_grad_(True)
    pred = critic_forward(interp, p, masks)
    grads = torch.autograd.grad(pred, interp, torch.ones_like(pred), create_graph=True, retain_graph=True)[0]
    norm = torch.sqrt(torch.sum(grads * grads, dim=(1, 2, 3)))
    return torch.mean((norm - 1.0) ** 2), norm, interp


class WganGpTrainer:
    def __init__(self, h: int, w: int, n_z: int = 128, d_steps: int = 3, gp_weight: float = 10.0, seed: int = 0):
        gen = torch.Generator().manual_seed(seed)
        self.h, self.w, self.n_z, self.d_steps, self.gp_weight = h, w, n_z, d_steps, gp_weight
        self.d = init_params(critic_spec(h, w), gen)
        self.g = init_params(generator_spec(h, w, n_z), gen)
        self.g_state = generator_state(h, w)
        for v in list(self.d.values()) + list(self.g.values()):
            v.requires_grad_(True)
        self.d_opt = L.KerasAdam(list(self.d.values()), lr=2e-4, beta_1=0.5, beta_2=0.9)
        self.g_opt = L.KerasAdam(list(self.g.values()), lr=2e-4, beta_1=0.5, beta_2=0.9)
        self.last_grads = {}

    def train_step(self, real, zs, alphas, masks):
        """real (N,H,W,1); zs: d_steps+1 latent tensors (N,n_z); alphas: d_steps tensors (N,1,1,1);
        masks: d_steps dicts {"fake", "real", "hat"} of critic mask dicts, then one {"gen"} dict for the generator step."""
        logs = {}
        for i in range(self.d_steps):
            fake = generator_forward(zs[i], self.g, self.g_state, True, self.h, self.w)
            fake_logits = critic_forward(fake, self.d, masks[i]["fake"])
            real_logits = critic_forward(real, self.d, masks[i]["real"])
            d_cost = fake_logits.mean() - real_logits.mean()
            gp, norm, _ = gradient_penalty(self.d, real, fake.detach(), alphas[i], masks[i]["hat"])
            d_loss = d_cost + gp * self.gp_weight
            for v in self.d.values():
                v.grad = None
            d_loss.backward()
            if i == self.d_steps - 1:
                self.last_grads["critic"] = OrderedDict((k, v.grad.clone()) for k, v in self.d.items())
            ps = list(self.d.values())
            self.d_opt.apply([q.grad for q in ps], ps)
            logs.update(d_loss=float(d_cost), d_total_loss=float(d_loss), grad_penalty=float(gp), grad_norm=float(norm.mean()))
        gen_img = generator_forward(zs[self.d_steps], self.g, self.g_state, True, self.h, self.w)
        g_loss = -critic_forward(gen_img, self.d, masks[self.d_steps]["gen"]).mean()
        for v in self.g.values():
            v.grad = None
        g_loss.backward()
        self.last_grads["generator"] = OrderedDict((k, v.grad.clone()) for k, v in self.g.items())
        ps = list(self.g.values())
        self.g_opt.apply([q.grad for q in ps], ps)
        logs["g_loss"] = float(g_loss)
        self.last_fake = gen_img.detach()
        return logs
