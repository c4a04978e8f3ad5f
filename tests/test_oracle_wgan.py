"""CPU checks of oracle/wgan.py (WassersteinGAN.py restated) and of the identity the CUDA path's gradient penalty rests on."""
import json
import os

import numpy as np
import torch

import sem_b200  # noqa: F401
from sem_b200.engine import Engine
from sem_b200.wgan_nets import WganCriticBuilder, WganGeneratorBuilder
from oracle import layers as OL, wgan as OW


def _setup(h=32, w=32, n=4, n_z=16, seed=0):
    tr = OW.WganGpTrainer(h, w, n_z=n_z, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    real = torch.rand(n, h, w, 1, generator=g) * 2 - 1
    zs = [torch.randn(n, n_z, generator=g) for _ in range(4)]
    alphas = [torch.randn(n, 1, 1, 1, generator=g) for _ in range(3)]
    masks = [{k: OW.draw_masks(n, h, w, g) for k in ("fake", "real", "hat")} for _ in range(3)] + [{"gen": OW.draw_masks(n, h, w, g)}]
    return tr, real, zs, alphas, masks


def test_shapes_and_inference_mode():
    tr, real, zs, alphas, masks = _setup()
    img = OW.generator_forward(zs[0], tr.g, dict(tr.g_state), False, 32, 32)
    assert tuple(img.shape) == (4, 32, 32, 1) and float(img.abs().max()) <= 1.0
    a = OW.critic_forward(real, tr.d, None)
    assert tuple(a.shape) == (4, 1)
    # Dropout is the identity in inference mode and changes the logits in training mode
    assert float((OW.critic_forward(real, tr.d, masks[0]["real"]) - a).abs().max()) > 0
    # parameter counts of the reference's models at 64x64 (WassersteinGAN.py:569-684): critic 4 conv blocks + Dense(1)
    n_d = sum(int(np.prod(s)) for _, s, _ in OW.critic_spec(64, 64))
    assert n_d == (25 * 1 * 64 + 64) + (25 * 64 * 128 + 128) + (25 * 128 * 256 + 256) + (25 * 256 * 512 + 512) + (4 * 4 * 512 + 1)


def test_train_step_moves_both_networks_and_updates_bn_statistics():
    tr, real, zs, alphas, masks = _setup()
    d0 = {k: v.detach().clone() for k, v in tr.d.items()}
    g0 = {k: v.detach().clone() for k, v in tr.g.items()}
    logs = tr.train_step(real, zs, alphas, masks)
    assert set(logs) == {"d_loss", "d_total_loss", "g_loss", "grad_penalty", "grad_norm"} and all(np.isfinite(v) for v in logs.values())
    assert abs(logs["d_total_loss"] - (logs["d_loss"] + 10.0 * logs["grad_penalty"])) < 1e-4 * max(1.0, abs(logs["d_total_loss"]))
    assert max(float((tr.d[k] - d0[k]).abs().max()) for k in d0) > 1e-5
    assert max(float((tr.g[k] - g0[k]).abs().max()) for k in g0) > 1e-5
    # four training-mode generator calls per step (three critic updates + the generator update), momentum 0.99
    assert float((tr.g_state["bn3/moving_variance"] - 1.0).abs().max()) > 0
    assert float(tr.last_grads["critic"]["dense/bias"].abs().max()) == 0.0      # mean(fake) - mean(real): the bias cancels


def test_penalty_gradient_equals_first_order_backprop_through_the_linearised_critic():
    """d(gp_weight * gp)/dW from torch's double backward (create_graph=True) == plain backprop of
    S(W) = <w_dense, T_W(u)> where T_W is the critic with its activations replaced by the fixed masks of the pass at x_hat
    (no biases) and u = d(gp_weight * gp)/d(grad_x D) -- what engine towers `hat` + `lin` compute on the GPU."""
    tr, real, zs, alphas, masks = _setup(seed=3)
    p, n = tr.d, real.shape[0]
    fake = torch.rand(real.shape, generator=torch.Generator().manual_seed(4)) * 2 - 1
    for v in p.values():
        v.grad = None
    gp, norm, interp = OW.gradient_penalty(p, real, fake, alphas[0], masks[0]["hat"])
    (10.0 * gp).backward()
    ref = {k: v.grad.clone() for k, v in p.items() if v.grad is not None}
    m = masks[0]["hat"]
    with torch.no_grad():
        a, M = interp.detach(), []
        for i in range(4):
            z = OL.conv2d(a, p[f"c{i}/kernel"], p[f"c{i}/bias"], 2, "same")
            mask = torch.where(z > 0, torch.ones_like(z), torch.full_like(z, 0.2))
            if i in (1, 2):
                mask = mask * m[i]
            if i == 3:
                mask = mask * m["flat"]
            M.append(mask)
            a = z * mask
    x = interp.detach().clone().requires_grad_(True)
    out = OW.critic_forward(x, {k: v.detach() for k, v in p.items()}, m)
    G = torch.autograd.grad(out, x, torch.ones_like(out))[0]
    nrm = G.flatten(1).norm(dim=1)
    u = (10.0 * 2.0 / n * (nrm - 1.0) / nrm).view(n, 1, 1, 1) * G
    W = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    t = u
    for i in range(4):
        t = OL.conv2d(t, W[f"c{i}/kernel"], None, 2, "same") * M[i]
    (t.reshape(n, -1) @ W["dense/kernel"]).sum().backward()
    for k, r in ref.items():
        if k.endswith("kernel"):
            assert float((W[k].grad - r).abs().max()) < 1e-5 * float(r.abs().max()), k
        else:
            assert float(r.abs().max()) == 0.0, k            # the penalty does not depend on the biases


def test_engine_builders_create_the_oracle_variables_in_keras_order():
    """Host logic only (dry engines): parameter names / logical shapes / order of the CUDA path's WGAN builders equal the oracle
    specs (= keras `model.weights` order: kernel, bias per layer; gamma, beta, moving_mean, moving_variance per BatchNormalization)."""
    e = Engine(4, "f32", dry=True)
    c = WganCriticBuilder(e, 64, 64)
    e.finalize()
    spec = OW.critic_spec(64, 64)
    assert c.creation_names == [n for n, _, _ in spec]
    assert [tuple(e.specs[n].logical_shape) for n in c.creation_names] == [tuple(s) for _, s, _ in spec]
    e2 = Engine(4, "f32", share=e, dry=True)
    lin = WganCriticBuilder(e2, 64, 64, like=c)                       # the linearised tower: same variables, shared masks, no biases
    e2.finalize()
    assert lin.creation_names == c.creation_names and all(a.like is b for a, b in zip(lin.masks, c.masks))
    from sem_b200.engine import ConvOp
    assert all(op.bias is None for op in e2.ops if isinstance(op, ConvOp)) and any(op.bias for op in e.ops if isinstance(op, ConvOp))
    e3 = Engine(4, "f32", dry=True)
    g = WganGeneratorBuilder(e3, 64, 64, 128)
    e3.finalize()
    want = []
    for name, shape, kind in OW.generator_spec(64, 64, 128):
        want.append((name, tuple(shape)))
        if kind == "beta":
            base = name.rsplit("/", 1)[0]
            want += [(base + "/moving_mean", tuple(shape)), (base + "/moving_variance", tuple(shape))]
    assert [(n, tuple(e3.specs[n].logical_shape)) for n in g.creation_names] == want
    assert g.out_hw == (64, 64) and c.feat == 4 * 4 * 512


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wgan_step_known_answers.json")


def _known_answers():
    tr, real, zs, alphas, masks = _setup(seed=0)
    logs = tr.train_step(real, zs, alphas, masks)
    out = {k: float(v) for k, v in logs.items()}
    out["critic_grad_l2"] = {k: float(v.norm()) for k, v in tr.last_grads["critic"].items()}
    out["generator_grad_l2"] = {k: float(v.norm()) for k, v in tr.last_grads["generator"].items()}
    out["fake_mean_abs"] = float(tr.last_fake.abs().mean())
    return out


def test_seeded_train_step_reproduces_the_committed_known_answers():
    """Pins oracle/wgan.py against silent changes: metrics and per-tensor gradient norms of one seeded train step (32x32, batch 4,
    latent 16; generated by `PYTHONPATH=. python tests/test_oracle_wgan.py --regen` with this very oracle -- a regression pin, not a reference
    vector: the reference cannot run here)."""
    want = json.load(open(GOLDEN))
    got = _known_answers()

    def close(a, b):
        return abs(a - b) <= 2e-4 * max(1e-3, abs(b))

    for k, v in want.items():
        if isinstance(v, dict):
            assert set(v) == set(got[k]) and all(close(got[k][n], v[n]) for n in v), k
        else:
            assert close(got[k], v), (k, got[k], v)


if __name__ == "__main__":
    import sys
    if "--regen" in sys.argv:
        with open(GOLDEN, "w") as fh:
            json.dump(_known_answers(), fh, indent=1, sort_keys=True)
        print("wrote", GOLDEN)
