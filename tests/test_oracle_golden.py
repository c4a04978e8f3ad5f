"""Pins the oracle: (1) against the committed golden vectors, (2) against the reference's shipped trained weights +
dataset (known-answer IoU inside the band the reference publishes, README.md:55-57), (3) against /root/reference
itself when that tree is present (build container only)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import pb_reader, unet as OU

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF_PB = "/root/reference/ImageJ Plugin/SEM_Particle_Segmentation_Models/TiO2_UNet_Masks_GAN.pb"
SHIP = os.path.join(os.path.dirname(os.path.dirname(__file__)), "baseline", "_ref")


def _weights():
    with np.load(os.path.join(GOLD, "unet_gan_weights.npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def test_spec_matches_shipped_graph_names_and_shapes():
    spec = OU.UNetSpec(16)
    w = _weights()
    assert list(w.keys()) == spec.names()
    for name, shape, _ in spec.entries:
        assert tuple(w[name].shape) == tuple(shape), name
    assert sum(int(np.prod(s)) for _, s, k in spec.entries if k not in ("bn_mean", "bn_var")) == 2414297
    assert sum(int(np.prod(s)) for _, s, _ in spec.entries) == 2429491


def test_golden_crops_regression_and_iou():
    with np.load(os.path.join(GOLD, "sem_crops.npz")) as z:
        crops, gold, masks = z["crops"].astype(np.float32), z["oracle_sigmoid"], np.unpackbits(z["masks"], axis=-1).astype(bool)
    x = np.stack([(c - c.min()) / (c - c.min()).max() for c in crops])[..., None]
    with torch.no_grad():
        y, _ = OU.unet_forward(torch.from_numpy(x), _weights(), training=False)
    y = y[..., 0].numpy()
    assert np.abs(y - gold).max() < 1e-4
    iou = [np.logical_and(a > 0.5, m).sum() / max(np.logical_or(a > 0.5, m).sum(), 1) for a, m in zip(y, masks)]
    assert np.mean(iou) > 0.75, iou      # trained weights of the reference segment its own images


def test_known_answers_are_inside_the_published_band():
    ka = json.load(open(os.path.join(GOLD, "pb_known_answers.json")))
    gan = ka["models"]["GAN"]["mean_iou"]
    assert ka["n_images"] == 40
    # README.md:55-57 publishes 0.8108 (v1.0.0), 0.8502 (torch), 0.8762 (tf) for the automatic workflow
    assert 0.80 < gan["0.5"] < 0.88 and abs(gan["0.5"] - 0.8369) < 5e-4
    assert abs(ka["models"]["Manual"]["mean_iou"]["0.5"] - 0.9295) < 5e-4
    assert abs(ka["models"]["TSEM"]["mean_iou"]["0.5"] - 0.9006) < 5e-4


def test_small_train_step_regression():
    with np.load(os.path.join(GOLD, "unet_small_step.npz")) as z:
        g = {k: z[k] for k in z.files}
    spec = OU.UNetSpec(16)
    tr = OU.UNetTrainer(spec, spec.init_params(seed=0), float(g["weighting"]))
    logs, yp = tr.train_step(torch.from_numpy(g["x"]), torch.from_numpy(g["y"]))
    assert abs(logs["loss"] - float(g["loss"])) < 1e-5 and abs(logs["acc"] - float(g["acc"])) < 1e-6
    assert np.abs(yp.numpy() - g["y_pred"]).max() < 1e-5
    for k in g:
        if k.startswith("grad:"):
            ref = g[k]
            assert np.abs(tr.last_grads[k[5:]].numpy() - ref).max() < 1e-3 * np.abs(ref).max() + 1e-7, k


@pytest.mark.skipif(not os.path.exists(REF_PB), reason="/root/reference only exists in the build container")
def test_weights_equal_the_reference_pb():
    ref = pb_reader.unet_params_from_pb(REF_PB)
    w = _weights()
    for k, v in w.items():
        assert np.array_equal(ref[k], v.numpy()), k


@pytest.mark.skipif(not os.path.exists(os.path.join(SHIP, "sem_dataset.npz")), reason="dataset slice not staged")
def test_full_image_iou_matches_known_answers_on_two_images():
    ka = json.load(open(os.path.join(GOLD, "pb_known_answers.json")))["models"]["GAN"]["per_image_iou_0.5"]
    with np.load(os.path.join(SHIP, "sem_dataset.npz")) as z:
        imgs, masks = z["images"], np.unpackbits(z["masks"], axis=-1).astype(bool)
    torch.set_num_threads(os.cpu_count())
    for i in (0, 17):
        x = imgs[i].astype(np.float32)
        x = (x - x.min()) / (x - x.min()).max()
        with torch.no_grad():
            y, _ = OU.unet_forward(torch.from_numpy(x)[None, :, :, None], _weights(), training=False)
        p = y[0, :, :, 0].numpy() > 0.5
        iou = np.logical_and(p, masks[i]).sum() / np.logical_or(p, masks[i]).sum()
        assert abs(iou - ka[i]) < 1e-4
