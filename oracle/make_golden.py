"""Generate the golden fixtures that pin the oracle.  TEST INFRASTRUCTURE.

Run once in the build container (needs /root/reference):
    python -m oracle.make_golden [--skip-iou]

Writes
  tests/golden/pb_known_answers.json   mean whole-image IoU of the reference's shipped
                                       UNet weights on the reference's Datasets/ (oracle fwd)
  tests/golden/unet_gan_weights.npz    TiO2_UNet_Masks_GAN.pb weights (fp32, oracle names)
  tests/golden/sem_crops.npz           4 SEM crops 256x256 (uint8) + manual masks + oracle outputs
  tests/golden/unet_small_step.npz     seeded tiny train step (loss, grads, post-Adam weights)
  baseline/_ref/sem_dataset.npz        (git-ignored, ships to the GPU box) 40 SEM images rows 0:704
                                       + manual masks, for the full-size GPU IoU parity run
  baseline/_ref/unet_{GAN,Manual,TSEM}_weights.npz

IoU definition: Archive/Other Scripts/Calculate_Scores.py:69-70
(sum(and)/sum(or)), ground truth Datasets/.../TiO2_Masks_Manual_4connected.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
from PIL import Image

from . import pb_reader, unet

REF = "/root/reference"
PB_DIR = os.path.join(REF, "ImageJ Plugin", "SEM_Particle_Segmentation_Models")
SEM_DIR = os.path.join(REF, "Datasets", "Electron Microscopy Images", "SEM")
MASK_DIR = os.path.join(REF, "Datasets", "Electron Microscopy Image Masks", "TiO2_Masks_Manual_4connected")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SHIP = os.path.join(ROOT, "baseline", "_ref")
ROWS = 704


def whole_image_iou(a, b):
    return float(np.logical_and(a, b).sum() / max(np.logical_or(a, b).sum(), 1))


def load_dataset():
    files = sorted(f for f in os.listdir(SEM_DIR) if f.endswith(".tif"))
    imgs, masks = [], []
    for f in files:
        im = np.array(Image.open(os.path.join(SEM_DIR, f)))[:ROWS]
        mk = np.array(Image.open(os.path.join(MASK_DIR, f.replace(".tif", "_m.tif"))))[:ROWS] > 0
        imgs.append(im.astype(np.uint8))
        masks.append(mk)
    return files, np.stack(imgs), np.stack(masks)


def normalise(img_u8):
    x = img_u8.astype(np.float32)
    x -= x.min()
    x /= x.max()
    return x


def to_torch_params(npd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in npd.items()}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-iou", action="store_true")
    ap.add_argument("--stage-only", action="store_true",
                    help="only (re)write the git-ignored baseline/_ref files that travel to the GPU box; tests/golden is untouched")
    args = ap.parse_args(argv)
    os.makedirs(GOLD, exist_ok=True)
    os.makedirs(SHIP, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    spec = unet.UNetSpec(16)

    files, imgs, masks = load_dataset()
    np.savez_compressed(os.path.join(SHIP, "sem_dataset.npz"), files=np.array(files), images=imgs,
                        masks=np.packbits(masks, axis=-1))
    weights = {}
    for tag in ("GAN", "Manual", "TSEM"):
        p = pb_reader.unet_params_from_pb(os.path.join(PB_DIR, f"TiO2_UNet_Masks_{tag}.pb"))
        p = {n: p[n] for n in spec.names()}
        weights[tag] = p
        np.savez_compressed(os.path.join(SHIP, f"unet_{tag}_weights.npz"), **p)
    if args.stage_only:
        print("staged", SHIP)
        return 0
    np.savez_compressed(os.path.join(GOLD, "unet_gan_weights.npz"), **weights["GAN"])

    # ---- known answers: mean IoU of the shipped weights on the shipped dataset
    if not args.skip_iou:
        ka = {"rows": ROWS, "n_images": len(files), "iou_definition": "Calculate_Scores.py:69-70",
              "preprocess": "palette index -> float32, per-image min-max to [0,1], whole image, plain threshold",
              "models": {}}
        for tag in ("GAN", "Manual", "TSEM"):
            P = to_torch_params(weights[tag])
            per = {0.3: [], 0.5: [], 0.7: []}
            with torch.no_grad():
                for i in range(len(files)):
                    x = torch.from_numpy(normalise(imgs[i]))[None, :, :, None]
                    y, _ = unet.unet_forward(x, P, training=False)
                    y = y[0, :, :, 0].numpy()
                    for t in per:
                        per[t].append(whole_image_iou(y > t, masks[i]))
                    print(tag, files[i], f"{per[0.5][-1]:.4f}", flush=True)
            ka["models"][tag] = {
                "mean_iou": {str(t): float(np.mean(v)) for t, v in per.items()},
                "min_iou_0.5": float(np.min(per[0.5])), "max_iou_0.5": float(np.max(per[0.5])),
                "per_image_iou_0.5": [round(float(v), 6) for v in per[0.5]],
            }
        with open(os.path.join(GOLD, "pb_known_answers.json"), "w") as f:
            json.dump(ka, f, indent=1)

    # ---- 4 crops with oracle outputs (GAN weights), the GPU parity fixture
    P = to_torch_params(weights["GAN"])
    sel = [(0, 100, 200), (7, 300, 600), (19, 448, 0), (33, 0, 768)]
    crops = np.stack([imgs[i][r:r + 256, c:c + 256] for i, r, c in sel])
    cmask = np.stack([masks[i][r:r + 256, c:c + 256] for i, r, c in sel])
    x = torch.from_numpy(np.stack([normalise(c) for c in crops]))[..., None]
    with torch.no_grad():
        y, _ = unet.unet_forward(x, P, training=False)
    np.savez_compressed(os.path.join(GOLD, "sem_crops.npz"), crops=crops, masks=np.packbits(cmask, axis=-1),
                        oracle_sigmoid=y[..., 0].numpy().astype(np.float32), sel=np.array(sel))

    # ---- tiny seeded train step (pins the oracle's training semantics against regressions)
    torch.manual_seed(0)
    p0 = spec.init_params(seed=0)
    gx = torch.Generator().manual_seed(123)
    xs = torch.rand(2, 32, 32, 1, generator=gx)
    ys = (torch.rand(2, 32, 32, 1, generator=gx) < 0.2).float()
    wgt = float((ys == 0).sum() / (ys == 1).sum())
    tr = unet.UNetTrainer(spec, p0, wgt)
    logs, yp = tr.train_step(xs, ys)
    out = {"x": xs.numpy(), "y": ys.numpy(), "weighting": np.float64(wgt), "loss": np.float64(logs["loss"]),
           "mae": np.float64(logs["mae"]), "acc": np.float64(logs["acc"]), "y_pred": yp.numpy()}
    for n in ("conv2d_1/kernel", "conv2d_2/kernel", "conv2d_20/kernel", "conv2d_transpose_1/kernel",
              "conv2d_transpose_1/bias", "batch_normalization_5/gamma", "batch_normalization_5/beta",
              "conv2d_57/kernel"):
        out["grad:" + n] = tr.last_grads[n].numpy()
        out["new:" + n] = tr.params[n].detach().numpy()
    out["new:batch_normalization_1/moving_mean"] = tr.params["batch_normalization_1/moving_mean"].numpy()
    out["new:batch_normalization_1/moving_variance"] = tr.params["batch_normalization_1/moving_variance"].numpy()
    np.savez_compressed(os.path.join(GOLD, "unet_small_step.npz"), **out)
    print("golden written to", GOLD)


if __name__ == "__main__":
    sys.exit(main())
