"""The reference's whole workflow (StartProcess.py:55-221: steps 0, 1, 2, 3, 4, 5, 6a, 6b) on a tiny synthetic data set, every step
through the drop-in step functions of sem_b200.StartProcess with the networks on the sm_100a engine.  Sizes and epoch counts are
shrunk (64x64 tiles, one epoch each, 8 base filters for the CycleGAN) so that the run takes about a minute; the step functions,
directory tree, file hand-over between steps and the classes' attribute protocol are the real ones.

One stub: a WGAN trained for ONE epoch draws noise, which the morphological clean-up of step 2 would delete entirely, so the
particles handed to `simulate_masks` are discs (the generator itself is exercised by step 1 and tests/test_wgan_gpu.py)."""
import os

import numpy as np
import pytest
from PIL import Image

pytestmark = pytest.mark.gpu


def _disc(size, r, cy=None, cx=None):
    yy, xx = np.mgrid[0:size, 0:size]
    cy, cx = size // 2 if cy is None else cy, size // 2 if cx is None else cx
    return (yy - cy) ** 2 + (xx - cx) ** 2 < r * r


def test_all_workflow_steps_on_a_tiny_dataset(tmp_path, monkeypatch):
    import random
    from sem_b200 import StartProcess as SP, WassersteinGAN
    root = str(tmp_path)
    rng = np.random.default_rng(0)
    os.makedirs(os.path.join(root, "Input_Images"))
    os.makedirs(os.path.join(root, "Input_Masks"))
    for i in range(2):                                      # "SEM images": bright discs on a dark, noisy background
        img = rng.normal(40, 8, (160, 160))
        for _ in range(14):
            cy, cx, r = rng.integers(10, 150), rng.integers(10, 150), rng.integers(7, 12)
            img[_disc(160, r, cy, cx)] = rng.normal(190, 10)
        Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(os.path.join(root, "Input_Images", f"sem{i}.tif"))
    for i, r in enumerate((9, 11, 13)):                     # single-particle masks for the WGAN
        Image.fromarray(_disc(32, r).astype(np.uint8) * 255).save(os.path.join(root, "Input_Masks", f"p{i}.tif"))
    cfg = dict(ROOT_DIR=root, INPUT_DIR_IMAGES=os.path.join(root, "Input_Images"), INPUT_DIR_MASKS=os.path.join(root, "Input_Masks"),
               OUTPUT_DIR_CYCLEGAN=os.path.join(root, "Output_Masks_CycleGAN"), OUTPUT_DIR_UNET=os.path.join(root, "Output_Masks_UNet"),
               TILE_SIZE_W=64, TILE_SIZE_H=64, NUM_SIMULATED_MASKS=8, RUN_INFERENCE_ON_WHOLE_IMAGE=True, WGAN_BATCH_SIZE=8, WGAN_EPOCHS=1,
               MAX_PARTICLE_OVERLAP=0.5, CYCLEGAN_BATCH_SIZE=2, CYCLEGAN_EPOCHS=1, CYCLEGAN_FILTERS=8, UNET_BATCH_SIZE=2, UNET_EPOCHS=1,
               UNET_FILTERS=16, USE_DATALOADER=True)
    for k, v in cfg.items():
        monkeypatch.setattr(SP, k, v)
    monkeypatch.setenv("SEMB_DTYPE", "bf16")
    random.seed(0)
    np.random.seed(0)
    ls = lambda *p: sorted(os.listdir(os.path.join(root, *p)))

    SP.start_step_0()
    assert len(ls("2_CycleGAN", "data", "trainA")) >= 8 and len(ls("2_CycleGAN", "data", "testA")) >= 1

    SP.start_step_1()
    (run,) = ls("1_WGAN", "Models")
    assert {"model.keras", "training_log.csv"} <= set(ls("1_WGAN", "Models", run))

    disc = _disc(32, 10).astype(np.uint8) * 255
    monkeypatch.setattr(WassersteinGAN.WGAN, "_generate_particles", lambda self, count: np.repeat(disc[None], count, 0))
    SP.start_step_2()
    masks = ls("2_CycleGAN", "data", "trainB")
    assert len(masks) >= 8 and len(ls("2_CycleGAN", "data", "testB")) == 5
    m0 = np.array(Image.open(os.path.join(root, "2_CycleGAN", "data", "trainB", masks[0])))
    assert m0.shape == (64, 64) and set(np.unique(m0)) <= {0, 255} and (m0 > 0).any()

    SP.start_step_3()
    (run,) = ls("2_CycleGAN", "Models")
    assert {"model.keras", "checkpoints_001.keras", "training_log.csv"} <= set(ls("2_CycleGAN", "Models", run))
    log = open(os.path.join(root, "2_CycleGAN", "Models", run, "training_log.csv")).read().strip().splitlines()
    assert len(log) == 2 and all(np.isfinite(float(v)) for v in log[1].split(";"))

    SP.start_step_4()
    assert ls("2_CycleGAN", "generate_images", "A") == masks                      # one fake image per simulated mask
    assert ls("2_CycleGAN", "generate_images", "B") == ["sem0.tif", "sem1.tif"]
    fake = np.array(Image.open(os.path.join(root, "2_CycleGAN", "generate_images", "A", masks[0])))
    assert fake.shape == (64, 64) and fake.dtype == np.uint8

    SP.start_step_5()
    assert ls("2_CycleGAN", "generate_images", "Synthetic_Masks_Filtered") == masks
    assert ls("Output_Masks_CycleGAN") == ["sem0.tif", "sem1.tif"]

    SP.start_step_6a()
    (run,) = ls("3_UNet", "Models")
    assert any(f.endswith(".keras") for f in ls("3_UNet", "Models", run))

    SP.start_step_6b()
    assert ls("Output_Masks_UNet") == ["sem0.tif", "sem0_raw.tif", "sem1.tif", "sem1_raw.tif"]      # segmentation + raw probability map
    out = np.array(Image.open(os.path.join(root, "Output_Masks_UNet", "sem0.tif")))
    assert out.shape == (160, 160) and set(np.unique(out)) <= {0, 255}
