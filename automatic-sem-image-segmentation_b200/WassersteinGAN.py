"""Drop-in for Releases/Version 1.2.0/WassersteinGAN.py (workflow step 1): `WGAN` with the reference's attributes and
methods, `WGAN_GP` = wgan_model.WganGpModel, `GANMonitor`.

WGAN.__init__ :288-357 (dataset: masks thresholded at 0.5, normalised to [-1, 1], four flips per mask, zero-padded to a common
size divisible by 16), start_training :359-373, create_model :700-722, get_discriminator_model / get_generator_model :569-684
(built inside the model, see wgan_nets.py), discriminator_loss / generator_loss :690-698.

Not part of the package: `simulate_masks` (:375-545, workflow step 2) -- host-side mask synthesis that needs the `opensimplex`
noise library; it raises with that reason.  Differences to the reference: the engine is specialised to one batch size, so an
epoch runs floor(len / batch_size) full batches of a reshuffled dataset (Keras' last partial batch is dropped; a dataset
smaller than one batch is tiled up to it); sample mosaics are written with PIL (no matplotlib)."""
from __future__ import annotations

import csv
import math
import os
import time

import numpy as np
from PIL import Image

from . import HelperFunctions, keras_io
from .wgan_model import METRICS, WganGpModel

WGAN_GP = WganGpModel


class GANMonitor:
    """reference :259-284: every `output_epochs` epochs a 3-column mosaic of `num_img` generated masks."""

    def __init__(self, output_dir, num_img=9, latent_dim=128, output_epochs=100):
        self.num_img, self.latent_dim, self.epochs, self.output_dir = num_img, latent_dim, output_epochs, output_dir
        self.model = None
        self.rng = np.random.default_rng(0)

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        if epoch % int(self.epochs) == 0:
            return self.plot_reconstruction(self.model, epoch, nex=self.num_img)

    def plot_reconstruction(self, model, epoch, nex=9):
        nex = min(nex, model.n)
        samples = model(self.rng.standard_normal((nex, self.latent_dim)).astype(np.float32))
        cols = 3
        rows = math.ceil(nex / float(cols))
        h, w = samples.shape[1:3]
        sheet = np.zeros((rows * h, cols * w), dtype=np.uint8)
        for i, s in enumerate(samples):
            r, c = divmod(i, cols)
            sheet[r * h:(r + 1) * h, c * w:(c + 1) * w] = np.clip(s[:, :, 0] * 127.5 + 127.5, 0, 255).astype(np.uint8)
        os.makedirs(self.output_dir, exist_ok=True)
        path = os.path.join(self.output_dir, "Epoch_{:05d}.png".format(epoch))
        Image.fromarray(sheet).save(path)
        return path


class WGAN:
    def __init__(self, root_dir, allow_memory_growth=True, use_gpus_no=(0,), dtype=None):
        self.root_dir = os.path.join(root_dir, '1_WGAN')
        self.input_dir = os.path.join(root_dir, 'Input_Masks')
        self.output_dir = os.path.join(self.root_dir, 'Output_Images')
        self.model_dir = os.path.join(self.root_dir, 'Models')
        self.generate_dir = os.path.join(root_dir, '2_CycleGAN', 'data', 'trainB')
        self.batch_size = 64
        self.epochs = 1000
        self.n_z = 128
        self.model = None
        self.allow_memory_growth, self.use_gpus_no = allow_memory_growth, use_gpus_no
        self.dtype = dtype or os.environ.get("SEMB_DTYPE", "bf16")
        self.train_images = []
        max_image_height = max_image_width = 0
        images = HelperFunctions.load_and_preprocess_images(input_dir_or_filelist=self.input_dir, threshold_value=0.5,
                                                            normalization_range=(-1, 1), output_channels=1, contrast_optimization_range=None)
        for image in images:
            max_image_height = max([max_image_height, image.shape[0]])
            max_image_width = max([max_image_height, image.shape[1]])        # sic (:333): the HEIGHT enters the width maximum
            self.train_images.append(image.copy())
            self.train_images.append(np.fliplr(image.copy()))
            self.train_images.append(np.flipud(image.copy()))
            self.train_images.append(np.flipud(np.fliplr(image.copy())))
        if max_image_height % 2 ** 4 != 0:
            max_image_height = (max_image_height // (2 ** 4) + 1) * 2 ** 4
        if max_image_width % 2 ** 4 != 0:
            max_image_width = (max_image_width // (2 ** 4) + 1) * 2 ** 4
        for i, image in enumerate(self.train_images):
            if image.shape[0] < max_image_height or image.shape[1] < max_image_width:
                img = np.zeros((max_image_height, max_image_width, 1), dtype='float32')
                t, l = (max_image_height - image.shape[0]) // 2, (max_image_width - image.shape[1]) // 2
                img[t:t + image.shape[0], l:l + image.shape[1], :] = image[:, :, :]
                self.train_images[i] = img
        self.train_images = np.asarray(self.train_images, dtype='float32')
        self.prefix = time.strftime('%Y-%m-%d_%H-%M-%S', time.localtime())

    # the loss functions of the reference (:690-698), on arrays of logits
    @staticmethod
    def discriminator_loss(real_img, fake_img):
        return float(np.mean(fake_img) - np.mean(real_img))

    @staticmethod
    def generator_loss(fake_img):
        return float(-np.mean(fake_img))

    def create_model(self):
        """:700-722: Adam(2e-4, beta_1 0.5, beta_2 0.9) for both networks, three critic updates per generator update."""
        h, w = self.train_images.shape[1:3]
        model = WGAN_GP(image_shape=(h, w, 1), batch_size=self.batch_size, latent_dim=self.n_z, discriminator_extra_steps=3,
                        dtype=self.dtype)
        model.compile(learning_rate=0.0002, beta_1=0.5, beta_2=0.9)
        return model

    def start_training(self):
        os.makedirs(os.path.join(self.model_dir, self.prefix), exist_ok=True)
        os.makedirs(os.path.join(self.output_dir, self.prefix), exist_ok=True)
        self.model = self.create_model()
        cbk = GANMonitor(output_dir=os.path.join(self.output_dir, self.prefix), num_img=9, latent_dim=self.n_z, output_epochs=20)
        cbk.set_model(self.model)
        log_path = os.path.join(self.model_dir, self.prefix, 'training_log.csv')
        data, bs = self.train_images, self.batch_size
        if len(data) < bs:
            data = np.concatenate([data] * math.ceil(bs / len(data)), 0)
        rng = np.random.default_rng(0)
        with open(log_path, "a", newline="") as fh:
            wr = csv.writer(fh)
            wr.writerow(["epoch"] + METRICS)
            for epoch in range(self.epochs):
                order = rng.permutation(len(data))
                sums, steps = dict.fromkeys(METRICS, 0.0), 0
                for s in range(len(data) // bs):
                    logs = self.model.train_step(data[order[s * bs:(s + 1) * bs]])
                    for k in METRICS:
                        sums[k] += logs[k]
                    steps += 1
                means = {k: sums[k] / max(steps, 1) for k in METRICS}      # keras.metrics.Mean over the epoch
                wr.writerow([epoch] + [means[k] for k in METRICS])
                fh.flush()
                cbk.on_epoch_end(epoch, means)
        self.save(os.path.join(self.model_dir, self.prefix, 'model.keras'))
        return self.model

    def save(self, path):
        named, order = {}, []
        for tag, net in self.model.nets.items():
            for n, wt in zip(net.names, net.get_weights()):
                named[f"{tag}/{n}"] = wt
                order.append(f"{tag}/{n}")
        h, w = self.train_images.shape[1:3]
        keras_io.save_keras(path, {"class_name": "WGAN_GP", "image_shape": [int(h), int(w), 1], "latent_dim": int(self.n_z),
                                   "discriminator_extra_steps": 3, "gp_weight": 10.0}, named, order, rename=lambda s: s)

    def simulate_masks(self, *args, **kwargs):
        raise NotImplementedError("simulate_masks (WassersteinGAN.py:375-545, workflow step 2) is host-side mask synthesis on top of "
                                  "the `opensimplex` noise library, which this package does not restate; run it with the reference on "
                                  "the generator trained here (weights: <model_dir>/<prefix>/model.keras)")
