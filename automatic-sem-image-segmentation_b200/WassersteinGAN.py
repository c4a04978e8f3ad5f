"""Drop-in for Releases/Version 1.2.0/WassersteinGAN.py (workflow step 1): `WGAN` with the reference's attributes and
methods, `WGAN_GP` = wgan_model.WganGpModel, `GANMonitor`.

WGAN.__init__ :288-357 (dataset: masks thresholded at 0.5, normalised to [-1, 1], four flips per mask, zero-padded to a common
size divisible by 16), start_training :359-373, create_model :700-722, get_discriminator_model / get_generator_model :569-684
(built inside the model, see wgan_nets.py), discriminator_loss / generator_loss :690-698.

`simulate_masks` (:375-545, workflow step 2) is host code (numpy / OpenCV / scipy.ndimage) around generator inference; the
`opensimplex` noise it clusters particles with is restated in simplex_noise.py.  Differences to the reference: the engine is specialised to one batch size, so an
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

    # ---- workflow step 2: synthetic masks from the trained generator (reference :375-545; HOST code) ---------------------------------
    def _load_latest_model(self):
        """:398-400: the most recent `model.keras` under model_dir."""
        runs = sorted(d for d in os.listdir(self.model_dir) if os.path.isfile(os.path.join(self.model_dir, d, 'model.keras')))
        if not runs:
            raise FileNotFoundError(f"no trained WGAN under {self.model_dir}: run start_training() first")
        cfg, named = keras_io.load_keras(os.path.join(self.model_dir, runs[-1], 'model.keras'), rename=lambda s: s)
        h, w = cfg["image_shape"][:2]
        model = WGAN_GP(image_shape=(h, w, 1), batch_size=self.batch_size, latent_dim=int(cfg["latent_dim"]), dtype=self.dtype)
        for tag, net in model.nets.items():
            net.set_named({n: named[f"{tag}/{n}"] for n in net.names})
        return model

    def _generate_particles(self, count):
        """`count` generator samples as uint8 images (:484-499), drawn batch by batch with training=False."""
        out = []
        while sum(len(o) for o in out) < count:
            z = np.random.standard_normal((self.batch_size, self.n_z)).astype(np.float32)
            out.append(self.model(z, training=False))
        samples = np.concatenate(out, 0)[:count]
        return (samples * 127.5 + 127.5)[:, :, :, 0].astype('uint8')

    @staticmethod
    def _grid_positions(grid_type, span_w, span_h, spacing_h, spacing_w, jitter_h, jitter_w):
        """Particle anchor points on a jittered hexagonal / cubic grid over [0, span] (:427-458).  spacing_* are the un-truncated
        grid_spacing_factor * particle size values: the reference truncates them differently in different places."""
        step_h, step_w = int(spacing_h), int(spacing_w)
        if grid_type == 'HEXAGONAL':
            shift = int(spacing_w / 2)                  # odd rows are offset by half a cell
            pts = []
            for k, y in enumerate(range(0, span_h, step_h)):
                for x in range(0, span_w, step_w):
                    if x + (k % 2) * shift > span_w:
                        break
                    pts.append((x + (k % 2) * shift, y))
            # The reference pre-allocates ceil(span_h / spacing_h) * ceil(span_w / spacing_w) + 1 slots (:428); the ones it never fills
            # stay at the origin and are jittered like real grid points -- kept (never fewer slots than points: the reference would
            # raise an IndexError there).
            slots = max(math.ceil(span_h / spacing_h) * math.ceil(span_w / spacing_w) + 1, len(pts))
            pos_x, pos_y = np.zeros(slots, dtype='int32'), np.zeros(slots, dtype='int32')
            if pts:
                pos_x[:len(pts)], pos_y[:len(pts)] = np.asarray(pts, dtype='int32').T
        else:
            # sic (:447): the reference steps BOTH axes by the spacing derived from the particle HEIGHT
            pos_y, pos_x = np.mgrid[0:span_h:step_h, 0:span_w:step_h]
            pos_x, pos_y = pos_x.flatten(), pos_y.flatten()
        pos_x = pos_x + np.random.randint(-jitter_w, jitter_w, pos_x.size)
        pos_y = pos_y + np.random.randint(-jitter_h, jitter_h, pos_y.size)
        return np.clip(pos_x, 0, span_w), np.clip(pos_y, 0, span_h)

    def simulate_masks(self, no_of_images=1, min_no_of_particles=100, max_no_of_particles=150, use_normal_distribution=False,
                       sigma=0.10, mu=1.0, min_scaling=0.75, max_scaling=1.25, use_perlin_noise=True, perlin_noise_threshold=0.5,
                       perlin_noise_frequency=4, use_random_rotation='DISABLE', max_overlap=0.01, grid_type='DISABLE',
                       grid_spacing_factor=0.125, grid_noise_factor=0.05, img_width=384, img_height=384):
        """Places rotated / scaled generator samples on a canvas (clustered by 2-D simplex noise, on a jittered grid when an
        overlap limit is set), resolves overlaps, crops the centre and writes `generate_dir/<index>.tif`; five random results
        are copied to ../testB.  Same parameters, defaults and quirks as the reference (:375-545)."""
        import random
        from shutil import copy

        import cv2
        from scipy import ndimage

        from . import simplex_noise

        ph, pw = self.train_images.shape[1], self.train_images.shape[2]
        d = math.ceil(math.sqrt((max_scaling * ph) ** 2 + (max_scaling * pw) ** 2))       # margin: one particle diagonal (before mu/sigma)
        if self.model is None:
            self.model = self._load_latest_model()
        if use_normal_distribution:
            min_scaling, max_scaling = mu - 3 * sigma, mu + 3 * sigma
        needs_noise = use_perlin_noise or use_random_rotation == 'PERLIN'
        if max_overlap is not None and grid_type not in ('HEXAGONAL', 'CUBIC'):
            grid_type = 'HEXAGONAL'                      # (:412-413) an overlap limit implies the hexagonal grid
        on_grid = grid_type in ('HEXAGONAL', 'CUBIC')
        span_w, span_h = img_width + 2 * d, img_height + 2 * d
        os.makedirs(self.generate_dir, exist_ok=True)
        for i in range(no_of_images):
            canvas = np.zeros((img_height + 3 * d, img_width + 3 * d), dtype='uint8')
            count = 0 if on_grid else random.randint(min_no_of_particles, max_no_of_particles)
            noise = None
            if needs_noise:
                simplex_noise.random_seed()
                f = perlin_noise_frequency
                ix, iy = np.arange(0, f, f / (img_width + 3 * d)), np.arange(0, f, f / (img_height + 3 * d))
                noise = simplex_noise.noise2array(iy, ix)                 # indexed [x, y] like the reference's call (:422)
                noise -= np.min(noise)
                noise /= np.max(noise) / 2
                noise = noise - 1                                         # range [-1, 1]
            level = 2 * perlin_noise_threshold - 1
            if on_grid:
                pos_x, pos_y = self._grid_positions(grid_type, span_w, span_h, grid_spacing_factor * ph, grid_spacing_factor * pw,
                                                    int(grid_noise_factor * ph), int(grid_noise_factor * pw))
                if use_perlin_noise:
                    keep = noise[pos_x, pos_y] > level
                    pos_x, pos_y = pos_x[keep], pos_y[keep]
                count = len(pos_x)
            elif use_perlin_noise:
                allowed = np.argwhere(noise > level)                      # every position above the threshold is equally likely
                pick = allowed[np.random.choice(len(allowed), count, replace=False)]
                pos_x, pos_y = pick[:, 0], pick[:, 1]
            else:
                pos_x, pos_y = np.random.randint(0, span_w, count), np.random.randint(0, span_h, count)
            if count == 0:
                raise ValueError("simulate_masks: no particle position left (perlin_noise_threshold too high for this grid?)")
            if use_normal_distribution:
                scalings = np.random.normal(mu, sigma, count)
            else:
                scalings = np.random.uniform(min_scaling, max_scaling, count)
            scalings = np.clip(scalings, min_scaling, max_scaling)
            if use_random_rotation == 'RANDOM':
                rotations = np.random.randint(0, 360, count)
            elif use_random_rotation == 'PERLIN':
                rotations = noise[pos_y, pos_x] * 180                     # sic (:477): [y, x] here, [x, y] everywhere else
            else:
                rotations = np.zeros(count)
            for j, p in enumerate(self._generate_particles(count)):
                height, width = p.shape
                centre = (width / 2, height / 2)
                mat = cv2.getRotationMatrix2D(centre, float(rotations[j]), float(scalings[j]))
                c, s_ = abs(mat[0, 0]), abs(mat[0, 1])
                out_w, out_h = int(width * c + height * s_), int(width * s_ + height * c)
                mat[0, 2] += out_w / 2 - centre[0]
                mat[1, 2] += out_h / 2 - centre[1]
                p = cv2.warpAffine(p, mat, (out_w, out_h)) > 127
                p = ndimage.binary_opening(ndimage.binary_fill_holes(p), structure=np.ones((9, 9)))
                core = ndimage.binary_erosion(p, iterations=2)
                if not core.any():
                    continue
                y0, x0 = int(pos_y[j]), int(pos_x[j])
                win = canvas[y0:y0 + p.shape[0], x0:x0 + p.shape[1]]
                if win.shape != p.shape:
                    # free (non-grid) positions are drawn over the whole canvas and can reach its far edge, where the reference
                    # fails with a shape error (:527-530); here the particle is clipped at the canvas border (outside the crop)
                    p, core = p[:win.shape[0], :win.shape[1]], core[:win.shape[0], :win.shape[1]]
                    if not core.any():
                        continue
                if max_overlap is not None and np.logical_and(win, core).sum() > max_overlap * core.sum():
                    continue
                win -= np.logical_and(win, p).astype('uint8')              # carve the full footprint, stamp the eroded one: a gap
                win += core.astype('uint8')                                #   of two pixels separates touching particles
            a, b = int((canvas.shape[0] - img_height) / 2), int((canvas.shape[1] - img_width) / 2)
            Image.fromarray(canvas[a:a + img_height, b:b + img_width] * 255).save(os.path.join(self.generate_dir, '{:05d}.tif'.format(i)))
        files = [f for f in os.listdir(self.generate_dir) if '.tif' in f or '.png' in f or '.bmp' in f]
        test_dir = os.path.join(self.generate_dir, '..', 'testB')
        os.makedirs(test_dir, exist_ok=True)
        for f in random.sample(files, min(5, len(files))):
            copy(os.path.join(self.generate_dir, f), test_dir)
