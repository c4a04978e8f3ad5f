"""Drop-in for Releases/Version 1.2.0/CycleGAN.py (class CycleGAN :20-317, DataLoader :454-479) on the sm_100a engine.

Same attribute-style configuration and method names; `create_model()` returns `sem_b200.CycleGanModel` (networks and
train step in hand-written CUDA).  Supported configuration = what StartProcess.py drives: no skip connection, no
Gaussian noise in the discriminators, transposed-conv upsampling, MAE cycle / identity losses (other switches raise).
Models are saved as `.keras`-style zip archives (keras_io.py).
"""
from __future__ import annotations

import collections
import os
import time

import numpy as np
from PIL import Image

from . import HelperFunctions, keras_compat
from .cyclegan_model import CycleGanModel, DiscriminatorModel, GeneratorModel, ImagePool
from .keras_compat import ReflectionPadding2D  # noqa: F401  (module-level class of the reference, :482-506)


class DataLoader:
    def __init__(self, train_a, train_b, batch_size=1, use_dataloader=False, scale_for_binary_crossentropy=False, invert_images=False, **kw):
        self.batch_size, self.train_a, self.train_b = batch_size, train_a, train_b
        self.use_dataloader, self.scale_for_binary_crossentropy, self.invert_images = use_dataloader, scale_for_binary_crossentropy, invert_images

    def __len__(self):
        return int(min(len(self.train_a), len(self.train_b)) / float(self.batch_size))

    def __getitem__(self, idx):
        s = slice(idx * self.batch_size, (idx + 1) * self.batch_size)
        a, b = self.train_a[s], self.train_b[s]
        if self.use_dataloader:
            a = CycleGAN.load_images(a, False, invert=self.invert_images)
            b = CycleGAN.load_images(b, self.scale_for_binary_crossentropy)
        return np.asarray(a), np.asarray(b)

    def on_epoch_end(self):
        np.random.shuffle(self.train_a)
        np.random.shuffle(self.train_b)


class CycleGAN:
    def __init__(self, root_dir="./", image_shape=(384, 384, 1), allow_memory_growth=True, use_gpus_no=(0,)):
        self.batch_size = 2
        self.epochs = 50
        self.learning_rate = 2e-4
        self.use_data_loader = False
        self.filters = 32
        self.num_downsampling_blocks_gen = 3
        self.num_residual_blocks_gen = 9
        self.num_upsampling_blocks_gen = 3
        self.num_downsampling_blocks_disc = 2
        self.allow_memory_growth, self.use_gpus_no = allow_memory_growth, use_gpus_no
        self.lambda_cycle_a = self.lambda_cycle_b = 10
        self.use_binary_crossentropy = False
        self.use_linear_decay = True
        self.decay_epoch = int(0.75 * self.epochs)
        self.lambda_identity_a = self.lambda_identity_b = 0.5
        self.use_skip_connection = True
        self.use_resize_convolution = False
        self.label_smoothing_factor = 0.0
        self.gaussian_noise_value = 0.15
        self.invert_images = False
        self.image_pool_size = 50
        self.inference_batch_size = 8       # tiles per engine call in run_inference(tile_images=True)
        self.model = self.data = None
        self.root_dir = root_dir
        self.model_dir = os.path.join(root_dir, "2_CycleGAN", "Models")
        self.image_shape = image_shape
        self.prefix = time.strftime("%Y-%m-%d_%H-%M-%S", time.localtime())
        self.dtype = os.environ.get("SEMB_DTYPE", "bf16")
        # the pools are created with the batch size known HERE (2); StartProcess assigns batch_size afterwards
        # (reference quirk) -- kept.
        self.image_pool_a = ImagePool(batch_size=self.batch_size, pool_size=self.image_pool_size)
        self.image_pool_b = ImagePool(batch_size=self.batch_size, pool_size=self.image_pool_size)
        d = os.path.join(root_dir, "2_CycleGAN", "data")
        ls = lambda s: HelperFunctions.get_image_file_paths_from_directory(os.path.join(d, s)) if os.path.isdir(os.path.join(d, s)) else []
        self.train_a, self.test_a, self.train_b, self.test_b = ls("trainA"), ls("testA"), ls("trainB"), ls("testB")

    def create_model(self) -> CycleGanModel:
        if self.use_binary_crossentropy:
            # CycleGAN.py:117-121: BinaryCrossentropy cycle / identity losses and a sigmoid head on gen_a
            raise NotImplementedError("use_binary_crossentropy=True is outside the accelerated path (MAE cycle / identity losses, "
                                      "tanh heads only); StartProcess.py leaves it off")
        if self.num_upsampling_blocks_gen != self.num_downsampling_blocks_gen:
            raise ValueError("num_upsampling_blocks_gen must equal num_downsampling_blocks_gen (the output would change size)")
        model = CycleGanModel(image_shape=self.image_shape, batch_size=self.batch_size, filters=self.filters, dtype=self.dtype,
                              n_res=self.num_residual_blocks_gen, lambda_cycle_a=self.lambda_cycle_a, lambda_cycle_b=self.lambda_cycle_b,
                              lambda_identity_a=self.lambda_identity_a, lambda_identity_b=self.lambda_identity_b,
                              image_pool_a=self.image_pool_a, image_pool_b=self.image_pool_b,
                              label_smoothing_factor=self.label_smoothing_factor, use_skip_connection=self.use_skip_connection,
                              use_resize_convolution=self.use_resize_convolution, gaussian_noise_value=self.gaussian_noise_value,
                              n_down=self.num_downsampling_blocks_gen, n_up=self.num_upsampling_blocks_gen,
                              n_down_disc=self.num_downsampling_blocks_disc)
        model.compile(learning_rate=self.learning_rate, beta_1=0.5)
        return model

    # ---- network builders (reference :323-451), same names and argument order ----------------------------------------------
    def get_resnet_generator(self, filters=64, num_downsampling_blocks=2, num_residual_blocks=9, num_upsample_blocks=2, name=None,
                             use_binary_crossentropy=False, padding="same"):
        """A stand-alone generator at `self.image_shape` (batch `self.batch_size`): callable, `get_weights / set_weights`."""
        if use_binary_crossentropy:
            raise NotImplementedError("sigmoid head / BinaryCrossentropy losses (CycleGAN.py:417-418) are not built")
        if padding != "same" or num_upsample_blocks != num_downsampling_blocks:
            raise NotImplementedError("generators are built with 'same' strided convs and as many up- as down-sampling blocks")
        return GeneratorModel(self.image_shape, self.batch_size, filters=filters, dtype=self.dtype, n_res=num_residual_blocks,
                              use_skip_connection=self.use_skip_connection, use_resize_convolution=self.use_resize_convolution,
                              n_down=num_downsampling_blocks, n_up=num_upsample_blocks)

    def get_discriminator(self, filters=32, num_downsampling_blocks=3, name=None, padding="same"):
        """A stand-alone PatchGAN (4x4 convs; the reference's create_model passes padding='valid', :148-150)."""
        if padding != "valid":
            raise NotImplementedError("the PatchGAN is built with padding='valid' (what create_model uses since 1.2.0, CycleGAN.py:148)")
        return DiscriminatorModel(self.image_shape, self.batch_size, filters=filters, dtype=self.dtype,
                                  gaussian_noise=self.gaussian_noise_value, n_down=num_downsampling_blocks)

    def residual_block(self, input_tensor, activation, kernel_size=(3, 3), strides=(1, 1), padding="valid", use_bias=False):
        """reference :323-337 on a symbolic tensor of a gan_nets graph (GT)."""
        if tuple(kernel_size) != (3, 3) or tuple(strides) != (1, 1) or padding != "valid" or use_bias:
            raise NotImplementedError("residual_block: 3x3, stride 1, reflect-padded 'valid', bias-free (the reference's own call)")
        return input_tensor.view.buf.eng._builder.residual_block(input_tensor, keras_compat.act_code(activation))

    def downsample(self, x, filters, activation, kernel_size=(3, 3), strides=(2, 2), padding="same", use_bias=False):
        """reference :339-345: 3x3 'same' (generator) or 4x4 'valid' (discriminator), stride 2, InstanceNorm, activation."""
        k = int(kernel_size[0])
        if tuple(strides) != (2, 2) or use_bias or (k, padding) not in ((3, "same"), (4, "valid")):
            raise NotImplementedError("downsample: 3x3 'same' or 4x4 'valid', stride 2, bias-free")
        return x.view.buf.eng._builder.downsample(x, int(filters), keras_compat.act_code(activation), k=k, padding=padding)

    def upsample(self, x, filters, activation, kernel_size=(3, 3), strides=(2, 2), padding="same", use_bias=False):
        """reference :347-358: Conv2DTranspose 3x3 stride 2 'same' (or the resize-convolution variant), InstanceNorm, activation."""
        if tuple(kernel_size) != (3, 3) or tuple(strides) != (2, 2) or padding != "same" or use_bias:
            raise NotImplementedError("upsample: 3x3, stride 2, 'same', bias-free")
        return x.view.buf.eng._builder.upsample(x, int(filters), keras_compat.act_code(activation),
                                                use_resize_convolution=self.use_resize_convolution)

    def generator_loss_fn(self, fake):
        t = (1.0 - self.label_smoothing_factor) + self.label_smoothing_factor / 2
        return float(np.mean((t - np.asarray(fake)) ** 2))

    def discriminator_loss_fn(self, real, fake):
        s = self.label_smoothing_factor
        real_loss = float(np.mean(((1.0 - s) + s / 2 - np.asarray(real)) ** 2))
        fake_loss = float(np.mean((s / 2 - np.asarray(fake)) ** 2))
        return (real_loss + fake_loss) * 0.5, real_loss, fake_loss

    def linear_decay(self, epoch, current_lr):
        if epoch < self.decay_epoch:
            return self.learning_rate
        return self.learning_rate * (1 - (epoch - self.decay_epoch) / float(self.epochs - self.decay_epoch))

    @staticmethod
    def load_images(image_list, scale_for_binary_crossentropy=False, invert=False):
        r = (0, 1) if scale_for_binary_crossentropy else (-1, 1)
        images = HelperFunctions.load_and_preprocess_images(image_list, threshold_value=None, normalization_range=r, output_channels=1)
        return images * -1.0 if invert else images

    def start_training(self):
        out_dir = os.path.join(self.model_dir, self.prefix)
        os.makedirs(out_dir, exist_ok=True)
        self.decay_epoch = int(0.75 * self.epochs)
        if not self.use_data_loader:
            self.train_a = self.load_images(self.train_a, False, invert=self.invert_images)
            self.train_b = self.load_images(self.train_b, self.use_binary_crossentropy)
        self.data = DataLoader(self.train_a, self.train_b, self.batch_size, self.use_data_loader, self.use_binary_crossentropy, self.invert_images)
        self.model = self.create_model()
        log = os.path.join(out_dir, "training_log.csv")
        for epoch in range(self.epochs):
            # NOTE: in the reference the LearningRateScheduler only touches the unused base optimizer of CycleGanModel,
            # not the four Adam instances; the four learning rates therefore stay constant.
            agg = {}
            for i in range(len(self.data)):
                for k, v in self.model.train_step(self.data[i]).items():
                    agg[k] = agg.get(k, 0.0) + v
            logs = {k: v / max(len(self.data), 1) for k, v in agg.items()}
            new = not os.path.exists(log)
            with open(log, "a") as fh:
                if new:
                    fh.write(";".join(["epoch"] + sorted(logs)) + "\n")
                fh.write(";".join([str(epoch)] + [repr(logs[k]) for k in sorted(logs)]) + "\n")
            print(f"Epoch {epoch + 1}/{self.epochs} - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()), flush=True)
            self.save(os.path.join(out_dir, f"checkpoints_{epoch + 1:03d}.keras"))
            self.data.on_epoch_end()
        self.save(os.path.join(out_dir, "model.keras"))
        return self.model

    def save(self, path):
        """One `.keras`-style zip per model (all four networks; keras_io), or a flat .npz for a path ending in .npz."""
        arrs, order = {}, []
        for name, net in self.model.nets.items():
            for n, w in zip(net.names, net.get_weights()):
                arrs[f"{name}/{n}"] = w
                order.append(f"{name}/{n}")
        if path.endswith(".npz"):
            np.savez(path, **arrs)
            return
        from . import keras_io
        cfg = {"class": "CycleGanModel", "image_shape": list(self.image_shape), "filters": self.filters,
               "use_skip_connection": self.use_skip_connection, "use_resize_convolution": self.use_resize_convolution,
               "num_residual_blocks_gen": self.num_residual_blocks_gen, "format": "semb200-keras-1"}
        keras_io.save_keras(path, cfg, arrs, order, rename=lambda s: s)

    @staticmethod
    def _read(path):
        from . import keras_io
        if keras_io.is_keras_archive(path):
            return keras_io.load_keras(path, rename=lambda s: s)[1]
        with np.load(path) as z:
            return {k: z[k] for k in z.files}

    def load(self, path):
        self.model = self.model or self.create_model()
        z = self._read(path)
        for name, net in self.model.nets.items():
            net.set_named({n: z[f"{name}/{n}"] for n in net.names})
        return self.model

    def run_inference(self, files, output_directory, source_domain, model=None, tile_images=False, min_overlap=2,
                      manage_overlap_mode=2, use_gpu=False):
        images = HelperFunctions.load_and_preprocess_images(files, normalization_range=(-1, 1))
        names = HelperFunctions.get_image_file_paths_from_directory(files)
        which = "gen_a" if "a" in source_domain.lower() else "gen_b"
        os.makedirs(output_directory, exist_ok=True)
        weights_from = model if isinstance(model, str) else None
        if self.model is None and weights_from is None:
            newest = sorted(os.listdir(self.model_dir))[-1]
            cand = [os.path.join(self.model_dir, newest, f) for f in ("model.keras", "model.npz")]
            weights_from = next((c for c in cand if os.path.exists(c)), cand[0])
        cache = collections.OrderedDict()          # at most two generator instances alive (LRU): sizes vary per image
        named = None
        if weights_from is not None:
            z = self._read(weights_from)
            named = {k[len(which) + 1:]: v for k, v in z.items() if k.startswith(which + "/")}
        from . import dp
        for i in dp.inference_indices(len(images)):           # every image here; a rank-strided share under torchrun
            img = images[i]
            if which == "gen_a" and self.invert_images:
                img = img * -1
            th, tw = (self.image_shape[0], self.image_shape[1]) if tile_images else (img.shape[0], img.shape[1])
            tiles = HelperFunctions.tile_image(img, tw, th, min_overlap=min_overlap) if tile_images else img[None]
            nb = min(len(tiles), self.inference_batch_size)
            key = (nb, th, tw)
            if key in cache:
                cache.move_to_end(key)
            else:          # generators are fully convolutional: ONE generator tower at this size, same weights
                while len(cache) >= 2:
                    cache.popitem(last=False)
                g = GeneratorModel((th, tw, 1), nb, filters=self.filters, dtype=self.dtype, n_res=self.num_residual_blocks_gen,
                                   use_skip_connection=self.use_skip_connection, use_resize_convolution=self.use_resize_convolution,
                                   n_down=self.num_downsampling_blocks_gen, n_up=self.num_upsampling_blocks_gen)
                if named is not None:
                    g.set_named(named)
                elif self.model is not None:
                    g.set_weights(self.model.nets[which].get_weights())
                cache[key] = g
            pred = cache[key].predict(tiles)
            out = (HelperFunctions.stitch_image(pred, img.shape[1], img.shape[0], min_overlap=min_overlap,
                                                manage_overlap_mode=manage_overlap_mode) if tile_images else pred[0])[:, :, 0].copy()
            if which == "gen_b" and self.invert_images:
                out *= -1
            out -= np.min(out)
            out /= np.max(out)
            Image.fromarray((out * 255).astype(np.uint8)).save(os.path.join(output_directory, os.path.split(names[i])[-1]))


class GANMonitor:
    """reference :810-905: after every epoch, A -> B -> A and B -> A -> B reconstructions of `num_img` test images as
    side-by-side mosaics (input | translated | cycled) written to `output_dir` (8-bit TIFF; the reference also overlays
    the mask on the image, which is presentation only)."""

    def __init__(self, test_a, test_b, output_dir, num_img=2):
        self.num_img, self.test_a, self.test_b, self.output_dir = num_img, test_a, test_b, output_dir
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        self.plot_reconstruction(self.model, epoch + 1, nex=self.num_img)

    def plot_reconstruction(self, model, epoch, nex=2):
        n = min(nex, len(self.test_a), len(self.test_b))
        if n == 0:
            return None
        a = np.asarray(self.test_a[:n], dtype=np.float32)
        b = np.asarray(self.test_b[:n], dtype=np.float32)

        def pad(x):                                  # the towers are specialised to the training batch size
            out = np.zeros((model.n,) + x.shape[1:], dtype=np.float32)
            out[:x.shape[0]] = x
            return out
        fake_b = model.generate("gen_a", pad(a))
        cyc_a = model.generate("gen_b", fake_b)
        fake_a = model.generate("gen_b", pad(b))
        cyc_b = model.generate("gen_a", fake_a)
        u8 = lambda x: np.clip((x[..., 0] + 1.0) * 127.5, 0, 255).astype(np.uint8)
        rows_aba = [np.concatenate([u8(a)[i], u8(fake_b)[i], u8(cyc_a)[i]], 1) for i in range(n)]
        rows_bab = [np.concatenate([u8(b)[i], u8(fake_a)[i], u8(cyc_b)[i]], 1) for i in range(n)]
        os.makedirs(self.output_dir, exist_ok=True)
        pa, pb = os.path.join(self.output_dir, f"epoch_{epoch:03d}_ABA.tif"), os.path.join(self.output_dir, f"epoch_{epoch:03d}_BAB.tif")
        Image.fromarray(np.concatenate(rows_aba, 0)).save(pa)
        Image.fromarray(np.concatenate(rows_bab, 0)).save(pb)
        return pa, pb
