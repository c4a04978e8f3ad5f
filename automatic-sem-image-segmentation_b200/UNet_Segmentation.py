"""Drop-in for Releases/Version 1.2.0/UNet_Segmentation.py on the sm_100a engine.

Same class / method / attribute names and argument meaning as the reference (UNet :147-562, ImageDataset :21-101,
DataLoader :104-122, DataSet :125-144); the Keras graph is replaced by `sem_b200.UNetModel` (hand-written CUDA behind
the C ABI).  Differences a caller can observe:
  * `model.keras` / `Checkpoint_Lowest_Loss.keras` are Keras-3 style zip archives whose weights store is the npz variant
    (`model.weights.npz`; h5py is not available) and whose config.json is this package's own description (keras_io.py);
  * `run_inference(..., use_gpu=False)` still runs on the GPU: there is no CPU path in this package;
  * `segment()` (Otsu + distance-transform watershed with lines) is restated from scikit-image's published algorithms
    (Measurements.py); scikit-image itself is not installable here.
"""
from __future__ import annotations

import math
import os
import random
import time
from datetime import datetime

import numpy as np
from PIL import Image

from . import HelperFunctions, keras_compat
from .keras_compat import ReflectionPadding2D  # noqa: F401  (module-level class of the reference, :565-589)
from .model import CSVLogger, LearningRateScheduler, ModelCheckpoint, UNetModel, load_model


class ImageDataset:
    """File list with the reference's deterministic 80/20 split (seed 1234) and x4 flip augmentation."""

    def __init__(self, image_dir, mask_dir, contrast_optimization_range=(0.5, 99.5), use_brightness_and_contrast_augmentation=False):
        self.image_ids, self.image_info, self.type = [], {}, ""
        self.image_dir, self.mask_dir = image_dir, mask_dir
        self.contrast_optimization_range = contrast_optimization_range
        self.use_brightness_and_contrast_augmentation = use_brightness_and_contrast_augmentation

    def add_image(self, image_id, path, mask, augmentation):
        self.image_info[image_id] = {"id": image_id, "image_path": path, "mask_path": mask, "augmentation": augmentation}
        self.image_ids.append(image_id)

    def initialize_images(self, subset, train_val_split=0.8, seed=1234):
        assert subset in ["train", "val"]
        files = HelperFunctions.get_image_file_paths_from_directory(self.image_dir)
        random.Random(seed).shuffle(files)
        cut = int(train_val_split * len(files))
        chosen = files[:cut] if subset == "train" else files[cut:]
        self.type = subset
        for i, path in enumerate(chosen):
            for aug in range(4):
                self.add_image(f"{i:05d}_augmentation_{aug}", path, path.replace(self.image_dir, self.mask_dir), aug)

    def load_from_file(self, image_ids, is_mask):
        if isinstance(image_ids, str):
            image_ids = [image_ids]
        out = []
        for image_id in image_ids:
            info = self.image_info[image_id]
            if is_mask:
                img = HelperFunctions.load_and_preprocess_images(info["mask_path"], normalization_range=(0, 1), threshold_value=0.5)[0]
            elif self.type == "train" and self.use_brightness_and_contrast_augmentation:
                c = random.random() * 2
                img = HelperFunctions.load_and_preprocess_images(info["image_path"], normalization_range=(0 - random.random(), 1 + random.random()),
                                                                 contrast_optimization_range=(c, c + 98))[0]
                img -= np.min(img)
                img /= np.max(img)
            else:
                img = HelperFunctions.load_and_preprocess_images(info["image_path"], normalization_range=(0, 1),
                                                                 contrast_optimization_range=self.contrast_optimization_range)[0]
            aug = info["augmentation"]
            if aug in (1, 3):
                img = np.fliplr(img)
            if aug in (2, 3):
                img = np.flipud(img)
            out.append(img)
        return np.asarray(out, dtype="float32")


class DataLoader:
    """keras.utils.Sequence over an ImageDataset (loads from disk per batch; ceil(len/batch) batches)."""

    def __init__(self, dataset, batch_size=1, shuffle=True, **kwargs):
        self.dataset, self.batch_size, self.shuffle = dataset, batch_size, shuffle
        self.all_image_ids = dataset.image_ids.copy()

    def __len__(self):
        return math.ceil(len(self.all_image_ids) / self.batch_size)

    def __getitem__(self, idx):
        ids = self.all_image_ids[idx * self.batch_size:(idx + 1) * self.batch_size]
        return self.dataset.load_from_file(ids, is_mask=False), self.dataset.load_from_file(ids, is_mask=True)

    def on_epoch_end(self):
        if self.shuffle:
            np.random.shuffle(self.all_image_ids)


class DataSet:
    """keras.utils.Sequence over in-memory arrays (floor(len/batch) batches: the last partial batch is dropped)."""

    def __init__(self, x, y, batch_size=1, shuffle=True, **kwargs):
        self.x, self.y, self.batch_size, self.shuffle = x, y, batch_size, shuffle

    def __len__(self):
        return self.x.shape[0] // self.batch_size

    def __getitem__(self, idx):
        s = slice(idx * self.batch_size, (idx + 1) * self.batch_size)
        return self.x[s], self.y[s]

    def on_epoch_end(self):
        if self.shuffle:
            perm = list(range(self.x.shape[0]))
            random.shuffle(perm)
            self.x, self.y = np.asarray(self.x[perm], dtype="float32"), np.asarray(self.y[perm], dtype="float32")


class UNet:
    def __init__(self, root_dir, image_dir, mask_dir, allow_memory_growth=True, use_gpus_no=(0,)):
        self.root_dir = os.path.join(root_dir, "3_UNet")
        self.model_dir = os.path.join(self.root_dir, "Models")
        self.image_dir, self.mask_dir = image_dir, mask_dir
        self.use_dataloader = False
        self.contrast_optimization_range = (1, 99)
        self.prefix = time.strftime("%Y-%m-%d_%H-%M-%S", time.localtime())
        self.batch_size = 1
        self.epochs = 100
        self.learning_rate = 0.001
        self.loss_function = "binary_crossentropy"
        self.lr_decay = "STEP_DECAY"
        self.image_shape = (384, 384, 1)
        self.filters = 16
        self.output_channels = 1
        self.allow_memory_growth = allow_memory_growth
        self.use_gpus_no = use_gpus_no
        self.dtype = os.environ.get("SEMB_DTYPE", "bf16")       # storage mode of the engine: bf16 (throughput) | f32 (parity)
        self.inference_batch_size = 16      # tiles per engine call in run_inference(tile_images=True)
        self.dataset_train = self.dataset_val = self.training_data = self.validation_data = self.model = None

    # ---- data ------------------------------------------------------------------------------------------
    def load_images(self, subset):
        assert subset in ["train", "val"]
        ds = self.dataset_train if subset == "train" else self.dataset_val
        if self.use_dataloader:
            return DataLoader(ds, self.batch_size)
        print(f"Importing {len(ds.image_ids)} augmented {subset} images: {datetime.now()}")
        x = ds.load_from_file(ds.image_ids, is_mask=False)
        y = ds.load_from_file(ds.image_ids, is_mask=True)
        return DataSet(x, y, self.batch_size)

    def step_decay(self, epoch, current_lr, drop=0.5, epochs_drop=10):
        return current_lr * drop if (epoch + 1) % epochs_drop == 0 else current_lr

    def linear_decay(self, epoch, current_lr):
        return self.learning_rate * (1 - epoch / float(self.epochs))

    # ---- model -----------------------------------------------------------------------------------------
    # The reference's static layer functions (:401-503), same names / argument order, operating on symbolic tensors of
    # keras_compat.Input(...): each call records engine ops.  Only what the MultiRes-UNet uses is accepted: square 1x1 / 3x3
    # kernels, 'same' padding, stride 1 (2x2 stride 2 for the transposed conv).
    @staticmethod
    def conv2d_bn(x, filters, num_row, num_col, padding="same", strides=(1, 1), activation="relu", name=None):
        if num_row != num_col or num_row not in (1, 3) or padding != "same" or tuple(strides) != (1, 1):
            raise NotImplementedError("conv2d_bn: the engine builds 1x1 / 3x3, 'same', stride-1 convolutions (what multi_res_unet uses)")
        return keras_compat.builder_of(x).conv2d_bn(x, int(filters), int(num_row), activation)

    @staticmethod
    def trans_conv2d_bn(x, filters, num_row, num_col, padding="same", strides=(2, 2), name=None):
        """Defined but never called by the reference's own graph (multi_res_unet uses a bare Conv2DTranspose, :542-551);
        the BatchNormalization after a transposed conv has no fused kernel here."""
        raise NotImplementedError("trans_conv2d_bn is unused by multi_res_unet (UNet_Segmentation.py:542-551) and not built")

    @staticmethod
    def multi_res_block(u, inp, alpha=1.67):
        return keras_compat.builder_of(inp).multi_res_block(int(u), inp, alpha=alpha)

    @staticmethod
    def res_path(filters, length, inp):
        return keras_compat.builder_of(inp).res_path(int(filters), int(length), inp)

    @staticmethod
    def multi_res_unet(inputs, output_channels=1, conv_filters=16, dtype=None, batch_size=None) -> UNetModel:
        """MultiResUNet (reference :505-562).  `inputs`: keras_compat.Input(shape=(H, W, 1), batch_size=, dtype=) like the
        reference's `keras.layers.Input`, or a plain (H, W, 1) shape tuple (then dtype / batch_size may be given here)."""
        if isinstance(inputs, keras_compat.T):
            e = inputs.view.buf.eng
            shape, dtype, batch_size = (inputs.h, inputs.w, inputs.layout.logical), dtype or e.dtype_name, batch_size or e.N
        else:
            shape = tuple(inputs)
        return UNetModel(shape, conv_filters, output_channels, dtype=dtype or "bf16", batch_size=batch_size or 1)

    def create_model(self):
        """weighting = #zeros/#ones over the training masks (:364-376); Adam(lr) + weighted BCE (:379-395)."""
        if self.use_dataloader:
            zeros = ones = 0
            tmp = None
            for image_id in self.dataset_train.image_ids:
                tmp = np.array(self.dataset_train.load_from_file(image_id, is_mask=True))
                zeros += np.count_nonzero(tmp == 0)
                ones += np.count_nonzero(tmp)
            self.image_shape = tmp.shape[1:3]
        else:
            y = self.training_data.y
            zeros, ones = np.count_nonzero(y == 0), np.count_nonzero(y)
            self.image_shape = y.shape[1:3]
        weighting = zeros / ones
        model = UNet.multi_res_unet((self.image_shape[0], self.image_shape[1], 1), self.output_channels, self.filters,
                                    dtype=self.dtype, batch_size=self.batch_size)
        model.compile(weighting=weighting, learning_rate=self.learning_rate)
        return model

    def run_training(self):
        out_dir = os.path.join(self.model_dir, self.prefix)
        os.makedirs(out_dir, exist_ok=True)
        self.dataset_train = ImageDataset(self.image_dir, self.mask_dir, self.contrast_optimization_range)
        self.dataset_val = ImageDataset(self.image_dir, self.mask_dir, self.contrast_optimization_range)
        self.dataset_train.initialize_images("train")
        self.dataset_val.initialize_images("val")
        self.training_data = self.load_images("train")
        self.validation_data = self.load_images("val")
        self.model = self.create_model()
        callbacks = [ModelCheckpoint(os.path.join(out_dir, "Checkpoint_Lowest_Loss.keras"), monitor="loss", verbose=1, save_best_only=True, mode="min"),
                     CSVLogger(os.path.join(out_dir, "training_log.csv"), separator=";", append=True)]
        if self.lr_decay == "STEP_DECAY":
            callbacks.append(LearningRateScheduler(self.step_decay))
        elif self.lr_decay == "LINEAR_DECAY":
            callbacks.append(LearningRateScheduler(self.linear_decay))
        print("Start training the model: " + str(datetime.now()))
        self.model.fit(self.training_data, batch_size=self.batch_size, epochs=self.epochs, verbose=1, callbacks=callbacks,
                       validation_data=self.validation_data)
        path = os.path.join(out_dir, "model.keras")
        print("Saving model to: " + path)
        self.model.save(path)
        return self.model

    def run_inference(self, files, output_directory, model=None, tile_images=False, threshold=-1, watershed_lines=True,
                      min_distance=9, min_overlap=2, manage_overlap_mode=2, use_gpu=False):
        if self.model is None:
            if model is None:
                newest = sorted(os.listdir(self.model_dir))[-1]
                cand = [os.path.join(self.model_dir, newest, f) for f in ("model.keras", "model.npz")]
                self.model = load_model(next((c for c in cand if os.path.exists(c)), cand[0]), dtype=self.dtype)
            elif isinstance(model, str):
                self.model = load_model(model, dtype=self.dtype)
            else:
                self.model = model
        elif model is not None:
            self.model = model
        images = HelperFunctions.load_and_preprocess_images(files, normalization_range=(0, 1),
                                                            contrast_optimization_range=self.contrast_optimization_range)
        names = HelperFunctions.get_image_file_paths_from_directory(files)
        os.makedirs(output_directory, exist_ok=True)
        from . import dp
        for i in dp.inference_indices(len(images)):           # every image here; a rank-strided share under torchrun
            img = images[i]
            if tile_images:
                th, tw = self.image_shape[0], self.image_shape[1]
                # tile grid, batched forward and stitching all on the device (the reference tiles / stitches on the host and
                # calls the model tile by tile, :338-340)
                out = self.model.predict_tiled(img, tw, th, min_overlap=min_overlap, manage_overlap_mode=manage_overlap_mode,
                                               batch_size=self.inference_batch_size)
            else:
                out = self.model(img[None], training=False).numpy()[0]      # fully convolutional: any H, W
            out = out[:, :, 0].copy()
            stem = os.path.splitext(os.path.split(names[i])[-1])[0]
            Image.fromarray(out).save(os.path.join(output_directory, stem + "_raw.tif"))
            out -= np.min(out)
            out /= np.max(out)
            mask = HelperFunctions.segment((out * 255).astype(np.uint8), threshold=threshold, watershed_lines=watershed_lines,
                                           min_distance=min_distance, use_four_connectivity=True)
            Image.fromarray(mask).save(os.path.join(output_directory, os.path.split(names[i])[-1]))

    @staticmethod
    def to_numpy_array(x):
        return np.asarray(x.cpu() if hasattr(x, "cpu") else x).copy()
