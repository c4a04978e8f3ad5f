"""Host-side helpers with the reference's names and argument meaning (Releases/Version 1.2.0/HelperFunctions.py).

Only what the conv-stack hot path touches is mirrored: image loading / normalisation (:290-329), the overlapping
tile grid and its inverse (:17-141) and a threshold-only `segment`.  The classical post-processing chain (watershed,
Li filter, particle measurements) is outside the hot path and is not reproduced here.
"""
from __future__ import annotations

import math
import os

import numpy as np
from PIL import Image

IMAGE_EXTENSIONS = (".tif", ".tiff", ".png", ".bmp", ".jpg", ".jpeg", ".gif")


def get_image_file_paths_from_directory(directory):
    return [os.path.join(directory, f) for f in os.listdir(directory) if f.endswith(IMAGE_EXTENSIONS)]


def load_and_preprocess_images(input_dir_or_filelist, threshold_value=None, normalization_range=(-1, 1), output_channels=1,
                               contrast_optimization_range=None):
    """HelperFunctions.py:296-329: optional percentile clipping, min-max to [0,1], optional threshold, affine to range."""
    if isinstance(input_dir_or_filelist, (str, os.PathLike)):
        files = (get_image_file_paths_from_directory(input_dir_or_filelist) if os.path.isdir(input_dir_or_filelist)
                 else [input_dir_or_filelist])
    else:
        files = list(input_dir_or_filelist)
    out = []
    for f in files:
        img = np.array(Image.open(f), dtype="float32")
        assert 2 <= img.ndim <= 3 and output_channels in (1, 3), "Invalid Image format"
        if img.ndim == 3 and output_channels == 1:
            img = np.average(img, -1)
        if img.ndim == 2:
            img = img[:, :, None]
        r = contrast_optimization_range
        if r is not None and r[0] > 0 and r[1] < 100:
            lo, hi = np.percentile(img, r[0]), np.percentile(img, r[1])
            img = np.clip(img, lo, hi)
        if normalization_range is not None:
            img = img - img.min()
            img = img / img.max()
            if threshold_value is not None:
                img = (img > threshold_value).astype("float32")
            img = normalization_range[0] + (normalization_range[1] - normalization_range[0]) * img
        out.append(img.astype("float32"))
    if len({im.shape for im in out}) > 1:
        # images of different sizes (the single-particle masks of the WGAN, WassersteinGAN.py:331-352): the numpy of the
        # reference's era returned an object array here, today's raises -- keep the old contract (an iterable of arrays)
        arr = np.empty(len(out), dtype=object)
        for i, im in enumerate(out):
            arr[i] = im
        return arr
    return np.array(out, dtype="float32")


def _grid(size: int, tile: int, min_overlap: int):
    """number of tiles and their offsets along one axis (HelperFunctions.py:21-47)"""
    n = math.ceil(size / tile)
    if n > 1 and (tile - size % tile) % tile <= min_overlap:
        n += 1
    if n == 1:
        return 1, [0]
    step = tile - (tile * n - size) / (n - 1)
    return n, [math.ceil(i * step) for i in range(n)]


def tile_image(img, tile_size_w, tile_size_h, min_overlap=2, normalization_range=None, normalize_tiles_individually=True):
    """Overlapping tiles, x-major order (outer loop over columns) like the reference; a tile reaching past the image
    edge is zero filled."""
    h, w = img.shape[0], img.shape[1]
    nx, xs = _grid(w, tile_size_w, min_overlap)
    ny, ys = _grid(h, tile_size_h, min_overlap)
    tiles = np.zeros((nx * ny, tile_size_h, tile_size_w, 1), dtype="float32")
    k = 0
    for ox in xs:
        for oy in ys:
            patch = img[oy:min(oy + tile_size_h, h), ox:min(ox + tile_size_w, w), :]
            tiles[k, :patch.shape[0], :patch.shape[1], :] = patch
            k += 1
    if normalization_range is not None:
        lo, hi = normalization_range
        if normalize_tiles_individually:
            for t in tiles:
                t -= t.min()
                t /= t.max()
                t *= (hi - lo)
                t += lo
        else:
            tiles -= img.min()
            tiles /= img.max()
            tiles = lo + (hi - lo) * tiles
    return tiles


def stitch_image(img, image_size_w, image_size_h, min_overlap=2, manage_overlap_mode=2, return_8_bit_image=False):
    """Inverse of tile_image; manage_overlap_mode 0 = maximum, 1 = average, 2 = crop half of the overlap from each tile."""
    th, tw, c = img.shape[1], img.shape[2], img.shape[-1]
    nx, xs = _grid(image_size_w, tw, min_overlap)
    ny, ys = _grid(image_size_h, th, min_overlap)
    out = np.zeros((image_size_h, image_size_w, c), dtype="float32")
    count = np.zeros_like(out, dtype="uint8")
    ovx = (tw * nx - image_size_w) // (2 * (nx - 1)) if nx > 1 else 0
    ovy = (th * ny - image_size_h) // (2 * (ny - 1)) if ny > 1 else 0
    k = 0
    for i, ox in enumerate(xs):
        for j, oy in enumerate(ys):
            y1, x1 = min(oy + th, image_size_h), min(ox + tw, image_size_w)
            if manage_overlap_mode == 0:
                out[oy:y1, ox:x1] = np.maximum(img[k, :y1 - oy, :x1 - ox], out[oy:y1, ox:x1])
            elif manage_overlap_mode == 1:
                out[oy:y1, ox:x1] += img[k, :y1 - oy, :x1 - ox]
                count[oy:y1, ox:x1] += 1
            else:
                cl = 0 if i == 0 else ovx
                cr = 0 if i == nx - 1 else ovx
                ct = 0 if j == 0 else ovy
                cb = 0 if j == ny - 1 else ovy
                yb, xb = min(oy + th - cb, image_size_h), min(ox + tw - cr, image_size_w)
                out[oy + ct:yb, ox + cl:xb] = img[k, ct:ct + (yb - oy - ct), cl:cl + (xb - ox - cl)]
            k += 1
    if manage_overlap_mode == 1:
        out = out / count
    out = np.asarray(out, dtype="float32")
    if return_8_bit_image:
        out = np.asarray(out * 255, dtype="uint8")
    return out


def eight_to_four_connected(img):
    """HelperFunctions.py:144-152: breaks diagonal-only (8-connected) contacts.  The reference's scan is sequential
    (every fix is visible to the following 2x2 windows), so it is reproduced pixel by pixel; `count > 2 or count <
    size - 2` is the reference's own guard."""
    if np.count_nonzero(img) > 2 or np.count_nonzero(img) < img.size - 2:
        a = img
        for x in range(a.shape[0] - 1):
            r0, r1 = a[x], a[x + 1]
            # candidate columns only: a diagonal pair of zeros next to a diagonal pair of non-zeros
            d1 = (r0[:-1] == 0) & (r1[1:] == 0) & (r1[:-1] != 0) & (r0[1:] != 0)
            d2 = (r1[:-1] == 0) & (r0[1:] == 0) & (r0[:-1] != 0) & (r1[1:] != 0)
            for y in np.flatnonzero(d1 | d2):
                if a[x, y] == 0 and a[x + 1, y + 1] == 0 and a[x + 1, y] != 0 and a[x, y + 1] != 0:
                    a[x + 1, y] = 0
                elif a[x + 1, y] == 0 and a[x, y + 1] == 0 and a[x, y] != 0 and a[x + 1, y + 1] != 0:
                    a[x, y] = 0
    return img


def threshold_otsu(image_u8: np.ndarray) -> float:
    """Otsu's threshold (skimage.filters.threshold_otsu semantics); see Measurements.threshold_otsu."""
    from .Measurements import threshold_otsu as _otsu
    return _otsu(image_u8)


def segment(image, threshold=-1, watershed_lines=True, min_distance=9, use_four_connectivity=True):
    """HelperFunctions.py:155-160: Measure.segment (threshold, Otsu when < 0; distance-transform watershed with lines) and
    the 8 -> 4 connectivity fix.  uint8 {0, 255}."""
    from .Measurements import Measure
    labels = Measure.segment(image, threshold, watershed_lines, min_distance, darkBackground=True)
    if use_four_connectivity:
        labels = eight_to_four_connected(labels)
    return labels


# ---- workflow plumbing of steps 0 and 5 (reference :188-287, :163-185): host-side file handling, no network math ---------
def initialize_directories(root_dir, output_dir_cyclegan, output_dir_unet):
    """The 1_WGAN / 2_CycleGAN / 3_UNet tree the steps hand their results through (HelperFunctions.py:188-238)."""
    for parts in (("1_WGAN", "Output_Images"), ("1_WGAN", "Models"), ("2_CycleGAN", "data", "testA"), ("2_CycleGAN", "data", "testB"),
                  ("2_CycleGAN", "data", "trainA"), ("2_CycleGAN", "data", "trainB"), ("2_CycleGAN", "generate_images", "A"),
                  ("2_CycleGAN", "generate_images", "B"), ("2_CycleGAN", "generate_images", "Synthetic_Masks_Filtered"),
                  ("2_CycleGAN", "images"), ("2_CycleGAN", "Models"), ("3_UNet", "Models")):
        os.makedirs(os.path.join(root_dir, *parts), exist_ok=True)
    os.makedirs(output_dir_cyclegan, exist_ok=True)
    os.makedirs(output_dir_unet, exist_ok=True)


def prepare_images_cycle_gan(root_dir, input_dir_images, tile_size_w=384, tile_size_h=384, num_simulated_masks=1000, dark_background=True):
    """Tiles the input images into 2_CycleGAN/data/trainA (tiles that are mostly background are skipped), copies five of
    them to testA and tops the set up with random flipped crops until there are `num_simulated_masks` (:241-287)."""
    import random
    from shutil import copy
    input_imgs = load_and_preprocess_images(input_dir_or_filelist=input_dir_images, normalization_range=None, output_channels=1)
    filenames = get_image_file_paths_from_directory(input_dir_images)
    train_a = os.path.join(root_dir, "2_CycleGAN", "data", "trainA")

    def foreground(tile, img):
        return (dark_background and np.mean(tile) >= 1.1 * np.mean(img)) or (not dark_background and np.mean(tile) <= 0.9 * np.mean(img))

    for i, img in enumerate(input_imgs):
        tiles = np.asarray(tile_image(img, tile_size_w, tile_size_h, normalization_range=(0, 255), min_overlap=0), dtype="uint8")
        f = os.path.split(filenames[i])[-1]
        ext = os.path.splitext(f)[-1]
        for j, t in enumerate(tiles):
            if foreground(t, img):
                Image.fromarray(t[:, :, 0]).save(os.path.join(train_a, f.replace(ext, f"-{j}{ext}")))
    tiles_a = get_image_file_paths_from_directory(train_a)
    for f in random.sample(tiles_a, min(5, len(tiles_a))):
        copy(f, os.path.join(root_dir, "2_CycleGAN", "data", "testA"))
    have, i = len(os.listdir(train_a)), 0
    while i < num_simulated_masks - have:
        r = random.randint(0, input_imgs.shape[0] - 1)
        f = os.path.split(filenames[r])[-1]
        ext = os.path.splitext(f)[-1]
        img = input_imgs[r]
        a = random.randint(0, img.shape[0] - tile_size_h - 1)
        b = random.randint(0, img.shape[1] - tile_size_w - 1)
        t = img[a:a + tile_size_h, b:b + tile_size_w]
        if random.random() > 0.5:
            t = np.fliplr(t)
        if random.random() > 0.5:
            t = np.flipud(t)
        if foreground(t, img):
            Image.fromarray(t[:, :, 0].astype("uint8")).save(os.path.join(train_a, f.replace(ext, f"-aug_{i}{ext}")))
            i += 1


def filter_gan_masks(img_path, msk_path, out_path, threshold_method=None, do_watershed_and_four_connectivity=True,
                     gaussian_blur_amount=0.0, dark_background=True):
    """Step 5 (:163-185): per file, optionally re-segment the mask (watershed + 4-connectivity), then keep only the
    particles whose mean grey value in the image lies on the particle side of a global threshold (Li by default)."""
    import cv2
    from PIL import ImageFilter
    from .Measurements import contours_filtered_by_mean_intensity, threshold_li
    threshold_method = threshold_method or threshold_li
    os.makedirs(out_path, exist_ok=True)
    for f in os.listdir(img_path):
        if not f.endswith(IMAGE_EXTENSIONS) or not os.path.exists(os.path.join(msk_path, f)):
            continue
        img = np.array(Image.open(os.path.join(img_path, f)), dtype="uint8")
        mask = np.array(Image.open(os.path.join(msk_path, f)), dtype="uint8")
        if do_watershed_and_four_connectivity:
            mask = segment(image=mask, threshold=-1, watershed_lines=True, use_four_connectivity=True)
        elif np.any((mask > 1) & (mask < 255)) or np.all(mask <= 1):
            from .Measurements import Measure
            mask = Measure.segment(mask, threshold=-1.0, applyWatershed=False, darkBackground=dark_background)
        t = threshold_method(img)
        keep = contours_filtered_by_mean_intensity(mask, img, min_value=t) if dark_background else \
            contours_filtered_by_mean_intensity(mask, img, min_value=0.0, max_value=t)
        out = np.zeros(img.shape[:2], dtype="uint8")
        cv2.drawContours(image=out, contours=keep, contourIdx=-1, color=(255, 255, 255), thickness=-1)
        Image.fromarray(out).save(os.path.join(out_path, f))
        if gaussian_blur_amount > 0:
            Image.fromarray(img).filter(ImageFilter.GaussianBlur(gaussian_blur_amount)).save(os.path.join(img_path, f))
