"""Classical post-processing of Releases/Version 1.2.0/Measurements.py that the UNet / CycleGAN inference paths call:
`Measure.segment` (:263-305) = Otsu threshold -> Euclidean distance map -> Gaussian smoothing -> local maxima ->
marker watershed with watershed lines.

HOST code (numpy / scipy): the reference runs this step on the CPU after the network (UNet_Segmentation.py:350), it is
not part of the accelerated conv path.  scikit-image is not installable in this environment, so
`threshold_otsu`, `peak_local_max` and `watershed` are restated here from scikit-image's published algorithms
(skimage.filters.threshold_otsu, skimage.feature.peak_local_max, skimage.segmentation.watershed: priority flood ordered
by (value, age), pixels labelled when they leave the queue, a pixel whose labelled neighbours disagree becomes a line).
Parity with scikit-image itself is UNPINNED (no copy of it to run here); the unit tests check the algebraic properties
(touching discs are split by a one-pixel line, labels never leave the mask, isolated blobs are untouched).
The particle metrology of the reference class (Feret diameters, ellipse fits, filters; :10-262, :307-654) is outside the
scope of this package.
"""
from __future__ import annotations

import heapq

import numpy as np
from scipy import ndimage


def threshold_otsu(image: np.ndarray, nbins: int = 256) -> float:
    """skimage.filters.threshold_otsu: histogram over `nbins` bins (integer images: one bin per value), returns the bin
    centre maximising the between-class variance."""
    image = np.asarray(image)
    if image.min() == image.max():
        return float(image.min())
    if np.issubdtype(image.dtype, np.integer):
        lo, hi = int(image.min()), int(image.max())
        hist = np.bincount(image.ravel() - lo, minlength=hi - lo + 1).astype(np.float64)
        centers = np.arange(lo, hi + 1, dtype=np.float64)
    else:
        hist, edges = np.histogram(image.ravel(), bins=nbins)
        hist = hist.astype(np.float64)
        centers = (edges[:-1] + edges[1:]) / 2.0
    w1 = np.cumsum(hist)
    w2 = np.cumsum(hist[::-1])[::-1]
    m1 = np.cumsum(hist * centers) / np.maximum(w1, 1e-300)
    m2 = (np.cumsum((hist * centers)[::-1]) / np.maximum(w2[::-1], 1e-300))[::-1]
    var12 = w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2
    return float(centers[int(np.argmax(var12))])


def peak_local_max(image: np.ndarray, min_distance: int = 1) -> np.ndarray:
    """skimage.feature.peak_local_max(image, min_distance) with its defaults: candidates are the pixels equal to the
    maximum over a (2*min_distance+1)^2 window and above the image minimum, not closer than min_distance to the border;
    peaks are then visited from the highest down and every peak within min_distance (Chebyshev) of an accepted one is
    dropped.  Returns (row, col) coordinates, highest peak first."""
    image = np.asarray(image, dtype=np.float64)
    size = 2 * min_distance + 1
    mx = ndimage.maximum_filter(image, size=size, mode="nearest")
    cand = (image == mx) & (image > image.min())
    if min_distance > 0:
        cand[:min_distance, :] = False
        cand[-min_distance:, :] = False
        cand[:, :min_distance] = False
        cand[:, -min_distance:] = False
    coords = np.argwhere(cand)
    if coords.shape[0] == 0:
        return coords
    order = np.argsort(-image[cand], kind="stable")
    coords = coords[order]
    taken = np.zeros(image.shape, dtype=bool)
    keep = []
    for r, c in coords:
        if taken[r, c]:
            continue
        keep.append((r, c))
        taken[max(r - min_distance, 0):r + min_distance + 1, max(c - min_distance, 0):c + min_distance + 1] = True
    return np.asarray(keep, dtype=np.int64).reshape(-1, 2)


_N8 = ((-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1))


def watershed(image: np.ndarray, markers: np.ndarray, mask: np.ndarray, watershed_line: bool = True) -> np.ndarray:
    """skimage.segmentation.watershed(image, markers, connectivity=ones((3,3)), mask=mask, watershed_line=...)."""
    h, w = image.shape
    img = np.pad(np.asarray(image, dtype=np.float64), 1, constant_values=np.inf)
    msk = np.pad(np.asarray(mask, dtype=bool), 1, constant_values=False)
    out = np.pad(np.asarray(markers, dtype=np.int64), 1, constant_values=0)
    out[~msk] = 0
    W = w + 2
    imgf, mskf, outf = img.ravel(), msk.ravel(), out.ravel()
    offs = [dr * W + dc for dr, dc in _N8]
    wsl = int(outf.max()) + 1
    heap = []
    age = 0
    for idx in np.flatnonzero(outf):
        heap.append((imgf[idx], 0, int(idx), int(idx)))
    heapq.heapify(heap)
    while heap:
        _, _, idx, src = heapq.heappop(heap)
        if watershed_line:
            if outf[idx] and idx != src:
                continue                        # already labelled from another neighbour
            first, clash = 0, False
            for o in offs:
                j = idx + o
                if mskf[j]:
                    lab = outf[j]
                    if lab == wsl:
                        continue
                    if not first:
                        first = lab
                    elif lab and lab != first:
                        clash = True
                        break
            if clash:
                outf[idx] = wsl
                continue
            outf[idx] = outf[src]
        for o in offs:
            j = idx + o
            if not mskf[j] or outf[j]:
                continue
            age += 1
            if not watershed_line:
                outf[j] = outf[idx]
            heapq.heappush(heap, (imgf[j], age, j, src))
    res = out[1:-1, 1:-1]
    res[res == wsl] = 0
    return res


class Measure:
    """Only the static segmentation entry point of the reference class is mirrored (Measurements.py:263-305)."""

    @staticmethod
    def segment(image, threshold=-1.0, applyWatershed=True, min_distance=9, darkBackground=False):
        img = np.asarray(image).copy()
        if threshold < 0:
            threshold = threshold_otsu(img)
        mask = img > threshold if darkBackground else img < threshold
        if not applyWatershed or np.min(mask) == np.max(mask):
            return np.asarray(mask * 255, dtype="uint8")
        distance = ndimage.distance_transform_edt(mask)
        distance = ndimage.gaussian_filter(distance, sigma=1)
        local_max = peak_local_max(distance, min_distance=min_distance)
        local_maxi = np.zeros(img.shape, dtype="uint8")
        if local_max.shape[0]:
            local_maxi[tuple(local_max.T)] = 1
        markers = ndimage.label(local_maxi)[0]
        labels = watershed(-distance, markers, mask=mask, watershed_line=applyWatershed)
        return np.asarray((labels > 0) * 255, dtype="uint8")


def threshold_li(image: np.ndarray, tolerance: float = None) -> float:
    """skimage.filters.threshold_li: Li's iterative minimum cross-entropy threshold (Li & Tam 1998), initial guess = mean."""
    image = np.asarray(image, dtype=np.float64)
    image = image[np.isfinite(image)]
    if image.size == 0 or image.min() == image.max():
        return float(image.min()) if image.size else 0.0
    shift = image.min()
    image = image - shift
    tolerance = tolerance or np.min(np.diff(np.unique(image))) / 2
    t_next = image.mean()
    t_curr = -2 * tolerance
    while abs(t_next - t_curr) > tolerance:
        t_curr = t_next
        fg = image > t_curr
        mean_fore, mean_back = image[fg].mean(), image[~fg].mean()
        if mean_back == 0:
            break
        t_next = (mean_back - mean_fore) / (np.log(mean_back) - np.log(mean_fore))
    return float(t_next + shift)


def contours_filtered_by_mean_intensity(mask_u8: np.ndarray, gray: np.ndarray, min_value: float = 0.0, max_value: float = -1.0):
    """Measure(mask, applyWatershed=False, excludeEdges=False, grayscaleImage=gray).calculateMeanIntensities() followed by
    filterResults('meanIntensity', minValue, maxValue) (Measurements.py:158-191, 321-342, 606-612): OpenCV contours
    (RETR_TREE), tiny ones (< 5 points and perimeter < 8) dropped, mean grey value over the pixels inside or on each
    contour, contours outside [min_value, max_value] removed.  Returns the surviving contours."""
    import cv2
    contours, _ = cv2.findContours(np.ascontiguousarray(mask_u8, dtype=np.uint8), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    keep = []
    for c in contours:
        if len(c) < 5:
            pts = c[:, 0, :].astype(np.float64)
            if np.sqrt(((np.roll(pts, -1, axis=0) - pts) ** 2).sum(1)).sum() < 8:
                continue
        x0, y0, w, h = cv2.boundingRect(c)
        sub = np.zeros((h, w), dtype=np.uint8)
        cv2.drawContours(sub, [c - np.array([[x0, y0]])], 0, 1, cv2.FILLED)      # filled polygon incl. its boundary
        vals = gray[y0:y0 + h, x0:x0 + w][sub > 0]
        total = float(vals.sum())
        mean = total / vals.size if total > 0 else 0.0
        if min_value == 0 and max_value < min_value:
            keep.append(c)
        elif not (mean < min_value or (mean > max_value and max_value >= min_value)):
            keep.append(c)
    return keep
