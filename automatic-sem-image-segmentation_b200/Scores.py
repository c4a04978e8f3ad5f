"""Segmentation scores as the reference computes them (Archive/Other Scripts/Calculate_Scores.py): whole-image IoU
(:69-70), instance IoU over OpenCV contours (:73-104), confusion rates / Youden index (:107-136) and the threshold sweep
0.0 ... 1.0 in steps of 0.1 with its best-of selection (:221-272).  HOST code (numpy / OpenCV); used to report the
`Datasets/` IoU of a trained UNet the reference's way.

Reference quirk kept on request only: calculateIoU accumulates threshold t/10 at list index t-1 (so t = 0 lands in the last
slot) but reports index/10 as the best threshold -- the reported threshold is 0.1 too low, the best IoU itself is right.
`sweep_iou(..., reference_indexing=True)` reproduces that report; the default reports the true threshold.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence

import numpy as np

from . import HelperFunctions
from .Measurements import Measure


def calculateWholeImageIoU(image1, image2) -> float:
    """Calculate_Scores.py:69-70."""
    return float(np.sum(np.logical_and(image1, image2)) / np.sum(np.logical_or(image1, image2)))


def polygon_area(x, y) -> float:
    """Shoelace formula (Calculate_Scores.py:139-151)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    x_, y_ = x - x.mean(), y - y.mean()
    correction = x_[-1] * y_[0] - y_[-1] * x_[0]
    main_area = np.dot(x_[:-1], y_[1:]) - np.dot(y_[:-1], x_[1:])
    return float(0.5 * np.abs(main_area + correction))


def calculateInstanceIoU(image1, image2, minArea: float = 0) -> float:
    """Mean over the contours of image1 (area > minArea) of the best IoU with any bounding-box-overlapping contour of
    image2 (Calculate_Scores.py:73-104)."""
    import cv2
    image1, image2 = np.ascontiguousarray(image1, dtype=np.uint8), np.ascontiguousarray(image2, dtype=np.uint8)
    c1, _ = cv2.findContours(image1, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    c2, _ = cv2.findContours(image2, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    boxes2 = [(c[:, 0, 0].min(), c[:, 0, 0].max(), c[:, 0, 1].min(), c[:, 0, 1].max()) for c in c2]
    drawn2: Dict[int, np.ndarray] = {}
    blank = np.zeros(image1.shape, dtype=np.uint8)
    scores = []
    for i, c in enumerate(c1):
        x1, y1 = c[:, 0, 0], c[:, 0, 1]
        if polygon_area(x1, y1) <= minArea:
            continue
        img1 = cv2.drawContours(blank.copy(), c1, i, 1, cv2.FILLED)
        best = 0.0
        for j, (xa, xb, ya, yb) in enumerate(boxes2):
            if xa > x1.max() or xb < x1.min() or ya > y1.max() or yb < y1.min():
                continue
            if j not in drawn2:
                drawn2[j] = cv2.drawContours(blank.copy(), c2, j, 1, cv2.FILLED)
            best = max(best, calculateWholeImageIoU(img1, drawn2[j]))
        scores.append(best)
    return float(np.mean(scores)) if scores else 0.0


def ROC(predicted, groundTruth):
    """(TPR, TNR, FPR, FNR) of two {0,1} images (Calculate_Scores.py:107-136), vectorised."""
    p, g = np.asarray(predicted).astype(np.int64), np.asarray(groundTruth).astype(np.int64)
    FP, FN = float(np.sum(p > g)), float(np.sum(p < g))
    TN, TP = float(np.sum((p == g) & (p == 0))), float(np.sum((p == g) & (p == 1)))
    TPR = TP / (TP + FN) if TP + FN > 0 else 0
    TNR = TN / (TN + FP) if TN + FP > 0 else 0
    FPR = FP / (TN + FP) if TN + FP > 0 else 0
    FNR = FN / (TP + FN) if TP + FN > 0 else 0
    return TPR, TNR, FPR, FNR


def segment(image, threshold, doWatershed: bool = True) -> np.ndarray:
    """Calculate_Scores.py:35-66: threshold (Otsu if < 0), optional watershed split, {0,1} uint8."""
    img = np.asarray(image)
    if img.dtype == bool:
        mask = img
    else:
        if threshold < 0:
            from .Measurements import threshold_otsu
            threshold = threshold_otsu(img)
        mask = img > threshold
    if np.min(mask) == np.max(mask) or not doWatershed:
        return np.asarray(mask > 0, dtype="uint8")
    from scipy import ndimage
    seg = Measure.segment(mask.astype(np.uint8), threshold=0.5, applyWatershed=True, min_distance=9, darkBackground=True) > 0
    seg = ndimage.binary_fill_holes(seg, structure=np.ones((3, 3)))
    return np.asarray(seg, dtype="uint8")


def sweep_iou(predictions: Sequence[np.ndarray], ground_truths: Sequence[np.ndarray], watershed: bool = False,
              instance: bool = False, reference_indexing: bool = False) -> Dict[str, object]:
    """calculateIoU (Calculate_Scores.py:221-272) over in-memory images: predictions in [0,1] (or 0..255), ground truths
    {0,1}; thresholds 0.0, 0.1, ..., 1.0; averages over the images; best average and its threshold."""
    n = float(len(predictions))
    whole = [0.0] * 11
    inst_all = [0.0] * 11
    inst_f = [0.0] * 11
    for pred, gt in zip(predictions, ground_truths):
        image = np.asarray(pred, dtype=np.float32)
        if image.max() > 1.0:
            image = image / 255.0
        gt = np.asarray(gt)
        gt = (gt // max(int(gt.max()), 1)).astype(np.uint8)
        for t in range(11):
            seg = segment(image, threshold=t / 10.0, doWatershed=watershed)
            seg = HelperFunctions.eight_to_four_connected(seg)
            slot = (t - 1) % 11 if reference_indexing else t
            whole[slot] += calculateWholeImageIoU(seg, gt) / n
            if instance:
                inst_all[slot] += calculateInstanceIoU(seg, gt, 0) / n
                inst_f[slot] += calculateInstanceIoU(seg, gt, 9) / n

    def best(v):
        i = int(np.argmax(v)) if max(v) > 0 else 0
        return float(v[i]), i / 10.0
    out = {"whole_image": whole, "best_whole_image": best(whole)}
    if instance:
        out.update({"instances_all": inst_all, "best_instances_all": best(inst_all), "instances_filtered": inst_f,
                    "best_instances_filtered": best(inst_f)})
    return out
