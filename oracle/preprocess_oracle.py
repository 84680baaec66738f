"""CPU oracle of the pre-processing step in front of the path (TEST INFRASTRUCTURE ONLY).

Reference: `get_image_tensor`, src/full_model/generate_reports_for_images.py:129-147 —
    cv2.imread(IMREAD_UNCHANGED) -> A.LongestMaxSize(512, INTER_AREA) -> A.PadIfNeeded(512, 512, BORDER_CONSTANT)
    -> A.Normalize(mean=0.471, std=0.302) -> ToTensorV2 -> [1, 1, 512, 512] fp32
albumentations==1.1.0 (requirements.txt:1) is not installed here, so its three transforms are restated from the pinned
version's published algorithm (functional.py: `longest_max_size` -> `py3round(dim * scale)` + `cv2.resize`;
`PadIfNeeded.update_params` centre padding; `normalize`: (img - mean*255) * reciprocal(std*255) in float32); the
resize itself is the REAL cv2.resize(..., INTER_AREA) the reference reaches (cv2 is in the image).

`resize_area_restated` additionally restates OpenCV's INTER_AREA arithmetic for 8-bit single-channel images
(modules/imgproc/src/resize.cpp: computeResizeAreaTab + resizeArea_ for fractional scales, ResizeAreaFast for integer
scales); tests pin it against cv2.resize bit for bit — it documents exactly what the CUDA kernel has to reproduce.
"""
import math

import numpy as np

IMAGE_INPUT_SIZE = 512  # generate_reports_for_images.py:26
MEAN, STD = 0.471, 0.302  # :29-30


def target_size(h: int, w: int, max_size: int = IMAGE_INPUT_SIZE):
    """albumentations 1.1.0 functional.longest_max_size / _func_max_size: scale = max_size / max(h, w);
    new dims = py3round(dim * scale) (Python round: half to even); unchanged if scale == 1."""
    scale = max_size / float(max(h, w))
    if scale == 1.0:
        return h, w
    return int(round(h * scale)), int(round(w * scale))


def pad_offsets(rows: int, cols: int, size: int = IMAGE_INPUT_SIZE):
    """PadIfNeeded.update_params (position = center): top = int((min - rows) / 2.0), bottom = the rest."""
    top = int((size - rows) / 2.0) if rows < size else 0
    left = int((size - cols) / 2.0) if cols < size else 0
    return top, left


def norm_constants():
    """albumentations functional.normalize: mean * 255 and reciprocal(std * 255), all in float32."""
    mean = np.array(MEAN, dtype=np.float32) * np.float32(255.0)
    std = np.array(STD, dtype=np.float32) * np.float32(255.0)
    return np.float32(mean), np.float32(np.reciprocal(std, dtype=np.float32))


def preprocess_reference(image: np.ndarray) -> np.ndarray:
    """uint8 [H, W] -> fp32 [1, 512, 512] exactly as the reference's transform pipeline computes it (real cv2.resize)."""
    import cv2

    assert image.dtype == np.uint8 and image.ndim == 2
    h, w = image.shape
    nh, nw = target_size(h, w)
    if (nh, nw) != (h, w):
        image = cv2.resize(image, dsize=(nw, nh), interpolation=cv2.INTER_AREA)
    top, left = pad_offsets(nh, nw)
    padded = np.zeros((IMAGE_INPUT_SIZE, IMAGE_INPUT_SIZE), dtype=np.uint8)  # BORDER_CONSTANT, value 0
    padded[top:top + nh, left:left + nw] = image
    mean, denom = norm_constants()
    out = padded.astype(np.float32)
    out -= mean
    out *= denom
    return out[None]


def area_tab(ssize: int, dsize: int):
    """OpenCV computeResizeAreaTab: list of (dst index, src index, float32 weight), in table order."""
    scale = 1.0 / (float(dsize) / float(ssize))  # resize(): inv_scale = dsize / ssize; scale = 1 / inv_scale (doubles)
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab, scale


def resize_area_restated(image: np.ndarray, nh: int, nw: int) -> np.ndarray:
    """cv2.resize(image, (nw, nh), INTER_AREA) for uint8 [H, W], down-scaling only, restated."""
    h, w = image.shape
    xtab, sx = area_tab(w, nw)
    ytab, sy = area_tab(h, nh)
    if abs(sx - round(sx)) < np.finfo(np.float64).eps and abs(sy - round(sy)) < np.finfo(np.float64).eps:
        ix, iy = int(round(sx)), int(round(sy))  # ResizeAreaFast: integer block sums
        s = image[:nh * iy, :nw * ix].astype(np.int64).reshape(nh, iy, nw, ix).sum(axis=(1, 3))
        if ix == 2 and iy == 2:
            return ((s + 2) >> 2).astype(np.uint8)  # the SIMD 2x2 path rounds half up
        return np.rint(s.astype(np.float32) * np.float32(1.0 / (ix * iy))).astype(np.uint8)  # saturate_cast: half to even
    # general path: horizontal pass per source row (fp32, table order), then vertical accumulation (fp32, table order)
    di = np.array([t[0] for t in xtab]); si = np.array([t[1] for t in xtab]); al = np.array([t[2] for t in xtab], dtype=np.float32)
    out = np.zeros((nh, nw), dtype=np.uint8)
    rows_of = {}
    for (dy, syi, beta) in ytab:
        rows_of.setdefault(dy, []).append((syi, beta))
    src = image.astype(np.float32)
    for dy, lst in rows_of.items():
        acc = None
        for (syi, beta) in lst:
            prod = src[syi, si] * al  # fp32 products
            buf = np.zeros(nw, dtype=np.float32)
            np.add.at(buf, di, prod)  # unbuffered, sequential: buf[di[k]] += prod[k] in table order, like the C loop
            term = (beta * buf).astype(np.float32)
            acc = term if acc is None else (acc + term).astype(np.float32)
        out[dy] = np.rint(acc).astype(np.uint8)
    return out
