"""CPU restatement of the input pre-processing that feeds TokenHMR.forward.  TEST
INFRASTRUCTURE: imported only by tests/, oracle/make_golden.py and never by the product.

What it restates (numpy, integer / float64 arithmetic, no cv2 / scipy / skimage calls):
  * ViTDetDataset.__init__/__getitem__          tokenhmr/lib/datasets/vitdet_dataset.py:17-88
  * expand_to_aspect_ratio                      tokenhmr/lib/datasets/utils.py:14-33
  * gen_trans_from_patch_cv                     tokenhmr/lib/datasets/utils.py:81-129
  * generate_image_patch_cv2 (no flip, rot 0)   tokenhmr/lib/datasets/utils.py:317-361
  * convert_cvimg_to_tensor                     tokenhmr/lib/datasets/utils.py:364-376
and the third-party routines those lines call, restated from their published algorithms:
  * cv2.getAffineTransform  (OpenCV imgproc/imgwarp.cpp: 6x6 system, hal::LU64f with partial pivoting)
  * cv2.warpAffine INTER_LINEAR / BORDER_CONSTANT (imgwarp.cpp WarpAffineInvoker + remapBilinear): inverse map in
    double, source coordinates in 10-bit fixed point rounded to 1/32 pixel; 8-bit images interpolate with the
    integer table 32*a*b (sum 2^15, rounding (v + 2^14) >> 15), float64 images with the float32 table
  * skimage.filters.gaussian(channel_axis=2, preserve_range=True) == scipy.ndimage.gaussian_filter over the two
    image axes, mode 'nearest', truncate 4.0, float64 (skimage is absent from this image; scipy is what it calls)

Pinned: tests/test_oracle_pinned.py compares every function with the real cv2 (4.13 here) / scipy calls and the
whole item with the LIVE reference ViTDetDataset (imported file by file through oracle/ref_import.py with a
skimage.filters shim that forwards to scipy.ndimage), and with tests/golden/preproc.npz written from it.
Scalar types follow NumPy >= 2 promotion (the version in this image): float32 arrays stay float32 against Python
scalars; the normalisation `(patch - mean) / std` with float64 mean/std runs in float64 and is stored as float32.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

DEFAULT_MEAN = (0.485, 0.456, 0.406)
DEFAULT_STD = (0.229, 0.224, 0.225)


# --------------------------------------------------------------------------------------------- box geometry
def box_center_scale(boxes: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """vitdet_dataset.py:35-38: float32 throughout."""
    boxes = boxes.astype(np.float32)
    center = (boxes[:, 2:4] + boxes[:, 0:2]) / np.float32(2.0)
    scale = (boxes[:, 2:4] - boxes[:, 0:2]) / np.float32(200.0)
    return center, scale


def expand_to_aspect_ratio(wh: np.ndarray, target: Optional[Sequence[int]]) -> np.ndarray:
    """utils.py:14-33 on a float32 (w, h) pair; the Python scalars h_t / w_t are weak (float32 comparison)."""
    if target is None:
        return wh
    w, h = np.float32(wh[0]), np.float32(wh[1])
    w_t, h_t = target
    if np.float32(h / w) < np.float32(h_t / w_t):
        h_new = max(np.float32(np.float32(w * np.float32(h_t)) / np.float32(w_t)), h)
        w_new = w
    else:
        h_new = h
        w_new = max(np.float32(np.float32(h * np.float32(w_t)) / np.float32(h_t)), w)
    return np.array([w_new, h_new], dtype=np.float32)


def bbox_size(scale: np.ndarray, bbox_shape: Optional[Sequence[int]]) -> np.float32:
    """vitdet_dataset.py:51-53: expand_to_aspect_ratio(scale*200, BBOX_SHAPE).max()."""
    return np.float32(expand_to_aspect_ratio(scale * np.float32(200), bbox_shape).max())


def blur_sigma(size: np.float32, patch: int) -> Optional[np.float32]:
    """vitdet_dataset.py:61-66: sigma of the anti-alias blur, or None when the box is small enough."""
    f = np.float32(np.float32(np.float32(size * np.float32(1.0)) / np.float32(patch)) / np.float32(2.0))
    if f > np.float32(1.1):
        return np.float32(np.float32(f - np.float32(1)) / np.float32(2))
    return None


# --------------------------------------------------------------------------------------------- affine map
def get_affine_transform(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """cv2.getAffineTransform: solve the 6x6 system with OpenCV's LU (partial pivoting), float64."""
    a = np.zeros((6, 6), dtype=np.float64)
    b = np.zeros(6, dtype=np.float64)
    for i in range(3):
        r0, r1 = 2 * i, 2 * i + 1
        a[r0, 0] = a[r1, 3] = float(src[i, 0])
        a[r0, 1] = a[r1, 4] = float(src[i, 1])
        a[r0, 2] = a[r1, 5] = 1.0
        b[r0], b[r1] = float(dst[i, 0]), float(dst[i, 1])
    m = 6
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(a[j, i]) > abs(a[k, i]):
                k = j
        if k != i:
            a[[i, k], i:] = a[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / a[i, i]
        for j in range(i + 1, m):
            alpha = a[j, i] * d
            for c in range(i + 1, m):
                a[j, c] += alpha * a[i, c]
            b[j] += alpha * b[i]
    for i in range(m - 1, -1, -1):
        s = b[i]
        for c in range(i + 1, m):
            s -= a[i, c] * b[c]
        b[i] = s / a[i, i]
    return b.reshape(2, 3)


def gen_trans(c_x: np.float32, c_y: np.float32, size: np.float32, patch: int) -> np.ndarray:
    """utils.py:81-129 with scale = 1, rot = 0 (the inference call, vitdet_dataset.py:69-73)."""
    half = np.float32(size * np.float32(0.5))
    src = np.zeros((3, 2), dtype=np.float32)
    src[0] = [c_x, c_y]
    src[1] = [c_x, np.float64(c_y) + np.float64(half)]     # float64 centre + float32 direction, stored as float32
    src[2] = [np.float64(c_x) + np.float64(half), c_y]
    h = np.float32(patch * 0.5)
    dst = np.array([[h, h], [h, h + h], [h + h, h]], dtype=np.float32)
    return get_affine_transform(src, dst)


def invert_affine(M: np.ndarray) -> np.ndarray:
    """cv2.warpAffine without WARP_INVERSE_MAP (imgwarp.cpp invertAffineTransform order of operations)."""
    M = M.astype(np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    iM = np.empty((2, 3), dtype=np.float64)
    iM[0, 0], iM[0, 1] = A11, M[0, 1] * (-D)
    iM[1, 0], iM[1, 1] = M[1, 0] * (-D), A22
    iM[0, 2] = -iM[0, 0] * M[0, 2] - iM[0, 1] * M[1, 2]
    iM[1, 2] = -iM[1, 0] * M[0, 2] - iM[1, 1] * M[1, 2]
    return iM


def _source_coords(iM: np.ndarray, W: int, H: int):
    """Fixed-point source coordinates of every destination pixel: integer part and 5-bit fractions."""
    x = np.arange(W, dtype=np.float64)
    y = np.arange(H, dtype=np.float64)
    ad = np.rint(iM[0, 0] * x * 1024).astype(np.int64)
    bd = np.rint(iM[1, 0] * x * 1024).astype(np.int64)
    X0 = np.rint((iM[0, 1] * y + iM[0, 2]) * 1024).astype(np.int64) + 16
    Y0 = np.rint((iM[1, 1] * y + iM[1, 2]) * 1024).astype(np.int64) + 16
    X = (X0[:, None] + ad[None, :]) >> 5
    Y = (Y0[:, None] + bd[None, :]) >> 5
    return X >> 5, Y >> 5, X & 31, Y & 31


def _gather(img: np.ndarray, yy: np.ndarray, xx: np.ndarray, dtype):
    Hs, Ws = img.shape[:2]
    ok = (yy >= 0) & (yy < Hs) & (xx >= 0) & (xx < Ws)
    v = img[np.clip(yy, 0, Hs - 1), np.clip(xx, 0, Ws - 1)].astype(dtype)
    return v * ok[..., None].astype(dtype)                  # BORDER_CONSTANT, borderValue 0


def warp_affine_u8(img: np.ndarray, M: np.ndarray, W: int, H: int) -> np.ndarray:
    """cv2.warpAffine(img uint8, M, (W,H), INTER_LINEAR, BORDER_CONSTANT, 0): bit exact."""
    sx, sy, fx, fy = _source_coords(invert_affine(M), W, H)
    w00 = ((32 - fx) * (32 - fy) * 32)[..., None]
    w01 = (fx * (32 - fy) * 32)[..., None]
    w10 = ((32 - fx) * fy * 32)[..., None]
    w11 = (fx * fy * 32)[..., None]
    acc = (_gather(img, sy, sx, np.int64) * w00 + _gather(img, sy, sx + 1, np.int64) * w01 +
           _gather(img, sy + 1, sx, np.int64) * w10 + _gather(img, sy + 1, sx + 1, np.int64) * w11)
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


def warp_affine_f64(img: np.ndarray, M: np.ndarray, W: int, H: int) -> np.ndarray:
    """The same call on a float64 image (the blurred one): float32 weight table, float64 sum, left to right."""
    sx, sy, fx, fy = _source_coords(invert_affine(M), W, H)
    tx1 = (fx.astype(np.float32) * np.float32(1.0 / 32))
    ty1 = (fy.astype(np.float32) * np.float32(1.0 / 32))
    tx0, ty0 = np.float32(1.0) - tx1, np.float32(1.0) - ty1
    w = [(ty0 * tx0).astype(np.float64)[..., None], (ty0 * tx1).astype(np.float64)[..., None],
         (ty1 * tx0).astype(np.float64)[..., None], (ty1 * tx1).astype(np.float64)[..., None]]
    return (_gather(img, sy, sx, np.float64) * w[0] + _gather(img, sy, sx + 1, np.float64) * w[1] +
            _gather(img, sy + 1, sx, np.float64) * w[2] + _gather(img, sy + 1, sx + 1, np.float64) * w[3])


# --------------------------------------------------------------------------------------------- blur
def gaussian_kernel1d(sigma: float) -> np.ndarray:
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, int(4.0 * sigma + 0.5)), float64."""
    radius = int(4.0 * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return phi / phi.sum()


def gaussian_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    """skimage.filters.gaussian(img, sigma, channel_axis=2, preserve_range=True): rows (axis 0) then columns
    (axis 1), edge pixels replicated, float64.  Tap order follows scipy's correlate1d for a symmetric kernel:
    centre tap first, then the mirrored pairs from the outside in ((left + right) * weight)."""
    w = gaussian_kernel1d(sigma)
    r = len(w) // 2
    out = img.astype(np.float64)
    for axis in (0, 1):
        n = out.shape[axis]
        idx = np.arange(n)
        acc = np.take(out, idx, axis=axis) * w[r]
        for k in range(-r, 0):
            lo = np.take(out, np.clip(idx + k, 0, n - 1), axis=axis)
            hi = np.take(out, np.clip(idx - k, 0, n - 1), axis=axis)
            acc = acc + (lo + hi) * w[k + r]
        out = acc
    return out


# --------------------------------------------------------------------------------------------- the item
def vitdet_item(img_bgr: np.ndarray, box: np.ndarray, image_size: int = 256,
                bbox_shape: Optional[Sequence[int]] = (192, 256), mean=DEFAULT_MEAN, std=DEFAULT_STD) -> Dict:
    """ViTDetDataset.__getitem__ for one box (x0, y0, x1, y1) of one BGR uint8 image."""
    center, scale = box_center_scale(np.asarray(box, dtype=np.float32)[None])
    size = bbox_size(scale[0], bbox_shape)
    sigma = blur_sigma(size, image_size)
    M = gen_trans(center[0, 0], center[0, 1], size, image_size)
    if sigma is None:
        patch = warp_affine_u8(img_bgr, M, image_size, image_size)
    else:
        patch = warp_affine_f64(gaussian_blur(img_bgr, float(sigma)), M, image_size, image_size)
    chw = np.transpose(patch[:, :, ::-1], (2, 0, 1)).astype(np.float32)
    m = 255.0 * np.array(mean, dtype=np.float64)
    s = 255.0 * np.array(std, dtype=np.float64)
    for c in range(3):
        chw[c] = ((chw[c].astype(np.float64) - m[c]) / s[c]).astype(np.float32)
    return {"img": chw, "patch": patch, "trans": M, "box_center": center[0].copy(), "box_size": size,
            "img_size": 1.0 * np.array([img_bgr.shape[1], img_bgr.shape[0]]), "sigma": sigma}
