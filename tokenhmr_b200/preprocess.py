"""GPU pre-processing of detector boxes into batch['img'].

Mirror of `ViTDetDataset` (tokenhmr/lib/datasets/vitdet_dataset.py:17-88) + the DataLoader collate that demo.py:69-72
puts in front of `model(batch)`: one call turns a BGR frame and its person boxes into the dict the reference loop
reads ('img', 'personid', 'box_center', 'box_size', 'img_size').  The crop, BGR->RGB flip and normalisation run in
libtokenhmr_b200.so (`thmr_preprocess_boxes`); Python only allocates tensors.  There is no CPU path.

    pre = ViTDetPreprocessor(image_size=256, bbox_shape=(192, 256))
    batch = pre(img_cv2, boxes)             # img_cv2: (H,W,3) uint8 BGR numpy array or CUDA tensor; boxes: (N,4) xyxy
    out = model(batch)
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

DEFAULT_MEAN = (0.485, 0.456, 0.406)      # model_config MODEL.IMAGE_MEAN (vitdet_dataset.py:14-15)
DEFAULT_STD = (0.229, 0.224, 0.225)


def _cfg_struct(image_size: int, bbox_shape: Optional[Sequence[int]], mean, std) -> _lib.PreprocCfg:
    c = _lib.PreprocCfg()
    c.image_size = int(image_size)
    c.bbox_w, c.bbox_h = (int(bbox_shape[0]), int(bbox_shape[1])) if bbox_shape else (0, 0)
    for i in range(3):
        c.mean[i], c.std[i] = float(mean[i]), float(std[i])
    return c


def plan_boxes(boxes: np.ndarray, image_size: int = 256, bbox_shape: Optional[Sequence[int]] = (192, 256)) -> Dict:
    """Host half only (no GPU): 'box_center', 'box_size', blur sigma and the inverse affine map of every box."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    n = boxes.shape[0]
    cfg = _cfg_struct(image_size, bbox_shape, DEFAULT_MEAN, DEFAULT_STD)
    center = np.empty((n, 2), np.float32)
    size = np.empty(n, np.float32)
    sigma = np.empty(n, np.float32)
    inv = np.empty((n, 6), np.float64)
    check(lib().thmr_preprocess_plan(boxes.ctypes.data, n, ctypes.byref(cfg), center.ctypes.data, size.ctypes.data,
                                     sigma.ctypes.data, inv.ctypes.data))
    return {"box_center": center, "box_size": size, "sigma": sigma, "inv_affine": inv.reshape(n, 2, 3)}


class ViTDetPreprocessor:
    def __init__(self, cfg=None, image_size: int = 256, bbox_shape: Optional[Sequence[int]] = (192, 256),
                 mean=DEFAULT_MEAN, std=DEFAULT_STD, device: str | torch.device = "cuda:0"):
        if cfg is not None:                  # the reference's CfgNode (vitdet_dataset.py:31-33,52)
            image_size = cfg.MODEL.IMAGE_SIZE
            mean, std = cfg.MODEL.IMAGE_MEAN, cfg.MODEL.IMAGE_STD
            bbox_shape = cfg.MODEL.get('BBOX_SHAPE', None)
        self.image_size = int(image_size)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.ThmrError("ViTDetPreprocessor needs a CUDA device (there is no CPU fallback)")
        lib()
        self._cfg = _cfg_struct(image_size, bbox_shape, mean, std)
        self._ws: Optional[torch.Tensor] = None

    @torch.no_grad()
    def __call__(self, img_cv2, boxes, return_patch: bool = False) -> Dict:
        dev = self.device
        if isinstance(img_cv2, np.ndarray):
            if img_cv2.dtype != np.uint8 or img_cv2.ndim != 3 or img_cv2.shape[2] != 3:
                raise _lib.ThmrError(f"img_cv2 must be (H,W,3) uint8 BGR, got {img_cv2.dtype} {img_cv2.shape}")
            img = torch.from_numpy(np.ascontiguousarray(img_cv2)).to(dev, non_blocking=True)
        else:
            img = img_cv2.to(dev)
            if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3 or img.stride(2) != 1 or img.stride(1) != 3:
                raise _lib.ThmrError("img tensor must be (H,W,3) uint8 with interleaved BGR pixels")
        H, W = int(img.shape[0]), int(img.shape[1])
        boxes_np = np.ascontiguousarray(boxes.detach().cpu().numpy() if torch.is_tensor(boxes) else boxes,
                                        dtype=np.float32).reshape(-1, 4)
        n = boxes_np.shape[0]
        S = self.image_size
        with torch.cuda.device(dev):
            need = lib().thmr_preprocess_workspace_bytes(H, W, n)
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            out = torch.empty(n, 3, S, S, dtype=torch.float32, device=dev)
            patch = torch.zeros(n, S, S, 3, dtype=torch.uint8, device=dev) if return_patch else None
            center = np.empty((n, 2), np.float32)
            size = np.empty(n, np.float32)
            sigma = np.empty(n, np.float32)
            check(lib().thmr_preprocess_boxes(img.data_ptr(), H, W, img.stride(0), boxes_np.ctypes.data, n,
                                              ctypes.byref(self._cfg), out.data_ptr(),
                                              patch.data_ptr() if patch is not None else None, center.ctypes.data,
                                              size.ctypes.data, sigma.ctypes.data, self._ws.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream))
        batch = {
            "img": out,
            "personid": torch.arange(n, dtype=torch.int64),
            "box_center": torch.from_numpy(center).to(dev),
            "box_size": torch.from_numpy(size).to(dev),
            "img_size": torch.tensor([[float(W), float(H)]] * n, dtype=torch.float64, device=dev),
        }
        if return_patch:
            batch["_patch_bgr_u8"], batch["_sigma"] = patch, sigma
        return batch
