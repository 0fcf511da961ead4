"""CPU restatement of the evaluation path and cam_crop_to_full (f3).  TEST INFRASTRUCTURE:
imported only by tests/ and oracle/make_golden.py; never by the product.

Pinned: tests/test_oracle_pinned.py compares these functions with the LIVE reference `lib/utils/pose_utils.py`
(importable in the build container through oracle/ref_import.py) and with tests/golden/evaluator.npz, which
oracle/make_golden.py wrote from the reference's own Evaluator.  cam_crop_to_full lives in renderer.py, whose module
imports pyrender/trimesh (absent): pinned through the same file with those two modules stubbed.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch


def compute_similarity_transform(S1: torch.Tensor, S2: torch.Tensor) -> torch.Tensor:
    """tokenhmr/lib/utils/pose_utils.py:61-114 (batched orthogonal Procrustes, S1 -> S2)."""
    S1 = S1.to(torch.float32).permute(0, 2, 1)
    S2 = S2.to(torch.float32).permute(0, 2, 1)
    mu1 = S1.mean(dim=2, keepdim=True)
    mu2 = S2.mean(dim=2, keepdim=True)
    X1, X2 = S1 - mu1, S2 - mu2
    var1 = (X1 ** 2).sum(dim=(1, 2))
    K = X1 @ X2.permute(0, 2, 1)
    U, s, Vh = torch.linalg.svd(K)              # torch.svd returns V; linalg.svd returns V^H
    V = Vh.permute(0, 2, 1)
    Z = torch.eye(3).unsqueeze(0).repeat(S1.shape[0], 1, 1)
    Z[:, -1, -1] *= torch.sign(torch.linalg.det(U @ Vh))
    R = V @ Z @ U.permute(0, 2, 1)
    trace = (R @ K).diagonal(offset=0, dim1=-1, dim2=-2).sum(dim=-1)
    scale = (trace / var1)[:, None, None]
    t = mu2 - scale * (R @ mu1)
    return (scale * (R @ S1) + t).permute(0, 2, 1)


def reconstruction_error(S1, S2):
    """pose_utils.py:116-127."""
    S1_hat = compute_similarity_transform(S1, S2)
    return torch.sqrt(((S1_hat - S2) ** 2).sum(dim=-1)).mean(dim=-1)


def eval_pose(pred_joints, gt_joints):
    """pose_utils.py:129-143: (MPJPE, PA-MPJPE) in mm."""
    mpjpe = torch.sqrt(((pred_joints - gt_joints) ** 2).sum(dim=-1)).mean(dim=-1)
    return 1000 * mpjpe, 1000 * reconstruction_error(pred_joints, gt_joints)


def evaluate_batch(output: Dict, batch: Dict, keypoint_list: List[int], pelvis_ind: int,
                   J_regressor_24_SMPL: Optional[torch.Tensor] = None, dataset: str = ''):
    """Evaluator.__call__ (pose_utils.py:201-275) for one batch, num_samples = 1; returns (mpjpe, re, pve) in mm.
    Unlike the reference it does not modify `output` in place."""
    if 'EMDB' in dataset:
        gt_vertices = batch['vertices']
        gt_kp = torch.matmul(J_regressor_24_SMPL, gt_vertices)
        gt_pelvis = (gt_kp[:, [1], :] + gt_kp[:, [2], :]) / 2.0
        gt_kp, gt_vertices = gt_kp - gt_pelvis, gt_vertices - gt_pelvis
        pred_vertices = output['pred_vertices']
        pred_kp = torch.matmul(J_regressor_24_SMPL, pred_vertices)
        pred_pelvis = (pred_kp[:, [1], :] + pred_kp[:, [2], :]) / 2.0
        pred_kp, pred_vertices = pred_kp - pred_pelvis, pred_vertices - pred_pelvis
    else:
        pred_kp = output['pred_keypoints_3d'].clone()
        gt_kp = batch['keypoints_3d'][:, :, :-1].clone()
        pred_pelvis = pred_kp[:, [pelvis_ind]]
        gt_pelvis = gt_kp[:, [pelvis_ind]]
        pred_kp, gt_kp = pred_kp - pred_pelvis, gt_kp - gt_pelvis
        pred_vertices = output['pred_vertices'] - pred_pelvis
        gt_vertices = batch['vertices'] - gt_pelvis
    mpjpe, re = eval_pose(pred_kp[:, keypoint_list], gt_kp[:, keypoint_list])
    pve = torch.sqrt(((pred_vertices - gt_vertices) ** 2).sum(dim=-1)).mean(dim=-1) * 1000.
    return mpjpe, re, pve


def cam_crop_to_full(cam_bbox, box_center, box_size, img_size, focal_length=5000.):
    """tokenhmr/lib/utils/renderer.py:13-23."""
    img_w, img_h = img_size[:, 0], img_size[:, 1]
    cx, cy, b = box_center[:, 0], box_center[:, 1], box_size
    w_2, h_2 = img_w / 2., img_h / 2.
    bs = b * cam_bbox[:, 0] + 1e-9
    tz = 2 * focal_length / bs
    tx = (2 * (cx - w_2) / bs) + cam_bbox[:, 1]
    ty = (2 * (cy - h_2) / bs) + cam_bbox[:, 2]
    return torch.stack([tx, ty, tz], dim=-1)


def synthetic_eval_batch(B: int, V: int = 6890, J: int = 44, seed: int = 0, noise: float = 0.03):
    """Seeded prediction/ground-truth pair shaped like the evaluation inputs: gt = random body-sized point sets,
    pred = a random similarity transform of gt plus noise (so Procrustes has something to undo)."""
    g = torch.Generator().manual_seed(seed)
    gt_v = torch.randn(B, V, 3, generator=g) * torch.tensor([0.3, 0.9, 0.15])
    gt_kp = torch.randn(B, J, 3, generator=g) * torch.tensor([0.3, 0.9, 0.15])
    ax = torch.randn(B, 3, generator=g) * 0.4
    ang = ax.norm(dim=-1, keepdim=True)
    k = ax / ang
    Kx = torch.zeros(B, 3, 3)
    Kx[:, 0, 1], Kx[:, 0, 2], Kx[:, 1, 0] = -k[:, 2], k[:, 1], k[:, 2]
    Kx[:, 1, 2], Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 0], -k[:, 1], k[:, 0]
    R = torch.eye(3) + torch.sin(ang)[..., None] * Kx + (1 - torch.cos(ang))[..., None] * (Kx @ Kx)
    s = 1 + 0.1 * torch.randn(B, 1, 1, generator=g)
    t = 0.2 * torch.randn(B, 1, 3, generator=g)
    pred_v = s * (gt_v @ R.transpose(1, 2)) + t + noise * torch.randn(B, V, 3, generator=g)
    pred_kp = s * (gt_kp @ R.transpose(1, 2)) + t + noise * torch.randn(B, J, 3, generator=g)
    conf = torch.ones(B, J, 1)
    return {"pred_vertices": pred_v, "pred_keypoints_3d": pred_kp}, \
           {"vertices": gt_v, "keypoints_3d": torch.cat([gt_kp, conf], -1), "imgname": [f"img{i}" for i in range(B)]}
