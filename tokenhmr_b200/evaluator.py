"""GPU-resident `Evaluator`, mirroring tokenhmr/lib/utils/pose_utils.py:145-275.

Same constructor, `__call__(output, batch)`, `log()`, `get_metrics_dict()`, `get_imgnames()` as the reference class,
so `eval.py:130-152` runs unchanged.  Differences, all deliberate:
  * every metric is computed by `thmr_eval_pose` on the GPU and stays there; the host sees numbers only in `log()` /
    `get_metrics_dict()` (the reference does `torch.svd` + `.cpu().numpy()` every batch);
  * per-sample results are kept as a list of device tensors instead of `np.zeros((dataset_length,))` per metric
    (`eval.py:134` passes `dataset_length=int(1e8)`, i.e. 800 MB per metric in the reference);
  * `output['pred_keypoints_3d']` is not modified (the reference's `-=` on a `.detach()` view pelvis-centres the
    caller's tensor in place, pose_utils.py:222,237 — nothing downstream relies on it).
There is no CPU path: tensors must be CUDA tensors.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops

_MM_METRICS = ("mode_mpjpe", "mode_re", "mode_pve")


class Evaluator:
    def __init__(self, dataset_length: int, keypoint_list: List, pelvis_ind: int,
                 metrics: List = ['mode_mpjpe', 'mode_re', 'model_pve'], J_regressor_24_SMPL=None, dataset=''):
        self.dataset_length = dataset_length
        self.keypoint_list = list(keypoint_list)
        self.pelvis_ind = pelvis_ind
        self.metrics = list(metrics)
        self.J_regressor_24_SMPL = J_regressor_24_SMPL
        self.dataset = dataset
        self._chunks: Dict[str, List[torch.Tensor]] = {m: [] for m in self.metrics}
        self._kp_list_dev: Optional[torch.Tensor] = None
        self.counter = 0
        self.imgnames: List = []

    # ------------------------------------------------------------------ reference-compatible accessors
    def _all(self, metric: str) -> torch.Tensor:
        ch = self._chunks[metric]
        if not ch:
            return torch.zeros(0)
        return torch.cat(ch)[:self.counter]

    def __getattr__(self, name):
        # the reference exposes each metric as an array attribute (setattr in pose_utils.py:168-169)
        chunks = self.__dict__.get("_chunks", {})
        if name in chunks:
            return self._all(name).cpu().numpy()
        raise AttributeError(name)

    def log(self):
        if self.counter == 0:
            print('Evaluation has not started')
            return
        print(f'{self.counter} / {self.dataset_length} samples')
        for metric in self.metrics:
            unit = 'mm' if metric in _MM_METRICS else ''
            print(f'{metric}: {self._all(metric).double().mean().item() if self._chunks[metric] else 0.0} {unit}')
        print('***')

    def get_metrics_dict(self) -> Dict:
        # metrics that were requested but never filled average to 0.0, as the reference's zero-initialised arrays do
        return {m: (self._all(m).double().mean().item() if self._chunks[m] else 0.0) for m in self.metrics}

    def get_imgnames(self):
        return self.imgnames

    # ------------------------------------------------------------------ one batch
    def __call__(self, output: Dict, batch: Dict):
        self.imgnames += list(batch['imgname'])
        pred_vertices = output['pred_vertices']
        dev = pred_vertices.device
        if self._kp_list_dev is None or self._kp_list_dev.device != dev:
            self._kp_list_dev = torch.tensor(self.keypoint_list, dtype=torch.int32, device=dev)
        want_pve = 'mode_pve' in self._chunks
        gt_vertices = batch['vertices'].to(dev, torch.float32)
        if 'EMDB' in self.dataset:
            jreg = self.J_regressor_24_SMPL.to(dev, torch.float32)
            gt_kp = ops.regress_joints(jreg, gt_vertices)
            pred_kp = ops.regress_joints(jreg, pred_vertices.float())
            pelvis = (1, 2)
        else:
            pred_kp = output['pred_keypoints_3d'].detach().float()
            gt_kp = batch['keypoints_3d'].to(dev, torch.float32)   # (B,J,4): the confidence column is skipped by the kernel
            pelvis = (self.pelvis_ind, self.pelvis_ind)
        mpjpe, re, pve = ops.eval_pose(pred_kp, gt_kp, self._kp_list_dev, pelvis,
                                       pred_vertices.float() if want_pve else None,
                                       gt_vertices if want_pve else None)
        batch_size = pred_kp.shape[0]
        if 'mode_mpjpe' in self._chunks:
            self._chunks['mode_mpjpe'].append(mpjpe)
        if 'mode_re' in self._chunks:
            self._chunks['mode_re'].append(re)
        if want_pve:
            self._chunks['mode_pve'].append(pve)
        self.counter += batch_size
        if 'mode_mpjpe' in self._chunks and 'mode_re' in self._chunks:
            return {'mode_mpjpe': mpjpe, 'mode_re': re}
        return None
