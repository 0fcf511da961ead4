"""CPU oracle for the SMPL stage.  TEST INFRASTRUCTURE ONLY (see tokenhmr_oracle.py header).

PARITY UNPINNED: the arithmetic lives in the third-party package smplx==0.1.28 (pinned by the reference's
requirements.txt:3 and tokenhmr/setup.py:13), which is neither vendored under /root/reference nor
installable offline.  The functions below restate the published algorithm of smplx/lbs.py,
smplx/body_models.py (SMPLLayer.forward) and smplx/vertex_joint_selector.py, and are anchored on the
reference's own call sites:
    tokenhmr/lib/models/smpl_wrapper.py:10,19-24,27-41   (SMPL(smplx.SMPLLayer), joint_map, extra regressor)
    tokenhmr/lib/models/tokenhmr.py:173-176              (self.smpl(..., pose2rot=False))
The one piece the reference tree can vouch for is cross-checked against it: batch_rodrigues agrees to 2e-6 with the
reference's own axis-angle converters (geometry.aa_to_rotmat, rotation_utils.axis_angle_to_matrix) through
tests/golden/rodrigues_ref.npz (oracle/make_golden.py --only rodrigues).
Known-answer checks lifted from the algorithm itself (tests/test_oracle_smpl.py): identity pose gives
v_template + shapedirs.beta; a global rotation rotates the rest mesh about the root joint; batch_rodrigues
of a zero vector is I; 90-degree rotations about the axes give the textbook matrices.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

Tensor = torch.Tensor


def batch_rodrigues(rot_vecs: Tensor, eps: float = 1e-8) -> Tensor:
    """smplx.lbs.batch_rodrigues: (N,3) axis-angle -> (N,3,3).
    angle = ||r + 1e-8||, K = skew(r / angle), R = I + sin(angle) K + (1 - cos(angle)) K K."""
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + eps, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.cos(angle).unsqueeze(1)
    sin = torch.sin(angle).unsqueeze(1)
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=rot_vecs.dtype)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(n, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype).unsqueeze(0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def batch_rigid_transform(rot_mats: Tensor, joints: Tensor, parents: Tensor) -> Tuple[Tensor, Tensor]:
    """smplx.lbs.batch_rigid_transform: rot_mats (B,J,3,3), joints (B,J,3) -> posed joints (B,J,3),
    relative transforms A (B,J,4,4) with A[:3,3] = G[:3,3] - G[:3,:3] @ J."""
    B, J = joints.shape[:2]
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] -= joints[:, parents[1:]]
    top = torch.cat([rot_mats, rel], dim=-1)                                  # (B,J,3,4)
    bottom = torch.tensor([0, 0, 0, 1], dtype=rot_mats.dtype).view(1, 1, 1, 4).expand(B, J, 1, 4)
    tm = torch.cat([top, bottom], dim=-2)                                     # (B,J,4,4)
    chain = [tm[:, 0]]
    for i in range(1, J):
        chain.append(torch.matmul(chain[int(parents[i])], tm[:, i]))
    G = torch.stack(chain, dim=1)
    posed = G[:, :, :3, 3]
    jh = torch.cat([joints, torch.zeros(B, J, 1, 1, dtype=rot_mats.dtype)], dim=2)   # (B,J,4,1)
    corr = torch.matmul(G, jh)                                                # (B,J,4,1)
    A = G - torch.cat([torch.zeros(B, J, 4, 3, dtype=rot_mats.dtype), corr], dim=-1)
    return posed, A


def lbs(betas: Tensor, pose: Tensor, v_template: Tensor, shapedirs: Tensor, posedirs: Tensor,
        J_regressor: Tensor, parents: Tensor, lbs_weights: Tensor, pose2rot: bool = True
        ) -> Tuple[Tensor, Tensor]:
    """smplx.lbs.lbs.  betas (B,10); pose (B,J*3) axis-angle if pose2rot else (B,J,3,3);
    returns vertices (B,V,3) and posed joints (B,J,3)."""
    B = betas.shape[0]
    dt = betas.dtype
    v_shaped = v_template.unsqueeze(0) + torch.einsum("bl,mkl->bmk", betas, shapedirs)      # blend_shapes
    Jr = torch.einsum("bik,ji->bjk", v_shaped, J_regressor)                                  # vertices2joints
    ident = torch.eye(3, dtype=dt)
    if pose2rot:
        rot_mats = batch_rodrigues(pose.reshape(-1, 3)).view(B, -1, 3, 3)
    else:
        rot_mats = pose.view(B, -1, 3, 3)
    pose_feature = (rot_mats[:, 1:] - ident).reshape(B, -1)                                  # (B,207)
    v_posed = v_shaped + torch.matmul(pose_feature, posedirs).view(B, -1, 3)
    J_transformed, A = batch_rigid_transform(rot_mats, Jr, parents)
    W = lbs_weights.unsqueeze(0).expand(B, -1, -1)
    nj = J_regressor.shape[0]
    T = torch.matmul(W, A.view(B, nj, 16)).view(B, -1, 4, 4)
    v_h = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=dt)], dim=2)
    verts = torch.matmul(T, v_h.unsqueeze(-1))[:, :, :3, 0]
    return verts, J_transformed


def smpl_forward(smpl: Dict[str, Tensor], global_orient: Tensor, body_pose: Tensor, betas: Tensor,
                 dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """tokenhmr SMPL wrapper forward (smpl_wrapper.py:27-41) on top of smplx.SMPLLayer.forward:
    rotation-matrix input (pose2rot ignored by SMPLLayer), 24 posed joints + 21 selected vertices = 45,
    remapped by joint_map to 25 OpenPose joints, plus 19 regressed extra joints -> (B,44,3)."""
    from tokenhmr_b200.config import SMPL_TO_OPENPOSE
    B = betas.shape[0]
    c = lambda t: t.to(dtype)
    full_pose = torch.cat([global_orient.reshape(B, -1, 3, 3), body_pose.reshape(B, -1, 3, 3)], dim=1)
    verts, joints = lbs(c(betas), c(full_pose), c(smpl["v_template"]), c(smpl["shapedirs"]),
                        c(smpl["posedirs"]), c(smpl["J_regressor"]), smpl["parents"], c(smpl["lbs_weights"]),
                        pose2rot=False)
    joints45 = torch.cat([joints, verts[:, smpl["extra_vertex_ids"]]], dim=1)                # VertexJointSelector
    j = joints45[:, torch.tensor(SMPL_TO_OPENPOSE)]                                          # smpl_wrapper.py:32
    extra = torch.einsum("bik,ji->bjk", verts, c(smpl["joint_regressor_extra"]))             # :38 vertices2joints
    return verts, torch.cat([j, extra], dim=1)
