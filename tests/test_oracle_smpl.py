"""Known-answer tests for the (unpinned) SMPL restatement: invariants lifted from the smplx algorithm."""
import math

import numpy as np
import torch

from oracle import smpl_oracle as S
from tokenhmr_b200 import synth
from tokenhmr_b200.config import tiny_config


def _smpl():
    return synth.make_smpl(tiny_config(num_verts=300), seed=3)


def test_rodrigues_known_answers():
    R = S.batch_rodrigues(torch.zeros(1, 3))
    torch.testing.assert_close(R[0], torch.eye(3), atol=1e-6, rtol=0)
    h = math.pi / 2
    R = S.batch_rodrigues(torch.tensor([[0., 0., h], [h, 0., 0.], [0., h, 0.]]))
    torch.testing.assert_close(R[0], torch.tensor([[0., -1., 0.], [1., 0., 0.], [0., 0., 1.]]), atol=1e-6, rtol=0)
    torch.testing.assert_close(R[1], torch.tensor([[1., 0., 0.], [0., 0., -1.], [0., 1., 0.]]), atol=1e-6, rtol=0)
    torch.testing.assert_close(R[2], torch.tensor([[0., 0., 1.], [0., 1., 0.], [-1., 0., 0.]]), atol=1e-6, rtol=0)
    r = torch.randn(50, 3)
    R = S.batch_rodrigues(r)
    torch.testing.assert_close(R @ R.transpose(1, 2), torch.eye(3).expand(50, 3, 3), atol=1e-5, rtol=0)
    torch.testing.assert_close(torch.linalg.det(R), torch.ones(50), atol=1e-5, rtol=0)


def test_identity_pose_gives_shaped_template():
    m = _smpl()
    betas = torch.randn(4, 10)
    R = torch.eye(3).expand(4, 24, 3, 3)
    v, j = S.lbs(betas, R, m["v_template"], m["shapedirs"], m["posedirs"], m["J_regressor"], m["parents"],
                 m["lbs_weights"], pose2rot=False)
    want = m["v_template"] + torch.einsum("bl,vkl->bvk", betas, m["shapedirs"])
    torch.testing.assert_close(v, want, atol=1e-5, rtol=0)
    torch.testing.assert_close(j, torch.einsum("bvk,jv->bjk", want, m["J_regressor"]), atol=1e-5, rtol=0)


def test_global_rotation_rotates_about_root():
    m = _smpl()
    betas = torch.zeros(1, 10)
    aa = torch.zeros(1, 24, 3)
    aa[0, 0] = torch.tensor([0.3, -0.5, 0.8])
    Rg = S.batch_rodrigues(aa[0, :1])[0]
    v, j = S.lbs(betas, aa.view(1, -1), m["v_template"], m["shapedirs"], m["posedirs"], m["J_regressor"], m["parents"],
                 m["lbs_weights"], pose2rot=True)
    J0 = m["J_regressor"] @ m["v_template"]
    # pose_feature ignores the root joint: pure rigid rotation about the root joint
    want = (m["v_template"] - J0[0]) @ Rg.T + J0[0]
    torch.testing.assert_close(v[0], want, atol=1e-5, rtol=0)
    torch.testing.assert_close(j[0], (J0 - J0[0]) @ Rg.T + J0[0], atol=1e-5, rtol=0)


def test_wrapper_joint_layout():
    cfg = tiny_config(num_verts=300)
    m = synth.make_smpl(cfg)
    R = S.batch_rodrigues(0.2 * torch.randn(2 * 24, 3)).view(2, 24, 3, 3)
    betas = torch.randn(2, 10)
    v, j = S.smpl_forward(m, R[:, :1], R[:, 1:], betas)
    assert v.shape == (2, 300, 3) and j.shape == (2, 44, 3)
    _, j24 = S.lbs(betas, R, m["v_template"], m["shapedirs"], m["posedirs"], m["J_regressor"], m["parents"],
                   m["lbs_weights"], pose2rot=False)
    torch.testing.assert_close(j[:, 8], j24[:, 0])            # OpenPose MidHip <- SMPL pelvis (joint_map[8] == 0)
    torch.testing.assert_close(j[:, 0], v[:, m["extra_vertex_ids"][0]])   # nose vertex (joint_map[0] == 24)
    torch.testing.assert_close(j[:, 25:], torch.einsum("bvk,jv->bjk", v, m["joint_regressor_extra"]), atol=1e-6, rtol=0)


def test_f64_fixture_is_reproduced(golden_dir):
    from tokenhmr_b200.config import release_config
    g = np.load(golden_dir / "smpl_lbs_f64.npz")
    m = synth.make_smpl(release_config(), 3)
    aa, betas = torch.from_numpy(g["aa"]), torch.from_numpy(g["betas"])
    R = S.batch_rodrigues(aa.view(-1, 3)).view(8, 24, 3, 3)
    v, j = S.smpl_forward(m, R[:, :1], R[:, 1:], betas)
    torch.testing.assert_close(v, torch.from_numpy(g["verts"]), atol=2e-6, rtol=0)
    torch.testing.assert_close(j, torch.from_numpy(g["joints"]), atol=2e-6, rtol=0)


def test_rodrigues_agrees_with_the_references_own_converters(golden_dir):
    """smplx is absent, but the reference carries two axis-angle -> matrix routines of its own (geometry.aa_to_rotmat,
    rotation_utils.axis_angle_to_matrix): the restated batch_rodrigues must agree with what they produced."""
    g = np.load(golden_dir / "rodrigues_ref.npz")
    R = S.batch_rodrigues(torch.from_numpy(g["aa"]))
    np.testing.assert_allclose(R.numpy(), g["R_aa_to_rotmat"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(R.numpy(), g["R_axis_angle_to_matrix"], rtol=0, atol=2e-6)
