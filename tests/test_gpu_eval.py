"""GPU parity of the evaluation row: thmr_eval_pose / thmr_regress_joints / thmr_cam_crop_to_full and
the Evaluator host class against the CPU oracle and the golden written by the live reference's Evaluator."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KL = list(range(25, 39))


@pytest.fixture(autouse=True)
def _flags(cuda_dev, built_lib):
    yield
    assert built_lib.thmr_check_device_flags() == 0, built_lib.thmr_last_error()


def _dev(d, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}


def test_eval_pose_matches_reference_golden(cuda_dev, golden_dir):
    from tokenhmr_b200.evaluator import Evaluator
    g = np.load(golden_dir / "evaluator.npz")
    t = lambda k: torch.from_numpy(g[k]).to(cuda_dev)
    out = {"pred_vertices": t("pred_vertices"), "pred_keypoints_3d": t("pred_keypoints_3d")}
    batch = {"vertices": t("gt_vertices"), "keypoints_3d": t("gt_keypoints_3d"), "imgname": [f"i{i}" for i in range(6)]}
    kp_before = out["pred_keypoints_3d"].clone()
    ev = Evaluator(int(1e8), list(g["keypoint_list"]), 39, metrics=['mode_re', 'mode_mpjpe', 'mode_pve'], dataset='3DPW-TEST')
    ret = ev(out, batch)
    assert torch.equal(out["pred_keypoints_3d"], kp_before)            # no in-place side effect
    np.testing.assert_allclose(ev.mode_mpjpe, g["mpjpe"], rtol=1e-5)
    np.testing.assert_allclose(ev.mode_re, g["re"], rtol=1e-4)         # tolerance: fp32 torch.svd vs fp64 Jacobi
    np.testing.assert_allclose(ev.mode_pve, g["pve"], rtol=1e-5)
    assert set(ret) == {"mode_mpjpe", "mode_re"} and ret["mode_re"].is_cuda
    d = ev.get_metrics_dict()
    assert abs(d["mode_mpjpe"] - g["mpjpe"].mean()) < 1e-2 and ev.counter == 6 and len(ev.get_imgnames()) == 6
    # EMDB branch: joints regressed from the vertices, pelvis = mid-hip of joints 1,2
    ev2 = Evaluator(int(1e8), list(range(24)), 39, metrics=['mode_re', 'mode_mpjpe', 'mode_pve'],
                    J_regressor_24_SMPL=t("jreg"), dataset='EMDB')
    ev2(out, batch)
    np.testing.assert_allclose(ev2.mode_mpjpe, g["emdb_mpjpe"], rtol=2e-5)
    np.testing.assert_allclose(ev2.mode_re, g["emdb_re"], rtol=1e-4)
    np.testing.assert_allclose(ev2.mode_pve, g["emdb_pve"], rtol=2e-5)


@pytest.mark.parametrize("B,V", [(1, 6890), (64, 6890), (7, 33)])
def test_eval_pose_vs_oracle(cuda_dev, B, V):
    from oracle import eval_oracle as E
    from tokenhmr_b200 import ops
    out, batch = E.synthetic_eval_batch(B, V=V, seed=B)
    m, r, p = E.evaluate_batch(out, batch, KL, 39)
    kl = torch.tensor(KL, dtype=torch.int32, device=cuda_dev)
    gm, gr, gp = ops.eval_pose(out["pred_keypoints_3d"].to(cuda_dev), batch["keypoints_3d"].to(cuda_dev), kl, (39, 39),
                               out["pred_vertices"].to(cuda_dev), batch["vertices"].to(cuda_dev))
    np.testing.assert_allclose(gm.cpu().numpy(), m.numpy(), rtol=1e-5)
    np.testing.assert_allclose(gr.cpu().numpy(), r.numpy(), rtol=1e-4)
    np.testing.assert_allclose(gp.cpu().numpy(), p.numpy(), rtol=2e-5)
    # without vertices: no PVE, same joint metrics; gt without the confidence column
    gm2, gr2, gp2 = ops.eval_pose(out["pred_keypoints_3d"].to(cuda_dev), batch["keypoints_3d"][..., :3].contiguous().to(cuda_dev),
                                  kl, (39, 39))
    assert gp2 is None and torch.equal(gm, gm2) and torch.equal(gr, gr2)


def test_procrustes_properties(cuda_dev):
    """Size-independent properties: a similarity-transformed copy has zero aligned error; a mirrored copy does not
    (the determinant fix forbids reflections); planar point sets (zero smallest singular value) stay finite."""
    from tokenhmr_b200 import ops
    g = torch.Generator().manual_seed(4)
    S2 = torch.randn(32, 14, 3, generator=g)
    c, s = np.cos(1.1), np.sin(1.1)
    R = torch.tensor([[c, 0., s], [0., 1., 0.], [-s, 0., c]], dtype=torch.float32)
    S1 = 0.6 * S2 @ R.T + torch.tensor([1., 2., -3.])
    kl = torch.arange(14, dtype=torch.int32, device=cuda_dev)
    _, re, _ = ops.eval_pose(S1.to(cuda_dev), S2.to(cuda_dev), kl, (0, 0))
    assert re.max().item() < 1e-2                                        # mm
    _, re_m, _ = ops.eval_pose((S2 * torch.tensor([1., 1., -1.])).to(cuda_dev), S2.to(cuda_dev), kl, (0, 0))
    assert re_m.min().item() > 50.0
    flat = S2.clone(); flat[..., 2] = 0
    _, re_f, _ = ops.eval_pose((flat @ R.T).contiguous().to(cuda_dev), flat.to(cuda_dev), kl, (0, 0))
    assert torch.isfinite(re_f).all() and re_f.max().item() < 1e-2


def test_regress_joints_and_cam_crop(cuda_dev, golden_dir):
    from tokenhmr_b200 import ops
    g = np.load(golden_dir / "evaluator.npz")
    t = lambda k: torch.from_numpy(g[k]).to(cuda_dev)
    j = ops.regress_joints(t("jreg"), t("pred_vertices"))
    ref = torch.matmul(torch.from_numpy(g["jreg"]), torch.from_numpy(g["pred_vertices"]))
    torch.testing.assert_close(j.cpu(), ref, rtol=1e-5, atol=1e-6)
    full = ops.cam_crop_to_full(t("cam"), t("center"), t("size"), t("img_size"))
    np.testing.assert_allclose(full.cpu().numpy(), g["full_cam"], rtol=2e-6, atol=1e-5)


def test_eval_rejects_bad_arguments(cuda_dev):
    from tokenhmr_b200 import ops
    from tokenhmr_b200._lib import ThmrError
    kp = torch.zeros(2, 44, 3, device=cuda_dev)
    with pytest.raises(ThmrError):
        ops.eval_pose(kp, kp, torch.arange(70, dtype=torch.int32, device=cuda_dev) % 44, (39, 39))   # K > 64
    with pytest.raises(ThmrError):
        ops.eval_pose(kp, kp, torch.arange(14, dtype=torch.int32, device=cuda_dev), (44, 44))        # pelvis out of range
    with pytest.raises(ThmrError):
        ops.eval_pose(kp.cpu(), kp, torch.arange(14, dtype=torch.int32, device=cuda_dev), (0, 0))     # no CPU path
