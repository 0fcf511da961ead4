"""Pins the CPU oracle: (1) against the golden vectors produced by the LIVE reference modules
(oracle/make_golden.py), (2) against the live modules themselves when /root/reference exists."""
import numpy as np
import pytest
import torch

from oracle import ref_import, smpl_oracle
from oracle import tokenhmr_oracle as O
from tokenhmr_b200 import synth
from tokenhmr_b200.config import release_config, tiny_config

W_SEED, SMPL_SEED, IMG_SEED = 1234, 3, 0


def _check_forward(golden, cfg, batch):
    g = np.load(golden)
    assert list(g["meta"][:4]) == [W_SEED, SMPL_SEED, IMG_SEED, batch]
    sd, smpl = synth.make_state_dict(cfg, W_SEED), synth.make_smpl(cfg, SMPL_SEED)
    img = synth.make_images(batch, cfg, IMG_SEED)
    with torch.no_grad():
        out = O.forward(sd, smpl, img, cfg, return_intermediates=True)
    t = lambda k: torch.from_numpy(g[k])
    tol = dict(rtol=0, atol=2e-6)
    torch.testing.assert_close(out["_vit_tokens"][:, ::8], t("vit_tokens_sub"), rtol=0, atol=2e-5)
    torch.testing.assert_close(out["cls_logits_softmax"][:, ::16], t("cls_probs_sub"), rtol=0, atol=1e-5)
    assert np.array_equal(out["cls_logits_softmax"].argmax(-1).numpy().astype(np.int16), g["cls_argmax"])
    for k in ("pred_cam", "focal_length"):
        torch.testing.assert_close(out[k], t(k), **tol)
    torch.testing.assert_close(out["pred_cam_t"], t("pred_cam_t"), rtol=1e-5, atol=1e-5)
    for k in ("global_orient", "body_pose", "betas"):
        torch.testing.assert_close(out["pred_smpl_params"][k], t(k), rtol=0, atol=5e-6)
    for k in ("pred_keypoints_3d", "pred_vertices", "pred_keypoints_2d"):
        torch.testing.assert_close(out[k], t(k), rtol=1e-5, atol=2e-5)


def test_forward_tiny_matches_reference_golden(golden_dir):
    _check_forward(golden_dir / "forward_tiny_d2.npz", tiny_config(vit_depth=2), 2)


def test_forward_release_matches_reference_golden(golden_dir):
    torch.set_num_threads(max(1, torch.get_num_threads()))
    _check_forward(golden_dir / "forward_release_d32.npz", release_config(), 2)


def test_vq_quantize_matches_reference_golden(golden_dir):
    g = np.load(golden_dir / "vq_quantize.npz")
    cb = torch.randn(2048, 256, generator=torch.Generator().manual_seed(1))
    x = torch.randn(4096, 256, generator=torch.Generator().manual_seed(2))
    assert np.array_equal(O.vq_quantize(x, cb).numpy(), g["idx_rand"])
    gg = torch.Generator().manual_seed(7)
    pick = torch.randint(0, 2048, (4096,), generator=gg)
    xn = cb[pick] + 0.05 * torch.randn(4096, 256, generator=gg)
    idx = O.vq_quantize(xn, cb).numpy()
    assert np.array_equal(idx, g["idx_near"]) and np.array_equal(idx, g["pick"])
    # quantize(codebook[i]) == i, dequantize(quantize(.)) returns codebook rows (invariants from the code)
    assert torch.equal(O.vq_quantize(cb, cb), torch.arange(2048))
    assert torch.equal(O.vq_dequantize(torch.from_numpy(idx), cb), cb[torch.from_numpy(idx)])
    torch.testing.assert_close(torch.from_numpy(g["logits"]) @ cb, torch.from_numpy(g["dequant_logits"]), rtol=1e-5, atol=1e-5)


def test_geometry_matches_reference_golden(golden_dir):
    g = np.load(golden_dir / "geometry.npz")
    torch.testing.assert_close(O.rot6d_to_rotmat(torch.from_numpy(g["x6"])), torch.from_numpy(g["rotmat"]), rtol=0, atol=1e-6)
    proj = O.perspective_projection(torch.from_numpy(g["pts"]), torch.from_numpy(g["tr"]), torch.from_numpy(g["fl"]))
    torch.testing.assert_close(proj, torch.from_numpy(g["proj"]), rtol=1e-5, atol=1e-5)
    eye6 = torch.tensor([[1., 0., 0., 0., 1., 0.]])
    torch.testing.assert_close(O.rot6d_to_rotmat(eye6)[0], torch.eye(3))


def test_upsample_index_matches_torch():
    import torch.nn as nn
    for lin, lout in [(160, 125), (125, 90), (90, 55), (55, 21)]:
        x = torch.arange(lin, dtype=torch.float32).view(1, 1, lin)
        want = nn.Upsample(lout)(x)[0, 0].long()
        assert torch.equal(O.upsample_nearest_index(lout, lin), want)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not mounted (GPU box)")
def test_restatement_equals_live_reference_modules():
    """Same seeds, other batch: the functional restatement reproduces the reference modules bit for bit."""
    cfg = tiny_config(vit_depth=2)
    sd, smpl = synth.make_state_dict(cfg, 99), synth.make_smpl(cfg, 5)
    img = synth.make_images(3, cfg, 11)
    ns = ref_import.load_modules()
    ref = ref_import.reference_forward(ns, ref_import.build_backbone(ns, sd, cfg), ref_import.build_head(ns, sd, cfg),
                                       smpl, img, cfg)
    with torch.no_grad():
        out = O.forward(sd, smpl, img, cfg, return_intermediates=True)
    for k in ("_vit_tokens", "cls_logits_softmax", "pred_cam", "pred_keypoints_3d", "pred_vertices", "pred_keypoints_2d"):
        torch.testing.assert_close(out[k], ref[k], rtol=0, atol=1e-6)
    # hard quantiser of the reference
    qz = ns.quantize_cnn.QuantizeEMAReset(64, 32)
    cb = torch.randn(64, 32)
    qz.codebook = cb
    x = torch.randn(500, 32)
    assert torch.equal(qz.quantize(x), O.vq_quantize(x, cb))


def test_fp16_emulation_stays_close_to_fp32():
    cfg = tiny_config(vit_depth=2)
    sd, smpl = synth.make_state_dict(cfg), synth.make_smpl(cfg)
    img = synth.make_images(2, cfg)
    with torch.no_grad():
        a = O.forward(sd, smpl, img, cfg, emulate_fp16=False)
        b = O.forward(sd, smpl, img, cfg, emulate_fp16=True)
    err = ((a["pred_vertices"] - b["pred_vertices"]).abs().max() / a["pred_vertices"].abs().max()).item()
    assert 0 < err < 2e-3
