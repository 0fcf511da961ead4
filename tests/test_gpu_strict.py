"""Strict mode (TokenHMREngine(strict=True): every contraction in split fp16, fp32-grade) against the fp32 reference:
the north-star bar "pose tokens exact; SMPL vertices/joints within 1e-4 rel fp32" end to end, and the bs=64
(BASELINE configs[1]) parity of both modes against a golden written by the LIVE reference modules."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

CONT_KEYS = ("pred_cam", "pred_cam_t", "pred_keypoints_3d", "pred_vertices", "pred_keypoints_2d")
# strict: measured ~1e-6..1e-5 (fp32 summation order, exp / erf implementations); the contract is 1e-4
TOL_STRICT = 1e-4
# pose tokens := argmax of cls_logits_softmax.  A token may differ from the fp32 reference only where the reference's
# own top-2 probabilities are closer than fp32 evaluation noise (the 2048-way softmax of x20-scaled logits amplifies a
# 1e-6 logit difference to ~1e-5 in probability): gap below this bound, and never more than 1 position in 160.
TIE_GAP = 2e-4


@pytest.fixture(autouse=True)
def _flags(cuda_dev, built_lib):
    yield
    assert built_lib.thmr_check_device_flags() == 0, built_lib.thmr_last_error()


def _check_tokens(probs_gpu: torch.Tensor, ref_argmax, ref_top1, ref_top2, max_frac: float, gap_bound: float):
    tok = probs_gpu.argmax(-1).cpu().numpy()
    bad = tok != np.asarray(ref_argmax)
    gap = (np.asarray(ref_top1) - np.asarray(ref_top2))[bad]
    assert bad.mean() <= max_frac, (float(bad.mean()), gap)
    assert (gap < gap_bound).all(), gap
    return int(bad.sum()), (float(gap.max()) if gap.size else 0.0)


def test_strict_tiny_vs_fp32_oracle(cuda_dev):
    from oracle import tokenhmr_oracle as O
    from tokenhmr_b200 import synth
    from tokenhmr_b200.config import tiny_config
    from tokenhmr_b200.engine import TokenHMREngine
    cfg = tiny_config(vit_depth=2)
    sd, smpl = synth.make_state_dict(cfg), synth.make_smpl(cfg)
    model = TokenHMREngine(cfg, sd, smpl, device=cuda_dev, use_cuda_graph=False, strict=True)
    for B in (1, 3):
        img = synth.make_images(B, cfg, seed=20 + B)
        out = model({"img": img}, return_taps=True)
        with torch.no_grad():
            ref = O.forward(sd, smpl, img, cfg, emulate_fp16=False, return_intermediates=True)
        errs = {k: rel_err(out[k], ref[k]) for k in CONT_KEYS + ("_vit_tokens", "_token_out", "_pred_body_pose_6d",
                                                                 "cls_logits_softmax")}
        print("strict tiny", B, {k: f"{v:.1e}" for k, v in errs.items()})
        for k, v in errs.items():
            assert v < (1e-3 if k == "cls_logits_softmax" else TOL_STRICT), (k, v)
        top2 = ref["cls_logits_softmax"].topk(2, dim=-1).values
        _check_tokens(out["cls_logits_softmax"], ref["cls_logits_softmax"].argmax(-1).numpy(), top2[..., 0].numpy(),
                      top2[..., 1].numpy(), 1 / 160, TIE_GAP)
    # CUDA graph replay of the strict forward is bit-identical to the eager run
    eager = model({"img": img})["pred_vertices"].clone()
    model.use_cuda_graph = True
    for _ in range(2):
        assert torch.equal(model({"img": img})["pred_vertices"], eager)


@pytest.fixture(scope="module")
def release_models(cuda_dev):
    from tokenhmr_b200 import synth
    from tokenhmr_b200.config import release_config
    from tokenhmr_b200.engine import TokenHMREngine
    cfg = release_config()
    sd, smpl = synth.make_state_dict(cfg, 1234), synth.make_smpl(cfg, 3)
    fast = TokenHMREngine(cfg, sd, smpl, device=cuda_dev)
    strict = TokenHMREngine(cfg, sd, smpl, device=cuda_dev, strict=True)
    return cfg, fast, strict


def test_strict_release_vs_reference_golden(release_models, golden_dir):
    """Full ViT-H/16 depth-32 forward, B=2, against the LIVE reference's fp32 outputs: 1e-4 / exact tokens."""
    from tokenhmr_b200 import synth
    cfg, _, strict = release_models
    g = np.load(golden_dir / "forward_release_d32.npz")
    out = strict({"img": synth.make_images(2, cfg, 0)}, return_taps=True)
    t = lambda k: torch.from_numpy(g[k])
    errs = {k: rel_err(out[k], t(k)) for k in CONT_KEYS}
    errs["vit_tokens"] = rel_err(out["_vit_tokens"][:, ::8], t("vit_tokens_sub"))
    errs["probs"] = rel_err(out["cls_logits_softmax"][:, ::16], t("cls_probs_sub"))
    nbad, gap = _check_tokens(out["cls_logits_softmax"], g["cls_argmax"], g["cls_maxprob"], g["cls_second_prob"],
                              1 / 320, TIE_GAP)
    print("strict release B=2", {k: f"{v:.1e}" for k, v in errs.items()}, "token mismatches", nbad, "max gap", gap)
    for k, v in errs.items():
        assert v < (1e-3 if k == "probs" else TOL_STRICT), (k, v)
    for k in ("global_orient", "body_pose", "betas"):
        assert rel_err(out["pred_smpl_params"][k], t(k)) < TOL_STRICT


def test_bs64_release_vs_reference_golden(release_models, golden_dir):
    """BASELINE configs[1] itself (bs = 64, the benchmarked shape) against the LIVE reference: the default mode within
    its fp16-operand tolerance, strict mode within 1e-4 with exact tokens."""
    from tokenhmr_b200 import synth
    cfg, fast, strict = release_models
    g = np.load(golden_dir / "forward_release_d32_bs64.npz")
    stride = int(g["meta"][6])
    img = synth.make_images(64, cfg, int(g["meta"][2]))
    t = lambda k: torch.from_numpy(g[k])
    for name, model, tol, frac, gapb in (("fast", fast, 2e-3, 0.02, 0.08), ("strict", strict, TOL_STRICT, 1 / 2000, TIE_GAP)):
        out = model({"img": img})
        errs = {k: rel_err(out[k], t(k)) for k in ("pred_cam", "pred_cam_t", "pred_keypoints_3d", "pred_keypoints_2d")}
        errs["pred_vertices"] = rel_err(out["pred_vertices"][:, ::stride], t("pred_vertices_sub"))
        errs["betas"] = rel_err(out["pred_smpl_params"]["betas"], t("betas"))
        errs["global_orient"] = rel_err(out["pred_smpl_params"]["global_orient"], t("global_orient"))
        nbad, gap = _check_tokens(out["cls_logits_softmax"], g["cls_argmax"], g["cls_maxprob"], g["cls_second_prob"],
                                  frac, gapb)
        print(f"bs64 {name}", {k: f"{v:.1e}" for k, v in errs.items()}, "token mismatches", nbad, "of", 64 * 160,
              "max gap", gap)
        for k, v in errs.items():
            assert v < tol, (name, k, v)
        assert torch.isfinite(out["pred_vertices"]).all()
        torch.testing.assert_close(out["cls_logits_softmax"].sum(-1), torch.ones(64, 160, device=out["pred_cam"].device),
                                   atol=1e-4, rtol=0)
        del out


def test_strict_rejects_out_of_range_weights(cuda_dev):
    from tokenhmr_b200 import synth
    from tokenhmr_b200._lib import ThmrError
    from tokenhmr_b200.config import tiny_config
    from tokenhmr_b200.engine import TokenHMREngine
    cfg = tiny_config(vit_depth=1)
    sd, smpl = synth.make_state_dict(cfg), synth.make_smpl(cfg)
    sd = dict(sd)
    sd["backbone.blocks.0.mlp.fc1.weight"] = sd["backbone.blocks.0.mlp.fc1.weight"].clone()
    sd["backbone.blocks.0.mlp.fc1.weight"][0, 0] = 300.0
    with pytest.raises(ThmrError):
        TokenHMREngine(cfg, sd, smpl, device=cuda_dev, strict=True)
