"""Host-side logic: config arithmetic, synthetic generators, checkpoint key handling, shard ranges."""
import torch

from tokenhmr_b200 import synth
from tokenhmr_b200.config import release_config, tiny_config
from tokenhmr_b200.dist import shard_range
from tokenhmr_b200.weights import strip_checkpoint


def test_release_config_matches_reference_numbers():
    c = release_config()
    assert (c.grid_h, c.grid_w, c.num_tokens, c.head_dim, c.crop_x0) == (16, 12, 192, 80, 32)
    assert c.upsample_sizes == [125, 90, 55, 21]          # vanilla_pose_vqvae.py:139
    assert c.dec_inner == 512 and c.npose == 144


def test_synthetic_state_dict_shapes_and_parameter_counts():
    c = tiny_config(vit_depth=2)
    sd = synth.make_state_dict(c)
    per_block = sum(v.numel() for k, v in sd.items() if k.startswith("backbone.blocks.0."))
    rest = sum(v.numel() for k, v in sd.items() if k.startswith("backbone.") and ".blocks." not in k)
    assert per_block * 32 + rest == 630_912_000          # SURVEY.md §8 header (ViT-H probe)
    dec = sum(v.numel() for k, v in sd.items() if k.startswith("smpl_head.transformer."))
    assert dec == 39_386_112
    cls = sum(v.numel() for k, v in sd.items() if k.startswith("smpl_head.decpose."))
    assert cls == 10_870_080
    tok = sum(v.numel() for k, v in sd.items() if k.startswith("tokenizer.decoder."))
    assert tok == 6_436_870
    assert sd["tokenizer.quantizer.codebook"].shape == (2048, 256)


def test_synthetic_generation_is_deterministic_and_nested():
    a = synth.make_state_dict(tiny_config(vit_depth=1))
    b = synth.make_state_dict(tiny_config(vit_depth=2))
    for k, v in a.items():
        assert torch.equal(v, b[k]), k


def test_synthetic_smpl_is_well_formed():
    m = synth.make_smpl(tiny_config(num_verts=500))
    assert torch.allclose(m["lbs_weights"].sum(1), torch.ones(500), atol=1e-6)
    assert (m["lbs_weights"] > 0).sum(1).max() <= 4
    assert torch.allclose(m["J_regressor"].sum(1), torch.ones(24), atol=1e-6)
    assert m["parents"][0] == -1 and int(m["extra_vertex_ids"].max()) < 500


def test_strip_checkpoint_prefixes():
    ck = {"backbone.pos_embed": torch.zeros(1), "smpl_head.deccam.bias": torch.zeros(3), "discriminator.x": torch.zeros(1)}
    net = {"decoder.decoder.0.weight": torch.zeros(1), "quantizer.codebook": torch.zeros(1),
           "decoder.body_model.faces": torch.zeros(1), "encoder.x": torch.zeros(1)}
    out = strip_checkpoint(ck, net)
    assert set(out) == {"backbone.pos_embed", "smpl_head.deccam.bias", "tokenizer.decoder.decoder.0.weight",
                        "tokenizer.quantizer.codebook"}


def test_shard_range_covers_batch_contiguously():
    for gb, world in [(512, 8), (10, 4), (3, 8), (64, 1)]:
        spans = [shard_range(gb, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_split_weight_layout_and_precision():
    """Host half of the split-precision contractions (strict mode, tokenizer encoder): w * 2^8 = hi + lo with
    hi = fp16(w * 2^8), lo = fp16(w * 2^8 - hi); packed per tap as [hi | hi | lo]."""
    import torch
    from tokenhmr_b200._lib import ThmrError
    from tokenhmr_b200.weights import STRICT_W_SCALE, split_weight
    g = torch.Generator().manual_seed(0)
    w = torch.randn(8, 3 * 16, generator=g) * 0.3
    w[0, 0], w[1, 5] = 1e-7, 200.0                                   # tiny and large magnitudes
    s = split_weight(w, taps=3)
    assert s.dtype == torch.float16 and tuple(s.shape) == (8, 3 * 3 * 16)
    s = s.float().view(8, 3, 3, 16)                                  # [out, tap, (hi | hi | lo), cin]
    assert torch.equal(s[:, :, 0], s[:, :, 1])
    rec = (s[:, :, 0] + s[:, :, 2]).reshape(8, 48) / STRICT_W_SCALE
    err = (rec - w).abs()
    assert (err <= w.abs() * 2.0 ** -21 + 2.0 ** -33).all(), err.max()
    one = split_weight(w, taps=1).float()
    assert torch.equal(one[:, :48], one[:, 48:96])                   # single tap: [hi(48) | hi(48) | lo(48)]
    import pytest
    with pytest.raises(ThmrError):
        split_weight(torch.full((2, 4), 300.0))


def _vq_screen_constants():
    """kVqScale / kVqScreenRel / kVqScreenAbs as compiled into the library (parsed from csrc/vq.cuh so that the
    emulation below cannot drift from the kernel)."""
    import re
    from pathlib import Path
    src = (Path(__file__).resolve().parent.parent / "tokenhmr_b200" / "csrc" / "vq.cuh").read_text()
    scale = float(re.search(r"kVqScale = ([0-9.]+)f", src).group(1))
    m = re.search(r"kVqScreenRel = ([0-9.]+)f \* ([0-9.]+)f", src)
    rel = float(m.group(1)) * float(m.group(2))
    ab = float(re.search(r"kVqScreenAbs = ([0-9.e+-]+)f", src).group(1))
    return scale, rel, ab


def test_vq_screen_margin_is_sound_numpy_emulation():
    """Emulates pass 1 of the screened arg-min (csrc/vq.cuh, gemm_tcgen05.cuh) in numpy: fp16-rounded operands, fp32
    accumulation, best / second best of e = |c|^2 - 2 x.c, margin tau.  Soundness: every row that pass 1 would NOT hand to the
    exact pass already has the fp64 arg-min.  Also: the error of the single-product e stays inside the bound eps = tau / 2
    the margin is built from, and the share of re-done rows is what DESIGN.md quotes (about a tenth of N(0,1) queries)."""
    import numpy as np
    scale, rel, ab = _vq_screen_constants()
    assert abs(rel - 2.0 ** -8 * 1.05) < 1e-9
    rng = np.random.default_rng(5)
    for xs, cs in ((1.0, 1.0), (6.0, 0.2), (1e-2, 3.0)):
        C = (cs * rng.standard_normal((2048, 256))).astype(np.float32)
        X = (xs * rng.standard_normal((3000, 256))).astype(np.float32)
        # near-ties on purpose: midpoints of two codes (+ a little noise)
        a, b = rng.integers(0, 2048, 300), rng.integers(0, 2048, 300)
        X[:300] = 0.5 * (C[a] + C[b]) + (1e-4 * cs * rng.standard_normal((300, 256))).astype(np.float32)
        c2 = (C.astype(np.float64) ** 2).sum(1)
        x2 = (X.astype(np.float64) ** 2).sum(1)
        exact = c2[None, :] - 2.0 * (X.astype(np.float64) @ C.astype(np.float64).T)
        Xh = (X * scale).astype(np.float16).astype(np.float32)
        Ch = (C * scale).astype(np.float16).astype(np.float32)
        acc = Xh @ Ch.T                                             # fp32 accumulate of exact fp16 x fp16 products
        e = (np.float32(-2.0 / (scale * scale)) * acc + c2.astype(np.float32)[None, :]).astype(np.float32)
        # the kernel packs the column's position in its 32-column chunk into the low mantissa byte of e (gemm_tcgen05.cuh)
        keys = ((e.view(np.uint32) & np.uint32(0xFFFFFF00)) | (np.arange(2048, dtype=np.uint32) & np.uint32(31))[None, :]).view(np.float32)
        order = np.argsort(keys, axis=1, kind="stable")
        best, second = keys[np.arange(len(e)), order[:, 0]], keys[np.arange(len(e)), order[:, 1]]
        cmax2 = np.float32(c2.max())
        tau = (np.float32(rel) * np.sqrt(x2.astype(np.float32) * cmax2) + np.float32(ab) * (x2.astype(np.float32) + cmax2)
               + np.float32(1e-12))
        redo = ~((second - best) > tau)
        final = ~redo
        assert np.array_equal(order[final, 0], exact.argmin(1)[final])
        assert (np.abs(keys - exact).max(1) <= 0.5 * tau).all()     # the bound the margin is built from
        assert redo[:300].mean() > 0.9                               # planted near-ties are re-done
        assert redo[300:].mean() < 0.3, redo[300:].mean()
