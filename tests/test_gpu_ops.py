"""GPU parity tests of the stand-alone operators (through the C ABI) against torch / the CPU oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _flags(cuda_dev, built_lib):
    yield
    assert built_lib.thmr_check_device_flags() == 0, built_lib.thmr_last_error()


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 520, 200), (64, 1024, 1024), (1000, 96, 160), (130, 6, 1536),
                                   (1536, 1280, 1280), (2, 1024, 1024), (0 + 1, 32, 64)])
@pytest.mark.parametrize("bn", [0, 256, 128, 64, 32])
def test_gemm_all_epilogues(cuda_dev, M, N, K, bn):
    from tokenhmr_b200 import ops
    torch.manual_seed(M * 7 + N)
    Kp = (K + 7) // 8 * 8
    A = torch.randn(M, Kp, device=cuda_dev).half()[:, :K]
    B = torch.randn(N, Kp, device=cuda_dev).half()[:, :K]
    A, B = A.contiguous() if K == Kp else A, B.contiguous() if K == Kp else B
    if K != Kp:   # ops.linear_f16 wants contiguous operands: keep the padded pitch by zero-extending K
        A = torch.cat([A, torch.zeros(M, Kp - K, device=cuda_dev, dtype=torch.float16)], 1).contiguous()
        B = torch.cat([B, torch.zeros(N, Kp - K, device=cuda_dev, dtype=torch.float16)], 1).contiguous()
    bias = torch.randn(N, device=cuda_dev)
    resid = torch.randn(M, N, device=cuda_dev)
    ref = A.float() @ B.float().t() + bias
    # generic epilogue: two outputs + separate residual + GELU on the fp16 output
    o32, o16 = ops.linear_f16(A, B, bias, resid, "gelu", out32=True, out16=True, block_n=bn)
    assert rel_err(o32, ref + resid) < 1e-5                    # fp32 accumulate: summation-order noise only
    assert rel_err(o16, F.gelu(ref + resid)) < 2e-3            # one fp16 rounding
    # fp16-only output (TMA-store epilogue when M >= 128 and N % 8 == 0)
    _, o16b = ops.linear_f16(A, B, bias, None, "relu", out32=False, out16=True, block_n=bn)
    assert rel_err(o16b, F.relu(ref)) < 2e-3
    # in-place residual add (TMA reduce-add epilogue when eligible)
    x = resid.clone()
    from tokenhmr_b200._lib import check, lib
    check(lib().thmr_gemm_f16(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, A.shape[1], bias.data_ptr(),
                              x.data_ptr(), N, 0, x.data_ptr(), N, None, 0, bn, torch.cuda.current_stream().cuda_stream))
    assert rel_err(x, ref + resid) < 1e-5


@pytest.mark.parametrize("M,N,K", [(12288, 1280, 1280), (12288, 1280, 5120), (384, 1280, 1280), (1024, 520, 320),
                                   (4608, 1280, 5120), (256, 256, 64), (9984, 1280, 704)])
def test_gemm_cta_pair_fp32_epilogues(cuda_dev, M, N, K):
    """fp32 epilogues (TMA reduce-add, TMA store) of the CTA-pair kernel at the ViT's proj / fc2 shapes (240 tiles on 74
    CTA pairs: stream-K, tiles shared by two clusters are reduce-added in a fixed order), with fewer tiles than CTA
    pairs, with a partial last column tile, and with a k-block count (11) that range boundaries snap against."""
    from tokenhmr_b200._lib import check, lib
    torch.manual_seed(M + N + K)
    A = (torch.randn(M, K, device=cuda_dev) * 0.5).half()
    B = (torch.randn(N, K, device=cuda_dev) * 0.5).half()
    bias = torch.randn(N, device=cuda_dev)
    resid = torch.randn(M, N, device=cuda_dev)
    ref = A.float() @ B.float().t() + bias
    st = torch.cuda.current_stream().cuda_stream

    def run():
        x = resid.clone()      # in-place residual add: TMA reduce-add epilogue
        check(lib().thmr_gemm_f16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, bias.data_ptr(), x.data_ptr(), N, 0,
                                  x.data_ptr(), N, None, 0, 512, st))
        y = torch.full((M, N), float("nan"), device=cuda_dev)   # plain fp32 store epilogue
        check(lib().thmr_gemm_f16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, bias.data_ptr(), None, 0, 0,
                                  y.data_ptr(), N, None, 0, 512, st))
        torch.cuda.synchronize()
        return x, y

    x, y = run()
    tol = 1e-5 if K <= 2048 else 3e-5          # fp32 summation-order noise grows with sqrt(K)
    assert rel_err(x, ref + resid) < tol
    assert rel_err(y, ref) < tol
    assert not torch.isnan(y).any()
    # bit-reproducible run to run (stream-K orders the two partial sums of a shared tile)
    for _ in range(3):
        x2, y2 = run()
        assert torch.equal(x2, x) and torch.equal(y2, y)


def test_gemm_linearity_at_full_size(cuda_dev):
    """Size-independent property at the ViT's real GEMM shape: f(a1 + a2) == f(a1) + f(a2) up to fp32 noise
    when the fp16 operands are exactly representable sums."""
    from tokenhmr_b200 import ops
    torch.manual_seed(1)
    M, N, K = 12288, 3840, 1280
    a1 = torch.randint(-8, 9, (M, K), device=cuda_dev).half()
    a2 = torch.randint(-8, 9, (M, K), device=cuda_dev).half()
    W = (torch.randint(-8, 9, (N, K), device=cuda_dev).float() / 8).half()
    y1, _ = ops.linear_f16(a1, W)
    y2, _ = ops.linear_f16(a2, W)
    y12, _ = ops.linear_f16(a1 + a2, W)
    assert torch.equal(y12, y1 + y2)      # small-integer products: every partial sum is exact in fp32


# ------------------------------------------------------------------------------------------------ LN / conv / attention
@pytest.mark.parametrize("R,C,eps", [(384, 1280, 1e-6), (5, 1024, 1e-5), (320, 64, 1e-5), (3, 10240, 1e-5), (1, 1280, 1e-6)])
def test_layernorm(cuda_dev, R, C, eps):
    from tokenhmr_b200 import ops
    torch.manual_seed(R)
    x = torch.randn(R, C, device=cuda_dev) * 3 + 1
    g, b = torch.randn(C, device=cuda_dev), torch.randn(C, device=cuda_dev)
    y16, y32 = ops.layernorm(x, g, b, eps, out16=True, out32=True)
    ref = F.layer_norm(x, (C,), g, b, eps)
    assert rel_err(y32, ref) < 2e-6 and rel_err(y16, ref) < 1e-3
    _, yr = ops.layernorm(x, g, b, eps, relu=True, out16=False, out32=True)
    assert rel_err(yr, F.relu(ref)) < 2e-6


@pytest.mark.parametrize("dil", [1, 3])
def test_conv1d_k3(cuda_dev, dil):
    from tokenhmr_b200 import ops
    torch.manual_seed(dil)
    B, L, pad, Cin, Cout = 3, 21, 3, 512, 512
    x = torch.zeros(B, L + 2 * pad, Cin, device=cuda_dev)
    x[:, pad:pad + L] = torch.randn(B, L, Cin, device=cuda_dev)
    w = torch.randn(Cout, Cin, 3, device=cuda_dev) * 0.05
    bias = torch.randn(Cout, device=cuda_dev)
    wt = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).half().contiguous()
    o32, o16 = ops.conv1d_k3_f16(x.half(), wt, bias, L, pad, dil, "relu")
    ref = F.conv1d(x[:, pad:pad + L].half().float().permute(0, 2, 1), w.half().float(), bias, padding=dil,
                   dilation=dil).permute(0, 2, 1)
    assert rel_err(o32[:, pad:pad + L], ref) < 1e-5
    assert rel_err(o16[:, pad:pad + L], F.relu(ref)) < 2e-3
    assert o32[:, :pad].abs().max() == 0 and o32[:, pad + L:].abs().max() == 0     # padding rows stay zero


@pytest.mark.parametrize("B,H", [(1, 1), (2, 16), (10, 16), (3, 5)])
def test_vit_attention(cuda_dev, B, H):
    from tokenhmr_b200 import ops
    torch.manual_seed(B * 31 + H)
    qkv = (torch.randn(B * 192, 3 * H * 80, device=cuda_dev) * 1.5).half()
    out, S = ops.vit_attention(qkv, B, H, return_scores=True)
    q, k, v = qkv.float().view(B, 192, 3, H, 80).permute(2, 0, 3, 1, 4)
    Sr = q @ k.transpose(-1, -2)
    assert rel_err(S.view(B, H, 192, 192), Sr) < 1e-5           # raw scores: fp32 accumulate of fp16 products
    s = Sr * 80 ** -0.5
    p = torch.exp(s - s.amax(-1, keepdim=True))
    o = ((p.half().float() @ v) / p.sum(-1, keepdim=True)).transpose(1, 2).reshape(B * 192, H * 80)
    assert rel_err(out, o) < 2e-3                                # fp16 probabilities + fp16 output rounding
    ref32 = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * 192, H * 80)   # vit.py:116-122 in fp32
    assert rel_err(out, ref32) < 3e-3


def test_vit_attention_rows_are_convex_combinations(cuda_dev):
    """Property at bs=64: with V == 1 everywhere the output is exactly 1 (softmax rows sum to one)."""
    from tokenhmr_b200 import ops
    B, H = 64, 16
    qkv = torch.randn(B * 192, 3 * H * 80, device=cuda_dev).half()
    qkv[:, 2 * H * 80:] = 1.0
    out = ops.vit_attention(qkv, B, H)
    assert (out.float() - 1).abs().max() < 2e-3


# ------------------------------------------------------------------------------------------------ VQ
def test_vq_quantize_golden_and_oracle(cuda_dev, golden_dir):
    from oracle import tokenhmr_oracle as O
    from tokenhmr_b200 import ops
    g = np.load(golden_dir / "vq_quantize.npz")
    cb = torch.randn(2048, 256, generator=torch.Generator().manual_seed(1))
    x = torch.randn(4096, 256, generator=torch.Generator().manual_seed(2))
    idx = ops.vq_quantize(x.to(cuda_dev), cb.to(cuda_dev)).cpu()
    assert idx.dtype == torch.int64
    mism = idx.numpy() != g["idx_rand"]
    # bit-exact indices except where the reference's own top-2 distance gap is below fp32 GEMM noise
    assert (g["gap_rand"][mism] < 1e-3).all() and mism.sum() <= 2, mism.sum()
    gg = torch.Generator().manual_seed(7)
    pick = torch.randint(0, 2048, (4096,), generator=gg)
    xn = cb[pick] + 0.05 * torch.randn(4096, 256, generator=gg)
    assert np.array_equal(ops.vq_quantize(xn.to(cuda_dev), cb.to(cuda_dev)).cpu().numpy(), g["idx_near"])
    # edge cases: the codebook itself, a single query, duplicated codes (first minimum wins), ragged Q
    assert torch.equal(ops.vq_quantize(cb.to(cuda_dev), cb.to(cuda_dev)).cpu(), torch.arange(2048))
    assert ops.vq_quantize(cb[5:6].to(cuda_dev), cb.to(cuda_dev)).item() == 5
    cb2 = cb.clone(); cb2[1000] = cb2[10]
    assert ops.vq_quantize(cb2[1000:1001].to(cuda_dev), cb2.to(cuda_dev)).item() == 10
    xr = torch.randn(777, 256)
    got = ops.vq_quantize(xr.to(cuda_dev), cb.to(cuda_dev)).cpu()
    want = O.vq_quantize(xr, cb)
    bad = got != want
    assert bad.sum() <= 1 and (O.vq_top2_gap(xr, cb)[bad] < 1e-3).all()
    # dequantize / dequantize_logits
    assert torch.equal(ops.vq_dequantize(idx.to(cuda_dev), cb.to(cuda_dev)).cpu(), cb[idx])
    deq = ops.vq_dequantize_logits(torch.from_numpy(g["logits"]).to(cuda_dev), cb.to(cuda_dev))
    assert rel_err(deq, torch.from_numpy(g["dequant_logits"])) < 2e-3      # fp16 operands, fp32 accumulate


def test_vq_quantize_1m_queries_round_trip(cuda_dev):
    """BASELINE config 4 size: 1 M queries x 2048 codes x 256 dims; property: quantising noisy codes returns the
    code they were drawn from, and a second quantise of the dequantised result is idempotent."""
    from tokenhmr_b200 import ops
    torch.manual_seed(3)
    cb = torch.randn(2048, 256, device=cuda_dev)
    pick = torch.randint(0, 2048, (1_000_000,), device=cuda_dev)
    x = cb[pick] + 0.05 * torch.randn(1_000_000, 256, device=cuda_dev)
    idx = ops.vq_quantize(x, cb)
    assert torch.equal(idx, pick)
    assert torch.equal(ops.vq_quantize(ops.vq_dequantize(idx, cb), cb), idx)


def test_vq_quantize_64k_vs_reference_golden(cuda_dev, golden_dir):
    """The default (screened, two-pass) schedule against the LIVE reference's QuantizeEMAReset.quantize on 65536 unstructured
    queries: identical indices except where the reference's own top-2 distance gap is below fp32 GEMM noise (4 of the 65536
    rows have a gap < 1e-3; a CPU emulation of the split-precision arithmetic reproduces the reference on all of them)."""
    from tokenhmr_b200 import ops
    g = np.load(golden_dir / "vq_quantize_64k.npz")
    cb = torch.randn(2048, 256, generator=torch.Generator().manual_seed(1))
    x = torch.randn(65536, 256, generator=torch.Generator().manual_seed(12))
    idx = ops.vq_quantize(x.to(cuda_dev), cb.to(cuda_dev)).cpu().numpy()
    mism = idx != g["idx"].astype(np.int64)
    assert (g["gap"][mism] < 1e-3).all(), (int(mism.sum()), float(g["gap"][mism].max()))


def _vq_both(x, cb):
    """(screened, exact) indices of the two arithmetically equivalent schedules of thmr_vq_argmin (csrc/vq.cuh)."""
    import os
    from tokenhmr_b200 import ops
    prev = os.environ.get("THMR_VQ_SCREEN")
    try:
        os.environ["THMR_VQ_SCREEN"] = "1"
        a = ops.vq_quantize(x, cb)
        os.environ["THMR_VQ_SCREEN"] = "0"
        b = ops.vq_quantize(x, cb)
    finally:
        if prev is None:
            os.environ.pop("THMR_VQ_SCREEN", None)
        else:
            os.environ["THMR_VQ_SCREEN"] = prev
    return a, b


def test_vq_screened_equals_exact(cuda_dev):
    """The screened arg-min (one fp16 product per pair, rows whose top-2 margin does not clear the rigorous error bound
    re-done by the exact 3-product pass) must return the SAME index as the exact pass on every row: unstructured queries
    (~11 % re-done), queries next to a code (none re-done), several pass-1 chunks, ragged row counts, a codebook of
    identical codes (every row re-done, more rows than one exact-pass round holds: first minimum = 0) and duplicated
    codes."""
    g = torch.Generator(cuda_dev).manual_seed(11)
    cb = torch.randn(2048, 256, device=cuda_dev, generator=g)
    for Q in (8192, 100_003, 300_001):
        x = torch.randn(Q, 256, device=cuda_dev, generator=g)
        a, b = _vq_both(x, cb)
        assert torch.equal(a, b), (Q, int((a != b).sum()))
    pick = torch.randint(0, 2048, (150_000,), device=cuda_dev, generator=g)
    near = cb[pick] + 0.05 * torch.randn(150_000, 256, device=cuda_dev, generator=g)
    a, b = _vq_both(near, cb)
    assert torch.equal(a, b) and torch.equal(a, pick)
    # scaled-up data (larger norms -> larger margins and bounds alike) and tiny data
    for scale in (7.0, 1e-3):
        x = scale * torch.randn(50_000, 256, device=cuda_dev, generator=g)
        a, b = _vq_both(x, scale * cb)
        assert torch.equal(a, b), scale
    same = cb[:1].expand(2048, 256).contiguous()
    x = torch.randn(300_001, 256, device=cuda_dev, generator=g)
    a, b = _vq_both(x, same)
    assert torch.equal(a, b) and int(a.abs().max()) == 0
    cb2 = cb.clone()
    cb2[1500] = cb2[20]
    x = cb2[torch.randint(0, 2048, (20_000,), device=cuda_dev, generator=g)] + 0.01 * torch.randn(20_000, 256, device=cuda_dev, generator=g)
    a, b = _vq_both(x, cb2)
    assert torch.equal(a, b) and int((a == 1500).sum()) == 0


# ------------------------------------------------------------------------------------------------ geometry / SMPL
def test_rot6d_golden(cuda_dev, golden_dir):
    from tokenhmr_b200 import ops
    g = np.load(golden_dir / "geometry.npz")
    R = ops.rot6d_to_rotmat(torch.from_numpy(g["x6"]).to(cuda_dev))
    # Gram-Schmidt subtracts nearly parallel vectors: fp32 rounding-order differences are amplified ~10x
    assert rel_err(R, torch.from_numpy(g["rotmat"])) < 5e-5
    eye = ops.rot6d_to_rotmat(torch.tensor([[1., 0, 0, 0, 1, 0]], device=cuda_dev))
    assert torch.equal(eye[0].cpu(), torch.eye(3))


@pytest.fixture(scope="module")
def smpl_pair(cuda_dev):
    from tokenhmr_b200 import ops, synth
    from tokenhmr_b200.config import release_config
    smpl = synth.make_smpl(release_config())
    return smpl, ops.SMPLModel(smpl, cuda_dev)


def test_lbs_vs_oracle_and_fixture(cuda_dev, smpl_pair, golden_dir):
    from oracle import smpl_oracle as S
    smpl, m = smpl_pair
    g = np.load(golden_dir / "smpl_lbs_f64.npz")
    aa, betas = torch.from_numpy(g["aa"]), torch.from_numpy(g["betas"])
    v, j = m.lbs(betas.to(cuda_dev), aa.to(cuda_dev), pose2rot=True)
    vr, jr = S.lbs(betas, aa.reshape(8, -1), smpl["v_template"], smpl["shapedirs"], smpl["posedirs"],
                   smpl["J_regressor"], smpl["parents"], smpl["lbs_weights"], pose2rot=True)
    assert rel_err(v, vr) < 1e-4 and (v.cpu() - vr).abs().max() < 1e-4     # north-star tolerance: 1e-4
    assert rel_err(j, jr) < 1e-4
    R = S.batch_rodrigues(aa.view(-1, 3)).view(8, 24, 3, 3)
    v2, j2 = m.lbs(betas.to(cuda_dev), R.to(cuda_dev), pose2rot=False)
    assert rel_err(v2, vr) < 1e-4 and rel_err(j2, jr) < 1e-4
    cam = torch.tensor([[0.9, 0.1, -0.05]]).repeat(8, 1)
    v3, j44, cam_t, focal, kp2d = m.forward(R[:, :1].to(cuda_dev), R[:, 1:].to(cuda_dev), betas.to(cuda_dev),
                                            pred_cam=cam.to(cuda_dev))
    assert rel_err(v3, torch.from_numpy(g["verts"])) < 1e-4 and rel_err(j44, torch.from_numpy(g["joints"])) < 1e-4
    from oracle import tokenhmr_oracle as O
    want_t = torch.stack([cam[:, 1], cam[:, 2], 2 * 5000.0 / (256 * cam[:, 0] + 1e-9)], -1)
    assert rel_err(cam_t, want_t) < 1e-6 and torch.equal(focal.cpu(), torch.full((8, 2), 5000.0))
    want2d = O.perspective_projection(torch.from_numpy(g["joints"]), want_t, torch.full((8, 2), 5000.0 / 256))
    assert rel_err(kp2d, want2d) < 1e-4


def test_lbs_skin_launch_shapes_agree(cuda_dev, smpl_pair):
    """The skinning kernel's two launch shapes (128 / 256 vertices per block, THMR_SKIN_THREADS) do the same arithmetic:
    bit-identical vertices, also across the 512-pose chunk boundary and for a ragged last pose tile."""
    import os
    smpl, m = smpl_pair
    torch.manual_seed(9)
    aa = 0.3 * torch.randn(1100 + 7, 24, 3, device=cuda_dev)
    betas = torch.randn(1100 + 7, 10, device=cuda_dev)
    got = {}
    prev = os.environ.get("THMR_SKIN_THREADS")
    try:
        for t in ("128", "256"):
            os.environ["THMR_SKIN_THREADS"] = t
            v, j = m.lbs(betas, aa, pose2rot=True)
            got[t] = (v.clone(), j.clone())
    finally:
        if prev is None:
            os.environ.pop("THMR_SKIN_THREADS", None)
        else:
            os.environ["THMR_SKIN_THREADS"] = prev
    assert torch.equal(got["128"][0], got["256"][0]) and torch.equal(got["128"][1], got["256"][1])


def test_lbs_identity_pose_and_batch_4096(cuda_dev, smpl_pair):
    """BASELINE config 5: 4096 poses vs the oracle within 1e-4, plus the rest-pose invariant."""
    from oracle import smpl_oracle as S
    smpl, m = smpl_pair
    betas = torch.randn(3, 10)
    eyeR = torch.eye(3).expand(3, 24, 3, 3).contiguous()
    v, _ = m.lbs(betas.to(cuda_dev), eyeR.to(cuda_dev), pose2rot=False)
    want = smpl["v_template"] + torch.einsum("bl,vkl->bvk", betas, smpl["shapedirs"])
    assert rel_err(v, want) < 1e-6
    torch.manual_seed(4)
    aa = 0.3 * torch.randn(4096, 24, 3)
    betas = torch.randn(4096, 10)
    v, j = m.lbs(betas.to(cuda_dev), aa.to(cuda_dev), pose2rot=True)
    vr, jr = S.lbs(betas, aa.reshape(4096, -1), smpl["v_template"], smpl["shapedirs"], smpl["posedirs"],
                   smpl["J_regressor"], smpl["parents"], smpl["lbs_weights"], pose2rot=True)
    assert rel_err(v, vr) < 1e-4 and (v.cpu() - vr).abs().max() < 1e-4 and rel_err(j, jr) < 1e-4
