"""Dev check of every kernel family on a real B200 against torch / the oracle."""
import sys, time, traceback
import numpy as np
import torch
import torch.nn.functional as F
from tokenhmr_b200 import ops, synth
from tokenhmr_b200._lib import lib
from tokenhmr_b200.config import tiny_config, release_config
from tokenhmr_b200.engine import TokenHMREngine
from oracle import tokenhmr_oracle as O, smpl_oracle

dev = torch.device("cuda:0")
torch.manual_seed(0)
L = lib()
which = sys.argv[1:] or ["ln", "attn", "conv", "vq", "smpl", "tiny", "release", "bench"]

def rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()

def flags():
    rc = L.thmr_check_device_flags()
    if rc != 0: print("!! device flag:", L.thmr_last_error())
    return rc

def run(name, fn):
    if name not in which: return
    print(f"==== {name}", flush=True)
    try:
        fn()
    except Exception:
        traceback.print_exc()
    flags()

def t_ln():
    for (R, C, eps) in [(384, 1280, 1e-6), (5, 1024, 1e-5), (320, 64, 1e-5), (3, 10240, 1e-5)]:
        x = torch.randn(R, C, device=dev) * 3 + 1
        g = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
        y16, y32 = ops.layernorm(x, g, b, eps, out16=True, out32=True)
        r = F.layer_norm(x, (C,), g, b, eps)
        print(f"LN R={R} C={C}: rel32={rel(y32, r):.2e} rel16={rel(y16, r):.2e}")

def t_attn():
    for (B, H) in [(1, 1), (2, 16), (10, 16)]:
        qkv = (torch.randn(B * 192, 3 * H * 80, device=dev) * 1.5).half()
        out, S = ops.vit_attention(qkv, B, H, return_scores=True)
        torch.cuda.synchronize()
        q, k, v = qkv.float().view(B, 192, 3, H, 80).permute(2, 0, 3, 1, 4)
        Sr = q @ k.transpose(-1, -2)
        s = Sr * 80 ** -0.5
        p = torch.exp(s - s.amax(-1, keepdim=True))
        o = (p.half().float() @ v) / p.sum(-1, keepdim=True)
        o = o.transpose(1, 2).reshape(B * 192, H * 80)
        print(f"attn B={B} H={H}: S rel={rel(S.view(B, H, 192, 192), Sr):.2e}  O rel={rel(out, o):.2e}")

def t_conv():
    B, Lq, pad, Cin, Cout = 3, 21, 3, 512, 512
    for dil in (1, 3):
        x = torch.zeros(B, Lq + 2 * pad, Cin, device=dev)
        x[:, pad:pad + Lq] = torch.randn(B, Lq, Cin, device=dev)
        w = torch.randn(Cout, Cin, 3, device=dev) * 0.05
        bias = torch.randn(Cout, device=dev)
        wt = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).half().contiguous()
        o32, o16 = ops.conv1d_k3_f16(x.half(), wt, bias, Lq, pad, dil, "relu")
        r = F.conv1d(x[:, pad:pad + Lq].half().float().permute(0, 2, 1), w.half().float(), bias, padding=dil, dilation=dil)
        r = r.permute(0, 2, 1)
        print(f"conv dil={dil}: rel32={rel(o32[:, pad:pad + Lq], r):.2e} rel16(relu)={rel(o16[:, pad:pad + Lq], F.relu(r)):.2e} "
              f"padrows={o32[:, :pad].abs().max().item():.1e}/{o32[:, pad + Lq:].abs().max().item():.1e}")

def t_vq():
    g = np.load("tests/golden/vq_quantize.npz")
    cb = torch.randn(2048, 256, generator=torch.Generator().manual_seed(1)).to(dev)
    xr = torch.randn(4096, 256, generator=torch.Generator().manual_seed(2)).to(dev)
    idx = ops.vq_quantize(xr, cb).cpu().numpy()
    mism = idx != g["idx_rand"]
    print(f"vq rand: mismatches {mism.sum()} / 4096; gaps at mismatches: {g['gap_rand'][mism][:5]}")
    gg = torch.Generator().manual_seed(7)
    pick = torch.randint(0, 2048, (4096,), generator=gg)
    xn = cb.cpu()[pick] + 0.05 * torch.randn(4096, 256, generator=gg)
    idx2 = ops.vq_quantize(xn.to(dev), cb).cpu().numpy()
    print(f"vq near: mismatches {(idx2 != g['idx_near']).sum()} / 4096, vs pick {(idx2 != g['pick']).sum()}")
    deq = ops.vq_dequantize_logits(torch.from_numpy(g["logits"]).to(dev), cb)
    print(f"vq dequant_logits rel={rel(deq, torch.from_numpy(g['dequant_logits']).to(dev)):.2e}")
    d = ops.vq_dequantize(torch.from_numpy(idx).to(dev), cb)
    print("vq dequantize exact:", torch.equal(d.cpu(), cb.cpu()[torch.from_numpy(idx)]))
    # perf: 1M queries
    Q = 1_000_000
    x = torch.randn(Q, 256, device=dev)
    for _ in range(2): ops.vq_quantize(x, cb)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(5): ops.vq_quantize(x, cb)
    torch.cuda.synchronize(); dt = (time.time() - t) / 5
    print(f"vq 1M queries: {dt * 1e3:.2f} ms -> {Q / dt / 1e6:.1f} Mq/s")

def t_smpl():
    cfg = release_config()
    smpl = synth.make_smpl(cfg)
    m = ops.SMPLModel(smpl, dev)
    g = np.load("tests/golden/smpl_lbs_f64.npz")
    aa = torch.from_numpy(g["aa"]); betas = torch.from_numpy(g["betas"])
    v, j = m.lbs(betas.to(dev), aa.to(dev), pose2rot=True)
    vr, jr = smpl_oracle.lbs(betas, aa.reshape(8, -1), smpl["v_template"], smpl["shapedirs"], smpl["posedirs"],
                             smpl["J_regressor"], smpl["parents"], smpl["lbs_weights"], pose2rot=True)
    print(f"lbs(aa): verts rel={rel(v.cpu(), vr):.2e} abs={(v.cpu() - vr).abs().max():.2e} joints rel={rel(j.cpu(), jr):.2e}")
    R = smpl_oracle.batch_rodrigues(aa.view(-1, 3)).view(8, 24, 3, 3)
    cam = torch.tensor([[0.9, 0.1, -0.05]]).repeat(8, 1)
    v2, j2, ct, fl, k2 = m.forward(R[:, :1].to(dev), R[:, 1:].to(dev), betas.to(dev), pred_cam=cam.to(dev))
    print(f"smpl fwd vs f64 golden: verts rel={rel(v2.cpu(), torch.from_numpy(g['verts'])):.2e} joints rel={rel(j2.cpu(), torch.from_numpy(g['joints'])):.2e}")
    # perf 4096 poses
    B = 4096
    aa = 0.3 * torch.randn(B, 24, 3, device=dev); be = torch.randn(B, 10, device=dev)
    for _ in range(2): m.lbs(be, aa)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(5): m.lbs(be, aa)
    torch.cuda.synchronize(); dt = (time.time() - t) / 5
    print(f"lbs 4096 poses: {dt * 1e3:.3f} ms -> {B / dt / 1e6:.2f} Mposes/s")

def cmp_out(o, r, keys):
    for k in keys:
        print(f"   {k:22s} rel={rel(o[k].cpu(), r[k]):.2e} abs={(o[k].cpu() - r[k]).abs().max().item():.2e}")

def t_tiny():
    cfg = tiny_config(vit_depth=2)
    sd = synth.make_state_dict(cfg); smpl = synth.make_smpl(cfg)
    img = synth.make_images(2, cfg)
    model = TokenHMREngine(cfg, sd, smpl, device=dev, use_cuda_graph=False)
    out = model({"img": img}, return_taps=True)
    torch.cuda.synchronize()
    keys = ["_vit_tokens", "_token_out", "cls_logits_softmax", "_pred_body_pose_6d", "pred_cam", "pred_cam_t", "pred_keypoints_3d", "pred_vertices", "pred_keypoints_2d"]
    r16 = O.forward(sd, smpl, img, cfg, emulate_fp16=True, return_intermediates=True)
    r32 = O.forward(sd, smpl, img, cfg, emulate_fp16=False, return_intermediates=True)
    print(" vs fp16-emulating oracle:"); cmp_out(out, r16, keys)
    print(" vs fp32 oracle:"); cmp_out(out, r32, keys)
    print(" tokens equal (emu):", (out["cls_logits_softmax"].argmax(-1).cpu() == r16["cls_logits_softmax"].argmax(-1)).float().mean().item(),
          " (fp32):", (out["cls_logits_softmax"].argmax(-1).cpu() == r32["cls_logits_softmax"].argmax(-1)).float().mean().item())
    model.use_cuda_graph = True
    out2 = model({"img": img}, return_taps=True)
    torch.cuda.synchronize()
    print(" graph replay identical:", all(torch.equal(out[k], out2[k]) for k in keys))

def t_release():
    cfg = release_config()
    t = time.time(); sd = synth.make_state_dict(cfg); smpl = synth.make_smpl(cfg); print(f" synth {time.time() - t:.1f}s")
    g = np.load("tests/golden/forward_release_d32.npz")
    img = synth.make_images(2, cfg)
    model = TokenHMREngine(cfg, sd, smpl, device=dev, use_cuda_graph=True)
    del sd
    out = model({"img": img}, return_taps=True)
    torch.cuda.synchronize()
    print(f" vit_tokens_sub rel={rel(out['_vit_tokens'][:, ::8].cpu(), torch.from_numpy(g['vit_tokens_sub'])):.2e}")
    for k in ["pred_cam", "pred_cam_t", "pred_keypoints_3d", "pred_vertices", "pred_keypoints_2d"]:
        print(f"   {k:22s} rel={rel(out[k].cpu(), torch.from_numpy(g[k])):.2e} abs={(out[k].cpu() - torch.from_numpy(g[k])).abs().max().item():.2e}")
    am = out["cls_logits_softmax"].argmax(-1).cpu().numpy()
    print("   tokens equal:", (am == g["cls_argmax"]).mean())
    if "bench" in which:
        B = 64
        img = synth.make_images(B, cfg).to(dev)
        for _ in range(3): model({"img": img})
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): model({"img": img})
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f" bs=64 forward: {ms:.2f} ms -> {B / ms * 1e3:.0f} img/s ({B * 252.1e9 / ms / 1e9:.0f} TFLOP/s), launches={model.num_launches()}")

run("ln", t_ln); run("attn", t_attn); run("conv", t_conv); run("vq", t_vq); run("smpl", t_smpl); run("tiny", t_tiny); run("release", t_release)
