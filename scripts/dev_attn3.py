"""Two-chain attention kernel: timing + per-role cycle counters."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

dev = torch.device("cuda:0"); torch.manual_seed(0)
cnt = torch.zeros(148 * 32, device=dev, dtype=torch.int64)
os.environ["THMR_ATTN_COUNTERS"] = hex(cnt.data_ptr())
from tokenhmr_b200 import ops

B, H = 64, 16
qkv = (torch.randn(B * 192, 3 * H * 80, device=dev)).half()
big = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def run(n, flush):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); tot = 0
    for _ in range(n):
        if flush: big.zero_()
        e0.record(); ops.vit_attention(qkv, B, H); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n


out = ops.vit_attention(qkv, B, H)
q, k, v = qkv.float().view(B, 192, 3, H, 80).permute(2, 0, 3, 1, 4)
ref = torch.softmax((q * 80 ** -0.5) @ k.transpose(-1, -2), -1) @ v
ref = ref.transpose(1, 2).reshape(B * 192, H * 80)
print("max abs err vs fp32 torch:", (out.float() - ref).abs().max().item(), "ref max", ref.abs().max().item())
run(3, False)
print(f"attention bs=64 knobs={os.environ.get('THMR_ATTN_TS', '0')}: L2-warm {run(20, False)*1e3:.1f} us, L2-flushed {run(20, True)*1e3:.1f} us per layer")
c = cnt.view(148, 32).float().mean(0).tolist()
print(f"MMA issuer 0: total {c[0]:.0f} wait o_empty {c[1]:.0f} qk_full {c[16]:.0f} p_full {c[17]:.0f} v_full {c[18]:.0f}")
print(f"MMA issuer 1: total {c[8]:.0f} wait o_empty {c[15]:.0f} qk_full {c[19]:.0f} p_full {c[20]:.0f} v_full {c[21]:.0f}")
for g, o in ((0, 2), (1, 9)):
    print(f"group {g} (warp q=2): total {c[o]:.0f} wait s_full {c[o+1]:.0f} wait o_full {c[o+2]:.0f} pass1 {c[o+3]:.0f} pass2 {c[o+4]:.0f} epi {c[o+5]:.0f} wait turn {c[22+g]:.0f}")
