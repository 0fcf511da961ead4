import torch, os
dev = torch.device("cuda:0"); torch.manual_seed(0)
cnt = torch.zeros(148 * 16, device=dev, dtype=torch.int64)
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
run(3, False)
print(f"attention bs=64: L2-warm {run(20, False)*1e3:.1f} us, L2-flushed {run(20, True)*1e3:.1f} us per layer")
c = cnt.view(148, 16).float().mean(0).tolist()
print(f"MMA thread: wait qk_full {c[0]:.0f} o_empty {c[1]:.0f} v_full {c[2]:.0f} p_full {c[3]:.0f} / total {c[4]:.0f}")
print(f"softmax detail: ldtm {c[5]:.0f} max+xchg {c[6]:.0f} exp {c[7]:.0f} (p_empty wait+STS+fence, from p_empty wait start) {c[13]:.0f} epi(o_full wait+ld+store) {c[14]:.0f}")
print(f"softmax warp(q=0,h=0): wait s_full {c[8]:.0f} p_empty {c[9]:.0f} o_full {c[10]:.0f} pairbar {c[11]:.0f} / total {c[12]:.0f}  (7 heads per CTA)")
