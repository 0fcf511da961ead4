import torch, os
from tokenhmr_b200 import ops
dev = torch.device("cuda:0"); torch.manual_seed(0)
B, H = 3, 16
qkv = (torch.randn(B * 192, 3 * H * 80, device=dev) * 1.5).half()
out = ops.vit_attention(qkv, B, H)
q, k, v = qkv.float().view(B, 192, 3, H, 80).permute(2, 0, 3, 1, 4)
s = (q @ k.transpose(-1, -2)) * 80 ** -0.5
p = torch.exp(s - s.amax(-1, keepdim=True))
o = ((p.half().float() @ v) / p.sum(-1, keepdim=True)).transpose(1, 2).reshape(B * 192, H * 80)
print("THMR_ATTN_TS =", os.environ.get("THMR_ATTN_TS"), "rel err", ((out.float() - o).abs().max() / o.abs().max()).item())
from tokenhmr_b200._lib import lib
print("flags", lib().thmr_check_device_flags())
