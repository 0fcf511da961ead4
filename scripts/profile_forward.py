"""Two eager (no CUDA graph) release-config forwards at bs=64 for ncu launch lists / captures."""
import sys, torch
from tokenhmr_b200 import synth
from tokenhmr_b200.config import release_config
from tokenhmr_b200.engine import TokenHMREngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = release_config()
model = TokenHMREngine(cfg, synth.make_state_dict(cfg), synth.make_smpl(cfg), use_cuda_graph=False)
img = synth.make_images(B, cfg).cuda()
for _ in range(n):
    model({"img": img})
torch.cuda.synchronize()
print("done", model.num_launches())
