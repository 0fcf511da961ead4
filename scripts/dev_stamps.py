"""Per-launch times inside the CUDA-graph replay (in-graph start stamps), averaged per position in the ViT block:
ln1, qkv, attention, proj, ln2, fc1, fc2.  usage: python scripts/dev_stamps.py [samples]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tokenhmr_b200 import synth
from tokenhmr_b200.config import release_config
from tokenhmr_b200.engine import TokenHMREngine
cfg = release_config()
model = TokenHMREngine(cfg, synth.make_state_dict(cfg), synth.make_smpl(cfg), use_cuda_graph=True)
img = synth.make_images(64, cfg).cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
acc = collections.defaultdict(list)
for _ in range(n):
    rows = model.profile_in_graph(img, replays=20)
    pos = collections.Counter()
    for name, ms, fl, by in rows:
        if name.startswith("vit.") and name not in ("vit.patch_im2col", "vit.patch_embed_gemm"):
            key = name
            if name == "vit.layernorm":
                key = "vit.layernorm#%d" % (pos[name] % 2 + 1)
            pos[name] += 1
            acc[key].append(ms * 1e3)
        else:
            acc[name].append(ms * 1e3)
tot = 0
for k, v in acc.items():
    per = sum(v) / len(v)
    cnt = len(v) / n
    tot += per * cnt
    print(f"{k:26s} launches/step {cnt:5.1f}  us/launch {per:8.2f}  ms/step {per*cnt/1e3:7.3f}")
print(f"sum {tot/1e3:.3f} ms")
