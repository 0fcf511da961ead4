import torch, time, sys
from tokenhmr_b200 import ops, synth
from tokenhmr_b200.config import release_config
dev = torch.device("cuda:0")
what = sys.argv[1:] or ["vq", "lbs"]
def ev(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
if "vq" in what:
    cb = torch.randn(2048, 256, device=dev); x = torch.randn(1_000_000, 256, device=dev)
    import os
    for mode in ("1", "0"):
        os.environ["THMR_VQ_SCREEN"] = mode
        ms = ev(lambda: ops.vq_quantize(x, cb))
        print(f"vq 1M ({'screened' if mode == '1' else 'exact'}): {ms:.2f} ms  {1e3/ms:.0f} Mq/s  ({2*1e6*2048*256/ms/1e9:.0f} TFLOP/s algorithmic)")
    os.environ.pop("THMR_VQ_SCREEN")
    pick = torch.randint(0, 2048, (100000,), device=dev)
    xn = cb[pick] + 0.05 * torch.randn(100000, 256, device=dev)
    print("vq near-code exact:", torch.equal(ops.vq_quantize(xn, cb), pick))
if "lbs" in what:
    cfg = release_config(); m = ops.SMPLModel(synth.make_smpl(cfg), dev)
    B = 4096
    aa = 0.3 * torch.randn(B, 24, 3, device=dev); be = torch.randn(B, 10, device=dev)
    ms = ev(lambda: m.lbs(be, aa))
    print(f"lbs 4096: {ms:.3f} ms {B/ms/1e3:.2f} Mposes/s  algorithmic {B*84.1e3/ms/1e6:.0f} GB/s")
