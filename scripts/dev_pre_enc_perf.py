"""Timing of the two 'next-row' components on the GPU (CUDA events, after warm-up):
   * thmr_preprocess_boxes: 1080p frame, 64 person boxes (8-bit path) and 8 large boxes (blurred path)
   * thmr_tok_encode: 4096 poses -> 655 360 pose tokens
Prints one JSON line per case (algorithmic bytes / time -> GB/s against the measured HBM peak)."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from tokenhmr_b200 import synth
from tokenhmr_b200.config import release_config
from tokenhmr_b200.preprocess import ViTDetPreprocessor
from tokenhmr_b200.tokenizer import EncodeTokens


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    H, W = 1080, 1920
    img = torch.from_numpy(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).to(dev)
    pre = ViTDetPreprocessor(device=dev)
    for name, n, lo, hi in (("u8", 64, 80, 500), ("blur", 8, 600, 1000)):
        cx, cy = rng.uniform(200, W - 200, n), rng.uniform(200, H - 200, n)
        w, h = rng.uniform(lo * 0.5, hi * 0.6, n), rng.uniform(lo, hi, n)
        boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)
        ms = timed(lambda: pre(img, boxes))
        out_bytes = n * 3 * 256 * 256 * 4
        print(json.dumps({"case": f"preprocess_{name}", "boxes": n, "ms": ms, "persons_per_s": n / ms * 1e3,
                          "out_GBps": out_bytes / ms / 1e6}))
    cfg = release_config()
    enc = EncodeTokens(cfg, synth.make_tokenizer_encoder_state_dict(cfg), device=dev)
    for B in (64, 4096):
        x = torch.randn(B, 21, 6, device=dev)
        ms = timed(lambda: enc(x), n=10)
        flops = 0
        L = 21
        flops += 2 * B * L * 512 * 18
        for Lc in (40, 80, 160, 320):
            flops += 2 * B * Lc * 512 * 1536
        flops += 2 * B * 160 * 512 * 2048 + 2 * (2 * B * 160 * 512 * (1536 + 512)) + 2 * B * 160 * 256 * 1536
        flops += 3 * 2 * B * 160 * 2048 * 256
        print(json.dumps({"case": "tok_encode", "poses": B, "ms": ms, "poses_per_s": B / ms * 1e3,
                          "tokens_per_s": B * 160 / ms * 1e3, "tflops": flops / ms / 1e9}))


if __name__ == "__main__":
    main()
