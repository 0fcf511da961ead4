"""Sustained (power-capped) rate of one GEMM shape: ours vs cuBLAS, back to back for a few seconds each, with
nvidia-smi clock / power samples.  usage: python scripts/dev_sustained.py [seconds]"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tokenhmr_b200._lib import lib, check
L = lib(); dev = torch.device("cuda:0"); torch.manual_seed(0)
SEC = float(sys.argv[1]) if len(sys.argv) > 1 else 2.5
st = lambda: torch.cuda.current_stream().cuda_stream
P = lambda t: None if t is None else t.data_ptr()
def gemm(A, B, bias, resid, act, o32, o16, bn=512):
    M, K = A.shape; N = B.shape[0]
    check(L.thmr_gemm_f16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, P(bias), P(resid), N, act, P(o32), N, P(o16), N, bn, st()))
rows = []
def sampler(stop):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        rows.append((time.time(),) + tuple(float(x) for x in line.split(",")))
        if stop.is_set(): break
    p.terminate()
stop = threading.Event(); th = threading.Thread(target=sampler, args=(stop,), daemon=True); th.start(); time.sleep(0.7)
def sustained(name, fn, flops):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    # calibrate iterations for SEC seconds
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(20)]; e1.record(); torch.cuda.synchronize()
    it = max(20, int(SEC * 1e3 / (e0.elapsed_time(e1) / 20)))
    t0 = time.time(); e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); t1 = time.time()
    ms = e0.elapsed_time(e1) / it
    win = [r for r in rows if t0 + 0.5 <= r[0] <= t1]
    clk = sorted(r[1] for r in win); pw = sorted(r[2] for r in win)
    med = lambda v: v[len(v) // 2] if v else float("nan")
    print(f"{name:34s} {ms*1e3:7.1f} us  {flops/ms/1e9:6.0f} TF/s  clk {med(clk):5.0f} MHz  power {med(pw):5.0f} W  -> {flops/ms/1e9/ (8192*148*med(clk)*1e6/1e12)*100:5.1f}% of peak at that clock, {flops/ms/1e9/med(pw):.2f} TF/s/W", flush=True)
    time.sleep(1.0)
M = 12288
for (name, N, K, act, mode) in [("qkv", 3840, 1280, 0, "s16"), ("fc1", 5120, 1280, 1, "s16"), ("fc2", 1280, 5120, 0, "add"), ("proj", 1280, 1280, 0, "add")]:
    A = torch.randn(M, K, device=dev).half(); B = (torch.randn(N, K, device=dev) * 0.03).half()
    bias = torch.randn(N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.float16); x = torch.zeros(M, N, device=dev)
    fl = 2.0 * M * N * K
    sustained(f"cublas {name} {N}x{K}", lambda: torch.matmul(A, B.t(), out=o16), fl)
    if mode == "s16":
        sustained(f"ours   {name} (act={act})", lambda: gemm(A, B, bias, None, act, None, o16), fl)
        if act: sustained(f"ours   {name} (no act)", lambda: gemm(A, B, bias, None, 0, None, o16), fl)
    else:
        sustained(f"ours   {name} reduce-add", lambda: gemm(A, B, bias, x, 0, x, None), fl)
        y = torch.empty(M, N, device=dev)
        sustained(f"ours   {name} fp32 store", lambda: gemm(A, B, bias, None, 0, y, None), fl)
stop.set()
