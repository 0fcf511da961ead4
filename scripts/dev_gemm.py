"""Dev check of the tcgen05 GEMM on a real B200 (correctness vs torch fp32, perf vs cuBLAS)."""
import sys
import torch
from tokenhmr_b200._lib import lib, check

L = lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)
st = lambda: torch.cuda.current_stream().cuda_stream
P = lambda t: None if t is None else t.data_ptr()

def gemm(A, B, bias=None, resid=None, act=0, o32=None, o16=None, bn=0):
    M, K = A.shape; N = B.shape[0]
    check(L.thmr_gemm_f16(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K, P(bias), P(resid), N, act,
                          P(o32), N, P(o16), N, bn, st()))

ok = True
def report(name, good, msg):
    global ok
    ok &= good
    print(f"{name}: {msg} {'OK' if good else 'FAIL'}", flush=True)

for (M, N, K) in [(128, 256, 64), (300, 520, 200), (1024, 1280, 5120), (64, 1024, 1024), (1000, 96, 160), (1536, 1280, 1280), (130, 6, 1536), (1000, 1280, 320)]:
    for bn in (0, 256, 128, 64, 32, 512):
        if bn == 512 and (M < 128 or N % 8): continue
        Kp = (K + 7) // 8 * 8
        A = torch.randn(M, Kp, device=dev).half()[:, :K]; B = torch.randn(N, Kp, device=dev).half()[:, :K]
        bias = torch.randn(N, device=dev); resid = torch.randn(M, N, device=dev)
        r = A.float() @ B.float().t() + bias
        # (a) both outputs, separate residual -> generic epilogue
        o32 = torch.empty(M, N, device=dev); o16 = torch.empty(M, N, device=dev, dtype=torch.float16)
        gemm(A, B, bias, resid, 1, o32, o16, 0 if bn == 512 else bn)
        e32 = ((o32 - (r + resid)).abs().max() / (r + resid).abs().max()).item()
        g = torch.nn.functional.gelu(r + resid)
        e16 = ((o16.float() - g).abs().max() / g.abs().max()).item()
        # (b) fp16-only output with GELU -> TMA store epilogue when eligible
        o16b = torch.zeros(M, N, device=dev, dtype=torch.float16)
        gemm(A, B, bias, None, 1, None, o16b, bn)
        gb = torch.nn.functional.gelu(r)
        e16b = ((o16b.float() - gb).abs().max() / gb.abs().max()).item()
        # (c) in-place residual add -> TMA reduce-add epilogue when eligible
        x = resid.clone()
        gemm(A, B, bias, x, 0, x, None, bn)
        ec = ((x - (r + resid)).abs().max() / (r + resid).abs().max()).item()
        rc = L.thmr_check_device_flags()
        report(f"M={M} N={N} K={K} bn={bn}", e32 < 2e-5 and e16 < 2e-3 and e16b < 2e-3 and ec < 2e-5 and rc == 0,
               f"gen32={e32:.1e} gen16={e16:.1e} store16={e16b:.1e} add32={ec:.1e} flags={rc}")
        if rc != 0: print(L.thmr_last_error()); sys.exit(1)
print("CORRECTNESS", "PASS" if ok else "FAIL", flush=True)

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

M = 12288
for (name, N, K, act, mode) in [("qkv", 3840, 1280, 0, "s16"), ("proj", 1280, 1280, 0, "add"), ("fc1", 5120, 1280, 1, "s16"),
                                ("fc2", 1280, 5120, 0, "add"), ("to_kv", 6144, 1280, 0, "s16")]:
    A = torch.randn(M, K, device=dev).half(); B = (torch.randn(N, K, device=dev) * 0.03).half()
    bias = torch.randn(N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.float16); x = torch.zeros(M, N, device=dev)
    tcb = timeit(lambda: torch.matmul(A, B.t()))
    line = f"{name} N={N} K={K}: cublas {tcb*1e3:.1f}us {2*M*N*K/tcb/1e9:.0f} TF |"
    for bn in (512, 256):
        if mode == "s16": fn = lambda: gemm(A, B, bias, None, act, None, o16, bn)
        else: fn = lambda: gemm(A, B, bias, x, 0, x, None, bn)
        t = timeit(fn)
        line += f" bn{bn} {t*1e3:.1f}us {2*M*N*K/t/1e9:.0f} TF |"
    print(line, flush=True)
print("flags", L.thmr_check_device_flags())
