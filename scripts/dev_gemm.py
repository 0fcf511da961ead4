"""Dev check of the tcgen05 GEMM on a real B200 (correctness vs torch fp32, perf vs cuBLAS)."""
import ctypes, sys, time
import torch
from tokenhmr_b200._lib import lib, check

L = lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)


def gemm(A, B, bias=None, resid=None, act=0, out32=True, out16=False, bn=0):
    M, K = A.shape
    N = B.shape[0]
    o32 = torch.empty(M, N, device=dev, dtype=torch.float32) if out32 else None
    o16 = torch.empty(M, N, device=dev, dtype=torch.float16) if out16 else None
    st = torch.cuda.current_stream().cuda_stream
    check(L.thmr_gemm_f16(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K,
                          bias.data_ptr() if bias is not None else None,
                          resid.data_ptr() if resid is not None else None, N, act,
                          o32.data_ptr() if out32 else None, N, o16.data_ptr() if out16 else None, N, bn, st))
    return o32, o16


def ref(A, B, bias=None, resid=None):
    r = A.float() @ B.float().t()
    if bias is not None: r = r + bias
    if resid is not None: r = r + resid
    return r

ok = True
for (M, N, K) in [(128, 256, 64), (128, 256, 128), (256, 512, 256), (300, 520, 200), (64, 1024, 1024), (1000, 96, 160),
                  (12288 // 8, 1280, 1280), (130, 6, 1536)]:
    for bn in (256, 128, 64, 32):
        Kp = (K + 7) // 8 * 8
        A = torch.randn(M, Kp, device=dev).half()[:, :K]
        B = torch.randn(N, Kp, device=dev).half()[:, :K]
        bias = torch.randn(N, device=dev)
        resid = torch.randn(M, N, device=dev)
        o32, o16 = gemm(A, B, bias, resid, act=1, out32=True, out16=True, bn=bn)
        torch.cuda.synchronize()
        rc = L.thmr_check_device_flags()
        r = ref(A, B, bias, resid)
        e32 = (o32 - r).abs().max().item() / (r.abs().max().item() + 1e-9)
        g = torch.nn.functional.gelu(r)
        e16 = (o16.float() - g).abs().max().item() / (g.abs().max().item() + 1e-9)
        good = e32 < 1e-5 and e16 < 2e-3 and rc == 0
        ok &= good
        print(f"M={M} N={N} K={K} bn={bn}: rel32={e32:.2e} rel16={e16:.2e} flags={rc} {'OK' if good else 'FAIL'}", flush=True)
        if rc != 0:
            print("timeout flag:", L.thmr_last_error()); sys.exit(1)

print("CORRECTNESS", "PASS" if ok else "FAIL", flush=True)

# perf
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

M = 12288
for (N, K) in [(3840, 1280), (1280, 1280), (5120, 1280), (1280, 5120), (6144, 1280)]:
    A = torch.randn(M, K, device=dev).half(); B = torch.randn(N, K, device=dev).half()
    bias = torch.randn(N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.float16)
    st = torch.cuda.current_stream().cuda_stream
    tcb = timeit(lambda: torch.matmul(A, B.t()))
    line = f"N={N} K={K}: cublas {tcb*1e3:.1f}us {2*M*N*K/tcb/1e9:.0f} TF |"
    for bn in (256, 128):
        t = timeit(lambda: check(L.thmr_gemm_f16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, bias.data_ptr(), None, 0, 0,
                                                   None, 0, o16.data_ptr(), N, bn, st)))
        line += f" bn{bn} {t*1e3:.1f}us {2*M*N*K/t/1e9:.0f} TF |"
    print(line, flush=True)
print("flags", L.thmr_check_device_flags())
