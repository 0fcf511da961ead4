"""proj-shaped GEMM (M=12288, N=1280, K=1280): which part costs the time?  fp32 reduce-add vs fp32 store vs fp16 store,
CTA-pair 256x256 tiles vs single-CTA 128x256 / 128x128 tiles, vs cuBLAS.  Burst (L2-warm) numbers."""
import torch, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokenhmr_b200._lib import lib, check
L = lib(); dev = torch.device("cuda:0"); torch.manual_seed(0)
st = lambda: torch.cuda.current_stream().cuda_stream
P = lambda t: None if t is None else t.data_ptr()
def gemm(A, B, bias, resid, act, o32, o16, bn):
    M, K = A.shape; N = B.shape[0]
    check(L.thmr_gemm_f16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, P(bias), P(resid), N, act, P(o32), N, P(o16), N, bn, st()))
def timeit(fn, iters=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
M = 12288
for (N, K) in [(1280, 1280), (1280, 2560), (2560, 1280)]:
    A = torch.randn(M, K, device=dev).half(); B = (torch.randn(N, K, device=dev) * 0.03).half()
    bias = torch.randn(N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.float16); x = torch.zeros(M, N, device=dev); y = torch.empty(M, N, device=dev)
    fl = 2.0 * M * N * K
    t = timeit(lambda: torch.matmul(A, B.t(), out=o16)); print(f"N={N} K={K} cublas fp16 out      {t*1e3:7.1f} us {fl/t/1e9:6.0f} TF")
    for bn in (512, 256, 128):
        for name, fn in (("reduce-add", lambda: gemm(A, B, bias, x, 0, x, None, bn)), ("fp32 store", lambda: gemm(A, B, bias, None, 0, y, None, bn)),
                         ("fp16 store", lambda: gemm(A, B, bias, None, 0, None, o16, bn))):
            t = timeit(fn); print(f"N={N} K={K} ours bn={bn:3d} {name:10s} {t*1e3:7.1f} us {fl/t/1e9:6.0f} TF", flush=True)
