import torch, os, sys
dev = torch.device("cuda:0"); torch.manual_seed(0)
cnt = torch.zeros(148 * 8, device=dev, dtype=torch.int64)
os.environ["THMR_GEMM_COUNTERS"] = hex(cnt.data_ptr())
from tokenhmr_b200._lib import lib, check
L = lib()
st = lambda: torch.cuda.current_stream().cuda_stream
M, N, K = 12288, int(sys.argv[1]) if len(sys.argv) > 1 else 3840, int(sys.argv[2]) if len(sys.argv) > 2 else 1280
A = torch.randn(M, K, device=dev).half(); B = (torch.randn(N, K, device=dev) * 0.03).half()
bias = torch.randn(N, device=dev); o16 = torch.empty(M, N, device=dev, dtype=torch.float16)
fn = lambda: check(L.thmr_gemm_f16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, bias.data_ptr(), None, N, 0, None, N, o16.data_ptr(), N, 256, st()))
for _ in range(10): fn()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
for _ in range(100): fn()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 100
c = cnt.view(148, 8).float().mean(0).tolist()
print(f"DBG={os.environ.get('THMR_GEMM_DBG')}: {t*1e3:.1f}us {2*M*N*K/t/1e9:.0f} TF | cycles: producer wait_empty {c[0]:.0f} / total {c[1]:.0f} | mma wait_tempty {c[2]:.0f} wait_full {c[3]:.0f} / total {c[4]:.0f} | epi wait_tfull {c[5]:.0f} / total {c[6]:.0f}")
