import torch, os, sys
dev = torch.device("cuda:0"); torch.manual_seed(0)
cnt = torch.zeros(148 * 16, device=dev, dtype=torch.int64)
os.environ["THMR_GEMM_COUNTERS"] = hex(cnt.data_ptr())
from tokenhmr_b200._lib import lib, check
L = lib()
st = lambda: torch.cuda.current_stream().cuda_stream
M, N, K = 12288, int(sys.argv[1]) if len(sys.argv) > 1 else 3840, int(sys.argv[2]) if len(sys.argv) > 2 else 1280
A = torch.randn(M, K, device=dev).half(); B = (torch.randn(N, K, device=dev) * 0.03).half()
bias = torch.randn(N, device=dev); o16 = torch.empty(M, N, device=dev, dtype=torch.float16)
x32 = torch.zeros(M, N, device=dev)
if os.environ.get('ADD'):
    fn = lambda: check(L.thmr_gemm_f16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, bias.data_ptr(), x32.data_ptr(), N, 0, x32.data_ptr(), N, None, N, int(os.environ.get('BN', '256')), st()))
else:
    fn = lambda: check(L.thmr_gemm_f16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, bias.data_ptr(), None, N, 0, None, N, o16.data_ptr(), N, int(os.environ.get('BN', '256')), st()))
for _ in range(10): fn()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
for _ in range(100): fn()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 100
cv = cnt.view(148, 16).float(); c = (cv[::2] if os.environ.get('BN') == '512' else cv).mean(0).tolist(); c1 = cv[1::2].mean(0).tolist(); print('peer producer wait_empty', c1[0], 'total', c1[1])
print(f"DBG={os.environ.get('THMR_GEMM_DBG')}: {t*1e3:.1f}us {2*M*N*K/t/1e9:.0f} TF | cycles: producer wait_empty {c[0]:.0f} / total {c[1]:.0f} | mma wait_tempty {c[2]:.0f} wait_full {c[3]:.0f} / total {c[4]:.0f} | epi wait_tfull {c[5]:.0f} wait_store_read {c[7]:.0f} / total {c[6]:.0f} | epi detail: ldtm {c[8]:.0f} alu {c[9]:.0f} sts(incl wait_read) {c[10]:.0f} fence {c[11]:.0f}")
