"""Where does the CTA-pair GEMM's MMA thread wait?  Cycle counters of the leader CTA (THMR_GEMM_COUNTERS):
w_tempty = waiting for a free accumulator (epilogue too slow), w_full = waiting for operands (TMA feed),
producer w_empty = TMA producer waiting for a free smem stage (MMA too slow).  One launch per shape."""
import os, torch
from tokenhmr_b200._lib import lib, check
L = lib(); dev = torch.device("cuda:0"); torch.manual_seed(0)
cnt = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
os.environ["THMR_GEMM_COUNTERS"] = hex(cnt.data_ptr())
st = lambda: torch.cuda.current_stream().cuda_stream
P = lambda t: None if t is None else t.data_ptr()
M = 12288
for (name, N, K, act, mode) in [("qkv", 3840, 1280, 0, "s16"), ("proj", 1280, 1280, 0, "add"), ("fc1", 5120, 1280, 1, "s16"), ("fc2", 1280, 5120, 0, "add")]:
    A = torch.randn(M, K, device=dev).half(); B = (torch.randn(N, K, device=dev) * 0.03).half()
    bias = torch.randn(N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.float16); x = torch.zeros(M, N, device=dev)
    def run():
        if mode == "s16":
            check(L.thmr_gemm_f16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, P(bias), None, N, act, None, N, P(o16), N, 512, st()))
        else:
            check(L.thmr_gemm_f16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, P(bias), P(x), N, 0, P(x), N, None, N, 512, st()))
    for _ in range(3): run()
    torch.cuda.synchronize(); cnt.zero_(); run(); torch.cuda.synchronize()
    c = cnt.view(148, 16)[0::2].double().cpu()     # leader CTAs
    peer = cnt.view(148, 16)[1::2].double().cpu()
    tot = c[:, 4].mean().item()
    print(f"{name}: mma-thread total {tot:.0f} clk | wait accumulator {c[:,2].mean().item()/tot*100:.1f}% | wait operands {c[:,3].mean().item()/tot*100:.1f}% | "
          f"producer total {c[:,1].mean().item():.0f} wait-free-stage {c[:,0].mean().item()/c[:,1].mean().item()*100:.1f}% (peer {peer[:,0].mean().item()/max(peer[:,1].mean().item(),1)*100:.1f}%)", flush=True)
