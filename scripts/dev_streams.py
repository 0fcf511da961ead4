"""How many step graphs in flight pay?  Device-resident replay on n streams and the end-to-end pipeline at several
(depth, streams) settings, same engine (concurrent=True), bs=64 release model.  One-stream numbers of the engine WITH
stream-K (the default engine) are printed next to them."""
import sys, time, json
import torch
from tokenhmr_b200 import synth
from tokenhmr_b200.config import release_config
from tokenhmr_b200.engine import TokenHMREngine, TokenHMRPipeline

dev = torch.device("cuda:0")
cfg = release_config()
sd, smpl = synth.make_state_dict(cfg), synth.make_smpl(cfg)
B, K = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 24
img_host = synth.make_images(B, cfg, seed=0).pin_memory()
img_dev = img_host.to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def resident(model, n):
    ss = [torch.cuda.Stream(dev) for _ in range(n)]
    def step(i):
        with torch.cuda.stream(ss[i % n]):
            model.forward({"img": img_dev}, alias_outputs=True, slot=i % n)
    for i in range(2 * n + 2):
        step(i)
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        main = torch.cuda.current_stream()
        e0.record(main)
        for s in ss:
            s.wait_event(e0)
        for i in range(K):
            step(i)
        for s in ss:
            ev = torch.cuda.Event(); ev.record(s); main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        best = ms if best is None else min(best, ms)
    return best


def e2e(model, depth, streams):
    pipe = TokenHMRPipeline(model, depth=depth, streams=streams)
    def run(n):
        tickets = []
        done = 0
        for _ in range(n):
            tickets.append(pipe.submit({"img": img_host}))
            if len(tickets) - done >= depth:
                pipe.result(tickets[done]); done += 1
        while done < len(tickets):
            out = pipe.result(tickets[done]); done += 1
        return out
    run(2 * depth)
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        e0.record(pipe.copy_stream)
        run(K)
        e1.record(pipe.join())
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        best = ms if best is None else min(best, ms)
    return best


out = {}
plain = TokenHMREngine(cfg, sd, smpl, device=dev, use_cuda_graph=True)
out["resident plain engine 1 stream"] = resident(plain, 1)
out["e2e plain engine depth2 streams1"] = e2e(plain, 2, 1)
del plain
torch.cuda.empty_cache()
conc = TokenHMREngine(cfg, sd, smpl, device=dev, use_cuda_graph=True, concurrent=True, max_cached_shapes=16)
for n in [int(a) for a in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1,2,3,4".split(","))]:
    out[f"resident concurrent engine {n} streams"] = resident(conc, n)
combos = [tuple(int(v) for v in c.split("x")) for c in (sys.argv[3].split(",") if len(sys.argv) > 3 else
                                                          "2x2,3x2,4x2,3x3,6x3,4x4".split(","))]
for depth, streams in combos:
    out[f"e2e concurrent engine depth{depth} streams{streams}"] = e2e(conc, depth, streams)
for k, v in out.items():
    print(f"{k:48s} {v:7.3f} ms/step  {B * 1e3 / v:7.0f} images/s")
print(json.dumps(out))
