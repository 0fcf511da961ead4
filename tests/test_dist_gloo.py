"""World-size-2 gloo test of the sharding + packed all-gather host logic (CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tokenhmr_b200.dist import GATHER_KEYS, ShardedTokenHMR, shard_range


def _fake_outputs(img: torch.Tensor) -> dict:
    """Deterministic per-image 'outputs' so that gathered results can be checked against the global batch."""
    B = img.shape[0]
    s = img.reshape(B, -1).sum(1)
    f = lambda *shape: s.view(B, *([1] * len(shape))).expand(B, *shape).clone() + torch.arange(
        int(torch.tensor(shape).prod())).view(1, *shape).float()
    return {"pred_vertices": f(20, 3), "pred_keypoints_3d": f(44, 3), "pred_keypoints_2d": f(44, 2), "pred_cam": f(3),
            "pred_cam_t": f(3), "focal_length": f(2),
            "pred_smpl_params": {"global_orient": f(1, 3, 3), "body_pose": f(23, 3, 3), "betas": f(10)}}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(0)
    global_img = torch.randn(6, 3, 4, 4, generator=gen)
    lo, hi = shard_range(6, rank, world)
    sharded = ShardedTokenHMR(lambda batch: _fake_outputs(batch["img"]))
    got = sharded({"img": global_img[lo:hi]})
    want = _fake_outputs(global_img)  # what one process would produce for the whole batch
    ok = all(torch.equal(got[k], want[k]) for k in GATHER_KEYS if k in want)
    ok &= all(torch.equal(got["pred_smpl_params"][k], want["pred_smpl_params"][k]) for k in want["pred_smpl_params"])
    q.put((rank, ok, tuple(got["pred_vertices"].shape)))
    dist.destroy_process_group()


def test_two_rank_allgather_reassembles_global_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (6, 20, 3) for _, _, shape in res)
