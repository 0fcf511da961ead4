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


def _worker(rank, world, port, q, global_batch):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(0)
    global_img = torch.randn(global_batch, 3, 4, 4, generator=gen)
    lo, hi = shard_range(global_batch, rank, world)
    sharded = ShardedTokenHMR(lambda batch: _fake_outputs(batch["img"]))
    assert sharded.transport == "torch"
    ok = True
    for _ in range(2):                 # second call reuses the cached shard sizes and buffers
        got = sharded({"img": global_img[lo:hi]})
        want = _fake_outputs(global_img)  # what one process would produce for the whole batch
        ok &= all(torch.equal(got[k], want[k]) for k in GATHER_KEYS if k in want)
        ok &= all(torch.equal(got["pred_smpl_params"][k], want["pred_smpl_params"][k]) for k in want["pred_smpl_params"])
    q.put((rank, ok, tuple(got["pred_vertices"].shape), sharded.sizes(hi - lo)))
    dist.destroy_process_group()


def _run(global_batch, world=2):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, global_batch)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    return res


def test_two_rank_allgather_reassembles_global_batch():
    res = _run(6)
    assert all(ok for _, ok, _, _ in res), res
    assert all(shape == (6, 20, 3) for _, _, shape, _ in res)


def test_uneven_shards_are_padded_and_trimmed():
    """7 images over 2 ranks = 4 + 3: every rank sends 4 rows per field, the padding row is dropped after the exchange."""
    res = _run(7)
    assert all(ok for _, ok, _, _ in res), res
    assert all(shape == (7, 20, 3) and sizes == [4, 3] for _, _, shape, sizes in res)


def test_shard_ranges_cover_the_batch():
    from tokenhmr_b200.dist import shard_sizes, trim
    for gb in (1, 7, 64, 65, 512):
        for world in (1, 2, 3, 8):
            rs = [shard_range(gb, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == gb and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert shard_sizes(gb, world) == [hi - lo for lo, hi in rs]
    x = torch.arange(8.0).view(8, 1)          # 2 ranks x 4 rows, sizes 4 + 3
    assert trim({"x": x}, 4, [4, 3])["x"].flatten().tolist() == [0, 1, 2, 3, 4, 5, 6]
    assert trim({"x": x}, 4, [4, 4])["x"] is x
