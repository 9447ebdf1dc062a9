"""World-size-2 gloo tests (CPU) for the multi-GPU host logic: weight broadcast and scene sharding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from omnivggt_official_b200 import OmniVGGT
    from omnivggt_official_b200.dist import broadcast_weights, max_over_ranks, shard_scenes
    m = OmniVGGT(img_size=56, embed_dim=128, depth=2, patch_embed="conv", dpt_features=128,
                 dpt_out_channels=(64, 128, 256, 256), dpt_layers=(0, 1, 0, 1), camera_heads=2, camera_trunk_depth=1,
                 init_seed=100 + rank)
    m.randomize_(seed=100 + rank)
    before = torch.cat([p.reshape(-1) for p in m.parameters()]).clone()
    nbytes = broadcast_weights(m, src=0, bucket_bytes=1 << 20)
    after = torch.cat([p.reshape(-1) for p in m.parameters()])
    gathered = [torch.empty_like(after) for _ in range(world)]
    dist.all_gather(gathered, after)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    changed = not torch.equal(before, after)
    mx = max_over_ranks(float(rank + 1), "cpu")
    q.put((rank, same, changed, nbytes, shard_scenes(7, rank, world), mx))
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)                      # identical replicas after the broadcast
    assert not res[0][2] and res[1][2]                 # rank 0 unchanged, rank 1 overwritten
    assert res[0][3] == res[1][3] > 0
    assert res[0][4] == [0, 2, 4, 6] and res[1][4] == [1, 3, 5]
    assert sorted(res[0][4] + res[1][4]) == list(range(7))
    assert res[0][5] == res[1][5] == 2.0


def _cp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from omnivggt_official_b200.context_parallel import ContextParallel
    cp = ContextParallel(torch.device("cpu"))
    out = (rank, cp.rank, cp.world, cp.local_views(8), cp.local_indices([0, 3, 4, 7], 8), cp.local_indices([], 8))
    try:
        cp.local_views(7)
        out += (False,)
    except ValueError:
        out += (True,)
    q.put(out)
    dist.destroy_process_group()


def test_context_parallel_view_partition_world2():
    """Host logic of the context-parallel path (the K / V exchange itself needs peer-mapped GPU memory: tools/cp_check.py,
    tests/test_cp_gpu.py): contiguous, equal view windows in rank order; index lists filtered to the owned views."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_cp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[3] for r in res] == [(0, 4), (4, 4)] and [r[2] for r in res] == [2, 2]
    assert res[0][4] == [0, 3] and res[1][4] == [0, 3]          # scene indices [0,3] / [4,7] relative to the window start
    assert res[0][5] == [] and all(r[6] for r in res)            # 7 views cannot be split over 2 ranks
