"""N > 1 path on CPU: two gloo ranks shard a batch by index exactly as bench.py does on GPUs (the decode
itself is stood in for by the oracle here, there is no GPU in this tier)."""
import hashlib
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, out_q):
    for p in (ROOT, os.path.join(ROOT, "video-coding_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist

    import synth
    from hcjpeg import shard
    from oracle import pyoracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_images, rank, world)
    hashes = []
    for i in range(lo, hi):
        jpg = orc.encode(synth.frame(i, 64, 48, 420), 64, 48, 420, 75, restart_interval=2 if i % 2 else 0)
        hashes.append((i, hashlib.sha256(orc.decode(jpg).yuv()).hexdigest()))
    shard.barrier(dist)
    t = shard.max_over_ranks(1.0 + rank, dist)  # the slowest rank defines the time
    allh = shard.gather_objects(hashes, dist)
    if rank == 0:
        out_q.put((t, [h for part in allh for h in part]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    sys.path.insert(0, os.path.join(ROOT, "video-coding_b200"))
    from hcjpeg import shard

    for total in (0, 1, 7, 1024, 8192):
        for world in (1, 2, 3, 4, 8):
            parts = [shard.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(8, 2, 2)


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    from oracle import pyoracle as orc

    orc.build()
    n = 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, hashes = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert t == 2.0
    want = []
    for i in range(n):
        jpg = orc.encode(synth.frame(i, 64, 48, 420), 64, 48, 420, 75, restart_interval=2 if i % 2 else 0)
        want.append((i, hashlib.sha256(orc.decode(jpg).yuv()).hexdigest()))
    assert hashes == want
