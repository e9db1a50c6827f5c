"""CPU tier, world_size 2 over gloo: the N > 1 host logic (block sharding of independent streams + the optional
edge gather of packets).  Per-rank compute is the CPU oracle here; on GPUs it is the Context of that rank."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import MODEL_DIR, ROOT
from lyra_b200 import sharding


def test_shard_ranges_cover_everything():
    for n, w in ((32768, 8), (4096, 1), (10, 4), (7, 8)):
        seen = []
        for r in range(w):
            first, cnt = sharding.shard_range(n, w, r)
            seen += list(range(first, first + cnt))
            for g in range(first, first + cnt):
                assert sharding.owner(g, n, w) == (r, g - first)
        assert seen == list(range(n))


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    first, cnt = sharding.shard_range(n, world, rank)
    rng = np.random.default_rng(123)
    pcm = rng.integers(-8192, 8192, size=(2, n, 320), dtype=np.int16)      # same on every rank
    codecs = [O.Codec(MODEL_DIR) for _ in range(cnt)]
    gathered = []
    for f in range(2):
        local = np.stack([np.frombuffer(codecs[k].encode(pcm[f, first + k], 64)[0], dtype=np.uint8) for k in range(cnt)])
        gathered.append(sharding.gather_packets(local, n))
    if rank == 0:
        np.save(out_path, np.stack(gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_encode_matches_single_process(tmp_path, oracle):
    n, world, port = 5, 2, 29500 + (os.getpid() % 2000)
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, port, n, out), nprocs=world, join=True)
    got = np.load(out)
    rng = np.random.default_rng(123)
    pcm = rng.integers(-8192, 8192, size=(2, n, 320), dtype=np.int16)
    codecs = [oracle.Codec(MODEL_DIR) for _ in range(n)]
    for f in range(2):
        for k in range(n):
            assert bytes(got[f, k]) == codecs[k].encode(pcm[f, k], 64)[0]
