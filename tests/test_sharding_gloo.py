"""CPU, world_size 2 over gloo: the N>1 path of the benchmark/driver -- frames are partitioned over ranks, there is
NO data-path collective, only a barrier and a MAX/SUM reduction of scalars.  The per-rank work is done by the CPU
oracle here (no GPU in this container); on the B200 box the same plumbing drives libvp_b200.so (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from vpb200 import shard


def test_partition_covers_every_item_exactly_once():
    for n in (0, 1, 7, 8, 64, 1001):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                s = shard.partition(n, world, r)
                seen += list(s.indices())
                assert len(s) in (n // world, n // world + 1)
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard.partition(4, 2, 2)
    assert [shard.camera_of_rank(r, 8) for r in range(8)] == list(range(8))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (os.path.join(root, "oracle"), os.path.join(root, "vision-processor_b200", "python"), here):
        sys.path.insert(0, p)
    import common
    import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_frames = 7
    mine = shard.partition(n_frames, world, rank)
    orc = O.Oracle("port")
    counters = {}
    for i in mine.indices():  # frame i is the same on every rank: seeded
        p, raw, _ = common.make_case(wq=64, hq=48, seed=100 + i)
        counters[i] = orc.detect(raw, p, want_images=False)["counter"].tolist()
    dist.barrier()
    elapsed = 0.25 if rank == 0 else 0.5          # pretend rank 1 is the slow one
    thr = shard.aggregate_throughput(len(mine), elapsed)
    worst = shard.max_over_ranks(elapsed)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([thr, worst, len(mine)]))
    import json
    json.dump(counters, open(os.path.join(out_dir, f"c{rank}.json"), "w"))
    dist.destroy_process_group()


def test_two_ranks_shard_frames_without_a_collective_on_the_data_path(tmp_path, port):
    import json
    import common
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = {}
    for r in range(world):
        thr, worst, n = np.load(tmp_path / f"r{r}.npy")
        assert worst == 0.5 and thr == 7 / 0.5           # all frames / slowest rank, identical on every rank
        got.update({int(k): v for k, v in json.load(open(tmp_path / f"c{r}.json")).items()})
    assert sorted(got) == list(range(7))
    for i in range(7):                                   # sharded result == single-process result
        p, raw, _ = common.make_case(wq=64, hq=48, seed=100 + i)
        assert got[i] == port.detect(raw, p, want_images=False)["counter"].tolist()
