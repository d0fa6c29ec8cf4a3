"""Multi-rank host logic on CPU: world_size-2 gloo processes (the N>1 path of bench.py / InferenceRunner)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_lpt_partition_balances_and_covers():
    from pst.distributed import lpt_partition, structure_cost

    rng = np.random.default_rng(0)
    lengths = list(rng.integers(64, 2048, 500))
    for world in (1, 2, 4, 8):
        shards = lpt_partition(lengths, world)
        assert sorted(i for s in shards for i in s) == list(range(500))
        loads = [sum(structure_cost(lengths[i]) for i in s) for s in shards]
        assert max(loads) / (sum(loads) / world) < 1.02
    assert lpt_partition(lengths, 4) == lpt_partition(lengths, 4)


def _worker(rank, world, port, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "protein-structure-tokenizer_b200"))
    import torch.distributed as dist
    from pst.distributed import tokenize_sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [64 + 7 * i for i in range(23)]

    def fake_tokenize(indices):  # stands in for StructureTokenizer.tokenize on this rank's shard
        return [(np.arange(lengths[i] // 2, dtype=np.uint32) * 31 + i) % 4096 for i in indices]

    out = tokenize_sharded(lengths, fake_tokenize, rank, world)
    if rank == 0:
        ok = all(np.array_equal(out[i], (np.arange(lengths[i] // 2, dtype=np.uint32) * 31 + i) % 4096) for i in range(len(lengths)))
        with open(os.path.join(out_dir, "result"), "w") as fh:
            fh.write("ok" if ok and len(out) == len(lengths) else "bad")
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_shard_and_gather_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "result").read_text() == "ok"


def test_single_rank_path_needs_no_process_group():
    from pst.distributed import tokenize_sharded

    lengths = [50, 80, 64]
    out = tokenize_sharded(lengths, lambda idx: [np.full(lengths[i], i, np.uint32) for i in idx], 0, 1)
    assert [o.shape[0] for o in out] == lengths and all((o == i).all() for i, o in enumerate(out))
