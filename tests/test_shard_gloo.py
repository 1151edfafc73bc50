"""Multi-GPU host logic on CPU: world_size-2 `gloo` run of teloscope_b200.shard.ShardedJob.

The scan has no exchange step, so what must hold for N > 1 is host logic only: the
partition is deterministic and balanced, every item is scanned by exactly one rank, and rank 0 gets
the per-item results back in submission order.  The per-rank scan is the CPU oracle here (test
infrastructure; on the GPU box the same ShardedJob is driven with Scanner.scan_batch by
tests/test_gpu_parity.py::test_sharded_job_equals_single_batch)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

from teloscope_b200.shard import ShardedJob, partition_items

HERE = Path(__file__).resolve().parent


def test_partition_is_deterministic_balanced_and_complete():
    rng = np.random.default_rng(3)
    lengths = rng.integers(1, 250_000_000, size=438).tolist()
    for world in (1, 2, 4, 8):
        parts = partition_items(lengths, world)
        again = partition_items(lengths, world)
        assert all(np.array_equal(a, b) for a, b in zip(parts, again))
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(len(lengths)))          # every item exactly once
        assert all(np.all(np.diff(p) > 0) for p in parts if len(p) > 1)      # submission order kept per rank
        loads = [int(np.sum(np.asarray(lengths)[p])) for p in parts]
        assert max(loads) - min(loads) <= max(lengths)                        # LPT bound
    assert [p.tolist() for p in partition_items([5, 5, 5], 2)] == [[0, 2], [1]]   # ties: lower rank, index order
    assert [len(p) for p in partition_items([], 3)] == [0, 0, 0]
    with pytest.raises(ValueError):
        partition_items([1], 0)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _items():
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
    seqs = []
    for k in range(7):
        body = rng.choice(acgt, size=int(rng.integers(800, 6000))).tobytes()
        seqs.append(b"CCCTAA" * (20 * (k % 3)) + body + b"TTAGGG" * (15 * ((k + 1) % 4)))
    seqs.append(b"")                                                          # an empty item travels too
    seqs.append(b"TTAGGG" * 3)
    abs_pos = [1000 * i for i in range(len(seqs))]
    return seqs, abs_pos


class _OracleBatch:
    def __init__(self, orc, seqs, abs_pos):
        self._res = [orc.scan_segment(s.upper(), int(a), False) for s, a in zip(seqs, abs_pos)]

    def item(self, k):
        return self._res[k]


def _worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, str(HERE)); sys.path.insert(0, str(HERE.parent))
    import torch.distributed as dist
    import oracle_lib as ol
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = ol.Oracle(ol.Config(window_size=500, step=500, terminal_limit=2000))
        calls = []

        def scan_fn(seqs, abs_pos, mode):
            calls.append(len(seqs))
            return _OracleBatch(orc, seqs, abs_pos)

        seqs, abs_pos = _items()
        out = ShardedJob(scan_fn, rank, world).run(seqs, abs_pos, 0)
        if rank == 0:
            want = _OracleBatch(orc, seqs, abs_pos)
            ok = len(out) == len(seqs)
            for i in range(len(seqs)):
                w, g = want.item(i), out[i]
                for k in ("terminal", "its", "canonical", "noncanonical", "windows"):
                    ok = ok and np.array_equal(w[k], g[k])
            q.put(("rank0", ok, calls))
        else:
            q.put((f"rank{rank}", out is None, calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_job_world2_gloo_equals_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in procs:
        name, ok, calls = q.get(timeout=240)
        got[name] = (ok, calls)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got["rank0"][0], "gathered results differ from the single-rank scan"
    assert got["rank1"][0], "non-destination ranks return None"
    n_items = len(_items()[0])
    assert sum(c for _, calls in got.values() for c in calls) == n_items       # every item scanned exactly once
