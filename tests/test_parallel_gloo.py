"""
N>1 host logic on CPU: world_size-2 gloo run of find_circ2_b200.parallel (record exchange by key hash + ordered junction
gather).  The GPU engine is replaced by a numpy test double with the same method names; the real engine's partition kernel is
covered by tests/test_gpu_parity.py::test_partition_by_key and the NCCL path by `bench.py --gpus N`.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests"))


class NumpyAgg:
    """numpy stand-in for the aggregation half of Engine (partition by key hash / append / records)"""

    def __init__(self, recs):
        self.recs = recs

    def agg_n_records(self):
        return len(self.recs)

    @staticmethod
    def _dest(recs, world):
        k = (recs["chrom"].astype(np.uint64) * np.uint64(1000003) + recs["start"].astype(np.uint64) * np.uint64(10007)
             + recs["end"].astype(np.uint64) * np.uint64(101) + (recs["sk"] & 3).astype(np.uint64))
        return (k % np.uint64(world)).astype(np.int64)

    def agg_partition(self, world, send, stream):
        d = self._dest(self.recs, world)
        order = np.argsort(d, kind="stable")
        raw = np.ascontiguousarray(self.recs[order]).view(np.uint8).reshape(-1)
        send[: raw.size] = torch.from_numpy(raw.copy())
        return np.bincount(d, minlength=world).astype(np.int64)

    def agg_reset(self):
        self.recs = self.recs[:0]

    def agg_append_device(self, n, recv, stream):
        from find_circ2_b200._lib import JREC_DTYPE

        self.recs = recv[: n * 48].numpy().view(JREC_DTYPE).copy()


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from find_circ2_b200 import parallel
    from find_circ2_b200._lib import JREC_DTYPE, JUNCTION_DTYPE

    rng = np.random.default_rng(100 + rank)
    n = 5000 + 777 * rank
    recs = np.zeros(n, dtype=JREC_DTYPE)
    recs["chrom"] = rng.integers(0, 3, n)
    recs["start"] = rng.integers(0, 200, n)
    recs["end"] = recs["start"] + rng.integers(1, 50, n)
    recs["sk"] = rng.integers(0, 4, n) | (1 << 8)
    recs["idx"] = (rank << 40) + np.arange(n)
    eng = NumpyAgg(recs.copy())
    sent, got = parallel.exchange_records(eng, dist, torch.device("cpu"), 0)
    mine = eng.recs
    # every record I hold hashes to me; stream order (idx ascending) survived the exchange
    assert (NumpyAgg._dest(mine, world) == rank).all()
    assert (np.diff(mine["idx"].astype(np.int64)) > 0).all()
    np.save(os.path.join(tmp, "recs_%d.npy" % rank), mine)
    np.save(os.path.join(tmp, "orig_%d.npy" % rank), recs)
    # ordered gather of per-rank junction tables
    j = np.zeros(3 + rank, dtype=JUNCTION_DTYPE)
    j["first_idx"] = (np.arange(len(j)) * 2 + rank).astype(np.uint64)
    j["chrom"] = rank
    allj = parallel.gather_junctions(j, dist, torch.device("cpu"))
    if rank == 0:
        assert len(allj) == sum(3 + r for r in range(world))
        assert (np.diff(allj["first_idx"].astype(np.int64)) >= 0).all()
    else:
        assert allj is None
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_and_gather_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(os.path.join(tmp_path, "recs_%d.npy" % r)) for r in range(world)])
    orig = np.concatenate([np.load(os.path.join(tmp_path, "orig_%d.npy" % r)) for r in range(world)])
    assert sorted(got["idx"].tolist()) == sorted(orig["idx"].tolist())
