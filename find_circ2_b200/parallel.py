"""
Multi-GPU plumbing: one process per GPU, anchor pairs sharded by rank, ONE exchange step.

The scan needs no communication (the genome is replicated per GPU).  Junction aggregation is a keyed reduce:
every rank turns its accepted spans into 48-byte records (fc_jrec), partitions them by hash(key) % world on the
device (fc_agg_partition, stable -> stream order survives), and a single all-to-all over NCCL/NVLink moves each
record to the rank that owns its key.  Ranks hold contiguous, ascending ranges of the input stream, and the
all-to-all output is ordered by source rank, so every receiver sees its records in global stream order --
which is what the discovery-order junction names and the sequential float sums of the reference depend on
(find_circ.py:684-686, 544).  Raw records (not partial sums) are exchanged because n_uniq / n_frags are distinct
counts (find_circ.py:584-590).  The reference's own multi-sample story has the same shape: independent runs
merged by merge_bed.py (merge_bed.py:117-130).
"""
from __future__ import annotations

import numpy as np

REC_BYTES = 48


_BUF = {}


def _buffer(tag, nbytes, dev):
    """persistent, growing byte buffers (no allocation inside the exchange once warmed up)"""
    import torch

    key = (tag, str(dev))
    t = _BUF.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=dev)
        _BUF[key] = t
    return t


def bind_to_gpu_numa(device_index: int) -> dict:
    """pin this process (and, by first touch, the pinned buffers it allocates afterwards) to the NUMA node its GPU hangs
    off: with one process per GPU the host-to-device copies of all ranks otherwise funnel through node 0's memory.
    Returns what was done ({} when the topology cannot be read)."""
    import os

    try:
        import torch

        pr = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception:
        bus = None
    try:
        if bus is None:
            import subprocess

            bus = subprocess.run(["nvidia-smi", "-i", str(device_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return {}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        return {}


def exchange_records(eng, dist, dev, stream=0, upper_bound=None):
    """hash-partition this rank's junction records and swap them with the other ranks; afterwards the engine's
    aggregator holds exactly the records whose keys this rank owns.  Returns (sent_bytes, received_bytes).
    Host synchronisations: two inside the partition (record count, per-destination counts) and one for the counts
    received from the peers."""
    import torch

    world = dist.get_world_size()
    if world == 1:
        return 0, 0
    ub = upper_bound if upper_bound is not None else eng.agg_n_records()
    send = _buffer("send", max(ub, 1) * REC_BYTES, dev)
    counts = eng.agg_partition(world, send, stream)  # int64[world], records per destination
    n = int(counts.sum())
    c_send = torch.from_numpy(np.ascontiguousarray(counts, dtype=np.int64)).to(dev, non_blocking=True)
    c_recv = torch.empty_like(c_send)
    dist.all_to_all_single(c_recv, c_send)
    recv_counts = c_recv.cpu().numpy()
    n_recv = int(recv_counts.sum())
    recv = _buffer("recv", max(n_recv, 1) * REC_BYTES, dev)
    dist.all_to_all_single(
        recv[: n_recv * REC_BYTES],
        send[: n * REC_BYTES],
        output_split_sizes=[int(c) * REC_BYTES for c in recv_counts],
        input_split_sizes=[int(c) * REC_BYTES for c in counts],
    )
    if hasattr(eng, "agg_replace_device"):
        eng.agg_replace_device(n_recv, recv, stream)
    else:
        eng.agg_reset()
        eng.agg_append_device(n_recv, recv, stream)
    return int(n * REC_BYTES), int(n_recv * REC_BYTES)


def p2p_setup(eng, dist, dev, capacity_records: int) -> bool:
    """export this rank's record buffer / counter with CUDA IPC, collect everybody's handles, open the peers'.
    Returns False (and leaves the engine on the NCCL all-to-all path) when IPC is not available."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    ok = 1
    try:
        handle = eng.p2p_export(capacity_records)
    except Exception:
        handle = np.zeros(128, dtype=np.uint8)
        ok = 0
    mine = torch.from_numpy(np.concatenate([handle, np.frombuffer(np.int64(capacity_records).tobytes(), dtype=np.uint8),
                                            np.array([ok], dtype=np.uint8)])).to(dev)
    allh = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    allh = np.stack([t.cpu().numpy() for t in allh])
    if not allh[:, 136].all():
        return False
    caps = np.array([np.frombuffer(allh[r, 128:136].tobytes(), dtype=np.int64)[0] for r in range(world)], dtype=np.int64)
    try:
        eng.p2p_connect(world, rank, allh[:, :128].copy(), caps)
        good = 1
    except Exception:
        good = 0
    flag = torch.tensor([good], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def stream_barrier(dist, dev, eng=None, stream=0, _cache={}):
    """a barrier that orders the STREAMS of all ranks without blocking the host.  With an engine whose peers are connected
    (p2p_setup) it is the library's own kernel over peer memory (fc_p2p_barrier, a few microseconds; it also ENDS a step of
    the peer emit: exactly one per step, see include/findcirc_b200.h); otherwise a one-element NCCL all-reduce"""
    if eng is not None:
        eng.p2p_barrier(stream)
        return
    import torch

    t = _cache.get(str(dev))
    if t is None:
        t = _cache[str(dev)] = torch.zeros(1, dtype=torch.int32, device=dev)
    dist.all_reduce(t)


def gather_junctions(junctions: np.ndarray, dist, dev):
    """ordered gather of every rank's junction table to rank 0: rows sorted by first_idx = discovery order, which is what
    the junction names count (find_circ.py:684-686).  Tables travel device to device (NCCL on GPUs, gloo on CPU) and are
    merged by one sort on rank 0's device; the other ranks get None."""
    import torch

    world = dist.get_world_size()
    if world == 1:
        return junctions
    rank = dist.get_rank()
    row = junctions.dtype.itemsize  # 64
    raw = torch.from_numpy(np.ascontiguousarray(junctions).view(np.uint8).reshape(-1).copy())
    n_local = torch.tensor([len(junctions)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    buf = torch.zeros(mx * row, dtype=torch.uint8, device=dev)
    buf[: raw.numel()] = raw.to(dev, non_blocking=True)
    bufs = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, bufs, dst=0)
    if rank != 0:
        return None
    allj = torch.cat([b[: s * row] for b, s in zip(bufs, sizes)]).view(torch.int64).reshape(-1, row // 8)
    # column 2 = first_idx (positions in the input stream, far below 2^63; equal values do not occur across ranks)
    order = torch.sort(allj[:, 2], stable=True).indices
    return allj[order].contiguous().view(torch.uint8).cpu().numpy().reshape(-1).view(junctions.dtype)
