#!/usr/bin/env python3
"""
bench.py -- anchor pairs / second through the breakpoint scan + junction merge (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5]     this repo's CUDA path
  python bench.py --impl reference ...                                       the reference algorithm on the host cores (oracle port)

Default workload = BASELINE.json configs[2], the configuration the metric's "1/2/4/8 B200" is quoted on: hg19-sized synthetic
genome (3.1 Gb, replicated per GPU), 100k planted circRNAs, 50 M anchor pairs from 2x100-nt mate pairs in the WHOLE job,
sharded over the ranks by contiguous ranges of the input stream (strong scaling).  A step is one pass of the hot path over
the rank's shard, resident in HBM as one batch.
  value     pairs/s, batch resident in HBM: scan kernel that writes the junction records on the way (N>1: straight into the
            owner ranks' buffers over NVLink), one barrier, aggregation (accumulate, rank, finish)
  e2e       the same through the host-buffer C-ABI call: pinned host SoA in, per-pair hits and the junction table out,
            copies inside the timed region
  roofline  the kernel with the longest launch in the step, roofline_other the second one.  scan: 82 (l=64) / 119 (l=114)
            algorithmic bytes per pair / CUDA-event time of the scan+emit kernel; aggregation: 48 B per record +
            64 B per junction / the library's own CUDA events around the accumulate kernel (fc_agg_get_timing)
  parity_checked  N=1: the whole drop-in (both ingest paths) on a prefix of the same workload against the oracle, all five
            outputs; N>1: the owners' junction tables of a peer-memory step, gathered to rank 0, against one context
            aggregating the union
  cpu_baseline    oracle (python restatement of find_circ.py, one numpy compare per split position) on a bounded prefix
  other_configs   (N=1) configs[1], configs[3] at 1/2/3 % errors and the per-GPU slice of configs[4], device-resident lines
Inputs are far larger than L2 (config 3: 3.8 GB of batch columns, 3.2 GB of genome tiles); configs whose batch fits L2 are
flushed (256 MiB memset) before every timed step.  Workloads: bench_workload.py (seeded, device independent).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "anchor pairs/sec (breakpoint scan + junction merge)"


def bytes_per_pair(cfg):
    l = cfg.read_len - 2 * (cfg.asize - cfg.margin)
    return 16 + (l + 3) // 4 + 2 * ((l + 2 + 3) // 4) + 16  # 82 at l = 64, 119 at l = 114


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU side (oracle)
def sample_sam(spec, cols):
    """SAM text lines (header first) of a prefix of the workload: BWA-MEM style two-segment records, mates flagged"""
    import bench_workload as W
    from find_circ2_b200 import synth

    pt = W.pair_table(cols)

    class G(object):
        names = spec.names
        sizes = [int(x) for x in spec.sizes]

    lines = synth.sam_header(G).splitlines(True)
    paired = spec.cfg.paired
    for i in range(len(pt)):
        flag = (0x41 if i % 2 == 0 else 0x81) if paired else 0
        lines += synth.bwa_records_for_pair(G, pt, i, "r%d" % int(pt.name_id[i]), base_flag=flag)
    return lines


def oracle_genome(spec, cols):
    """the oracle's genome object over the stretches of the (never materialised) genome that the sample touches"""
    import bench_workload as W
    from oracle import find_circ_oracle as O

    og = O.Genome.__new__(O.Genome)
    og.names = list(spec.names)
    og.seqs = W.sparse_genome(spec, cols)
    return og


_CPU_STATE = None  # (oracle genome, @SQ names, options): inherited by forked workers instead of being pickled


def _cpu_shard(recs):
    from oracle import find_circ_oracle as O

    og, names, opt = _CPU_STATE
    t0 = time.perf_counter()
    r = O.Run(og, names, opt)
    r.process(recs)
    out = r.outputs()
    return time.perf_counter() - t0, out


def _cpu_shard_time(recs):
    return _cpu_shard(recs)[0]


def oracle_options(cfg):
    from oracle import find_circ_oracle as O

    return O.Options(asize=cfg.asize, margin=cfg.margin, maxdist=cfg.maxdist, min_uniq_qual=cfg.min_uniq, name="bench",
                     halfunique=cfg.halfunique, report_nobridges=cfg.report_nobridges)


def cpu_run(spec, cols, lines, procs=1):
    """scan + aggregation of the oracle on the sample; SAM decoding and the genome stretches are ready before the clock starts.
    Returns (seconds, outputs of the single-process run or None)"""
    global _CPU_STATE
    from oracle import find_circ_oracle as O

    names, recs = O.read_sam(lines)
    recs = list(recs)
    _CPU_STATE = (oracle_genome(spec, cols), names, oracle_options(spec.cfg))
    if procs <= 1:
        return _cpu_shard(recs)
    import multiprocessing as mp

    # shard by fragment: 2 records per read, 4 per mate pair -- cut on multiples of 4
    per = (len(recs) // 4 + procs - 1) // procs * 4
    shards = [recs[k: k + per] for k in range(0, len(recs), per)]
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(len, shards[:1])  # workers are up before the clock starts
        t0 = time.perf_counter()
        pool.map(_cpu_shard_time, shards, chunksize=1)
        dt = time.perf_counter() - t0
    return dt, None


def run_reference(args):
    import bench_workload as W

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = W.configs()[args.config]
    cores = os.cpu_count() or 1
    spec = W.Spec(cfg)
    # a bounded prefix of the workload per step, sized so that the whole run takes about two minutes: the oracle scans
    # ~1.7 k pairs/s per core (one numpy compare per split position, as the reference does)
    n_sample = int(1500 * cores * 120 / (args.steps + 1 + (1 if args.warmup else 0)))
    n_sample = max(4000, min(n_sample, 2000000, args.cpu_sample * cores * 4)) // 4 * 4
    cols = W.make_pairs(spec, 0, n_sample, "cpu")
    lines = sample_sam(spec, cols)
    if args.warmup:
        cpu_run(spec, cols, lines[: 2000 + len(spec.names) + 3], 1)
    secs = [cpu_run(spec, cols, lines, cores)[0] for _ in range(args.steps)]
    v = float(np.mean([n_sample / s for s in secs]))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": cfg.scaling,
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": {"workload": cfg.workload},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": "the first %d pairs of the workload per step, sharded %d ways by fragment with multiprocessing; SAM decode untimed" % (n_sample, cores)},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU side
class Shard(object):
    """one rank's part of a workload, resident on the device in the layout the library scans"""

    def __init__(self, eng, spec, i0, i1, dev, flat, error_rate=None):
        import torch

        import bench_workload as W

        cfg = spec.cfg
        cols = W.make_pairs(spec, i0, i1, dev, flat=flat, error_rate=error_rate)
        s = W.soa(cols, cfg)
        del cols
        self.n = n = int(s["row"].numel())
        self.n_in = i1 - i0
        self.max_l = int(s["l"].max().item()) if n else 0
        self.n_words = max(1, (self.max_l + 31) // 32)
        self.stride = int(s["internal"].shape[1])
        stream = torch.cuda.current_stream().cuda_stream
        s["read_hash"] = torch.empty(n, dtype=torch.int64, device=dev)
        eng.hash_reads_device(s["reads"], cfg.read_len, s["read_hash"], stream)
        nm = s["name_id"]
        s["qname_hash"] = (nm * (0x9E3779B97F4A7C15 - (1 << 64))) ^ ((nm >> 7) & ((1 << 57) - 1))
        torch.cuda.synchronize()
        # rows of one fragment (same read name) are neighbours: how many come before / after each row (fc_batch descriptors,
        # bits 28-31); read names are unique per fragment in this workload, so the scan kernel may settle n_frags locally
        back = torch.zeros(n, dtype=torch.int32, device=dev)
        fwd = torch.zeros(n, dtype=torch.int32, device=dev)
        for k in (1, 2, 3, 4):
            if n > k:
                eq = (nm[k:] == nm[:-k]).to(torch.int32)
                back[k:] += eq
                fwd[:-k] += eq
        frag = torch.where((back > 3) | (fwd > 3) | (back + fwd > 3), torch.full_like(back, 15), back | (fwd << 2)).to(torch.uint8)
        del s["reads"], s["name_id"], s["row"], back, fwd
        self.planes = torch.zeros(3 * self.n_words * n, dtype=torch.int32, device=dev)
        eng.pack_reads(s["internal"], self.stride, s["l"], self.n_words, self.planes, s["flags"], stream)
        torch.cuda.synchronize()
        self.d = s
        self.hits = torch.zeros(n * 4, dtype=torch.int32, device=dev)
        self.pairs = self._pairs(eng, n)
        self.pay = (s["wden"], s["q_a"], s["q_b"], s["read_hash"], s["qname_hash"])
        # the packed form the scan kernels read (what a native ingest writes directly): 16-byte descriptors, read rows, q words
        self.nw = eng.batch_words(self.max_l)
        self.meta = torch.zeros(n * 4, dtype=torch.int32, device=dev)
        self.rows = torch.zeros(n * 2 * self.nw, dtype=torch.int32, device=dev)
        self.rn_rows = torch.zeros(n * self.nw, dtype=torch.int32, device=dev)
        self.q = torch.zeros(n, dtype=torch.int32, device=dev)
        self.batch = eng.pack_batch(self.pairs, self.meta, self.rows, self.rn_rows, s["wden"], s["q_a"], s["q_b"], self.q, frag, stream)
        torch.cuda.synchronize()

    def batch_prefix(self, m):
        from find_circ2_b200._lib import Batch

        b = self.batch
        return Batch(m, b.d_meta, b.d_reads, b.d_rn, b.n_words, b.max_l)

    def _pairs(self, eng, m):
        """fc_pairs over the first m rows (the planes keep the stride of the whole shard)"""
        from find_circ2_b200._lib import Pairs, ptr

        d = self.d
        base, step = ptr(self.planes), 4 * self.n * self.n_words
        return Pairs(m, ptr(d["chrom"]), ptr(d["a_start"]), ptr(d["b_end"]), ptr(d["l"]), ptr(d["flags"]), base, base + step,
                     base + 2 * step, self.n_words, self.max_l, self.n)


def cli_argv(cfg):
    return (["-G", "unused", "-a", str(cfg.asize), "-m", str(cfg.margin), "-d", str(cfg.maxdist), "-n", "bench", "--min-uniq-qual",
             str(cfg.min_uniq)] + (["--halfuniq"] if cfg.halfunique else []) + (["--report_nobridge"] if cfg.report_nobridges else []))


def dropin_parity(eng, spec, cfg, n_sample, large=True, error_rate=None):
    """the whole drop-in (SAM text -> native / python ingest -> GPU -> five outputs) on a prefix of the workload against the
    oracle; also what `ingest` (host decoding rate) and `cpu_baseline` are taken from"""
    import bench_workload as W
    from find_circ2_b200 import cli
    from oracle import find_circ_oracle as O

    cols = W.make_pairs(spec, 0, n_sample, "cpu", error_rate=error_rate)
    lines = sample_sam(spec, cols)
    cpu_s, want = cpu_run(spec, cols, lines, 1)
    res = {}
    ok = True
    with tempfile.TemporaryDirectory() as tmp:
        sam = os.path.join(tmp, "sample.sam")
        with open(sam, "w") as fh:
            fh.writelines(lines)
        for tag, native in (("python", False), ("native", True)):
            opt = cli.parse_args(cli_argv(cfg))[0]
            out = cli.run_to_strings(opt, sam, engine=eng, native=native)
            same = (O.canonical_bed(out["circ"]) == O.canonical_bed(want.circ_bed) and O.canonical_bed(out["lin"]) == O.canonical_bed(want.lin_bed)
                    and out["reads"] == want.reads_fastq and O.canonical_multi(out["multi"]) == O.canonical_multi(want.multi_events)
                    and out["counters"] == want.counters)
            ok = ok and same
            host_s = out["seconds_total"] - out["seconds_gpu_calls"]
            res[tag] = {"pairs_per_s": n_sample / host_s, "seconds_host": host_s, "seconds_gpu_calls": out["seconds_gpu_calls"], "equals_oracle": same}
        # host ingest at a size where fixed costs vanish: the SAM text of the sample (the workload's own fragments: mate
        # pairs with both mates spliced in configs[2]) replicated under fresh read names
        if not large:
            n_rows = len(want.circ_bed.splitlines()) + len(want.lin_bed.splitlines()) - 2
            return ({"ok": ok and n_rows > 0, "against": "oracle, all five outputs of the drop-in, native and python ingest", "pairs": n_sample,
                     "junction_rows": n_rows}, None, None)

        body = "".join(lines).encode()
        head_end = body.find(b"\nr") + 1  # (header lines start with @, read names with r)
        header, recs = body[:head_end], b"\n" + body[head_end:]
        big = os.path.join(tmp, "ingest.sam")
        copies = 100
        with open(big, "wb") as fh:
            fh.write(header)
            for k in range(copies):
                fh.write(recs.replace(b"\nr", b"\nc%dr" % k)[1:])
        opt = cli.parse_args(cli_argv(cfg))[0]
        out = cli.run_to_strings(opt, big, engine=eng, native=True)
        host_s = out["seconds_ingest_and_scan"] - out["seconds_gpu_calls"]
        threads = opt.ingest_threads or max(1, min(16, os.cpu_count() or 1))
        res["native_large"] = {"pairs_per_s": n_sample * copies / host_s, "reads": n_sample * copies, "sam_bytes": os.path.getsize(big),
                               "parser_threads": threads, "seconds_host": host_s, "seconds_gpu_calls": out["seconds_gpu_calls"],
                               "seconds_main_thread": out.get("seconds_ingest_stages"),
                               "seconds_writers": out["seconds_total"] - out["seconds_ingest_and_scan"],
                               "seconds_writer_stages": out.get("seconds_writer_stages"),
                               "seconds_wall": out["seconds_total"], "junction_rows": out["circ"].count("\n") + out["lin"].count("\n") - 2}
    n_rows = len(want.circ_bed.splitlines()) + len(want.lin_bed.splitlines()) - 2
    parity = {"ok": ok and n_rows > 0, "against": "oracle (oracle/find_circ_oracle.py), all five outputs of the drop-in, native and python ingest",
              "pairs": n_sample, "junction_rows": n_rows}
    ingest = {"value": res["native_large"]["pairs_per_s"], "unit": "pairs/s",
              "kind": "host time, GPU calls excluded.  value = native_large: SAM text -> mates -> fragments -> rows and fragment records "
                      "for the GPU (csrc/ingest.cu on parser threads) -> evidence rules over the batch (pipeline._native_batch), on the "
                      "workload's own fragments (%s); the junction aggregation and the text writers that follow are seconds_writers.  "
                      "native / python = the whole host path (writers included) on the %d-pair prefix alone"
                      % ("mate pairs, both mates spliced" if cfg.paired else "single-end reads, one span each", n_sample),
              "sample": "%d reads (the %d-pair prefix x %d)" % (res["native_large"]["reads"], n_sample, copies), "native_large": res["native_large"], "native": res["native"],
              "python": res["python"]}
    cpu = {"value": n_sample / cpu_s, "unit": "pairs/s", "cores": 1, "kind": "port",
           "sample": "first %d pairs of the same workload, oracle scan+aggregation single process (%.1f s); SAM decode untimed" % (n_sample, cpu_s)}
    return parity, ingest, cpu


def multi_gpu_parity(eng, sh, dist, dev, idx_base, total, m=100000):
    """a peer-memory step on the first m rows of every rank's shard: owners' tables gathered to rank 0 == a second context
    on rank 0 that scans every rank's prefix itself and aggregates the union"""
    import torch

    from find_circ2_b200 import parallel
    from find_circ2_b200.engine import Engine

    world, rank = dist.get_world_size(), dist.get_rank()
    m = min(m, sh.n)
    stream = torch.cuda.current_stream().cuda_stream
    eng.agg_reset_async(stream)
    eng.agg_set_idx_range(0, total)
    eng.scan_emit_batch(sh.batch_prefix(m), sh.hits, sh.q, sh.d["read_hash"], None, idx_base, stream)
    parallel.stream_barrier(dist, dev, eng, stream)
    nj = eng.agg_finalize(stream)
    got = parallel.gather_junctions(eng.agg_fetch(nj), dist, dev)
    # rank 0 alone: every rank ships the columns of its prefix (plain NCCL gathers of tensors)
    keys = ("chrom", "a_start", "b_end", "l", "flags", "internal", "wden", "q_a", "q_b", "read_hash", "qname_hash")
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([m], dtype=torch.int64, device=dev))
    sizes = [int(x.item()) for x in sizes]
    mx = max(sizes)
    parts = {}
    for k in keys:
        v = sh.d[k][:m]
        pad = torch.zeros((mx,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
        pad[:m] = v
        raw = pad.view(torch.uint8).reshape(-1)  # (NCCL has no 16-bit integer type: ship bytes)
        bufs = [torch.zeros_like(raw) for _ in range(world)] if rank == 0 else None
        dist.gather(raw, bufs, dst=0)
        parts[k] = [b.view(v.dtype).reshape(pad.shape) for b in bufs] if rank == 0 else None
    bases = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(bases, torch.tensor([idx_base], dtype=torch.int64, device=dev))
    ok, n_junc = True, 0
    if rank == 0:
        e2 = Engine(device=dev.index, asize=eng.params.asize, margin=eng.params.margin, maxdist=eng.params.maxdist)
        e2.share_genome(eng)
        e2.agg_reset()
        for r, mr in enumerate(sizes):
            if mr == 0:
                continue
            c = {k: parts[k][r][:mr].contiguous() for k in parts}
            planes = torch.zeros(3 * sh.n_words * mr, dtype=torch.int32, device=dev)
            e2.pack_reads(c["internal"], sh.stride, c["l"], sh.n_words, planes, c["flags"], stream)
            pr = e2.make_pairs(mr, c["chrom"], c["a_start"], c["b_end"], c["l"], c["flags"], planes, sh.n_words, sh.max_l)
            hits = torch.zeros(mr * 4, dtype=torch.int32, device=dev)
            e2.scan_emit(pr, hits, c["wden"], c["q_a"], c["q_b"], c["read_hash"], c["qname_hash"], int(bases[r].item()), stream)
            torch.cuda.synchronize()
        want = e2.agg_fetch(e2.agg_finalize(stream))
        e2.close()
        ok, n_junc = (got.tobytes() == want.tobytes() and len(want) > 0), len(want)
    flag = torch.tensor([1 if ok else 0, n_junc], dtype=torch.int64, device=dev)
    dist.broadcast(flag, 0)
    return bool(flag[0].item()), int(flag[1].item())


def bench_config(args, cfg, dist, dev, primary, error_rate=None):
    import torch

    import bench_workload as W
    from find_circ2_b200 import parallel
    from find_circ2_b200._lib import HIT_DTYPE
    from find_circ2_b200.engine import Engine

    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    local = dev.index
    t_setup = time.perf_counter()
    eng = Engine(device=local, asize=cfg.asize, margin=cfg.margin, maxdist=cfg.maxdist)
    spec = W.Spec(cfg)
    flat = spec.materialize(dev)
    eng.load_genome_arrays(spec.names, spec.chrom_arrays(flat))
    total = args.pairs if (primary and args.pairs) else cfg.n_pairs
    if cfg.scaling == "weak":
        total *= world
    per = total // world // 2 * 2
    total = per * world
    sh = Shard(eng, spec, rank * per, (rank + 1) * per, dev, flat, error_rate=error_rate)
    del flat
    torch.cuda.empty_cache()
    n = sh.n
    d = sh.d
    idx_base = rank * per  # position of the rank's shard in the whole input stream (dense over all ranks)
    stream = torch.cuda.current_stream().cuda_stream
    bpp = bytes_per_pair(cfg)
    batch_bytes = n * (17 + 12 * sh.n_words + 21)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if batch_bytes < (512 << 20) else None
    setup_s = time.perf_counter() - t_setup

    # multi-GPU: records travel to the rank that owns their key, written by the scan kernel straight into the owner's
    # buffer over NVLink (CUDA IPC peer memory); fallback: partition + NCCL all-to-all
    use_p2p = world > 1 and not args.no_p2p and parallel.p2p_setup(eng, dist, dev, int(2.2 * n) + (1 << 16))

    def step_device(ev=None):
        eng.agg_reset_async(stream)  # stays behind the previous step in the stream: no host round trip before the scan
        if world > 1:
            eng.agg_set_idx_range(0, total)  # lets the owner rank the junctions without a sort
        if ev:
            ev[0].record()
        if world == 1:
            # scan + record in one kernel (fc_scan_emit_batch); no name hashes: the descriptors carry the fragment fields
            eng.scan_emit_batch(sh.batch, sh.hits, sh.q, d["read_hash"], None, idx_base, stream)
            if ev:
                ev[1].record()
                ev[2].record()
        elif use_p2p:
            # the same kernel, records land in the owner ranks' buffers (the context is connected to its peers)
            eng.scan_emit_batch(sh.batch, sh.hits, sh.q, d["read_hash"], None, idx_base, stream)
            if ev:
                ev[1].record()
            parallel.stream_barrier(dist, dev, eng, stream)  # ends the step: counts published, every rank's records have landed
            if ev:
                ev[2].record()
        else:
            eng.scan(sh.pairs, sh.hits, stream)
            eng.agg_emit(n, sh.hits, d["chrom"], d["flags"], *sh.pay, idx_base, stream)
            if ev:
                ev[1].record()
            parallel.exchange_records(eng, dist, dev, stream, upper_bound=n)
            if ev:
                ev[2].record()
        return eng.agg_finalize(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def do_flush():
        if flush is not None:
            flush.zero_()

    steps = args.steps if primary else max(3, args.steps // 4)
    warm = max(args.warmup, 3) if primary else 3
    for _ in range(warm):
        do_flush()
        nj = step_device()
    barrier()
    launches0 = eng.launch_count()
    sampler = None
    if primary:
        sampler = ClockSampler(local)
        sampler.start()
        time.sleep(0.2)
    # ---- timed: device-resident
    tot_ms = scan_ms = xchg_ms = 0.0
    barrier()
    for _ in range(steps):
        do_flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e0.record()
        nj = step_device(ev)
        e1.record()
        torch.cuda.synchronize()
        tot_ms += e0.elapsed_time(e1)
        scan_ms += ev[0].elapsed_time(ev[1])
        xchg_ms += ev[1].elapsed_time(ev[2])
    barrier()
    launches = eng.launch_count() - launches0
    n_rec = eng.agg_n_records()
    # ---- the aggregation kernel alone: a few more steps with the library's own CUDA events around its stages
    eng.agg_set_timing(True)
    stage_sum = {}
    k_acc = max(2, steps // 4)
    for _ in range(k_acc):
        do_flush()
        step_device()
        torch.cuda.synchronize()
        for k, v in eng.agg_get_timing().items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
    eng.agg_set_timing(False)
    stages_us = {k: round(v / k_acc, 1) for k, v in stage_sum.items()}
    acc_step_ms = stages_us["accumulate"] * 1e-3
    barrier()
    # ---- one step with the ordered gather of the junction tables to rank 0 inside (find_circ.py:684-686 across ranks)
    gather_ms = None
    if world > 1 and primary:
        for it in range(2):
            barrier()
            t0 = time.perf_counter()
            nj = step_device()
            parallel.gather_junctions(eng.agg_fetch(nj), dist, dev)
            torch.cuda.synchronize()
            gather_ms = (time.perf_counter() - t0) * 1e3
    ms_step = tot_ms / steps
    scan_step = scan_ms / steps

    # ---- timed: end to end through host buffers.  The batch sits in pinned host memory in the layout an ingest writes
    # (fc_batch descriptors, read rows, q words, read hashes, the N planes as a sparse list); a step streams it through the
    # slots of an fc_stream in chunks -- copy in, scan + record, compact results back, all overlapped -- then ends the
    # step (N>1: barrier), reduces and fetches the junction table.
    e2e = None
    if primary:
        from find_circ2_b200.engine import HostStream

        def pin(t):
            h = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            h.copy_(t)
            return h.numpy()
        h_meta = pin(sh.meta)
        h_rows = pin(sh.rows)
        h_q = pin(sh.q)
        h_rh = pin(d["read_hash"])
        flagged = torch.nonzero((d["flags"] & 4) != 0)[:, 0]
        h_rn_idx = flagged.to(torch.int32).cpu().numpy()
        h_rn_rows = sh.rn_rows.view(n, sh.nw)[flagged].contiguous().cpu().numpy()
        h_se = torch.empty(n * 2, dtype=torch.int32).pin_memory().numpy()
        mw_total = (n + 31) // 32 + 8
        h_hm = torch.empty(mw_total, dtype=torch.int32).pin_memory().numpy()
        h_sm = torch.empty(mw_total, dtype=torch.int32).pin_memory().numpy()
        cap = min(n, args.e2e_chunk) // 32 * 32 or n
        n_slots = 3
        hs = HostStream(eng, n_slots, cap, sh.nw)
        chunks = []
        for c0 in range(0, n, cap):
            c1 = min(n, c0 + cap)
            lo, hi = np.searchsorted(h_rn_idx, [c0, c1])
            chunks.append((c0, c1, (h_rn_idx[lo:hi] - c0).astype(np.uint32), np.ascontiguousarray(h_rn_rows[lo:hi])))
        h2d = h_meta.nbytes + h_rows.nbytes + h_q.nbytes + h_rh.nbytes + h_rn_idx.nbytes + h_rn_rows.nbytes
        parts = {"reset": 0.0, "stream": 0.0, "finalize": 0.0, "fetch": 0.0}

        def step_e2e():
            tp = [time.perf_counter()]
            eng.agg_reset()
            if world > 1:
                eng.agg_set_idx_range(0, total)
            tp.append(time.perf_counter())
            for k, (c0, c1, rn_i, rn_r) in enumerate(chunks):
                slot = k % n_slots
                hs.wait(slot)
                hs.submit(slot, c1 - c0, h_meta[4 * c0:], h_rows[2 * sh.nw * c0:], sh.nw, sh.max_l, q=h_q[c0:], read_hash=h_rh[c0:],
                          idx_base=idx_base + c0, rn_idx=rn_i if len(rn_i) else None, rn_rows=rn_r if len(rn_i) else None, emit=True,
                          out_mode=2, out_hits=h_se[2 * c0:], out_hit_mask=h_hm[c0 // 32:], out_strand_mask=h_sm[c0 // 32:])
            hs.wait_all()
            tp.append(time.perf_counter())
            if use_p2p:
                parallel.stream_barrier(dist, dev, eng, 0)
            elif world > 1:
                parallel.exchange_records(eng, dist, dev, 0, upper_bound=n)
            njj = eng.agg_finalize(0)
            tp.append(time.perf_counter())
            junc = eng.agg_fetch(njj, copy=False)  # read in place (pinned buffer of the engine)
            tp.append(time.perf_counter())
            for k, name in enumerate(("reset", "stream", "finalize", "fetch")):
                parts[name] += tp[k + 1] - tp[k]
            return njj, junc

        e2e_steps = max(3, steps // 2)
        for _ in range(2):
            nj2, junc = step_e2e()
        # the streamed path gives the junction table of the device-resident path, and the compact hits are the hits
        e2e_ok = int(nj2) == int(nj)
        hits_dev = sh.hits.view(n, 4)
        got_mask = torch.from_numpy(h_hm[: (n + 31) // 32].copy()).to(dev)
        want_hit = (hits_dev[:, 2] & 0xFFFF) != 0
        bits = ((got_mask[torch.arange(n, device=dev) // 32] >> (torch.arange(n, device=dev) % 32)) & 1).bool()
        e2e_ok = e2e_ok and bool(torch.equal(bits, want_hit)) and bool(torch.equal(torch.from_numpy(h_se.copy()).to(dev).view(n, 2)[want_hit], hits_dev[:, :2][want_hit]))
        del got_mask, bits, want_hit
        barrier()
        for k in parts:
            parts[k] = 0.0
        e2e_s = 0.0
        for _ in range(e2e_steps):
            do_flush()
            barrier()
            t0 = time.perf_counter()
            nj2, junc = step_e2e()
            e2e_s += time.perf_counter() - t0
        barrier()
        hs.close()
        e2e = {"s_step": e2e_s / e2e_steps, "h2d": int(h2d), "d2h": n * 8 + 8 * ((n + 31) // 32) + int(nj2) * 64, "ok": e2e_ok,
               "chunk_rows": cap, "slots": n_slots, "calls_ms": {k: round(v / e2e_steps * 1e3, 3) for k, v in parts.items()}}
        del h_meta, h_rows, h_q, h_rh, h_se, h_hm, h_sm
    if sampler is not None:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    # ---- parity
    parity = ingest = cpu = None
    if primary or args.check_all:
        if world > 1 and use_p2p:
            ok, n_union = multi_gpu_parity(eng, sh, dist, dev, idx_base, total)
            parity = {"ok": ok, "against": "one context aggregating the union of all ranks' prefixes (rank 0), byte for byte after the ordered gather",
                      "pairs": "first 100000 rows of every rank's shard", "junction_rows": n_union}
        elif world == 1:
            parity, ingest, cpu = dropin_parity(eng, spec, cfg, min(args.cpu_sample, per) // 4 * 4)
    elif world == 1:
        # the other configs: a short prefix of each through the whole drop-in against the oracle
        parity, _, _ = dropin_parity(eng, spec, cfg, 4000, large=False, error_rate=error_rate)

    # ---- reductions over ranks
    e2e_step = e2e["s_step"] if e2e else 0.0
    if world > 1:
        v = torch.tensor([ms_step, scan_step, e2e_step, acc_step_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms_step, scan_step, e2e_step, acc_step_ms = [float(x) for x in v.tolist()]
        cnt = torch.tensor([n, n_rec, int(nj)], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt)
        total_pairs, total_rec, total_junc = [int(x) for x in cnt.tolist()]
    else:
        total_pairs, total_rec, total_junc = n, n_rec, int(nj)
    peak, peak_src = peaks()
    # DRAM bytes per launch from the one `ncu --set full` capture of this workload (profiles/r02_kernels_ncu_summary_final.txt):
    # only quoted for the launch that was captured (config 3 at full size on one GPU)
    captured = primary and world == 1 and cfg.name == "3" and n == 49037197
    traffic_scan = 3.601469e9 + 2.857509e9 if captured else None
    traffic_acc = 2.269204e9 + 0.883749e9 if captured else None
    achieved = bpp * n / (scan_step * 1e-3) / 1e9
    merge_bytes = 48.0 * n_rec + 64.0 * int(nj)
    acc_achieved = merge_bytes / (max(acc_step_ms, 1e-9) * 1e-3) / 1e9
    fused = world == 1 or use_p2p
    roof_scan = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_scan,
                 "kernel": ("scan_emit%s_kernel (csrc/scan.cu): breakpoint scan; also writes a 48-byte record per accepted pair" % ("" if world == 1 else "_p2p")
                            if fused else "scan_kernel (csrc/scan.cu)"),
                 "bytes_per_pair": bpp, "pairs": n, "ms": scan_step, "peak_source": peak_src}
    distinct_ms = stages_us.get("distinct", 0.0) * 1e-3
    roof_acc = {"bound": "hbm", "achieved": acc_achieved, "peak": peak, "unit": "GB/s", "frac": acc_achieved / peak, "traffic": traffic_acc,
                "kernel": "fused_accumulate_kernel (csrc/agg.cu)" + ("; large inputs: the distinct counts follow in distinct_parts_kernel (%.3f ms, not in `ms`)" % distinct_ms if distinct_ms > 0.01 else ""),
                "ms": acc_step_ms,
                "bytes": "48 B x %d records + 64 B x %d junctions (rank 0)" % (n_rec, int(nj)), "peak_source": peak_src}
    dominant, other = (roof_acc, roof_scan) if acc_step_ms > scan_step else (roof_scan, roof_acc)
    res = {
        "value": total_pairs / (ms_step * 1e-3), "ms_per_step": ms_step, "steps": steps, "warmup": warm,
        "workload": cfg.workload + (" [substitution rate %.0f%%]" % (100 * error_rate) if error_rate is not None else ""),
        "detail": {
            "pairs_scanned": total_pairs, "pairs_scanned_rank0": n, "pairs_in": total, "records": total_rec, "junctions": total_junc,
            "l2": ("flushed (256 MiB memset) before every timed step" if flush is not None else "inputs larger than L2 (%.1f GB of batch columns per GPU)" % (batch_bytes / 1e9)),
            "timing": "per-step CUDA events summed over the steps; max over ranks",
            "exchange": ("none" if world == 1 else ("scan kernel writes records into the owners' slices over peer memory (CUDA IPC, NVLink), one barrier per step" if use_p2p else "partition + NCCL all-to-all")),
            "scan_ms": scan_step, "exchange_barrier_ms": xchg_ms / steps, "merge_ms": ms_step - scan_step - xchg_ms / steps,
            "accumulate_kernel_ms": acc_step_ms, "merge_stages_us": stages_us, "step_with_ordered_gather_ms": gather_ms,
            "setup_s": round(setup_s, 1), "seeds": {"workload": cfg.seed},
        },
        "roofline": dominant, "roofline_other": other, "gpu_launches": int(launches),
        "parity_checked": bool(parity and parity["ok"]), "parity": parity,
    }
    if e2e:
        res["e2e"] = {"value": total_pairs / e2e_step, "unit": "pairs/s", "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                      "ms_per_step": e2e_step * 1e3, "calls_ms": e2e["calls_ms"],
                      "equals_device_path": e2e["ok"], "chunk_rows": e2e["chunk_rows"], "slots": e2e["slots"],
                      "call": "fc_stream_submit / fc_stream_wait over %d slots + fc_agg_finalize + fc_agg_fetch: pinned host batch in fc_batch layout (16-B descriptors, read rows, q, read hash; N planes as a sparse list), compact hits (start, end, 2 bit masks) back" % e2e["slots"]}
    if sampler is not None:
        res["clocks"] = sampler.summary()
    if ingest:
        res["ingest"] = ingest
    if cpu:
        res["cpu_baseline"] = cpu
    if parity is not None and not parity["ok"]:
        sys.stderr.write("bench.py: PARITY CHECK FAILED for config %s: %r\n" % (cfg.name, parity))
    eng.close()
    del sh
    torch.cuda.empty_cache()
    return res


def _brief(r):
    return {"workload": r["workload"], "value": r["value"], "unit": "pairs/s", "ms_per_step": r["ms_per_step"], "steps": r["steps"],
            "roofline": {k: r["roofline"][k] for k in ("kernel", "frac", "achieved", "ms")},
            "roofline_other": {k: r["roofline_other"][k] for k in ("kernel", "frac", "achieved", "ms")},
            "pairs_scanned": r["detail"]["pairs_scanned"], "records": r["detail"]["records"], "junctions": r["detail"]["junctions"],
            "merge_stages_us": r["detail"]["merge_stages_us"], "parity_checked": r["parity_checked"]}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import bench_workload as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from find_circ2_b200 import parallel

    numa = parallel.bind_to_gpu_numa(local) if world > 1 else {}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfgs = W.configs()
    cfg = cfgs[args.config]
    r = bench_config(args, cfg, dist if world > 1 else None, dev, primary=True)
    line = {
        "metric": METRIC, "value": r["value"], "unit": "pairs/s", "n_gpus": world, "steps": r["steps"], "warmup": r["warmup"],
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": cfg.scaling, "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": {"workload": r["workload"]},
    }
    for k in ("detail", "roofline", "roofline_other", "e2e", "gpu_launches", "clocks", "parity_checked", "parity", "ingest", "cpu_baseline"):
        if k in r:
            line[k] = r[k]
    line.setdefault("e2e", None)
    if numa:
        line["detail"]["host_binding"] = numa
    if world == 1 and not args.no_extra:
        others = {}
        for name in [c for c in ("2", "4", "5") if c != args.config]:
            c2 = cfgs[name]
            if name == "4":
                for er in (0.01, 0.02, 0.03):
                    others["4@%d%%" % round(100 * er)] = _brief(bench_config(args, c2, None, dev, primary=False, error_rate=er))
            else:
                others[name] = _brief(bench_config(args, c2, None, dev, primary=False))
        line["other_configs"] = others
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="3", choices=["2", "3", "4", "5"], help="BASELINE.json configs[1..4]")
    ap.add_argument("--pairs", type=int, default=0, help="override the pair count of the primary config")
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--e2e-chunk", type=int, default=4 << 20, help="rows per streamed host batch of the e2e path")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: partition + NCCL all-to-all instead of peer-memory emit")
    ap.add_argument("--no-extra", action="store_true", help="N=1: skip the lines of the other configs")
    ap.add_argument("--check-all", action="store_true", help="run the parity check on the other configs too")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
