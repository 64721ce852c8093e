#!/usr/bin/env python3
"""
bench.py -- anchor pairs / second through the breakpoint scan + junction merge (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference ...                          the reference algorithm on the host cores (oracle port)

A step is one pass of the hot path over one batch: config 2 of BASELINE.json -- synthetic 100 Mb genome
(20 x 5 Mb, 0.5 % N), 10 000 planted circRNAs, 1 M anchor pairs from 100-nt reads (a=20, m=2, d=2), seeds fixed.
  value   pairs/s with the batch already resident in HBM: one step = four launches -- the scan kernel that writes the
          junction records on the way (fc_scan_emit), the one-pass aggregation kernel and the two small kernels that rank
          and convert the junctions (fc_agg_finalize)
  e2e     the same through the host-buffer C-ABI call fc_batch_host + fc_agg_finalize + fc_agg_fetch:
          pinned host SoA in, per-pair hits and the junction table out, copies inside the timed region
  roofline  the kernel with the longest launch in the step, roofline_other the second one.  scan kernel: 82
            algorithmic bytes per pair / CUDA-event time of the kernel; aggregation kernel: 48 B per
            record + 64 B per junction / the library's own CUDA events around the kernel (fc_agg_get_timing)
  cpu_baseline  oracle (python restatement of find_circ.py, one numpy compare per split position) on a bounded sample
L2 is flushed (256 MiB memset) before every timed step; per-step CUDA events are summed.
Experiments only: FC_BENCH_ZIPF=<s> changes the popularity law of the planted junctions (default 1.0, the top junction
holds 9 % of the records; 0 = uniform); FC_AGG_TIMING=1 prints the stage times of every fc_agg_finalize on stderr.
Multi-GPU: pairs are sharded by rank (weak scaling: every rank scans its own 1 M pairs), junction records are
hash-partitioned by key and written straight into the owner's buffer over NVLink (fallback: one all-to-all), every
rank reduces its keys.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ASIZE, MARGIN, MAXDIST = 20, 2, 2
READ_LEN = 100
BYTES_PER_PAIR = 16 + (READ_LEN - 2 * (ASIZE - MARGIN) + 3) // 4 + 2 * ((READ_LEN - 2 * (ASIZE - MARGIN) + 2 + 3) // 4) + 16  # 82


def workload(rank, n_pairs, genome_mb, n_circ):
    from find_circ2_b200 import synth

    per = genome_mb * 1000000 // 20
    g = synth.make_genome([per] * 20, seed=1, n_frac=0.005, n_run=(50, 5000), soft_frac=0.0)
    J = synth.plant_junctions(g, n_circ, max(n_circ // 20, 1), seed=2, span=(200, 50000), margin=400)
    t = synth.make_pairs(g, J, n_pairs, read_len=READ_LEN, asize=ASIZE, seed=3 + 1000 * rank, error_rate=0.005, zipf=float(os.environ.get("FC_BENCH_ZIPF", "1.0")),
                         frac_decoy=0.10, frac_nonuniq=0.02, frac_edge=0.01)
    return g, J, t


def soa_from_table(t, eng):
    """what ingest produces for 2-segment reads (find_circ.py:1058-1140, 821-848): scan inputs + aggregation payload"""
    eff = ASIZE - MARGIN
    R = t.read_len
    uniq_a = np.where(t.xs_a >= 0, t.as_a - t.xs_a, t.as_a)
    uniq_b = np.where(t.xs_b >= 0, t.as_b - t.xs_b, t.as_b)
    keep = np.minimum(uniq_a, uniq_b) >= 2  # is_uniq filter happens before the scan (find_circ.py:1299-1301)
    idx = np.nonzero(keep)[0]
    n = len(idx)
    soa = dict(
        chrom=t.chrom[idx].astype(np.int32),
        a_start=(t.a_pos[idx] + eff).astype(np.int32),
        b_end=(t.b_pos[idx] + t.b_len[idx] - eff).astype(np.int32),
        l=np.full(n, R - 2 * eff, dtype=np.int32),
        flags=(((t.b_pos[idx] - (t.a_pos[idx] + t.a_len[idx])) < 0).astype(np.uint8) | (t.reverse[idx].astype(np.uint8) << 1)),
        internal=np.ascontiguousarray(t.reads[idx, eff : R - eff]),
        wden=np.ones(n, dtype=np.uint8),
        q_a=(t.as_a[idx] - np.maximum(t.xs_a[idx], 0)).astype(np.int16),
        q_b=(t.as_b[idx] - np.maximum(t.xs_b[idx], 0)).astype(np.int16),
    )
    soa["read_hash"] = eng.hash_reads(t.reads[idx], np.full(n, R, dtype=np.int32))
    names = t.name_id[idx].astype(np.uint64)
    soa["qname_hash"] = (names * np.uint64(0x9E3779B97F4A7C15)) ^ (names >> np.uint64(7))
    return soa, idx


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
            time.sleep(0.1)

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


def ascii_to_planes(internal, n_words):
    """[n, L] ASCII -> (lo, hi, n) bit planes, word-major [n_words * n] uint32 (the layout of fc_pairs.rlo/rhi/rn)"""
    n, L = internal.shape
    code = np.full(internal.shape, 4, dtype=np.uint8)
    up = internal & 0xDF
    for k, ch in enumerate(b"ACGT"):
        code[up == ch] = k
    out = []
    for plane in ((code & 1) & (code < 4), ((code >> 1) & 1) & (code < 4), code == 4):
        bits = np.zeros((n, n_words * 32), dtype=np.uint8)
        bits[:, :L] = plane
        words = np.packbits(bits, axis=1, bitorder="little").view(np.uint32)  # [n, n_words]
        out.append(np.ascontiguousarray(words.T).reshape(-1))
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU side
def oracle_records(g, t, idx):
    """SAM records (decoded, untimed) for the sample rows"""
    from find_circ2_b200 import synth
    from oracle import find_circ_oracle as O

    lines = []
    for i in idx:
        lines += synth.bwa_records_for_pair(g, t, int(i), "r%d" % int(t.name_id[i]))
    names, recs = O.read_sam(synth.sam_header(g).splitlines(True) + lines)
    return names, list(recs)


def oracle_genome(g):
    from oracle import find_circ_oracle as O

    og = O.Genome.__new__(O.Genome)
    og.names = list(g.names)
    og.seqs = {n: s.tobytes().decode() for n, s in zip(g.names, g.seqs)}
    return og


_CPU_GENOME = None  # (oracle genome, @SQ names): inherited by forked workers instead of being pickled


def _cpu_shard(recs):
    from oracle import find_circ_oracle as O

    og, names = _CPU_GENOME
    t0 = time.perf_counter()
    r = O.Run(og, names, O.Options(asize=ASIZE, margin=MARGIN, maxdist=MAXDIST))
    r.process(recs)
    r.outputs()
    return time.perf_counter() - t0, r.n_spans


def cpu_baseline(g, t, n_sample, procs=1):
    """scan + aggregation of the oracle on n_sample pairs; SAM decoding is done before the clock starts"""
    global _CPU_GENOME
    idx = np.arange(min(n_sample, len(t)))
    names, recs = oracle_records(g, t, idx)
    _CPU_GENOME = (oracle_genome(g), names)
    if procs <= 1:
        dt, spans = _cpu_shard(recs)
        return len(idx) / dt, dt
    import multiprocessing as mp

    # shard by fragment (records of one read stay together: 2 records per read here)
    per = (len(recs) // 2 + procs - 1) // procs * 2
    shards = [recs[k : k + per] for k in range(0, len(recs), per)]
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(len, shards[:1])  # workers are up before the clock starts
        t0 = time.perf_counter()
        pool.map(_cpu_shard, shards, chunksize=1)
        dt = time.perf_counter() - t0
    return len(idx) / dt, dt


def host_ingest(g, t, n_sample, eng):
    """host ingest reported separately (north_star): SAM text -> fragments -> struct-of-arrays batches in the python host of
    the drop-in (find_circ2_b200/pipeline.py), measured on a sample by running the whole CLI path and subtracting the time
    spent inside GPU calls"""
    import tempfile

    from find_circ2_b200 import cli, synth

    idx = np.arange(min(n_sample, len(t)))
    with tempfile.TemporaryDirectory() as tmp:
        sam = os.path.join(tmp, "sample.sam")
        with open(sam, "w") as fh:
            fh.write(synth.sam_header(g))
            for i in idx:
                fh.writelines(synth.bwa_records_for_pair(g, t, int(i), "r%d" % int(t.name_id[i])))
        res = {}
        for tag, native in (("python", False), ("native", True)):
            opt = cli.parse_args(["-G", "unused", "-a", str(ASIZE), "-n", "bench"])[0]
            out = cli.run_to_strings(opt, sam, engine=eng, native=native)
            host_s = out["seconds_total"] - out["seconds_gpu_calls"]
            res[tag] = {"pairs_per_s": len(idx) / host_s, "seconds_host": host_s, "seconds_gpu_calls": out["seconds_gpu_calls"]}
    return {"value": res["native"]["pairs_per_s"], "unit": "pairs/s",
            "kind": "SAM text -> fragments -> SoA batches -> writers on the host, GPU calls excluded; native = csrc/ingest.cu (C++), python = pipeline.py",
            "sample": "%d reads" % len(idx), "native": res["native"], "python": res["python"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    g, J, t = workload(0, max(args.cpu_sample * cores, 1000), args.genome_mb, args.n_circ)
    n_sample = len(t)
    vals = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_baseline(g, t, min(2000, n_sample), 1)
    for _ in range(args.steps):
        v, dt = cpu_baseline(g, t, n_sample, cores)
        vals.append((v, dt))
    v = float(np.mean([x[0] for x in vals]))
    ms = float(np.mean([x[1] for x in vals])) * 1e3
    line = {
        "impl": "reference", "metric": "anchor pairs/sec (breakpoint scan + junction merge)", "value": v, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "configs[1]: synthetic %d Mb genome, %d planted circRNAs, 100-nt reads, a=20 m=2 d=2" % (args.genome_mb, args.n_circ),
                   "sample_pairs_per_step": n_sample},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": "%d pairs per step, sharded %d ways by read with multiprocessing; SAM decode untimed" % (n_sample, cores)},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU side
def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from find_circ2_b200.engine import Engine
    from find_circ2_b200._lib import HIT_DTYPE, JREC_DTYPE
    from find_circ2_b200 import parallel

    eng = Engine(device=local, asize=ASIZE, margin=MARGIN, maxdist=MAXDIST)
    g, J, t = workload(rank, args.pairs, args.genome_mb, args.n_circ)
    eng.load_genome_arrays(g.names, g.seqs)
    soa, idx = soa_from_table(t, eng)
    n = len(idx)
    max_l = int(soa["l"].max())
    n_words = max(1, (max_l + 31) // 32)
    stride = soa["internal"].shape[1]

    # ---- device-resident copy of the batch (for `value`)
    tn = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d = {k: tn(v if v.dtype != np.uint64 else v.view(np.int64)) for k, v in soa.items()}
    planes = torch.zeros(3 * n_words * n, dtype=torch.int32, device=dev)
    hits = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    eng.pack_reads(d["internal"], stride, d["l"], n_words, planes, d["flags"], stream)
    pairs = eng.make_pairs(n, d["chrom"], d["a_start"], d["b_end"], d["l"], d["flags"], planes, n_words, max_l)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    idx_base = rank * args.pairs  # position of the rank's shard in the whole input stream (dense over all ranks)

    # multi-GPU: records travel to the rank that owns their key.  Preferred: the emit kernel writes them straight into
    # the owner's buffer over NVLink (CUDA IPC peer memory); fallback: partition + NCCL all-to-all.
    use_p2p = world > 1 and not args.no_p2p and parallel.p2p_setup(eng, dist, dev, int(2.5 * n) + (1 << 16))

    def step_device(ev_scan=None):
        if use_p2p:
            eng.agg_reset_async(stream)
            eng.agg_set_idx_range(0, world * args.pairs)  # lets the owner rank the junctions without a sort
        else:
            eng.agg_reset_async(stream)  # stays behind the L2-flush kernel in the stream: no host round trip before the scan
        if ev_scan:
            ev_scan[0].record()
        if world == 1:
            # scan + record in one kernel (fc_scan_emit)
            eng.scan_emit(pairs, hits, d["wden"], d["q_a"], d["q_b"], d["read_hash"], d["qname_hash"], idx_base, stream)
            if ev_scan:
                ev_scan[1].record()
                ev_scan[2].record()
                ev_scan[3].record()
            return eng.agg_finalize(stream)
        if use_p2p:
            # scan + records straight into the owner ranks' buffers, one kernel (fc_scan_emit_p2p)
            eng.scan_emit_p2p(pairs, hits, d["wden"], d["q_a"], d["q_b"], d["read_hash"], d["qname_hash"], idx_base, stream)
            if ev_scan:
                ev_scan[1].record()
                ev_scan[2].record()
            parallel.stream_barrier(dist, dev, eng, stream)  # every rank's records have landed
        else:
            eng.scan(pairs, hits, stream)
            if ev_scan:
                ev_scan[1].record()
            eng.agg_emit(n, hits, d["chrom"], d["flags"], d["wden"], d["q_a"], d["q_b"], d["read_hash"], d["qname_hash"], idx_base, stream)
            if ev_scan:
                ev_scan[2].record()
            if world > 1:
                parallel.exchange_records(eng, dist, dev, stream, upper_bound=n)
        if ev_scan:
            ev_scan[3].record()
        return eng.agg_finalize(stream)

    # ---- pinned host copy of the batch (for `e2e`)
    def pin(a):
        tt = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return tt, tt.numpy()
    pinned = {k: pin(v) for k, v in soa.items()}
    h_hits_t = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
    h_hits = h_hits_t.numpy().view(HIT_DTYPE)
    h2d = sum(v[1].nbytes for v in pinned.values())

    # what the native ingest hands over: the internal read part as bit planes (word-major), 3 x 4 x n_words bytes per pair
    planes_np = ascii_to_planes(soa["internal"], n_words)
    pl_pin = [pin(planes_np[k]) for k in range(3)]
    h2d_planes = sum(v[1].nbytes for k, v in pinned.items() if k != "internal") + sum(v[1].nbytes for v in pl_pin)

    e2e_parts = {"reset": 0.0, "batch": 0.0, "finalize": 0.0, "fetch": 0.0}

    def step_e2e(ascii_reads=False):
        tp = [time.perf_counter()]
        eng.agg_reset()
        tp.append(time.perf_counter())
        p = {k: v[1] for k, v in pinned.items()}
        if ascii_reads:
            eng.batch_host(p["chrom"], p["a_start"], p["b_end"], p["l"], p["flags"], p["internal"], p["wden"], p["q_a"], p["q_b"],
                           p["read_hash"], p["qname_hash"], idx_base, emit=True, out=h_hits)
        else:
            eng.batch_host_planes(n, p["chrom"], p["a_start"], p["b_end"], p["l"], p["flags"], pl_pin[0][1], pl_pin[1][1],
                                  pl_pin[2][1], n_words, n, max_l, p["wden"], p["q_a"], p["q_b"], p["read_hash"],
                                  p["qname_hash"], idx=None, idx_base=idx_base, emit=True, out=h_hits)
        tp.append(time.perf_counter())
        if world > 1:
            parallel.exchange_records(eng, dist, dev, 0, upper_bound=n)
        nj = eng.agg_finalize(0)
        tp.append(time.perf_counter())
        junc = eng.agg_fetch(nj, copy=False)  # read in place (pinned buffer of the engine)
        tp.append(time.perf_counter())
        if not ascii_reads:
            for k, name in enumerate(("reset", "batch", "finalize", "fetch")):
                e2e_parts[name] += tp[k + 1] - tp[k]
        return nj, junc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        nj = step_device()
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    # ---- timed: device-resident
    tot_ms, scan_ms, emit_ms, pre_ms, xchg_ms = 0.0, 0.0, 0.0, 0.0, 0.0
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0, s1, s2, s3 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e0.record()
        nj = step_device((s0, s1, s2, s3))
        e1.record()
        torch.cuda.synchronize()
        tot_ms += e0.elapsed_time(e1)
        scan_ms += s0.elapsed_time(s1)
        emit_ms += s1.elapsed_time(s2)
        pre_ms += e0.elapsed_time(s0)   # reset (+ the first barrier on the multi-GPU path)
        xchg_ms += s2.elapsed_time(s3)  # second barrier / all-to-all (multi-GPU)
    barrier()
    launches = eng.launch_count() - launches0
    # ---- the aggregation kernel alone: a few more steps with the library's own CUDA events around its stages
    eng.agg_set_timing(True)
    acc_us, stage_sum = 0.0, {}
    for _ in range(args.steps):
        flush.zero_()
        step_device()
        torch.cuda.synchronize()
        st = eng.agg_get_timing()
        acc_us += st["accumulate"]
        for k, v in st.items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
    eng.agg_set_timing(False)
    acc_step_ms = acc_us / args.steps * 1e-3
    stages_us = {k: round(v / args.steps, 1) for k, v in stage_sum.items()}
    barrier()
    # ---- timed: end to end through host buffers (wall clock brackets synchronous calls; device idle otherwise)
    for _ in range(2):
        step_e2e()
    barrier()
    e2e_s = 0.0
    for k in e2e_parts:
        e2e_parts[k] = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nj2, junc = step_e2e()
        e2e_s += time.perf_counter() - t0
    barrier()
    # the same with ASCII read bases in the host buffers (what the python ingest produces)
    e2e_ascii_s = 0.0
    step_e2e(True)
    for _ in range(max(args.steps // 2, 1)):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_e2e(True)
        e2e_ascii_s += time.perf_counter() - t0
    e2e_ascii_step = e2e_ascii_s / max(args.steps // 2, 1)
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    d2h = n * 16 + int(nj2) * 64

    ms_step = tot_ms / args.steps
    scan_step = scan_ms / args.steps
    e2e_step = e2e_s / args.steps
    if world > 1:
        v = torch.tensor([ms_step, scan_step, e2e_step], dtype=torch.float64, device=dev)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms_step, scan_step, e2e_step = [float(x) for x in v.tolist()]
        cnt = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt)
        total_pairs = int(cnt.item())
    else:
        total_pairs = n

    n_rec = eng.agg_n_records() if world == 1 else n
    if rank == 0:
        peak, peak_src = peaks()
        # single GPU: the scan kernel also writes the junction records (48 B per accepted pair) and reads their payload
        # columns (22 B per pair: wden, q_a, q_b, read hash, name hash)
        fused_scan = world == 1 or use_p2p  # the scan kernel writes the records itself
        n_emitted = n_rec if world == 1 else 0.9 * n  # (multi-GPU: n_rec counts what this rank RECEIVED; ~90 % of the pairs emit)
        scan_bytes = BYTES_PER_PAIR * n + ((48.0 * n_emitted + 22.0 * n) if fused_scan else 0.0)
        achieved = scan_bytes / (scan_step * 1e-3) / 1e9
        merge_bytes = 48.0 * n_rec + 64.0 * int(nj)
        acc_achieved = merge_bytes / (max(acc_step_ms, 1e-9) * 1e-3) / 1e9
        roof_scan = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": (87.6e6 if fused_scan else 46.4e6),
                     "kernel": ("scan_emit%s_kernel<NP=3,T=1> (csrc/scan.cu): scan + 48-byte record per accepted pair" % ("" if world == 1 else "_p2p")
                                if fused_scan else "scan_kernel<NP=3,T=1> (csrc/scan.cu)"),
                     "bytes_per_pair": scan_bytes / n, "ms": scan_step,
                     "peak_source": peak_src, "traffic_source": "ncu --set full, profiles/r01_kernels_ncu_summary.txt"}
        roof_acc = {"bound": "hbm", "achieved": acc_achieved, "peak": peak, "unit": "GB/s", "frac": acc_achieved / peak,
                    "traffic": 130.7e6, "kernel": "fused_accumulate_kernel (csrc/agg.cu)", "ms": acc_step_ms,
                    "bytes": "48 B x %d records + 64 B x %d junctions (rank 0)" % (n_rec, int(nj)), "peak_source": peak_src,
                    "traffic_source": "ncu --set full, profiles/r01_kernels_ncu_summary.txt",
                    "note": "random 16/32-byte accesses to hash tables: bound by L1/LSU wavefronts and L2 atomics, not by bytes"}
        dominant, other = (roof_acc, roof_scan) if acc_step_ms > scan_step else (roof_scan, roof_acc)
        line = {
            "metric": "anchor pairs/sec (breakpoint scan + junction merge)", "value": total_pairs / (ms_step * 1e-3),
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {
                "workload": "configs[1]: synthetic %d Mb genome (20 chrom, 0.5%% N), %d planted circRNAs, %d anchor pairs/GPU from 100-nt reads, a=20 m=2 d=2, 0.5%% substitutions, 10%% decoys"
                            % (args.genome_mb, args.n_circ, args.pairs),
                "pairs_scanned_per_gpu": n, "junctions_rank0": int(nj), "l2": "flushed (256 MiB memset) before every timed step",
                "timing": "per-step CUDA events summed over the steps; max over ranks",
                "exchange": ("none" if world == 1 else ("fused emit+exchange over peer memory (CUDA IPC, NVLink)" if use_p2p else "partition + NCCL all-to-all")),
                "scan_ms": scan_step, "merge_ms": ms_step - scan_step, "accumulate_kernel_ms": acc_step_ms,
                "emit_ms": emit_ms / args.steps, "reset_barrier_ms": pre_ms / args.steps, "exchange_barrier_ms": xchg_ms / args.steps, "seeds": {"genome": 1, "junctions": 2, "pairs": "3+1000*rank"},
            },
            "roofline": dominant,        # the kernel with the longest launch inside the step
            "roofline_other": other,     # the second kernel of the path
            "merge_stages_us": stages_us,
            "e2e": {"value": total_pairs / e2e_step, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d_planes), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_step * 1e3,
                    "calls_ms": {k: round(v / args.steps * 1e3, 3) for k, v in e2e_parts.items()},
                    "call": "fc_batch_host_planes + fc_agg_finalize + fc_agg_fetch (pinned host SoA with bit-plane reads, as csrc/ingest.cu emits them)",
                    "ascii_reads": {"value": n * world / e2e_ascii_step, "h2d_bytes_per_step": int(h2d), "ms_per_step": e2e_ascii_step * 1e3,
                                    "call": "fc_batch_host (ASCII read bases, packed on the device)"}},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        if world == 1:
            line["ingest"] = host_ingest(g, t, min(args.cpu_sample, 20000), eng)
            v, dt = cpu_baseline(g, t, args.cpu_sample, 1)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": 1, "kind": "port",
                                    "sample": "first %d pairs of the same workload, oracle scan+aggregation single process (%.1f s); SAM decode untimed" % (args.cpu_sample, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1000000)
    ap.add_argument("--genome-mb", type=int, default=100)
    ap.add_argument("--n-circ", type=int, default=10000)
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: use partition + NCCL all-to-all instead of peer-memory emit")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
