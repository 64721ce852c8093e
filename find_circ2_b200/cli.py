"""
Command line of the drop-in: the option table of the shipped find_circ.py (find_circ.py:383-413), plus the
README (v1.2) spellings of the same switches (README.md:283-337), the same output directory layout
(find_circ.py:420-458) and the same run.log counter dump (find_circ.py:1605-1607).

    bwa mem ... | find_circ.py --genome genome.fa -n sample -o out_dir
    find_circ.py --genome genome.fa -o out_dir alignments.bam
"""
from __future__ import annotations

import gzip
import logging
import optparse
import os
import sys
import time
import traceback

from . import samio
from .pipeline import VERSION, Options, Run

USAGE = """
   bwa mem -t<threads> [-p] -A2 -B10 -k 15 -T 1 $GENOME_INDEX reads.fastq.gz | %prog [options]

   OR:

   %prog [options] <bwa_mem_genome_alignments.bam>
"""


def build_parser() -> optparse.OptionParser:
    p = optparse.OptionParser(usage=USAGE)
    a = p.add_option
    a("-v", "--version", dest="version", action="store_true", default=False, help="print the version and exit")
    a("-S", "--system", dest="system", type=str, default="", help="model system database (not supported: needs the byo library)")
    a("-G", "--genome", dest="genome", type=str, default="", help="genome as ONE multi-sequence FASTA file")
    a("", "--known-circ", dest="known_circ", type=str, default="", help="BED6 file of known circRNA junctions: they keep their names in the output")
    a("", "--known-lin", dest="known_lin", type=str, default="", help="BED6 file of known linear splice junctions: they keep their names in the output")
    a("-o", "--output", dest="output", default="find_circ_run", help="output directory (created if missing)")
    a("-q", "--silent", dest="silent", default=False, action="store_true", help="no summary lines on stdout")
    a("", "--stdout", dest="stdout", default=None, choices=["circs", "lins", "reads", "multi", "test"],
      help="write this output to stdout instead of its file")
    a("-n", "--name", dest="name", default="unknown", help="sample name used in junction names and the tissues column [unknown]")
    a("", "--min-uniq-qual", "--min_uniq_qual", dest="min_uniq_qual", type=int, default=2,
      help="smallest AS-XS margin of both segments for a pair to be scanned [2]")
    a("-a", "--anchor", dest="asize", type=int, default=15, help="minimal aligned length of a segment [15]")
    a("-m", "--margin", dest="margin", type=int, default=2, help="how far a breakpoint may lie inside a segment [2]")
    a("-d", "--max-mismatch", "--maxdist", dest="maxdist", type=int, default=2,
      help="mismatches tolerated when extending the segments to the breakpoint [2]")
    a("", "--short-threshold", dest="short_threshold", type=int, default=100, help="junctions spanning less are labelled SHORT [100]")
    a("", "--huge-threshold", dest="huge_threshold", type=int, default=100000, help="junctions spanning more are labelled HUGE [100000]")
    a("", "--debug", dest="debug", default=False, action="store_true", help="(accepted, no effect)")
    a("", "--profile", dest="profile", default=False, action="store_true", help="(accepted, no effect)")
    a("", "--non-canonical", "--noncanonical", dest="noncanonical", default=False, action="store_true",
      help="do not insist on GT/AG (CT/AC) at the breakpoint")
    a("", "--all-hits", "--allhits", dest="allhits", default=False, action="store_true", help="record every tied breakpoint, not only the first")
    a("", "--stranded", dest="stranded", default=False, action="store_true", help="not supported (crashes in the reference: find_circ.py:533)")
    a("", "--strand-pref", "--strandpref", dest="strandpref", default=False, action="store_true",
      help="break ties in favour of the strand of the read")
    a("", "--half-unique", "--halfunique", "--halfuniq", dest="halfunique", default=False, action="store_true",
      help="keep junctions with a single uniquely placed side")
    a("", "--report-nobridges", "--report_nobridges", "--report_nobridge", dest="report_nobridges", default=False, action="store_true",
      help="keep junctions without any read whose two segments are both unique")
    a("-B", "--bam", dest="bam", default=False, action="store_true", help="not supported")
    a("-t", "--throughput", dest="throughput", default=False, action="store_true", help="accepted for compatibility")
    a("", "--chunk-size", "--chunksize", dest="chunksize", type=int, default=100000, help="accepted for compatibility")
    a("", "--noop", dest="noop", default=False, action="store_true", help="decode the alignments only, no junction search")
    a("", "--test", dest="test", default=False, action="store_true",
      help="write test_results.tsv: every fragment against the structure encoded in its read name (simulated reads)")
    a("", "--no-linear", dest="nolinear", default=False, action="store_true", help="ignore linear junctions of fragments without a back-splice")
    a("", "--no-multi", dest="multi_events", default=True, action="store_false", help="do not write multi_events.tsv rows")
    a("", "--batch-pairs", dest="batch_pairs", type=int, default=1 << 18, help="anchor pairs per GPU batch (default 262144)")
    a("", "--device", dest="device", type=int, default=0, help="CUDA device (default 0)")
    a("", "--gpus", dest="gpus", type=int, default=1,
      help="GPUs of this node to use (default 1): the run is re-launched as one process per GPU; needs a SAM text file as input")
    a("", "--python-ingest", dest="native", default=True, action="store_false",
      help="decode SAM text in python instead of the native (C++) ingest")
    return p


def parse_args(argv):
    o, args = build_parser().parse_args(list(argv))
    for bad in ("stranded", "bam"):
        if getattr(o, bad):
            raise SystemExit("option --%s is not supported by this build" % bad)
    if o.system:
        raise SystemExit("-S/--system is not supported by this build")
    opt = Options(
        genome=o.genome, output=o.output, name=o.name, min_uniq_qual=o.min_uniq_qual, asize=o.asize, margin=o.margin,
        maxdist=o.maxdist, short_threshold=o.short_threshold, huge_threshold=o.huge_threshold, noncanonical=o.noncanonical,
        allhits=o.allhits, strandpref=o.strandpref, halfunique=o.halfunique, report_nobridges=o.report_nobridges,
        nolinear=o.nolinear, multi_events=o.multi_events, throughput=o.throughput, chunksize=o.chunksize, noop=o.noop,
        silent=o.silent, stdout=o.stdout, batch_pairs=o.batch_pairs, device=o.device, native=o.native,
        known_circ=o.known_circ, known_lin=o.known_lin, test=o.test,
    )
    return opt, args, o


class GzipMembers(object):
    """spliced_reads.fastq.gz writer (find_circ.py:427, 1442-1447): text in, gzip out.  The text is cut into blocks that
    worker threads compress on their own (zlib releases the GIL) and that are written one after the other as members of
    ONE gzip file -- every gzip reader concatenates members.  Level 9 on a single thread, the reference's way, took two
    thirds of the run time of the whole drop-in."""

    BLOCK = 8 << 20

    def __init__(self, path: str, level: int = 6, threads: int = 0):
        self.fh = open(path, "wb")
        self.level = level
        self.threads = threads or max(1, min(16, (os.cpu_count() or 1)))
        self.pending = []
        self.size = 0

    def write(self, text: str):
        if text:
            self.pending.append(text)
            self.size += len(text)
            if self.size >= 16 * self.BLOCK:
                self._flush()

    def _flush(self):
        if not self.pending:
            return
        data = "".join(self.pending).encode("latin-1")
        self.pending, self.size = [], 0
        blocks = [data[i : i + self.BLOCK] for i in range(0, len(data), self.BLOCK)]
        if len(blocks) == 1 or self.threads == 1:
            parts = [gzip.compress(b, compresslevel=self.level) for b in blocks]
        else:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(self.threads) as pool:
                parts = list(pool.map(lambda b: gzip.compress(b, compresslevel=self.level), blocks))
        for part in parts:
            self.fh.write(part)

    def close(self):
        self._flush()
        if self.fh.tell() == 0:
            self.fh.write(gzip.compress(b""))  # an empty but valid gzip file
        self.fh.close()


def native_ok(opt: Options, path) -> bool:
    """the native ingest covers SAM text (a file whose name ends in 'sam', or stdin) and BAM (any other name:
    find_circ.py:461-469); --all-hits / --noop / --test use the python reader"""
    return not opt.allhits and not opt.noop and not opt.test and opt.native


class _Prefixed(object):
    """a binary stream with some bytes put back in front of it"""

    def __init__(self, prefix: bytes, fh):
        self.prefix, self.fh = prefix, fh

    def read(self, n: int) -> bytes:
        if self.prefix:
            head, self.prefix = self.prefix[:n], self.prefix[n:]
            if len(head) == n:
                return head
            return head + self.fh.read(n - len(head))
        return self.fh.read(n)


def _stream_header(fh):
    """@SQ names from the head of a SAM text stream; returns (names, stream positioned at the first record)"""
    names, first = [], b""
    while True:
        line = fh.readline()
        if not line:
            break
        if not line.startswith(b"@"):
            first = line
            break
        if line.startswith(b"@SQ"):
            for x in line.rstrip(b"\r\n").split(b"\t")[1:]:
                if x.startswith(b"SN:"):
                    names.append(x[3:].decode("latin-1"))
    return names, _Prefixed(first, fh)


class _Range(object):
    """bytes [start, end) of a binary file as a stream"""

    def __init__(self, fh, start: int, end: int):
        self.fh, self.left = fh, max(0, end - start)
        fh.seek(start)

    def read(self, n: int) -> bytes:
        if self.left <= 0:
            return b""
        data = self.fh.read(min(n, self.left))
        self.left -= len(data)
        return data


def _fragment_boundary(fh, pos: int, body_start: int, size: int) -> int:
    """where the part of the stream that starts nominally at byte `pos` really starts: the first line at or after pos whose
    read name differs from the name of the line before it (a fragment = consecutive records of one name,
    find_circ.py:1450-1486).  Every rank applies the same rule to its start and to its end."""
    if pos <= body_start:
        return body_start
    if pos >= size:
        return size
    fh.seek(pos - 1)
    fh.readline()  # finish the line that pos falls into (or the newline right before pos)
    first = fh.readline()
    if not first:
        return size
    name = first.split(b"\t", 1)[0]
    while True:
        at = fh.tell()
        line = fh.readline()
        if not line:
            return size
        if line.split(b"\t", 1)[0] != name:
            return at


def sam_ranges(path: str, world: int):
    """(header names, [(start, end)] * world, body start): contiguous byte ranges of a SAM text file cut on fragment boundaries"""
    size = os.path.getsize(path)
    with open(path, "rb") as fh:
        names = []
        while True:
            at = fh.tell()
            line = fh.readline()
            if not line or not line.startswith(b"@"):
                body = at
                break
            if line.startswith(b"@SQ"):
                for x in line.rstrip(b"\r\n").split(b"\t")[1:]:
                    if x.startswith(b"SN:"):
                        names.append(x[3:].decode("latin-1"))
        cuts = [_fragment_boundary(fh, body + (size - body) * r // world, body, size) for r in range(world)] + [size]
    for r in range(1, world + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return names, [(cuts[r], cuts[r + 1]) for r in range(world)], body


def _merge_host_state(run: Run, parts):
    """rank 0: fold the host-side companions of the other ranks' parts of the stream into `run` (parts in rank order =
    stream order): counters, per-junction flags, spliced reads, multi-event rows"""
    from collections import defaultdict

    run.N = defaultdict(float)
    run._ev_py, run.ev_native, run._info_cache = [], [], None
    run.reads_out, run.native_reads, run.multi_out, run.native_multi, run.test_out = [], [], [], [], []
    run.n_fragments = run.n_pairs_scanned = 0
    run.t_scan = 0.0
    for p in parts:
        for k, v in p["N"].items():
            run.N[k] += v
        run.ev_native.append(p["events"])
        run.reads_out.extend(p["reads_out"])
        run.native_reads.extend(p["native_reads"])
        run.multi_out.extend(p["multi_out"])
        run.native_multi.extend(p["native_multi"])
        run.test_out.extend(p["test_out"])
        run.n_fragments += p["n_fragments"]
        run.n_pairs_scanned += p["n_pairs_scanned"]
        run.t_scan = max(run.t_scan, p["t_scan"])


def run_distributed(opt: Options, path, dist, torch_dev, engine=None):
    """the whole run on world_size GPUs of one node (one process each, launched by torchrun or by --gpus N): every rank takes
    a contiguous part of the SAM text file, cut where the read name changes; stream positions stay global (rank r numbers
    its fragments from r * stride), so junction names (find_circ.py:684-686) and every per-junction column come out as in
    a single-process run; junction records travel to the rank that owns their key, junction rows and the host-side
    companions are gathered to rank 0, which alone returns the outputs (the other ranks return None)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if not path or path == "-" or not path.endswith("sam"):
        raise ValueError("a multi-GPU run needs a SAM text FILE (every rank reads its own byte range of it)")
    if not native_ok(opt, path):
        raise ValueError("a multi-GPU run uses the native ingest: --all-hits, --noop, --test and --python-ingest are single-GPU options")
    names, ranges, body = sam_ranges(path, world)
    stride = max(1, max(e - s for s, e in ranges) // 16 + 1)  # more than the fragments any rank can hold (a SAM line is > 16 bytes)
    run = Run(opt, names, engine)
    try:
        t0 = time.perf_counter()
        start, end = ranges[rank]
        run.cur_seq = run.n_fragments = 0
        with open(path, "rb") as fh:
            run.process_native(_Range(fh, start, end), first_fragment=rank * stride, at_stream_start=(start == body))
        t1 = time.perf_counter()
        run.finalize(dist, torch_dev)
        part = {"N": dict(run.N), "events": run.events(), "native_multi": run.native_multi,
                "reads_out": run.reads_out, "native_reads": run.native_reads, "multi_out": run.multi_out, "test_out": run.test_out,
                "n_fragments": run.n_fragments, "n_pairs_scanned": run.n_pairs_scanned, "t_scan": run.t_scan}
        parts = [None] * world if rank == 0 else None
        dist.gather_object(part, parts, dst=0)
        if rank != 0:
            return None
        _merge_host_state(run, parts)
        return {
            "circ": run.bed_text(0), "lin": run.bed_text(1), "reads": run.reads_text(), "multi": run.multi_text(), "test": run.test_text(),
            "counters": run.counters_text(), "n_fragments": run.n_fragments, "n_pairs_scanned": run.n_pairs_scanned,
            "seconds_ingest_and_scan": t1 - t0, "seconds_gpu_calls": run.t_scan, "seconds_total": time.perf_counter() - t0,
        }
    finally:
        run.close()


def distributed_context(device_ok: bool = True):
    """(dist, torch device) when this process is one rank of a torchrun launch (RANK / WORLD_SIZE > 1 in the environment)"""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return None, None
    import torch
    import torch.distributed as dist

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if device_ok and torch.cuda.is_available():
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        if not dist.is_initialized():
            dist.init_process_group("gloo")
    return dist, dev


def run_to_strings(opt: Options, path=None, engine=None, native=None):
    """the whole run, outputs as strings (tests and the single-process CLI share this)"""
    native = native_ok(opt, path) if native is None else native
    stream = None
    if native and (not path or path == "-"):
        names, stream = _stream_header(sys.stdin.buffer)
        records = None
    elif native and not path.endswith("sam"):
        from .ingest import BamText

        stream = BamText(path)  # BAM: inflated and turned into SAM text records in C++ (csrc/bam.cu)
        names = stream.names
        records = None
    elif native:
        names = samio.sam_header_names(path)
        records = None
    else:
        names, lengths, records = samio.open_alignments(path)
    run = Run(opt, names, engine)
    try:
        t0 = time.perf_counter()
        if stream is not None:
            try:
                run.process_native(stream)
            finally:
                if hasattr(stream, "close"):
                    stream.close()
        elif native:
            with open(path, "rb") as fh:
                run.process_native(fh)
        else:
            run.process(records)
        t1 = time.perf_counter()
        stages = {}

        def timed(name, fn, *a):
            t = time.perf_counter()
            r = fn(*a)
            stages[name] = round(time.perf_counter() - t, 4)
            return r

        timed("aggregate", run.finalize)
        out = {
            "circ": timed("circ_bed", run.bed_text, 0),
            "lin": timed("lin_bed", run.bed_text, 1),
            "reads": timed("reads", run.reads_text),
            "multi": timed("multi_events", run.multi_text),
            "test": run.test_text(),
            "counters": run.counters_text(),
            "seconds_writer_stages": stages,
            "n_fragments": run.n_fragments,
            "n_pairs_scanned": run.n_pairs_scanned,
            "seconds_ingest_and_scan": t1 - t0,
            "seconds_gpu_calls": run.t_scan,
            "seconds_ingest_stages": {k: round(v, 4) for k, v in getattr(run, "t_ingest", {}).items()},
            "seconds_total": time.perf_counter() - t0,
        }
    finally:
        run.close()
    return out


V12_ONLY = ("-p", "--prefix", "-s", "--stats", "-R", "--reads", "-w", "--wiggle", "-r", "--reads2samples", "--format=1.2")


def wants_v12(argv) -> bool:
    """the two generations of the command line share `-q` with different meanings (README.md:303 vs find_circ.py:390);
    the v1.2 face is selected by any option only it has, or by --format=1.2"""
    for a in argv:
        head = a.split("=")[0]
        if a in V12_ONLY or head in V12_ONLY:
            return True
    return False


def build_parser_v12() -> optparse.OptionParser:
    """README.md:283-337"""
    p = optparse.OptionParser(usage="\n  bowtie2 [mapping options] anchors.fastq.gz | %prog [options] > candidates.bed\n")
    a = p.add_option
    a("-v", "--version", dest="version", action="store_true", default=False)
    a("-G", "--genome", dest="genome", type=str, default="")
    a("-n", "--name", dest="name", default="unknown")
    a("-p", "--prefix", dest="prefix", default="")
    a("-q", "--min_uniq_qual", dest="min_uniq_qual", type=int, default=2)
    a("-a", "--anchor", dest="asize", type=int, default=20)
    a("-m", "--margin", dest="margin", type=int, default=2)
    a("-d", "--maxdist", dest="maxdist", type=int, default=2)
    a("-w", "--wiggle", dest="wiggle", type=int, default=2)
    a("", "--noncanonical", dest="noncanonical", default=False, action="store_true")
    a("", "--allhits", dest="allhits", default=False, action="store_true")
    a("", "--halfunique", "--halfuniq", dest="halfunique", default=False, action="store_true")
    a("", "--report_nobridges", "--report_nobridge", dest="report_nobridges", default=False, action="store_true")
    a("-R", "--reads", dest="reads", default=None)
    a("-s", "--stats", dest="stats", default="runstats.log")
    a("", "--format", dest="format", default="1.2")
    a("", "--batch-pairs", dest="batch_pairs", type=int, default=1 << 18)
    a("", "--device", dest="device", type=int, default=0)
    for unsupported in ("--randomize", "--stranded", "--strandpref"):
        a("", unsupported, dest=unsupported.strip("-"), default=False, action="store_true")
    a("-B", "--bam", dest="bam", default=None)
    a("-r", "--reads2samples", dest="reads2samples", default="")
    a("-S", "--system", dest="system", default="")
    return p


def run_v12_to_strings(opt, path=None, engine=None):
    from .v12 import RunV12

    names, lengths, records = samio.open_alignments(path)
    run = RunV12(opt, names, engine)
    try:
        run.process(records)
        run.finalize()
        bed, reads, stats = run.bed_and_reads()
    finally:
        run.close()
    return {"bed": bed, "reads": reads, "stats": stats, "n_pairs_scanned": run.n_pairs_scanned}


def parse_args_v12(argv):
    from .v12 import OptionsV12

    o, args = build_parser_v12().parse_args(list(argv))
    for bad in ("randomize", "stranded", "strandpref"):
        if getattr(o, bad):
            raise SystemExit("option --%s is not supported by this build" % bad)
    if o.bam or o.reads2samples or o.system:
        raise SystemExit("-B, -r and -S are not supported by this build")
    opt = OptionsV12(genome=o.genome, name=o.name, prefix=o.prefix, min_uniq_qual=o.min_uniq_qual, asize=o.asize,
                     margin=o.margin, maxdist=o.maxdist, wiggle=o.wiggle, noncanonical=o.noncanonical, allhits=o.allhits,
                     halfunique=o.halfunique, report_nobridges=o.report_nobridges, reads=o.reads, stats=o.stats,
                     batch_pairs=o.batch_pairs, device=o.device)
    return opt, args, o


def main_v12(argv) -> int:
    opt, args, raw = parse_args_v12(argv)
    if raw.version:
        print("find_circ.py version %s (v1.2 command line)" % VERSION)
        return 0
    if not opt.genome:
        print("need to specify either model system database (-S) or genome FASTA file (-G).")
        return 1
    try:
        out = run_v12_to_strings(opt, args[0] if args else None)
    except Exception:
        sys.stderr.write(traceback.format_exc())
        return 1
    sys.stdout.write(out["bed"])
    if opt.reads:
        with open(opt.reads, "w") as fh:
            fh.write(out["reads"])
    else:
        sys.stderr.write(out["reads"])
    with open(opt.stats, "w") as fh:
        fh.write(out["stats"])
    return 0


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if wants_v12(argv):
        return main_v12(argv)
    opt, args, raw = parse_args(argv)
    if raw.version:
        print("find_circ.py version %s (B200-native breakpoint scan + junction merge)" % VERSION)
        return 0
    if not opt.genome:
        print("need to specify either model system database (-S) or genome FASTA file (-G).")
        return 1
    if raw.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        # one process per GPU: hand the same command line to torchrun (rank 0 writes the outputs)
        import socket
        import subprocess

        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        script = os.path.abspath(sys.argv[0]) if sys.argv and os.path.exists(sys.argv[0]) else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "find_circ.py")
        return subprocess.call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(raw.gpus), "--master-addr",
                                "127.0.0.1", "--master-port", str(port), script] + list(argv))
    dist, torch_dev = distributed_context()
    rank = dist.get_rank() if dist is not None else 0
    if dist is not None:
        from . import parallel

        opt.device = torch_dev.index if torch_dev.type == "cuda" else 0
        if torch_dev.type == "cuda":
            parallel.bind_to_gpu_numa(opt.device)
    if rank != 0:
        # the other ranks scan their part of the input and hand everything to rank 0
        try:
            run_distributed(opt, args[0] if args else None, dist, torch_dev)
            return 0
        except Exception:
            sys.stderr.write(traceback.format_exc())
            return 1
    if not os.path.isdir(opt.output):
        os.makedirs(opt.output)
    fmt = "%(asctime)-20s\t%(levelname)s\t%(name)s\t%(message)s"
    logging.basicConfig(level=logging.INFO, format=fmt, filename=os.path.join(opt.output, "run.log"), filemode="w")
    log = logging.getLogger("find_circ")
    log.info("find_circ %s invoked as '%s'" % (VERSION, " ".join(sys.argv)))
    files = {
        "circs": open(os.path.join(opt.output, "circ_splice_sites.bed"), "w"),
        "lins": open(os.path.join(opt.output, "lin_splice_sites.bed"), "w"),
        "reads": GzipMembers(os.path.join(opt.output, "spliced_reads.fastq.gz")),
        "multi": open(os.path.join(opt.output, "multi_events.tsv"), "w"),
    }
    if opt.stdout and opt.stdout in files:
        files[opt.stdout].write("# redirected to stdout\n")
        files[opt.stdout].close()
        files[opt.stdout] = sys.stdout
        log.info("redirected %s to stdout" % opt.stdout)
    path = args[0] if args else None
    log.info("reading from %s" % (path or "stdin"))
    t0 = time.time()
    try:
        out = run_distributed(opt, path, dist, torch_dev) if dist is not None else run_to_strings(opt, path)
    except KeyboardInterrupt:
        logging.warning("KeyboardInterrupt by user")
        return 1
    except Exception:
        exc = traceback.format_exc()
        logging.error("Unhandled exception raised while processing input")
        logging.error(exc)
        sys.stderr.write(exc)
        return 1
    t1 = time.time()
    n = out["n_fragments"]
    txt = "processed %.2fM (paired or single end) reads in %.1f minutes (overall %.2fk reads/second on average)" % (
        n / 1e6, (t1 - t0) / 60.0, n / max(t1 - t0, 1e-9) / 1000.0)
    log.info(txt)
    log.info("anchor pairs scanned on the GPU: %d, seconds inside GPU calls: %.3f" % (out["n_pairs_scanned"], out["seconds_gpu_calls"]))
    if not opt.silent and not opt.stdout:
        print("#", txt)
        print("# results stored in '%s'" % opt.output)
    log.info("run finished")
    for line in out["counters"].splitlines():
        log.info(line)
    files["circs"].write(out["circ"])
    files["lins"].write(out["lin"])
    files["reads"].write(out["reads"])
    files["multi"].write(out["multi"])
    if opt.test:
        with open(os.path.join(opt.output, "test_results.tsv"), "w") as fh:
            fh.write(out["test"])
    for f in files.values():
        if f is not sys.stdout:
            f.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
