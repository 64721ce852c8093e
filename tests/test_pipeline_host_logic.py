"""
CPU coverage of the product's HOST logic (find_circ2_b200/pipeline.py, cli.py, samio.py): with the GPU engine replaced
by tests/fake_engine.py the pipeline must reproduce the reference goldens.  The real engine is covered by
tests/test_gpu_pipeline.py (-m gpu).
"""
import os

import pytest

from conftest import golden_cases, golden_ids
import helpers as H
from fake_engine import FakeEngine
from oracle import find_circ_oracle as O


@pytest.mark.parametrize("native", [False, True], ids=["python-ingest", "native-ingest"])
@pytest.mark.parametrize("case_dir,ref_dir,argv", golden_cases(), ids=golden_ids())
def test_host_logic_matches_reference(case_dir, ref_dir, argv, native):
    from find_circ2_b200 import cli

    opt = cli.parse_args(["-G", os.path.join(case_dir, "genome.fa")] + argv)[0]
    opt.batch_pairs = 211
    eng = FakeEngine(0, opt.asize, opt.margin, opt.maxdist, opt.noncanonical, opt.strandpref)
    eng.load_genome_fasta(opt.genome)
    if native and (opt.allhits or opt.noop or opt.test):
        pytest.skip("--all-hits, --noop and --test use the python ingest")
    out = cli.run_to_strings(opt, os.path.join(case_dir, "input.sam"), engine=eng, native=native)
    H.compare_outputs(out["circ"], out["lin"], out["reads"], out["multi"], out["counters"], ref_dir, argv, out["test"])


@pytest.mark.parametrize("threads", [2, 5])
@pytest.mark.parametrize("case,tag", [("synth_a", "default"), ("synth_a", "a20"), ("synth_b", "default"), ("kat3", "default"),
                                      ("synth_a", "uniq0_half_nobridge"), ("synth_a", "nolinear_nomulti")])
def test_native_ingest_on_parser_threads(case, tag, threads):
    """the SAM text of a chunk is cut into pieces on fragment boundaries and parsed by several threads (own parser handle and
    fragment numbering each): same outputs as the reference, whatever the cut"""
    from conftest import GOLDEN, golden_cases
    from find_circ2_b200 import cli

    case_dir, ref_dir = os.path.join(GOLDEN, case), os.path.join(GOLDEN, case, "ref_" + tag)
    argv = [a for c, r, a in golden_cases() if r == ref_dir][0]
    opt = cli.parse_args(["-G", os.path.join(case_dir, "genome.fa")] + argv)[0]
    opt.batch_pairs, opt.ingest_threads, opt.ingest_piece_bytes = 149, threads, 1
    eng = FakeEngine(0, opt.asize, opt.margin, opt.maxdist, opt.noncanonical, opt.strandpref)
    eng.load_genome_fasta(opt.genome)
    out = cli.run_to_strings(opt, os.path.join(case_dir, "input.sam"), engine=eng, native=True)
    H.compare_outputs(out["circ"], out["lin"], out["reads"], out["multi"], out["counters"], ref_dir, argv, out["test"])


@pytest.mark.parametrize("case,tag", [("synth_a", "default"), ("synth_a", "a20"), ("synth_a", "m4_d3"), ("synth_b", "default"),
                                      ("kat3", "default"), ("cdr1as", "default"), ("synth_a", "known")])
def test_native_ingest_from_bam(tmp_path, case, tag):
    """BAM input goes through the native ingest as well: csrc/bam.cu inflates the file and hands the records to the C++
    parser as SAM text (find_circ.py:461-469 reads BAM through pysam)"""
    from conftest import GOLDEN, golden_cases
    from find_circ2_b200 import cli
    from test_cli_files import sam_to_bam

    case_dir, ref_dir = os.path.join(GOLDEN, case), os.path.join(GOLDEN, case, "ref_" + tag)
    argv = [a for c, r, a in golden_cases() if r == ref_dir][0]
    bam = str(tmp_path / "input.bam")
    # (real BGZF with small blocks for half of the cases -- inflated on several threads --, plain gzip members for the others)
    sam_to_bam(os.path.join(case_dir, "input.sam"), bam, bgzf=3000 if tag in ("default", "known") else 0)
    opt = cli.parse_args(["-G", os.path.join(case_dir, "genome.fa")] + argv)[0]
    opt.batch_pairs = 131
    assert cli.native_ok(opt, bam)
    eng = FakeEngine(0, opt.asize, opt.margin, opt.maxdist, opt.noncanonical, opt.strandpref)
    eng.load_genome_fasta(opt.genome)
    out = cli.run_to_strings(opt, bam, engine=eng)
    H.compare_outputs(out["circ"], out["lin"], out["reads"], out["multi"], out["counters"], ref_dir, argv, out["test"])


def test_native_ingest_from_stdin(monkeypatch):
    """SAM text piped in (`bwa mem ... | find_circ.py`, find_circ.py:461-469) goes through the native ingest too: the
    header is read off the stream, the body is parsed in chunks"""
    import io
    import sys

    from conftest import GOLDEN
    from find_circ2_b200 import cli

    case_dir = os.path.join(GOLDEN, "synth_a")
    ref_dir = os.path.join(case_dir, "ref_default")
    opt = cli.parse_args(["-G", os.path.join(case_dir, "genome.fa"), "-n", "test"])[0]
    opt.batch_pairs = 97
    assert cli.native_ok(opt, None) and cli.native_ok(opt, "-") and cli.native_ok(opt, "x.bam")
    eng = FakeEngine(0, opt.asize, opt.margin, opt.maxdist, opt.noncanonical, opt.strandpref)
    eng.load_genome_fasta(opt.genome)

    class Stdin(object):
        buffer = io.BufferedReader(io.BytesIO(open(os.path.join(case_dir, "input.sam"), "rb").read()))

    monkeypatch.setattr(sys, "stdin", Stdin())
    out = cli.run_to_strings(opt, None, engine=eng)
    rd = lambda n: open(os.path.join(ref_dir, n)).read()  # noqa: E731
    assert O.canonical_bed(out["circ"]) == O.canonical_bed(rd("circ_splice_sites.bed"))
    assert O.canonical_bed(out["lin"]) == O.canonical_bed(rd("lin_splice_sites.bed"))
    assert out["reads"] == rd("spliced_reads.fastq")
    assert out["counters"] == rd("counters.txt")


def test_gzip_members_writer(tmp_path):
    """spliced_reads.fastq.gz is written as several gzip members compressed on worker threads: any gzip reader must see
    the concatenated text; an empty run still leaves a valid file"""
    import gzip

    from find_circ2_b200.cli import GzipMembers

    rec = "@r%d circ_000001 \nACGTACGTAC\n+r%d circ_000001 \nIIIIIIIIII\n"
    text = "".join(rec % (k, k) for k in range(120000))
    p = str(tmp_path / "reads.fastq.gz")
    w = GzipMembers(p, threads=3)
    w.BLOCK = 1 << 20  # several members
    for i in range(0, len(text), 777777):
        w.write(text[i:i + 777777])
    w.close()
    assert gzip.open(p, "rt").read() == text
    raw = open(p, "rb").read()
    assert raw.count(b"\x1f\x8b\x08") >= 4  # more than one member
    q = str(tmp_path / "empty.fastq.gz")
    GzipMembers(q).close()
    assert gzip.open(q, "rt").read() == ""


def test_bam_text_stream_errors_and_close(tmp_path):
    """csrc/bam.cu through ingest.BamText: header names, a truncated file raises, close() with chunks still queued returns"""
    from conftest import GOLDEN
    from find_circ2_b200.ingest import BamText
    from test_cli_files import sam_to_bam

    sam = os.path.join(GOLDEN, "synth_a", "input.sam")
    bam = str(tmp_path / "a.bam")
    sam_to_bam(sam, bam)
    b = BamText(bam)
    assert b.names == ["chr1", "chr2", "chr3"] and b.lengths == [30000, 22000, 9000]
    first = b.read(1 << 17)
    n_records = sum(1 for line in open(sam) if not line.startswith("@"))
    text = first
    while True:
        c = b.read(1 << 17)
        if not c:
            break
        text += c
    assert text.count(b"\n") == n_records and b.read(10) == b""
    b.close()
    # closing early (the reader thread holds queued chunks) must not hang
    b = BamText(bam)
    b.read(1 << 17)
    b.close()
    # truncated file: the stream fails instead of ending silently
    raw = open(bam, "rb").read()
    cut = str(tmp_path / "cut.bam")
    open(cut, "wb").write(raw[: len(raw) * 2 // 3])
    b = BamText(cut)
    with pytest.raises(IOError):
        while b.read(1 << 17):
            pass
    b.close()
    with pytest.raises(IOError):
        BamText(sam)  # not a BAM file


@pytest.mark.parametrize("threads", [1, 3])
def test_native_ingest_chunk_boundaries_and_a_fragment_larger_than_the_head_room(tmp_path, threads):
    """process_native reads the stream in chunks behind 1 MB of head room for the unfinished fragment a chunk leaves over.
    Small chunks put many fragments across boundaries; one fragment of 1.3 MB of text (thousands of unmapped records under
    one name) does not fit the head room and takes the concatenating way.  Outputs must equal those of the
    python ingest on the same file."""
    from conftest import GOLDEN
    from find_circ2_b200 import cli, pipeline, samio

    case_dir = os.path.join(GOLDEN, "synth_a")
    lines = open(os.path.join(case_dir, "input.sam")).read().splitlines(True)
    head = [ln for ln in lines if ln.startswith("@")]
    body = [ln for ln in lines if not ln.startswith("@")]
    # (one mapped record, then thousands of unmapped ones under the same name: counted and skipped, find_circ.py:1466-1467)
    first = [ln for ln in body if ln.split("\t")[2] != "*"][0]
    unmapped = "giant\t4\t*\t0\t0\t*\t*\t0\t0\t" + "ACGT" * 25 + "\t" + "I" * 100 + "\n"
    giant = ["giant\t" + first.split("\t", 1)[1]] + [unmapped] * (1300000 // len(unmapped))
    sam = str(tmp_path / "with_giant.sam")
    with open(sam, "w") as fh:
        fh.writelines(head + body[:3000] + giant + body[3000:])
    argv = ["-a", "15", "-n", "t", "-q"]
    outs = []
    for native in (False, True):
        opt = cli.parse_args(["-G", os.path.join(case_dir, "genome.fa")] + argv)[0]
        opt.batch_pairs, opt.ingest_threads, opt.ingest_piece_bytes = 211, threads, 1 << 15
        eng = FakeEngine(0, opt.asize, opt.margin, opt.maxdist, opt.noncanonical, opt.strandpref)
        eng.load_genome_fasta(opt.genome)
        if not native:
            outs.append(cli.run_to_strings(opt, sam, engine=eng, native=False))
            continue
        names = samio.sam_header_names(sam)
        run = pipeline.Run(opt, names, eng)
        with open(sam, "rb") as fh:
            run.process_native(fh, chunk_bytes=200000)
        run.finalize()
        outs.append({"circ": run.bed_text(0), "lin": run.bed_text(1), "reads": run.reads_text(), "multi": run.multi_text(),
                     "counters": run.counters_text()})
        run.close()
    for k in ("circ", "lin", "reads", "multi", "counters"):
        assert outs[0][k] == outs[1][k], k
    assert outs[1]["circ"].count("\n") > 5
