"""
The drop-in as a PROCESS: ./find_circ.py run the way the reference is run (file argument, stdin pipe, BAM input),
its output directory compared with the reference goldens.  Needs the GPU; the BAM reader itself is also checked on CPU.
"""
import gzip
import os
import struct
import subprocess
import sys

import pytest

from conftest import GOLDEN, ROOT
from oracle import find_circ_oracle as O


def sam_to_bam(sam_path, bam_path, bgzf=0):
    """minimal BAM writer (test helper).  bgzf=0: plain gzip members of 60 kB (valid for every gzip reader, but without the block
    sizes of BGZF: the library reads it through one zlib stream); bgzf=n: real BGZF with blocks of n bytes"""
    names, lens, recs = [], [], []
    text = ""
    for line in open(sam_path):
        if line.startswith("@"):
            text += line
            if line.startswith("@SQ"):
                d = dict(x.split(":", 1) for x in line.rstrip("\n").split("\t")[1:])
                names.append(d["SN"])
                lens.append(int(d["LN"]))
        elif line.strip():
            recs.append(line.rstrip("\n").split("\t"))
    out = bytearray(b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(names)))
    for n, ln in zip(names, lens):
        out += struct.pack("<i", len(n) + 1) + n.encode() + b"\x00" + struct.pack("<i", ln)
    code = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    ops = {c: i for i, c in enumerate("MIDNSHP=X")}
    import re

    for f in recs:
        qname = f[0].encode() + b"\x00"
        tid = names.index(f[2]) if f[2] != "*" else -1
        cig = [] if f[5] == "*" else [(int(n), ops[c]) for n, c in re.findall(r"(\d+)([MIDNSHP=X])", f[5])]
        seq = "" if f[9] == "*" else f[9]
        sb = bytearray()
        for i in range(0, len(seq), 2):
            hi = code[seq[i]]
            lo = code[seq[i + 1]] if i + 1 < len(seq) else 0
            sb.append((hi << 4) | lo)
        qual = bytes([0xFF] * len(seq)) if f[10] == "*" else bytes(ord(c) - 33 for c in f[10])
        tags = bytearray()
        for t in f[11:]:
            tag, typ, val = t.split(":", 2)
            if typ == "i":
                v = int(val)
                tags += tag.encode() + (b"c" + struct.pack("<b", v) if -128 <= v < 128 else b"i" + struct.pack("<i", v))
            else:
                tags += tag.encode() + b"Z" + val.encode() + b"\x00"
        body = struct.pack("<iiBBHHHiiii", tid, int(f[3]) - 1, len(qname), int(f[4]), 4680, len(cig), int(f[1]), len(seq), -1, -1, 0)
        body += qname + b"".join(struct.pack("<I", (n << 4) | o) for n, o in cig) + bytes(sb) + qual + bytes(tags)
        out += struct.pack("<i", len(body)) + body
    with open(bam_path, "wb") as fh:
        if bgzf:
            fh.write(bgzf_compress(bytes(out), bgzf))
        else:
            for i in range(0, len(out), 60000):
                fh.write(gzip.compress(bytes(out[i : i + 60000])))


def bgzf_compress(data: bytes, block: int = 60000) -> bytes:
    """BGZF as samtools / htslib write it: gzip members of at most 64 KiB whose extra field 'BC' holds the member's size, closed
    by the empty end-of-file member"""
    import zlib

    out = bytearray()
    for i in list(range(0, len(data), block)) + [None]:
        chunk = b"" if i is None else data[i : i + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(chunk) + c.flush()
        size = 18 + len(body) + 8
        out += struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, ord("B"), ord("C"), 2, size - 1)
        out += body + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
    return bytes(out)


def test_bam_reader_equals_sam_reader(tmp_path):
    from find_circ2_b200 import samio

    sam = os.path.join(GOLDEN, "synth_a", "input.sam")
    bam = str(tmp_path / "input.bam")
    sam_to_bam(sam, bam)
    n1, l1, r1 = samio.read_sam(open(sam))
    n2, l2, r2 = samio.read_bam(open(bam, "rb"))
    assert n1 == n2 and l1 == l2
    a, b = list(r1), list(r2)
    assert len(a) == len(b) > 2000
    for x, y in zip(a, b):
        assert (x.qname, x.flag, x.tid, x.pos, x.cigar, x.seq, x.qual, x.AS, x.XS, x.aend) == (
            y.qname, y.flag, y.tid, y.pos, y.cigar, y.seq, y.qual, y.AS, y.XS, y.aend)


def _compare_dir(out_dir, ref_dir):
    rd = lambda n: open(os.path.join(ref_dir, n)).read()  # noqa: E731
    assert O.canonical_bed(open(os.path.join(out_dir, "circ_splice_sites.bed")).read()) == O.canonical_bed(rd("circ_splice_sites.bed"))
    assert O.canonical_bed(open(os.path.join(out_dir, "lin_splice_sites.bed")).read()) == O.canonical_bed(rd("lin_splice_sites.bed"))
    assert gzip.open(os.path.join(out_dir, "spliced_reads.fastq.gz"), "rt").read() == rd("spliced_reads.fastq")
    assert O.canonical_multi(open(os.path.join(out_dir, "multi_events.tsv")).read()) == O.canonical_multi(rd("multi_events.tsv"))
    log = open(os.path.join(out_dir, "run.log")).read().split("\n")
    k = [i for i, l in enumerate(log) if l.endswith("run finished")][0]
    counters = "".join(l.split("\t")[-1] + "\n" for l in log[k + 1 :] if "=" in l)
    assert counters == rd("counters.txt")


@pytest.mark.gpu
@pytest.mark.parametrize("how", ["file", "stdin", "bam"])
def test_find_circ_process(tmp_path, how):
    case = os.path.join(GOLDEN, "synth_a")
    ref = os.path.join(case, "ref_default")
    out = str(tmp_path / "run")
    cmd = [sys.executable, os.path.join(ROOT, "find_circ.py"), "-G", os.path.join(case, "genome.fa"), "-n", "test", "-o", out, "-q"]
    sam = os.path.join(case, "input.sam")
    if how == "file":
        r = subprocess.run(cmd + [sam], capture_output=True, text=True)
    elif how == "stdin":
        r = subprocess.run(cmd, stdin=open(sam), capture_output=True, text=True)
    else:
        bam = str(tmp_path / "input.bam")
        sam_to_bam(sam, bam)
        r = subprocess.run(cmd + [bam], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    _compare_dir(out, ref)


@pytest.mark.gpu
def test_find_circ_process_errors(tmp_path):
    exe = [sys.executable, os.path.join(ROOT, "find_circ.py")]
    r = subprocess.run(exe + ["-o", str(tmp_path / "x")], capture_output=True, text=True, stdin=subprocess.DEVNULL)
    assert r.returncode == 1 and "need to specify" in r.stdout          # find_circ.py:438-440
    r = subprocess.run(exe + ["-v"], capture_output=True, text=True)
    assert r.returncode == 0 and "version" in r.stdout                  # find_circ.py:416-418
    # unknown chromosome in the alignments -> the run aborts with exit 1 (KeyError at find_circ.py:193)
    case = os.path.join(GOLDEN, "kat3")
    other = os.path.join(GOLDEN, "cdr1as", "genome.fa")
    r = subprocess.run(exe + ["-G", other, "-o", str(tmp_path / "y"), "-q", os.path.join(case, "input.sam")], capture_output=True, text=True)
    assert r.returncode == 1 and "KeyError" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("case_name,tag", [("synth_b", "noop"), ("synth_b", "stdout_circs"), ("synth_b", "stdout_reads"),
                                           ("kat3", "selftest"), ("synth_a", "selftest")])
def test_find_circ_process_rare_switches(tmp_path, case_name, tag):
    """--noop (alignments are only grouped, find_circ.py:1554-1558), --stdout NAME (that output goes to stdout, its file
    keeps one comment line, find_circ.py:453-458), -t / --chunk-size (progress on stderr), --test (test_results.tsv,
    find_circ.py:1148-1273, 1380-1394) -- against reference runs"""
    case = os.path.join(GOLDEN, case_name)
    ref = os.path.join(case, "ref_" + tag)
    argv = open(os.path.join(ref, "cmdline.txt")).read().split("\n")[0].split()
    out = str(tmp_path / "run")
    cmd = [sys.executable, os.path.join(ROOT, "find_circ.py"), "-G", os.path.join(case, "genome.fa"), "-o", out, "-q"] + argv
    r = subprocess.run(cmd + [os.path.join(case, "input.sam")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    rd = lambda d, n: open(os.path.join(d, n)).read()  # noqa: E731
    got = {"circs": rd(out, "circ_splice_sites.bed"), "lins": rd(out, "lin_splice_sites.bed"),
           "reads": gzip.open(os.path.join(out, "spliced_reads.fastq.gz"), "rt").read(), "multi": rd(out, "multi_events.tsv")}
    if "--stdout" in argv:
        name = argv[argv.index("--stdout") + 1]
        assert got[name] == "# redirected to stdout\n"  # (for `reads` a complete gzip stream; the reference leaves it truncated)
        got[name] = r.stdout
    else:
        assert r.stdout == ""
    log = rd(out, "run.log").split("\n")
    k = [i for i, l in enumerate(log) if l.endswith("run finished")][0]
    counters = "".join(l.split("\t")[-1] + "\n" for l in log[k + 1:] if "=" in l)
    import helpers as H

    tests = rd(out, "test_results.tsv") if "--test" in argv else None
    H.compare_outputs(got["circs"], got["lins"], got["reads"], got["multi"], counters, ref, argv, tests)


@pytest.mark.parametrize("block,threads", [(60000, "1"), (60000, "4"), (700, "3"), (65280, "16")])
def test_bgzf_on_several_threads_gives_the_same_text(tmp_path, monkeypatch, block, threads):
    """BAM written as real BGZF is inflated and formatted on several threads (csrc/bam.cu): the SAM text must be the text the
    one-stream reader gives for the same records, whatever the block size (records straddle members and batches of 512
    members) and the thread count"""
    from find_circ2_b200.ingest import BamText

    sam = os.path.join(GOLDEN, "synth_a", "input.sam")
    big = str(tmp_path / "big.sam")
    lines = open(sam).read().splitlines(True)
    head = [ln for ln in lines if ln.startswith("@")]
    body = [ln for ln in lines if not ln.startswith("@")]
    with open(big, "w") as fh:
        fh.writelines(head)
        for k in range(12):  # ~30 000 records: more than one batch of members at the small block size
            fh.writelines(ln.replace("\t", "_%d\t" % k, 1) for ln in body)
    plain, bg = str(tmp_path / "plain.bam"), str(tmp_path / "bgzf.bam")
    sam_to_bam(big, plain)
    sam_to_bam(big, bg, bgzf=block)
    monkeypatch.setenv("FC_BAM_THREADS", threads)

    def text(path, n):
        t = BamText(path)
        names = list(t.names)
        out = []
        while True:
            piece = t.read(n)
            if not piece:
                break
            assert piece.endswith(b"\n")
            out.append(piece)
        t.close()
        return names, b"".join(out)

    n1, t1 = text(plain, 1 << 20)
    n2, t2 = text(bg, 1 << 20)
    n3, t3 = text(bg, 1 << 16)  # small reads: pieces are cut on line boundaries
    assert n1 == n2 == n3 and len(t1) > 3000000
    assert t2 == t1 and t3 == t1
    # a file that ends inside a member, and one whose member is damaged
    raw = open(bg, "rb").read()
    for bad in (raw[: len(raw) // 2], raw[:5000] + bytes([raw[5000] ^ 0x55]) + raw[5001:]):
        p = str(tmp_path / "bad.bam")
        open(p, "wb").write(bad)
        with pytest.raises((IOError, OSError, RuntimeError)):
            t = BamText(p)
            while t.read(1 << 20):
                pass
