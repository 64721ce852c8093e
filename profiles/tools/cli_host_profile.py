#!/usr/bin/env python3
"""Where does the host time of the drop-in go?  Builds a synthetic SAM file (2-segment reads of the bench workload), runs
the whole CLI path on it under cProfile and prints the top of the profile.  usage: cli_host_profile.py [n_reads]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from find_circ2_b200 import cli, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
g = synth.make_genome([5000000] * 4, seed=1, n_frac=0.005, n_run=(50, 5000), soft_frac=0.0)
J = synth.plant_junctions(g, 2000, 100, seed=2, span=(200, 50000), margin=400)
t = synth.make_pairs(g, J, n, read_len=100, asize=20, seed=3, error_rate=0.005, zipf=1.0, frac_decoy=0.1, frac_nonuniq=0.02,
                     frac_edge=0.01)
with tempfile.TemporaryDirectory() as tmp:
    fa, sam = os.path.join(tmp, "g.fa"), os.path.join(tmp, "in.sam")
    g.write_fasta(fa)
    with open(sam, "w") as fh:
        fh.write(synth.sam_header(g))
        for i in range(len(t)):
            fh.writelines(synth.bwa_records_for_pair(g, t, i, "r%d" % i))
    argv = ["-G", fa, "-a", "20", "-n", "prof", "-o", os.path.join(tmp, "out"), "-q", sam]
    t0 = time.perf_counter()
    cli.main(argv)  # warm-up: CUDA context, genome upload
    print("first run  %.2f s" % (time.perf_counter() - t0))
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    cli.main(["-G", fa, "-a", "20", "-n", "prof", "-o", os.path.join(tmp, "out2"), "-q", sam])
    pr.disable()
    dt = time.perf_counter() - t0
    print("second run %.2f s  -> %.0f reads/s" % (dt, n / dt))
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
    pstats.Stats(pr).sort_stats("cumulative").print_stats("find_circ2_b200|numpy", 30)
