"""The reference's own host (main.c, gfa-io.c, bseq.c, format.c, kthread.c ... compiled unmodified) linked against
libmgb200.so must print the same bytes as the reference: the drop-in claim of include/mgb200.h."""
import os
import subprocess

import pytest

import cases
import mgtest as T

pytestmark = pytest.mark.gpu
DROPIN = os.path.join(T.REPO, "oracle", "_ref", "minigraph_b200")


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/minigraph_b200 not built (make -C oracle dropin)")
def test_dropin_binary_matches_reference(workdir):
    hap, reads = os.path.join(workdir, "mt.hap.fa"), os.path.join(workdir, "mt.reads.fa")
    T.sim_mt_haps(hap)
    T.sim_reads(hap, reads, 24, 10000, "ont", 11)
    gfa = os.path.join(T.FIX, "MT.gfa")
    got = subprocess.run([DROPIN, "-cx", "lr", "-t", "4", gfa, reads], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    want = cases.golden("c2_MT_24x10k_ont_s11.lr.gaf")
    assert got == want, cases.first_diff(got, want)
    got = subprocess.run([DROPIN, "-cx", "lr", gfa, os.path.join(T.FIX, "MT-orangA.fa")], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    assert got == cases.golden("c1_MT_orangA.lr.gaf")
