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


REF_BIN = os.path.join(T.REPO, "oracle", "_ref", "minigraph")


@pytest.mark.skipif(not (os.path.exists(DROPIN) and os.path.exists(REF_BIN)), reason="oracle/_ref binaries not built")
def test_dropin_binary_paired_short_reads(workdir):
    """-x sr with two query files: read pairs become two-segment fragments (gmap.c:70-97) and reach the GPU through
    mg_map_batch_frag(); same bytes as the reference binary"""
    hap, frags = os.path.join(workdir, "mt.hap.fa"), os.path.join(workdir, "mt.frag.fa")
    T.sim_mt_haps(hap)
    T.sim_reads(hap, frags, 300, 500, "hifi", 67)
    names, seqs = T.read_fasta(frags)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    r1, r2 = os.path.join(workdir, "sr_1.fa"), os.path.join(workdir, "sr_2.fa")
    with open(r1, "wb") as f1, open(r2, "wb") as f2:
        for nm, s in zip(names, seqs):
            f1.write(b">" + nm + b"/1\n" + s[:150] + b"\n")
            f2.write(b">" + nm + b"/2\n" + s[-150:].translate(comp)[::-1] + b"\n")
    gfa = os.path.join(T.FIX, "MT.gfa")
    want = subprocess.run([REF_BIN, "-x", "sr", "-t", "4", gfa, r1, r2], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    got = subprocess.run([DROPIN, "-x", "sr", "-t", "4", gfa, r1, r2], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    assert len(want) > 1000
    assert got == want, cases.first_diff(got, want)


@pytest.mark.skipif(not (os.path.exists(DROPIN) and os.path.exists(REF_BIN)), reason="oracle/_ref binaries not built")
def test_dropin_other_consumers_of_the_boundary(workdir):
    """the reference's other users of mg_index()/mg_map() -- incremental graph generation (-cxggs: ggen.c:36 maps every sample
    through mg_map() from kt_for threads and re-indexes the grown graph), --call (asm-call.c:21) and --cov (cal_cov.c:8) -- run
    unchanged on the library and print the reference's bytes"""
    hum, chimp, orang = (os.path.join(T.FIX, f) for f in ("MT-human.fa", "MT-chimp.fa", "MT-orangA.fa"))
    gfa = os.path.join(T.FIX, "MT.gfa")
    hap, reads = os.path.join(workdir, "mt.hap.fa"), os.path.join(workdir, "mt.cov.fa")
    T.sim_mt_haps(hap)
    T.sim_reads(hap, reads, 60, 8000, "ont", 73)
    pre = os.path.join(workdir, "svcall")
    T.sim_graph(pre, 200000, 3, 41)  # --call wants bubbles: an SV graph and one of its haplotypes as the assembly (a 200 kb query, -x asm)
    hn, hs = T.read_fasta(pre + ".hap.fa")
    sample = os.path.join(workdir, "svcall.h2.fa")
    T.write_fasta(sample, hn[2:3], hs[2:3])
    for args in (["-cxggs", "-t", "3", hum, chimp, orang], ["-cxasm", "--call", "-t", "2", pre + ".gfa", sample], ["-cxasm", "--cov", gfa, chimp],
                 ["-cxlr", "--cov", "-t", "4", gfa, reads]):
        want = subprocess.run([REF_BIN] + args, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
        got = subprocess.run([DROPIN] + args, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
        assert len(want) > 100, args
        assert got == want, (args, cases.first_diff(got, want))


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/minigraph_b200 not built (make -C oracle dropin)")
def test_dropin_binary_on_several_devices(workdir):
    """MGB_DEVICES: the reference host drives every listed GPU through the one mg_map_batch_frag() call per mini-batch"""
    import torch
    hap, reads = os.path.join(workdir, "mt.hap.fa"), os.path.join(workdir, "mt.reads.fa")
    T.sim_mt_haps(hap)
    T.sim_reads(hap, reads, 24, 10000, "ont", 11)
    env = dict(os.environ, MGB_DEVICES="0,1" if torch.cuda.device_count() >= 2 else "0,0")
    got = subprocess.run([DROPIN, "-cx", "lr", "-t", "4", os.path.join(T.FIX, "MT.gfa"), reads], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env).stdout
    want = cases.golden("c2_MT_24x10k_ont_s11.lr.gaf")
    assert got == want, cases.first_diff(got, want)
