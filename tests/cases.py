"""Parity cases shared by the CPU (hostsim) and GPU (libmgb200) test modules."""
import hashlib
import os

import mgtest as T

G = os.path.join(T.REPO, "tests", "golden")


def golden(name):
    with open(os.path.join(G, name), "rb") as f:
        return f.read()


def first_diff(a, b):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            fx, fy = x.split(b"\t"), y.split(b"\t")
            for j, (p, q) in enumerate(zip(fx, fy)):
                if p != q:
                    return "line %d field %d: %r != %r" % (i, j, p[:120], q[:120])
            return "line %d: field count %d != %d" % (i, len(fx), len(fy))
    return "line count %d != %d" % (len(la), len(lb))


def check_gaf(lib, gfa, fasta, preset, want, flag_extra=0):
    names, seqs = T.read_fasta(fasta)
    got, st = T.gaf_with_engine(lib, gfa, names, seqs, preset, flag_extra=flag_extra)
    assert got == want, first_diff(got, want)
    return st


def case_c1(lib, workdir):
    """config 1: test/MT.gfa <- test/MT-orangA.fa, -cx lr; md5 pinned in SURVEY.md section 8c."""
    st = check_gaf(lib, os.path.join(T.FIX, "MT.gfa"), os.path.join(T.FIX, "MT-orangA.fa"), "lr", golden("c1_MT_orangA.lr.gaf"))
    assert hashlib.md5(golden("c1_MT_orangA.lr.gaf")).hexdigest() == "22bf23ebe2039e8353f56f4a324a2eaa"
    check_gaf(lib, os.path.join(T.FIX, "MT.gfa"), os.path.join(T.FIX, "MT-chimp.fa"), "lr", golden("c1_MT_chimp.lr.gaf"))
    return st


def case_c2(lib, workdir):
    hap, reads = os.path.join(workdir, "mt.hap.fa"), os.path.join(workdir, "mt.reads.fa")
    T.sim_mt_haps(hap)
    T.sim_reads(hap, reads, 24, 10000, "ont", 11)
    return check_gaf(lib, os.path.join(T.FIX, "MT.gfa"), reads, "lr", golden("c2_MT_24x10k_ont_s11.lr.gaf"))


def case_c3(lib, workdir):
    pre, reads = os.path.join(workdir, "sv"), os.path.join(workdir, "sv.reads.fa")
    T.sim_graph(pre, 300000, 3, 7)
    T.sim_reads(pre + ".hap.fa", reads, 24, 15000, "ont", 5)
    return check_gaf(lib, pre + ".gfa", reads, "lr", golden("c3_sv300k_h3_s7_24x15k_ont_s5.lr.gaf"))


def case_c4(lib, workdir):
    reads = os.path.join(workdir, "mth.reads.fa")
    T.sim_reads(os.path.join(T.FIX, "MT-human.fa"), reads, 12, 20000, "hifi", 13, circular=True)
    return check_gaf(lib, os.path.join(T.FIX, "MT-human.fa"), reads, "asm", golden("c4_MThuman_12x20k_hifi_s13.asm.gaf"))


def case_edge(lib, workdir):
    """empty, tiny, all-N, unmappable and lower-case-free reads: same objects as the reference (map-algo.c:356-360)."""
    import ctypes as C
    from minigraph_b200 import capi, options
    gfa = os.path.join(T.FIX, "MT.gfa")
    _, hs = T.read_fasta(os.path.join(T.FIX, "MT-human.fa"))
    names = [b"empty", b"tiny", b"allN", b"random", b"short_ok", b"with_N"]
    seqs = [b"", b"ACGT", b"N" * 500, (b"ACGTTGCA" * 200)[:1500], hs[0][1000:1300], hs[0][2900:3300]]
    got, _, _ = T.map_with_engine(lib, gfa, names, seqs, "lr")
    assert got[0] is None                      # qlen == 0 -> no result object
    for r in got[1:4]:
        assert r is not None and r["n_gc"] == 0
    if T.have_ref():
        want, _ = T.map_with_ref(gfa, names, seqs, "lr")
        for i, (a, b) in enumerate(zip(want, got)):
            assert T.diff_results(a, b) is None, (i, T.diff_results(a, b))


def case_struct_random(lib, workdir, n_reads=150, seed=23):
    """field-by-field comparison of mg_gchains_t (incl. anchors, lchains, CIGAR, ds offsets) against the reference library."""
    pre, reads = os.path.join(workdir, "svb"), os.path.join(workdir, "svb.reads.fa")
    T.sim_graph(pre, 400000, 4, 31)
    T.sim_reads(pre + ".hap.fa", reads, n_reads, 12000, "ont", seed)
    names, seqs = T.read_fasta(reads)
    want, mo_r = T.map_with_ref(pre + ".gfa", names, seqs, "lr")
    got, mo_e, st = T.map_with_engine(lib, pre + ".gfa", names, seqs, "lr")
    assert (mo_r.occ_max1, mo_r.lc_max_occ) == (mo_e.occ_max1, mo_e.lc_max_occ)
    for i, (a, b) in enumerate(zip(want, got)):
        d = T.diff_results(a, b)
        assert d is None, "read %d (%s): %s" % (i, names[i], d)
    return st
