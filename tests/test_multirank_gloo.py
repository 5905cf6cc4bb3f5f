"""world_size-2 test of the sharding + offset all-gather on gloo (CPU): each rank maps its block of reads with the
single-lane simulator, writes its GAF at the gathered offset into a shared file; the file must equal the golden GAF."""
import os
import sys

import pytest
import torch.multiprocessing as mp

import cases
import mgtest as T


def _rank_main(rank, world, port, workdir, out_path):
    import torch.distributed as dist
    sys.path.insert(0, T.REPO)
    from minigraph_b200 import dist as mdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = T.load_hostsim()
    reads = os.path.join(workdir, "mt.reads.fa")
    names, seqs = T.read_fasta(reads)
    b = mdist.shard_bounds([len(s) for s in seqs], world)
    lo, hi = b[rank], b[rank + 1]
    text, _ = T.gaf_with_engine(lib, os.path.join(T.FIX, "MT.gfa"), names[lo:hi], seqs[lo:hi], "lr")
    off, counts = mdist.gaf_offsets(len(text))
    assert sum(counts[:rank]) == off
    dist.barrier()
    with open(out_path, "r+b") as f:
        f.seek(off)
        f.write(text)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_single_process_output(workdir):
    hap, reads = os.path.join(workdir, "mt.hap.fa"), os.path.join(workdir, "mt.reads.fa")
    T.sim_mt_haps(hap)
    T.sim_reads(hap, reads, 24, 10000, "ont", 11)
    T.load_hostsim()  # build once before forking
    out = os.path.join(workdir, "two_ranks.gaf")
    want = cases.golden("c2_MT_24x10k_ont_s11.lr.gaf")
    with open(out, "wb") as f:
        f.write(b"\0" * len(want))
    mp.spawn(_rank_main, args=(2, 29731, workdir, out), nprocs=2, join=True)
    with open(out, "rb") as f:
        got = f.read()
    assert got == want, cases.first_diff(got, want)


def test_shard_bounds_cover_everything():
    from minigraph_b200 import dist as mdist
    for world in (1, 2, 3, 8):
        ls = [1000, 10, 5000, 5000, 20, 300, 7000, 1, 1, 9000]
        b = mdist.shard_bounds(ls, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == len(ls) and all(x <= y for x, y in zip(b, b[1:]))
    assert mdist.shard_bounds([], 4) == [0, 0, 0, 0, 0]
