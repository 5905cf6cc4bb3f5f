#!/usr/bin/env python3
"""Randomised parity battery on the CPU simulators of the device code (no GPU): random SV graphs (mgsim), random read sets, lr / asm
presets; the GAF text of every read against the unmodified reference binary's.

    python tools/random_parity.py {1|32} SEED N_CASES      # one-lane or 32-lane (fibre) simulator

TEST INFRASTRUCTURE: executes oracle/_ref/minigraph as the checker.  A differing case leaves <case>.got.gaf / .want.gaf in /tmp."""
import os
import random
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import mgtest as T  # noqa: E402


def main():
    lib = T.load_hostsim32() if sys.argv[1] == "32" else T.load_hostsim()
    seed0, n_cases = int(sys.argv[2]), int(sys.argv[3])
    wd = "/tmp/mgb_random_parity_%d" % seed0
    os.makedirs(wd, exist_ok=True)
    n_ok = 0
    for k in range(n_cases):
        seed = seed0 * 100 + k
        rng = random.Random(seed)
        glen, nh, rl = rng.choice([300000, 1000000, 3000000]), rng.choice([2, 4, 8]), rng.choice([3000, 9000, 15000, 30000])
        preset = rng.choice(["lr", "lr", "asm"])
        err = "hifi" if preset == "asm" else rng.choice(["ont", "hifi"])
        pre = os.path.join(wd, "g%d" % k)
        T.sim_graph(pre, glen, nh, seed)
        fa = pre + ".reads.fa"
        nr = max(20, min(300, 2000000 // rl))
        T.sim_reads(pre + ".hap.fa", fa, nr, rl, err, seed + 1)
        names, seqs = T.read_fasta(fa)
        t0 = time.time()
        got, _ = T.gaf_with_engine(lib, pre + ".gfa", names, seqs, preset)
        want = T.gaf_with_ref_binary(pre + ".gfa", fa, preset, threads=16)
        ok = got == want
        print("seed %d: graph %d bp x %d haplotypes, %d reads x %d bp (%s), -cx %s: %s  (%.0f s)" % (
            seed, glen, nh, nr, rl, err, preset, "identical" if ok else "DIFFERENT", time.time() - t0), flush=True)
        if not ok:
            open(pre + ".got.gaf", "wb").write(got)
            open(pre + ".want.gaf", "wb").write(want)
        n_ok += ok
    print("%d of %d cases identical" % (n_ok, n_cases))
    return 0 if n_ok == n_cases else 1


if __name__ == "__main__":
    sys.exit(main())
