#!/usr/bin/env python
"""Function-level golden vectors from the UNMODIFIED reference (oracle/_ref/libmgref.so, built from /root/reference):
mg_sketch() on fixture slices (incl. the N of MT-human and a homopolymer/tandem torture string) and radix_sort_128x()
on arrays full of ties.  Output: tests/golden/vectors.json (small).  The reference ships no vectors."""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import mgtest as T  # noqa: E402


class mg128_v(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(T.capi.mg128_t))]


def ref_sketch(ref, seq, w, k, rid):
    v = mg128_v()
    ref.mg_sketch(None, seq, len(seq), w, k, rid, C.byref(v))
    out = [[v.a[i].x, v.a[i].y] for i in range(v.n)]
    C.CDLL(None).free(v.a)
    return out


def splitmix(seed):
    s = seed
    while True:
        s = (s + 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        yield z ^ (z >> 31)


def sort_inputs():
    g = splitmix(99)
    cases = []
    for n, nkeys in ((5, 2), (64, 3), (65, 3), (300, 5), (1000, 40), (3000, 7), (2000, 1 << 40)):
        xs = [next(g) % nkeys if nkeys < (1 << 30) else (next(g) >> 8) for _ in range(n)]
        if nkeys == 40:
            xs = [x << 33 | (next(g) & 1) << 32 | (next(g) % 3) for x in xs]  # seed-like keys with ties
        cases.append([[x, i] for i, x in enumerate(xs)])
    return cases


def main():
    ref = T.load_ref()
    ref.mg_sketch.restype = None
    ref.mg_sketch.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.POINTER(mg128_v)]
    ref.radix_sort_128x.restype = None
    ref.radix_sort_128x.argtypes = [C.POINTER(T.capi.mg128_t), C.POINTER(T.capi.mg128_t)]
    _, hs = T.read_fasta(os.path.join(T.FIX, "MT-human.fa"))
    h = hs[0]
    seqs = {
        "human_0_1500": h[:1500], "human_N_region": h[2900:3400], "tandem": b"ACACACACACACACACACACACACACACACACACACACACACACACACAC" * 4 + h[100:300],
        "homopolymer": b"A" * 120 + h[500:620] + b"T" * 80, "short": h[10:40], "with_Ns": b"ACGTNNNN" + h[700:900] + b"N" + h[900:1000],
    }
    out = {"sketch": [], "sort": []}
    for name, s in seqs.items():
        for (w, k) in ((11, 17), (10, 19), (5, 15)):
            out["sketch"].append({"name": name, "seq": s.decode(), "w": w, "k": k, "rid": 3, "mz": ref_sketch(ref, s, w, k, 3)})
    for arr in sort_inputs():
        a = (T.capi.mg128_t * len(arr))()
        for i, (x, y) in enumerate(arr):
            a[i].x, a[i].y = x, y
        ref.radix_sort_128x(a, C.cast(C.byref(a, C.sizeof(a)), C.POINTER(T.capi.mg128_t)))
        out["sort"].append({"in": arr, "out_y": [a[i].y for i in range(len(arr))]})
    with open(os.path.join(HERE, "vectors.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("sketch cases", len(out["sketch"]), "sort cases", len(out["sort"]))


if __name__ == "__main__":
    main()
