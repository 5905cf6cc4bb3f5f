"""Mapping presets, restating options.c:5-108 of the reference (mg_idxopt_init, mg_mapopt_init, mg_opt_set).

Only the presets of the mapping path are provided (lr = default, asm); the graph-generation options are out of
scope. Values are kept in the C structs of the ABI so that they can be handed to mg_index()/mg_map_batch()."""
from .capi import mg_idxopt_t, mg_mapopt_t, MG_M_RMQ, MG_M_CIGAR

MG_M_SR, MG_M_FRAG_MODE, MG_M_FRAG_MERGE, MG_M_HEAP_SORT, MG_M_2_IO_THREADS = 0x20, 0x40, 0x80, 0x400, 0x80000  # minigraph.h:10-24


def idxopt_init():
    io = mg_idxopt_t()
    io.k, io.w, io.bucket_bits = 17, 11, 14  # options.c:8-10
    return io


def mapopt_init():
    mo = mg_mapopt_t()  # zero-initialised like memset (options.c:15)
    mo.seed = 11
    mo.occ_max1, mo.occ_max1_cap = 50, 250
    mo.occ_max1_frac = 2e-4
    mo.max_gap, mo.max_gap_ref, mo.max_gap_pre = 5000, -1, 1000
    mo.max_lc_skip = mo.max_gc_skip = 25
    mo.max_lc_iter = 5000
    mo.bw, mo.bw_long = 500, 20000
    mo.rmq_size_cap, mo.rmq_rescue_size, mo.rmq_rescue_ratio = 100000, 1000, 0.1
    mo.mini_batch_size = 500000000
    mo.div = 0.1
    mo.chn_pen_gap, mo.chn_pen_skip = 1.0, 0.05
    mo.min_lc_cnt, mo.min_lc_score = 5, 40
    mo.min_gc_cnt, mo.min_gc_score = 5, 50
    mo.gdp_max_ed = 10000
    mo.lc_max_trim, mo.lc_max_occ = 50, 2
    mo.mask_level, mo.sub_diff, mo.best_n, mo.pri_ratio = 0.5, 6, 5, 0.8
    mo.ref_bonus, mo.pe_ori = 0, 0
    mo.min_cov_mapq, mo.min_cov_blen = 20, 1000
    mo.cap_kalloc = 1000000000
    return mo


def opt_set(preset=None, cigar=True):
    """Return (idxopt, mapopt) for `-x preset` (options.c:65-108); `cigar` adds -c (main.c:148)."""
    io, mo = idxopt_init(), mapopt_init()
    if preset in (None, "lr"):
        pass
    elif preset == "asm":
        io.k, io.w = 19, 10
        mo.flag |= MG_M_RMQ
        mo.occ_max1, mo.occ_max1_cap = 10, 100
        mo.bw, mo.bw_long = 1000, 150000
        mo.max_gap, mo.max_gap_pre = 10000, 1000
        mo.min_lc_cnt, mo.min_lc_score = 5, 40
        mo.min_gc_cnt, mo.min_gc_score = 5, 1000
        mo.min_cov_mapq, mo.min_cov_blen = 5, 100000
        mo.max_lc_skip = mo.max_gc_skip = 50
        mo.div = 0.01
        mo.mini_batch_size = 4000000000
    elif preset in ("se", "sr"):  # options.c:87-104
        io.k, io.w = 21, 10
        mo.flag |= MG_M_SR | MG_M_HEAP_SORT | MG_M_2_IO_THREADS
        mo.occ_max1, mo.occ_max1_cap = 1000, 2500
        mo.max_gap = 100
        mo.bw = mo.bw_long = 100
        mo.max_frag_len = 800
        mo.pri_ratio = 0.5
        mo.min_lc_cnt, mo.min_lc_score = 2, 25
        mo.min_gc_cnt, mo.min_gc_score = 3, 40
        mo.mini_batch_size = 50000000
        mo.min_cov_blen = 50
        mo.chn_pen_gap = 0.2
        mo.ref_bonus = 1
        if preset == "sr":
            mo.flag |= MG_M_FRAG_MODE | MG_M_FRAG_MERGE
            mo.pe_ori = 0 << 1 | 1
    else:
        raise ValueError("unsupported preset %r (mapping presets: lr, asm, sr, se)" % (preset,))
    if cigar:
        mo.flag |= MG_M_CIGAR
    return io, mo
