// mgb_galign.cuh -- base-level alignment of graph chains and the per-read result blob.
//   gchain_cigar()  (reference: galign.c:39-145 mg_gchain_cigar)
//   gchain_ds()     (reference: galign.c:182-293 mg_gchain_gen_ds)
//   stage_align()   K6-K8 for one read (reference: map-algo.c:455-479)
#pragma once
#include "mgb_pipeline.cuh"
#include "mgb_gchain.cuh"
#include "mgb_wfa.cuh"

namespace mgb {

MG_HD inline int cigar_append1(Arena &A, AVec<uint64_t> &c, int32_t op, int32_t len)
{
	if (c.n > 0 && (int32_t)(c.a[c.n - 1] & 0xf) == op) c.a[c.n - 1] += (uint64_t)(int64_t)len << 4;
	else {
		uint64_t x = (uint64_t)(int64_t)len << 4 | (uint64_t)op;
		MGB_TRY(avec_push(A, c, x));
	}
	return 0;
}

MG_HD inline int cigar_append(Arena &A, AVec<uint64_t> &c, int32_t n_cigar, const uint32_t *cigar)
{
	if (n_cigar == 0) return 0;
	MGB_TRY(cigar_append1(A, c, (int32_t)(cigar[0] & 0xf), (int32_t)(cigar[0] >> 4)));
	MGB_TRY(avec_reserve(A, c, c.n + n_cigar - 1));
	for (int32_t k = 0; k < n_cigar - 1; ++k) c.a[c.n + k] = cigar[1 + k];
	c.n += n_cigar - 1;
	return 0;
}

struct CigarOut { uint64_t *cigar; int32_t n; };

// One CIGAR per graph chain: stitch the target across the walk between consecutive kept anchors and align each gap.
// out[i].cigar is allocated in the arena (kept until the caller releases its mark).
MG_HD inline int gchain_cigar(Arena &A, const GraphDev &g, const char *qseq, GcSet &gt, CigarOut *out)
{
	for (int32_t i = 0; i < gt.n_gc; ++i) {
		GChain *gc = &gt.gc[i];
		int32_t l0 = gc->off;
		const int32_t off_a0 = gt.lc[l0].off;
		int32_t j, j0 = 0, k, l, l_seq;
		AVec<uint64_t> cigar;
		avec_init(cigar);
		MGB_TRY(avec_reserve(A, cigar, 64));
		MGB_TRY(cigar_append1(A, cigar, 7, (int32_t)(gt.a[off_a0].y >> 32 & 0xff)));
		for (j = 1; j < gc->n_anchor; ++j) {
			const u128 *q, *p = &gt.a[off_a0 + j];
			if ((p->y & SEED_IGNORE) && j != gc->n_anchor - 1) continue;
			q = &gt.a[off_a0 + j0];
			for (l = l0; l < gc->off + gc->cnt; ++l) {
				const LLChain *r = &gt.lc[l];
				if (off_a0 + j >= r->off && off_a0 + j < r->off + r->cnt) break;
			}
			if (l >= gc->off + gc->cnt) return MGB_E_INTERNAL;
			uint64_t mark = A.top;
			const char *tseq;
			if (l == l0) { // same vertex: the target is a slice of the stored sequence
				l_seq = (int32_t)p->x - (int32_t)q->x;
				tseq = g_vseq(g, gt.lc[l0].v) + ((int32_t)q->x + 1);
			} else {
				uint32_t v = gt.lc[l0].v;
				int32_t tot = g.seg_len[v >> 1] - (int32_t)q->x - 1;
				for (k = l0 + 1; k < l; ++k) tot += g_vlen(g, gt.lc[k].v);
				tot += (int32_t)p->x + 1;
				char *seq;
				MGB_ALLOC(A, seq, char, tot + 1);
				l_seq = g.seg_len[v >> 1] - (int32_t)q->x - 1;
				{
					const char *s = g_vseq(g, v) + ((int32_t)q->x + 1);
					for (int32_t x = 0; x < l_seq; ++x) seq[x] = s[x];
				}
				for (k = l0 + 1; k < l; ++k) {
					v = gt.lc[k].v;
					const char *s = g_vseq(g, v);
					int32_t vl = g_vlen(g, v);
					for (int32_t x = 0; x < vl; ++x) seq[l_seq + x] = s[x];
					l_seq += vl;
				}
				{
					const char *s = g_vseq(g, gt.lc[l].v);
					int32_t n = (int32_t)p->x + 1;
					for (int32_t x = 0; x < n; ++x) seq[l_seq + x] = s[x];
					l_seq += n;
				}
				tseq = seq;
			}
			{
				int32_t qlen = (int32_t)p->y - (int32_t)q->y;
				const char *qs = &qseq[(int32_t)q->y + 1];
				if (!(l_seq > 0 || qlen > 0)) return MGB_E_INTERNAL;
				if (l_seq == 0) { A.top = mark; MGB_TRY(cigar_append1(A, cigar, 1, qlen)); }
				else if (qlen == 0) { A.top = mark; MGB_TRY(cigar_append1(A, cigar, 2, l_seq)); }
				else if (l_seq == qlen && (uint64_t)(int64_t)qlen <= (q->y >> 32 & 0xff)) { A.top = mark; MGB_TRY(cigar_append1(A, cigar, 7, qlen)); }
				else {
					WfResult rst;
					MGB_TRY(wfa_exact(A, l_seq, tseq, qlen, qs, 100000000LL, &rst));
					if (rst.s < 0) return MGB_E_UNSUPPORTED; // TODO(round 2): chaining heuristic of the reference (miniwfa.c:776-834)
					// the gap CIGAR sits above `cigar`; when the vector has to grow it moves above the gap CIGAR, and
					// the hole is reclaimed with the read.  Copy the ops first if growth is impossible in place.
					if (cigar.n + rst.n_cigar > cigar.m) {
						// release the gap scratch by moving the ops to a temporary that survives the regrowth
						MGB_TRY(cigar_append(A, cigar, rst.n_cigar, rst.cigar));
					} else {
						MGB_TRY(cigar_append(A, cigar, rst.n_cigar, rst.cigar));
						A.top = mark;
					}
				}
			}
			j0 = j, l0 = l;
		}
		out[i].cigar = cigar.a, out[i].n = (int32_t)cigar.n;
		gc->has_cigar = 1;
		gc->n_cigar = (int32_t)cigar.n;
		gc->c_ss = (int32_t)gt.a[off_a0].x + 1 - (int32_t)(gt.a[off_a0].y >> 32 & 0xff);
		gc->c_ee = (int32_t)gt.a[off_a0 + gc->n_anchor - 1].x + 1;
		gc->c_mlen = gc->c_blen = gc->c_aplen = 0;
		for (j = 0, l = 0; j < gc->n_cigar; ++j) {
			int32_t op = (int32_t)(cigar.a[j] & 0xf), len = (int32_t)(cigar.a[j] >> 4);
			if (op == 7) gc->c_mlen += len, gc->c_blen += len;
			else gc->c_blen += len;
			if (op != 1) gc->c_aplen += len;
			if (op != 2) l += len;
		}
		if (!(l == gc->qe - gc->qs && gc->c_aplen == gc->pe - gc->ps)) return MGB_E_INTERNAL;
	}
	return 0;
}

// ---- ds:Z difference string ----

struct DsOut { char *ds; int32_t len; int32_t *off; int32_t n_off; };

MG_HD inline char ds_nt(const char *s, int64_t i) { return "acgtn"[nt4((uint8_t)s[i])]; }

MG_HD inline int ds_putc(Arena &A, AVec<char> &s, char c) { return avec_push(A, s, c); }

MG_HD inline int ds_putint(Arena &A, AVec<char> &s, int32_t c)
{
	char buf[16];
	int l = 0;
	uint32_t x = c >= 0? (uint32_t)c : (uint32_t)(-c);
	do { buf[l++] = (char)(x % 10 + '0'); x /= 10; } while (x > 0);
	if (c < 0) buf[l++] = '-';
	for (int i = l - 1; i >= 0; --i) MGB_TRY(avec_push(A, s, buf[i]));
	return 0;
}

// reference: galign.c:153-180 write_indel
MG_HD inline int ds_write_indel(Arena &A, AVec<char> &str, int64_t len, const char *seq, int64_t ll, int64_t lr)
{
	int64_t i;
	if (ll + lr >= len) {
		MGB_TRY(ds_putc(A, str, '['));
		for (i = 0; i < len; ++i) MGB_TRY(ds_putc(A, str, ds_nt(seq, i)));
		MGB_TRY(ds_putc(A, str, ']'));
	} else {
		int64_t k = 0;
		if (ll > 0) {
			MGB_TRY(ds_putc(A, str, '['));
			for (i = 0; i < ll; ++i) MGB_TRY(ds_putc(A, str, ds_nt(seq, k + i)));
			MGB_TRY(ds_putc(A, str, ']'));
			k += ll;
		}
		for (i = 0; i < len - lr - ll; ++i) MGB_TRY(ds_putc(A, str, ds_nt(seq, k + i)));
		k += len - lr - ll;
		if (lr > 0) {
			MGB_TRY(ds_putc(A, str, '['));
			for (i = 0; i < lr; ++i) MGB_TRY(ds_putc(A, str, ds_nt(seq, k + i)));
			MGB_TRY(ds_putc(A, str, ']'));
		}
	}
	return 0;
}

MG_HD inline int gchain_ds(Arena &A, const GraphDev &g, const char *qseq, GcSet &gt, const CigarOut *cg, DsOut *out)
{
	for (int32_t i = 0; i < gt.n_gc; ++i) {
		GChain *gc = &gt.gc[i];
		int32_t j;
		int64_t x, y;
		AVec<char> str;
		AVec<int32_t> off;
		char *seq;
		int64_t seq_l = 0;
		avec_init(str), avec_init(off);
		MGB_ALLOC(A, seq, char, gc->c_aplen + 1);
		for (j = 0; j < gc->cnt; ++j) { // the aligned part of the walk
			int32_t k = gc->off + j;
			uint32_t v = gt.lc[k].v;
			int32_t slen = g_vlen(g, v);
			int32_t st = j > 0? 0 : gc->c_ss;
			int32_t en = j < gc->cnt - 1? slen : gc->c_ee;
			if (seq_l + (en - st) > gc->c_aplen) return MGB_E_INTERNAL;
			const char *s = g_vseq(g, v) + st;
			for (int32_t t = 0; t < en - st; ++t) seq[seq_l + t] = s[t];
			seq_l += en - st;
		}
		if (seq_l != gc->c_aplen) return MGB_E_INTERNAL;
		// both vectors grow; interleaved growth wastes arena but stays correct
		MGB_TRY(avec_reserve(A, off, 64));
		MGB_TRY(avec_reserve(A, str, 256));
		for (j = 0, x = 0, y = gc->qs; j < gc->n_cigar; ++j) {
			int64_t op = (int64_t)(cg[i].cigar[j] & 0xf), len = (int64_t)(cg[i].cigar[j] >> 4);
			if (op == 0 || op == 7 || op == 8) {
				int64_t z;
				int32_t l = 0;
				for (z = 0; z < len; ++z) {
					uint8_t cx = (uint8_t)nt4((uint8_t)seq[x + z]);
					uint8_t cy = (uint8_t)nt4((uint8_t)qseq[y + z]);
					if (cx != cy) {
						if (l > 0) {
							MGB_TRY(avec_push(A, off, (int32_t)str.n));
							MGB_TRY(ds_putc(A, str, ':'));
							MGB_TRY(ds_putint(A, str, l));
						}
						MGB_TRY(avec_push(A, off, (int32_t)str.n));
						MGB_TRY(ds_putc(A, str, '*'));
						MGB_TRY(ds_putc(A, str, "acgtn"[cx]));
						MGB_TRY(ds_putc(A, str, "acgtn"[cy]));
						l = 0;
					} else ++l;
				}
				if (l > 0) {
					MGB_TRY(avec_push(A, off, (int32_t)str.n));
					MGB_TRY(ds_putc(A, str, ':'));
					MGB_TRY(ds_putint(A, str, l));
				}
				x += len, y += len;
			} else if (op == 1) {
				int64_t z, ll, lr;
				for (z = 1; z <= len; ++z)
					if (y - z < gc->qs || qseq[y + len - z] != qseq[y - z]) break;
				lr = z - 1;
				for (z = 0; z < len; ++z)
					if (y + len + z >= gc->qe || qseq[y + len + z] != qseq[y + z]) break;
				ll = z;
				MGB_TRY(avec_push(A, off, (int32_t)str.n));
				MGB_TRY(ds_putc(A, str, '+'));
				MGB_TRY(ds_write_indel(A, str, len, &qseq[y], ll, lr));
				y += len;
			} else if (op == 2) {
				int64_t z, ll, lr;
				for (z = 1; z <= len; ++z)
					if (x - z < 0 || seq[x + len - z] != seq[x - z]) break;
				lr = z - 1;
				for (z = 0; z < len; ++z)
					if (x + len + z >= gc->c_aplen || seq[x + z] != seq[x + len + z]) break;
				ll = z;
				MGB_TRY(avec_push(A, off, (int32_t)str.n));
				MGB_TRY(ds_putc(A, str, '-'));
				MGB_TRY(ds_write_indel(A, str, len, &seq[x], ll, lr));
				x += len;
			}
		}
		out[i].ds = str.a, out[i].len = (int32_t)str.n, out[i].off = off.a, out[i].n_off = (int32_t)off.n;
		gc->ds_len = (int32_t)str.n, gc->n_dsoff = (int32_t)off.n;
	}
	return 0;
}

// per-read result header, one per read, in an array parallel to ReadMeta
struct ReadOut {
	int32_t status;
	int32_t n_gc, n_lc, n_a, rep_len;
	int32_t n_mz;
	uint32_t blob_size;
	int64_t blob_off;   // byte offset into the output pool; blob = GChain[n_gc] | LLChain[n_lc] | u128 a[n_a] | cigars | ds | ds offsets
};

MG_HD inline uint64_t align8(uint64_t x) { return (x + 7) & ~(uint64_t)7; }

// K6-K8 for one read.
MG_HD inline int stage_align(const PipeCtx &c, ReadOut *routs, int rid, Arena &A)
{
	ReadMeta &m = c.meta[rid];
	ReadOut &ro = routs[rid];
	const MapOptDev &o = c.opt;
	ro.status = 0, ro.n_gc = ro.n_lc = ro.n_a = 0, ro.rep_len = m.rep_len, ro.n_mz = m.n_mz, ro.blob_size = 0, ro.blob_off = 0;
	if (m.status != 0) { ro.status = m.status; return 0; } // status 1: read skipped (empty or too long) -> no result object
	uint64_t mark = A.top;
	const char *qseq = c.b.seq + c.b.seq_off[rid];
	const int32_t qlen = c.b.seq_len[rid];
	const u128 *a = c.anchor + m.a_off;
	int32_t n_lc = m.n_lc, n_u = 0;
	uint64_t *u = 0;
	LChain *lc;
	MGB_ALLOC(A, lc, LChain, n_lc);
	for (int32_t i = 0; i < n_lc; ++i) lc[i] = c.lchain[m.lc_off + i];
	MGB_TRY(gchain1_dp(A, c.g, &n_lc, lc, qlen, o.bw_long, o.bw_long, o.bw_long, o.max_gc_skip, o.ref_bonus, o.chn_pen_gap, o.mask_level, a, &u, &n_u));
	GcSet gs;
	MGB_TRY(gchain_gen(A, c.g, n_u, u, lc, a, m.hash, o.min_gc_cnt, o.min_gc_score, o.gdp_max_ed, 1, qseq, gs));
	gs.rep_len = m.rep_len;
	MGB_TRY(gchain_set_parent(A, o.mask_level, gs.n_gc, gs.gc, o.sub_diff));
	gchain_flt_sub(o.pri_ratio, c.ix.k * 2, o.best_n, gs.n_gc, gs.gc);
	MGB_TRY(gchain_drop_flt(A, gs));
	MGB_TRY(gchain_set_mapq(o, gs, qlen, m.n_mz, o.min_gc_score));
	CigarOut *cg = 0;
	DsOut *ds = 0;
	if ((o.flag & F_CIGAR) && gs.n_gc > 0) {
		MGB_ALLOC(A, cg, CigarOut, gs.n_gc);
		MGB_ALLOC(A, ds, DsOut, gs.n_gc);
		MGB_TRY(gchain_cigar(A, c.g, qseq, gs, cg));
		MGB_TRY(gchain_ds(A, c.g, qseq, gs, cg, ds));
	}
	// ---- serialise ----
	uint64_t sz = 0;
	uint64_t off_gc = 0; sz += align8((uint64_t)gs.n_gc * sizeof(GChain));
	uint64_t off_lc = sz; sz += align8((uint64_t)gs.n_lc * sizeof(LLChain));
	uint64_t off_a = sz; sz += align8((uint64_t)gs.n_a * sizeof(u128));
	for (int32_t i = 0; i < gs.n_gc; ++i) {
		GChain *gc = &gs.gc[i];
		if (cg) {
			gc->cigar_off = (int64_t)sz; sz += align8((uint64_t)cg[i].n * 8);
			gc->ds_off = (int64_t)sz; sz += align8((uint64_t)ds[i].len + 1);
			gc->dsoff_off = (int64_t)sz; sz += align8((uint64_t)ds[i].n_off * 4);
		} else gc->has_cigar = 0, gc->n_cigar = 0, gc->cigar_off = gc->ds_off = gc->dsoff_off = 0, gc->ds_len = gc->n_dsoff = 0;
	}
	int64_t boff = pool_alloc(c.pool_out, sz);
	if (boff < 0) return MGB_E_POOL;
	char *blob = c.out + boff;
	{
		GChain *d = (GChain*)(blob + off_gc);
		for (int32_t i = 0; i < gs.n_gc; ++i) d[i] = gs.gc[i];
		LLChain *dl = (LLChain*)(blob + off_lc);
		for (int32_t i = 0; i < gs.n_lc; ++i) dl[i] = gs.lc[i];
		u128 *da = (u128*)(blob + off_a);
		for (int32_t i = 0; i < gs.n_a; ++i) da[i] = gs.a[i];
		if (cg) {
			for (int32_t i = 0; i < gs.n_gc; ++i) {
				uint64_t *dc = (uint64_t*)(blob + gs.gc[i].cigar_off);
				for (int32_t k = 0; k < cg[i].n; ++k) dc[k] = cg[i].cigar[k];
				char *dd = blob + gs.gc[i].ds_off;
				for (int32_t k = 0; k < ds[i].len; ++k) dd[k] = ds[i].ds[k];
				dd[ds[i].len] = 0;
				int32_t *dof = (int32_t*)(blob + gs.gc[i].dsoff_off);
				for (int32_t k = 0; k < ds[i].n_off; ++k) dof[k] = ds[i].off[k];
			}
		}
	}
	ro.n_gc = gs.n_gc, ro.n_lc = gs.n_lc, ro.n_a = gs.n_a, ro.blob_size = (uint32_t)sz, ro.blob_off = boff;
	A.top = mark;
	return 0;
}

} // namespace mgb
