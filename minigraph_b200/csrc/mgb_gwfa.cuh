// mgb_gwfa.cuh -- graph wavefront edit distance with walk traceback, used to bridge two linear chains.
// (reference: gfa-ed.c:117-593 gwf_ed_extend / gfa_ed_step).  The *walk* that is reported depends on the order in
// which equal-cost diagonals are generated, merged and pruned, so the control flow below replays the
// reference step by step: batch extension of consecutive diagonals, the FIFO of boundary diagonals, first-visitor
// wins at vertex ends, the out-of-order split sort, dedup keeping the first maximum and the periodic prune.
// Containers are arena vectors; the two hash tables only need set/map semantics.
#pragma once
#include "mgb_model.cuh"

namespace mgb {

static const int32_t GWF_DIAG_SHIFT = 0x40000000;

struct GwfDiag { // one diagonal of the wavefront (reference: gfa-ed.c:117-123)
	uint64_t vd;   // vertex<<32 | diagonal + GWF_DIAG_SHIFT
	int32_t k, len;
	uint32_t xo;   // anti-diagonal<<1 | out-of-order flag
	int32_t t;     // traceback node
};
struct GwfIntv { uint64_t vd0, vd1; };
struct GwfTrace { int32_t v, pre; };

struct KeyDiagVd { MG_HD uint64_t operator()(const GwfDiag &p) const { return p.vd; } };
struct KeyIntvVd0 { MG_HD uint64_t operator()(const GwfIntv &p) const { return p.vd0; } };

MG_HD inline uint64_t gwf_gen_vd(uint32_t v, int32_t d) { return (uint64_t)v << 32 | (uint64_t)(uint32_t)(GWF_DIAG_SHIFT + d); }

// open-addressing u64 -> int32 table with epoch-based clearing
struct U64Tab {
	uint64_t *key;
	int32_t *val, *ep;
	int32_t bits, epoch;
	int64_t n;
};

MG_HD inline uint32_t u64tab_hash(uint64_t key)
{
	key = ~key + (key << 21);
	key = key ^ key >> 24;
	key = (key + (key << 3)) + (key << 8);
	key = key ^ key >> 14;
	key = (key + (key << 2)) + (key << 4);
	key = key ^ key >> 28;
	key = key + (key << 31);
	return (uint32_t)key;
}

MG_HD MG_NOINLINE inline int u64tab_alloc(Arena &A, U64Tab &t, int bits)
{
	MGB_ALLOC(A, t.key, uint64_t, 1LL << bits);
	MGB_ALLOC(A, t.val, int32_t, 1LL << bits);
	MGB_ALLOC(A, t.ep, int32_t, 1LL << bits);
	MGB_NO_UNROLL
	for (int64_t i = 0; i < (1LL << bits); ++i) t.ep[i] = 0;
	t.bits = bits;
	return 0;
}
MG_HD inline int u64tab_init(Arena &A, U64Tab &t, int bits) { t.epoch = 1, t.n = 0; return u64tab_alloc(A, t, bits); }
MG_HD inline void u64tab_clear(U64Tab &t) { ++t.epoch, t.n = 0; }

// returns slot index; *absent tells whether the key was inserted by this call
// twice the slots, entries re-inserted in slot order (out of line: once per doubling)
MG_HD MG_NOINLINE inline int u64tab_grow(Arena &A, U64Tab &t)
{
	U64Tab o = t;
	MGB_TRY(u64tab_alloc(A, t, o.bits + 1));
	t.epoch = 1;
	uint64_t nmask = (1ULL << t.bits) - 1;
	MGB_NO_UNROLL
	for (int64_t i = 0; i < (1LL << o.bits); ++i) {
		if (o.ep[i] != o.epoch) continue;
		uint64_t g = u64tab_hash(o.key[i]) & nmask;
		while (t.ep[g] == t.epoch) g = (g + 1) & nmask;
		t.key[g] = o.key[i], t.val[g] = o.val[i], t.ep[g] = t.epoch;
	}
	return 0;
}

MG_HD inline int u64tab_put(Arena &A, U64Tab &t, uint64_t key, int *absent, int64_t *slot)
{
	for (;;) {
		uint64_t mask = (1ULL << t.bits) - 1, h = u64tab_hash(key) & mask;
		while (t.ep[h] == t.epoch && t.key[h] != key) h = (h + 1) & mask;
		if (t.ep[h] == t.epoch) { *absent = 0, *slot = (int64_t)h; return 0; }
		if ((uint64_t)(t.n + 1) * 2 > (1ULL << t.bits)) { // grow
			MGB_TRY(u64tab_grow(A, t));
			continue;
		}
		t.key[h] = key, t.val[h] = 0, t.ep[h] = t.epoch, ++t.n;
		*absent = 1, *slot = (int64_t)h;
		return 0;
	}
}

struct GwfOpt { int32_t traceback, bw_dyn, max_lag, max_chk, s_term; int64_t i_term; };

struct GwfResult {
	int32_t s;
	uint32_t end_v;
	int32_t end_off, wlen, nv;
	int64_t n_iter;
	int32_t *v;
};

struct GwfState {
	const GraphDev *g;
	int32_t ql;
	const char *q;
	AVec<GwfDiag> a, B, ooo, Q;
	int64_t q_head;
	AVec<GwfIntv> intv, tmp, swap;
	AVec<GwfTrace> t;
	U64Tab ha, ht;
	int32_t s, end_tb;
};

MG_HD inline int gwf_trace_push(Arena &A, GwfState &z, int32_t v, int32_t pre, int32_t *idx)
{
	uint64_t key = (uint64_t)(uint32_t)v << 32 | (uint32_t)pre;
	int absent; int64_t slot;
	MGB_TRY(u64tab_put(A, z.ht, key, &absent, &slot));
	if (absent) {
		GwfTrace x; x.v = v, x.pre = pre;
		MGB_TRY(avec_push_c(A, z.t, x));
		z.ht.val[slot] = (int32_t)z.t.n - 1;
	}
	*idx = z.ht.val[slot];
	return 0;
}

MG_HD inline int gwf_diag_push(Arena &A, AVec<GwfDiag> &B, uint32_t v, int32_t d, int32_t k, uint32_t x, uint32_t ooo, int32_t t)
{
	GwfDiag p;
	p.vd = gwf_gen_vd(v, d), p.k = k, p.len = 0, p.xo = x << 1 | ooo, p.t = t;
	return avec_push_c(A, B, p);
}

MG_HD inline int32_t gwf_diag_update(GwfDiag *p, uint32_t v, int32_t d, int32_t k, uint32_t x, uint32_t ooo, int32_t t)
{
	uint64_t vd = gwf_gen_vd(v, d);
	if (p->vd == vd) {
		p->xo = p->k > k? p->xo : x << 1 | ooo;
		p->t = p->k > k? p->t : t;
		p->k = p->k > k? p->k : k;
		return 0;
	}
	return 1;
}

// furthest k reachable along diagonal d by exact matches (reference: gfa-ed.c:305-329 gwf_extend1)
MG_HD inline int32_t gwf_extend1(int32_t d, int32_t k, int32_t vl, const char *ts, int32_t ql, const char *qs)
{
	int32_t max_k = (ql - d < vl? ql - d : vl) - 1;
	const char *ts_ = ts + 1, *qs_ = qs + d + 1;
	while (k + 4 <= max_k) { // four bases per step; both buffers carry >= 8 bytes of slack behind their last base
		uint32_t x = ld32_unaligned(ts_ + k) ^ ld32_unaligned(qs_ + k);
		if (x) return k + (ctz32(x) >> 3);
		k += 4;
	}
	while (k < max_k && ts_[k] == qs_[k]) ++k;
	return k;
}

// the two klib sorts of this file, out of line (large inputs only in the warp-wide form below)
MG_HD MG_NOINLINE inline int gwf_sort_diag_cold(Arena &A, GwfDiag *c, int64_t n) { return radix_sort_exact(A, c, n, 8, KeyDiagVd()); }
MG_HD MG_NOINLINE inline int gwf_sort_intv_cold(Arena &A, GwfIntv *a, int64_t n) { return radix_sort_exact(A, a, n, 8, KeyIntvVd0()); }

// split sort exploiting the out-of-order flag (reference: gfa-ed.c:162-187)
MG_HD inline int gwf_diag_sort(Arena &A, GwfState &z, int32_t n_a, GwfDiag *a)
{
	int32_t i, j, k, n_b, n_c = 0;
	MGB_TRY(avec_reserve_c(A, z.ooo, n_a));
	for (i = 0; i < n_a; ++i) if (a[i].xo & 1) ++n_c;
	n_b = n_a - n_c;
	GwfDiag *b = z.ooo.a, *c = b + n_b;
	for (i = j = k = 0; i < n_a; ++i) {
		if (a[i].xo & 1) c[k++] = a[i];
		else b[j++] = a[i];
	}
	MGB_TRY(gwf_sort_diag_cold(A, c, n_c));
	for (k = 0; k < n_c; ++k) c[k].xo &= 0xfffffffeU;
	i = j = k = 0;
	while (i < n_b && j < n_c) {
		if (b[i].vd <= c[j].vd) a[k++] = b[i++];
		else a[k++] = c[j++];
	}
	while (i < n_b) a[k++] = b[i++];
	while (j < n_c) a[k++] = c[j++];
	return 0;
}

// keep one diagonal per (vertex,diag): the first one reaching furthest (reference: gfa-ed.c:190-206)
MG_HD inline int gwf_diag_dedup(Arena &A, GwfState &z, int32_t n_a, GwfDiag *a, int32_t *n_out)
{
	int32_t i, n, st;
	for (i = 1; i < n_a; ++i) if (a[i-1].vd > a[i].vd) break;
	if (i < n_a) MGB_TRY(gwf_diag_sort(A, z, n_a, a));
	for (i = 1, st = 0, n = 0; i <= n_a; ++i) {
		if (i == n_a || a[i].vd != a[st].vd) {
			int32_t max_j = st;
			for (int32_t j = st + 1; j < i; ++j)
				if (a[max_j].k < a[j].k) max_j = j;
			a[n++] = a[max_j];
			st = i;
		}
	}
	*n_out = n;
	return 0;
}

MG_HD inline int64_t gwf_intv_merge_adj(int64_t n, GwfIntv *a)
{
	int64_t i, k;
	uint64_t st, en;
	if (n == 0) return 0;
	st = a[0].vd0, en = a[0].vd1;
	for (i = 1, k = 0; i < n; ++i) {
		if (a[i].vd0 > en) {
			a[k].vd0 = st, a[k++].vd1 = en;
			st = a[i].vd0, en = a[i].vd1;
		} else en = en > a[i].vd1? en : a[i].vd1;
	}
	a[k].vd0 = st, a[k++].vd1 = en;
	return k;
}

// reference: gfa-ed.c:264-278 gwf_dedup
MG_HD inline int gwf_dedup(Arena &A, GwfState &z, int32_t n_a, GwfDiag *a, int32_t *n_out)
{
	if (z.intv.n + z.tmp.n > 0) {
		int64_t i;
		for (i = 1; i < z.tmp.n; ++i) if (z.tmp.a[i-1].vd0 > z.tmp.a[i].vd0) break;
		if (i < z.tmp.n) MGB_TRY(gwf_sort_intv_cold(A, z.tmp.a, z.tmp.n));
		MGB_TRY(avec_reserve_c(A, z.swap, z.intv.n));
		for (i = 0; i < z.intv.n; ++i) z.swap.a[i] = z.intv.a[i];
		z.swap.n = z.intv.n;
		MGB_TRY(avec_reserve_c(A, z.intv, z.intv.n + z.tmp.n));
		{ // merge two sorted lists, then fuse overlapping intervals
			int64_t x = 0, y = 0, k = 0;
			const GwfIntv *b = z.swap.a, *c = z.tmp.a;
			GwfIntv *o = z.intv.a;
			while (x < z.swap.n && y < z.tmp.n) {
				if (b[x].vd0 <= c[y].vd0) o[k++] = b[x++];
				else o[k++] = c[y++];
			}
			while (x < z.swap.n) o[k++] = b[x++];
			while (y < z.tmp.n) o[k++] = c[y++];
			z.intv.n = gwf_intv_merge_adj(k, o);
		}
	}
	MGB_TRY(gwf_diag_dedup(A, z, n_a, a, &n_a));
	if (z.intv.n > 0) { // drop diagonals inside forbidden bands (reference: gfa-ed.c:209-219)
		int32_t i = 0, k = 0;
		int64_t j = 0;
		const GwfIntv *b = z.intv.a;
		while (i < n_a && j < z.intv.n) {
			if (a[i].vd >= b[j].vd0 && a[i].vd < b[j].vd1) ++i;
			else if (a[i].vd >= b[j].vd1) ++j;
			else a[k++] = a[i++];
		}
		while (i < n_a) a[k++] = a[i++];
		n_a = k;
	}
	*n_out = n_a;
	return 0;
}

// reference: gfa-ed.c:281-302 gwf_prune
MG_HD inline int32_t gwf_prune(int32_t n_a, GwfDiag *a, uint32_t max_lag, int32_t bw_dyn)
{
	int32_t i, j, iq, dq, max_i = 0;
	uint32_t max_x = 0;
	MGB_NO_UNROLL
	for (i = 0; i < n_a; ++i)
		if (a[i].xo >> 1 > max_x) max_x = a[i].xo >> 1, max_i = i;
	const GwfDiag *q = &a[max_i];
	iq = (int32_t)q->vd - GWF_DIAG_SHIFT + q->k;
	dq = (int32_t)(q->xo >> 1) - iq - iq;
	MGB_NO_UNROLL
	for (i = j = 0; i < n_a; ++i) {
		GwfDiag *p = &a[i];
		int32_t ip = (int32_t)p->vd - GWF_DIAG_SHIFT + p->k;
		int32_t dp = (int32_t)(p->xo >> 1) - ip - ip;
		int32_t w = dp > dq? dp - dq : dq - dp;
		if (bw_dyn >= 0 && w > bw_dyn) continue;
		if ((p->xo >> 1) + max_lag < max_x) continue;
		a[j++] = *p;
	}
	return j;
}

// Landau-Vishkin style step on a run of consecutive diagonals of one vertex (reference: gfa-ed.c:332-402)
MG_HD inline int gwf_extend_batch(Arena &A, GwfState &z, int32_t n, GwfDiag *a)
{
	const GraphDev &g = *z.g;
	int32_t j, m;
	uint32_t v = (uint32_t)(a->vd >> 32);
	int32_t vl = g_vlen(g, v);
	const char *ts = g_vseq(g, v);
	for (j = 0; j < n; ++j) {
		int32_t k = gwf_extend1((int32_t)a[j].vd - GWF_DIAG_SHIFT, a[j].k, vl, ts, z.ql, z.q);
		a[j].len = k - a[j].k;
		a[j].xo += (uint32_t)a[j].len << 2;
		a[j].k = k;
	}
	MGB_TRY(avec_reserve_c(A, z.B, z.B.n + n + 2));
	GwfDiag *b = &z.B.a[z.B.n];
	b[0].vd = a[0].vd - 1;
	b[0].xo = a[0].xo + 2;
	b[0].k = a[0].k + 1;
	b[0].t = a[0].t;
	b[0].len = 0;
	b[1].vd = a[0].vd;
	b[1].xo = n == 1 || a[0].k > a[1].k? a[0].xo + 4 : a[1].xo + 2;
	b[1].t = n == 1 || a[0].k > a[1].k? a[0].t : a[1].t;
	b[1].k = (n == 1 || a[0].k > a[1].k? a[0].k : a[1].k) + 1;
	b[1].len = 0;
	for (j = 1; j < n - 1; ++j) {
		uint32_t x = a[j-1].xo + 2;
		int32_t k = a[j-1].k, t = a[j-1].t;
		x = k > a[j].k + 1? x : a[j].xo + 4;
		t = k > a[j].k + 1? t : a[j].t;
		k = k > a[j].k + 1? k : a[j].k + 1;
		x = k > a[j+1].k + 1? x : a[j+1].xo + 2;
		t = k > a[j+1].k + 1? t : a[j+1].t;
		k = k > a[j+1].k + 1? k : a[j+1].k + 1;
		b[j+1].vd = a[j].vd, b[j+1].k = k, b[j+1].xo = x, b[j+1].t = t, b[j+1].len = 0;
	}
	if (n >= 2) {
		b[n].vd = a[n-1].vd;
		b[n].xo = a[n-2].k > a[n-1].k + 1? a[n-2].xo + 2 : a[n-1].xo + 4;
		b[n].t = a[n-2].k > a[n-1].k + 1? a[n-2].t : a[n-1].t;
		b[n].k = a[n-2].k > a[n-1].k + 1? a[n-2].k : a[n-1].k + 1;
		b[n].len = 0;
	}
	b[n+1].vd = a[n-1].vd + 1;
	b[n+1].xo = a[n-1].xo + 2;
	b[n+1].t = a[n-1].t;
	b[n+1].k = a[n-1].k;
	b[n+1].len = 0;
	// diagonals touching the end of the vertex or of the query go to the queue
	for (j = 0; j < n; ++j) {
		GwfDiag *p = &a[j];
		if (p->k == vl - 1 || (int32_t)p->vd - GWF_DIAG_SHIFT + p->k == z.ql - 1) {
			p->xo |= 1;
			MGB_TRY(avec_push_c(A, z.Q, *p));
		}
	}
	for (j = 0, m = 0; j < n + 2; ++j) {
		GwfDiag *p = &b[j];
		int32_t d = (int32_t)p->vd - GWF_DIAG_SHIFT;
		if (d + p->k < z.ql && p->k < vl) {
			b[m++] = *p;
		} else if (p->k == vl) {
			GwfIntv iv;
			iv.vd0 = gwf_gen_vd(v, d), iv.vd1 = iv.vd0 + 1;
			MGB_TRY(avec_push_c(A, z.tmp, iv));
		}
	}
	z.B.n += m;
	return 0;
}

// one edit-distance step: extend + next (reference: gfa-ed.c:405-507 gwf_ed_extend). On return z.a is the new wavefront.
MG_HD inline int gwf_ed_extend(Arena &A, GwfState &z, const GwfOpt &opt, uint32_t v1, int32_t off1, GwfResult *r)
{
	const GraphDev &g = *z.g;
	const int32_t ql = z.ql;
	const char *q = z.q;
	int32_t i, x, n = (int32_t)z.a.n, do_dedup = 1;
	r->end_v = (uint32_t)-1;
	r->end_off = z.end_tb = -1;
	z.tmp.n = 0;
	u64tab_clear(z.ha);
	z.Q.n = 0, z.q_head = 0;
	z.B.n = 0;
	MGB_TRY(avec_reserve_c(A, z.B, (int64_t)n * 2));
	GwfDiag *a = z.a.a;
	for (x = 0, i = 1; i <= n; ++i) {
		if (i == n || a[i].vd != a[i-1].vd + 1) {
			MGB_TRY(gwf_extend_batch(A, z, i - x, &a[x]));
			x = i;
		}
	}
	if (z.Q.n == 0) do_dedup = 0;

	while (z.q_head < z.Q.n) {
		GwfDiag t = z.Q.a[z.q_head++];
		uint32_t v, x0;
		int32_t ooo, d, k, vl;
		ooo = t.xo & 1, v = (uint32_t)(t.vd >> 32);
		d = (int32_t)t.vd - GWF_DIAG_SHIFT;
		k = t.k;
		vl = g_vlen(g, v);
		k = gwf_extend1(d, k, vl, g_vseq(g, v), ql, q);
		i = k + d;
		x0 = (t.xo >> 1) + ((uint32_t)(k - t.k) << 1);

		if (k + 1 < vl && i + 1 < ql) { // wavefront in the middle of the vertex
			int32_t push1 = 1, push2 = 1;
			if (z.B.n >= 2) push1 = gwf_diag_update(&z.B.a[z.B.n - 2], v, d-1, k+1, x0 + 1, ooo, t.t);
			if (z.B.n >= 1) push2 = gwf_diag_update(&z.B.a[z.B.n - 1], v, d,   k+1, x0 + 2, ooo, t.t);
			if (push1)          MGB_TRY(gwf_diag_push(A, z.B, v, d-1, k+1, x0 + 1, 1, t.t));
			if (push2 || push1) MGB_TRY(gwf_diag_push(A, z.B, v, d,   k+1, x0 + 2, 1, t.t));
			MGB_TRY(gwf_diag_push(A, z.B, v, d+1, k, x0 + 1, ooo, t.t));
		} else if (i + 1 < ql) { // end of the vertex, not the end of the query
			int32_t nv = g_arc_n(g, v), j, n_ext = 0, tw = -1;
			const DevArc *av = g_arc_a(g, v);
			GwfIntv iv;
			iv.vd0 = gwf_gen_vd(v, d), iv.vd1 = iv.vd0 + 1;
			MGB_TRY(avec_push_c(A, z.tmp, iv));
			if (opt.traceback) MGB_TRY(gwf_trace_push(A, z, (int32_t)v, t.t, &tw));
			for (j = 0; j < nv; ++j) {
				uint32_t w = av[j].w;
				int32_t ol = av[j].ow;
				int absent; int64_t slot;
				MGB_TRY(u64tab_put(A, z.ha, (uint64_t)w << 32 | (uint64_t)(uint32_t)(i + 1), &absent, &slot));
				if (q[i + 1] == g_vseq(g, w)[ol]) {
					++n_ext;
					if (absent) {
						GwfDiag p;
						p.vd = gwf_gen_vd(w, i + 1 - ol), p.k = ol, p.xo = (x0 + 2) << 1 | 1, p.t = tw, p.len = 0;
						MGB_TRY(avec_push_c(A, z.Q, p));
					}
				} else if (absent) {
					MGB_TRY(gwf_diag_push(A, z.B, w, i - ol,     ol, x0 + 1, 1, tw));
					MGB_TRY(gwf_diag_push(A, z.B, w, i + 1 - ol, ol, x0 + 2, 1, tw));
				}
			}
			if (nv == 0 || n_ext != nv)
				MGB_TRY(gwf_diag_push(A, z.B, v, d+1, k, x0 + 1, 1, t.t));
		} else if (v1 == (uint32_t)-1 || (v == v1 && k == off1)) { // end of the query at the wanted position
			r->end_v = v, r->end_off = k, r->wlen = (int32_t)(x0 - (uint32_t)i - 1), z.end_tb = t.t;
			z.a.n = 0;
			return 0;
		} else if (k + 1 < vl) { // end of the query, not the end of the vertex
			MGB_TRY(gwf_diag_push(A, z.B, v, d-1, k+1, x0 + 1, ooo, t.t));
		} else if (v != v1) { // end of both, but not on the last vertex
			int32_t nv = g_arc_n(g, v), j, tw = -1;
			const DevArc *av = g_arc_a(g, v);
			if (opt.traceback) MGB_TRY(gwf_trace_push(A, z, (int32_t)v, t.t, &tw));
			for (j = 0; j < nv; ++j)
				MGB_TRY(gwf_diag_push(A, z.B, av[j].w, i - av[j].ow, av[j].ow, x0 + 1, 1, tw));
		}
	}
	n = (int32_t)z.B.n;
	if (do_dedup) MGB_TRY(gwf_dedup(A, z, n, z.B.a, &n));
	if (opt.max_lag > 0 && n > opt.max_chk && ((z.s + 1) & 0xf) == 0)
		n = gwf_prune(n, z.B.a, (uint32_t)opt.max_lag, opt.bw_dyn);
	z.B.n = n;
	{ AVec<GwfDiag> sw = z.a; z.a = z.B; z.B = sw; } // the new wavefront becomes current; the old buffer is recycled
	return 0;
}

// Edit distance from (v0,off0) to (v1,off1) with the walk (reference: gfa-ed.c:552-593 gfa_ed_init + gfa_ed_step).
// r->v (nv vertices) is allocated at the caller's mark; everything else is released.
MG_HD inline int gwf_align(Arena &A, const GraphDev &g, const GwfOpt &opt, int32_t ql, const char *q, uint32_t v0, int32_t off0,
						   uint32_t v1, int32_t off1, int32_t s_term, GwfResult *r)
{
	uint64_t mark = A.top;
	GwfState z;
	z.g = &g, z.ql = ql, z.q = q, z.s = 0, z.end_tb = -1, z.q_head = 0;
	avec_init(z.a), avec_init(z.B), avec_init(z.ooo), avec_init(z.Q), avec_init(z.intv), avec_init(z.tmp), avec_init(z.swap), avec_init(z.t);
	uint64_t mark_keep = mark;
	MGB_TRY(u64tab_init(A, z.ha, 6));
	MGB_TRY(u64tab_init(A, z.ht, 6));
	MGB_TRY(avec_reserve_c(A, z.t, 16));
	{
		GwfDiag d0;
		d0.vd = gwf_gen_vd(v0, -off0), d0.k = off0 - 1, d0.xo = 0, d0.len = 0, d0.t = 0;
		if (opt.traceback) MGB_TRY(gwf_trace_push(A, z, -1, -1, &d0.t));
		MGB_TRY(avec_push_c(A, z.a, d0));
	}
	if (s_term < 0 && opt.s_term >= 0) s_term = opt.s_term;
	r->n_iter = 0, r->nv = 0, r->v = 0, r->end_v = (uint32_t)-1, r->end_off = -1, r->wlen = 0;
	while (z.a.n > 0) {
		MGB_TRY(gwf_ed_extend(A, z, opt, v1, off1, r));
		r->n_iter += z.a.n;
		if (r->end_off >= 0 || z.a.n == 0) break;
		if (s_term >= 0 && z.s >= s_term) break;
		if (opt.i_term > 0 && r->n_iter > opt.i_term) break;
		++z.s;
	}
	if (opt.traceback && r->end_off >= 0) { // reference: gfa-ed.c:509-522 gwf_traceback
		int32_t i = z.end_tb, n = 1;
		while (i >= 0 && z.t.a[i].v >= 0) ++n, i = z.t.a[i].pre;
		int32_t *tmpw, *walk = (int32_t*)(A.base + mark); // the walk is built above the scratch, then moved down to the mark
		MGB_ALLOC(A, tmpw, int32_t, n);
		i = z.end_tb, n = 0;
		tmpw[n++] = (int32_t)r->end_v;
		while (i >= 0 && z.t.a[i].v >= 0) tmpw[n++] = z.t.a[i].v, i = z.t.a[i].pre;
		r->nv = n;
		for (i = 0; i < n >> 1; ++i) { int32_t x = tmpw[i]; tmpw[i] = tmpw[n - 1 - i], tmpw[n - 1 - i] = x; }
		for (i = 0; i < n; ++i) { int32_t x = tmpw[i]; walk[i] = x; } // forward copy: the destination lies below the source
		r->v = walk;
		mark_keep = mark + (((uint64_t)n * 4 + 15) & ~(uint64_t)15);
	}
	r->s = r->end_v != (uint32_t)-1? z.s : -1;
	A.top = mark_keep;
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-cooperative form.  The state lives in one place (shared memory on the device) and is seen by all lanes; the
// per-diagonal work of a wavefront step (extension, the Landau-Vishkin recurrence, the two order-preserving
// compactions) is spread over the lanes, while the vertex-crossing queue, the dedup and the pruning stay on lane 0.
// Results are identical to gwf_align(): every list is written in the order the sequential code writes it.
// ---------------------------------------------------------------------------------------------------------------

struct GwfShared {
	Arena A;
	GwfState z;
	GwfResult r;
};

// one run of n consecutive diagonals of one vertex (gwf_extend_batch); all lanes enter
MG_HD inline int gwf_extend_batch_w(GwfShared *sh, int32_t n, GwfDiag *a, int lane)
{
	GwfState &z = sh->z;
	const GraphDev &g = *z.g;
	const uint32_t v = (uint32_t)(a->vd >> 32);
	const int32_t vl = g_vlen(g, v), ql = z.ql;
	const char *ts = g_vseq(g, v);
	MGB_NO_UNROLL
	for (int32_t j = lane; j < n; j += MGB_W) {
		GwfDiag p = a[j];
		int32_t k = gwf_extend1((int32_t)p.vd - GWF_DIAG_SHIFT, p.k, vl, ts, ql, z.q);
		p.len = k - p.k;
		p.xo += (uint32_t)p.len << 2;
		p.k = k;
		a[j] = p;
	}
	int rc = 0; // codes travel by shuffle, not through shared memory: a flag there could be rewritten before a slow lane has read it
	if (lane == 0) {
		rc = avec_reserve_c(sh->A, z.B, z.B.n + n + 2);
		if (rc == 0) rc = avec_reserve_c(sh->A, z.Q, z.Q.n + n);
		if (rc == 0) rc = avec_reserve_c(sh->A, z.tmp, z.tmp.n + n + 2);
	}
	warp_sync();
	rc = warp_bcast_i32(rc, 0);
	if (rc < 0) return rc;
	GwfDiag *b = &z.B.a[z.B.n];
	GwfDiag *Q = z.Q.a;
	GwfIntv *T = z.tmp.a;
	int64_t qn = z.Q.n, tn = z.tmp.n;
	int32_t m = 0;
	// next-score candidates b[0..n+1], kept in order when still inside the vertex and the query
	MGB_NO_UNROLL
	for (int32_t base = 0; base < n + 2; base += MGB_W) {
		const int32_t j = base + lane;
		GwfDiag p;
		int keep = 0, ends = 0;
		p.vd = 0, p.k = 0, p.xo = 0, p.t = 0, p.len = 0;
		if (j < n + 2) {
			if (j == 0) {
				p.vd = a[0].vd - 1, p.xo = a[0].xo + 2, p.k = a[0].k + 1, p.t = a[0].t;
			} else if (j == n + 1) {
				p.vd = a[n-1].vd + 1, p.xo = a[n-1].xo + 2, p.k = a[n-1].k, p.t = a[n-1].t;
			} else if (j == 1) {
				const int first = n == 1 || a[0].k > a[1].k;
				p.vd = a[0].vd;
				p.xo = first? a[0].xo + 4 : a[1].xo + 2;
				p.t = first? a[0].t : a[1].t;
				p.k = (first? a[0].k : a[1].k) + 1;
			} else if (j == n) {
				const int first = a[n-2].k > a[n-1].k + 1;
				p.vd = a[n-1].vd;
				p.xo = first? a[n-2].xo + 2 : a[n-1].xo + 4;
				p.t = first? a[n-2].t : a[n-1].t;
				p.k = first? a[n-2].k : a[n-1].k + 1;
			} else {
				const GwfDiag l = a[j-2], c = a[j-1], rr = a[j];
				uint32_t x = l.xo + 2;
				int32_t k = l.k, t = l.t;
				x = k > c.k + 1? x : c.xo + 4;
				t = k > c.k + 1? t : c.t;
				k = k > c.k + 1? k : c.k + 1;
				x = k > rr.k + 1? x : rr.xo + 2;
				t = k > rr.k + 1? t : rr.t;
				k = k > rr.k + 1? k : rr.k + 1;
				p.vd = c.vd, p.k = k, p.xo = x, p.t = t;
			}
			const int32_t d = (int32_t)p.vd - GWF_DIAG_SHIFT;
			if (d + p.k < ql && p.k < vl) keep = 1;
			else if (p.k == vl) ends = 1;
		}
		const uint32_t mk = warp_ballot(keep), me = warp_ballot(ends);
		if (keep) b[m + mask_rank(mk, lane)] = p;
		if (ends) {
			GwfIntv iv;
			iv.vd0 = gwf_gen_vd(v, (int32_t)p.vd - GWF_DIAG_SHIFT), iv.vd1 = iv.vd0 + 1;
			T[tn + mask_rank(me, lane)] = iv;
		}
		m += mask_count(mk), tn += mask_count(me);
	}
	// diagonals touching the end of the vertex or of the query go to the queue
	MGB_NO_UNROLL
	for (int32_t base = 0; base < n; base += MGB_W) {
		const int32_t j = base + lane;
		GwfDiag p;
		int f = 0;
		p.vd = 0, p.k = 0, p.xo = 0, p.t = 0, p.len = 0;
		if (j < n) {
			p = a[j];
			f = p.k == vl - 1 || (int32_t)p.vd - GWF_DIAG_SHIFT + p.k == ql - 1;
		}
		const uint32_t mq = warp_ballot(f);
		if (f) {
			p.xo |= 1;
			a[j].xo = p.xo;
			Q[qn + mask_rank(mq, lane)] = p;
		}
		qn += mask_count(mq);
	}
	warp_sync();
	if (lane == 0) z.B.n += m, z.Q.n = qn, z.tmp.n = tn;
	warp_sync();
	return 0;
}

// lane 0: the vertex-crossing queue of gwf_ed_extend()
MG_HD inline int gwf_ed_queue(Arena &A, GwfState &z, const GwfOpt &opt, uint32_t v1, int32_t off1, GwfResult *r, int *want_dedup)
{
	const GraphDev &g = *z.g;
	const int32_t ql = z.ql;
	const char *q = z.q;
	int32_t i, n, do_dedup = z.Q.n != 0;
	MGB_NO_UNROLL
	while (z.q_head < z.Q.n) {
		GwfDiag t = z.Q.a[z.q_head++];
		uint32_t v, x0;
		int32_t ooo, d, k, vl;
		ooo = t.xo & 1, v = (uint32_t)(t.vd >> 32);
		d = (int32_t)t.vd - GWF_DIAG_SHIFT;
		k = t.k;
		vl = g_vlen(g, v);
		k = gwf_extend1(d, k, vl, g_vseq(g, v), ql, q);
		i = k + d;
		x0 = (t.xo >> 1) + ((uint32_t)(k - t.k) << 1);
		if (k + 1 < vl && i + 1 < ql) { // wavefront in the middle of the vertex
			int32_t push1 = 1, push2 = 1;
			if (z.B.n >= 2) push1 = gwf_diag_update(&z.B.a[z.B.n - 2], v, d-1, k+1, x0 + 1, ooo, t.t);
			if (z.B.n >= 1) push2 = gwf_diag_update(&z.B.a[z.B.n - 1], v, d,   k+1, x0 + 2, ooo, t.t);
			if (push1)          MGB_TRY(gwf_diag_push(A, z.B, v, d-1, k+1, x0 + 1, 1, t.t));
			if (push2 || push1) MGB_TRY(gwf_diag_push(A, z.B, v, d,   k+1, x0 + 2, 1, t.t));
			MGB_TRY(gwf_diag_push(A, z.B, v, d+1, k, x0 + 1, ooo, t.t));
		} else if (i + 1 < ql) { // end of the vertex, not the end of the query
			int32_t nv = g_arc_n(g, v), j, n_ext = 0, tw = -1;
			const DevArc *av = g_arc_a(g, v);
			GwfIntv iv;
			iv.vd0 = gwf_gen_vd(v, d), iv.vd1 = iv.vd0 + 1;
			MGB_TRY(avec_push_c(A, z.tmp, iv));
			if (opt.traceback) MGB_TRY(gwf_trace_push(A, z, (int32_t)v, t.t, &tw));
			MGB_NO_UNROLL
			for (j = 0; j < nv; ++j) {
				uint32_t w = av[j].w;
				int32_t ol = av[j].ow;
				int absent; int64_t slot;
				MGB_TRY(u64tab_put(A, z.ha, (uint64_t)w << 32 | (uint64_t)(uint32_t)(i + 1), &absent, &slot));
				if (q[i + 1] == g_vseq(g, w)[ol]) {
					++n_ext;
					if (absent) {
						GwfDiag p;
						p.vd = gwf_gen_vd(w, i + 1 - ol), p.k = ol, p.xo = (x0 + 2) << 1 | 1, p.t = tw, p.len = 0;
						MGB_TRY(avec_push_c(A, z.Q, p));
					}
				} else if (absent) {
					MGB_TRY(gwf_diag_push(A, z.B, w, i - ol,     ol, x0 + 1, 1, tw));
					MGB_TRY(gwf_diag_push(A, z.B, w, i + 1 - ol, ol, x0 + 2, 1, tw));
				}
			}
			if (nv == 0 || n_ext != nv)
				MGB_TRY(gwf_diag_push(A, z.B, v, d+1, k, x0 + 1, 1, t.t));
		} else if (v1 == (uint32_t)-1 || (v == v1 && k == off1)) { // end of the query at the wanted position
			r->end_v = v, r->end_off = k, r->wlen = (int32_t)(x0 - (uint32_t)i - 1), z.end_tb = t.t;
			z.a.n = 0;
			*want_dedup = -1; // done
			return 0;
		} else if (k + 1 < vl) { // end of the query, not the end of the vertex
			MGB_TRY(gwf_diag_push(A, z.B, v, d-1, k+1, x0 + 1, ooo, t.t));
		} else if (v != v1) { // end of both, but not on the last vertex
			int32_t nv = g_arc_n(g, v), j, tw = -1;
			const DevArc *av = g_arc_a(g, v);
			if (opt.traceback) MGB_TRY(gwf_trace_push(A, z, (int32_t)v, t.t, &tw));
			MGB_NO_UNROLL
			for (j = 0; j < nv; ++j)
				MGB_TRY(gwf_diag_push(A, z.B, av[j].w, i - av[j].ow, av[j].ow, x0 + 1, 1, tw));
		}
	}
	(void)n;
	*want_dedup = do_dedup;
	return 0;
}

// lane 0: what follows the dedup in gwf_ed_extend()
MG_HD inline void gwf_ed_finish(GwfState &z, const GwfOpt &opt)
{
	int32_t n = (int32_t)z.B.n;
	if (opt.max_lag > 0 && n > opt.max_chk && ((z.s + 1) & 0xf) == 0)
		n = gwf_prune(n, z.B.a, (uint32_t)opt.max_lag, opt.bw_dyn);
	z.B.n = n;
	{ AVec<GwfDiag> sw = z.a; z.a = z.B; z.B = sw; }
}

MG_HD inline int64_t gwf_lower_bound_vd(const GwfDiag *a, int64_t n, uint64_t key) // number of elements with vd < key
{
	int64_t lo = 0, hi = n;
	MGB_NO_UNROLL
	while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a[mid].vd < key) lo = mid + 1; else hi = mid; }
	return lo;
}
MG_HD inline int64_t gwf_upper_bound_vd(const GwfDiag *a, int64_t n, uint64_t key) // number of elements with vd <= key
{
	int64_t lo = 0, hi = n;
	MGB_NO_UNROLL
	while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a[mid].vd <= key) lo = mid + 1; else hi = mid; }
	return lo;
}

// gwf_dedup() on z.B entered by all lanes: same lists, same order
MG_HD inline int gwf_dedup_w(GwfShared *sh, int lane)
{
	GwfState &z = sh->z;
	Arena &A = sh->A;
	// forbidden intervals: fold the new ones in (nothing changes when there are no new ones)
	const int64_t n_old = z.intv.n, n_new = z.tmp.n;
	int rc0 = 0, tmp_big_unsorted = 0;
	if (n_new > 1) { // the new intervals in order (klib sorts up to 64 elements by insertion: a stable sort, done here by counting)
		GwfIntv *tn = z.tmp.a;
		int uns = 0;
		MGB_NO_UNROLL
		for (int64_t i = 1 + lane; i < n_new; i += MGB_W) if (tn[i-1].vd0 > tn[i].vd0) uns = 1;
		uns = warp_any(uns);
		if (uns && n_new <= 64) small_sort_stable_w(tn, n_new, KeyIntvVd0(), lane);
		else tmp_big_unsorted = uns;
	}
	if (lane == 0) {
		int rc = 0;
		if (n_new > 0) {
			if (tmp_big_unsorted) rc = gwf_sort_intv_cold(A, z.tmp.a, n_new);
			if (rc == 0) rc = avec_reserve_c(A, z.swap, n_old);
			if (rc == 0) z.swap.n = n_old, rc = avec_reserve_c(A, z.intv, n_old + n_new);
		}
		if (rc == 0) rc = avec_reserve_c(A, z.ooo, z.B.n);
		rc0 = rc;
	}
	warp_sync();
	rc0 = warp_bcast_i32(rc0, 0);
	if (rc0 < 0) return rc0;
	if (n_new > 0) {
		GwfIntv *b = z.swap.a, *c = z.tmp.a, *o = z.intv.a;
		MGB_NO_UNROLL
		for (int64_t i = lane; i < n_old; i += MGB_W) b[i] = o[i]; // o may have moved: avec_reserve copied the old content
		warp_sync();
		// stable merge of two sorted lists (the old one first on ties), every element finds its place by bisection
		MGB_NO_UNROLL
		for (int64_t x = lane; x < n_old; x += MGB_W) {
			const uint64_t key = b[x].vd0;
			int64_t lo = 0, hi = n_new;
			MGB_NO_UNROLL
			while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (c[mid].vd0 < key) lo = mid + 1; else hi = mid; }
			o[x + lo] = b[x];
		}
		MGB_NO_UNROLL
		for (int64_t y = lane; y < n_new; y += MGB_W) {
			const uint64_t key = c[y].vd0;
			int64_t lo = 0, hi = n_old;
			MGB_NO_UNROLL
			while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (b[mid].vd0 <= key) lo = mid + 1; else hi = mid; }
			o[y + lo] = c[y];
		}
		warp_sync();
		// gwf_intv_merge_adj: an interval opens a new run when it starts beyond everything seen so far
		const int64_t n_m = n_old + n_new;
		int64_t k = 0;
		uint64_t carry = 0; // running maximum of vd1 over the elements of the earlier rounds
		MGB_NO_UNROLL
		for (int64_t base = 0; base < n_m; base += MGB_W) {
			const int64_t i = base + lane;
			GwfIntv e;
			e.vd0 = e.vd1 = 0;
			if (i < n_m) e = o[i];
			uint64_t pm = warp_incl_scan_max_u64(i < n_m? e.vd1 : 0, lane);
			if (pm < carry) pm = carry;
			uint64_t before = warp_shfl_up1_u64(pm);
			if (lane == 0) before = carry;
			const int head = i < n_m && (i == 0 || e.vd0 > before);
			const uint32_t mh = warp_ballot(head);
			if (head) {
				const int64_t g = k + mask_rank(mh, lane);
				o[g].vd0 = e.vd0;
				if (g > 0) o[g - 1].vd1 = before;
			}
			k += mask_count(mh);
			carry = warp_bcast_u64(pm, MGB_W - 1);
			warp_sync();
		}
		if (lane == 0) o[k - 1].vd1 = carry, z.intv.n = k;
		warp_sync();
	}
	GwfDiag *a = z.B.a;
	const int32_t n_a = (int32_t)z.B.n;
	// ---- gwf_diag_dedup: sort if needed ----
	int unsorted = 0;
	MGB_NO_UNROLL
	for (int32_t i = 1 + lane; i < n_a; i += MGB_W) if (a[i-1].vd > a[i].vd) unsorted = 1;
	unsorted = warp_any(unsorted);
	if (unsorted) { // gwf_diag_sort: in-order part and out-of-order part, the latter sorted, then a stable merge
		GwfDiag *b = z.ooo.a;
		int32_t n_c = 0;
		MGB_NO_UNROLL
		for (int32_t base = 0; base < n_a; base += MGB_W) {
			const int32_t i = base + lane;
			n_c += mask_count(warp_ballot(i < n_a && (a[i].xo & 1)));
		}
		const int32_t n_b = n_a - n_c;
		GwfDiag *c = b + n_b;
		int32_t jb = 0, jc = 0;
		MGB_NO_UNROLL
		for (int32_t base = 0; base < n_a; base += MGB_W) {
			const int32_t i = base + lane;
			GwfDiag p;
			p.vd = 0, p.k = 0, p.xo = 0, p.t = 0, p.len = 0;
			if (i < n_a) p = a[i];
			const uint32_t mc = warp_ballot(i < n_a && (p.xo & 1)), mb = warp_ballot(i < n_a && !(p.xo & 1));
			if (i < n_a) {
				if (p.xo & 1) c[jc + mask_rank(mc, lane)] = p;
				else b[jb + mask_rank(mb, lane)] = p;
			}
			jc += mask_count(mc), jb += mask_count(mb);
		}
		warp_sync();
		if (n_c <= 64) small_sort_stable_w(c, n_c, KeyDiagVd(), lane); // what klib's insertion sort of a short list yields
		else {
			if (lane == 0) rc0 = gwf_sort_diag_cold(A, c, n_c);
			warp_sync();
			rc0 = warp_bcast_i32(rc0, 0);
			if (rc0 < 0) return rc0;
		}
		MGB_NO_UNROLL
		for (int32_t j = lane; j < n_c; j += MGB_W) c[j].xo &= 0xfffffffeU;
		warp_sync();
		MGB_NO_UNROLL
		for (int32_t i = lane; i < n_b; i += MGB_W) a[i + gwf_lower_bound_vd(c, n_c, b[i].vd)] = b[i];
		MGB_NO_UNROLL
		for (int32_t j = lane; j < n_c; j += MGB_W) a[j + gwf_upper_bound_vd(b, n_b, c[j].vd)] = c[j];
		warp_sync();
	}
	// ---- one diagonal per (vertex,diag): the first one reaching furthest ----
	int32_t n = 0;
	MGB_NO_UNROLL
	for (int32_t base = 0; base < n_a; base += MGB_W) {
		const int32_t i = base + lane;
		GwfDiag best;
		int head = 0;
		best.vd = 0, best.k = 0, best.xo = 0, best.t = 0, best.len = 0;
		if (i < n_a) {
			best = a[i];
			head = i == 0 || a[i-1].vd != best.vd;
			if (head)
				MGB_NO_UNROLL
				for (int32_t j = i + 1; j < n_a && a[j].vd == best.vd; ++j)
					if (best.k < a[j].k) best = a[j];
		}
		const uint32_t mh = warp_ballot(head);
		warp_sync();
		if (head) a[n + mask_rank(mh, lane)] = best;
		n += mask_count(mh);
		warp_sync();
	}
	// ---- drop diagonals inside forbidden bands ----
	if (z.intv.n > 0) {
		const GwfIntv *iv = z.intv.a;
		const int64_t n_iv = z.intv.n;
		int32_t k = 0;
		MGB_NO_UNROLL
		for (int32_t base = 0; base < n; base += MGB_W) {
			const int32_t i = base + lane;
			GwfDiag p;
			int keep = 0;
			p.vd = 0, p.k = 0, p.xo = 0, p.t = 0, p.len = 0;
			if (i < n) {
				p = a[i];
				int64_t lo = 0, hi = n_iv; // intervals are disjoint and sorted: the last one starting at or before vd decides
				MGB_NO_UNROLL
				while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (iv[mid].vd0 <= p.vd) lo = mid + 1; else hi = mid; }
				keep = !(lo > 0 && p.vd < iv[lo - 1].vd1);
			}
			const uint32_t mk = warp_ballot(keep);
			warp_sync();
			if (keep) a[k + mask_rank(mk, lane)] = p;
			k += mask_count(mk);
			warp_sync();
		}
		n = k;
	}
	warp_sync();
	if (lane == 0) z.B.n = n;
	warp_sync();
	return 0;
}

// gwf_align() entered by all lanes of a warp.  `sh` is visible to every lane; sh->A is the arena to work in.  On return
// sh->r holds the result (r.v at the arena's mark, as gwf_align leaves it) and the return code is the same on all lanes.
MG_HD inline int gwf_align_w(GwfShared *sh, const GraphDev &g, const GwfOpt &opt, int32_t ql, const char *q, uint32_t v0, int32_t off0,
							 uint32_t v1, int32_t off1, int32_t s_term, int lane)
{
	Arena &A = sh->A;
	GwfState &z = sh->z;
	GwfResult *r = &sh->r;
	const uint64_t mark = A.top;
	if (s_term < 0 && opt.s_term >= 0) s_term = opt.s_term;
	int code = 0, go = 0; // replicated on every lane, set by lane 0 and spread by shuffle
	if (lane == 0) {
		int rc = 0;
		z.g = &g, z.ql = ql, z.q = q, z.s = 0, z.end_tb = -1, z.q_head = 0;
		avec_init(z.a), avec_init(z.B), avec_init(z.ooo), avec_init(z.Q), avec_init(z.intv), avec_init(z.tmp), avec_init(z.swap), avec_init(z.t);
		rc = u64tab_init(A, z.ha, 6);
		if (rc == 0) rc = u64tab_init(A, z.ht, 6);
		if (rc == 0) rc = avec_reserve_c(A, z.t, 16);
		if (rc == 0 && A.cap >= (1u << 16)) { // in the worker's arena: room for a typical wavefront up front (growth abandons the old block and copies on one lane)
			rc = avec_reserve_c(A, z.a, 96);
			if (rc == 0) rc = avec_reserve_c(A, z.B, 192);
			if (rc == 0) rc = avec_reserve_c(A, z.ooo, 192);
			if (rc == 0) rc = avec_reserve_c(A, z.Q, 96);
			if (rc == 0) rc = avec_reserve_c(A, z.tmp, 64);
			if (rc == 0) rc = avec_reserve_c(A, z.intv, 64);
			if (rc == 0) rc = avec_reserve_c(A, z.swap, 64);
			if (rc == 0) rc = avec_reserve_c(A, z.t, 64);
		}
		if (rc == 0) {
			GwfDiag d0;
			d0.vd = gwf_gen_vd(v0, -off0), d0.k = off0 - 1, d0.xo = 0, d0.len = 0, d0.t = 0;
			if (opt.traceback) rc = gwf_trace_push(A, z, -1, -1, &d0.t);
			if (rc == 0) rc = avec_push_c(A, z.a, d0);
		}
		r->n_iter = 0, r->nv = 0, r->v = 0, r->end_v = (uint32_t)-1, r->end_off = -1, r->wlen = 0;
		code = rc, go = rc == 0 && z.a.n > 0;
	}
	warp_sync();
	code = warp_bcast_i32(code, 0), go = warp_bcast_i32(go, 0);
	MGB_NO_UNROLL
	while (go) {
		// ---- one edit-distance step (gwf_ed_extend) ----
		if (lane == 0) {
			r->end_v = (uint32_t)-1;
			r->end_off = z.end_tb = -1;
			z.tmp.n = 0;
			u64tab_clear(z.ha);
			z.Q.n = 0, z.q_head = 0;
			z.B.n = 0;
			code = avec_reserve_c(A, z.B, z.a.n * 2);
		}
		warp_sync();
		code = warp_bcast_i32(code, 0);
		if (code < 0) break;
		{
			const int32_t n = (int32_t)z.a.n;
			GwfDiag *a = z.a.a;
			int32_t x = 0, rc = 0;
			MGB_NO_UNROLL
			for (int32_t base = 1; base <= n && rc == 0; base += MGB_W) { // runs of consecutive diagonals, in order
				const int32_t i = base + lane;
				uint32_t mask = warp_ballot(i <= n && (i == n || a[i].vd != a[i-1].vd + 1));
				MGB_NO_UNROLL
				while (mask && rc == 0) {
					const int32_t e = base + ctz32(mask);
					mask &= mask - 1;
					rc = gwf_extend_batch_w(sh, e - x, &a[x], lane);
					x = e;
				}
			}
			if (rc < 0) { code = rc; break; }
		}
		int dd = 0;
		if (lane == 0) code = gwf_ed_queue(A, z, opt, v1, off1, r, &dd);
		warp_sync();
		code = warp_bcast_i32(code, 0), dd = warp_bcast_i32(dd, 0);
		if (code < 0) break;
		if (dd > 0 && (code = gwf_dedup_w(sh, lane)) < 0) break;
		if (lane == 0) {
			if (dd >= 0) gwf_ed_finish(z, opt);
			go = 1;
			r->n_iter += z.a.n;
			if (r->end_off >= 0 || z.a.n == 0) go = 0;
			else if (s_term >= 0 && z.s >= s_term) go = 0;
			else if (opt.i_term > 0 && r->n_iter > opt.i_term) go = 0;
			else ++z.s;
		}
		warp_sync();
		go = warp_bcast_i32(go, 0);
	}
	warp_sync();
	if (code < 0) { if (lane == 0) A.top = mark; warp_sync(); return code; }
	if (lane == 0) {
		uint64_t mark_keep = mark;
		int rc = 0;
		if (opt.traceback && r->end_off >= 0) { // reference: gfa-ed.c:509-522 gwf_traceback
			int32_t i = z.end_tb, n = 1;
			MGB_NO_UNROLL
			while (i >= 0 && z.t.a[i].v >= 0) ++n, i = z.t.a[i].pre;
			int32_t *walk = (int32_t*)(A.base + mark);
			int32_t *tmpw = (int32_t*)arena_alloc(A, (uint64_t)n * sizeof(int32_t));
			if (tmpw == 0) rc = MGB_E_ARENA;
			else {
				i = z.end_tb, n = 0;
				tmpw[n++] = (int32_t)r->end_v;
				MGB_NO_UNROLL
				while (i >= 0 && z.t.a[i].v >= 0) tmpw[n++] = z.t.a[i].v, i = z.t.a[i].pre;
				r->nv = n;
				MGB_NO_UNROLL
				for (i = 0; i < n >> 1; ++i) { int32_t t = tmpw[i]; tmpw[i] = tmpw[n - 1 - i], tmpw[n - 1 - i] = t; }
				MGB_NO_UNROLL
				for (i = 0; i < n; ++i) { int32_t t = tmpw[i]; walk[i] = t; } // forward copy: the destination lies below the source
				r->v = walk;
				mark_keep = mark + (((uint64_t)n * 4 + 15) & ~(uint64_t)15);
			}
		}
		r->s = r->end_v != (uint32_t)-1? z.s : -1;
		A.top = rc == 0? mark_keep : mark;
		code = rc;
	}
	warp_sync();
	return warp_bcast_i32(code, 0);
}

} // namespace mgb
