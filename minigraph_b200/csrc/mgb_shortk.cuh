// mgb_shortk.cuh -- bounded k-shortest walks from one vertex to a set of destination vertices.
// (reference: shortk.c:41-242 mg_shortest_k).  Best-first search over the flattened arc CSR; every vertex keeps at
// most MAX_SHORT_K arrivals.  The frontier key is dist<<32|id with `id` a global insertion counter, so keys are
// unique and any exact priority queue pops in the reference's order: the AVL tree becomes an indexed
// binary heap, the visited-vertex hash becomes an open-addressing table in the worker arena.
#pragma once
#include "mgb_model.cuh"

namespace mgb {

// reference: mgpriv.h:40-52 mg_path_dst_t
struct PathDst {
	uint32_t v;
	int32_t target_dist;
	uint32_t target_hash;
	int32_t meta, check_hash, inner;
	int32_t qlen;
	int32_t n_path, is_0;
	int32_t path_end;
	int32_t dist;
	uint32_t hash;
};

// reference: mgpriv.h:54-57 mg_pathv_t
struct PathV {
	uint32_t v, d;
	int32_t pre;
};

struct SpNode {
	uint64_t di;       // dist<<32 | unique id (later: dist<<32 | position in out[])
	uint32_t v;
	int32_t pre;
	uint32_t hash;
	int32_t is_0;
	int32_t heap_pos;  // position in the frontier heap, -1 when not in it
};

struct SpTopK {
	uint32_t v;
	int32_t k;
	int32_t p[MAX_SHORT_K];  // max-heap of node indices by di ...
	uint64_t d[MAX_SHORT_K]; // ... with the keys beside them (a comparison then stays inside this record)
};

struct SpHeapEnt { uint64_t di; int32_t node, pad; }; // frontier entry: the key travels with the node index

struct SpState {
	AVec<SpNode> nd;
	AVec<SpHeapEnt> heap;   // frontier: min-heap by di (keys are unique, so any exact heap pops in the reference's order)
	AVec<SpTopK> topk;
	int64_t *htab;          // vertex -> index into topk (open addressing): vertex<<32 | index, -1 empty
	int32_t htab_bits;
};

MG_HD inline void sp_heap_set(SpState &S, int32_t i, const SpHeapEnt &e) { S.heap.a[i] = e; S.nd.a[e.node].heap_pos = i; }
// 4-ary heap: half the levels of a binary one, and the four children of a node are 64 contiguous bytes
MG_HD inline void sp_heap_up(SpState &S, int32_t i)
{
	const SpHeapEnt e = S.heap.a[i];
	while (i > 0) {
		const int32_t par = (i - 1) >> 2;
		const SpHeapEnt pe = S.heap.a[par];
		if (pe.di <= e.di) break;
		sp_heap_set(S, i, pe);
		i = par;
	}
	sp_heap_set(S, i, e);
}
MG_HD inline void sp_heap_down(SpState &S, int32_t i)
{
	const int32_t n = (int32_t)S.heap.n;
	const SpHeapEnt e = S.heap.a[i];
	for (;;) {
		const int32_t c0 = 4 * i + 1;
		if (c0 >= n) break;
		const int32_t nc = n - c0 < 4? n - c0 : 4;
		SpHeapEnt c = S.heap.a[c0];
		int32_t m = c0;
		for (int32_t j = 1; j < nc; ++j) { const SpHeapEnt x = S.heap.a[c0 + j]; if (x.di < c.di) c = x, m = c0 + j; }
		if (e.di <= c.di) break;
		sp_heap_set(S, i, c);
		i = m;
	}
	sp_heap_set(S, i, e);
}
MG_HD inline int sp_heap_push(Arena &A, SpState &S, int32_t node)
{
	SpHeapEnt e;
	e.di = S.nd.a[node].di, e.node = node, e.pad = 0;
	MGB_TRY(avec_push(A, S.heap, e));
	S.nd.a[node].heap_pos = (int32_t)S.heap.n - 1;
	sp_heap_up(S, (int32_t)S.heap.n - 1);
	return 0;
}
MG_HD inline void sp_heap_remove_at(SpState &S, int32_t pos)
{
	const int32_t last = (int32_t)S.heap.n - 1;
	const int32_t node = S.heap.a[pos].node;
	--S.heap.n;
	if (pos != last) {
		sp_heap_set(S, pos, S.heap.a[last]);
		sp_heap_up(S, pos);
		sp_heap_down(S, S.nd.a[S.heap.a[last].node].heap_pos); // a[last] still holds the moved entry's node
	}
	S.nd.a[node].heap_pos = -1;
}

MG_HD inline int sp_htab_get(Arena &A, SpState &S, uint32_t v, int *absent, int32_t *idx_)
{
	for (;;) {
		uint32_t mask = (1u << S.htab_bits) - 1, h = hash32(v) & mask;
		int64_t e;
		while ((e = S.htab[h]) >= 0 && (uint32_t)(e >> 32) != v) h = (h + 1) & mask;
		if (e >= 0) { *absent = 0, *idx_ = (int32_t)e; return 0; }
		if ((uint64_t)(S.topk.n + 1) * 2 > (1ULL << S.htab_bits)) { // grow and rehash
			int32_t nb = S.htab_bits + 1;
			int64_t *nt;
			MGB_ALLOC(A, nt, int64_t, 1LL << nb);
			uint32_t nmask = (1u << nb) - 1;
			for (int64_t i = 0; i < (1LL << nb); ++i) nt[i] = -1;
			for (int64_t i = 0; i < S.topk.n; ++i) {
				uint32_t g = hash32(S.topk.a[i].v) & nmask;
				while (nt[g] >= 0) g = (g + 1) & nmask;
				nt[g] = (int64_t)S.topk.a[i].v << 32 | i;
			}
			S.htab = nt, S.htab_bits = nb;
			continue;
		}
		SpTopK t;
		t.v = v, t.k = 0;
		MGB_TRY(avec_push(A, S.topk, t));
		S.htab[h] = (int64_t)v << 32 | (S.topk.n - 1);
		*absent = 1, *idx_ = (int32_t)S.topk.n - 1;
		return 0;
	}
}

// per-vertex top-k max-heap on di (reference: ksort.h:42-65 ks_heapup/ks_heapdown with sp_node_lt); only its root, the
// longest of the kept arrivals, is ever consulted, and keys are unique
MG_HD inline void sp_topk_up(int32_t n, int32_t *l, uint64_t *d)
{
	int32_t k = n - 1, tmp = l[k];
	const uint64_t td = d[k];
	while (k) {
		int32_t i = (k - 1) >> 1;
		if (td < d[i]) break;
		l[k] = l[i], d[k] = d[i], k = i;
	}
	l[k] = tmp, d[k] = td;
}
MG_HD inline void sp_topk_down(int32_t i, int32_t n, int32_t *l, uint64_t *d)
{
	int32_t k = i, tmp = l[i];
	const uint64_t td = d[i];
	while ((k = (k << 1) + 1) < n) {
		if (k != n - 1 && d[k] < d[k+1]) ++k;
		if (d[k] < td) break;
		l[i] = l[k], d[i] = d[k], i = k;
	}
	l[i] = tmp, d[i] = td;
}

// Returns 0 or an error code.  If pathv_/n_pathv_ are non-null the compacted backtrack array is produced in the
// arena (above the caller's mark); dst[].path_end then indexes it.
MG_HD inline int shortest_k(Arena &A, const GraphDev &g, uint32_t src, int32_t n_dst, PathDst *dst, int32_t max_dist, int32_t max_k,
							PathV **pathv_, int32_t *n_pathv_)
{
	if (n_pathv_) *n_pathv_ = 0;
	if (pathv_) *pathv_ = 0;
	if (n_dst <= 0) return 0;
	for (int32_t i = 0; i < n_dst; ++i) {
		PathDst *t = &dst[i];
		if (t->inner) t->dist = 0, t->n_path = 1, t->path_end = -1;
		else t->dist = -1, t->n_path = 0, t->path_end = -1;
	}
	if (max_k > MAX_SHORT_K) max_k = MAX_SHORT_K;
	// the result (if requested) must sit below the scratch: reserve it lazily at the end by copying down
	uint64_t mark = A.top;
	int8_t *dst_done;
	uint64_t *dst_group;
	MGB_ALLOC(A, dst_done, int8_t, n_dst);
	MGB_ALLOC(A, dst_group, uint64_t, n_dst);
	for (int32_t i = 0; i < n_dst; ++i) dst_done[i] = 0, dst_group[i] = (uint64_t)dst[i].v << 32 | (uint64_t)i;
	MGB_TRY(radix_sort_64(A, dst_group, n_dst));

	SpState S;
	avec_init(S.nd), avec_init(S.heap), avec_init(S.topk);
	MGB_TRY(avec_reserve(A, S.nd, 1024)); // growth abandons the old block and copies: start where most searches end
	MGB_TRY(avec_reserve(A, S.heap, 256));
	MGB_TRY(avec_reserve(A, S.topk, 64));
	S.htab_bits = 7;
	MGB_ALLOC(A, S.htab, int64_t, 1 << S.htab_bits);
	for (int i = 0; i < (1 << S.htab_bits); ++i) S.htab[i] = -1;
	AVec<int32_t> out;
	avec_init(out);
	MGB_TRY(avec_reserve(A, out, 1024));

	uint32_t id = 0;
	{
		SpNode p;
		int absent; int32_t qi;
		p.v = src, p.di = (uint64_t)0 << 32 | id++, p.pre = -1, p.is_0 = 1, p.hash = hash32(src), p.heap_pos = -1;
		MGB_TRY(avec_push(A, S.nd, p));
		MGB_TRY(sp_heap_push(A, S, 0));
		MGB_TRY(sp_htab_get(A, S, src, &absent, &qi));
		S.topk.a[qi].k = 1, S.topk.a[qi].p[0] = 0, S.topk.a[qi].d[0] = S.nd.a[0].di;
	}
	int32_t n_done = 0;
	while (S.heap.n > 0) {
		int32_t ri = S.heap.a[0].node;
		sp_heap_remove_at(S, 0);
		int32_t n_out = (int32_t)out.n;
		S.nd.a[ri].di = S.nd.a[ri].di >> 32 << 32 | (uint64_t)(uint32_t)n_out;
		MGB_TRY(avec_push(A, out, ri));
		n_out = (int32_t)out.n;
		const uint32_t rv = S.nd.a[ri].v, rhash = S.nd.a[ri].hash;
		const int32_t rdist = (int32_t)(S.nd.a[ri].di >> 32), ris0 = S.nd.a[ri].is_0;
		{ // is rv a destination?  (binary search the grouped list)
			int32_t lo = 0, hi = n_dst;
			while (lo < hi) { int32_t mid = (lo + hi) >> 1; if ((uint32_t)(dst_group[mid] >> 32) < rv) lo = mid + 1; else hi = mid; }
			if (lo < n_dst && (uint32_t)(dst_group[lo] >> 32) == rv) {
				int32_t off = lo, cnt = 0;
				while (off + cnt < n_dst && (uint32_t)(dst_group[off + cnt] >> 32) == rv) ++cnt;
				for (int32_t j = 0; j < cnt; ++j) {
					PathDst *t = &dst[(int32_t)dst_group[off + j]];
					int32_t done = 0;
					if (t->inner) {
						done = 1;
					} else {
						int32_t copy = 0;
						if (t->n_path == 0) {
							copy = 1;
						} else if (t->target_dist >= 0) {
							if (rdist == t->target_dist && t->check_hash && rhash == t->target_hash) {
								copy = 1, done = 1;
							} else {
								int32_t d0 = t->dist, d1 = rdist;
								d0 = d0 > t->target_dist? d0 - t->target_dist : t->target_dist - d0;
								d1 = d1 > t->target_dist? d1 - t->target_dist : t->target_dist - d1;
								if (d1 < d0) copy = 1;
							}
						}
						if (copy) {
							t->path_end = n_out - 1, t->dist = rdist, t->hash = rhash, t->is_0 = ris0;
							if (t->target_dist >= 0) {
								if (rdist == t->target_dist && t->check_hash && rhash == t->target_hash) done = 1;
								else if (rdist > t->target_dist + 1000) done = 1; // MG_SHORT_K_EXT
							}
						}
						++t->n_path;
						if (t->n_path >= max_k) done = 1;
					}
					if (dst_done[off + j] == 0 && done) dst_done[off + j] = 1, ++n_done;
				}
				if (n_done == n_dst) break;
			}
		}
		int32_t nv = g_arc_n(g, rv);
		const DevArc *av = g_arc_a(g, rv);
		for (int32_t i = 0; i < nv; ++i) {
			const DevArc *ai = &av[i];
			int32_t d = (int32_t)((uint32_t)rdist + ai->lv);
			if (d > max_dist) continue;
			int absent; int32_t qi;
			MGB_TRY(sp_htab_get(A, S, ai->w, &absent, &qi));
			SpTopK *q = &S.topk.a[qi];
			if (q->k < max_k) {
				SpNode p;
				p.v = ai->w, p.di = (uint64_t)(uint32_t)d << 32 | id++, p.pre = n_out - 1;
				p.hash = rhash + hash32(ai->w);
				p.is_0 = ris0;
				if (ai->rank > 0) p.is_0 = 0;
				p.heap_pos = -1;
				MGB_TRY(avec_push(A, S.nd, p));
				int32_t pi = (int32_t)S.nd.n - 1;
				MGB_TRY(sp_heap_push(A, S, pi));
				q = &S.topk.a[qi];
				q->p[q->k] = pi, q->d[q->k] = p.di, ++q->k;
				sp_topk_up(q->k, q->p, q->d);
			} else if ((int64_t)(q->d[0] >> 32) > (int64_t)d) {
				int32_t pi = q->p[0];
				if (S.nd.a[pi].heap_pos < 0) { A.top = mark; return MGB_E_INTERNAL; } // "logical bug" branch of the reference
				sp_heap_remove_at(S, S.nd.a[pi].heap_pos);
				SpNode *p = &S.nd.a[pi];
				p->di = (uint64_t)(uint32_t)d << 32 | id++;
				p->pre = n_out - 1;
				p->hash = rhash + hash32(ai->w);
				p->is_0 = ris0;
				if (ai->rank > 0) p->is_0 = 0;
				MGB_TRY(sp_heap_push(A, S, pi));
				q->d[0] = p->di;
				sp_topk_down(0, q->k, q->p, q->d);
			}
		}
	}

	int32_t n_found = 0;
	for (int32_t i = 0; i < n_dst; ++i) if (dst[i].n_path > 0) ++n_found;
	PathV *ret = 0;
	int32_t n_ret = 0;
	if (n_found > 0 && n_pathv_) { // backtrack array (reference: shortk.c:198-232)
		int32_t n_out = (int32_t)out.n, n = 0, *trans;
		MGB_ALLOC(A, trans, int32_t, n_out);
		for (int32_t i = 0; i < n_out; ++i) trans[i] = 0;
		for (int32_t i = 0; i < n_dst; ++i) {
			PathDst *t = &dst[i];
			if (t->n_path > 0 && t->target_dist >= 0 && t->path_end >= 0)
				trans[(int32_t)S.nd.a[out.a[t->path_end]].di] = 1;
		}
		for (int32_t i = 0; i < n_out; ++i) {
			uint32_t ov = S.nd.a[out.a[i]].v;
			int32_t lo = 0, hi = n_dst;
			while (lo < hi) { int32_t mid = (lo + hi) >> 1; if ((uint32_t)(dst_group[mid] >> 32) < ov) lo = mid + 1; else hi = mid; }
			if (lo < n_dst && (uint32_t)(dst_group[lo] >> 32) == ov) {
				int32_t off = lo, cnt = 0;
				while (off + cnt < n_dst && (uint32_t)(dst_group[off + cnt] >> 32) == ov) ++cnt;
				for (int32_t j = off; j < off + cnt; ++j) // NB: the reference indexes dst[] by group position here
					if (dst[j].target_dist < 0) trans[i] = 1;
			}
		}
		for (int32_t i = n_out - 1; i >= 0; --i)
			if (trans[i] && S.nd.a[out.a[i]].pre >= 0) trans[S.nd.a[out.a[i]].pre] = 1;
		for (int32_t i = 0; i < n_out; ++i) {
			if (trans[i]) trans[i] = n++;
			else trans[i] = -1;
		}
		n_ret = n;
		PathV *tmp;
		MGB_ALLOC(A, tmp, PathV, n);
		for (int32_t i = 0; i < n_out; ++i) {
			if (trans[i] < 0) continue;
			PathV *p = &tmp[trans[i]];
			const SpNode &o = S.nd.a[out.a[i]];
			p->v = o.v, p->d = (uint32_t)(o.di >> 32);
			p->pre = o.pre < 0? o.pre : trans[o.pre];
		}
		for (int32_t i = 0; i < n_dst; ++i)
			if (dst[i].path_end >= 0) dst[i].path_end = trans[dst[i].path_end];
		// move the result down to the caller's mark so that all scratch can be released
		ret = (PathV*)(A.base + mark);
		for (int32_t i = 0; i < n; ++i) { PathV x = tmp[i]; ret[i] = x; } // forward copy is safe: destination is below the source
		A.top = mark + (((uint64_t)n * sizeof(PathV) + 15) & ~(uint64_t)15);
		if (A.top > A.peak) A.peak = A.top;
	} else A.top = mark;
	if (pathv_) *pathv_ = ret;
	if (n_pathv_) *n_pathv_ = n_ret;
	return 0;
}

} // namespace mgb
