// mgb_gclabel.cuh -- reachability labels of the graph: what graph chaining asks mg_shortest_k() for, computed once per source vertex.
//
// The reference calls mg_shortest_k(src, dst[], max_dist, 15) once per linear chain of every read (gchain1.c:174) and that search
// is where graph chaining spends its time (shortk.c:41-242).  What it returns for one destination vertex t depends on nothing
// but the walks from src to t:
//   * the search pops walk ends in (dist, insertion id) order and keeps at most MAX_SHORT_K arrivals per vertex; a push is
//     dropped when dist > max_dist.  Dropped pushes have no effect on the order of the pushes that are kept, so the pop sequence
//     of a search bounded by D is the pop sequence of a search bounded by D' >= D with the pops beyond D left out;
//   * a destination's result (dist, hash, is_0, n_path != 0) is the first of its arrivals, in pop order, that minimises
//     |dist - target_dist| (shortk.c:116-133: a later arrival replaces the kept one only when it is strictly closer), and the
//     search stops only when every destination is final (15 arrivals seen, or the kept one is beyond target_dist + 1000 and so
//     cannot be improved by the longer ones that follow) or the frontier is empty.
// Hence the per-vertex arrival lists of ONE exhaustive search from src, bounded by the largest max_dist any read can ask for
// (max_dist_g + len(src)), answer every query a read can make from that source.  A source is searched the first time a batch
// needs it and its labels stay in HBM for the batches that follow (at 30x coverage a source serves thousands of reads).
//
// Record of one source in the label pool:  LabRec | u32 tv[n_tv] (ascending) | u32 start[n_tv + 1] | pad to 8 | u64 lab[n_lab]
// with lab = (dist << 1 | is_0) << 32 | hash, the arrivals of tv[i] being lab[start[i] .. start[i + 1]) in pop order.
#pragma once
#include "mgb_model.cuh"

namespace mgb {

static const int64_t LAB_NONE = -1, LAB_WANTED = -2;

struct LabRec { int32_t n_tv, n_lab; };

struct LabTab {
	long long *src_off;     // [2 * n_seg] LAB_NONE, LAB_WANTED (listed in new_src by this batch) or the byte offset of the source's record
	Pool *pool_hdr;
	char *pool;
	int32_t *new_src;       // sources wanted for the first time by the batch in flight
	unsigned int *n_new;    // n_new[0]: their number; n_new[1]: how many of them outgrew a thread's share of the arena (listed from the top of new_src[])
	int32_t cap_new;        // entries of new_src[]
	int32_t max_dist_g;     // the bound the table was built for
};

MG_HD inline uint64_t lab_rec_bytes(int32_t n_tv, int32_t n_lab)
{
	uint64_t head = sizeof(LabRec) + (uint64_t)n_tv * 4 + ((uint64_t)n_tv + 1) * 4;
	head = (head + 7) & ~(uint64_t)7;
	return head + (uint64_t)n_lab * 8;
}
MG_HD inline const uint32_t *lab_tv(const char *rec) { return (const uint32_t*)(rec + sizeof(LabRec)); }
MG_HD inline const uint32_t *lab_start(const char *rec) { return lab_tv(rec) + ((const LabRec*)rec)->n_tv; }
MG_HD inline const uint64_t *lab_arr(const char *rec)
{
	const LabRec *h = (const LabRec*)rec;
	uint64_t head = sizeof(LabRec) + (uint64_t)h->n_tv * 4 + ((uint64_t)h->n_tv + 1) * 4;
	return (const uint64_t*)(rec + ((head + 7) & ~(uint64_t)7));
}

// a linear chain far from both ends of its segment, or small against that distance, is left out of graph chaining (gchain1.c:79-91)
MG_HD inline bool gc_isolated(const GraphDev &g, const LChain &r, int32_t max_dist_g)
{
	const int32_t tail = g.seg_len[r.v >> 1] - r.re, end_dist = r.rs < tail? r.rs : tail;
	return end_dist > max_dist_g || (end_dist >> 3) > r.score;
}

// a read needs the labels of source v: list it for this batch's search kernel unless somebody already has
MG_HD inline void lab_want(const LabTab &T, uint32_t v)
{
	if (T.src_off == 0) return;
#if MGB_ON_DEVICE
	if (T.src_off[v] != LAB_NONE) return;
	if (atomicCAS((unsigned long long*)&T.src_off[v], (unsigned long long)LAB_NONE, (unsigned long long)LAB_WANTED) == (unsigned long long)LAB_NONE)
		T.new_src[atomicAdd(T.n_new, 1u)] = (int32_t)v;
#else
	if (T.src_off[v] == LAB_NONE) T.src_off[v] = LAB_WANTED, T.new_src[(*T.n_new)++] = (int32_t)v;
#endif
}

// The answer mg_shortest_k() gives for destination vertex t (target distance `target` >= 0) in a search bounded by max_dist.
// Returns 0 when t is not reached within max_dist (the reference's n_path == 0).
MG_HD inline int lab_query(const char *rec, uint32_t t, int32_t max_dist, int32_t target, int32_t *dist, uint32_t *hash, int32_t *is_0)
{
	const LabRec *h = (const LabRec*)rec;
	const uint32_t *tv = lab_tv(rec);
	int32_t lo = 0, hi = h->n_tv;
	while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (tv[mid] < t) lo = mid + 1; else hi = mid; }
	if (lo >= h->n_tv || tv[lo] != t) return 0;
	const uint32_t *st = lab_start(rec);
	const uint64_t *lab = lab_arr(rec);
	int32_t best = -1, best_gap = 0;
	for (uint32_t i = st[lo]; i < st[lo + 1]; ++i) {
		const uint32_t di = (uint32_t)(lab[i] >> 32);
		const int32_t d = (int32_t)(di >> 1);
		if (d > max_dist) break; // arrivals are in ascending distance
		const int32_t gap = d > target? d - target : target - d;
		if (best < 0 || gap < best_gap) best = (int32_t)i, best_gap = gap;
	}
	if (best < 0) return 0;
	const uint32_t di = (uint32_t)(lab[best] >> 32);
	*dist = (int32_t)(di >> 1), *is_0 = (int32_t)(di & 1), *hash = (uint32_t)lab[best];
	return 1;
}

// ---- the exhaustive bounded search from one source ----

struct LsNode { uint64_t di; uint32_t v, hash; int32_t is_0, pad; };          // one walk end: dist << 32 | insertion id
struct LsEnt { uint64_t di; int32_t node, pad; };                             // frontier entry; stale once the node was re-keyed
struct LsVtx { uint32_t v; int32_t k; int32_t node[MAX_SHORT_K]; uint64_t d[MAX_SHORT_K]; }; // the arrivals kept for one vertex

struct LsState {
	AVec<LsNode> nd;
	AVec<LsEnt> heap;
	AVec<LsVtx> vt;
	int64_t *htab; // vertex << 32 | index into vt, -1 empty
	int32_t htab_bits;
};

MG_HD inline int ls_heap_push(Arena &A, LsState &S, uint64_t di, int32_t node)
{
	LsEnt e;
	e.di = di, e.node = node, e.pad = 0;
	MGB_TRY(avec_push(A, S.heap, e));
	LsEnt *h = S.heap.a;
	int64_t i = S.heap.n - 1;
	while (i > 0) {
		const int64_t par = (i - 1) >> 1;
		if (h[par].di <= di) break;
		h[i] = h[par], i = par;
	}
	h[i] = e;
	return 0;
}
MG_HD inline LsEnt ls_heap_pop(LsState &S)
{
	LsEnt *h = S.heap.a;
	const LsEnt top = h[0], last = h[--S.heap.n];
	const int64_t n = S.heap.n;
	int64_t i = 0;
	for (;;) {
		int64_t c = 2 * i + 1;
		if (c >= n) break;
		if (c + 1 < n && h[c + 1].di < h[c].di) ++c;
		if (last.di <= h[c].di) break;
		h[i] = h[c], i = c;
	}
	if (n > 0) h[i] = last;
	return top;
}

MG_HD inline int ls_vtx_get(Arena &A, LsState &S, uint32_t v, int32_t *idx_)
{
	for (;;) {
		const uint32_t mask = (1u << S.htab_bits) - 1;
		uint32_t h = hash32(v) & mask;
		int64_t e;
		while ((e = S.htab[h]) >= 0 && (uint32_t)(e >> 32) != v) h = (h + 1) & mask;
		if (e >= 0) { *idx_ = (int32_t)e; return 0; }
		if ((uint64_t)(S.vt.n + 1) * 2 > (1ULL << S.htab_bits)) { // keep the load at or below one half
			const int32_t nb = S.htab_bits + 1;
			int64_t *nt;
			MGB_ALLOC(A, nt, int64_t, 1LL << nb);
			const uint32_t nmask = (1u << nb) - 1;
			for (int64_t i = 0; i < (1LL << nb); ++i) nt[i] = -1;
			for (int64_t i = 0; i < S.vt.n; ++i) {
				uint32_t g = hash32(S.vt.a[i].v) & nmask;
				while (nt[g] >= 0) g = (g + 1) & nmask;
				nt[g] = (int64_t)S.vt.a[i].v << 32 | i;
			}
			S.htab = nt, S.htab_bits = nb;
			continue;
		}
		if (S.vt.n == S.vt.m) MGB_TRY(avec_reserve(A, S.vt, S.vt.n + 1));
		LsVtx *q = &S.vt.a[S.vt.n];
		q->v = v, q->k = 0;
		S.htab[h] = (int64_t)v << 32 | S.vt.n;
		*idx_ = (int32_t)S.vt.n++;
		return 0;
	}
}

// in-place heap sort (no scratch, no stack arrays: this runs one search per thread)
MG_HD inline void ls_sort_u64(uint64_t *a, int32_t n)
{
	for (int32_t s = n / 2 - 1, e = n; ; ) {
		uint64_t x;
		if (s >= 0) x = a[s];                       // phase 1: build the max-heap
		else { if (--e <= 0) break; x = a[e], a[e] = a[0]; } // phase 2: move the maximum behind the heap
		int32_t i = s >= 0? s : 0;
		const int32_t m = s >= 0? n : e;
		for (;;) {
			int32_t c = 2 * i + 1;
			if (c >= m) break;
			if (c + 1 < m && a[c + 1] > a[c]) ++c;
			if (a[c] <= x) break;
			a[i] = a[c], i = c;
		}
		a[i] = x;
		if (s >= 0) --s;
	}
}

// Search from src with every push beyond max_dist dropped; the record is built in the arena (at *rec_, *bytes_ long) and is
// valid until the caller releases its mark.  One lane.
MG_HD inline int label_search(Arena &A, const GraphDev &g, uint32_t src, int32_t max_dist, char **rec_, uint64_t *bytes_)
{
	LsState S;
	avec_init(S.nd), avec_init(S.heap), avec_init(S.vt);
	MGB_TRY(avec_reserve(A, S.nd, 512));
	MGB_TRY(avec_reserve(A, S.heap, 512));
	MGB_TRY(avec_reserve(A, S.vt, 48));
	S.htab_bits = 7;
	MGB_ALLOC(A, S.htab, int64_t, 1 << S.htab_bits);
	for (int i = 0; i < (1 << S.htab_bits); ++i) S.htab[i] = -1;
	uint32_t id = 0;
	{
		LsNode p;
		int32_t qi;
		p.v = src, p.di = (uint64_t)id++, p.hash = hash32(src), p.is_0 = 1, p.pad = 0;
		MGB_TRY(avec_push(A, S.nd, p));
		MGB_TRY(ls_heap_push(A, S, p.di, 0));
		MGB_TRY(ls_vtx_get(A, S, src, &qi));
		S.vt.a[qi].k = 1, S.vt.a[qi].node[0] = 0, S.vt.a[qi].d[0] = p.di;
	}
	while (S.heap.n > 0) {
		const LsEnt e = ls_heap_pop(S);
		const LsNode r = S.nd.a[e.node];
		if (r.di != e.di) continue; // this arrival was replaced by a shorter one after it entered the frontier
		const int32_t rdist = (int32_t)(r.di >> 32);
		const int32_t nv = g_arc_n(g, r.v);
		const DevArc *av = g_arc_a(g, r.v);
		for (int32_t i = 0; i < nv; ++i) {
			const DevArc ai = av[i];
			const int32_t d = (int32_t)((uint32_t)rdist + ai.lv);
			if (d > max_dist) continue;
			int32_t qi;
			MGB_TRY(ls_vtx_get(A, S, ai.w, &qi));
			LsVtx *q = &S.vt.a[qi];
			LsNode p;
			p.v = ai.w, p.hash = r.hash + hash32(ai.w), p.is_0 = ai.rank > 0? 0 : r.is_0, p.pad = 0;
			if (q->k < MAX_SHORT_K) {
				p.di = (uint64_t)(uint32_t)d << 32 | id++;
				MGB_TRY(avec_push(A, S.nd, p));
				const int32_t pi = (int32_t)S.nd.n - 1;
				MGB_TRY(ls_heap_push(A, S, p.di, pi));
				q = &S.vt.a[qi];
				q->node[q->k] = pi, q->d[q->k] = p.di, ++q->k;
			} else { // the longest kept arrival gives way to a shorter walk (it cannot have been popped: its distance exceeds d >= rdist)
				int m = 0;
				for (int j = 1; j < MAX_SHORT_K; ++j) if (q->d[j] > q->d[m]) m = j;
				if ((int64_t)(q->d[m] >> 32) > (int64_t)d) {
					const int32_t pi = q->node[m];
					p.di = (uint64_t)(uint32_t)d << 32 | id++;
					S.nd.a[pi] = p;
					MGB_TRY(ls_heap_push(A, S, p.di, pi));
					q->d[m] = p.di;
				}
			}
		}
	}
	// ---- the record: vertices ascending, arrivals of a vertex in pop order (= ascending dist << 32 | id) ----
	const int32_t n_tv = (int32_t)S.vt.n;
	int32_t n_lab = 0;
	for (int32_t i = 0; i < n_tv; ++i) n_lab += S.vt.a[i].k;
	uint64_t *ord;
	MGB_ALLOC(A, ord, uint64_t, n_tv);
	for (int32_t i = 0; i < n_tv; ++i) ord[i] = (uint64_t)S.vt.a[i].v << 32 | (uint64_t)i;
	ls_sort_u64(ord, n_tv); // keys are distinct: any exact sort
	const uint64_t bytes = lab_rec_bytes(n_tv, n_lab);
	char *rec = (char*)arena_alloc(A, bytes);
	if (rec == 0) return MGB_E_ARENA;
	LabRec *h = (LabRec*)rec;
	h->n_tv = n_tv, h->n_lab = n_lab;
	uint32_t *tv = (uint32_t*)(rec + sizeof(LabRec)), *st = tv + n_tv;
	uint64_t *lab = (uint64_t*)lab_arr(rec);
	uint32_t n = 0;
	for (int32_t i = 0; i < n_tv; ++i) {
		LsVtx *q = &S.vt.a[(int32_t)(uint32_t)ord[i]];
		tv[i] = q->v, st[i] = n;
		for (int a = 1; a < q->k; ++a) { // insertion sort of at most 15 keys
			const uint64_t kd = q->d[a];
			const int32_t kn = q->node[a];
			int b = a;
			for (; b > 0 && q->d[b - 1] > kd; --b) q->d[b] = q->d[b - 1], q->node[b] = q->node[b - 1];
			q->d[b] = kd, q->node[b] = kn;
		}
		for (int a = 0; a < q->k; ++a) {
			const LsNode &x = S.nd.a[q->node[a]];
			lab[n++] = (uint64_t)((uint32_t)(x.di >> 32) << 1 | (uint32_t)(x.is_0 != 0)) << 32 | x.hash;
		}
	}
	st[n_tv] = n;
	*rec_ = rec, *bytes_ = bytes;
	return 0;
}

// One item of the label kernel: search a source this batch asked for and publish its record.  One lane.
// A source that cannot be finished here (arena or label pool too small) goes back to LAB_NONE: graph chaining then searches it
// on the spot, and the host grows the pool between batches.
MG_HD inline int label_job(Arena &A, const GraphDev &g, const LabTab &T, int item, int second_pass)
{
	const uint32_t v = (uint32_t)T.new_src[second_pass? T.cap_new - 1 - item : item];
	if (T.src_off[v] >= 0) return 0;
	const uint64_t mark = A.top;
	char *rec;
	uint64_t bytes;
	int rc = label_search(A, g, v, T.max_dist_g + g_vlen(g, v), &rec, &bytes);
	if (rc == MGB_E_ARENA && !second_pass) { // a large neighbourhood: once more with a whole worker arena (k_gc_labels_big)
		unsigned int at;
#if MGB_ON_DEVICE
		at = atomicAdd(&T.n_new[1], 1u);
#else
		at = T.n_new[1]++;
#endif
		T.new_src[T.cap_new - 1 - (int32_t)at] = (int32_t)v; // the two lists together hold at most one entry per vertex
		A.top = mark;
		return 0;
	}
	int64_t off = -1;
	if (rc == 0) off = pool_alloc(T.pool_hdr, bytes);
	if (off >= 0) {
		uint64_t *d = (uint64_t*)(T.pool + off);
		const uint64_t *s = (const uint64_t*)rec;
		for (uint64_t i = 0; i < bytes / 8; ++i) d[i] = s[i];
#if MGB_ON_DEVICE
		__threadfence();
#endif
	}
	T.src_off[v] = off >= 0? (long long)off : (long long)LAB_NONE;
	A.top = mark;
	return 0;
}

} // namespace mgb
