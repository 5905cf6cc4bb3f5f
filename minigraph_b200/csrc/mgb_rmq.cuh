// mgb_rmq.cuh -- ordered set with range-minimum queries, used by the RMQ chaining DP.
//
// The reference keeps active anchors in an AVL tree augmented with subtree minima (krmq.h) and
// asks for *a* minimum of `pri` over a closed key interval; when priorities tie, the answer
// depends on the shape of the tree.  To stay bit-exact under ties this is an
// index-based tree that performs the same AVL insert/erase rebalancing steps and the same
// LCA-walk query as krmq.h:110-327, so that it always reaches the same node.
// Nodes live in a flat array inside the worker arena; links are int32 indices (NIL = -1).
#pragma once
#include "mgb_common.cuh"

namespace mgb {

struct RmqNode {
	double pri;
	int32_t y, i;      // ordering key: (y, i)
	int32_t c[2];      // children
	int32_t s;         // index of the node holding the minimum pri in this subtree
	uint32_t size;
	int32_t balance;
};

struct RmqTree {
	RmqNode *nd;       // nd[0] is scratch (the "fake" super-root used while erasing)
	int32_t n_nd, m_nd;
	int32_t root;
};

static const int RMQ_NIL = -1;
static const int RMQ_MAX_DEPTH = 64;

MG_HD inline int rmq_cmp(const RmqNode &a, const RmqNode &b)
{
	return a.y < b.y? -1 : a.y > b.y? 1 : (a.i > b.i) - (a.i < b.i);
}
MG_HD inline int rmq_cmp_key(int32_t y, int32_t i, const RmqNode &b)
{
	return y < b.y? -1 : y > b.y? 1 : (i > b.i) - (i < b.i);
}

MG_HD inline int rmq_init(Arena &A, RmqTree &t, int32_t max_nodes)
{
	MGB_ALLOC(A, t.nd, RmqNode, max_nodes + 1);
	t.n_nd = 1, t.m_nd = max_nodes + 1, t.root = RMQ_NIL;
	return 0;
}

MG_HD inline uint32_t rmq_size(const RmqTree &t) { return t.root == RMQ_NIL? 0 : t.nd[t.root].size; }
MG_HD inline uint32_t rmq_csize(const RmqTree &t, int32_t q, int d) { int32_t c = t.nd[q].c[d]; return c == RMQ_NIL? 0 : t.nd[c].size; }

MG_HD inline void rmq_update_min(RmqTree &t, int32_t p, int32_t q, int32_t r)
{
	RmqNode *nd = t.nd;
	int32_t s = (q == RMQ_NIL || nd[p].pri < nd[nd[q].s].pri)? p : nd[q].s;
	s = (r == RMQ_NIL || nd[s].pri < nd[nd[r].s].pri)? s : nd[r].s;
	nd[p].s = s;
}

// one rotation: (a,(b,c)q)p => ((a,b)p,c)q
MG_HD inline int32_t rmq_rotate1(RmqTree &t, int32_t p, int dir)
{
	RmqNode *nd = t.nd;
	int opp = 1 - dir;
	int32_t q = nd[p].c[opp], s = nd[p].s;
	uint32_t size_p = nd[p].size;
	nd[p].size -= nd[q].size - rmq_csize(t, q, dir);
	nd[q].size = size_p;
	rmq_update_min(t, p, nd[p].c[dir], nd[q].c[dir]);
	nd[q].s = s;
	nd[p].c[opp] = nd[q].c[dir];
	nd[q].c[dir] = p;
	return q;
}

// two rotations: (a,((b,c)r,d)q)p => ((a,b)p,(c,d)q)r
MG_HD inline int32_t rmq_rotate2(RmqTree &t, int32_t p, int dir)
{
	RmqNode *nd = t.nd;
	int opp = 1 - dir, b1;
	int32_t q = nd[p].c[opp], r = nd[q].c[dir], s = nd[p].s;
	uint32_t size_x_dir = rmq_csize(t, r, dir);
	nd[r].size = nd[p].size;
	nd[p].size -= nd[q].size - size_x_dir;
	nd[q].size -= size_x_dir + 1;
	rmq_update_min(t, p, nd[p].c[dir], nd[r].c[dir]);
	rmq_update_min(t, q, nd[q].c[opp], nd[r].c[opp]);
	nd[r].s = s;
	nd[p].c[opp] = nd[r].c[dir];
	nd[r].c[dir] = p;
	nd[q].c[dir] = nd[r].c[opp];
	nd[r].c[opp] = q;
	b1 = dir == 0? +1 : -1;
	if (nd[r].balance == b1) nd[q].balance = 0, nd[p].balance = -b1;
	else if (nd[r].balance == 0) nd[q].balance = nd[p].balance = 0;
	else nd[q].balance = b1, nd[p].balance = 0;
	nd[r].balance = 0;
	return r;
}

// insert a new element; returns its node index (keys are unique in the chaining DP), <0 on failure
MG_HD inline int32_t rmq_insert(RmqTree &t, int32_t y, int32_t i, double pri)
{
	RmqNode *nd = t.nd;
	unsigned char stack[RMQ_MAX_DEPTH];
	int32_t path[RMQ_MAX_DEPTH];
	int32_t bp, bq, p, q, r = RMQ_NIL, x;
	int which = 0, top, b1, path_len, k;
	if (t.n_nd >= t.m_nd) return -2;
	bp = t.root, bq = RMQ_NIL;
	for (p = bp, q = bq, top = path_len = 0; p != RMQ_NIL; q = p, p = nd[p].c[which]) {
		int cmp = rmq_cmp_key(y, i, nd[p]);
		if (cmp == 0) return p;
		if (nd[p].balance != 0) bq = q, bp = p, top = 0;
		stack[top++] = (unsigned char)(which = (cmp > 0));
		path[path_len++] = p;
		if (path_len >= RMQ_MAX_DEPTH - 1) return -3;
	}
	x = t.n_nd++;
	nd[x].y = y, nd[x].i = i, nd[x].pri = pri;
	nd[x].balance = 0, nd[x].size = 1, nd[x].c[0] = nd[x].c[1] = RMQ_NIL, nd[x].s = x;
	if (q == RMQ_NIL) t.root = x;
	else nd[q].c[which] = x;
	if (bp == RMQ_NIL) return x;
	for (k = 0; k < path_len; ++k) ++nd[path[k]].size;
	for (k = path_len - 1; k >= 0; --k) {
		rmq_update_min(t, path[k], nd[path[k]].c[0], nd[path[k]].c[1]);
		if (nd[path[k]].s != x) break;
	}
	for (p = bp, top = 0; p != x; p = nd[p].c[stack[top]], ++top) {
		if (stack[top] == 0) --nd[p].balance;
		else ++nd[p].balance;
	}
	if (nd[bp].balance > -2 && nd[bp].balance < 2) return x;
	which = (nd[bp].balance < 0);
	b1 = which == 0? +1 : -1;
	q = nd[bp].c[1 - which];
	if (nd[q].balance == b1) {
		r = rmq_rotate1(t, bp, which);
		nd[q].balance = nd[bp].balance = 0;
	} else r = rmq_rotate2(t, bp, which);
	if (bq == RMQ_NIL) t.root = r;
	else nd[bq].c[bp != nd[bq].c[0]] = r;
	return x;
}

MG_HD inline int32_t rmq_find(const RmqTree &t, int32_t y, int32_t i)
{
	int32_t p = t.root;
	while (p != RMQ_NIL) {
		int cmp = rmq_cmp_key(y, i, t.nd[p]);
		if (cmp < 0) p = t.nd[p].c[0];
		else if (cmp > 0) p = t.nd[p].c[1];
		else break;
	}
	return p;
}

// erase the element with key (y,i); returns 1 if it was present
MG_HD inline int rmq_erase(RmqTree &t, int32_t y, int32_t i)
{
	RmqNode *nd = t.nd;
	int32_t p, path[RMQ_MAX_DEPTH];
	unsigned char dir[RMQ_MAX_DEPTH];
	int k, d = 0, cmp;
	const int32_t fake = 0;
	if (t.root == RMQ_NIL) return 0;
	nd[fake] = nd[t.root], nd[fake].c[0] = t.root, nd[fake].c[1] = RMQ_NIL;
	for (cmp = -1, p = fake; cmp; cmp = rmq_cmp_key(y, i, nd[p])) {
		int which = (cmp > 0);
		dir[d] = (unsigned char)which;
		path[d++] = p;
		p = nd[p].c[which];
		if (p == RMQ_NIL) return 0;
		if (d >= RMQ_MAX_DEPTH - 2) return 0;
	}
	for (k = 1; k < d; ++k) --nd[path[k]].size;
	if (nd[p].c[1] == RMQ_NIL) {
		nd[path[d-1]].c[dir[d-1]] = nd[p].c[0];
	} else {
		int32_t q = nd[p].c[1];
		if (nd[q].c[0] == RMQ_NIL) {
			nd[q].c[0] = nd[p].c[0];
			nd[q].balance = nd[p].balance;
			nd[path[d-1]].c[dir[d-1]] = q;
			path[d] = q, dir[d++] = 1;
			nd[q].size = nd[p].size - 1;
		} else {
			int32_t r;
			int e = d++;
			for (;;) {
				dir[d] = 0;
				path[d++] = q;
				r = nd[q].c[0];
				if (nd[r].c[0] == RMQ_NIL) break;
				q = r;
			}
			nd[r].c[0] = nd[p].c[0];
			nd[q].c[0] = nd[r].c[1];
			nd[r].c[1] = nd[p].c[1];
			nd[r].balance = nd[p].balance;
			nd[path[e-1]].c[dir[e-1]] = r;
			path[e] = r, dir[e] = 1;
			for (k = e + 1; k < d; ++k) --nd[path[k]].size;
			nd[r].size = nd[p].size - 1;
		}
	}
	for (k = d - 1; k >= 0; --k)
		rmq_update_min(t, path[k], nd[path[k]].c[0], nd[path[k]].c[1]);
	while (--d > 0) {
		int32_t q = path[d];
		int which, other, b1 = 1, b2 = 2;
		which = dir[d], other = 1 - which;
		if (which) b1 = -b1, b2 = -b2;
		nd[q].balance += b1;
		if (nd[q].balance == b1) break;
		else if (nd[q].balance == b2) {
			int32_t r = nd[q].c[other];
			if (nd[r].balance == -b1) {
				nd[path[d-1]].c[dir[d-1]] = rmq_rotate2(t, q, which);
			} else {
				nd[path[d-1]].c[dir[d-1]] = rmq_rotate1(t, q, which);
				if (nd[r].balance == 0) {
					nd[r].balance = -b1;
					nd[q].balance = b1;
					break;
				} else nd[r].balance = nd[q].balance = 0;
			}
		}
	}
	t.root = nd[fake].c[0];
	return 1;
}

// a minimum-pri node with key in the CLOSED interval [(lo_y,lo_i), (up_y,up_i)], or NIL
MG_HD inline int32_t rmq_query(const RmqTree &t, int32_t lo_y, int32_t lo_i, int32_t up_y, int32_t up_i)
{
	const RmqNode *nd = t.nd;
	int32_t p, path[2][RMQ_MAX_DEPTH], mn;
	int plen[2] = {0, 0}, pcmp[2][RMQ_MAX_DEPTH], k, cmp, lca;
	if (t.root == RMQ_NIL) return RMQ_NIL;
	p = t.root;
	while (p != RMQ_NIL) {
		cmp = rmq_cmp_key(lo_y, lo_i, nd[p]);
		path[0][plen[0]] = p, pcmp[0][plen[0]++] = cmp;
		if (cmp < 0) p = nd[p].c[0];
		else if (cmp > 0) p = nd[p].c[1];
		else break;
	}
	p = t.root;
	while (p != RMQ_NIL) {
		cmp = rmq_cmp_key(up_y, up_i, nd[p]);
		path[1][plen[1]] = p, pcmp[1][plen[1]++] = cmp;
		if (cmp < 0) p = nd[p].c[0];
		else if (cmp > 0) p = nd[p].c[1];
		else break;
	}
	for (k = 0; k < plen[0] && k < plen[1]; ++k)
		if (path[0][k] == path[1][k] && pcmp[0][k] <= 0 && pcmp[1][k] >= 0) break;
	if (k == plen[0] || k == plen[1]) return RMQ_NIL;
	lca = k, mn = path[0][lca];
	for (k = lca + 1; k < plen[0]; ++k) {
		if (pcmp[0][k] <= 0) {
			int32_t c = nd[path[0][k]].c[1];
			if (nd[path[0][k]].pri < nd[mn].pri) mn = path[0][k];
			if (c != RMQ_NIL && nd[nd[c].s].pri < nd[mn].pri) mn = nd[c].s;
		}
	}
	for (k = lca + 1; k < plen[1]; ++k) {
		if (pcmp[1][k] >= 0) {
			int32_t c = nd[path[1][k]].c[0];
			if (nd[path[1][k]].pri < nd[mn].pri) mn = path[1][k];
			if (c != RMQ_NIL && nd[nd[c].s].pri < nd[mn].pri) mn = nd[c].s;
		}
	}
	return mn;
}

// largest element <= key (reference: krmq_interval(), lower bound only)
MG_HD inline int32_t rmq_lower(const RmqTree &t, int32_t y, int32_t i)
{
	int32_t p = t.root, l = RMQ_NIL;
	while (p != RMQ_NIL) {
		int cmp = rmq_cmp_key(y, i, t.nd[p]);
		if (cmp < 0) p = t.nd[p].c[0];
		else if (cmp > 0) l = p, p = t.nd[p].c[1];
		else { l = p; break; }
	}
	return l;
}

// in-order iterator walking towards smaller keys (reference: krmq_itr_find + krmq_itr_prev)
struct RmqItr {
	int32_t stack[RMQ_MAX_DEPTH];
	int top; // index of the top entry, -1 when exhausted
};

MG_HD inline void rmq_itr_find(const RmqTree &t, int32_t x, RmqItr &it)
{
	int32_t p = t.root;
	const RmqNode &key = t.nd[x];
	it.top = -1;
	while (p != RMQ_NIL) {
		it.stack[++it.top] = p;
		int cmp = rmq_cmp(key, t.nd[p]);
		if (cmp < 0) p = t.nd[p].c[0];
		else if (cmp > 0) p = t.nd[p].c[1];
		else break;
	}
}

MG_HD inline int32_t rmq_itr_at(const RmqItr &it) { return it.top < 0? RMQ_NIL : it.stack[it.top]; }

MG_HD inline int rmq_itr_prev(const RmqTree &t, RmqItr &it)
{
	int32_t p;
	if (it.top < 0) return 0;
	p = t.nd[it.stack[it.top]].c[0];
	if (p != RMQ_NIL) {
		for (; p != RMQ_NIL; p = t.nd[p].c[1]) it.stack[++it.top] = p;
		return 1;
	} else {
		do {
			p = it.stack[it.top--];
		} while (it.top >= 0 && p == t.nd[it.stack[it.top]].c[0]);
		return it.top < 0? 0 : 1;
	}
}

} // namespace mgb
