/* mgoracle.c -- plain-C restatement of leaf functions of minigraph's seed-chain path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed from the product
 * (minigraph_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Each function states the reference lines it follows.  The restatement is pinned two ways (tests/test_oracle.py):
 * against the unmodified reference compiled as oracle/_ref/libmgref.so when present, and against the golden vectors
 * committed under tests/golden/ (generated from the reference by tests/golden/make_vectors.py).
 * The complete path (chaining, graph chaining, GWFA, WFA, GAF) is checked against oracle/_ref directly. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* khashl.h:321-329 kh_hash_uint32 */
uint32_t orc_hash32(uint32_t key)
{
	key += ~(key << 15);
	key ^= key >> 10;
	key += key << 3;
	key ^= key >> 6;
	key += ~(key << 11);
	key ^= key >> 16;
	return key;
}

/* khashl.h:342-346 kh_hash_str (X31 hash on signed chars) */
uint32_t orc_hash_str(const char *s)
{
	uint32_t h = (uint32_t)(int32_t)*s;
	if (h == 0) return 0;
	for (++s; *s; ++s) h = h * 31u + (uint32_t)(int32_t)*s;
	return h;
}

/* sketch.c:28-38 hash64: invertible mix restricted to 2k bits */
uint64_t orc_hash64(uint64_t key, uint64_t mask)
{
	key = (~key + (key << 21)) & mask;
	key ^= key >> 24;
	key = (key * 265) & mask;
	key ^= key >> 14;
	key = (key * 21) & mask;
	key ^= key >> 28;
	key = (key * 2147483649ULL) & mask; /* key + (key << 31) */
	return key;
}

/* map-algo.c:362-364: per-read hash from the name hash, the query length and the seed */
uint32_t orc_read_hash(uint32_t name_hash, int32_t qlen, int32_t seed)
{
	uint32_t h = name_hash;
	h ^= orc_hash32((uint32_t)qlen) + orc_hash32((uint32_t)seed);
	return orc_hash32(h);
}

static int base_code(unsigned char c) /* sketch.c:9-26 seq_nt4_table */
{
	switch (c) {
	case 'A': case 'a': case 0: return 0;
	case 'C': case 'c': case 1: return 1;
	case 'G': case 'g': case 2: return 2;
	case 'T': case 't': case 'U': case 'u': case 3: return 3;
	}
	return 4;
}

/* sketch.c:56-109 mg_sketch.  The window is kept as an explicit list of the last w slots (oldest first) instead of a
 * ring buffer, which makes the "rightmost minimum" and "report equal-hash copies in window order" rules explicit.
 * Returns the number of minimizers; at most cap are stored. */
typedef struct { uint64_t x, y; } slot_t;

int64_t orc_sketch(const char *str, int32_t len, int32_t w, int32_t k, uint32_t rid, uint64_t *ox, uint64_t *oy, int64_t cap)
{
	const uint64_t mask = (1ULL << 2 * k) - 1, none = ~0ULL;
	uint64_t fw = 0, rv = 0;
	slot_t win[256], cur_min = { ~0ULL, ~0ULL };
	int32_t n_win = 0;     /* number of slots filled so far, saturates at w */
	int32_t min_age = -1;  /* how many slots ago the current minimum entered (0 = newest); -1 = none yet */
	int32_t run = 0;       /* valid, strand-resolved k-mers since the last ambiguous base */
	int64_t n = 0;
	int32_t i, j;
#define EMIT(s) do { if (n < cap) ox[n] = (s).x, oy[n] = (s).y; ++n; } while (0)
	for (i = 0; i < len; ++i) {
		int c = base_code((unsigned char)str[i]);
		slot_t info = { ~0ULL, ~0ULL };
		if (c < 4) {
			int32_t span = run + 1 < k? run + 1 : k;
			fw = (fw << 2 | (uint64_t)c) & mask;
			rv = rv >> 2 | (uint64_t)(3 - c) << (2 * (k - 1));
			if (fw == rv) continue; /* palindromic k-mer: the slot is not consumed (sketch.c:74) */
			++run;
			if (run >= k && span < 256) {
				int z = fw < rv? 0 : 1;
				info.x = orc_hash64(z? rv : fw, mask) << 8 | (uint64_t)span;
				info.y = (uint64_t)rid << 32 | (uint64_t)((uint32_t)i << 1) | (uint64_t)z;
			}
		} else run = 0;
		/* push the slot; win[0] is the oldest */
		if (n_win == w) { memmove(win, win + 1, (size_t)(w - 1) * sizeof(slot_t)); --n_win; }
		win[n_win++] = info;
		if (min_age >= 0) ++min_age;
		if (run == w + k - 1 && cur_min.x != none) /* first complete window: earlier copies of the minimum (sketch.c:83-88) */
			for (j = 0; j < n_win - 1; ++j)
				if (win[j].x == cur_min.x && win[j].y != cur_min.y) EMIT(win[j]);
		if (info.x <= cur_min.x) { /* new minimum, ties go to the newer k-mer (sketch.c:89-91) */
			if (run >= w + k && cur_min.x != none) EMIT(cur_min);
			cur_min = info, min_age = 0;
		} else if (min_age == w) { /* the minimum has just left the window (sketch.c:92-104) */
			int32_t best = -1;
			if (run >= w + k - 1 && cur_min.x != none) EMIT(cur_min);
			cur_min.x = none;
			for (j = 0; j < n_win; ++j) /* oldest to newest with >=: the newest of equal hashes wins */
				if (cur_min.x >= win[j].x) cur_min = win[j], best = j;
			min_age = n_win - 1 - best;
			if (run >= w + k - 1 && cur_min.x != none)
				for (j = 0; j < n_win; ++j)
					if (win[j].x == cur_min.x && win[j].y != cur_min.y) EMIT(win[j]);
		}
	}
	if (cur_min.x != none) EMIT(cur_min);
#undef EMIT
	return n;
}

/* ksort.h:112-162 KRADIX_SORT_INIT(128x, mg128_t, x, 8): in-place MSD radix sort on 8-bit digits of .x with
 * American-flag cycle permutation (NOT stable), insertion sort for buckets of at most 64 elements. */
typedef struct { uint64_t x, y; } pair_t;

static void ins_sort(pair_t *beg, pair_t *end)
{
	pair_t *i, *j, t;
	for (i = beg + 1; i < end; ++i) {
		if (i->x >= (i - 1)->x) continue;
		t = *i;
		for (j = i; j > beg && t.x < (j - 1)->x; --j) *j = *(j - 1);
		*j = t;
	}
}

static void flag_sort(pair_t *beg, pair_t *end, int shift)
{
	pair_t *head[256], *tail[256];
	size_t cnt[256];
	pair_t *p;
	int b;
	memset(cnt, 0, sizeof(cnt));
	for (p = beg; p != end; ++p) ++cnt[p->x >> shift & 255];
	for (b = 0, p = beg; b < 256; ++b) head[b] = p, p += cnt[b], tail[b] = p;
	for (b = 0; b < 256;) { /* walk the buckets; follow displacement cycles until the element belongs here */
		if (head[b] == tail[b]) { ++b; continue; }
		int d = (int)(head[b]->x >> shift & 255);
		if (d == b) { ++head[b]; continue; }
		pair_t carry = *head[b], tmp;
		do {
			tmp = *head[d]; *head[d]++ = carry; carry = tmp;
			d = (int)(carry.x >> shift & 255);
		} while (d != b);
		*head[b]++ = carry;
	}
	if (shift == 0) return;
	for (b = 0, p = beg; b < 256; ++b) {
		size_t m = cnt[b];
		if (m > 64) flag_sort(p, p + m, shift > 8? shift - 8 : 0);
		else if (m > 1) ins_sort(p, p + m);
		p += m;
	}
}

void orc_radix_sort_128x(uint64_t *x, uint64_t *y, int64_t n)
{
	pair_t *a = (pair_t*)malloc((size_t)(n > 0? n : 1) * sizeof(pair_t));
	int64_t i;
	for (i = 0; i < n; ++i) a[i].x = x[i], a[i].y = y[i];
	if (n <= 64) ins_sort(a, a + n);
	else flag_sort(a, a + n, 56);
	for (i = 0; i < n; ++i) x[i] = a[i].x, y[i] = a[i].y;
	free(a);
}

/* mgpriv.h:63-71 mg_log2 (bit trick, valid for x >= 2) */
float orc_log2(float x)
{
	uint32_t u;
	float f, r;
	memcpy(&u, &x, 4);
	r = (float)(int)((u >> 23 & 255) - 128);
	u = (u & ~(255u << 23)) + (127u << 23);
	memcpy(&f, &u, 4);
	r += (-0.34484843f * f + 2.02466578f) * f - 0.67487759f;
	return r;
}

/* lchain.c:114-139 comput_sc for single-segment reads (n_seg == 1, is_cdna == 0); INT32_MIN = not chainable */
int32_t orc_chain_score(uint64_t xi, uint64_t yi, uint64_t xj, uint64_t yj, int32_t max_dist_x, int32_t max_dist_y, int32_t bw, float pen_gap, float pen_skip)
{
	int32_t dq = (int32_t)yi - (int32_t)yj, dr, dd, dg, span, sc;
	if (dq <= 0 || dq > max_dist_x) return INT32_MIN;
	dr = (int32_t)(xi - xj);
	if (dr == 0 || dq > max_dist_y) return INT32_MIN;
	dd = dr > dq? dr - dq : dq - dr;
	if (dd > bw) return INT32_MIN;
	dg = dr < dq? dr : dq;
	span = (int32_t)(yj >> 32 & 0xff);
	sc = span < dg? span : dg;
	if (dd || dg > span) {
		float lin = pen_gap * (float)dd + pen_skip * (float)dg;
		float lg = dd >= 1? orc_log2((float)(dd + 1)) : 0.0f;
		sc -= (int)(lin + .5f * lg);
	}
	return sc;
}
