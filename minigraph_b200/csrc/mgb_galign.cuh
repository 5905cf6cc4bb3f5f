// mgb_galign.cuh -- base-level alignment of graph chains and the per-read result blob.
//   gchain_cigar()  (reference: galign.c:39-145 mg_gchain_cigar)
//   gchain_ds()     (reference: galign.c:182-293 mg_gchain_gen_ds)
//   stage_align()   K6-K8 for one read (reference: map-algo.c:455-479)
#pragma once
#include "mgb_pipeline.cuh"
#include "mgb_gchain.cuh"
#include "mgb_wfa_tiers.cuh"
#if defined(MGB_HOSTSIM)
#include <stdio.h>
#include <stdlib.h>
#endif

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

// One gap between two kept anchors that needs a real alignment.  Produced by the planning pass of K6/K7 (one lane
// per read), consumed by the WFA kernel K8a (one warp per job), stitched back by the finishing pass.
struct WfaJob {
	int32_t rid, gc;
	int32_t l0, l;        // first and last llchain of the gap (indices into the read's LLChain array)
	int32_t t_beg;        // first target base on lc[l0].v  (= q->x + 1)
	int32_t t_last;       // last target base on lc[l].v    (= p->x)
	int32_t tl, ql, q_off;
	int32_t n_cigar, status;
	int64_t lc_off;       // byte offset of the read's LLChain array inside the output pool
	int64_t cig_off;      // byte offset of the job's CIGAR (uint32 len<<4|op) inside the cigar pool
};

static const uint64_t PLAN_JOB = 1ULL << 63;
// The state of a warp's slice {next free byte, end} sits in the 96 bytes between the end of the tier's shared-memory layout and its
// stride (a static array would cost k_wfa_mid its seventh block per SM); tier 3 has no shared memory and few jobs: it asks per gap.
MG_HD inline unsigned long long *wfa_cig_chunk(int32_t *smem, int tier)
{
	static_assert(WfTier1::BYTES % 8 == 0 && WfTier1::BYTES + 16 <= WfTier1::STRIDE && WfTier2::BYTES % 8 == 0 && WfTier2::BYTES + 16 <= WfTier2::STRIDE, "no room for the slice state behind the tier layouts");
	return smem == 0 || tier > 2? (unsigned long long*)0 : (unsigned long long*)((char*)smem + (tier == 1? WfTier1::BYTES : WfTier2::BYTES));
}
static const unsigned long long CIG_CHUNK_BYTES = 2048; // a warp of a WFA kernel takes the CIGAR pool in slices of this size (the host sizes the pool with one slice per worker and kernel to spare)

// Planning pass of mg_gchain_cigar() (reference: galign.c:39-124): walk the kept anchors of every graph chain, emit
// literal CIGAR items for the trivial gaps (galign.c:98-100) and a WfaJob for the others.

// K8a: align one gap.  Warp-uniform (all lanes enter with identical arguments).
// tier 1: small gaps, wavefronts + traceback bytes in shared memory; tier 2: mid-size gaps, wavefronts in shared
// memory; tier 3: anything, wavefronts in the worker arena.  A job that does not fit a tier is appended to the queue
// of the next one (jobq[tier-1]); the host launches the next tier over that queue.
MG_HD inline int wfa_job_run(Arena &A, const PipeCtx &c, int64_t job_idx, int lane, int32_t *smem, int tier)
{
	WfaJob *J = &c.jobs[job_idx];
	if (J->rid < 0) return 0; // a slot no read wrote (its allocation ran over the end of the pool; the batch is re-run with a larger one)
	const int32_t rid = J->rid, l0 = J->l0, l = J->l, tl = J->tl, ql = J->ql;
	if (c.meta[rid].status < 0) return 0; // the read already failed elsewhere; it will be redone as a whole
	// Routing: a gap that cannot finish in an on-chip tier costs that tier up to a full window of cells before it gives up.
	// Which lengths fail depends on the error rate of the reads, so it is learned: one gap in 64 tries every tier and
	// reports where it finished; the host turns the counts of one batch into the two thresholds of the next.  The
	// result of a gap does not depend on the tier that computes it.
	const int32_t mlen = tl > ql? tl : ql;
	const int explore = (job_idx & 63) == 0;
	if (!explore && ((tier == 1 && mlen >= c.skip1_len) || (tier == 2 && mlen >= c.skip2_len))) {
		if (lane == 0) {
			unsigned int at;
#if MGB_ON_DEVICE
			at = atomicAdd(&c.jobq_n[tier - 1], 1u);
#else
			at = c.jobq_n[tier - 1]++;
#endif
			c.jobq[tier - 1][at] = (int32_t)job_idx;
		}
		return 0;
	}
	if ((tier == 1 && (tl > WfTier1::MAXLEN_ || ql > WfTier1::MAXLEN_)) || (tier == 2 && (tl > WfTier2::MAXLEN_ || ql > WfTier2::MAXLEN_))) {
		if (lane == 0) {
			unsigned int at;
#if MGB_ON_DEVICE
			at = atomicAdd(&c.jobq_n[tier - 1], 1u);
#else
			at = c.jobq_n[tier - 1]++;
#endif
			c.jobq[tier - 1][at] = (int32_t)job_idx;
		}
		return 0;
	}
	const GraphDev &g = c.g;
	const LLChain *lc = (const LLChain*)(c.out + J->lc_off);
	const char *qs = c.b.seq + c.b.seq_off[rid] + J->q_off;
	const char *tseq;
	if (l == l0) tseq = g_vseq(g, lc[l0].v) + J->t_beg;
	else { // stitch the target across the walk (reference: galign.c:76-93)
		char *seq;
		MGB_ALLOC(A, seq, char, tl + 1);
		int32_t n = g_vlen(g, lc[l0].v) - J->t_beg, at = 0;
		const char *s = g_vseq(g, lc[l0].v) + J->t_beg;
		for (int32_t x = lane; x < n; x += MGB_W) seq[at + x] = s[x];
		at += n;
		for (int32_t k = l0 + 1; k < l; ++k) {
			s = g_vseq(g, lc[k].v), n = g_vlen(g, lc[k].v);
			for (int32_t x = lane; x < n; x += MGB_W) seq[at + x] = s[x];
			at += n;
		}
		s = g_vseq(g, lc[l].v), n = J->t_last + 1;
		for (int32_t x = lane; x < n; x += MGB_W) seq[at + x] = s[x];
		at += n;
		if (at != tl) return MGB_E_INTERNAL;
		warp_sync();
		tseq = seq;
	}
	WfResult rst;
	unsigned long long pt0 = prof_clock();
	int rc;
	if (tier == 1) rc = wfa_smem<WfTier1::W_, WfTier1::MAXLEN_, WfTier1::TBCAP_>(A, smem, tl, tseq, ql, qs, &rst, lane);
	else if (tier == 2) rc = wfa_smem<WfTier2::W_, WfTier2::MAXLEN_, WfTier2::TBCAP_>(A, smem, tl, tseq, ql, qs, &rst, lane);
	else rc = wfa_exact(A, tl, tseq, ql, qs, 100000000LL, &rst, lane);
	if (rc < 0) return rc;
	if (lane == 0) {
		unsigned long long dt = prof_clock() - pt0;
		int slot = tier == 1? PROF_WFA_FAST_CYC : tier == 2? PROF_WFA_MID_CYC : PROF_WFA_SLOW_CYC;
		prof_add(c, slot, dt), prof_add(c, slot + 1, 1);
		prof_max(c, PROF_WFA_MAX_CYC, dt << 16 | (unsigned long long)(mlen < 65535? mlen : 65535)); // cycles of the slowest gap, its length in the low 16 bits
		if (rc == 0) prof_add(c, PROF_WFA_CELLS, (unsigned long long)rst.n_iter);
		if (rc == 0 && explore && c.tier_hist) {
			unsigned int *h = &c.tier_hist[(mlen >> 4 < 31? mlen >> 4 : 31) * 4 + tier];
#if MGB_ON_DEVICE
			atomicAdd(h, 1u);
#else
			++*h;
#endif
		}
	}
	if (rc == 1) { // does not fit this tier
		if (lane == 0) {
			unsigned int at;
#if MGB_ON_DEVICE
			at = atomicAdd(&c.jobq_n[tier - 1], 1u);
#else
			at = c.jobq_n[tier - 1]++;
#endif
			c.jobq[tier - 1][at] = (int32_t)job_idx;
		}
		return 0;
	}
#if !MGB_ON_DEVICE && defined(MGB_HOSTSIM)
	if (getenv("MGB_DUMP_JOBS")) fprintf(stderr, "JOB\t%d\t%d\t%d\t%ld\t%d\t%d\n", tl, ql, rst.s, (long)rst.n_iter, rst.n_cigar, tier);
#endif
	if (rst.s < 0) return MGB_E_INTERNAL;
	int64_t coff = 0;
	if (lane == 0) {
#if MGB_ON_DEVICE
		unsigned long long *ck = wfa_cig_chunk(smem, tier); // the warp's slice of the pool
		const unsigned long long need = ((unsigned long long)rst.n_cigar * 4 + 15) & ~15ULL;
		if (ck == 0) coff = pool_alloc(c.pool_cig, (uint64_t)rst.n_cigar * 4);
		else if (ck[0] + need > ck[1]) {
			const unsigned long long get = need > CIG_CHUNK_BYTES? need : CIG_CHUNK_BYTES;
			const int64_t off = pool_alloc(c.pool_cig, get);
			if (off < 0) ck[0] = ck[1] = 0;
			else ck[0] = (unsigned long long)off, ck[1] = (unsigned long long)off + get;
			coff = off;
		}
		if (ck && coff >= 0) coff = (int64_t)ck[0], ck[0] += need;
#else
		coff = pool_alloc(c.pool_cig, (uint64_t)rst.n_cigar * 4);
#endif
	}
	coff = (int64_t)warp_bcast_u64((uint64_t)coff, 0);
	if (coff < 0) return MGB_E_POOL;
	uint32_t *dst = (uint32_t*)((char*)c.cig + coff);
	for (int32_t x = lane; x < rst.n_cigar; x += MGB_W) dst[x] = rst.cigar[x];
	if (lane == 0) J->n_cigar = rst.n_cigar, J->cig_off = coff;
	return 0;
}

// Finishing pass of mg_gchain_cigar() (reference: galign.c:125-141): concatenate plan items and job CIGARs.

#if MGB_ON_DEVICE
MG_D inline void lane_atomic_add_u64(uint64_t *p, uint64_t v) { atomicAdd((unsigned long long*)p, (unsigned long long)v); }
#else
inline void lane_atomic_add_u64(uint64_t *p, uint64_t v) { *p += v; } // lanes of the simulators never run concurrently
#endif

// gchain_cigar_finish() entered by all lanes of a warp (stage_finish<1>, parameter "fin_v2").  The reference appends item after
// item and merges the first operation of an item into the last one written when they are of the same kind (galign.c:125-141 via
// mg_cigar_push); nothing else is ever merged.  Here: (1) a prefix sum over the items gives every item its place in a flat list,
// (2) the lanes copy the items there, the first operation of each item tagged, (3) a tagged operation that equals its left
// neighbour continues that neighbour's run, every other operation starts a run; the runs are numbered by a ballot prefix
// count and their lengths added up.  Same list as the sequential append, operation for operation.
MG_HD inline int gchain_cigar_finish_w(Arena &A, const PipeCtx &c, GcSet &gt, CigarOut *out, int lane)
{
	const uint64_t FIRST = 1ULL << 62; // tag: first operation of an item (lengths stay far below 2^58)
	for (int32_t i = 0; i < gt.n_gc; ++i) {
		GChain *gc = &gt.gc[i];
		const int32_t off_a0 = gt.lc[gc->off].off, n_plan = gc->n_plan;
		const uint64_t *plan = c.plan + gc->plan_off;
		int32_t *ioff;
		MGB_ALLOC(A, ioff, int32_t, n_plan + 1);
		int32_t tot = 0;
		for (int32_t base = 0; base < n_plan; base += MGB_W) {
			const int32_t t = base + lane;
			int32_t cnt = 0;
			if (t < n_plan) cnt = (plan[t] & PLAN_JOB)? c.jobs[plan[t] & ~PLAN_JOB].n_cigar : 1;
			const int32_t incl = warp_incl_scan_i32(cnt, lane);
			if (t < n_plan) ioff[t] = tot + incl - cnt;
			tot += warp_bcast_i32(incl, MGB_W - 1);
		}
		uint64_t *F, *O;
		int32_t *oidx;
		MGB_ALLOC(A, F, uint64_t, tot + 1);
		MGB_ALLOC(A, O, uint64_t, tot + 1);
		MGB_ALLOC(A, oidx, int32_t, tot + 1);
		warp_sync();
		for (int32_t t = lane; t < n_plan; t += MGB_W) {
			uint64_t *dst = F + ioff[t];
			if (plan[t] & PLAN_JOB) {
				const WfaJob *J = &c.jobs[plan[t] & ~PLAN_JOB];
				const uint32_t *cig = (const uint32_t*)((const char*)c.cig + J->cig_off);
				for (int32_t k = 0; k < J->n_cigar; ++k) dst[k] = (uint64_t)cig[k] | (k == 0? FIRST : 0);
			} else dst[0] = plan[t] | FIRST;
		}
		warp_sync();
		int32_t n_out = 0;
		for (int32_t base = 0; base < tot; base += MGB_W) {
			const int32_t j = base + lane;
			int head = 0;
			if (j < tot) head = j == 0 || !((F[j] & FIRST) && (F[j] & 0xf) == (F[j - 1] & 0xf));
			const uint32_t mh = warp_ballot(head);
			if (j < tot) {
				const int32_t o = n_out + mask_rank(mh, lane) + head - 1;
				oidx[j] = o;
				if (head) O[o] = F[j] & 0xf;
			}
			n_out += mask_count(mh);
		}
		warp_sync();
		for (int32_t j = lane; j < tot; j += MGB_W) lane_atomic_add_u64(&O[oidx[j]], (F[j] & ~FIRST) >> 4 << 4);
		warp_sync();
		int32_t mlen = 0, blen = 0, aplen = 0, l = 0;
		for (int32_t j = lane; j < n_out; j += MGB_W) {
			const int32_t op = (int32_t)(O[j] & 0xf), len = (int32_t)(O[j] >> 4);
			if (op == 7) mlen += len;
			blen += len;
			if (op != 1) aplen += len;
			if (op != 2) l += len;
		}
		mlen = warp_sum_i32(mlen), blen = warp_sum_i32(blen), aplen = warp_sum_i32(aplen), l = warp_sum_i32(l);
		if (lane == 0) {
			out[i].cigar = O, out[i].n = n_out;
			gc->has_cigar = 1;
			gc->n_cigar = n_out;
			gc->c_ss = (int32_t)gt.a[off_a0].x + 1 - (int32_t)(gt.a[off_a0].y >> 32 & 0xff);
			gc->c_ee = (int32_t)gt.a[off_a0 + gc->n_anchor - 1].x + 1;
			gc->c_mlen = mlen, gc->c_blen = blen, gc->c_aplen = aplen;
		}
		if (!(l == gc->qe - gc->qs && aplen == gc->pe - gc->ps)) return MGB_E_INTERNAL;
		warp_sync();
	}
	return 0;
}

// ---- ds:Z difference string ----

struct DsOut { char *ds; int32_t len; int32_t *off; int32_t n_off; };

MG_HD inline char ds_nt(const char *s, int64_t i) { return "acgtn"[nt4((uint8_t)s[i])]; }

// ---- raw writers for the chunked ds generation: the caller has sized the buffers from per-operation upper bounds ----
MG_HD inline char *ds_w_int(char *w, int32_t c)
{
	char buf[16];
	int l = 0;
	uint32_t x = c >= 0? (uint32_t)c : (uint32_t)(-c);
	do { buf[l++] = (char)(x % 10 + '0'); x /= 10; } while (x > 0);
	if (c < 0) buf[l++] = '-';
	for (int i = l - 1; i >= 0; --i) *w++ = buf[i];
	return w;
}

// reference: galign.c:153-180 write_indel
MG_HD inline char *ds_w_indel(char *w, int64_t len, const char *seq, int64_t ll, int64_t lr)
{
	int64_t i;
	if (ll + lr >= len) {
		*w++ = '[';
		for (i = 0; i < len; ++i) *w++ = ds_nt(seq, i);
		*w++ = ']';
	} else {
		int64_t k = 0;
		if (ll > 0) {
			*w++ = '[';
			for (i = 0; i < ll; ++i) *w++ = ds_nt(seq, k + i);
			*w++ = ']';
			k += ll;
		}
		for (i = 0; i < len - lr - ll; ++i) *w++ = ds_nt(seq, k + i);
		k += len - lr - ll;
		if (lr > 0) {
			*w++ = '[';
			for (i = 0; i < lr; ++i) *w++ = ds_nt(seq, k + i);
			*w++ = ']';
		}
	}
	return w;
}

static const int DS_CHUNKS = 32;
struct DsChunk { int64_t dx, dy; int32_t cap_b, cap_o, n_b, n_o; };

// The ds:Z string of every chain (reference: galign.c:182-262 mg_gchain_gen_ds).  What one CIGAR operation contributes
// depends only on the operation and on where it starts on the walk and on the read, so the operations are cut into
// DS_CHUNKS runs that are written independently (by different lanes on the device) and then joined in order.
// Warp-uniform: all lanes enter; out[] is filled identically on every lane.
MG_HD inline int gchain_ds_w(Arena &A, const GraphDev &g, const char *qseq, GcSet &gt, const CigarOut *cg, DsOut *out, int lane)
{
	for (int32_t i = 0; i < gt.n_gc; ++i) {
		GChain *gc = &gt.gc[i];
		const uint64_t *cigar = cg[i].cigar;
		const int32_t n_cigar = gc->n_cigar, aplen = gc->c_aplen;
		char *seq;
		int64_t seq_l = 0;
		MGB_ALLOC(A, seq, char, aplen + 1);
		for (int32_t j = 0; j < gc->cnt; ++j) { // the aligned part of the walk
			const int32_t k = gc->off + j;
			const uint32_t v = gt.lc[k].v;
			const int32_t slen = g_vlen(g, v);
			const int32_t st = j > 0? 0 : gc->c_ss;
			const int32_t en = j < gc->cnt - 1? slen : gc->c_ee;
			if (seq_l + (en - st) > aplen) return MGB_E_INTERNAL;
			const char *s = g_vseq(g, v) + st;
			for (int32_t t = lane; t < en - st; t += MGB_W) seq[seq_l + t] = s[t];
			seq_l += en - st;
		}
		if (seq_l != aplen) return MGB_E_INTERNAL;
		DsChunk *ch;
		MGB_ALLOC(A, ch, DsChunk, DS_CHUNKS);
		const int32_t per = (n_cigar + DS_CHUNKS - 1) / DS_CHUNKS;
		// pass 0: how far each run advances, and how much it can write at most
		for (int c = lane; c < DS_CHUNKS; c += MGB_W) {
			const int32_t j0 = c * per < n_cigar? c * per : n_cigar, j1 = j0 + per < n_cigar? j0 + per : n_cigar;
			DsChunk d;
			d.dx = d.dy = 0, d.cap_b = d.cap_o = d.n_b = d.n_o = 0;
			for (int32_t j = j0; j < j1; ++j) {
				const int64_t op = (int64_t)(cigar[j] & 0xf), len = (int64_t)(cigar[j] >> 4);
				if (op == 7) d.dx += len, d.dy += len, d.cap_b += 12, d.cap_o += 1;
				else if (op == 0 || op == 8) d.dx += len, d.dy += len, d.cap_b += (int32_t)(14 * len + 12), d.cap_o += (int32_t)(2 * len + 1);
				else if (op == 1) d.dy += len, d.cap_b += (int32_t)(len + 6), d.cap_o += 1;
				else if (op == 2) d.dx += len, d.cap_b += (int32_t)(len + 6), d.cap_o += 1;
			}
			ch[c] = d;
		}
		warp_sync();
		int64_t tot_cb = 0, tot_co = 0;
		for (int c = 0; c < DS_CHUNKS; ++c) tot_cb += ch[c].cap_b, tot_co += ch[c].cap_o;
		char *tmp_b;
		int32_t *tmp_o;
		MGB_ALLOC(A, tmp_b, char, tot_cb);
		MGB_ALLOC(A, tmp_o, int32_t, tot_co);
		// pass 1: every run writes its text and its (run-relative) offsets
		int bad = 0;
		for (int c = lane; c < DS_CHUNKS; c += MGB_W) {
			const int32_t j0 = c * per < n_cigar? c * per : n_cigar, j1 = j0 + per < n_cigar? j0 + per : n_cigar;
			int64_t x = 0, y = gc->qs, cb = 0, co = 0;
			for (int q = 0; q < c; ++q) x += ch[q].dx, y += ch[q].dy, cb += ch[q].cap_b, co += ch[q].cap_o;
			char *const w0 = tmp_b + cb;
			char *w = w0;
			int32_t *const o0 = tmp_o + co;
			int32_t *o = o0;
			for (int32_t j = j0; j < j1; ++j) {
				const int64_t op = (int64_t)(cigar[j] & 0xf), len = (int64_t)(cigar[j] >> 4);
				if (op == 0 || op == 7 || op == 8) {
					int64_t z;
					int32_t l = 0;
					if (op == 7) l = (int32_t)len, z = len; // '=' runs hold identical characters: nothing to look at
					else z = 0;
					for (; z < len; ++z) {
						const uint8_t cx = (uint8_t)nt4((uint8_t)seq[x + z]);
						const uint8_t cy = (uint8_t)nt4((uint8_t)qseq[y + z]);
						if (cx != cy) {
							if (l > 0) { *o++ = (int32_t)(w - w0); *w++ = ':'; w = ds_w_int(w, l); }
							*o++ = (int32_t)(w - w0);
							*w++ = '*', *w++ = "acgtn"[cx], *w++ = "acgtn"[cy];
							l = 0;
						} else ++l;
					}
					if (l > 0) { *o++ = (int32_t)(w - w0); *w++ = ':'; w = ds_w_int(w, l); }
					x += len, y += len;
				} else if (op == 1) {
					int64_t z, ll, lr;
					for (z = 1; z <= len; ++z)
						if (y - z < gc->qs || qseq[y + len - z] != qseq[y - z]) break;
					lr = z - 1;
					for (z = 0; z < len; ++z)
						if (y + len + z >= gc->qe || qseq[y + len + z] != qseq[y + z]) break;
					ll = z;
					*o++ = (int32_t)(w - w0);
					*w++ = '+';
					w = ds_w_indel(w, len, &qseq[y], ll, lr);
					y += len;
				} else if (op == 2) {
					int64_t z, ll, lr;
					for (z = 1; z <= len; ++z)
						if (x - z < 0 || seq[x + len - z] != seq[x - z]) break;
					lr = z - 1;
					for (z = 0; z < len; ++z)
						if (x + len + z >= aplen || seq[x + z] != seq[x + len + z]) break;
					ll = z;
					*o++ = (int32_t)(w - w0);
					*w++ = '-';
					w = ds_w_indel(w, len, &seq[x], ll, lr);
					x += len;
				}
			}
			ch[c].n_b = (int32_t)(w - w0), ch[c].n_o = (int32_t)(o - o0);
			if (ch[c].n_b > ch[c].cap_b || ch[c].n_o > ch[c].cap_o) bad = 1; // cannot happen: the bounds are per operation
		}
		if (warp_any(bad)) return MGB_E_INTERNAL;
		warp_sync();
		// join
		int64_t tot_b = 0, tot_o = 0;
		for (int c = 0; c < DS_CHUNKS; ++c) tot_b += ch[c].n_b, tot_o += ch[c].n_o;
		char *str;
		int32_t *off;
		MGB_ALLOC(A, str, char, tot_b + 1);
		MGB_ALLOC(A, off, int32_t, tot_o);
		{
			int64_t cb = 0, co = 0, ab = 0, ao = 0;
			for (int c = 0; c < DS_CHUNKS; ++c) {
				const char *sb = tmp_b + cb;
				const int32_t *so = tmp_o + co;
				for (int32_t t = lane; t < ch[c].n_b; t += MGB_W) str[ab + t] = sb[t];
				for (int32_t t = lane; t < ch[c].n_o; t += MGB_W) off[ao + t] = so[t] + (int32_t)ab;
				cb += ch[c].cap_b, co += ch[c].cap_o, ab += ch[c].n_b, ao += ch[c].n_o;
			}
		}
		warp_sync();
		out[i].ds = str, out[i].len = (int32_t)tot_b, out[i].off = off, out[i].n_off = (int32_t)tot_o;
		if (lane == 0) gc->ds_len = (int32_t)tot_b, gc->n_dsoff = (int32_t)tot_o;
	}
	return 0;
}

// per-read result header, one per read, in an array parallel to ReadMeta
struct ReadOut {
	int32_t status;
	int32_t n_gc, n_lc, n_a, rep_len;
	int32_t n_mz;
	uint32_t blob_size, blob2_size;
	int64_t blob_off;   // output pool: GChain[n_gc] | LLChain[n_lc] | u128 a[n_a]          (written by stage_gchain)
	int64_t blob2_off;  // output pool: per chain CIGAR (u64) | ds text | ds offsets        (written by stage_finish)
};

MG_HD inline uint64_t align8(uint64_t x) { return (x + 7) & ~(uint64_t)7; }

// state handed from the graph-chaining DP pass to the materialisation pass
struct GState {
	int32_t n_lc, n_u, n_gc, n_jobs;
	int64_t job_first;
	// followed by LChain lc[n_lc] | uint64 u[n_u] | uint32 gc_hash[n_gc]
};

// K6 for one read: graph chaining DP, overlap resolution, bridging plan (one lane).
// bridging plan of one read and the state K7 picks up again; one lane
MG_HD inline int stage_gchain_plan(const PipeCtx &c, ReadMeta &m, int rid, Arena &A, int32_t n_lc, LChain *lc, int32_t n_u, const uint64_t *u, const u128 *a, uint32_t *gc_hash)
{
	const MapOptDev &o = c.opt;
	int32_t n_gc = 0;
	// a read's jobs must be contiguous in the pool (they are consumed in order): reserve the worst case, one job per
	// linear chain, up front and mark the unused slots
	int64_t job_first;
	{
		int64_t off = pool_alloc(c.pool_gjobs, (uint64_t)(n_lc > 0? n_lc : 1) * sizeof(GwfaJob));
		if (off < 0) return MGB_E_POOL;
		job_first = off / (int64_t)sizeof(GwfaJob);
	}
	struct LocalEmit {
		GwfaJob *dst; int rid; int32_t n, cap;
		MG_HD int operator()(GwfaJob &J) { if (n >= cap) return MGB_E_INTERNAL; J.rid = rid; dst[n++] = J; return 0; }
	} le;
	le.dst = c.gjobs + job_first, le.rid = rid, le.n = 0, le.cap = n_lc > 0? n_lc : 1;
	MGB_TRY(gchain_prep(c.g, n_u, u, lc, a, m.hash, o.min_gc_cnt, o.min_gc_score, o.gdp_max_ed, batch_n_seg(c.b, rid), gc_hash, &n_gc, le));
	for (int32_t i = le.n; i < le.cap; ++i) le.dst[i].rid = -1; // unused reserved slots: skipped by the job kernel
	// persist
	uint64_t sz = sizeof(GState) + align8((uint64_t)n_lc * sizeof(LChain)) + (uint64_t)n_u * 8 + align8((uint64_t)n_gc * 4);
	int64_t goff = pool_alloc(c.pool_gstate, sz);
	if (goff < 0) return MGB_E_POOL;
	GState *gsb = (GState*)(c.gstate + goff);
	gsb->n_lc = n_lc, gsb->n_u = n_u, gsb->n_gc = n_gc, gsb->n_jobs = le.n, gsb->job_first = job_first;
	LChain *dlc = (LChain*)(gsb + 1);
	for (int32_t i = 0; i < n_lc; ++i) dlc[i] = lc[i];
	uint64_t *du = (uint64_t*)((char*)dlc + align8((uint64_t)n_lc * sizeof(LChain)));
	for (int32_t i = 0; i < n_u; ++i) du[i] = u[i];
	uint32_t *dh = (uint32_t*)(du + n_u);
	for (int32_t i = 0; i < n_gc; ++i) dh[i] = gc_hash[i];
	m.gstate_off = goff;
	return 0;
}

// K6 for one read, entered by all lanes of a warp: the graph-chaining DP is warp-wide (gchain_dp_w), the bridging plan and the
// hand-over to K7 (a few dozen words per read) are written by lane 0.
MG_HD inline int stage_gchain(const PipeCtx &c, ReadOut *routs, int rid, Arena &A, int lane)
{
	ReadMeta &m = c.meta[rid];
	ReadOut &ro = routs[rid];
	const MapOptDev &o = c.opt;
	if (lane == 0) ro.status = 0, ro.n_gc = ro.n_lc = ro.n_a = 0, ro.rep_len = m.rep_len, ro.n_mz = m.n_mz, ro.blob_size = ro.blob2_size = 0, ro.blob_off = ro.blob2_off = 0;
	if (m.status != 0) { if (lane == 0) ro.status = m.status; return 0; } // status 1: read skipped (empty or too long) -> no result object
	const uint64_t mark = A.top;
	const int32_t qlen = c.b.seq_len[rid];
	const u128 *a = c.anchor + m.a_off;
	int32_t n_lc = m.n_lc, n_u = 0;
	uint64_t *u = 0;
	LChain *lc;
	uint32_t *gc_hash;
	MGB_ALLOC(A, lc, LChain, n_lc);
	MGB_ALLOC(A, gc_hash, uint32_t, n_lc);
	for (int32_t i = lane; i < n_lc; i += MGB_W) lc[i] = c.lchain[m.lc_off + i];
	warp_sync();
	unsigned long long pt0 = prof_clock();
	GcParam gp;
	gp.max_dist_g = gp.max_dist_q = gp.bw = o.bw_long, gp.ref_bonus = o.ref_bonus, gp.chn_pen_gap = o.chn_pen_gap, gp.mask_level = o.mask_level; // reference: map-algo.c:461-462
	MGB_TRY(gchain_dp_w(A, c.g, c.lab, &n_lc, lc, qlen, gp, o.max_gc_skip, a, &u, &n_u, lane));
	if (lane == 0) { unsigned long long dt = prof_clock() - pt0; prof_add(c, PROF_GC_DP_CYC, dt); prof_max(c, PROF_GC_DP_MAX_CYC, dt); }
	int rc = 0;
	if (lane == 0) {
		Arena B = A;
		rc = stage_gchain_plan(c, m, rid, B, n_lc, lc, n_u, u, a, gc_hash);
		if (B.peak > A.peak) A.peak = B.peak;
	}
	rc = warp_bcast_i32(rc, 0);
	A.top = mark;
	return rc;
}

// K7a: one bridging alignment (reference: gchain1.c:349-381).  The wavefront containers of a typical bridge (tens of
// diagonals) fit a small per-warp arena in SHARED memory, which removes the global-memory latency from the sequential
// control flow; a bridge that outgrows it is redone with the worker's arena in HBM.  Lane 0 runs the alignment.
#ifndef MGB_GWFA_SMEM_KB
#define MGB_GWFA_SMEM_KB 1
#endif
#ifndef MGB_GWFA_SMEM_QL
#define MGB_GWFA_SMEM_QL 0
#endif
static const int GWFA_SMEM_ARENA = MGB_GWFA_SMEM_KB * 1024;
static const int GWFA_SMEM_MAX_QL = MGB_GWFA_SMEM_QL;

MG_HD inline int gwfa_job_run(Arena &A, const PipeCtx &c, int64_t job_idx, int lane, int32_t *smem)
{
	GwfaJob *J = &c.gjobs[job_idx];
	if (J->rid < 0) return 0;
	if (c.meta[J->rid].status < 0) return 0;
	GwfShared *sh = (GwfShared*)smem; // the one copy of the alignment state, seen by all lanes
	static_assert(sizeof(GwfShared) + 16 <= GWFA_SMEM_ARENA, "the alignment state has to fit the warp's slice of shared memory");
	const uint64_t sh_bytes = (sizeof(GwfShared) + 15) & ~(uint64_t)15;
	GwfOpt opt;
	opt.traceback = 1, opt.max_chk = 1000, opt.bw_dyn = 1000, opt.max_lag = J->max_ed / 2, opt.s_term = -1;
	opt.i_term = 500000000LL;
	const char *qseq = c.b.seq + c.b.seq_off[J->rid];
	unsigned long long t0 = prof_clock();
	// one call site for both attempts (the kernel holds one copy of the alignment code): first in the shared-memory arena, then,
	// if that was outgrown or not worth trying, in the worker's arena in HBM
	int rc = MGB_E_ARENA, in_smem = 0;
	uint64_t smem_peak = 0;
	for (int pass = J->ql < GWFA_SMEM_MAX_QL? 0 : 1; pass < 2; ++pass) { // longer bridges nearly always outgrow the shared-memory arena: do not try
		uint64_t top0 = 0;
		if (lane == 0) {
			if (pass == 0) arena_init(sh->A, (char*)smem + sh_bytes, GWFA_SMEM_ARENA - sh_bytes);
			else sh->A = A, sh->A.peak = A.top;
		}
		warp_sync();
		if (pass == 1) top0 = sh->A.top;
		rc = gwf_align_w(sh, c.g, opt, J->ql, qseq + J->qs, J->v0, J->end0, J->v1, J->end1, J->max_ed, lane);
		if (pass == 0) {
			if (rc != MGB_E_ARENA) { in_smem = 1, smem_peak = sh->A.peak; break; }
			warp_sync();
			continue;
		}
#if !MGB_ON_DEVICE && defined(MGB_HOSTSIM)
		if (getenv("MGB_DUMP_JOBS")) fprintf(stderr, "GWFAG\t%d\t%lu\n", J->ql, (unsigned long)(sh->A.peak - top0));
#endif
		(void)top0;
		if (sh->A.peak > A.peak) A.peak = sh->A.peak;
	}
	(void)smem_peak, (void)in_smem;
	if (lane == 0) {
		const GwfResult &r = sh->r;
		{ unsigned long long dt = prof_clock() - t0; prof_add(c, PROF_GC_GWFA_CYC, dt); prof_max(c, PROF_GWFA_MAX_CYC, dt << 16 | (unsigned long long)(J->ql < 65535? J->ql : 65535)); }
#if !MGB_ON_DEVICE && defined(MGB_HOSTSIM)
		if (getenv("MGB_DUMP_JOBS")) fprintf(stderr, "GWFA\t%d\t%d\t%ld\t%d\t%lu\t%d\n", J->ql, r.s, (long)r.n_iter, r.nv, (unsigned long)smem_peak, in_smem);
#endif
		if (rc == 0) {
			J->s = r.s, J->nv = r.s >= 0? r.nv : 0;
			if (r.s >= 0) {
				int64_t woff = pool_alloc(c.pool_walk, (uint64_t)r.nv * 4);
				if (woff < 0) rc = MGB_E_POOL;
				else {
					J->walk_off = woff / 4;
					for (int32_t i = 0; i < r.nv; ++i) c.walk[J->walk_off + i] = r.v[i];
				}
			}
		}
	}
	warp_sync();
	rc = warp_bcast_i32(rc, 0);
	return rc;
}

// What the sequential head of stage_gchain_gen<1>() hands to the warp-wide tail (stage_gchain_gen_w, parameter "gen_v2")
struct GenHand {
	GcSet gs;
	int64_t boff;
	uint64_t off_lc, off_a, sz, mark;
	unsigned long long pt3;
	int32_t want_plan, skip;
};

// gchain_cigar_plan() entered by all lanes of a warp.  Every kept anchor but the first yields exactly one item, which depends on
// the anchor, on the kept anchor in front of it and on the linear chains the two sit on: the kept anchors are listed by an
// ordered compaction, then every lane makes the item of one of them; the jobs of a chunk are allocated with one pool request.
MG_HD inline int gchain_cigar_plan_w(Arena &A, const PipeCtx &c, int rid, const GraphDev &g, GcSet &gt, int64_t lc_off_bytes, int lane)
{
	for (int32_t i = 0; i < gt.n_gc; ++i) {
		GChain *gc = &gt.gc[i];
		const int32_t off_a0 = gt.lc[gc->off].off, n_anchor = gc->n_anchor, l_beg = gc->off, l_end = gc->off + gc->cnt;
		uint64_t mark = A.top;
		int32_t *kj, *kl; // kept anchors (index into the chain's anchors) and the linear chain each sits on; entry 0 is the first anchor
		uint64_t *plan;
		MGB_ALLOC(A, kj, int32_t, n_anchor + 1);
		MGB_ALLOC(A, kl, int32_t, n_anchor + 1);
		MGB_ALLOC(A, plan, uint64_t, n_anchor + 8);
		if (lane == 0) kj[0] = 0, kl[0] = l_beg;
		int32_t nk = 1, bad = 0;
		for (int32_t base = 1; base < n_anchor; base += MGB_W) {
			const int32_t j = base + lane;
			int keep = 0, l = -1;
			if (j < n_anchor) {
				keep = !((gt.a[off_a0 + j].y & SEED_IGNORE) && j != n_anchor - 1);
				if (keep) { // the linear chains of a graph chain own disjoint, ascending runs of its anchors
					for (l = l_beg; l < l_end; ++l)
						if (off_a0 + j >= gt.lc[l].off && off_a0 + j < gt.lc[l].off + gt.lc[l].cnt) break;
					if (l >= l_end) bad = 1;
				}
			}
			const uint32_t mk = warp_ballot(keep);
			if (keep) { const int32_t at = nk + mask_rank(mk, lane); kj[at] = j, kl[at] = l; }
			nk += mask_count(mk);
		}
		if (warp_any(bad)) return MGB_E_INTERNAL;
		warp_sync();
		if (lane == 0) plan[0] = (uint64_t)(gt.a[off_a0].y >> 32 & 0xff) << 4 | 7;
		for (int32_t base = 1; base < nk; base += MGB_W) {
			const int32_t k = base + lane;
			uint64_t item = 0;
			int is_job = 0;
			int32_t l0 = 0, l = 0, l_seq = 0, qlen = 0;
			u128 p, q;
			p.x = p.y = q.x = q.y = 0;
			if (k < nk) {
				p = gt.a[off_a0 + kj[k]], q = gt.a[off_a0 + kj[k - 1]];
				l = kl[k], l0 = kl[k - 1];
				if (l == l0) l_seq = (int32_t)p.x - (int32_t)q.x;
				else {
					l_seq = g.seg_len[gt.lc[l0].v >> 1] - (int32_t)q.x - 1;
					for (int32_t t = l0 + 1; t < l; ++t) l_seq += g_vlen(g, gt.lc[t].v);
					l_seq += (int32_t)p.x + 1;
				}
				qlen = (int32_t)p.y - (int32_t)q.y;
				if (!(l_seq > 0 || qlen > 0)) bad = 1;
				else if (l_seq == 0) item = (uint64_t)(int64_t)qlen << 4 | 1;
				else if (qlen == 0) item = (uint64_t)(int64_t)l_seq << 4 | 2;
				else if (l_seq == qlen && (uint64_t)(int64_t)qlen <= (q.y >> 32 & 0xff)) item = (uint64_t)(int64_t)qlen << 4 | 7;
				else is_job = 1;
			}
			if (warp_any(bad)) return MGB_E_INTERNAL;
			const uint32_t mj = warp_ballot(is_job);
			int64_t jbase = 0;
			if (mj) {
				if (lane == 0) jbase = pool_alloc(c.pool_jobs, (uint64_t)mask_count(mj) * sizeof(WfaJob));
				jbase = (int64_t)warp_bcast_u64((uint64_t)jbase, 0);
				if (jbase < 0) return MGB_E_POOL;
			}
			if (is_job) {
				const int64_t joff = jbase + (int64_t)mask_rank(mj, lane) * (int64_t)sizeof(WfaJob);
				WfaJob *J = (WfaJob*)((char*)c.jobs + joff);
				J->rid = rid, J->gc = i, J->l0 = l0, J->l = l, J->t_beg = (int32_t)q.x + 1, J->t_last = (int32_t)p.x;
				J->tl = l_seq, J->ql = qlen, J->q_off = (int32_t)q.y + 1, J->n_cigar = 0, J->status = 0, J->lc_off = lc_off_bytes, J->cig_off = 0;
				item = PLAN_JOB | (uint64_t)(joff / (int64_t)sizeof(WfaJob));
			}
			if (k < nk) plan[k] = item;
		}
		warp_sync();
		int64_t poff = 0;
		if (lane == 0) poff = pool_alloc(c.pool_plan, (uint64_t)nk * 8);
		poff = (int64_t)warp_bcast_u64((uint64_t)poff, 0);
		if (poff < 0) return MGB_E_POOL;
		uint64_t *dst = c.plan + poff / 8;
		for (int32_t t = lane; t < nk; t += MGB_W) dst[t] = plan[t];
		if (lane == 0) gc->plan_off = poff / 8, gc->n_plan = nk;
		warp_sync();
		A.top = mark;
	}
	return 0;
}

// K7b for one read: materialise graph chains from the DP and the bridging results, post filters, alignment plan.
MG_HD inline int stage_gchain_gen_head(const PipeCtx &c, ReadOut *routs, int rid, Arena &A, GenHand *hand)
{
	ReadMeta &m = c.meta[rid];
	ReadOut &ro = routs[rid];
	const MapOptDev &o = c.opt;
	if (m.status != 0) { ro.status = m.status; return 0; }
	uint64_t mark = A.top;
	const char *qseq = c.b.seq + c.b.seq_off[rid];
	const int32_t qlen = c.b.seq_len[rid];
	const u128 *a = c.anchor + m.a_off;
	const GState *gsb = (const GState*)(c.gstate + m.gstate_off);
	const int32_t n_lc = gsb->n_lc, n_u = gsb->n_u;
	LChain *lc;
	MGB_ALLOC(A, lc, LChain, n_lc);
	{
		const LChain *slc = (const LChain*)(gsb + 1);
		for (int32_t i = 0; i < n_lc; ++i) lc[i] = slc[i];
	}
	const uint64_t *u = (const uint64_t*)((const char*)(gsb + 1) + align8((uint64_t)n_lc * sizeof(LChain)));
	const uint32_t *gc_hash = (const uint32_t*)(u + n_u);
	GwfaFeed feed;
	feed.job = c.gjobs + gsb->job_first, feed.walk_pool = c.walk, feed.next = 0, feed.n = gsb->n_jobs;
	GcSet gs;
	unsigned long long pt1 = prof_clock();
	MGB_TRY(gchain_gen(A, c.g, n_u, u, lc, a, m.hash, o.min_gc_cnt, o.min_gc_score, o.gdp_max_ed, batch_n_seg(c.b, rid), qseq, gs, &feed, gc_hash));
	gs.rep_len = m.rep_len;
	unsigned long long pt2 = prof_clock();
	prof_add(c, PROF_GC_GEN_CYC, pt2 - pt1);
	prof_add(c, PROF_GC_SHORTK_CYC, gs.cyc_shortk), prof_add(c, PROF_GC_EXTRA_CYC, gs.cyc_extra);
	MGB_TRY(gchain_set_parent(A, o.mask_level, gs.n_gc, gs.gc, o.sub_diff));
	gchain_flt_sub(o.pri_ratio, c.ix.k * 2, o.best_n, gs.n_gc, gs.gc);
	MGB_TRY(gchain_drop_flt(A, gs));
	MGB_TRY(gchain_set_mapq(o, gs, qlen, m.n_mz, o.min_gc_score));
	unsigned long long pt3 = prof_clock();
	prof_add(c, PROF_GC_POST_CYC, pt3 - pt2);
	// ---- part 1 of the result ----
	uint64_t off_lc = align8((uint64_t)gs.n_gc * sizeof(GChain));
	uint64_t off_a = off_lc + align8((uint64_t)gs.n_lc * sizeof(LLChain));
	uint64_t sz = off_a + align8((uint64_t)gs.n_a * sizeof(u128));
	int64_t boff = pool_alloc(c.pool_out, sz);
	if (boff < 0) return MGB_E_POOL;
	for (int32_t i = 0; i < gs.n_gc; ++i) {
		GChain *gc = &gs.gc[i];
		gc->has_cigar = 0, gc->n_cigar = 0, gc->cigar_off = gc->ds_off = gc->dsoff_off = 0, gc->ds_len = gc->n_dsoff = 0, gc->plan_off = 0, gc->n_plan = 0;
	}
	hand->gs = gs, hand->boff = boff, hand->off_lc = off_lc, hand->off_a = off_a, hand->sz = sz, hand->mark = mark, hand->pt3 = pt3;
	hand->want_plan = (o.flag & F_CIGAR) && gs.n_gc > 0 && batch_n_seg(c.b, rid) == 1, hand->skip = 0; // reference: map-algo.c:475
	return 0;
}

// stage_gchain_gen() entered by all lanes of a warp (parameter "gen_v2"): lane 0 runs the sequential head, then the plan and
// the copies of the first part of the result are shared by the lanes.
MG_HD inline int stage_gchain_gen(const PipeCtx &c, ReadOut *routs, int rid, Arena &A, int lane)
{
	GenHand h;
	h.gs.n_gc = h.gs.n_lc = h.gs.n_a = h.gs.rep_len = 0, h.gs.gc = 0, h.gs.lc = 0, h.gs.a = 0;
	h.gs.cyc_gwfa = h.gs.cyc_shortk = h.gs.cyc_extra = 0;
	h.boff = 0, h.off_lc = h.off_a = h.sz = 0, h.mark = A.top, h.pt3 = 0, h.want_plan = 0, h.skip = 1;
	int rc = 0;
	Arena B = A;
	if (lane == 0) rc = stage_gchain_gen_head(c, routs, rid, B, &h);
	rc = warp_bcast_i32(rc, 0);
	A.top = warp_bcast_u64(B.top, 0), A.peak = warp_bcast_u64(B.peak, 0);
	h.skip = warp_bcast_i32(h.skip, 0);
	warp_sync();
	if (rc < 0 || h.skip) { A.top = h.mark; return rc; }
	h.gs.n_gc = warp_bcast_i32(h.gs.n_gc, 0), h.gs.n_lc = warp_bcast_i32(h.gs.n_lc, 0), h.gs.n_a = warp_bcast_i32(h.gs.n_a, 0);
	h.gs.gc = (GChain*)warp_bcast_u64((uint64_t)h.gs.gc, 0), h.gs.lc = (LLChain*)warp_bcast_u64((uint64_t)h.gs.lc, 0), h.gs.a = (u128*)warp_bcast_u64((uint64_t)h.gs.a, 0);
	h.boff = (int64_t)warp_bcast_u64((uint64_t)h.boff, 0), h.off_lc = warp_bcast_u64(h.off_lc, 0), h.off_a = warp_bcast_u64(h.off_a, 0), h.sz = warp_bcast_u64(h.sz, 0);
	h.want_plan = warp_bcast_i32(h.want_plan, 0);
	if (h.want_plan) MGB_TRY(gchain_cigar_plan_w(A, c, rid, c.g, h.gs, h.boff + (int64_t)h.off_lc, lane));
	char *blob = c.out + h.boff;
	{
		GChain *d = (GChain*)blob;
		for (int32_t i = lane; i < h.gs.n_gc; i += MGB_W) d[i] = h.gs.gc[i];
		LLChain *dl = (LLChain*)(blob + h.off_lc);
		for (int32_t i = lane; i < h.gs.n_lc; i += MGB_W) dl[i] = h.gs.lc[i];
		u128 *da = (u128*)(blob + h.off_a);
		for (int32_t i = lane; i < h.gs.n_a; i += MGB_W) da[i] = h.gs.a[i];
	}
	if (lane == 0) {
		ReadOut &ro = routs[rid];
		ro.n_gc = h.gs.n_gc, ro.n_lc = h.gs.n_lc, ro.n_a = h.gs.n_a, ro.blob_size = (uint32_t)h.sz, ro.blob_off = h.boff;
		prof_add(c, PROF_GC_PLAN_CYC, prof_clock() - h.pt3);
	}
	warp_sync();
	A.top = h.mark;
	return 0;
}

// K8b for one read: stitch CIGARs, ds strings, part 2 of the result (one lane).
// Warp-uniform: all lanes enter.  The CIGAR stitching runs on lane 0, the ds strings and the copies on all lanes.
MG_HD inline int stage_finish(const PipeCtx &c, ReadOut *routs, int rid, Arena &A, int lane)
{
	ReadMeta &m = c.meta[rid];
	ReadOut &ro = routs[rid];
	if (m.status != 0) { if (lane == 0) ro.status = m.status; return 0; }
	if (!(c.opt.flag & F_CIGAR) || ro.n_gc == 0 || batch_n_seg(c.b, rid) != 1) return 0;
	uint64_t mark = A.top;
	const char *qseq = c.b.seq + c.b.seq_off[rid];
	char *blob = c.out + ro.blob_off;
	GcSet gs;
	gs.n_gc = ro.n_gc, gs.n_lc = ro.n_lc, gs.n_a = ro.n_a, gs.rep_len = ro.rep_len;
	gs.gc = (GChain*)blob;
	gs.lc = (LLChain*)(blob + align8((uint64_t)gs.n_gc * sizeof(GChain)));
	gs.a = (u128*)((char*)gs.lc + align8((uint64_t)gs.n_lc * sizeof(LLChain)));
	CigarOut *cg;
	DsOut *ds;
	MGB_ALLOC(A, cg, CigarOut, gs.n_gc);
	MGB_ALLOC(A, ds, DsOut, gs.n_gc);
	unsigned long long pt0 = prof_clock();
	MGB_TRY(gchain_cigar_finish_w(A, c, gs, cg, lane));
	unsigned long long pt1 = prof_clock();
	MGB_TRY(gchain_ds_w(A, c.g, qseq, gs, cg, ds, lane));
	if (lane == 0) prof_add(c, PROF_FIN_CIGAR_CYC, pt1 - pt0), prof_add(c, PROF_FIN_DS_CYC, prof_clock() - pt1);
	uint64_t sz = 0;
	for (int32_t i = 0; i < gs.n_gc; ++i)
		sz += align8((uint64_t)cg[i].n * 8) + align8((uint64_t)ds[i].len + 1) + align8((uint64_t)ds[i].n_off * 4);
	int64_t boff = 0;
	if (lane == 0) boff = pool_alloc(c.pool_out, sz);
	boff = (int64_t)warp_bcast_u64((uint64_t)boff, 0);
	if (boff < 0) return MGB_E_POOL;
	uint64_t at = (uint64_t)boff;
	for (int32_t i = 0; i < gs.n_gc; ++i) {
		GChain *gc = &gs.gc[i];
		const int64_t cigar_off = (int64_t)at; at += align8((uint64_t)cg[i].n * 8);
		const int64_t ds_off = (int64_t)at; at += align8((uint64_t)ds[i].len + 1);
		const int64_t dsoff_off = (int64_t)at; at += align8((uint64_t)ds[i].n_off * 4);
		if (lane == 0) gc->cigar_off = cigar_off, gc->ds_off = ds_off, gc->dsoff_off = dsoff_off;
		uint64_t *dc = (uint64_t*)(c.out + cigar_off);
		for (int32_t k = lane; k < cg[i].n; k += MGB_W) dc[k] = cg[i].cigar[k];
		char *dd = c.out + ds_off;
		for (int32_t k = lane; k < ds[i].len; k += MGB_W) dd[k] = ds[i].ds[k];
		if (lane == 0) dd[ds[i].len] = 0;
		int32_t *dof = (int32_t*)(c.out + dsoff_off);
		for (int32_t k = lane; k < ds[i].n_off; k += MGB_W) dof[k] = ds[i].off[k];
	}
	if (lane == 0) ro.blob2_off = boff, ro.blob2_size = (uint32_t)sz;
	A.top = mark;
	return 0;
}

} // namespace mgb
