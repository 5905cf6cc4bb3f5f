// mgb_wfa_cta.cuh -- the tier-3 gap alignment entered by a whole thread block: the diagonals of a wavefront are spread over
// MGB_CTA_T threads instead of the lanes of one warp.  It is meant for the handful of long gaps whose single-warp time is
// what the tier-3 kernel waits for at its end.  Same scheme and same results as wfa_ring_g() (mgb_wfa.cuh), whose body
// the function below mirrors with block-level votes (mgb_cta.cuh); a gap it cannot take is left to the warp kernel.
// Off by default (engine parameter "cta_len").
#pragma once
#include "mgb_galign.cuh"
#include "mgb_cta.cuh"

namespace mgb {

// returns 0: r filled (r->s < 0 when the cell cap was hit), 1: not applicable, < 0: error
MG_HD inline int wfa_ring_cta(Arena &A, CtaCtx &cx, int32_t tl, const char *ts, int32_t ql, const char *qs, int64_t max_iter, WfResult *r,
							 uint32_t *cig_store, int64_t max_cigar, int tid)
{
	if (tl + ql > 16000 || tl <= 0 || ql <= 0) return 1;
	uint64_t mark = A.top;
	int32_t W = 64;
	while (W < tl + ql + 2) W <<= 1;
	const int32_t mask = W - 1;
	wf_cell_t *cells;
	MGB_ALLOC(A, cells, wf_cell_t, (int64_t)5 * 17 * W);
	wf_cell_t *H = cells, *E1 = H + 17 * W, *F1 = E1 + 17 * W, *E2 = F1 + 17 * W, *F2 = E2 + 17 * W;
	const int32_t max_rows = tl + ql + 64; // a score never exceeds the cost of deleting one string and inserting the other: one row per score
	WfTbRow *rows;
	MGB_ALLOC(A, rows, WfTbRow, max_rows);
	int32_t n_rows = 0;
	int32_t wlo = 0, whi = 0, last_state = 0, s = 0, stopped = 0;
	int64_t n_iter = 0;
	int hs = 0;
#define MGB_WF_RG(lo_, hi_) ((uint32_t)((hi_) + 0x8000) << 16 | (uint32_t)((lo_) + 0x8000))
	uint32_t g0 = MGB_WF_RG(0, 0), g1 = MGB_WF_RG(1, 0), g2 = g1, g3 = g1, g4 = g1, g5 = g1, g6 = g1, g7 = g1, g8 = g1, g9 = g1, g10 = g1, g11 = g1, g12 = g1, g13 = g1, g14 = g1, g15 = g1, g16 = g1;
	int hit = 0, hit_noext = 0;
	if (tid == 0) { // score 0: the main diagonal, extended from the corner
		const int32_t c0 = (1 << 20) & mask;
		E1[c0] = F1[c0] = E2[c0] = F2[c0] = (wf_cell_t)WF_NEG_INF16;
		int32_t k0 = -1, k = -1;
		k = wf_extend(ts, qs, k0, 0);
		if (k == tl - 1 && k == ql - 1) hit = 1, hit_noext = (k == k0), k = k0;
		H[c0] = (wf_cell_t)k;
	}
	{
		const uint32_t vb = cta_or_bits(cx, (hit? 1u : 0u) | (hit_noext? 2u : 0u), tid);
		hit = vb & 1, hit_noext = vb >> 1 & 1;
	}
	for (;;) {
		if (hit) {
			if (hit_noext) { WfTbArena t; t.row = rows; last_state = t.get(n_rows - 1, ql - tl) & 7; }
			break;
		}
		const int32_t lo = wlo > -tl? wlo - 1 : -tl;
		const int32_t hi = whi < ql? whi + 1 : ql;
		const int32_t width = hi - lo + 1;
		const int32_t ns = s + 1;
		const int nhs = hs + 1 == 17? 0 : hs + 1;
		if (n_rows >= max_rows) { A.top = mark; return 1; }
		uint8_t *x;
		MGB_ALLOC(A, x, uint8_t, width);
		if (tid == 0) rows[n_rows].lo = lo, rows[n_rows].hi = hi, rows[n_rows].x = x;
		++n_rows;
		uint8_t *ax = x - lo;
		const int r4 = nhs >= WF_X? nhs - WF_X : nhs - WF_X + 17, r6 = nhs >= WF_O1 + WF_E1? nhs - (WF_O1 + WF_E1) : nhs - (WF_O1 + WF_E1) + 17;
		const int r16 = nhs >= WF_O2 + WF_E2? nhs - (WF_O2 + WF_E2) : nhs - (WF_O2 + WF_E2) + 17, r2 = nhs >= WF_E1? nhs - WF_E1 : nhs - WF_E1 + 17;
		WfSrcG sHx, sHo1, sHo2, sE1, sF1, sE2, sF2;
#define MGB_WF_SRC(dst, arr, slot, g) (dst).p = (arr) + (int64_t)(slot) * W, (dst).lo = (int32_t)((g) & 0xffffu) - 0x8000, (dst).hi = (int32_t)((g) >> 16) - 0x8000
		MGB_WF_SRC(sHx, H, r4, g3); MGB_WF_SRC(sHo1, H, r6, g5); MGB_WF_SRC(sHo2, H, r16, g15);
		MGB_WF_SRC(sE1, E1, r2, g1); MGB_WF_SRC(sF1, F1, r2, g1); MGB_WF_SRC(sE2, E2, hs, g0); MGB_WF_SRC(sF2, F2, hs, g0);
#undef MGB_WF_SRC
		wf_cell_t *nH = H + (int64_t)nhs * W, *nE1 = E1 + (int64_t)nhs * W, *nF1 = F1 + (int64_t)nhs * W, *nE2 = E2 + (int64_t)nhs * W, *nF2 = F2 + (int64_t)nhs * W;
		int grow_lo = 0, grow_hi = 0;
		for (int32_t d = lo + tid; d <= hi; d += MGB_CTA_T) { // reference: miniwfa.c:281-308 wf_next_tb, then :212-226 on the new cell
			int32_t h, f, e, e1, e2, f1, f2, a0, b0;
			uint8_t xb = 0, ze, zf, z;
			a0 = wfg_at(sHo1, d - 1, mask), b0 = wfg_at(sE1, d - 1, mask);
			xb |= a0 >= b0? 0 : 0x08; e1 = MGB_WF_MAX(a0, b0);
			a0 = wfg_at(sHo2, d - 1, mask), b0 = wfg_at(sE2, d - 1, mask);
			xb |= a0 >= b0? 0 : 0x20; e2 = MGB_WF_MAX(a0, b0);
			ze = e1 >= e2? 1 : 3;
			e = MGB_WF_MAX(e1, e2);
			a0 = wfg_at(sHo1, d + 1, mask), b0 = wfg_at(sF1, d + 1, mask);
			xb |= a0 >= b0? 0 : 0x10; f1 = MGB_WF_MAX(a0, b0) + 1;
			a0 = wfg_at(sHo2, d + 1, mask), b0 = wfg_at(sF2, d + 1, mask);
			xb |= a0 >= b0? 0 : 0x40; f2 = MGB_WF_MAX(a0, b0) + 1;
			zf = f1 >= f2? 2 : 4;
			f = MGB_WF_MAX(f1, f2);
			z = e >= f? ze : zf;
			h = MGB_WF_MAX(e, f);
			a0 = wfg_at(sHx, d, mask) + 1;
			z = a0 >= h? 0 : z;
			h = MGB_WF_MAX(a0, h);
			ax[d] = xb | z;
			if (h >= -1 || e1 >= -1 || f1 >= -1 || e2 >= -1 || f2 >= -1) {
				if (d == lo) grow_lo = 1;
				if (d == hi) grow_hi = 1;
			}
			if (!(h < -1 || d + h < -1 || h >= tl || d + h >= ql)) {
				const int32_t k = wf_extend(ts, qs, h, d);
				if (k == tl - 1 && d + k == ql - 1) hit = 1, hit_noext = (k == h);
				else h = k;
			}
			const int32_t c = (d + (1 << 20)) & mask;
			nE1[c] = (wf_cell_t)e1, nF1[c] = (wf_cell_t)f1, nE2[c] = (wf_cell_t)e2, nF2[c] = (wf_cell_t)f2, nH[c] = (wf_cell_t)h;
		}
		{ // one vote for the four flags; its barrier also closes the writes of this wavefront
			const uint32_t vb = cta_or_bits(cx, (grow_lo? 1u : 0u) | (grow_hi? 2u : 0u) | (hit? 4u : 0u) | (hit && hit_noext? 8u : 0u), tid);
			if (vb & 1) wlo = lo;
			if (vb & 2) whi = hi;
			hit = vb >> 2 & 1, hit_noext = vb >> 3 & 1;
		}
		g16 = g15, g15 = g14, g14 = g13, g13 = g12, g12 = g11, g11 = g10, g10 = g9, g9 = g8, g8 = g7, g7 = g6, g6 = g5, g5 = g4, g4 = g3, g3 = g2, g2 = g1, g1 = g0;
		g0 = MGB_WF_RG(lo, hi);
		s = ns, hs = nhs;
		if ((s & 0xff) == 0) { // reference: miniwfa.c:144-171 wf_stripe_shrink: keep the diagonals on which one of the 17 wavefronts still has a cell inside the matrix
			const uint32_t gg[17] = { g0, g1, g2, g3, g4, g5, g6, g7, g8, g9, g10, g11, g12, g13, g14, g15, g16 };
			int32_t nlo = 0, nhi = 0, found = 0;
			for (int pass = 0; pass < 2; ++pass) {
				found = 0;
				for (int32_t base = pass == 0? wlo : whi; pass == 0? base <= whi : base >= wlo; base += pass == 0? MGB_CTA_T : -MGB_CTA_T) {
					const int32_t d = pass == 0? base + tid : base - tid;
					int good = 0;
					if (d >= wlo && d <= whi) {
						const int32_t c = (d + (1 << 20)) & mask;
						for (int j = 0; j < 17 && !good; ++j) {
							const int32_t jl = (int32_t)(gg[j] & 0xffffu) - 0x8000, jh = (int32_t)(gg[j] >> 16) - 0x8000;
							if (d < jl || d > jh) continue;
							const int slot = hs >= j? hs - j : hs - j + 17;
							const int64_t o = (int64_t)slot * W + c;
							good = wf_good_diag(d, H[o], tl, ql) || wf_good_diag(d, E1[o], tl, ql) || wf_good_diag(d, F1[o], tl, ql) || wf_good_diag(d, E2[o], tl, ql) || wf_good_diag(d, F2[o], tl, ql);
						}
					}
					const int32_t first = cta_min_i32(cx, good? tid : INT32_MAX, tid); // the first thread of this round that holds a live diagonal
					if (first != INT32_MAX) { found = 1; if (pass == 0) nlo = base + first; else nhi = base - first; break; }
				}
				if (!found) break;
			}
			if (!found) { A.top = mark; return MGB_E_INTERNAL; }
			wlo = nlo, whi = nhi;
		}
		n_iter += width;
		if (max_iter > 0 && n_iter > max_iter) { stopped = 1; break; }
	}
#undef MGB_WF_RG
	r->n_iter = n_iter;
	r->s = stopped? -1 : s;
	if (!stopped) {
		int rc = 0;
		int32_t n_cig = 0;
		int64_t first = 0;
		if (tid == 0) {
			WfTbArena t; t.row = rows;
			rc = wf_traceback(t, n_rows, tl, ts, ql, qs, last_state, cig_store, max_cigar, &n_cig, &first);
		}
		rc = cta_bcast_i32(cx, rc, tid), n_cig = cta_bcast_i32(cx, n_cig, tid), first = (int64_t)cta_bcast_u64(cx, (uint64_t)first, tid);
		if (rc < 0) { A.top = mark; return rc; }
		r->n_cigar = n_cig, r->cigar = cig_store + first;
	}
	A.top = mark;
	return 0;
}

MG_HD inline void wf_stage_seq_cta(char *dst, const char *src, int32_t len, uint8_t sentinel, int tid)
{
	for (int32_t i = tid; i < len; i += MGB_CTA_T) dst[i] = src[i];
	for (int32_t i = len + tid; i < len + WF_SEQ_PAD; i += MGB_CTA_T) dst[i] = (char)sentinel;
}

// One entry of the tier-3 queue, taken by a block (block-uniform: all threads enter with identical arguments).  Does what
// wfa_job_run() does for tier 3 when the gap is long enough to be worth a block and small enough for the 16-bit ring; the
// finished entry is struck from the queue (-1).  Anything else -- short gaps, gaps over the ring's limit, the cell cap, an
// arena that is too small -- is left in the queue for the warp kernel that runs next, so this pass can only take work away.
MG_HD inline int wfa_job_cta(Arena &A, CtaCtx &cx, const PipeCtx &c, int32_t qidx, int tid)
{
	const int32_t job_idx = c.jobq[1][qidx];
	if (job_idx < 0) return 0;
	WfaJob *J = &c.jobs[job_idx];
	const int32_t rid = J->rid, l0 = J->l0, l = J->l, tl = J->tl, ql = J->ql;
	if (c.meta[rid].status < 0) return 0;
	if (c.cta_len <= 0 || tl + ql < c.cta_len || tl + ql > 16000 || tl <= 0 || ql <= 0) return 0;
	const GraphDev &g = c.g;
	const LLChain *lc = (const LLChain*)(c.out + J->lc_off);
	const char *qs_g = c.b.seq + c.b.seq_off[rid] + J->q_off;
	char *ts, *qs;
	uint32_t *cig_store;
	const int64_t max_cigar = (int64_t)tl + ql + 2;
	ts = (char*)arena_alloc(A, (uint64_t)(tl + WF_SEQ_PAD + 4));
	qs = (char*)arena_alloc(A, (uint64_t)(ql + WF_SEQ_PAD + 4));
	cig_store = (uint32_t*)arena_alloc(A, (uint64_t)max_cigar * 4);
	if (ts == 0 || qs == 0 || cig_store == 0) return 0;
	if (l == l0) wf_stage_seq_cta(ts, g_vseq(g, lc[l0].v) + J->t_beg, tl, 0xfe, tid);
	else { // the target stitched across the walk (reference: galign.c:76-93), straight into the staging buffer
		int32_t n = g_vlen(g, lc[l0].v) - J->t_beg, at = 0;
		const char *s = g_vseq(g, lc[l0].v) + J->t_beg;
		for (int32_t x = tid; x < n; x += MGB_CTA_T) ts[at + x] = s[x];
		at += n;
		for (int32_t k = l0 + 1; k < l; ++k) {
			s = g_vseq(g, lc[k].v), n = g_vlen(g, lc[k].v);
			for (int32_t x = tid; x < n; x += MGB_CTA_T) ts[at + x] = s[x];
			at += n;
		}
		s = g_vseq(g, lc[l].v), n = J->t_last + 1;
		for (int32_t x = tid; x < n; x += MGB_CTA_T) ts[at + x] = s[x];
		at += n;
		if (at != tl) return 0; // the warp kernel reports it
		for (int32_t i = tl + tid; i < tl + WF_SEQ_PAD; i += MGB_CTA_T) ts[i] = (char)0xfe;
	}
	wf_stage_seq_cta(qs, qs_g, ql, 0xff, tid);
	cta_sync();
	WfResult rst;
	rst.s = -1, rst.n_cigar = 0, rst.n_iter = 0, rst.cigar = 0;
	unsigned long long pt0 = prof_clock();
	const int rc = wfa_ring_cta(A, cx, tl, ts, ql, qs, 100000000LL, &rst, cig_store, max_cigar, tid);
	if (rc != 0 || rst.s < 0) return 0;
	int64_t coff = 0;
	if (tid == 0) coff = pool_alloc(c.pool_cig, (uint64_t)rst.n_cigar * 4);
	coff = (int64_t)cta_bcast_u64(cx, (uint64_t)coff, tid);
	if (coff < 0) return 0; // the warp kernel will hit the same wall and report it
	uint32_t *dst = (uint32_t*)((char*)c.cig + coff);
	for (int32_t x = tid; x < rst.n_cigar; x += MGB_CTA_T) dst[x] = rst.cigar[x];
	if (tid == 0) {
		J->n_cigar = rst.n_cigar, J->cig_off = coff;
		c.jobq[1][qidx] = -1;
		unsigned long long dt = prof_clock() - pt0;
		prof_add(c, PROF_WFA_CTA_CYC, dt), prof_add(c, PROF_WFA_CTA_CYC + 1, 1);
		prof_add(c, PROF_WFA_CELLS, (unsigned long long)rst.n_iter);
	}
	return 0;
}

} // namespace mgb
