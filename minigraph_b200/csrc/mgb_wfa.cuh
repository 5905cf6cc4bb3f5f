// mgb_wfa.cuh -- global alignment of one inter-anchor gap with the 2-piece affine wavefront algorithm.
// (reference: miniwfa.c:380-435 mwf_wfa_core, :281-308 wf_next_tb, :329-377 wf_traceback, :144-171 wf_stripe_shrink)
// Penalties: mismatch 4, gap1 4+2l, gap2 15+1l.  CIGAR bytes depend on the tie preferences in the recurrence
// (>= everywhere, E before F, mismatch before gap) and on the traceback state machine; both are kept.
//
// Layout: the last max_pen+1 = 17 wavefronts live in fixed slots of the worker arena, five int32 lanes each
// (H,E1,F1,E2,F2) padded by 17 cells of -inf on both sides; one traceback byte per (score,diagonal).
#pragma once
#include "mgb_common.cuh"

namespace mgb {

static const int32_t WF_NEG_INF = -0x40000000;
static const int WF_X = 4, WF_O1 = 4, WF_E1 = 2, WF_O2 = 15, WF_E2 = 1;
static const int WF_MAX_PEN = 16;           // max(x, o1+e1, o2+e2)
static const int WF_NSLICE = WF_MAX_PEN + 1;
static const int WF_PAD = WF_MAX_PEN + 1;   // m1 in the reference

struct WfSlice {
	int32_t lo, hi;
	int32_t *H, *E1, *F1, *E2, *F2; // indexable by diagonal d in [lo-PAD, hi+PAD]
};

struct WfTb1 { int32_t lo, hi; int64_t off; };

struct WfState {
	WfSlice sl[WF_NSLICE];
	int32_t *mem;       // WF_NSLICE slots
	int64_t slot_stride; // int32 elements per slot
	int64_t lane_stride; // int32 elements per lane inside a slot
	int32_t s, top, lo, hi;
};

MG_HD inline void wf_slice_bind(WfState &wf, int slot, int32_t lo, int32_t hi, int lane)
{
	WfSlice &f = wf.sl[slot];
	int32_t n = hi - lo + 1;
	int32_t *base = wf.mem + (int64_t)slot * wf.slot_stride;
	f.lo = lo, f.hi = hi;
	f.H = base + WF_PAD;
	f.E1 = f.H + wf.lane_stride;
	f.F1 = f.E1 + wf.lane_stride;
	f.E2 = f.F1 + wf.lane_stride;
	f.F2 = f.E2 + wf.lane_stride;
	for (int32_t i = -WF_PAD + lane; i < 0; i += MGB_W) f.H[i] = f.E1[i] = f.E2[i] = f.F1[i] = f.F2[i] = WF_NEG_INF;
	for (int32_t i = n + lane; i < n + WF_PAD; i += MGB_W) f.H[i] = f.E1[i] = f.E2[i] = f.F1[i] = f.F2[i] = WF_NEG_INF;
	f.H -= lo, f.E1 -= lo, f.E2 -= lo, f.F1 -= lo, f.F2 -= lo;
}

// reference: miniwfa.c:80-101 wf_stripe_add
MG_HD inline WfSlice &wf_stripe_add(WfState &wf, int32_t lo, int32_t hi, int lane)
{
	++wf.s;
	++wf.top;
	if (wf.top == WF_NSLICE) wf.top = 0;
	wf_slice_bind(wf, wf.top, lo, hi, lane);
	return wf.sl[wf.top];
}

MG_HD inline const WfSlice &wf_stripe_get(const WfState &wf, int32_t x)
{
	int32_t y = wf.top - x;
	if (y < 0) y += WF_NSLICE;
	return wf.sl[y];
}

MG_HD inline int wf_good_diag(int32_t d, int32_t k, int32_t tl, int32_t ql)
{
	return ((k >= -1 && k < tl) && (d + k >= -1 && d + k < ql));
}

// narrow [lo,hi] to diagonals that still hold a cell inside the DP matrix (reference: miniwfa.c:144-171)
MG_HD inline int wf_stripe_shrink(WfState &wf, int32_t tl, int32_t ql)
{
	int32_t j, d;
	for (d = wf.lo; d <= wf.hi; ++d) {
		for (j = 0; j < WF_NSLICE; ++j) {
			const WfSlice *p = &wf.sl[(wf.top + 1 + j) % WF_NSLICE];
			if (d < p->lo || d > p->hi) continue;
			if (wf_good_diag(d, p->H[d], tl, ql)) break;
			if (wf_good_diag(d, p->E1[d], tl, ql) || wf_good_diag(d, p->F1[d], tl, ql)) break;
			if (wf_good_diag(d, p->E2[d], tl, ql) || wf_good_diag(d, p->F2[d], tl, ql)) break;
		}
		if (j < WF_NSLICE) break;
	}
	if (d > wf.hi) return MGB_E_INTERNAL;
	wf.lo = d;
	for (d = wf.hi; d >= wf.lo; --d) {
		for (j = 0; j < WF_NSLICE; ++j) {
			const WfSlice *p = &wf.sl[(wf.top + 1 + j) % WF_NSLICE];
			if (d < p->lo || d > p->hi) continue;
			if (wf_good_diag(d, p->H[d], tl, ql)) break;
			if (wf_good_diag(d, p->E1[d], tl, ql) || wf_good_diag(d, p->F1[d], tl, ql)) break;
			if (wf_good_diag(d, p->E2[d], tl, ql) || wf_good_diag(d, p->F2[d], tl, ql)) break;
		}
		if (j < WF_NSLICE) break;
	}
	if (d < wf.lo) return MGB_E_INTERNAL;
	wf.hi = d;
	return 0;
}

#define MGB_WF_MAX(a, b) ((a) >= (b)? (a) : (b))

// one cell of the recurrence with its traceback byte (reference: miniwfa.c:281-308 wf_next_tb)
MG_HD inline uint8_t wf_cell(int32_t d, int32_t *H, int32_t *E1, int32_t *F1, int32_t *E2, int32_t *F2,
							 const int32_t *pHx, const int32_t *pHo1, const int32_t *pHo2,
							 const int32_t *pE1, const int32_t *pF1, const int32_t *pE2, const int32_t *pF2)
{
	int32_t h, f, e, e1, e2, f1, f2;
	uint8_t x = 0, ze, zf, z;
	x |= pHo1[d-1] >= pE1[d-1]? 0 : 0x08;
	e1 = MGB_WF_MAX(pHo1[d-1], pE1[d-1]);
	x |= pHo2[d-1] >= pE2[d-1]? 0 : 0x20;
	e2 = MGB_WF_MAX(pHo2[d-1], pE2[d-1]);
	ze = e1 >= e2? 1 : 3;
	e = MGB_WF_MAX(e1, e2);
	x |= pHo1[d+1] >= pF1[d+1]? 0 : 0x10;
	f1 = MGB_WF_MAX(pHo1[d+1], pF1[d+1]) + 1;
	x |= pHo2[d+1] >= pF2[d+1]? 0 : 0x40;
	f2 = MGB_WF_MAX(pHo2[d+1], pF2[d+1]) + 1;
	zf = f1 >= f2? 2 : 4;
	f = MGB_WF_MAX(f1, f2);
	z = e >= f? ze : zf;
	h = MGB_WF_MAX(e, f);
	z = pHx[d] + 1 >= h? 0 : z;
	E1[d] = e1, E2[d] = e2, F1[d] = f1, F2[d] = f2;
	H[d] = MGB_WF_MAX(pHx[d] + 1, h);
	return x | z;
}

struct WfResult {
	int32_t s;        // score, -1 if the iteration cap was hit
	int32_t n_cigar;
	int64_t n_iter;
	uint32_t *cigar;  // len<<4|op  (op: 7 '=', 8 'X', 1 'I', 2 'D'), allocated at the caller's mark
};

MG_HD inline int wf_cigar_push1(Arena &A, AVec<uint32_t> &c, int32_t op, int32_t len)
{
	if (c.n && (uint32_t)op == (c.a[c.n-1] & 0xf)) c.a[c.n-1] += (uint32_t)len << 4;
	else {
		uint32_t x = (uint32_t)len << 4 | (uint32_t)op;
		MGB_TRY(avec_push(A, c, x));
	}
	return 0;
}

// sequential traceback on one lane; writes the CIGAR in input order to cig_store
MG_HD inline int wf_traceback_lane0(Arena &A, int32_t tl, const char *ts, int32_t ql, const char *qs, const AVec<WfTb1> &tb, const AVec<uint8_t> &tbx,
									int32_t last_state, uint32_t *cig_store, int64_t max_cigar, int32_t *n_cigar)
{
	AVec<uint32_t> cigar;
	avec_init(cigar);
	int32_t i = ql - 1, k = tl - 1, s = (int32_t)tb.n - 1, last = last_state;
	while (i >= 0 && k >= 0) {
		int32_t k0 = k, j, x, state, ext;
		if (last == 0) {
			while (i >= 0 && k >= 0 && qs[i] == ts[k]) --i, --k;
			if (k0 - k > 0) MGB_TRY(wf_cigar_push1(A, cigar, 7, k0 - k));
			if (i < 0 || k < 0) break;
		}
		if (s < 0) return MGB_E_INTERNAL;
		j = i - k - tb.a[s].lo;
		if (j < 0 || j > tb.a[s].hi - tb.a[s].lo) return MGB_E_INTERNAL;
		x = tbx.a[tb.a[s].off + j];
		state = last == 0? x & 7 : last;
		ext = state > 0? x >> (state + 2) & 1 : 0;
		if (state == 0) {
			MGB_TRY(wf_cigar_push1(A, cigar, 8, 1));
			--i, --k, s -= WF_X;
		} else if (state == 1) {
			MGB_TRY(wf_cigar_push1(A, cigar, 1, 1));
			--i, s -= ext? WF_E1 : WF_O1 + WF_E1;
		} else if (state == 3) {
			MGB_TRY(wf_cigar_push1(A, cigar, 1, 1));
			--i, s -= ext? WF_E2 : WF_O2 + WF_E2;
		} else if (state == 2) {
			MGB_TRY(wf_cigar_push1(A, cigar, 2, 1));
			--k, s -= ext? WF_E1 : WF_O1 + WF_E1;
		} else if (state == 4) {
			MGB_TRY(wf_cigar_push1(A, cigar, 2, 1));
			--k, s -= ext? WF_E2 : WF_O2 + WF_E2;
		} else return MGB_E_INTERNAL;
		last = state > 0 && ext? state : 0;
	}
	if (i >= 0) MGB_TRY(wf_cigar_push1(A, cigar, 1, i + 1));
	else if (k >= 0) MGB_TRY(wf_cigar_push1(A, cigar, 2, k + 1));
	if (cigar.n > max_cigar) return MGB_E_INTERNAL;
	for (int64_t c = 0; c < cigar.n; ++c) cig_store[c] = cigar.a[cigar.n - 1 - c]; // back to input order
	*n_cigar = (int32_t)cigar.n;
	return 0;
}

// Exact WFA with traceback (reference: miniwfa.c:380-435 + :603-615 with opt.step == 0).
// ts/qs need not be padded: the extension loop checks the sequence ends explicitly, which is what the reference's
// distinct padding characters achieve (miniwfa.c:182-226).
//
// Warp-uniform: all lanes enter with identical arguments and identical arena state.  Lanes own diagonals
// (d = lo + lane, + 32, ...) in the three data-parallel phases of a score step -- pad initialisation, exact-match
// extension, and the recurrence -- while the scalar bookkeeping (ring of 17 wavefronts, [lo,hi] tracking, the
// every-256-scores shrink, iteration counting) is replicated.  The traceback is sequential and runs on lane 0.
// Only one diagonal (d = ql - tl) can reach the end of both sequences, so "first diagonal that finishes" of the
// sequential reference needs no ordering between lanes.
MG_HD inline int wfa_exact(Arena &A, int32_t tl, const char *ts, int32_t ql, const char *qs, int64_t max_iter, WfResult *r, int lane)
{
	uint64_t mark = A.top;
	WfState wf;
	int32_t last_state = 0, stopped = 0;
	r->s = -1, r->n_cigar = 0, r->n_iter = 0, r->cigar = 0;
	// the CIGAR is built bottom-up at the caller's mark after the scratch is released; reserve its worst case first
	uint32_t *cig_store;
	const int64_t max_cigar = (int64_t)tl + ql + 2;
	MGB_ALLOC(A, cig_store, uint32_t, max_cigar);
	uint64_t mark_keep = A.top;
	{
		int64_t maxw = (int64_t)tl + ql + 1;
		wf.lane_stride = maxw + 2 * WF_PAD;
		wf.slot_stride = 5 * wf.lane_stride;
		MGB_ALLOC(A, wf.mem, int32_t, wf.slot_stride * WF_NSLICE);
	}
	AVec<WfTb1> tb; AVec<uint8_t> tbx;
	avec_init(tb), avec_init(tbx);
	// reference: miniwfa.c:103-121 wf_stripe_init
	wf.s = 0, wf.top = 0, wf.lo = wf.hi = 0;
	for (int i = 0; i < WF_NSLICE; ++i) {
		WfSlice &f = wf_stripe_add(wf, 0, 0, lane);
		if (lane == 0) f.H[0] = f.E1[0] = f.E2[0] = f.F1[0] = f.F2[0] = WF_NEG_INF;
	}
	wf.s = 0;
	if (lane == 0) wf.sl[wf.top].H[0] = -1;
	warp_sync();

	for (;;) {
		WfSlice *p = &wf.sl[wf.top];
		int32_t lo, hi, *H = p->H;
		int hit = 0, hit_noext = 0;
		for (int32_t d = p->lo + lane; d <= p->hi; d += MGB_W) { // extension along exact matches
			int32_t k = H[d], k0 = k;
			if (k < -1 || d + k < -1 || k >= tl || d + k >= ql) continue;
			while (k + 1 < tl && d + k + 1 < ql && ts[k + 1] == qs[d + k + 1]) ++k;
			if (k == tl - 1 && d + k == ql - 1) hit = 1, hit_noext = (k == k0);
			else H[d] = k;
		}
		warp_sync();
		if (warp_any(hit)) {
			if (warp_any(hit && hit_noext)) {
				const WfTb1 &t1 = tb.a[tb.n - 1];
				last_state = tbx.a[t1.off + ((ql - tl) - t1.lo)] & 7;
			}
			break;
		}
		lo = wf.lo > -tl? wf.lo - 1 : -tl;
		hi = wf.hi < ql? wf.hi + 1 : ql;
		{ // reference: miniwfa.c:313-327 wf_next_basic (traceback variant)
			const WfSlice &ft = wf_stripe_add(wf, lo, hi, lane);
			const WfSlice &fx = wf_stripe_get(wf, WF_X);
			const WfSlice &fo1 = wf_stripe_get(wf, WF_O1 + WF_E1);
			const WfSlice &fo2 = wf_stripe_get(wf, WF_O2 + WF_E2);
			const WfSlice &fe1 = wf_stripe_get(wf, WF_E1);
			const WfSlice &fe2 = wf_stripe_get(wf, WF_E2);
			WfTb1 t1;
			t1.lo = lo, t1.hi = hi, t1.off = tbx.n;
			MGB_TRY(avec_reserve_w(A, tb, tb.n + 1, lane));
			if (lane == 0) tb.a[tb.n] = t1;
			++tb.n;
			MGB_TRY(avec_reserve_w(A, tbx, tbx.n + (hi - lo + 1), lane));
			uint8_t *ax = tbx.a + tbx.n - lo;
			tbx.n += hi - lo + 1;
			for (int32_t dd = lo + lane; dd <= hi; dd += MGB_W)
				ax[dd] = wf_cell(dd, ft.H, ft.E1, ft.F1, ft.E2, ft.F2, fx.H, fo1.H, fo2.H, fe1.E1, fe1.F1, fe2.E2, fe2.F2);
			warp_sync();
			if (ft.H[lo] >= -1 || ft.E1[lo] >= -1 || ft.F1[lo] >= -1 || ft.E2[lo] >= -1 || ft.F2[lo] >= -1) wf.lo = lo;
			if (ft.H[hi] >= -1 || ft.E1[hi] >= -1 || ft.F1[hi] >= -1 || ft.E2[hi] >= -1 || ft.F2[hi] >= -1) wf.hi = hi;
		}
		if ((wf.s & 0xff) == 0) MGB_TRY(wf_stripe_shrink(wf, tl, ql));
		r->n_iter += hi - lo + 1;
		if (max_iter > 0 && r->n_iter > max_iter) { stopped = 1; break; }
	}
	r->s = stopped? -1 : wf.s;
	if (!stopped) { // reference: miniwfa.c:329-377 wf_traceback (lane 0; the outcome is broadcast)
		int rc = 0;
		int32_t n_cig = 0;
		if (lane == 0) {
			rc = wf_traceback_lane0(A, tl, ts, ql, qs, tb, tbx, last_state, cig_store, max_cigar, &n_cig);
		}
		rc = warp_bcast_i32(rc, 0), n_cig = warp_bcast_i32(n_cig, 0);
		warp_sync();
		if (rc < 0) { A.top = mark; return rc; }
		r->n_cigar = n_cig, r->cigar = cig_store;
	}
	A.top = mark_keep;
	return 0;
}

} // namespace mgb
