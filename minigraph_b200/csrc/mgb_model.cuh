// mgb_model.cuh -- read-only device model (graph, index, options) and per-read records.
#pragma once
#include "mgb_common.cuh"

namespace mgb {

// flattened arc, copied verbatim and in host order from gfa_t::arc[] (reference: gfa.h:33-39; order matters)
struct DevArc {
	uint32_t w;     // head vertex of the arc's target
	uint32_t lv;    // low 32 bits of v_lv = len(v) - ov
	int32_t rank;
	int32_t ow;
};

struct GraphDev {
	int32_t n_seg;
	const int32_t *seg_len;    // [n_seg]
	const uint64_t *vseq_off;  // [2*n_seg] offset of vertex v's sequence in seq[] (v odd = reverse complement copy)
	const char *seq;           // upper-case ASCII; every segment stored on both strands (reference: gfa-ed.c:24-42)
	const uint64_t *arc_idx;   // [2*n_seg] start<<32 | n   (reference: gfa_t::idx)
	const DevArc *arc;
	// for MG_M_NO_DIAG (map-algo.c:167-178): the id of the name a segment goes by (its stable sequence if it has one) and its offset there
	const int32_t *seg_name_id; // [n_seg]
	const int32_t *seg_soff;    // [n_seg], 0 when the segment is named by itself
};

MG_HD inline int32_t g_vlen(const GraphDev &g, uint32_t v) { return g.seg_len[v >> 1]; }
MG_HD inline const char *g_vseq(const GraphDev &g, uint32_t v) { return g.seq + g.vseq_off[v]; }
MG_HD inline int32_t g_arc_n(const GraphDev &g, uint32_t v) { return (int32_t)(uint32_t)g.arc_idx[v]; }
MG_HD inline const DevArc *g_arc_a(const GraphDev &g, uint32_t v) { return g.arc + (g.arc_idx[v] >> 32); }

// minimizer index: open addressing, one 16-byte slot per distinct minimizer.
//   slot.x = minimizer<<1 | singleton      (all ones = empty)
//   slot.y = singleton ? position : start<<32 | n  into pos[]
// position = seg<<32 | lastPos<<1 | strand, occurrence lists ascending (reference: index.c:115-165)
struct IndexDev {
	int32_t k, w;
	uint64_t n_slots_mask;
	const u128 *slot;
	const uint64_t *pos;
};

MG_HD inline uint64_t idx_slot_hash(uint64_t minier) { return (minier * 0x9E3779B97F4A7C15ULL) >> 20; }

// reference: index.c:67-72 mg_idx_get(); returns pointer to the occurrence list and its length
MG_HD inline const uint64_t *idx_get(const IndexDev &ix, uint64_t minier, int *n)
{
	uint64_t h = idx_slot_hash(minier) & ix.n_slots_mask;
	for (;;) {
		const u128 *s = &ix.slot[h];
		uint64_t kx = s->x;
		if (kx == ~0ULL) { *n = 0; return 0; }
		if (kx >> 1 == minier) {
			if (kx & 1) { *n = 1; return &s->y; }
			*n = (int)(uint32_t)s->y;
			return ix.pos + (s->y >> 32);
		}
		h = (h + 1) & ix.n_slots_mask;
	}
}

// the scalars of mg_mapopt_t consumed on the hot path (reference: minigraph.h:51-77), plus host-computed floats
struct MapOptDev {
	uint64_t flag;
	int32_t seed, max_qlen;
	int32_t occ_max1;
	int32_t bw, bw_long;
	int32_t rmq_size_cap, rmq_rescue_size;
	float rmq_rescue_ratio;
	int32_t max_gap_pre, max_gap, max_gap_ref, max_frag_len;
	float chn_pen_gap, chn_pen_skip; // ALREADY multiplied by expf(-div*k) on the host (glibc)
	int32_t max_lc_skip, max_lc_iter, max_gc_skip;
	int32_t min_lc_cnt, min_lc_score;
	int32_t min_gc_cnt, min_gc_score;
	int32_t gdp_max_ed, lc_max_trim, lc_max_occ;
	float mask_level;
	int32_t sub_diff, best_n;
	float pri_ratio;
	int32_t ref_bonus;
	// glibc logf() of small integers, tabulated by the host so that mapq is bit-exact (reference: gcmisc.c:216-217)
	const float *logf_tab;
	int32_t n_logf_tab;
};

static const uint64_t F_SPLICE = 0x10, F_SR = 0x20, F_HEAP_SORT = 0x400, F_RMQ = 0x8000, F_NO_DIAG = 0x400000, F_CIGAR = 0x4000000;

// linear chain (reference: minigraph.h:100-106 mg_lchain_t)
struct LChain {
	int32_t off, cnt;
	uint32_t v;
	int32_t rs, re, qs, qe;
	int32_t score, dist_pre;
	uint32_t hash_pre;
	int32_t inner_pre;
};

// per-read bookkeeping that travels between the stage kernels
struct ReadMeta {
	int32_t status;
	int32_t n_mz;        // number of minimizers (mv.n)
	int32_t rep_len;
	int32_t n_a;         // seeds (after k_seed) / anchors in chains (after k_chain)
	int64_t a_off;       // element offset into the anchor pool
	int32_t n_mp;        // n_mini_pos
	int64_t mp_off;      // element offset into the mini_pos pool
	int32_t n_lc;
	int64_t lc_off;      // element offset into the lchain pool
	uint32_t hash;       // per-read hash (reference: map-algo.c:362-364)
	int32_t n_seed0;     // seeds before chaining (for the byte model)
	int32_t n_u0;        // chains out of the chaining DP (for the byte model)
	uint64_t arena_peak;
	int64_t gstate_off;  // byte offset of the GState blob (between k_gchain and k_gchain_gen)
};

struct BatchDev {
	int32_t n_reads;
	const char *seq;            // concatenated upper-case read sequences
	const uint64_t *seq_off;    // [n_reads]
	const int32_t *seq_len;     // [n_reads]
	const uint32_t *name_hash;  // [n_reads] kh_hash_str(qname) computed by the host, 0 if no name
	const int32_t *self_id;     // [n_reads] MG_M_NO_DIAG only: GraphDev::seg_name_id value the read's name equals, or -1 (NULL otherwise)
	// multi-segment fragments (paired reads; reference: map-algo.c:34-45,356-360): a read's sequence is the concatenation of its
	// segments, seg_len[seg_off[r] .. seg_off[r+1]) their lengths.  NULL for the usual batch of single-segment reads.
	const int32_t *seg_off;     // [n_reads + 1]
	const int32_t *seg_len;
	// the reads as they cross PCIe: 2 bits per base (A/C/G/T only; k_unpack writes the ASCII copy above, the sketch reads the codes).
	// NULL when the batch was uploaded as ASCII (reads with other letters in it)
	const uint64_t *pk;         // 32 bases per word; a read's words start at pk_off[r]
	const uint64_t *pk_off;     // [n_reads] word offsets
};
MG_HD inline int32_t batch_n_seg(const BatchDev &b, int rid) { return b.seg_off? b.seg_off[rid + 1] - b.seg_off[rid] : 1; }

} // namespace mgb
