// mgb_cta.cuh -- block-uniform helpers: the counterpart of the warp_* helpers of mgb_common.cuh for code that is entered by
// ALL threads of a thread block with identical arguments (one work item per block instead of one per warp).
//
// Same rules as for warp-uniform code: scalars are replicated on every thread, loops are strided by MGB_CTA_T, threads
// exchange values only through the helpers below, and a helper that votes is also the barrier that separates two memory
// phases.  Three implementations: the device (barriers + a few words of shared memory), the fibre simulator of the CPU
// tests (MGB_SIM_LANES: a "warp" of MGB_CTA_T fibres, mgb_simlanes.h) and the single-lane CPU build (identities).
#pragma once
#include "mgb_common.cuh"

namespace mgb {

struct CtaScratch {      // lives in shared memory, one per block
	uint32_t bits[3];    // cta_or_bits: three slots used in turn, so that one barrier per vote is enough
	int32_t wmin[32];    // cta_min_i32: one partial result per warp
	uint64_t bc;         // cta_bcast_*
};
struct CtaCtx {          // replicated per thread
	CtaScratch *s;
	uint32_t k;          // votes taken so far
};

#define MGB_CTA_THREADS 128 // block size of the kernels built on these helpers (the same number in the host and the device pass of nvcc)
#if MGB_ON_DEVICE
#define MGB_CTA_T MGB_CTA_THREADS
MG_D inline void cta_init(CtaCtx &x, CtaScratch *s, int tid)
{
	x.s = s, x.k = 0;
	if (tid == 0) s->bits[0] = s->bits[1] = s->bits[2] = 0;
	__syncthreads();
}
MG_D inline void cta_sync() { __syncthreads(); }
// bitwise OR of v over the block; also a barrier.  Slot k%3 collects vote k; thread 0 clears the slot of vote k+1 on its way
// into vote k: that slot was last read in vote k-2, and every thread has left vote k-2 before any thread can pass the
// barrier of vote k-1.
MG_D inline uint32_t cta_or_bits(CtaCtx &x, uint32_t v, int tid)
{
	const uint32_t i = x.k % 3u, nx = i == 2u? 0u : i + 1u;
	++x.k;
	const uint32_t w = __reduce_or_sync(0xffffffffu, v);
	if (tid == 0) x.s->bits[nx] = 0;
	if ((tid & 31) == 0 && w) atomicOr(&x.s->bits[i], w);
	__syncthreads();
	return x.s->bits[i];
}
MG_D inline int32_t cta_min_i32(CtaCtx &x, int32_t v, int tid)
{
	const int32_t m = warp_min_i32(v);
	if ((tid & 31) == 0) x.s->wmin[tid >> 5] = m;
	__syncthreads();
	int32_t r = x.s->wmin[0];
	for (int i = 1; i < MGB_CTA_T / 32; ++i) r = x.s->wmin[i] < r? x.s->wmin[i] : r;
	__syncthreads();
	return r;
}
MG_D inline uint64_t cta_bcast_u64(CtaCtx &x, uint64_t v, int tid) // the value of thread 0
{
	if (tid == 0) x.s->bc = v;
	__syncthreads();
	const uint64_t r = x.s->bc;
	__syncthreads();
	return r;
}
#elif defined(MGB_SIM_LANES)
#define MGB_CTA_T 128
inline void cta_init(CtaCtx &x, CtaScratch *s, int) { x.s = s, x.k = 0; }
inline void cta_sync() { uint64_t o[MGB_CTA_T]; sim::exchange(0, o, 20, MGB_CTA_T); }
inline uint32_t cta_or_bits(CtaCtx &x, uint32_t v, int) { ++x.k; uint64_t o[MGB_CTA_T]; sim::exchange(v, o, 21, MGB_CTA_T); uint32_t r = 0; for (int i = 0; i < MGB_CTA_T; ++i) r |= (uint32_t)o[i]; return r; }
inline int32_t cta_min_i32(CtaCtx &, int32_t v, int) { uint64_t o[MGB_CTA_T]; sim::exchange((uint64_t)(uint32_t)v, o, 22, MGB_CTA_T); int32_t r = v; for (int i = 0; i < MGB_CTA_T; ++i) if ((int32_t)(uint32_t)o[i] < r) r = (int32_t)(uint32_t)o[i]; return r; }
inline uint64_t cta_bcast_u64(CtaCtx &, uint64_t v, int) { uint64_t o[MGB_CTA_T]; sim::exchange(v, o, 23, MGB_CTA_T); return o[0]; }
#else
#define MGB_CTA_T 1
inline void cta_init(CtaCtx &x, CtaScratch *s, int) { x.s = s, x.k = 0; }
inline void cta_sync() {}
inline uint32_t cta_or_bits(CtaCtx &x, uint32_t v, int) { ++x.k; return v; }
inline int32_t cta_min_i32(CtaCtx &, int32_t v, int) { return v; }
inline uint64_t cta_bcast_u64(CtaCtx &, uint64_t v, int) { return v; }
#endif
MG_HD inline int32_t cta_bcast_i32(CtaCtx &x, int32_t v, int tid) { return (int32_t)(uint32_t)cta_bcast_u64(x, (uint64_t)(uint32_t)v, tid); }

} // namespace mgb
