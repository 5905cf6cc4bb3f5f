// mgb_simlanes.h -- TEST INFRASTRUCTURE ONLY: a 32-lane warp on the CPU, for tests/hostsim/libmgb_hostsim32.so.
//
// The device code is written as warp-uniform functions (mgb_common.cuh): all lanes enter together, exchange values only
// through the warp_* helpers and separate memory phases with warp_sync().  That makes a warp easy to run on one CPU
// thread: every lane is a fibre (ucontext), a lane runs until it reaches a helper, parks its contribution and yields, and
// the last lane to arrive completes the exchange.  The simulator therefore executes the very ballots, prefix scans and
// order-preserving compactions the GPU executes, with 32 lanes, and it stops loudly when lanes do not meet at the same
// helper (a divergence bug that would hang or corrupt on the device).  It says nothing about memory races between two
// synchronisation points: lanes never run concurrently here.
#pragma once
#include <ucontext.h>
#include <execinfo.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>
#include <functional>
#include <vector>

namespace mgb { namespace sim {

enum { MAX_LANES = 128, STACK_BYTES = 1 << 20 }; // 32 lanes for a warp, 128 for the block-uniform code (mgb_cta.cuh)

// A fibre context.  glibc's swapcontext() makes a system call per switch (it saves the signal mask); on x86-64 a lane switch
// is instead six pushes, a stack swap and six pops -- the callee-saved registers of the SysV ABI -- about fifty times cheaper.
#if defined(__x86_64__) && defined(__GNUC__)
struct Ctx { void *sp; };
__attribute__((naked, noinline)) inline void ctx_switch_raw(void **, void *)
{
	__asm__ volatile(
		"pushq %rbp\n\tpushq %rbx\n\tpushq %r12\n\tpushq %r13\n\tpushq %r14\n\tpushq %r15\n\t"
		"movq %rsp, (%rdi)\n\t"
		"movq %rsi, %rsp\n\t"
		"popq %r15\n\tpopq %r14\n\tpopq %r13\n\tpopq %r12\n\tpopq %rbx\n\tpopq %rbp\n\t"
		"ret\n");
}
inline void ctx_switch(Ctx &from, Ctx &to) { ctx_switch_raw(&from.sp, to.sp); }
inline void ctx_make(Ctx &c, Ctx &, char *stack, size_t bytes, void (*entry)())
{
	uintptr_t top = ((uintptr_t)stack + bytes) & ~(uintptr_t)15;
	void **p = (void**)(top - 16); // after the `ret` below rsp is top-8: what a function sees on entry
	*p = (void*)entry;
	for (int i = 0; i < 6; ++i) *--p = 0;
	c.sp = (void*)p;
}
#else
struct Ctx { ucontext_t uc; };
inline void ctx_switch(Ctx &from, Ctx &to) { swapcontext(&from.uc, &to.uc); }
inline void ctx_make(Ctx &c, Ctx &back, char *stack, size_t bytes, void (*entry)())
{
	getcontext(&c.uc);
	c.uc.uc_stack.ss_sp = stack, c.uc.uc_stack.ss_size = bytes, c.uc.uc_link = &back.uc;
	makecontext(&c.uc, entry, 0);
}
#endif

struct Warp {
	int n = 0, cur = 0;
	Ctx sched, ctx[MAX_LANES];
	char *stack[MAX_LANES];
	bool done[MAX_LANES];
	uint64_t slot[2][MAX_LANES];
	int kind[2][MAX_LANES];       // which helper each lane entered the exchange from: lanes of a warp-uniform program use the same one
	long gen[MAX_LANES];          // exchanges entered so far, per lane
	long wait_gen[MAX_LANES];     // >= 0: the lane is parked in that exchange and need not be resumed before it is complete
	long slot_gen[2] = {-1, -1};  // the exchange a buffer currently belongs to
	int count[2] = {0, 0};        // lanes that have contributed to it
	long ready[2] = {-1, -1};     // set to the exchange number once every lane has contributed
	unsigned long progress = 0;
	int dumper = -1;              // lane that asked for the dump (resumes after each printing lane), -1: the scheduler
	bool dump = false;            // set when the warp is stuck: the waiting lanes print where they are
	const std::function<void(int)> *fn = 0;
};

inline Warp *&current() { static thread_local Warp *w = 0; return w; }
inline int *tag() { static thread_local int t[2] = {-1, -1}; return t; } // what the warp is working on (stage, item), for the messages
inline int lane() { Warp *w = current(); return w? w->cur : 0; }

inline void trampoline()
{
	Warp *w = current();
	const int me = w->cur;
	(*w->fn)(me);
	w->done[me] = true, ++w->progress;
	ctx_switch(w->ctx[me], w->sched); // never resumed
}

inline void on_segv(int)
{
	Warp *w = current();
	void *bt[32];
	int nb = backtrace(bt, 32);
	fprintf(stderr, "[mgb::sim] SIGSEGV in stage %d item %d lane %d\n", tag()[0], tag()[1], w? w->cur : -1);
	backtrace_symbols_fd(bt, nb, 2);
	_exit(139);
}
inline void install_segv_handler()
{
	static bool done = false;
	if (done) return;
	done = true;
	static char alt[1 << 16];
	stack_t ss;
	ss.ss_sp = alt, ss.ss_size = sizeof(alt), ss.ss_flags = 0;
	sigaltstack(&ss, 0);
	struct sigaction sa;
	memset(&sa, 0, sizeof(sa));
	sa.sa_handler = on_segv, sa.sa_flags = SA_ONSTACK;
	sigaction(SIGSEGV, &sa, 0);
}

// Run fn(lane) for lanes 0..n-1 as one warp.
inline void run_warp(int n, const std::function<void(int)> &fn)
{
	if (getenv("MGB_SIM_SEGV_TRACE")) install_segv_handler();
	static thread_local unsigned long long seed = getenv("MGB_SIM_SEED")? strtoull(getenv("MGB_SIM_SEED"), 0, 10) : 0;
	Warp w;
	w.n = n, w.fn = &fn;
	Warp *outer = current();
	current() = &w;
	for (int l = 0; l < n; ++l) {
		w.done[l] = false, w.gen[l] = 0, w.wait_gen[l] = -1;
		w.stack[l] = (char*)malloc(STACK_BYTES); // not cleared: only the pages a lane touches are ever mapped
		ctx_make(w.ctx[l], w.sched, w.stack[l], STACK_BYTES, trampoline);
	}
	for (;;) { // round robin; a full round without any lane moving on means the lanes wait for different things
		int alive = 0;
		const unsigned long before = w.progress;
		int order[MAX_LANES];
		for (int l = 0; l < n; ++l) order[l] = l;
		if (seed) // MGB_SIM_SEED=<n>: lanes are resumed in a different random order every round (other interleavings, same results expected)
			for (int l = n - 1; l > 0; --l) { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; int r = (int)((seed >> 33) % (unsigned)(l + 1)); int t = order[l]; order[l] = order[r], order[r] = t; }
		for (int q = 0; q < n; ++q) {
			const int l = order[q];
			if (w.done[l]) continue;
			++alive;
			if (w.wait_gen[l] >= 0 && w.ready[w.wait_gen[l] & 1] != w.wait_gen[l]) continue; // still waiting for the others
			w.cur = l;
			ctx_switch(w.sched, w.ctx[l]);
		}
		if (alive == 0) break;
		if (w.progress == before) {
			fprintf(stderr, "[mgb::sim] stage %d item %d: warp stuck, lanes did not meet at the same warp_* helper (exchange counts:", tag()[0], tag()[1]);
			for (int l = 0; l < n; ++l) fprintf(stderr, " %ld%s", w.gen[l], w.done[l]? "x" : "");
			fprintf(stderr, ")\n");
			w.dump = true;
			for (int l = 0; l < n; ++l) if (!w.done[l]) { w.cur = l; ctx_switch(w.sched, w.ctx[l]); }
			abort();
		}
	}
	for (int l = 0; l < n; ++l) free(w.stack[l]);
	current() = outer;
}

// Every lane contributes v and receives the contributions of all lanes; out[] has room for `cap` of them.
inline void exchange(uint64_t v, uint64_t *out, int kind, int cap = 32)
{
	Warp *w = current();
	if (w == 0) { fprintf(stderr, "[mgb::sim] a warp helper was called outside run_warp()\n"); abort(); }
	if (w->n > cap) { fprintf(stderr, "[mgb::sim] a %d-lane helper (kind %d) was called by a group of %d lanes\n", cap, kind, w->n); abort(); }
	const int me = w->cur;
	const long g = w->gen[me]++;
	const int b = (int)(g & 1);
	if (w->slot_gen[b] != g) w->slot_gen[b] = g, w->count[b] = 0;
	w->slot[b][me] = v;
	w->kind[b][me] = kind;
	++w->progress;
	if (++w->count[b] == w->n) {
		for (int l = 1; l < w->n; ++l)
			if (w->kind[b][l] != w->kind[b][0]) {
				fprintf(stderr, "[mgb::sim] stage %d item %d: lanes 0 and %d meet in different warp_* helpers (kinds %d and %d, exchange %ld)\n", tag()[0], tag()[1], l, w->kind[b][0], w->kind[b][l], g);
				{ void *bt[24]; int nb = backtrace(bt, 24); fprintf(stderr, "[mgb::sim] lane %d is at:\n", me); backtrace_symbols_fd(bt, nb, 2); }
				w->dump = true, w->dumper = me;
				for (int k = 0; k < w->n; ++k) if (k != me && (k == 0 || k == l)) { w->cur = k; ctx_switch(w->ctx[me], w->ctx[k]); }
				abort();
			}
		w->ready[b] = g;
	}
	while (w->ready[b] != g) {
		w->wait_gen[me] = g;
		ctx_switch(w->ctx[me], w->sched);
		w->wait_gen[me] = -1;
		if (w->dump) { void *bt[24]; int nb = backtrace(bt, 24); fprintf(stderr, "[mgb::sim] lane %d waits at:\n", me); backtrace_symbols_fd(bt, nb, 2); ctx_switch(w->ctx[me], w->dumper >= 0? w->ctx[w->dumper] : w->sched); }
	}
	memcpy(out, w->slot[b], sizeof(uint64_t) * (size_t)w->n);
}

} } // namespace mgb::sim
