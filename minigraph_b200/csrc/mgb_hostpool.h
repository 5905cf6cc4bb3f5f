// mgb_hostpool.h -- a small persistent pool of host threads.  The batch path fans work out to host threads four times per
// batch (packing, result assembly, GAF formatting, GAF copy); spawning ~50 threads each time costs more than a millisecond
// per fan-out, so the workers are kept.  One pool serves one caller at a time (run() holds the pool's mutex); the engine
// and the GAF writer own one pool each, because bench.py drives them from two threads at once.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <cstdint>

namespace mgb {

class HostPool {
public:
	~HostPool()
	{
		{ std::lock_guard<std::mutex> lk(m_); quit_ = true; }
		cv_.notify_all();
		for (auto &t : th_) t.join();
	}
	// fn(i) for i in [0, n), cut into nt contiguous ranges (the caller's thread takes the first one)
	void run(int64_t n, int nt, const std::function<void(int64_t)> &fn)
	{
		if (n <= 0) return;
		if (nt > n) nt = (int)n;
		if (nt <= 1) { for (int64_t i = 0; i < n; ++i) fn(i); return; }
		std::lock_guard<std::mutex> one_caller(run_m_);
		{
			std::unique_lock<std::mutex> lk(m_);
			while ((int)th_.size() < nt - 1) { int id = (int)th_.size(); th_.emplace_back([this, id]() { worker(id); }); seen_.push_back(gen_); }
			fn_ = &fn, n_ = n, nt_ = nt, pending_ = nt - 1, ++gen_;
		}
		cv_.notify_all();
		range(0);
		std::unique_lock<std::mutex> lk(m_);
		done_.wait(lk, [this]() { return pending_ == 0; });
		fn_ = 0;
	}
private:
	void range(int t)
	{
		const int64_t chunk = (n_ + nt_ - 1) / nt_, b = t * chunk, e = b + chunk < n_? b + chunk : n_;
		for (int64_t i = b; i < e; ++i) (*fn_)(i);
	}
	void worker(int id)
	{
		std::unique_lock<std::mutex> lk(m_);
		for (;;) {
			cv_.wait(lk, [this, id]() { return quit_ || seen_[id] != gen_; });
			if (quit_) return;
			seen_[id] = gen_;
			if (id + 1 < nt_) { // this round uses workers 0 .. nt_-2 (ranges 1 .. nt_-1)
				lk.unlock();
				range(id + 1);
				lk.lock();
				if (--pending_ == 0) done_.notify_one();
			}
		}
	}
	std::mutex m_, run_m_;
	std::condition_variable cv_, done_;
	std::vector<std::thread> th_;
	std::vector<uint64_t> seen_;
	const std::function<void(int64_t)> *fn_ = 0;
	int64_t n_ = 0;
	int nt_ = 0, pending_ = 0;
	uint64_t gen_ = 0;
	bool quit_ = false;
};

} // namespace mgb
