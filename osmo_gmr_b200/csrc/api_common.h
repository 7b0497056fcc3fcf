// api_common.h - plumbing shared by the C-ABI translation units: error reporting, lazy device
// init, and the Stage helper that lets every entry point take host OR device pointers.
#pragma once
#include <cuda_runtime.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <vector>

namespace gmr1 {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int set_err(int rc, const char *what, cudaError_t e = cudaSuccess)
{
	if (e != cudaSuccess)
		snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
	else
		snprintf(g_err, sizeof(g_err), "%s", what);
	return rc;
}

inline int cuda_rc(cudaError_t e, const char *what)
{
	if (e == cudaSuccess)
		return 0;
	if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
		return set_err(-ENODEV, what, e);
	if (e == cudaErrorMemoryAllocation)
		return set_err(-ENOMEM, what, e);
	return set_err(-EIO, what, e);
}

// Stage: resolves each pointer argument of a batched call to a device pointer.
//   in(p, bytes)   host  -> device scratch + H2D copy on the stream;   device -> p itself
//   out(p, bytes)  host  -> device scratch, D2H copy queued for finish(); device -> p itself
//   finish()       runs the queued D2H copies, synchronises the stream iff any host buffer was
//                  involved, releases scratch (stream-ordered)
class Stage {
public:
	explicit Stage(void *stream) : st_((cudaStream_t)stream) {}
	~Stage() { release(); }

	template <class T> const T *in(const T *p, size_t count)
	{
		if (!p || !count || failed_)
			return p;
		if (is_device(p))
			return p;
		void *d = scratch(count * sizeof(T));
		if (!d)
			return nullptr;
		check(cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, st_), "H2D copy");
		return (const T *)d;
	}

	template <class T> T *out(T *p, size_t count)
	{
		if (!p || !count || failed_)
			return p;
		if (is_device(p))
			return p;
		void *d = scratch(count * sizeof(T));
		if (!d)
			return nullptr;
		backs_.push_back({p, d, count * sizeof(T)});
		return (T *)d;
	}

	// device-only scratch that lives until finish()
	template <class T> T *tmp(size_t count) { return failed_ ? nullptr : (T *)scratch(count * sizeof(T)); }

	bool failed() const { return failed_; }
	int rc() const { return rc_; }

	int finish(cudaError_t launch_err, const char *what)
	{
		if (launch_err != cudaSuccess && !failed_) {
			failed_ = true;
			rc_ = cuda_rc(launch_err, what);
		}
		if (!failed_)
			for (auto &b : backs_)
				check(cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, st_), "D2H copy");
		if (host_involved_ || failed_) {
			cudaError_t e = cudaStreamSynchronize(st_);
			if (e != cudaSuccess && !failed_) {
				failed_ = true;
				rc_ = cuda_rc(e, what);
			}
		}
		release();
		return rc_;
	}

private:
	struct Back { void *host, *dev; size_t bytes; };
	cudaStream_t st_;
	std::vector<void *> tmp_;
	std::vector<Back> backs_;
	bool host_involved_ = false, failed_ = false;
	int rc_ = 0;

	void check(cudaError_t e, const char *what)
	{
		if (e != cudaSuccess && !failed_) {
			failed_ = true;
			rc_ = cuda_rc(e, what);
		}
	}
	bool is_device(const void *p)
	{
		cudaPointerAttributes at;
		cudaError_t e = cudaPointerGetAttributes(&at, p);
		if (e != cudaSuccess) {
			cudaGetLastError();
			host_involved_ = true;
			return false;
		}
		if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)
			return true;
		host_involved_ = true;
		return false;
	}
	void *scratch(size_t bytes)
	{
		// keep freed scratch in the stream-ordered pool instead of returning it to the driver at
		// every synchronisation (the default release threshold is 0)
		static std::atomic<bool> pool_set[64];
		int dev = 0;
		if (cudaGetDevice(&dev) == cudaSuccess && dev < 64 && !pool_set[dev].exchange(true)) {
			cudaMemPool_t pool;
			if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
				uint64_t thr = UINT64_MAX;
				cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
			}
		}
		void *d = nullptr;
		cudaError_t e = cudaMallocAsync(&d, bytes, st_);
		if (e != cudaSuccess) {
			check(e, "device scratch allocation");
			return nullptr;
		}
		tmp_.push_back(d);
		return d;
	}
	void release()
	{
		for (void *d : tmp_)
			cudaFreeAsync(d, st_);
		tmp_.clear();
		backs_.clear();
	}
};

}  // namespace gmr1
