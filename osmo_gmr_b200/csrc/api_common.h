// api_common.h - plumbing shared by the C-ABI translation units: error reporting, lazy device
// init, and the Stage helper that lets every entry point take host OR device pointers.
#pragma once
#include <stdlib.h>
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

// Per-thread staging arena for callers that pass HOST pointers only and say so (gmr1b200_host_hint, set by the n = 1
// compat wrappers of compat.c): one page-locked host buffer and one device buffer of the same size, kept for the life
// of the thread.  A call then costs one H2D copy per input, the kernel launches, ONE D2H copy for all outputs and one
// synchronisation - no cudaPointerGetAttributes, no cudaMallocAsync / cudaFreeAsync, no pageable-memory staging by the
// driver.  (The link-level drop-in spent ~1 ms per n = 1 call on those; profiles/README.md has the breakdown.)
struct HostArena {
	char  *h = nullptr, *d = nullptr;
	size_t cap = 0, want = 0;
	bool   busy = false;         // one Stage at a time (a nested Stage stages the ordinary way)
};
extern thread_local int g_host_hint;
HostArena &host_arena();          // this thread's arena for the current device (allocated / grown between calls)

// true when this thread can dereference p (plain or pinned host memory); device and managed pointers: false
inline bool host_pointer(const void *p)
{
	if (g_host_hint)
		return true;
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
		cudaGetLastError();
		return true;
	}
	return at.type == cudaMemoryTypeUnregistered || at.type == cudaMemoryTypeHost;
}

// recordings [rec_ofs[i], rec_ofs[i] + rec_len[i]) must lie inside iq_len samples: checked when the host can read the
// descriptors (device-resident descriptors are the caller's responsibility; the kernels bound every window by rec_len)
inline bool recordings_in_range(const int64_t *rec_ofs, const int32_t *rec_len, int n, int64_t iq_len)
{
	if (!host_pointer(rec_ofs) || !host_pointer(rec_len))
		return true;
	for (int i = 0; i < n; i++)
		if (rec_ofs[i] < 0 || rec_len[i] < 0 || rec_ofs[i] + rec_len[i] > iq_len)
			return false;
	return true;
}

// Stage: resolves each pointer argument of a batched call to a device pointer.
//   in(p, bytes)   host  -> device scratch + H2D copy on the stream;   device -> p itself
//   out(p, bytes)  host  -> device scratch, D2H copy queued for finish(); device -> p itself
//   finish()       runs the queued D2H copies, synchronises the stream iff any host buffer was
//                  involved, releases scratch (stream-ordered)
class Stage {
public:
	explicit Stage(void *stream) : st_((cudaStream_t)stream)
	{
		if (g_host_hint) {
			ar_ = &host_arena();
			if (!ar_->h || ar_->busy)
				ar_ = nullptr;
			else
				ar_->busy = true;
		}
	}
	~Stage()
	{
		release();
		if (ar_)
			ar_->busy = false;
	}

	template <class T> const T *in(const T *p, size_t count)
	{
		if (!p || !count || failed_)
			return p;
		if (is_device(p))
			return p;
		const size_t bytes = count * sizeof(T);
		if (char *a = arena(bytes)) {
			memcpy(ar_->h + (a - ar_->d), p, bytes);
			check(cudaMemcpyAsync(a, ar_->h + (a - ar_->d), bytes, cudaMemcpyHostToDevice, st_), "H2D copy");
			return (const T *)a;
		}
		void *d = scratch(bytes);
		if (!d)
			return nullptr;
		check(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, st_), "H2D copy");
		return (const T *)d;
	}

	template <class T> T *out(T *p, size_t count)
	{
		if (!p || !count || failed_)
			return p;
		if (is_device(p))
			return p;
		const size_t bytes = count * sizeof(T);
		if (char *a = arena(bytes)) {
			const size_t off = (size_t)(a - ar_->d);
			out_lo_ = off < out_lo_ ? off : out_lo_;
			out_hi_ = off + bytes > out_hi_ ? off + bytes : out_hi_;
			backs_.push_back({p, a, bytes, true});
			return (T *)a;
		}
		void *d = scratch(bytes);
		if (!d)
			return nullptr;
		backs_.push_back({p, d, bytes, false});
		return (T *)d;
	}

	// device-only scratch that lives until finish()
	template <class T> T *tmp(size_t count)
	{
		if (failed_)
			return nullptr;
		if (char *a = arena(count * sizeof(T)))
			return (T *)a;
		return (T *)scratch(count * sizeof(T));
	}

	bool failed() const { return failed_; }
	int rc() const { return rc_; }

	int finish(cudaError_t launch_err, const char *what)
	{
		if (launch_err != cudaSuccess && !failed_) {
			failed_ = true;
			rc_ = cuda_rc(launch_err, what);
		}
		if (!failed_) {
			if (out_hi_ > out_lo_)       // every arena output in one copy
				check(cudaMemcpyAsync(ar_->h + out_lo_, ar_->d + out_lo_, out_hi_ - out_lo_, cudaMemcpyDeviceToHost, st_),
				      "D2H copy");
			for (auto &b : backs_)
				if (!b.arena)
					check(cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, st_), "D2H copy");
		}
		if (host_involved_ || failed_) {
			cudaError_t e = cudaStreamSynchronize(st_);
			if (e != cudaSuccess && !failed_) {
				failed_ = true;
				rc_ = cuda_rc(e, what);
			}
		}
		if (!failed_)
			for (auto &b : backs_)
				if (b.arena)
					memcpy(b.host, ar_->h + ((char *)b.dev - ar_->d), b.bytes);
		release();
		return rc_;
	}

private:
	struct Back { void *host, *dev; size_t bytes; bool arena; };
	cudaStream_t st_;
	std::vector<void *> tmp_;
	std::vector<Back> backs_;
	bool host_involved_ = false, failed_ = false;
	int rc_ = 0;
	HostArena *ar_ = nullptr;
	size_t cur_ = 0, out_lo_ = (size_t)-1, out_hi_ = 0;

	// bump allocation in the thread's arena (256-byte granules); NULL: not in arena mode or the arena is full - the
	// arena then grows before the thread's next call
	char *arena(size_t bytes)
	{
		if (!ar_)
			return nullptr;
		const size_t need = (bytes + 255) & ~(size_t)255;
		ar_->want = ar_->want > cur_ + need ? ar_->want : cur_ + need;
		if (cur_ + need > ar_->cap)
			return nullptr;
		char *p = ar_->d + cur_;
		cur_ += need;
		return p;
	}
	void check(cudaError_t e, const char *what)
	{
		if (e != cudaSuccess && !failed_) {
			failed_ = true;
			rc_ = cuda_rc(e, what);
		}
	}
	bool is_device(const void *p)
	{
		if (g_host_hint) {              // the caller vouches: host memory
			host_involved_ = true;
			return false;
		}
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
				// no reuse of a block whose free is still pending on ANOTHER stream: the runtime would make the
				// allocating stream wait for that stream's work, which serialises callers that keep several
				// streams busy (two wideband recordings in flight ran at half speed that way); the pool grows by
				// a block per busy stream instead.  GMR1B200_POOL_INTERNAL_DEPS=1 restores the default.
				const char *ev = getenv("GMR1B200_POOL_INTERNAL_DEPS");
				int dep = ev && atoi(ev) != 0 ? 1 : 0;
				cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &dep);
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
