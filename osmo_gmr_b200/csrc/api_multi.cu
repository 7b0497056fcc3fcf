// api_multi.cu - C ABI: several GPUs of one box from ONE process (include/gmr1_b200.h, "device pool").
//
// north_star: "ARFCNs are independent, so channels are partitioned across the 8 GPUs of one box with per-GPU
// streams and a host-side result gather, and NCCL is not needed."  The reference walks its channels one after the
// other (src/gmr1_rx.c:732-741, one chan_desc per ARFCN); here ARFCN a goes to device a mod G, every device has one
// feeder thread (bound to the CPUs next to the GPU when sysfs says which they are), a few streams and a ring of
// device buffers, and the host IQ of its ARFCNs travels in chunks: strided H2D copy (cudaMemcpy2DAsync picks every
// G-th ARFCN) -> the same batched entry points a single-GPU caller uses, on device pointers -> strided D2H copy of
// the results straight into the caller's arrays.  No collective, no peer traffic.
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gmr1_b200.h"
#include "api_common.h"
#include "gmr1_tables.h"

using namespace gmr1;

namespace {

struct Slot {                 // one stream + its device buffers
	cudaStream_t st = nullptr;
	float *iq = nullptr;      // chunk of windows
	uint8_t *l2 = nullptr;
	int32_t *i32 = nullptr;   // crc | conv (xcch);  rough | align (fcch)
	float *f32 = nullptr;     // toa (xcch); ferr (fcch)
};

struct Dev {
	int id = 0;
	std::vector<Slot> slots;
	std::vector<int> cpus;    // CPUs local to the GPU (empty: unknown)
};

}  // namespace

struct gmr1b200_pool {
	std::vector<Dev> devs;
	int64_t chunk_bytes = 0;  // IQ bytes per slot
	int64_t chunk_units = 0;  // result rows per slot
};

static void local_cpus(int dev, std::vector<int> &out)
{
	char bus[32] = {0};
	if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess)
		return;
	for (char *c = bus; *c; c++)
		*c = (char)tolower(*c);
	std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
	FILE *f = fopen(path.c_str(), "r");
	if (!f)
		return;
	char line[4096] = {0};
	if (fgets(line, sizeof(line), f)) {
		for (char *tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
			int a = 0, b = 0;
			if (sscanf(tok, "%d-%d", &a, &b) == 2) {
				for (int i = a; i <= b; i++)
					out.push_back(i);
			} else if (sscanf(tok, "%d", &a) == 1)
				out.push_back(a);
		}
	}
	fclose(f);
}

static void bind_thread(const std::vector<int> &cpus)
{
	if (cpus.empty())
		return;
	cpu_set_t set;
	CPU_ZERO(&set);
	for (int c : cpus)
		if (c >= 0 && c < CPU_SETSIZE)
			CPU_SET(c, &set);
	pthread_setaffinity_np(pthread_self(), sizeof(set), &set);      // best effort
}

// pool_create / pool_free visit every device from the CALLER's thread: its current device is put back on the way out
struct DeviceGuard {
	int dev = -1;
	DeviceGuard() { if (cudaGetDevice(&dev) != cudaSuccess) { dev = -1; cudaGetLastError(); } }
	~DeviceGuard() { if (dev >= 0) cudaSetDevice(dev); }
};

static void pool_free(gmr1b200_pool *p)
{
	DeviceGuard keep;
	for (Dev &d : p->devs) {
		cudaSetDevice(d.id);
		for (Slot &s : d.slots) {
			if (s.st) cudaStreamSynchronize(s.st);
			cudaFree(s.iq);
			cudaFree(s.l2);
			cudaFree(s.i32);
			cudaFree(s.f32);
			if (s.st) cudaStreamDestroy(s.st);
		}
	}
	delete p;
}

// run fn(dev, dev_index) on one thread per device; first non-zero return code wins
template <class F>
static int per_device(gmr1b200_pool *p, F fn)
{
	const int G = (int)p->devs.size();
	std::vector<int> rc(G, 0);
	std::vector<std::string> msg(G);
	std::vector<std::thread> th;
	for (int g = 0; g < G; g++)
		th.emplace_back([&, g] {
			bind_thread(p->devs[g].cpus);
			cudaError_t e = cudaSetDevice(p->devs[g].id);
			if (e != cudaSuccess) {
				rc[g] = cuda_rc(e, "pool: cudaSetDevice");
			} else {
				rc[g] = fn(p->devs[g], g);
				for (Slot &s : p->devs[g].slots) {
					e = cudaStreamSynchronize(s.st);
					if (e != cudaSuccess && !rc[g])
						rc[g] = cuda_rc(e, "pool: stream");
				}
			}
			if (rc[g])
				msg[g] = g_err;            // thread-local error text of the worker
		});
	for (auto &t : th)
		t.join();
	for (int g = 0; g < G; g++)
		if (rc[g])
			return set_err(rc[g], msg[g].c_str());
	return 0;
}

extern "C" {

void *gmr1b200_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

void gmr1b200_host_free(void *p)
{
	if (p)
		cudaFreeHost(p);
}

int gmr1b200_pool_create(const int *devices, int n_dev, int streams_per_dev, int64_t chunk_bytes,
                         struct gmr1b200_pool **out)
{
	if (!out || n_dev < 1 || n_dev > 64 || streams_per_dev < 1 || streams_per_dev > 8 || chunk_bytes < (1 << 20))
		return set_err(-EINVAL, "pool_create: bad argument");
	int have = 0;
	cudaError_t e = cudaGetDeviceCount(&have);
	if (e != cudaSuccess)
		return cuda_rc(e, "pool_create: cudaGetDeviceCount");
	DeviceGuard keep;
	gmr1b200_pool *p = new gmr1b200_pool;
	p->chunk_bytes = chunk_bytes;
	p->chunk_units = chunk_bytes / 1024;             // >= one result row per KB of IQ (the shortest window is 3.8 KB)
	for (int g = 0; g < n_dev; g++) {
		Dev d;
		d.id = devices ? devices[g] : g;
		if (d.id < 0 || d.id >= have) {
			pool_free(p);
			return set_err(-ENODEV, "pool_create: no such device");
		}
		local_cpus(d.id, d.cpus);
		if ((e = cudaSetDevice(d.id)) != cudaSuccess) {
			pool_free(p);
			return cuda_rc(e, "pool_create: cudaSetDevice");
		}
		d.slots.resize(streams_per_dev);
		p->devs.push_back(d);
		for (Slot &s : p->devs.back().slots) {
			e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking);
			if (e == cudaSuccess) e = cudaMalloc(&s.iq, (size_t)chunk_bytes);
			if (e == cudaSuccess) e = cudaMalloc(&s.l2, (size_t)p->chunk_units * 24);
			if (e == cudaSuccess) e = cudaMalloc(&s.i32, (size_t)p->chunk_units * 8);
			if (e == cudaSuccess) e = cudaMalloc(&s.f32, (size_t)p->chunk_units * 4);
			if (e != cudaSuccess) {
				pool_free(p);
				return cuda_rc(e, "pool_create: allocation");
			}
		}
	}
	*out = p;
	return 0;
}

void gmr1b200_pool_destroy(struct gmr1b200_pool *p)
{
	if (p)
		pool_free(p);
}

int gmr1b200_pool_size(const struct gmr1b200_pool *p)
{
	return p ? (int)p->devs.size() : -EINVAL;
}

int gmr1b200_pool_rx_xcch(struct gmr1b200_pool *p, int chan, const float *host_iq, int n_arfcn, int per_arfcn,
                          int win_len, int sps, float freq_shift0, uint8_t *l2, int32_t *crc, int32_t *conv, float *toa)
{
	if (!p || !host_iq || !l2 || n_arfcn < 0 || per_arfcn < 1 || win_len < 1)
		return set_err(-EINVAL, "pool_rx_xcch: bad argument");
	const int G = (int)p->devs.size();
	const int64_t arfcn_samples = (int64_t)per_arfcn * win_len;
	const size_t arfcn_bytes = (size_t)arfcn_samples * 8;
	const int64_t per_chunk = std::min<int64_t>(p->chunk_bytes / (int64_t)arfcn_bytes, p->chunk_units / per_arfcn);
	if (per_chunk < 1)
		return set_err(-EINVAL, "pool_rx_xcch: one ARFCN does not fit a chunk (raise chunk_bytes)");
	return per_device(p, [&](Dev &d, int g) -> int {
		const int n_local = n_arfcn > g ? (n_arfcn - g + G - 1) / G : 0;        // ARFCNs g, g+G, g+2G, ...
		int k = 0;
		for (int a0 = 0; a0 < n_local; a0 += (int)per_chunk, k++) {
			Slot &s = d.slots[k % d.slots.size()];
			const int m = (int)std::min<int64_t>(per_chunk, n_local - a0);
			const int64_t first = g + (int64_t)a0 * G;                          // global ARFCN of the chunk's first row
			cudaError_t e = cudaMemcpy2DAsync(s.iq, arfcn_bytes, (const char *)host_iq + (size_t)first * arfcn_bytes,
			                                  (size_t)G * arfcn_bytes, arfcn_bytes, (size_t)m, cudaMemcpyHostToDevice, s.st);
			if (e != cudaSuccess)
				return cuda_rc(e, "pool_rx_xcch: H2D");
			const int n = m * per_arfcn;
			int rc = gmr1b200_rx_xcch_batch(chan, s.iq, (int64_t)m * arfcn_samples, nullptr, win_len, win_len, sps, nullptr,
			                                freq_shift0, s.l2, crc ? s.i32 : nullptr, conv ? s.i32 + n : nullptr,
			                                toa ? s.f32 : nullptr, nullptr, n, s.st);
			if (rc)
				return rc;
			auto back = [&](void *host, const void *dev, size_t row) -> cudaError_t {      // row: bytes per ARFCN
				return cudaMemcpy2DAsync((char *)host + (size_t)first * row, (size_t)G * row, dev, row, row, (size_t)m,
				                         cudaMemcpyDeviceToHost, s.st);
			};
			e = back(l2, s.l2, (size_t)per_arfcn * 24);
			if (e == cudaSuccess && crc) e = back(crc, s.i32, (size_t)per_arfcn * 4);
			if (e == cudaSuccess && conv) e = back(conv, s.i32 + n, (size_t)per_arfcn * 4);
			if (e == cudaSuccess && toa) e = back(toa, s.f32, (size_t)per_arfcn * 4);
			if (e != cudaSuccess)
				return cuda_rc(e, "pool_rx_xcch: D2H");
		}
		return 0;
	});
}

int gmr1b200_pool_fcch_acquire(struct gmr1b200_pool *p, int fcch_type, const float *host_iq, int n_arfcn, int win_len,
                               int sps, int32_t *rough, int32_t *align, float *freq_error)
{
	if (!p || !host_iq || !align || !freq_error || n_arfcn < 0 || win_len < 1)
		return set_err(-EINVAL, "pool_fcch_acquire: bad argument");
	const int G = (int)p->devs.size();
	const size_t row_bytes = (size_t)win_len * 8;
	const int64_t per_chunk = std::min<int64_t>(p->chunk_bytes / (int64_t)row_bytes, p->chunk_units);
	if (per_chunk < 1)
		return set_err(-EINVAL, "pool_fcch_acquire: one window does not fit a chunk (raise chunk_bytes)");
	return per_device(p, [&](Dev &d, int g) -> int {
		const int n_local = n_arfcn > g ? (n_arfcn - g + G - 1) / G : 0;
		int k = 0;
		for (int a0 = 0; a0 < n_local; a0 += (int)per_chunk, k++) {
			Slot &s = d.slots[k % d.slots.size()];
			const int m = (int)std::min<int64_t>(per_chunk, n_local - a0);
			const int64_t first = g + (int64_t)a0 * G;
			cudaError_t e = cudaMemcpy2DAsync(s.iq, row_bytes, (const char *)host_iq + (size_t)first * row_bytes,
			                                  (size_t)G * row_bytes, row_bytes, (size_t)m, cudaMemcpyHostToDevice, s.st);
			if (e != cudaSuccess)
				return cuda_rc(e, "pool_fcch_acquire: H2D");
			int rc = gmr1b200_fcch_acquire_batch(fcch_type, s.iq, (int64_t)m * win_len, nullptr, win_len, win_len, sps,
			                                     s.i32, s.i32 + m, s.f32, m, s.st);
			if (rc)
				return rc;
			auto back = [&](void *host, const void *dev) -> cudaError_t {
				return cudaMemcpy2DAsync((char *)host + (size_t)first * 4, (size_t)G * 4, dev, 4, 4, (size_t)m,
				                         cudaMemcpyDeviceToHost, s.st);
			};
			e = back(align, s.i32 + m);
			if (e == cudaSuccess) e = back(freq_error, s.f32);
			if (e == cudaSuccess && rough) e = back(rough, s.i32);
			if (e != cudaSuccess)
				return cuda_rc(e, "pool_fcch_acquire: D2H");
		}
		return 0;
	});
}

}  // extern "C"
