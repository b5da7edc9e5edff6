// download.cuh -- device-driven copy-out of a batch's renderer-facing arrays (GLChunk::format_data's p_data / n_data / c_data
// + index buffer, GLChunk.cpp:278-296) straight into pinned, device-mapped HOST buffers.
//
// Why a kernel and not cudaMemcpyAsync: the sizes of a batch's arrays are only known on the device (the emitters run
// against arena capacities and never wait for the host), so a copy-engine transfer needs a host round trip -- wait for the
// totals, then enqueue the copies -- between the last kernel and the first byte on the wire.  This kernel reads the totals
// where they are and stores through the mapped pointers (posted PCIe writes, 16 bytes per thread, fully coalesced), so the
// download is enqueued right behind the batch without ever synchronising, and it can narrow the data on the way:
//   * chunk-local indices are < n_verts of their chunk; when every chunk of the batch has < 65536 vertices they travel as
//     uint16 (half the bytes of the largest stream),
//   * streams that are provably constant on this path (colour == 1 everywhere; normal == 0 when no step writes it) are
//     skipped -- the chunk table says so (bmf_chunk_info.flags) and the consumer synthesises them.
// The grid is small on purpose (the PCIe link, not the SMs, is the limit): the CTAs sit beside the next batch's compute
// kernels of a second context instead of displacing them.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "extract.cuh"

namespace bmf
{

// slots of the device-side totals block (unsigned long long each) shared by all kernels of a batch
enum
{
	TOT_CELLS = 0, TOT_VERTS = 1, TOT_INDS = 2,
	TOT_OVERFLOW = 3, // a total reached 2^32
	TOT_LIST0 = 4, TOT_LIST1 = 5, // surface-cell list counters (k_bases)
	TOT_WORK = 6,     // chunk work counter (k_smooth_chunks)
	TOT_SMALL = 7,    // an output arena is too small for this batch: the emitters did not run
	TOT_MAXV = 8,     // largest n_verts of any chunk of the batch
	TOT_DLERR = 9,    // download: 1 = host buffers too small, 2 = uint16 indices requested but a chunk has >= 65536 vertices
	TOT_SLOTS = 16
};

struct DownloadArgs
{
	// device-visible addresses of the caller's pinned host buffers (null = stream not wanted)
	float* pos;
	float* normal;
	float* color;
	uint8_t* boundary;
	uint8_t* valence;
	uint32_t* inds32;
	uint16_t* inds16;
	unsigned long long cap_verts, cap_inds;
	// batch arrays on the device
	const float* d_pos;
	const float* d_normal;
	const float* d_color;
	const uint8_t* d_boundary;
	const uint8_t* d_valence;
	const uint32_t* d_inds;
};

// n_bytes from src (device, 16-byte aligned) to dst (mapped host): 16-byte stores when dst is 16-byte aligned, 4-byte otherwise
__device__ __forceinline__ void copy_stream(void* dst, const void* src, size_t n_bytes, size_t tid, size_t nthreads)
{
	if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)
	{
		const size_t n16 = n_bytes >> 4;
		const uint4* s = reinterpret_cast<const uint4*>(src);
		uint4* d = reinterpret_cast<uint4*>(dst);
		// four independent 16-byte loads in flight per thread before the first store
		size_t i = tid;
		for (; i + 3 * nthreads < n16; i += 4 * nthreads)
		{
			const uint4 a = __ldcs(s + i), b = __ldcs(s + i + nthreads), c = __ldcs(s + i + 2 * nthreads), e = __ldcs(s + i + 3 * nthreads);
			d[i] = a; d[i + nthreads] = b; d[i + 2 * nthreads] = c; d[i + 3 * nthreads] = e;
		}
		for (; i < n16; i += nthreads) d[i] = __ldcs(s + i);
		for (size_t b = (n16 << 4) + tid; b < n_bytes; b += nthreads) reinterpret_cast<uint8_t*>(dst)[b] = reinterpret_cast<const uint8_t*>(src)[b];
	}
	else
	{
		const size_t n4 = ((reinterpret_cast<uintptr_t>(dst) & 3) == 0) ? (n_bytes >> 2) : 0;
		for (size_t i = tid; i < n4; i += nthreads) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
		for (size_t b = (n4 << 2) + tid; b < n_bytes; b += nthreads) reinterpret_cast<uint8_t*>(dst)[b] = reinterpret_cast<const uint8_t*>(src)[b];
	}
}

__global__ void __launch_bounds__(CTA) k_download(DownloadArgs A, unsigned long long* __restrict__ tot, unsigned long long* __restrict__ tot_host)
{
	if (tot[TOT_SMALL]) return; // the emitters did not run: the host grows the arenas, re-launches them and this kernel
	const unsigned long long V = tot[TOT_VERTS], I = tot[TOT_INDS];
	unsigned long long err = 0;
	const bool want_v = A.pos || A.normal || A.color || A.boundary || A.valence, want_i = A.inds32 || A.inds16;
	if ((want_v && V > A.cap_verts) || (want_i && I > A.cap_inds)) err = 1;
	else if (A.inds16 && tot[TOT_MAXV] > 65535ull) err = 2;
	if (blockIdx.x == 0 && threadIdx.x == 0)
	{
		tot[TOT_DLERR] = err;
		tot_host[TOT_DLERR] = err;
	}
	if (err) return;
	const size_t tid = (size_t)blockIdx.x * CTA + threadIdx.x, nthreads = (size_t)gridDim.x * CTA;
	if (A.pos) copy_stream(A.pos, A.d_pos, 12 * (size_t)V, tid, nthreads);
	if (A.normal) copy_stream(A.normal, A.d_normal, 12 * (size_t)V, tid, nthreads);
	if (A.color) copy_stream(A.color, A.d_color, 12 * (size_t)V, tid, nthreads);
	if (A.boundary) copy_stream(A.boundary, A.d_boundary, (size_t)V, tid, nthreads);
	if (A.valence) copy_stream(A.valence, A.d_valence, (size_t)V, tid, nthreads);
	if (A.inds32) copy_stream(A.inds32, A.d_inds, 4 * (size_t)I, tid, nthreads);
	if (A.inds16)
	{
		// 8 indices (two 16-byte loads) -> one 16-byte store
		const size_t n8 = (size_t)I >> 3;
		const uint4* s = reinterpret_cast<const uint4*>(A.d_inds);
		if ((reinterpret_cast<uintptr_t>(A.inds16) & 15) == 0)
		{
			uint4* d = reinterpret_cast<uint4*>(A.inds16);
			for (size_t i = tid; i < n8; i += nthreads)
			{
				const uint4 a = __ldcs(s + 2 * i), b = __ldcs(s + 2 * i + 1);
				d[i] = make_uint4(a.x | (a.y << 16), a.z | (a.w << 16), b.x | (b.y << 16), b.z | (b.w << 16));
			}
			for (size_t i = (n8 << 3) + tid; i < (size_t)I; i += nthreads) A.inds16[i] = (uint16_t)A.d_inds[i];
		}
		else
			for (size_t i = tid; i < (size_t)I; i += nthreads) A.inds16[i] = (uint16_t)A.d_inds[i];
	}
	// the stream's completion (event / synchronize) orders these stores for the host; nothing else to do
}

// chunk-local indices -> uint16 in device memory (the copy engine cannot narrow): 8 indices (two 16-byte loads) -> one 16-byte store
__global__ void __launch_bounds__(CTA) k_pack_indices16(const uint32_t* __restrict__ inds, size_t n, uint16_t* __restrict__ out)
{
	const size_t tid = (size_t)blockIdx.x * CTA + threadIdx.x, nthreads = (size_t)gridDim.x * CTA;
	const size_t n8 = n >> 3;
	const uint4* s = reinterpret_cast<const uint4*>(inds);
	uint4* d = reinterpret_cast<uint4*>(out);
	for (size_t i = tid; i < n8; i += nthreads)
	{
		const uint4 a = __ldcs(s + 2 * i), b = __ldcs(s + 2 * i + 1);
		d[i] = make_uint4(a.x | (a.y << 16), a.z | (a.w << 16), b.x | (b.y << 16), b.z | (b.w << 16));
	}
	for (size_t i = (n8 << 3) + tid; i < n; i += nthreads) out[i] = (uint16_t)inds[i];
}

} // namespace bmf
