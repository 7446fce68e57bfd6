// ort_kernels.cuh -- every __global__ function of libort_b200.so (sm_100a): the trace kernels (default and the
// selectable experiment variants), the delta scatter, the fixture noise kernels and the roofline diagnostic.
// Included by ort_device.cu only; the per-ray traversal itself lives in ort_trace.cuh.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "ort_noise.h"
#include "ort_trace.cuh"

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

namespace ort {

// delta upload: one thread per 16-byte half node
__global__ void scatter_nodes_kernel(uint4* __restrict__ nodes, const uint32_t* __restrict__ ids, const uint4* __restrict__ src, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < 2u * n)
		nodes[2u * (ids[i >> 1] - 1u) + (i & 1u)] = src[i];
}

__global__ void fill_miss_kernel(uint32_t* __restrict__ voxel, uint8_t* __restrict__ face, float* __restrict__ t, uint16_t* __restrict__ npush, size_t n)
{
	const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i < n)
	{
		voxel[i] = 0;
		face[i] = 6;
		t[i] = __uint_as_float(0x7F800000u);
		if (npush) npush[i] = 0;
	}
}

__global__ void fill_u32_kernel(uint32_t* __restrict__ dst, uint32_t v, size_t n)
{
	const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i < n) dst[i] = v;
}

// explicit rays: thread i traces ray i
template<int VARIANT, bool COUNT>
__global__ void __launch_bounds__(256)
trace_rays_kernel(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, RcpTable rt,
                  const float* __restrict__ o3, int o_stride, const float* __restrict__ d3, size_t n,
                  uint32_t* __restrict__ voxel, uint8_t* __restrict__ face, float* __restrict__ t, uint16_t* __restrict__ npush)
{
	const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* o = o3 + i * static_cast<size_t>(o_stride);
	const float* d = d3 + i * 3;
	const float ox = __ldg(o), oy = __ldg(o + 1), oz = __ldg(o + 2);
	const Ray r = ray_setup(rt, ox, oy, oz, __ldg(d), __ldg(d + 1), __ldg(d + 2));
	const Hit h = traverse_variant<VARIANT, COUNT>(nodes_m1, root, depth, miss_t, ox, oy, oz, r);
	voxel[i] = h.voxel;
	face[i] = static_cast<uint8_t>(h.face);
	t[i] = h.t;
	if (COUNT) npush[i] = static_cast<uint16_t>(min(h.npush, 65535u));
}

// FrameRows / frame_row (strip-local row -> frame row) live in ort_trace.cuh, next to the camera

// camera rays: a warp owns an 8 x 4 pixel tile (coherent rays -> shared upper-level nodes),
// a 256-thread block a 16 x 16 pixel tile.  SHAPED = true is the measurement build that also takes other warp tiles
// (fr.tile_shape) and block heights (blockDim.x / 16 rows); the default build has the mapping fixed, which is worth
// ~2 % (no runtime branches or special-register reads in the prologue).
template<int VARIANT, bool COUNT, bool SHAPED>
__global__ void __launch_bounds__(256)
trace_frame_kernel(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, RcpTable rt, Camera cam, FrameRows fr,
                   uint32_t* __restrict__ voxel, uint8_t* __restrict__ face, float* __restrict__ t, uint16_t* __restrict__ npush)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	int x, r;
	if (SHAPED && fr.tile_shape == 1)      { x = blockIdx.x * 16 + (lane & 15);                   r = blockIdx.y * 16 + warp * 2 + (lane >> 4); }
	else if (SHAPED && fr.tile_shape == 2) { x = blockIdx.x * 16 + (warp & 3) * 4 + (lane & 3);   r = blockIdx.y * 16 + (warp >> 2) * 8 + (lane >> 2); }
	else if (SHAPED && fr.tile_shape == 3) { x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);   r = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3); }   // 32 x 8 block: stays inside one 8-row strip
	else if (SHAPED)                       { x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);   r = blockIdx.y * static_cast<int>(blockDim.x >> 4) + (warp >> 1) * 4 + (lane >> 3); }
	else
	{
		unsigned by = blockIdx.y + static_cast<unsigned>(fr.band_rotate);
		if (by >= gridDim.y) by -= gridDim.y;
		x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
		r = static_cast<int>(by) * 16 + (warp >> 1) * 4 + (lane >> 3);
	}
	if (x >= fr.W || r >= fr.rows) return;
	const int y = frame_row(fr, r);

	float dx, dy, dz;
	camera_ray(cam, x, y, dx, dy, dz);
	const Ray ray = ray_setup(rt, cam.ox, cam.oy, cam.oz, dx, dy, dz);
	const Hit h = traverse_variant<VARIANT, COUNT>(nodes_m1, root, depth, miss_t, cam.ox, cam.oy, cam.oz, ray);

	const size_t i = static_cast<size_t>(r) * fr.W + x;
	voxel[i] = h.voxel;
	face[i] = static_cast<uint8_t>(h.face);
	t[i] = h.t;
	if (COUNT) npush[i] = static_cast<uint16_t>(min(h.npush, 65535u));
}


// Several frame jobs in ONE launch (blockIdx.z = job): strips of different frames, or the views of a multi-camera
// rig.  Separate launches each end with the latency tail of their longest rays; here the blocks of all jobs stream
// through the SMs back to back and only the last job's tail is exposed.  The jobs travel in the kernel parameters.
constexpr int kMaxJobs = 32;

struct FrameJob
{
	Camera cam;
	FrameRows fr;
	uint32_t* voxel;
	uint8_t*  face;
	float*    t;
	uint16_t* npush;    // may be null
};

struct FrameJobBatch
{
	FrameJob job[kMaxJobs];
};

template<bool COUNT>          // COUNT: some job of the batch wants per-ray PUSH counts (jobs without an npush pointer skip the store)
__global__ void __launch_bounds__(256)
trace_frames_kernel(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, RcpTable rt, const __grid_constant__ FrameJobBatch batch)
{
	const FrameJob& jb = batch.job[blockIdx.z];
	const FrameRows& fr = jb.fr;
	const int bands = (fr.rows + 15) >> 4;
	if (static_cast<int>(blockIdx.y) >= bands) return;                  // the grid is sized for the largest job
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	int by = static_cast<int>(blockIdx.y) + fr.band_rotate;
	if (by >= bands) by -= bands;
	const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
	const int r = by * 16 + (warp >> 1) * 4 + (lane >> 3);
	if (x >= fr.W || r >= fr.rows) return;
	const int y = frame_row(fr, r);

	float dx, dy, dz;
	camera_ray(jb.cam, x, y, dx, dy, dz);
	const Ray ray = ray_setup(rt, jb.cam.ox, jb.cam.oy, jb.cam.oz, dx, dy, dz);
	const Hit h = traverse_variant<1, COUNT>(nodes_m1, root, depth, miss_t, jb.cam.ox, jb.cam.oy, jb.cam.oz, ray);

	const size_t i = static_cast<size_t>(r) * fr.W + x;
	jb.voxel[i] = h.voxel;
	jb.face[i] = static_cast<uint8_t>(h.face);
	jb.t[i] = h.t;
	if (COUNT && jb.npush) jb.npush[i] = static_cast<uint16_t>(min(h.npush, 65535u));
}

// Shading epilogue (tree_camera::trace_pixel, test_och_h_octree.cpp:76-84) fused into the frame kernel: the hit is
// turned into the pixel the demo would Draw() -- colours[6 * (voxel - 1) + face], the sky colour on exit, the
// "inside" colour when the origin sits in a solid voxel -- and only that uint32 leaves the SM (4 B per pixel
// instead of 9).  Voxel types beyond the palette (the reference reads past its array there) shade as 0.
struct Palette
{
	const uint32_t* colours;
	uint32_t n_voxels;
	uint32_t exit_rgba, inside_rgba;
};

__global__ void __launch_bounds__(256)
trace_frame_rgba_kernel(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, RcpTable rt, Camera cam, FrameRows fr,
                        Palette pal, uint32_t* __restrict__ rgba)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned by = blockIdx.y + static_cast<unsigned>(fr.band_rotate);
	if (by >= gridDim.y) by -= gridDim.y;
	const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
	const int r = static_cast<int>(by) * 16 + (warp >> 1) * 4 + (lane >> 3);
	if (x >= fr.W || r >= fr.rows) return;
	const int y = frame_row(fr, r);

	float dx, dy, dz;
	camera_ray(cam, x, y, dx, dy, dz);
	const Ray ray = ray_setup(rt, cam.ox, cam.oy, cam.oz, dx, dy, dz);
	const Hit h = traverse_variant<1, false>(nodes_m1, root, depth, miss_t, cam.ox, cam.oy, cam.oz, ray);

	uint32_t px;
	if (h.face == 6u) px = pal.exit_rgba;
	else if (h.face == 7u) px = pal.inside_rgba;
	else px = (h.voxel - 1u < pal.n_voxels) ? __ldg(pal.colours + 6u * (h.voxel - 1u) + h.face) : 0u;
	rgba[static_cast<size_t>(r) * fr.W + x] = px;
}

// Variants 5 / 6: TightWalker (leaner bookkeeping per round, see ort_trace.cuh).  WW = false keeps the
// "if-if" round of the default kernel (one child load, then descend OR advance); WW = true is the "while-while"
// shape: every lane first advances over empty child slots until it holds a non-empty child (or leaves the tree),
// then the whole warp descends together.
template<bool COUNT, bool WW>
__global__ void __launch_bounds__(256)
trace_frame_tight_kernel(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, RcpTable rt, Camera cam, FrameRows fr,
                         uint32_t* __restrict__ voxel, uint8_t* __restrict__ face, float* __restrict__ t, uint16_t* __restrict__ npush)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
	const int r = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
	if (x >= fr.W || r >= fr.rows) return;
	const int y = frame_row(fr, r);

	float dx, dy, dz;
	camera_ray(cam, x, y, dx, dy, dz);
	const Ray ray = ray_setup(rt, cam.ox, cam.oy, cam.oz, dx, dy, dz);
	uint32_t stack[kMaxDepth];
	Hit h;
	if (fast_path_ok(cam.ox, cam.oy, cam.oz, ray))
	{
		TightWalker<COUNT> w;
		w.start(root, miss_t, ray);
		if (WW)
		{
			for (;;)
			{
				uint32_t child;
				bool done = false;
				while ((child = w.load_child(nodes_m1)) == 0u)
					if (w.advance(stack)) { done = true; break; }
				if (done || w.descend(child, depth, stack))
					break;
			}
		}
		else
		{
			for (;;)
			{
				const uint32_t child = w.load_child(nodes_m1);
				if (child ? w.descend(child, depth, stack) : w.advance(stack))
					break;
			}
		}
		h = w.hit;
	}
	else
		h = traverse(nodes_m1, root, depth, miss_t, ray, stack);

	const size_t i = static_cast<size_t>(r) * fr.W + x;
	voxel[i] = h.voxel;
	face[i] = static_cast<uint8_t>(h.face);
	t[i] = h.t;
	if (COUNT) npush[i] = static_cast<uint16_t>(min(h.npush, 65535u));
}

// Variant 12: persistent warps over tiles.  The default kernel's blocks retire when their slowest warp does, which
// leaves warp slots empty (achieved occupancy 83 %).  Here a resident grid is launched once and every WARP draws its
// next 8 x 4 tile from a global counter as soon as it is done, in the order the default kernel would have used
// (8 consecutive tiles = one 16 x 16 block tile), so slots never wait for a block mate.
template<bool COUNT>
__global__ void __launch_bounds__(256, 8)
trace_frame_tiles_kernel(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, RcpTable rt, Camera cam, FrameRows fr,
                         unsigned int n_tiles, unsigned int* __restrict__ counter,
                         uint32_t* __restrict__ voxel, uint8_t* __restrict__ face, float* __restrict__ t, uint16_t* __restrict__ npush)
{
	const unsigned lane = threadIdx.x & 31u;
	const unsigned blocks_x = (fr.W + 15) / 16;
	for (;;)
	{
		unsigned tile = 0;
		if (lane == 0) tile = atomicAdd(counter, 1u);
		tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
		if (tile >= n_tiles) return;
		const unsigned blk = tile >> 3, sub = tile & 7u;
		const int x = static_cast<int>(blk % blocks_x) * 16 + static_cast<int>(sub & 1u) * 8 + static_cast<int>(lane & 7u);
		const int r = static_cast<int>(blk / blocks_x) * 16 + static_cast<int>(sub >> 1) * 4 + static_cast<int>(lane >> 3);
		if (x >= fr.W || r >= fr.rows) continue;
		const int y = frame_row(fr, r);

		float dx, dy, dz;
		camera_ray(cam, x, y, dx, dy, dz);
		const Ray ray = ray_setup(rt, cam.ox, cam.oy, cam.oz, dx, dy, dz);
		const Hit h = traverse_variant<1, COUNT>(nodes_m1, root, depth, miss_t, cam.ox, cam.oy, cam.oz, ray);

		const size_t i = static_cast<size_t>(r) * fr.W + x;
		voxel[i] = h.voxel;
		face[i] = static_cast<uint8_t>(h.face);
		t[i] = h.t;
		if (COUNT) npush[i] = static_cast<uint16_t>(min(h.npush, 65535u));
	}
}

// Variant 7: PipeWalker (ALU-lean bookkeeping, see ort_trace.cuh).
template<bool COUNT>
__global__ void __launch_bounds__(256)
trace_frame_pipe_kernel(const uint32_t* __restrict__ nodes_m1, unsigned long long base_biased, uint32_t root, int depth, float miss_t, RcpTable rt, Camera cam, FrameRows fr,
                        uint32_t* __restrict__ voxel, uint8_t* __restrict__ face, float* __restrict__ t, uint16_t* __restrict__ npush)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
	const int r = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
	if (x >= fr.W || r >= fr.rows) return;
	const int y = frame_row(fr, r);

	float dx, dy, dz;
	camera_ray(cam, x, y, dx, dy, dz);
	const Ray ray = ray_setup(rt, cam.ox, cam.oy, cam.oz, dx, dy, dz);
	Hit h;
	if (fast_path_ok(cam.ox, cam.oy, cam.oz, ray))
	{
		uint32_t stack_n[kMaxDepth];
		float stack_f[kMaxDepth];
		PipeWalker<COUNT> w;
		w.start(root, miss_t, ray);
		for (;;)
		{
			const uint32_t child = w.load_child(base_biased);
			if (child ? w.descend(child, depth, stack_n, stack_f) : w.advance(stack_n, stack_f))
				break;
		}
		h = w.hit;
	}
	else
	{
		uint32_t stack[kMaxDepth];
		h = traverse(nodes_m1, root, depth, miss_t, ray, stack);
	}

	const size_t i = static_cast<size_t>(r) * fr.W + x;
	voxel[i] = h.voxel;
	face[i] = static_cast<uint8_t>(h.face);
	t[i] = h.t;
	if (COUNT) npush[i] = static_cast<uint16_t>(min(h.npush, 65535u));
}

// Probe kernels (variants 8 / 9): the default walk plus PROBE_FMA dependent-free FMA-pipe instructions or PROBE_ALU
// ALU-pipe instructions per round, on dummy accumulators that are folded into the result only if they take an
// impossible value.  They answer "which resource binds the loop?": extra work on a unit that has slack is free.
template<int PROBE_FMA, int PROBE_ALU>
__global__ void __launch_bounds__(256)
trace_frame_probe_kernel(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, RcpTable rt, Camera cam, FrameRows fr,
                         uint32_t* __restrict__ voxel, uint8_t* __restrict__ face, float* __restrict__ t)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
	const int r = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
	if (x >= fr.W || r >= fr.rows) return;
	const int y = frame_row(fr, r);

	float dx, dy, dz;
	camera_ray(cam, x, y, dx, dy, dz);
	const Ray ray = ray_setup(rt, cam.ox, cam.oy, cam.oz, dx, dy, dz);
	uint32_t stack[kMaxDepth];
	Hit h;
	float facc[3] = { dx, dy, dz };
	uint32_t iacc[3] = { ray.px, ray.py, ray.pz };
	if (fast_path_ok(cam.ox, cam.oy, cam.oz, ray))
	{
		FastWalker<false> w;
		w.start(root, miss_t, ray);
		for (;;)
		{
			const uint32_t child = w.load_child(nodes_m1);
#pragma unroll
			for (int k = 0; k < PROBE_FMA; ++k)
				asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(facc[k % 3]) : "f"(w.cx), "f"(w.bx));
#pragma unroll
			for (int k = 0; k < PROBE_ALU; ++k)
				asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(iacc[k % 3]) : "r"(w.idx), "r"(w.inv));
			if (child ? w.descend(child, depth, stack) : w.advance(stack))
				break;
		}
		h = w.hit;
	}
	else
		h = traverse(nodes_m1, root, depth, miss_t, ray, stack);
	if (facc[0] + facc[1] + facc[2] == 1.2345e-30f || (iacc[0] ^ iacc[1] ^ iacc[2]) == 0xDEADBEEFu)
		h.voxel ^= 0x80000000u;                                                  // never true; keeps the probes alive

	const size_t i = static_cast<size_t>(r) * fr.W + x;
	voxel[i] = h.voxel;
	face[i] = static_cast<uint8_t>(h.face);
	t[i] = h.t;
}

// Experiment kernel for the SIMT-efficiency question (variant 4): "deferred phases".  In the default kernel every
// round of the loop runs the descend block for the lanes whose child exists AND the advance block for the lanes
// whose child is empty -- each with about two thirds of the warp.  Here a phase that fewer than `threshold` lanes
// want is postponed (those lanes keep their loaded child and wait) as long as the other phase has enough takers, in
// the hope that the stragglers' phase fills up.  Costs two ballots per round.
template<bool COUNT>
__global__ void __launch_bounds__(256)
trace_frame_deferred_kernel(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, RcpTable rt, Camera cam, FrameRows fr,
                            int threshold,
                            uint32_t* __restrict__ voxel, uint8_t* __restrict__ face, float* __restrict__ t, uint16_t* __restrict__ npush)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
	const int r = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
	const bool valid = x < fr.W && r < fr.rows;
	const int y = frame_row(fr, r);

	float dx, dy, dz;
	camera_ray(cam, x, y, dx, dy, dz);
	const Ray ray = ray_setup(rt, cam.ox, cam.oy, cam.oz, dx, dy, dz);
	uint32_t stack[kMaxDepth];
	FastWalker<COUNT> w;
	w.start(root, miss_t, ray);
	int st = 0;                       // 0 load next child, 1 wants descend (child held), 2 wants advance, 3 finished
	uint32_t child = 0;
	if (!valid)
		st = 3;
	else if (!fast_path_ok(cam.ox, cam.oy, cam.oz, ray))
	{
		w.hit = traverse(nodes_m1, root, depth, miss_t, ray, stack);
		st = 3;
	}

	for (;;)
	{
		if (st == 0)
		{
			child = w.load_child(nodes_m1);
			st = child ? 1 : 2;
		}
		const unsigned md = __ballot_sync(0xFFFFFFFFu, st == 1), ma = __ballot_sync(0xFFFFFFFFu, st == 2);
		if ((md | ma) == 0u)
			break;
		const int nd = __popc(md), na = __popc(ma);
		const bool run_d = nd >= threshold || na < threshold;
		const bool run_a = na >= threshold || nd < threshold;
		if (run_d && st == 1) st = w.descend(child, depth, stack) ? 3 : 0;
		if (run_a && st == 2) st = w.advance(stack) ? 3 : 0;
	}

	if (valid)
	{
		const size_t i = static_cast<size_t>(r) * fr.W + x;
		voxel[i] = w.hit.voxel;
		face[i] = static_cast<uint8_t>(w.hit.face);
		t[i] = w.hit.t;
		if (COUNT) npush[i] = static_cast<uint16_t>(min(w.hit.npush, 65535u));
	}
}

// Experiment kernel for the "upper levels in shared memory" question: 1024-thread blocks (a 32 x 32 pixel tile,
// warps still 8 x 4) copy the first n_staged nodes -- the top levels, a contiguous prefix of the level-ordered
// array -- into shared memory and serve PUSHes on those nodes from there.  Only valid in the h_octree layout.
// Kept selectable (variant 3) so that the decision can be re-measured; see DESIGN.md section 4 for the numbers.
template<bool COUNT>
__global__ void __launch_bounds__(1024)
trace_frame_staged_kernel(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, RcpTable rt, Camera cam, FrameRows fr,
                          uint32_t n_staged,
                          uint32_t* __restrict__ voxel, uint8_t* __restrict__ face, float* __restrict__ t, uint16_t* __restrict__ npush)
{
	extern __shared__ uint4 s_raw[];
	uint32_t* s_nodes = reinterpret_cast<uint32_t*>(s_raw);
	{
		const uint4* src = reinterpret_cast<const uint4*>(nodes_m1 + 8);            // id 1
		for (uint32_t i = threadIdx.x; i < 2u * n_staged; i += blockDim.x) s_raw[i] = __ldg(src + i);
	}
	__syncthreads();
	const uint32_t* s_nodes_m1 = s_nodes - 8;

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
	const int r = blockIdx.y * 32 + (warp >> 2) * 4 + (lane >> 3);
	if (x >= fr.W || r >= fr.rows) return;
	const int y = frame_row(fr, r);

	float dx, dy, dz;
	camera_ray(cam, x, y, dx, dy, dz);
	const Ray ray = ray_setup(rt, cam.ox, cam.oy, cam.oz, dx, dy, dz);
	uint32_t stack[kMaxDepth];
	Hit h;
	if (fast_path_ok(cam.ox, cam.oy, cam.oz, ray))
	{
		FastWalker<COUNT> w;
		w.start(root, miss_t, ray);
		while (!w.iterate_staged(nodes_m1, depth, stack, s_nodes_m1, n_staged)) {}
		h = w.hit;
	}
	else
		h = traverse(nodes_m1, root, depth, miss_t, ray, stack);

	const size_t i = static_cast<size_t>(r) * fr.W + x;
	voxel[i] = h.voxel;
	face[i] = static_cast<uint8_t>(h.face);
	t[i] = h.t;
	if (COUNT) npush[i] = static_cast<uint16_t>(min(h.npush, 65535u));
}

// Fixture kernels: the noise evaluations of the demo's terrain set-up, one thread per column / voxel.
// get_terrain_heigth over the whole map (test_och_h_octree.cpp:561-566, :587-592)
__global__ void __launch_bounds__(256)
fixture_heightmap_kernel(uint16_t* __restrict__ heights, int dim)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
	if (x < dim) heights[static_cast<size_t>(y) * dim + x] = ort_noise::terrain_height(x, y, dim);
}

// remove(tree, splatter_noise(-0.5F, .., 1/16))'s dim^3 test (:735-743, :755-763) for the voxels at or below the
// surface: bit (y * dim + x) of slab z = "carved".  A warp covers 32 consecutive x and writes one 32-bit word.
__global__ void __launch_bounds__(256)
fixture_carve_kernel(const uint16_t* __restrict__ heights, int dim, uint32_t* __restrict__ bits, size_t words32_per_slab)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
	const bool carved = x < dim && z <= static_cast<int>(heights[static_cast<size_t>(y) * dim + x]) && ort_noise::carve_test(x, y, z);
	const unsigned w = __ballot_sync(0xFFFFFFFFu, carved);
	if ((threadIdx.x & 31u) == 0u && x < dim)
		bits[static_cast<size_t>(z) * words32_per_slab + ((static_cast<size_t>(y) * dim + x) >> 5)] = w;
}

// Diagnostic: random 32-byte-sector gather over an L2-resident buffer -- the memory-side ceiling of a
// traversal whose nodes live in L2 (one 4-byte child read moves one sector).  Independent loads, 8 in
// flight per thread, addresses from a counter hash so that L1 cannot help.
__global__ void __launch_bounds__(256)
gather_peak_kernel(const uint32_t* __restrict__ buf, uint32_t n_sectors, uint32_t iters, uint32_t* __restrict__ sink)
{
	uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	uint32_t acc = 0;
	for (uint32_t i = 0; i < iters; ++i)
	{
		uint32_t v[8];
#pragma unroll
		for (int k = 0; k < 8; ++k)
		{
			x = x * 1664525u + 1013904223u;
			const uint32_t s = __umulhi(x ^ (x >> 15), n_sectors);            // uniform in [0, n_sectors)
			v[k] = __ldg(buf + (static_cast<size_t>(s) << 3) + (x & 7u));
		}
#pragma unroll
		for (int k = 0; k < 8; ++k) acc ^= v[k];
	}
	if (acc == 0x9E3779B9u) *sink = acc;                                      // keep the loads alive
}

// ------------------------------------------------------------------------------------------------
// Persistent warps with lane refill ("warp-level ray compaction").
//
// A fixed grid (one resident wave) walks the ray list: each warp draws batches of ray indices from a global
// counter and keeps its 32 lanes busy -- when a lane's ray ends, the lane writes its result and goes idle;
// once the number of busy lanes (ballot + popc) falls to `low_water` and rays remain, the idle lanes are
// refilled before traversal continues.  This trades a ballot per traversal round and scattered result
// writes for lanes that no longer wait on the slowest ray of their warp: worth little on coherent camera
// rays, a lot on incoherent rays (BASELINE config 3: 10 PUSHes on average, 600 worst case).
// FRAME = true enumerates the pixels of the strip in 8x4 tile order (index = 32 * tile + lane-in-tile) so
// that consecutive indices stay spatially coherent.
// ------------------------------------------------------------------------------------------------

constexpr unsigned kBatch = 128;    // ray indices a warp draws per atomicAdd

// MINB = minimum resident blocks per SM the compiler must allow for (register budget): 1 -> 47 registers, 57 %
// occupancy; 6 -> 40 registers (a few spilled words), 75 %; 8 -> 32 registers, 100 %.
template<bool COUNT, bool FRAME, int MINB = 1>
__global__ void __launch_bounds__(256, MINB)
trace_persistent_kernel(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, RcpTable rt,
                        const float* __restrict__ o3, int o_stride, const float* __restrict__ d3, Camera cam, FrameRows fr,
                        unsigned long long n, unsigned long long* __restrict__ counter, int low_water,
                        uint32_t* __restrict__ voxel, uint8_t* __restrict__ face, float* __restrict__ t, uint16_t* __restrict__ npush)
{
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt_mask = (1u << lane) - 1u;
	const unsigned tiles_x = FRAME ? (fr.W + 7) / 8 : 0;

	uint32_t stack[kMaxDepth];
	FastWalker<COUNT> w;
	bool active = false;
	size_t out = 0;                              // where this lane's result goes
	unsigned long long next = 0, end = 0;        // the warp's current batch (uniform)
	bool exhausted = false;                      // (uniform)

	auto store = [&](size_t i, const Hit& h) {
		voxel[i] = h.voxel;
		face[i] = static_cast<uint8_t>(h.face);
		t[i] = h.t;
		if (COUNT) npush[i] = static_cast<uint16_t>(min(h.npush, 65535u));
	};

	for (;;)
	{
		// ---- refill idle lanes ---------------------------------------------------------------
		unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
		while (idle != 0u && !exhausted)
		{
			if (next == end)
			{
				unsigned long long b = 0;
				if (lane == 0) b = atomicAdd(counter, static_cast<unsigned long long>(kBatch));
				b = __shfl_sync(0xFFFFFFFFu, b, 0);
				if (b >= n) { exhausted = true; break; }
				next = b;
				end = b + kBatch < n ? b + kBatch : n;
			}
			const unsigned long long left = end - next;
			const unsigned n_idle = __popc(idle);
			const unsigned avail = left < n_idle ? static_cast<unsigned>(left) : n_idle;
			const unsigned rank = __popc(idle & lt_mask);
			if (!active && rank < avail)
			{
				const unsigned long long i = next + rank;
				float ox, oy, oz, dx, dy, dz;
				bool valid = true;
				if (FRAME)
				{
					const unsigned tile = static_cast<unsigned>(i >> 5), l = static_cast<unsigned>(i) & 31u;
					const int x = static_cast<int>(tile % tiles_x) * 8 + static_cast<int>(l & 7u);
					const int r = static_cast<int>(tile / tiles_x) * 4 + static_cast<int>(l >> 3);
					valid = x < fr.W && r < fr.rows;
					const int y = frame_row(fr, r);
					ox = cam.ox; oy = cam.oy; oz = cam.oz;
					camera_ray(cam, x, y, dx, dy, dz);
					out = static_cast<size_t>(r) * fr.W + x;
				}
				else
				{
					const float* o = o3 + i * static_cast<unsigned long long>(o_stride);
					const float* d = d3 + i * 3ull;
					ox = __ldg(o); oy = __ldg(o + 1); oz = __ldg(o + 2);
					dx = __ldg(d); dy = __ldg(d + 1); dz = __ldg(d + 2);
					out = static_cast<size_t>(i);
				}
				if (valid)
				{
					const Ray ray = ray_setup(rt, ox, oy, oz, dx, dy, dz);
					if (fast_path_ok(ox, oy, oz, ray))
					{
						w.start(root, miss_t, ray);
						active = true;
					}
					else
						store(out, traverse(nodes_m1, root, depth, miss_t, ray, stack));       // out-of-domain ray: the generic walk, right away
				}
			}
			next += avail;
			idle = __ballot_sync(0xFFFFFFFFu, !active);
		}

		if (__ballot_sync(0xFFFFFFFFu, active) == 0u)
			break;

		// ---- traverse until too few lanes are busy ---------------------------------------------
		for (;;)
		{
			if (active && w.iterate(nodes_m1, depth, stack))
			{
				store(out, w.hit);
				active = false;
			}
			const unsigned busy = __ballot_sync(0xFFFFFFFFu, active);
			if (busy == 0u || (!exhausted && __popc(busy) <= low_water))
				break;
		}
	}
}

}  // namespace ort
