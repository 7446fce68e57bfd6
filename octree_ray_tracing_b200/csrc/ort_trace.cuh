// ort_trace.cuh -- the per-ray DAG traversal, device side (sm_100a).
//
// Computes exactly what och::h_octree<L,D>::sse_trace computes (och_h_octree.h:292-447): an adapted
// Laine-Karras traversal with (i) the "dimension bit" -- one float-mantissa bit OR-ed into / masked
// out of the position instead of a scale pair -- and (ii) the early POP branch taken before any
// position update, so no overstep correction exists.  Everything that is an integer trick on float
// bit patterns in the SSE original is integer arithmetic here; the three float operations that
// decide results keep their exact IEEE form:
//     coef  = RCPPS(d)            -> table model (ort_rcp_model), bit-exact with the CPU instruction
//     bias  = -(coef * o)         -> __fmul_rn, never contracted
//     t     = fma(pos, coef, bias)-> __fmaf_rn (single rounding, like _mm_fmadd_ps)
// NaN results (axis-parallel rays: coef = -inf, bias = +inf) are canonicalised to x86's default NaN
// 0xFFC00000 because the reference orders t values by their raw bits as UNSIGNED integers (:384-406).
#pragma once

#include <cstdint>
#ifndef ORT_HOST_EMU          // tests/host_emu compiles this header for the host with its own stand-ins for the intrinsics
#include <cuda_runtime.h>
#endif

namespace ort {

constexpr int kMaxDepth = 16;

struct RcpTable
{
	const uint32_t* tab;   // 1 << log2n entries
	int shift;             // 23 - log2n
};

__device__ __forceinline__ uint32_t rcp_model(const RcpTable rt, uint32_t x)
{
	const uint32_t sign = x & 0x80000000u;
	const uint32_t e = (x >> 23) & 0xFFu;
	const uint32_t m = x & 0x7FFFFFu;
	if (e - 1u < 252u)
	{
		// common case: the result exponent field (>= 126 - 125 = 1 for Intel's table) cannot underflow
		const uint32_t r = __ldg(rt.tab + (m >> rt.shift));
		const int re = static_cast<int>(r >> 23) - (static_cast<int>(e) - 127);
		if (re > 0) return sign | (r - ((e - 127u) << 23));
		return sign;
	}
	if (e == 255u) return m ? (x | 0x00400000u) : sign;      // NaN stays NaN, inf -> 0
	if (e == 0u) return sign | 0x7F800000u;                  // 0 and denormals -> inf
	const uint32_t r = __ldg(rt.tab + (m >> rt.shift));
	const int re = static_cast<int>(r >> 23) - (static_cast<int>(e) - 127);
	if (re <= 0) return sign;                                // would be denormal -> 0
	return sign | (static_cast<uint32_t>(re) << 23) | (r & 0x7FFFFFu);
}

// x86 orders NaN (default NaN 0xFFC00000) after every number when t bits are compared as unsigned
__device__ __forceinline__ uint32_t t_bits(float t)
{
	const uint32_t b = __float_as_uint(t);
	return (b & 0x7FFFFFFFu) > 0x7F800000u ? 0xFFC00000u : b;
}

struct Ray
{
	float cx, cy, cz;        // coef
	float bx, by, bz;        // bias
	uint32_t px, py, pz;     // pos (float bit patterns in [1,2))
	uint32_t inv;            // inv_signs
	uint32_t idx;
};

__device__ __forceinline__ void ray_axis(const RcpTable rt, float o, float d, int a, float& coef, float& bias, uint32_t& pos, uint32_t& inv, uint32_t& idx)
{
	const bool sg = 0.0f < d;                                                  // :310
	inv |= static_cast<uint32_t>(sg) << a;                                     // :322
	const uint32_t dn = __float_as_uint(d) | 0x80000000u;                      // :312
	const float oa = fabsf(__fsub_rn(sg ? 3.0f : 0.0f, o));                    // :314
	coef = __uint_as_float(rcp_model(rt, dn));                                 // :316
	bias = __uint_as_float(__float_as_uint(__fmul_rn(coef, oa)) ^ 0x80000000u); // :318
	pos = __float_as_uint(oa) & 0x3FC00000u;                                   // :320
	idx |= static_cast<uint32_t>(pos == 0x3FC00000u) << a;                     // :324
}

__device__ __forceinline__ Ray ray_setup(const RcpTable rt, float ox, float oy, float oz, float dx, float dy, float dz)
{
	Ray r;
	r.inv = 0;
	r.idx = 0;
	ray_axis(rt, ox, dx, 0, r.cx, r.bx, r.px, r.inv, r.idx);
	ray_axis(rt, oy, dy, 1, r.cy, r.by, r.py, r.inv, r.idx);
	ray_axis(rt, oz, dz, 2, r.cz, r.bz, r.pz, r.inv, r.idx);
	return r;
}

struct Hit
{
	uint32_t voxel;
	uint32_t face;
	float    t;
	uint32_t npush;
};

// nodes_m1: node id i lives at nodes_m1[8*i .. 8*i+7] (h_octree layout: nodes - 8, ids 1-based; och::octree pool:
// the pool itself, raw rows, root = 0).  miss_t = hit_time reported by a MISS.
// Baseline variant: one thread walks one ray from start to end, parent stack in local memory.
__device__ __forceinline__ Hit traverse(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, Ray r, uint32_t* stack)
{
	uint32_t node = root;
	uint32_t dim = 1u << 22;                                                   // :326
	int      level = 1;                                                        // :334
	uint32_t mti = 8;                                                          // :336
	float    tmin = 0.0f;                                                      // :338
	uint32_t idx = r.idx;
	uint32_t px = r.px, py = r.py, pz = r.pz;
	Hit h;
	h.npush = 0;

	for (;;)
	{
		// PUSH (:342-376)
		++h.npush;
		const uint32_t child = __ldg(nodes_m1 + (static_cast<size_t>(node) << 3) + ((idx ^ r.inv) & 7u));

		if (child)
		{
			if (level == depth)                                                // :346 HIT
			{
				h.voxel = child;
				h.face = (mti >> 1) + 3u * ((r.inv & mti) == 0u);
				h.t = tmin;
				return h;
			}
			stack[level - 1] = node;                                           // :357
			++level;
			node = child;
			dim >>= 1;                                                         // :361
			const float tx = __fmaf_rn(__uint_as_float(px | dim), r.cx, r.bx); // :363-365
			const float ty = __fmaf_rn(__uint_as_float(py | dim), r.cy, r.by);
			const float tz = __fmaf_rn(__uint_as_float(pz | dim), r.cz, r.bz);
			const bool ux = tx >= tmin, uy = ty >= tmin, uz = tz >= tmin;      // :367 ordered compare, false on NaN
			idx = static_cast<uint32_t>(ux) | (static_cast<uint32_t>(uy) << 1) | (static_cast<uint32_t>(uz) << 2);
			px |= ux ? dim : 0u;                                               // :371-373
			py |= uy ? dim : 0u;
			pz |= uz ? dim : 0u;
			continue;
		}

		for (;;)
		{
			// STEP (:378-419)
			const uint32_t tx = t_bits(__fmaf_rn(__uint_as_float(px), r.cx, r.bx));
			const uint32_t ty = t_bits(__fmaf_rn(__uint_as_float(py), r.cy, r.by));
			const uint32_t tz = t_bits(__fmaf_rn(__uint_as_float(pz), r.cz, r.bz));
			const uint32_t tm = min(tx, min(ty, tz));                          // :388-406: unsigned argmin, ties -> x, y, z
			mti = tx == tm ? 1u : (ty == tm ? 2u : 4u);
			tmin = __uint_as_float(tm);

			if (idx & mti)                                                     // :410-419 step to the sibling
			{
				px &= ~(mti & 1u ? dim : 0u);
				py &= ~(mti & 2u ? dim : 0u);
				pz &= ~(mti & 4u ? dim : 0u);
				idx ^= mti;
				break;
			}

			// POP (:421-446)
			if (--level == 0)                                                  // :423 MISS
			{
				h.voxel = 0;
				h.face = 6;
				h.t = miss_t;
				return h;
			}
			node = stack[level - 1];                                           // :434
			px &= ~dim; py &= ~dim; pz &= ~dim;                                // :436
			dim <<= 1;                                                         // :438
			idx = static_cast<uint32_t>((px & dim) != 0u) | (static_cast<uint32_t>((py & dim) != 0u) << 1) | (static_cast<uint32_t>((pz & dim) != 0u) << 2); // :440-444
		}
	}
}


// ------------------------------------------------------------------------------------------------
// Fast variant.  Same decisions, same FMAs, different bookkeeping -- chosen from the ncu profile of the
// baseline (profiles/r1_v1_ncu_full.md: ALU pipe 81 % busy, FMA pipe 17 %):
//   * the position stays in FLOAT registers.  OR-ing the level's mantissa bit into a position whose lower
//     bits are clear is an exact float add of the cell size (and AND-NOT an exact subtract), so the bit
//     masks of the SSE original become FADDs on the idle FMA pipe -- valid because the origin is in [1,2)
//     (callers outside that domain are routed to traverse(), which is exact for any bit pattern);
//   * axes with d == +-0 or denormal (coef = -inf) would produce NaN t values, which x86 orders LAST as
//     unsigned bits (0xFFC00000).  Such an axis gets coef = 0, bias = -inf instead: t = -inf = 0xFF800000,
//     still after every finite value in unsigned order and still failing `t >= tmin`, so every decision is
//     unchanged and the per-STEP NaN canonicalisation disappears.  (A ray with all three components
//     degenerate would see tmin = -inf; it is routed to traverse() as well.);
//   * the child index of the level being left rides in the top 3 bits of the parent-stack entry (compact
//     ids stay below 2^29), so POP restores idx with one shift instead of three bit tests.
// ------------------------------------------------------------------------------------------------

constexpr uint32_t kIdMask = 0x1FFFFFFFu;

__device__ __forceinline__ bool in_unit_cube(float ox, float oy, float oz)
{
	// all three in [1, 2): exponent field 127, sign 0
	return ((__float_as_uint(ox) >> 23) == 127u) & ((__float_as_uint(oy) >> 23) == 127u) & ((__float_as_uint(oz) >> 23) == 127u);
}

// State of one ray in flight.  iterate() runs ONE round of the reference's PUSH label (one child-slot load)
// plus whatever STEP/POP work follows it, and returns true when the ray is finished (result in `hit`).
// Kept as a resumable object so that the persistent kernel can interleave lane refills with traversal.
template<bool COUNT>
struct FastWalker
{
	uint32_t node, idx, inv, mti;
	int      level;
	float    px, py, pz, dimf, tmin;
	float    cx, cy, cz, bx, by, bz;
	float    miss_t;
	Hit      hit;

	__device__ __forceinline__ void start(uint32_t root, float miss_time, const Ray& r)
	{
		node = root;
		miss_t = miss_time;
		level = 1;
		idx = r.idx;
		inv = r.inv;
		px = __uint_as_float(r.px); py = __uint_as_float(r.py); pz = __uint_as_float(r.pz);
		dimf = 0.5f;                 // size of the children of the current node = value of the dimension bit
		tmin = 0.0f;
		mti = 8;
		hit.npush = 0;
		// degenerate axes (see above)
		cx = r.cx; cy = r.cy; cz = r.cz; bx = r.bx; by = r.by; bz = r.bz;
		const float ninf = __uint_as_float(0xFF800000u);
		if (cx == ninf) { cx = 0.0f; bx = ninf; }
		if (cy == ninf) { cy = 0.0f; by = ninf; }
		if (cz == ninf) { cz = 0.0f; bz = ninf; }
	}

	__device__ __forceinline__ void miss()
	{
		hit.voxel = 0;
		hit.face = 6;
		hit.t = miss_t;
	}

	// nodes_m1: see traverse(); stack = this ray's parent stack (kMaxDepth entries, caller-owned so
	// that it stays a plain local array)
	__device__ __forceinline__ bool iterate(const uint32_t* __restrict__ nodes_m1, int depth, uint32_t* stack)
	{
		return iterate_staged(nodes_m1, depth, stack, nullptr, 0u);
	}

	// s_nodes_m1 / n_staged: the first n_staged node ids (the DAG's upper levels: a contiguous prefix of the
	// level-ordered array) are also held in shared memory, laid out like nodes_m1.  n_staged = 0: global only.
	__device__ __forceinline__ bool iterate_staged(const uint32_t* __restrict__ nodes_m1, int depth, uint32_t* stack,
	                                               const uint32_t* s_nodes_m1, uint32_t n_staged)
	{
		const uint32_t child = load_child(nodes_m1, s_nodes_m1, n_staged);
		return child ? descend(child, depth, stack) : advance(stack);
	}

	// PUSH's load (och_h_octree.h:344)
	__device__ __forceinline__ uint32_t load_child(const uint32_t* __restrict__ nodes_m1, const uint32_t* s_nodes_m1 = nullptr, uint32_t n_staged = 0u)
	{
		if (COUNT) ++hit.npush;
		const uint32_t word = node * 8u + (idx ^ inv);                             // id < 2^29: the word index fits 32 bits
		if (n_staged != 0u && node <= n_staged)
			return s_nodes_m1[word];
		return __ldg(nodes_m1 + word);
	}

	// PUSH with a non-empty child: HIT at the last level (returns true), else go down one level
	__device__ __forceinline__ bool descend(uint32_t child, int depth, uint32_t* stack)
	{
		if (level == depth)
		{
			hit.voxel = child;
			hit.face = (mti >> 1) + 3u * ((inv & mti) == 0u);
			hit.t = tmin;
			return true;
		}
		stack[level - 1] = node + (idx << 29);
		++level;
		node = child;
		dimf *= 0.5f;
		const float mx = px + dimf, my = py + dimf, mz = pz + dimf;          // exact
		const float tx = __fmaf_rn(mx, cx, bx);
		const float ty = __fmaf_rn(my, cy, by);
		const float tz = __fmaf_rn(mz, cz, bz);
		idx = 0;
		if (tx >= tmin) { px = mx; idx += 1u; }
		if (ty >= tmin) { py = my; idx += 2u; }
		if (tz >= tmin) { pz = mz; idx += 4u; }
		return false;
	}

	// PUSH with an empty child: STEP to the sibling across the nearest exit plane, POPping as far as needed;
	// returns true on MISS
	__device__ __forceinline__ bool advance(uint32_t* stack)
	{
		// the child slot is empty: leave this cell through its nearest exit plane
		bool ax, ay;
		for (;;)
		{
			const uint32_t tx = __float_as_uint(__fmaf_rn(px, cx, bx));
			const uint32_t ty = __float_as_uint(__fmaf_rn(py, cy, by));
			const uint32_t tz = __float_as_uint(__fmaf_rn(pz, cz, bz));
			const uint32_t tm = min(tx, min(ty, tz));
			tmin = __uint_as_float(tm);
			ax = tx == tm;
			ay = !ax && ty == tm;
			mti = ax ? 1u : (ay ? 2u : 4u);

			if (idx & mti)
				break;                                                          // a sibling lies that way

			if (((tx | ty | tz) & 0x80000000u) == 0u)
			{
				// Multi-level POP.  The reference pops ONE level and re-runs STEP at the parent's corner.  While
				// every t is a non-negative float (unsigned bit order == float order) that re-run cannot change
				// anything: on the exit axis a* the parent's corner coordinate is the child's (its idx bit is 0),
				// on the other axes the corner only moves towards smaller coordinates, i.e. t only grows, and a
				// lower-indexed axis that lost strictly keeps losing.  So tmin and a* are fixed and the chain of
				// POPs runs exactly until an ancestor whose idx bit on a* is 1 -- the lowest set mantissa bit of
				// pos[a*] above the current level.  One FLO replaces the loop; rays with a negative or -inf t in
				// play take the one-level path below, which is the reference's sequence verbatim.
				const uint32_t pa = __float_as_uint(ax ? px : (ay ? py : pz));
				// bits of levels level-1, level-2, .. on axis a*; the exponent (127, odd) lands right above them,
				// so "no ancestor bit set" shows up as level reaching 0
				level -= __ffs(static_cast<int>(pa >> (24 - level)));
				if (level == 0)
				{
					miss();                                                         // popped through the root
					return true;
				}
				const uint32_t keep = 0xFFFFFFFFu << (23 - level);              // drop the position bits of the levels left
				px = __uint_as_float(__float_as_uint(px) & keep);
				py = __uint_as_float(__float_as_uint(py) & keep);
				pz = __uint_as_float(__float_as_uint(pz) & keep);
				dimf = __uint_as_float(static_cast<uint32_t>(127 - level) << 23);
				const uint32_t e = stack[level - 1];
				node = e & kIdMask;
				idx = e >> 29;                                                  // bit a* is set there
				break;
			}

			if (--level == 0)
			{
				miss();
				return true;
			}
			if (idx & 1u) px -= dimf;                                           // back to the parent's corner
			if (idx & 2u) py -= dimf;
			if (idx & 4u) pz -= dimf;
			dimf += dimf;
			const uint32_t e = stack[level - 1];
			node = e & kIdMask;
			idx = e >> 29;
		}

		// step to the sibling across the exit plane (exact: the bit is set)
		if (ax) px -= dimf;
		else if (ay) py -= dimf;
		else pz -= dimf;
		idx ^= mti;
		return false;
	}
};

// ------------------------------------------------------------------------------------------------
// Tight variant.  Same decisions and the same FMAs as FastWalker; what changes is bookkeeping that the SASS of
// FastWalker's loop showed to be avoidable (profiles/r1_v4_ncu_full.md: the loop is issue-bound, so every
// instruction of the round counts):
//   * the node is carried as its word index (id * 8): the child-slot address is one LOP3 (node8 | (idx ^ inv))
//     plus the 64-bit scale-add, and a parent-stack entry is node8 | idx (the low three bits are free), so POP
//     restores node and idx with two masks and ids are no longer squeezed below 2^29 by the stack format
//     (the 32-bit word index still caps ids at 2^29);
//   * the step to the sibling is three predicated FADDs instead of an if / else-if ladder;
//   * the child index after a descend is assembled from the three compares without a branch.
// ------------------------------------------------------------------------------------------------
template<bool COUNT>
struct TightWalker
{
	uint32_t node8, idx, inv, mti;
	int      level;
	float    px, py, pz, dimf, tmin;
	float    cx, cy, cz, bx, by, bz;
	float    miss_t;
	Hit      hit;

	__device__ __forceinline__ void start(uint32_t root, float miss_time, const Ray& r)
	{
		node8 = root << 3;
		miss_t = miss_time;
		level = 1;
		idx = r.idx;
		inv = r.inv;
		px = __uint_as_float(r.px); py = __uint_as_float(r.py); pz = __uint_as_float(r.pz);
		dimf = 0.5f;
		tmin = 0.0f;
		mti = 8;
		hit.npush = 0;
		cx = r.cx; cy = r.cy; cz = r.cz; bx = r.bx; by = r.by; bz = r.bz;
		const float ninf = __uint_as_float(0xFF800000u);      // degenerate axes: see FastWalker
		if (cx == ninf) { cx = 0.0f; bx = ninf; }
		if (cy == ninf) { cy = 0.0f; by = ninf; }
		if (cz == ninf) { cz = 0.0f; bz = ninf; }
	}

	__device__ __forceinline__ void miss()
	{
		hit.voxel = 0;
		hit.face = 6;
		hit.t = miss_t;
	}

	// PUSH's load (och_h_octree.h:344)
	__device__ __forceinline__ uint32_t load_child(const uint32_t* __restrict__ nodes_m1)
	{
		if (COUNT) ++hit.npush;
		return __ldg(nodes_m1 + (node8 | (idx ^ inv)));
	}

	// PUSH with a non-empty child: HIT at the last level (returns true), else go down one level
	__device__ __forceinline__ bool descend(uint32_t child, int depth, uint32_t* stack)
	{
		if (level == depth)
		{
			hit.voxel = child;
			hit.face = (mti >> 1) + 3u * ((inv & mti) == 0u);
			hit.t = tmin;
			return true;
		}
		stack[level - 1] = node8 | idx;
		++level;
		node8 = child << 3;
		dimf *= 0.5f;
		const float mx = px + dimf, my = py + dimf, mz = pz + dimf;          // exact
		const bool ux = __fmaf_rn(mx, cx, bx) >= tmin;
		const bool uy = __fmaf_rn(my, cy, by) >= tmin;
		const bool uz = __fmaf_rn(mz, cz, bz) >= tmin;
		px = ux ? mx : px;
		py = uy ? my : py;
		pz = uz ? mz : pz;
		idx = static_cast<uint32_t>(ux) | (static_cast<uint32_t>(uy) << 1) | (static_cast<uint32_t>(uz) << 2);
		return false;
	}

	// PUSH with an empty child: STEP to the sibling across the nearest exit plane, POPping as far as needed;
	// returns true on MISS
	__device__ __forceinline__ bool advance(const uint32_t* stack)
	{
		bool ax, ay;
		for (;;)
		{
			const uint32_t tx = __float_as_uint(__fmaf_rn(px, cx, bx));
			const uint32_t ty = __float_as_uint(__fmaf_rn(py, cy, by));
			const uint32_t tz = __float_as_uint(__fmaf_rn(pz, cz, bz));
			const uint32_t tyz = min(ty, tz);
			ax = tx <= tyz;                                                     // unsigned argmin, ties -> x, y, z (:388-406)
			ay = !ax && ty <= tz;
			tmin = __uint_as_float(min(tx, tyz));
			mti = ax ? 1u : (ay ? 2u : 4u);

			if (idx & mti)
				break;                                                          // a sibling lies that way

			if (((tx | ty | tz) & 0x80000000u) == 0u)
			{
				// multi-level POP (proof in FastWalker::advance)
				const uint32_t pa = __float_as_uint(ax ? px : (ay ? py : pz));
				level -= __ffs(static_cast<int>(pa >> (24 - level)));
				if (level == 0)
				{
					miss();
					return true;
				}
				const uint32_t keep = 0xFFFFFFFFu << (23 - level);
				px = __uint_as_float(__float_as_uint(px) & keep);
				py = __uint_as_float(__float_as_uint(py) & keep);
				pz = __uint_as_float(__float_as_uint(pz) & keep);
				dimf = __uint_as_float(static_cast<uint32_t>(127 - level) << 23);
				const uint32_t e = stack[level - 1];
				node8 = e & ~7u;
				idx = e & 7u;                                                   // bit a* is set there
				break;
			}

			if (--level == 0)
			{
				miss();
				return true;
			}
			if (idx & 1u) px -= dimf;                                           // back to the parent's corner
			if (idx & 2u) py -= dimf;
			if (idx & 4u) pz -= dimf;
			dimf += dimf;
			const uint32_t e = stack[level - 1];
			node8 = e & ~7u;
			idx = e & 7u;
		}

		// step to the sibling across the exit plane (exact: the bit is set)
		const bool az = !(ax | ay);
		if (ax) px -= dimf;
		if (ay) py -= dimf;
		if (az) pz -= dimf;
		idx ^= mti;
		return false;
	}
};

// ------------------------------------------------------------------------------------------------
// ALU-lean variant ("PipeWalker").  Probe kernels (ort_kernels.cuh, variants 8-11) showed how the loop is bound:
// six extra FMA-pipe instructions per round cost +7.9 %, six extra ALU-pipe instructions +16.7 % -- an ALU-pipe
// instruction (LOP3, SEL, FSEL, ISETP, FSETP, VIMNMX, SHF ...: one warp instruction per two cycles) costs twice an
// FMA-pipe one.  Same decisions and the same t values as FastWalker; the bookkeeping moves off the ALU pipe:
//   * child pick after a descend: `set.ge.f32` yields 1.0f / 0.0f, the position takes the half step by an exact
//     FMA (p + u * size) and the child-slot index is accumulated in float, already XORed with inv_signs:
//     idx' = inv + sum_a w_a u_a with w_a = +-2^a, kept as F = 2^23 + idx' (every partial sum is a small integer at
//     ulp 1, so the FMAs are exact).  bits(F) = 0x4B000000 | idx' goes straight into the 64-bit address IMAD; the
//     constant is folded into the base pointer.  No FSETP / FSEL / SEL / LOP3 for the pick;
//   * a sibling step is three predicated FADD pairs (position, F);
//   * the parent stack holds node id and F in two local arrays -- POP restores both with loads, no unpacking;
//   * multi-level POP: the position bits below the current level are zero and the current level's bit on the exit
//     axis is clear (that is why we pop), so the ancestor to resume at is simply the LOWEST set bit b of the exit
//     axis' position word (the exponent's lowest bit is the sentinel for "through the root"):  b = pa & -pa,
//     positions &= -b, cell size = as_float(0x3F800000 | b) - 1, level = 23 - flo(b).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kMagicBits = 0x4B000000u;     // bits of 8388608.0f = 2^23

__device__ __forceinline__ float set_ge(float a, float b)    // 1.0f if a >= b (ordered) else 0.0f -- one instruction
{
#ifdef ORT_HOST_EMU
	return a >= b ? 1.0f : 0.0f;
#else
	float r;
	asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
	return r;
#endif
}

template<bool COUNT>
struct PipeWalker
{
	uint32_t node, inv, mti;
	int      level;
	float    F;                      // 2^23 + ((child index) ^ inv)
	float    px, py, pz, dimf, tmin;
	float    cx, cy, cz, bx, by, bz;
	float    wx, wy, wz, c0;         // idx' = inv + wx*ux + wy*uy + wz*uz;  c0 = 2^23 + inv
	float    miss_t;
	Hit      hit;

	__device__ __forceinline__ void start(uint32_t root, float miss_time, const Ray& r)
	{
		node = root;
		miss_t = miss_time;
		level = 1;
		inv = r.inv;
		px = __uint_as_float(r.px); py = __uint_as_float(r.py); pz = __uint_as_float(r.pz);
		dimf = 0.5f;
		tmin = 0.0f;
		mti = 8;
		hit.npush = 0;
		cx = r.cx; cy = r.cy; cz = r.cz; bx = r.bx; by = r.by; bz = r.bz;
		const float ninf = __uint_as_float(0xFF800000u);      // degenerate axes: see FastWalker
		if (cx == ninf) { cx = 0.0f; bx = ninf; }
		if (cy == ninf) { cy = 0.0f; by = ninf; }
		if (cz == ninf) { cz = 0.0f; bz = ninf; }
		wx = (inv & 1u) ? -1.0f : 1.0f;
		wy = (inv & 2u) ? -2.0f : 2.0f;
		wz = (inv & 4u) ? -4.0f : 4.0f;
		c0 = __uint_as_float(kMagicBits | inv);
		F = __uint_as_float(kMagicBits | (r.idx ^ inv));
	}

	__device__ __forceinline__ void miss()
	{
		hit.voxel = 0;
		hit.face = 6;
		hit.t = miss_t;
	}

	// PUSH's load (och_h_octree.h:344).  base_biased = address of nodes_m1 minus 4 * kMagicBits (computed on the host),
	// so that bits(F) = kMagicBits | idx' can be used as the word offset as it is: two 64-bit IMADs, no logic op
	__device__ __forceinline__ uint32_t load_child(unsigned long long base_biased)
	{
		if (COUNT) ++hit.npush;
		const unsigned long long a = base_biased + static_cast<unsigned long long>(node) * 32ull + static_cast<unsigned long long>(__float_as_uint(F)) * 4ull;
		return __ldg(reinterpret_cast<const uint32_t*>(a));
	}

	// PUSH with a non-empty child: HIT at the last level (returns true), else go down one level
	__device__ __forceinline__ bool descend(uint32_t child, int depth, uint32_t* stack_n, float* stack_f)
	{
		if (level == depth)
		{
			hit.voxel = child;
			hit.face = (mti >> 1) + 3u * ((inv & mti) == 0u);
			hit.t = tmin;
			return true;
		}
		stack_n[level - 1] = node;
		stack_f[level - 1] = F;
		++level;
		node = child;
		dimf *= 0.5f;
		const float ux = set_ge(__fmaf_rn(px + dimf, cx, bx), tmin);            // px + dimf is exact
		const float uy = set_ge(__fmaf_rn(py + dimf, cy, by), tmin);
		const float uz = set_ge(__fmaf_rn(pz + dimf, cz, bz), tmin);
		px = __fmaf_rn(ux, dimf, px);                                              // exact: + size or + 0
		py = __fmaf_rn(uy, dimf, py);
		pz = __fmaf_rn(uz, dimf, pz);
		F = __fmaf_rn(uz, wz, __fmaf_rn(uy, wy, __fmaf_rn(ux, wx, c0)));           // exact small integers at ulp 1
		return false;
	}

	// PUSH with an empty child: STEP to the sibling across the nearest exit plane, POPping as far as needed;
	// returns true on MISS
	__device__ __forceinline__ bool advance(const uint32_t* stack_n, const float* stack_f)
	{
		bool ax, ay;
		for (;;)
		{
			const uint32_t tx = __float_as_uint(__fmaf_rn(px, cx, bx));
			const uint32_t ty = __float_as_uint(__fmaf_rn(py, cy, by));
			const uint32_t tz = __float_as_uint(__fmaf_rn(pz, cz, bz));
			const uint32_t tyz = min(ty, tz);
			ax = tx <= tyz;                                                     // unsigned argmin, ties -> x, y, z (:388-406)
			ay = !ax && ty <= tz;
			tmin = __uint_as_float(min(tx, tyz));
			mti = 4u;
			if (ay) mti = 2u;
			if (ax) mti = 1u;

			if (((__float_as_uint(F) ^ inv) & mti) != 0u)
				break;                                                          // a sibling lies that way

			if (((tx | ty | tz) & 0x80000000u) == 0u)
			{
				// multi-level POP (proof of the shortcut in FastWalker::advance; the bit trick is explained above)
				uint32_t pa = __float_as_uint(pz);
				if (ay) pa = __float_as_uint(py);
				if (ax) pa = __float_as_uint(px);
				const uint32_t b = pa & (0u - pa);
				if (b == 0x00800000u)
				{
					miss();                                                         // popped through the root
					return true;
				}
				const uint32_t keep = 0u - b;
				px = __uint_as_float(__float_as_uint(px) & keep);
				py = __uint_as_float(__float_as_uint(py) & keep);
				pz = __uint_as_float(__float_as_uint(pz) & keep);
				dimf = __uint_as_float(0x3F800000u | b) - 1.0f;                 // b * 2^-23, exact
				level = __clz(static_cast<int>(b)) - 8;                         // 23 - flo(b)
				node = stack_n[level - 1];
				F = stack_f[level - 1];                                         // bit a* (un-XORed) is set there
				break;
			}

			// a negative or -inf t is in play: the reference's sequence verbatim, one level at a time
			if (--level == 0)
			{
				miss();
				return true;
			}
			const uint32_t u = __float_as_uint(F) ^ inv;
			if (u & 1u) px -= dimf;                                             // back to the parent's corner
			if (u & 2u) py -= dimf;
			if (u & 4u) pz -= dimf;
			dimf += dimf;
			node = stack_n[level - 1];
			F = stack_f[level - 1];
		}

		// step to the sibling across the exit plane (exact: the bit is set)
		if (ax) { px -= dimf; F -= wx; }
		else if (ay) { py -= dimf; F -= wy; }
		else { pz -= dimf; F -= wz; }
		return false;
	}
};

template<bool COUNT>
__device__ __forceinline__ Hit traverse_fast(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, const Ray& r, uint32_t* stack)
{
	FastWalker<COUNT> w;
	w.start(root, miss_t, r);
	while (!w.iterate(nodes_m1, depth, stack)) {}
	return w.hit;
}

// the fast path's preconditions: origin inside [1,2)^3, a start position in [1,2)^3, no NaN reciprocal and at least one
// non-degenerate direction component.
//  * The second is not implied by the first: a coordinate of exactly 1.0f on an axis travelled in the positive direction
//    is mirrored to |3 - 1| = 2.0f, whose masked bits (och_h_octree.h:320) are 0 -- a position the reference then walks
//    as raw bit patterns (denormals), which only traverse() reproduces.
//  * A NaN direction component gives a NaN coef and NaN t values.  x86 keeps the operand's sign (negative here: bit
//    patterns that sort last), the GPU's FMA returns 0x7FFFFFFF (sorts before every negative t); traverse()
//    canonicalises NaN like x86, the fast walkers do not.
// coef always carries the sign bit (och_h_octree.h:312), so as unsigned bits: regular < 0xFF800000 (-inf) < NaN.
__device__ __forceinline__ bool fast_path_ok(float ox, float oy, float oz, const Ray& r)
{
	const uint32_t ninf = 0xFF800000u;
	const uint32_t cx = __float_as_uint(r.cx), cy = __float_as_uint(r.cy), cz = __float_as_uint(r.cz);
	const bool no_nan = max(cx, max(cy, cz)) <= ninf;
	const bool some_regular = min(cx, min(cy, cz)) < ninf;                            // not all three degenerate (-inf)
	const bool pos_in_cube = (r.px & r.py & r.pz & 0x3F800000u) == 0x3F800000u;      // pos is masked to 0x3FC00000: exponent field 127
	return in_unit_cube(ox, oy, oz) & pos_in_cube & no_nan & some_regular;
}

// The walk of one ray, start to end.  VARIANT 0: the baseline transliteration; otherwise FastWalker where its
// preconditions hold (origin in [1,2)^3, a non-degenerate direction component) and the baseline for the rest.
// The loop is spelled out here -- load, then descend or advance -- rather than through FastWalker::iterate: the same
// instructions, but this shape schedules ~2 % better (measured: 15.2 vs 14.9 Grays/s on the bench step).
template<int VARIANT, bool COUNT>
__device__ __forceinline__ Hit traverse_variant(const uint32_t* __restrict__ nodes_m1, uint32_t root, int depth, float miss_t, float ox, float oy, float oz, const Ray& r)
{
	uint32_t stack[kMaxDepth];       // parent stack, local memory (dynamically indexed)
	if (VARIANT != 0 && fast_path_ok(ox, oy, oz, r))
	{
		FastWalker<COUNT> w;
		w.start(root, miss_t, r);
		for (;;)
		{
			const uint32_t child = w.load_child(nodes_m1);
			if (child ? w.descend(child, depth, stack) : w.advance(stack))
				break;
		}
		return w.hit;
	}
	return traverse(nodes_m1, root, depth, miss_t, r, stack);
}

// Camera ray of pixel (x, y) -- tree_camera::update_position (test_och_h_octree.cpp:119-137) with
// every operation rounded separately, in source order.
struct Camera
{
	float ox, oy, oz;
	float r[9];
	float fov;
	float aspect, vfx, vfy;
};

__device__ __forceinline__ void camera_ray(const Camera& c, int x, int y, float& dx, float& dy, float& dz)
{
	const float u = __fmul_rn(c.aspect, __fsub_rn(__fmul_rn(c.vfx, static_cast<float>(x)), 1.0f));   // :123
	const float v = __fsub_rn(__fmul_rn(c.vfy, static_cast<float>(y)), 1.0f);                        // :125
	const float ru = __fadd_rn(__fadd_rn(__fmul_rn(u, c.r[0]), __fmul_rn(v, c.r[1])), __fmul_rn(c.fov, c.r[2])); // :129-131
	const float rv = __fadd_rn(__fadd_rn(__fmul_rn(u, c.r[3]), __fmul_rn(v, c.r[4])), __fmul_rn(c.fov, c.r[5]));
	const float rw = __fadd_rn(__fadd_rn(__fmul_rn(u, c.r[6]), __fmul_rn(v, c.r[7])), __fmul_rn(c.fov, c.r[8]));
	const float s = __fadd_rn(__fadd_rn(__fmul_rn(ru, ru), __fmul_rn(rv, rv)), __fmul_rn(rw, rw));
	const float rm = __fdiv_rn(1.0f, __fsqrt_rn(s));                                                  // :133
	dx = __fmul_rn(rw, rm);                                                                           // :135
	dy = __fmul_rn(ru, rm);
	dz = __fmul_rn(-rv, rm);
}

// Rows of a frame one launch traces (ort_trace_frame: y0, rows, tile_rows, tile_step) and how the kernels map them.
struct FrameRows
{
	int W, H;
	int y0, rows, tile_rows, tile_step;
	int tile_shape;     // warp tile: 0 = 8x4, 1 = 16x2, 2 = 4x8 (trace_frame_kernel only)
	int band_rotate;    // block row b is traced by blockIdx.y = (b - band_rotate) mod gridDim.y: which 16-row band starts first
	int tile_shift;     // log2(tile_rows) + 1 when tile_rows is a power of two (the row mapping then needs no division), else 0
};

// FrameRows::tile_shift for a tile height (host side)
inline int tile_shift_of(int tile_rows)
{
	if (tile_rows <= 0 || (tile_rows & (tile_rows - 1)) != 0) return 0;
	int s = 0;
	while ((1 << s) < tile_rows) ++s;
	return s + 1;
}

// strip-local row r -> frame row: contiguous strip, or tiles of tile_rows rows every tile_rows * tile_step rows
// (uniform branches: every thread of a launch takes the same one)
__device__ __forceinline__ int frame_row(const FrameRows& fr, int r)
{
	if (fr.tile_step == 1)
		return fr.y0 + r;
	if (fr.tile_shift)
	{
		const int m = (1 << (fr.tile_shift - 1)) - 1;
		return fr.y0 + (r & ~m) * fr.tile_step + (r & m);
	}
	return fr.y0 + (r / fr.tile_rows) * fr.tile_rows * fr.tile_step + r % fr.tile_rows;
}

}  // namespace ort
