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
#include <cstring>
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

// RCPPS as a table function: the top log2n mantissa bits pick the entry, the exponent is handled
// arithmetically.  Every entry is the reciprocal of a number in [1, 2), i.e. lies in (0.5, 1] with exponent field 126 or
// 127 -- ort_set_rcp_table() refuses tables that break this -- which is what lets the common case run without a check
// on the result's exponent.
__device__ __forceinline__ uint32_t rcp_model(const RcpTable rt, uint32_t x)
{
	const uint32_t xe = x & 0x7F800000u;                                   // the exponent field e, in place
	if (xe - 0x00800000u < 0x7E000000u)
	{
		// e in 1..252: result exponent field = (126 | 127) - (e - 127) >= 1, no underflow; r - ((e - 127) << 23) in one add
		const uint32_t r = __ldg(rt.tab + ((x & 0x7FFFFFu) >> rt.shift));
		return (x & 0x80000000u) | (r + 0x3F800000u - xe);
	}
	const uint32_t sign = x & 0x80000000u;
	const uint32_t e = xe >> 23;
	const uint32_t m = x & 0x7FFFFFu;
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
	uint32_t t0or;           // OR over the axes of bits(fma(o_a, coef_a, bias_a)): the sign bit tells whether a t value can ever be negative (LeanWalker)
};

// t0or: the t of the plane through the origin itself only exists as a cell plane when the origin coordinate lies on the
// finest level's grid, i.e. when its mantissa bits below that level (plane_mask) are clear (see LeanWalker /
// lean_path_ok).  CAMERA = false: tested per ray with `arg` = plane_mask.  CAMERA = true: the rays of a launch share the
// origin, `arg` != 0 says "on the grid" for the whole launch (a warp-uniform branch; camera_origin_flags()).
template<bool CAMERA>
__device__ __forceinline__ void ray_axis(const RcpTable rt, float o, float d, int a, float& coef, float& bias, uint32_t& pos, uint32_t& inv, uint32_t& idx, uint32_t& t0or, uint32_t arg)
{
	const bool sg = 0.0f < d;                                                  // :310
	inv |= static_cast<uint32_t>(sg) << a;                                     // :322
	const uint32_t dn = __float_as_uint(d) | 0x80000000u;                      // :312
	const float oa = fabsf(__fsub_rn(sg ? 3.0f : 0.0f, o));                    // :314
	coef = __uint_as_float(rcp_model(rt, dn));                                 // :316
	bias = __uint_as_float(__float_as_uint(__fmul_rn(coef, oa)) ^ 0x80000000u); // :318
	pos = __float_as_uint(oa) & 0x3FC00000u;                                   // :320
	idx |= static_cast<uint32_t>(pos == 0x3FC00000u) << a;                     // :324
	if (CAMERA ? arg != 0u : (__float_as_uint(oa) & arg) == 0u)
		t0or |= __float_as_uint(__fmaf_rn(oa, coef, bias));
}

// plane_mask = (1 << (23 - depth)) - 1, the mantissa bits below the finest level's grid (0: treat every origin as on-grid)
__device__ __forceinline__ Ray ray_setup(const RcpTable rt, float ox, float oy, float oz, float dx, float dy, float dz, uint32_t plane_mask = 0u)
{
	Ray r;
	r.inv = 0;
	r.idx = 0;
	r.t0or = 0;
	ray_axis<false>(rt, ox, dx, 0, r.cx, r.bx, r.px, r.inv, r.idx, r.t0or, plane_mask);
	ray_axis<false>(rt, oy, dy, 1, r.cy, r.by, r.py, r.inv, r.idx, r.t0or, plane_mask);
	ray_axis<false>(rt, oz, dz, 2, r.cz, r.bz, r.pz, r.inv, r.idx, r.t0or, plane_mask);
	return r;
}

// the rays of a camera frame: origin_flags = camera_origin_flags() of the shared origin (bits 0-2: coordinate on the grid)
__device__ __forceinline__ Ray ray_setup_camera(const RcpTable rt, float ox, float oy, float oz, float dx, float dy, float dz, uint32_t origin_flags)
{
	Ray r;
	r.inv = 0;
	r.idx = 0;
	r.t0or = 0;
	ray_axis<true>(rt, ox, dx, 0, r.cx, r.bx, r.px, r.inv, r.idx, r.t0or, origin_flags & 1u);
	ray_axis<true>(rt, oy, dy, 1, r.cy, r.by, r.py, r.inv, r.idx, r.t0or, origin_flags & 2u);
	ray_axis<true>(rt, oz, dz, 2, r.cz, r.bz, r.pz, r.inv, r.idx, r.t0or, origin_flags & 4u);
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

// child pick without predicates: set.ge.f32 and a 2^23 'magic' float that carries the slot index (LeanWalker; PipeWalker in
// ort_trace_experiments.cuh)
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

// ------------------------------------------------------------------------------------------------
// Lean variant (round 2, the default for frames).  Same decisions, same FMAs, same t values as FastWalker; built from
// what round 1's profile said about the loop: it is bound by warp-instruction issue (5.8 eligible warps per cycle,
// ALU pipe at half rate), so every instruction of the round counts, ALU-pipe instructions twice.
//   * ONE word of state for "where am I": w = node * 8 + ((idx ^ inv_signs) | 2^23-magic), the word offset of the
//     current child slot.  The PUSH load is base[w] (one IMAD.WIDE + LDG), a parent-stack entry is w itself (no
//     packing on the way in, no unpacking on the way out), the sibling step is w ^= min_t_idx, and the sibling test reads
//     bit a* of (w ^ inv).  The child pick after a descend needs no predicate: set.ge.f32 gives 1.0f / 0.0f, position and
//     slot index take the half step through exact FMAs (FMA pipe, full rate), and bits(F) = magic | idx' is added to
//     child * 8 by the same IMAD that scales the id; the magic constant is folded into the base pointer on the host.
//   * no `level` register: the size of the current node's children (dimf = 2^-level) says it all.  HIT test: dimf ==
//     2^-depth.  Parent-stack address: the exponent field of dimf, scaled by a shift (one LEA.HI), into a per-thread
//     column of shared memory (no 64-bit local addressing, no unpack).  POP restores dimf from the lowest set position
//     bit and finds its stack entry the same way.
//   * no negative-t branch in the loop.  FastWalker tests "is a negative t in play" before every multi-level POP and
//     carries the reference's one-level POP sequence for that case.  A t value is fma(X, coef, bias) for a cell plane
//     X on the ray's side of the (mirrored) origin o: X <= o.  For X < o the exact value |coef| * (o - X) - e, with e the
//     rounding error of bias = -(coef * o), is positive, because o - X >= 2^-23 and |e| < 2^-23 * |coef|; coef = -0
//     (huge |d|) gives t = +0 and an overflowing bias gives t = +inf.  So a negative t can only be the t of the plane
//     through the origin itself, X == o, and that one value per axis is computed at ray set-up (Ray::t0or).  Rays
//     whose three values are sign-clear and whose three reciprocals are regular (no -inf: d = +-0 / denormal) can
//     never see a negative t: they take this walker, the others take FastWalker (exact for those, as before).
// ------------------------------------------------------------------------------------------------

// parent stack of one thread.  Device: a column of shared memory, entry of level L (children of size 2^-L) at byte
// address base + (bits(2^-L) >> SHIFT), i.e. exponent field (127 - L) scaled by the column pitch 2^(23 - SHIFT) bytes.
// Host emulation: a plain array indexed by the exponent field.
template<int SHIFT>
struct LeanStack
{
#ifdef ORT_HOST_EMU
	uint32_t* e;                                     // kMaxDepth entries
	__device__ __forceinline__ void store(float dimf, uint32_t w) const { e[(__float_as_uint(dimf) >> 23) - (127u - kMaxDepth)] = w; }
	__device__ __forceinline__ uint32_t load(float dimf) const { return e[(__float_as_uint(dimf) >> 23) - (127u - kMaxDepth)]; }
#else
	uint32_t base;                                   // shared-space byte address, biased by the lowest exponent in use
	__device__ __forceinline__ void store(float dimf, uint32_t w) const
	{
		asm volatile("st.shared.u32 [%0], %1;" :: "r"(base + (__float_as_uint(dimf) >> SHIFT)), "r"(w) : "memory");
	}
	__device__ __forceinline__ uint32_t load(float dimf) const
	{
		uint32_t w;
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(base + (__float_as_uint(dimf) >> SHIFT)) : "memory");
		return w;
	}
#endif
};

template<bool COUNT>
struct LeanWalker
{
	uint32_t w, mti;
	float    px, py, pz, dimf, tmin;
	float    cx, cy, cz, bx, by, bz;
	float    wx, wy, wz, c0;         // slot index of a picked child = inv + wx*ux + wy*uy + wz*uz;  c0 = 2^23 + inv (PipeWalker)
	Hit      hit;

	__device__ __forceinline__ void start(uint32_t root, const Ray& r)
	{
		w = root * 8u + (kMagicBits | (r.idx ^ r.inv));
		px = __uint_as_float(r.px); py = __uint_as_float(r.py); pz = __uint_as_float(r.pz);
		dimf = 0.5f;
		tmin = 0.0f;
		mti = 8;
		hit.npush = 0;
		cx = r.cx; cy = r.cy; cz = r.cz; bx = r.bx; by = r.by; bz = r.bz;
		wx = (r.inv & 1u) ? -1.0f : 1.0f;
		wy = (r.inv & 2u) ? -2.0f : 2.0f;
		wz = (r.inv & 4u) ? -4.0f : 4.0f;
		c0 = __uint_as_float(kMagicBits | r.inv);
	}

	// Re-entry into the walk at time tau > 0 (beam start, ort_beam.cuh): descend from the root choosing, per axis, the
	// half whose mid plane the ray has not crossed by tau -- the same `t_mid >= tmin` test descend() applies, with tmin =
	// tau.  The caller continues with the ordinary loop; the first empty slot it meets is the empty slot S_i of the
	// reference walk with the smallest exit time e_i >= tau (proof in ort_beam.cuh), with the same slot word, position,
	// cell size and parent stack; advance() then overwrites tmin and mti before anything reads them.  The stack entry
	// of level 0 does not exist, so the first half step is spelled out here instead of calling descend().
	__device__ __forceinline__ void start_at(uint32_t root, const Ray& r, float tau)
	{
		start(root, r);
		tmin = tau;
		const float ux = set_ge(__fmaf_rn(1.5f, cx, bx), tau);
		const float uy = set_ge(__fmaf_rn(1.5f, cy, by), tau);
		const float uz = set_ge(__fmaf_rn(1.5f, cz, bz), tau);
		px = __fmaf_rn(ux, 0.5f, 1.0f);
		py = __fmaf_rn(uy, 0.5f, 1.0f);
		pz = __fmaf_rn(uz, 0.5f, 1.0f);
		const float F = __fmaf_rn(uz, wz, __fmaf_rn(uy, wy, __fmaf_rn(ux, wx, c0)));
		w = root * 8u + __float_as_uint(F);
	}

	// the time the ray leaves the cube [1,2)^3: the exit time of the walk's last advance (the one that pops through the
	// root) -- the smallest t of the three lower planes X = 1.0
	__device__ __forceinline__ float cube_exit_time() const
	{
		return fminf(__fmaf_rn(1.0f, cx, bx), fminf(__fmaf_rn(1.0f, cy, by), __fmaf_rn(1.0f, cz, bz)));
	}

	// PUSH's load (och_h_octree.h:344).  base_biased = address of nodes_m1 minus 4 * kMagicBits
	__device__ __forceinline__ uint32_t load_child(unsigned long long base_biased)
	{
		if (COUNT) ++hit.npush;
		return __ldg(reinterpret_cast<const uint32_t*>(base_biased + static_cast<unsigned long long>(w) * 4ull));
	}

	// PUSH with a non-empty child (:346-376): HIT at the last level (returns true), else go down one level.
	// leaf_dimf = 2^-depth
	template<class STACK>
	__device__ __forceinline__ bool descend(uint32_t child, float leaf_dimf, const STACK st)
	{
		if (dimf == leaf_dimf)                                                      // :346 HIT
		{
			const uint32_t inv = __float_as_uint(c0);
			hit.voxel = child;
			hit.face = (mti >> 1) + 3u * ((inv & mti) == 0u);
			hit.t = tmin;
			return true;
		}
		st.store(dimf, w);                                                          // :357
		dimf *= 0.5f;                                                               // :361
		const float ux = set_ge(__fmaf_rn(px + dimf, cx, bx), tmin);                // :363-367; px + dimf is exact
		const float uy = set_ge(__fmaf_rn(py + dimf, cy, by), tmin);
		const float uz = set_ge(__fmaf_rn(pz + dimf, cz, bz), tmin);
		px = __fmaf_rn(ux, dimf, px);                                               // :371-373, exact: + size or + 0
		py = __fmaf_rn(uy, dimf, py);
		pz = __fmaf_rn(uz, dimf, pz);
		const float F = __fmaf_rn(uz, wz, __fmaf_rn(uy, wy, __fmaf_rn(ux, wx, c0)));   // exact small integers at ulp 1
		w = child * 8u + __float_as_uint(F);
		return false;
	}

	// PUSH with an empty child: STEP (:378-419) to the sibling across the nearest exit plane, POPping (:421-446) as far as
	// needed; returns true on MISS.  All t are non-negative and none is NaN here (see above): the reference's unsigned
	// order of the bit patterns is the order of the float values, so the argmin is a float min3 + two equality tests
	// (ties -> x, then y, as :388-406), and the multi-level POP (proof in FastWalker::advance) always applies.
	template<class STACK>
	__device__ __forceinline__ bool advance(float miss_t, const STACK st)
	{
		const float tx = __fmaf_rn(px, cx, bx);
		const float ty = __fmaf_rn(py, cy, by);
		const float tz = __fmaf_rn(pz, cz, bz);
		tmin = fminf(tx, fminf(ty, tz));
		const bool ax = tx == tmin;
		const bool ay = !ax && ty == tmin;
		mti = 4u;
		if (ay) mti = 2u;
		if (ax) mti = 1u;

		if (((w ^ __float_as_uint(c0)) & mti) == 0u)                                // :410 no sibling that way
		{
			// The position bits below the current level are zero and the current level's bit on the exit axis is clear
			// (that is why we pop), so the ancestor to resume at is the LOWEST set bit b of the exit axis' position word;
			// the exponent's lowest bit (127 is odd) is the sentinel for "popped through the root".
			uint32_t pa = __float_as_uint(pz);
			if (ay) pa = __float_as_uint(py);
			if (ax) pa = __float_as_uint(px);
			const uint32_t na = 0u - pa;
			const uint32_t b = pa & na;
			if (b == 0x00800000u)                                                   // :423
			{
				hit.voxel = 0;
				hit.face = 6;
				hit.t = miss_t;
				return true;
			}
			const uint32_t keep = pa | na;                                          // = -b: drops the position bits of the levels left
			px = __uint_as_float(__float_as_uint(px) & keep);
			py = __uint_as_float(__float_as_uint(py) & keep);
			pz = __uint_as_float(__float_as_uint(pz) & keep);
			dimf = __uint_as_float(b) * 0x1p126f;                                   // b * 2^-23 (b as a denormal x 2^126), exact
			w = st.load(dimf);                                                      // bit a* of its index is set
		}

		// step to the sibling across the exit plane (:414-419; exact: the bit is set)
		if (ax) px -= dimf;
		if (ay) py -= dimf;
		if (!(ax | ay)) pz -= dimf;
		w ^= mti;
		return false;
	}

	// One round of the reference's PUSH label plus the STEP / POP work that follows an empty slot.  Returns true when
	// the ray is finished (result in `hit`).
	template<class STACK>
	__device__ __forceinline__ bool round(unsigned long long base_biased, float leaf_dimf, float miss_t, const STACK st)
	{
		const uint32_t child = load_child(base_biased);
		return child ? descend(child, leaf_dimf, st) : advance(miss_t, st);
	}
};

// the origin part of fast_path_ok: origin inside [1,2)^3 and a start position in [1,2)^3 (see there)
__device__ __forceinline__ bool origin_in_cube(float ox, float oy, float oz, const Ray& r)
{
	return in_unit_cube(ox, oy, oz) & ((r.px & r.py & r.pz & 0x3F800000u) == 0x3F800000u);
}

// LeanWalker's preconditions on top of origin_in_cube: no degenerate axis (every coef regular, i.e. below -inf as unsigned
// bits -- which also rules out NaN) and no negative t at the planes through the origin (Ray::t0or; the argument is in
// LeanWalker's header).  Together they imply fast_path_ok.
__device__ __forceinline__ bool lean_path_ok(const Ray& r)
{
	const uint32_t ninf = 0xFF800000u;
	const uint32_t cx = __float_as_uint(r.cx), cy = __float_as_uint(r.cy), cz = __float_as_uint(r.cz);
	return (max(cx, max(cy, cz)) < ninf) & ((r.t0or & 0x80000000u) == 0u);
}

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
	uint32_t origin_flags;   // camera_origin_flags(): what holds for EVERY ray of the frame because they share the origin
};

// Origin facts of a camera frame, computed once on the host instead of once per ray:
//   bits 0-2  coordinate a lies on the finest level's grid (its mantissa bits below that level are clear) -- only then
//             does a cell plane pass through the origin and the t0 test of the lean tier apply (Ray::t0or);
//   bit 3     the origin passes fast_path_ok's origin tests for every direction: all coordinates in [1, 2) and none
//             exactly 1.0f (which mirrors to 2.0f on an axis travelled in the positive direction and leaves the cube).
constexpr uint32_t kOriginInCube = 8u;
inline uint32_t camera_origin_flags(float ox, float oy, float oz, uint32_t plane_mask)
{
	const float o[3] = { ox, oy, oz };
	uint32_t f = kOriginInCube;
	for (int a = 0; a < 3; ++a)
	{
		uint32_t b;
#ifdef ORT_HOST_EMU
		b = __float_as_uint(o[a]);
#else
		memcpy(&b, &o[a], 4);
#endif
		if ((b >> 23) != 127u || b == 0x3F800000u) f &= ~kOriginInCube;
		if ((b & plane_mask) == 0u) f |= 1u << a;       // (3 - o has the same low mantissa bits clear as o, for o in [1, 2))
	}
	return f;
}

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
	// band schedule from an earlier launch of the same view (ort_device.cu: BandMap): blockIdx.y = k traces band band_order[k]
	// (a permutation of the launch's bands, most expensive first; null: the band_rotate rule), and every warp leaves the
	// longest time one of its lanes took in band_cost[band] (null: not recorded)
	const uint16_t* band_order;
	unsigned*       band_cost;
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
