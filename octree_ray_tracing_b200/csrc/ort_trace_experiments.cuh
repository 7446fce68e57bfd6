// ort_trace_experiments.cuh -- walkers that were measured and NOT adopted (round 1: variants 5/6 TightWalker, 7 PipeWalker).
// Compiled into libort_b200_exp.so (-DORT_EXPERIMENTS) and into the host emulation (tests/host_emu) only; the product
// library carries FastWalker / LeanWalker / traverse() from ort_trace.cuh.
#pragma once

#include "ort_trace.cuh"

namespace ort {

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


// ------------------------------------------------------------------------------------------------
// Round-2 experiments on top of LeanWalker (same state, same preconditions, same results).
// ------------------------------------------------------------------------------------------------

// Variant 14, "straight-line round": the descend / advance if-else of the round removed, as the round-1 review asked.
// The position takes the half step only when the child exists (h = child ? size/2 : 0), ONE FMA triple serves both
// the child pick and the exit-plane t values, and everything else of both sides is computed for all lanes and
// committed through selects; the multi-level POP stays the only real branch.
template<bool COUNT>
struct FlatWalker : LeanWalker<COUNT>
{
	using B = LeanWalker<COUNT>;

	template<class STACK>
	__device__ __forceinline__ bool round(unsigned long long base_biased, float leaf_dimf, float miss_t, const STACK st)
	{
		const uint32_t child = B::load_child(base_biased);
		const bool has = child != 0u;
		if (has && B::dimf == leaf_dimf)                                            // HIT (och_h_octree.h:346)
		{
			const uint32_t inv = __float_as_uint(B::c0);
			B::hit.voxel = child;
			B::hit.face = (B::mti >> 1) + 3u * ((inv & B::mti) == 0u);
			B::hit.t = B::tmin;
			return true;
		}
		if (has) st.store(B::dimf, B::w);
		const float half = B::dimf * 0.5f;
		const float h = has ? half : 0.0f;
		const float qx = B::px + h, qy = B::py + h, qz = B::pz + h;                 // exact
		const float tx = __fmaf_rn(qx, B::cx, B::bx);                               // the one FMA triple
		const float ty = __fmaf_rn(qy, B::cy, B::by);
		const float tz = __fmaf_rn(qz, B::cz, B::bz);
		// descend side (h = 0 on the advance side: the position does not move there)
		const float ux = set_ge(tx, B::tmin), uy = set_ge(ty, B::tmin), uz = set_ge(tz, B::tmin);
		B::px = __fmaf_rn(ux, h, B::px);
		B::py = __fmaf_rn(uy, h, B::py);
		B::pz = __fmaf_rn(uz, h, B::pz);
		const float F = __fmaf_rn(uz, B::wz, __fmaf_rn(uy, B::wy, __fmaf_rn(ux, B::wx, B::c0)));
		const uint32_t w_down = child * 8u + __float_as_uint(F);
		// advance side
		const uint32_t ix = __float_as_uint(tx), iy = __float_as_uint(ty), iz = __float_as_uint(tz);
		const uint32_t iyz = min(iy, iz);
		const bool ax = ix <= iyz;
		const bool ay = !ax && iy <= iz;
		const uint32_t m = ax ? 1u : (ay ? 2u : 4u);
		const bool sib = ((B::w ^ __float_as_uint(B::c0)) & m) != 0u;
		// commit
		B::tmin = has ? B::tmin : __uint_as_float(min(ix, iyz));
		B::mti = has ? B::mti : m;
		B::dimf = has ? half : B::dimf;
		if (!has && !sib)
		{
			uint32_t pa = __float_as_uint(B::pz);
			if (ay) pa = __float_as_uint(B::py);
			if (ax) pa = __float_as_uint(B::px);
			const uint32_t b = pa & (0u - pa);
			if (b == 0x00800000u)
			{
				B::hit.voxel = 0;
				B::hit.face = 6;
				B::hit.t = miss_t;
				return true;
			}
			const uint32_t keep = 0u - b;
			B::px = __uint_as_float(__float_as_uint(B::px) & keep);
			B::py = __uint_as_float(__float_as_uint(B::py) & keep);
			B::pz = __uint_as_float(__float_as_uint(B::pz) & keep);
			B::dimf = __uint_as_float(0x3F800000u | b) - 1.0f;
			B::w = st.load(B::dimf);
		}
		const float s = has ? 0.0f : B::dimf;                                       // sibling step only on the advance side
		if (ax) B::px -= s;
		if (ay) B::py -= s;
		if (!(ax | ay)) B::pz -= s;
		B::w = has ? w_down : (B::w ^ m);
		return false;
	}
};

// Variant 15, "128-bit node fetches": the PUSH load fetches the 16-byte half node that holds the child slot with one
// ld.global.v4 and picks the child with the two low bits of the slot index.
template<bool COUNT>
struct V4Walker : LeanWalker<COUNT>
{
	using B = LeanWalker<COUNT>;

	template<class STACK>
	__device__ __forceinline__ bool round(unsigned long long base_biased, float leaf_dimf, float miss_t, const STACK st)
	{
		if (COUNT) ++B::hit.npush;
		const uint4 hn = __ldg(reinterpret_cast<const uint4*>(base_biased + static_cast<unsigned long long>(B::w & ~3u) * 4ull));
		const uint32_t lo = (B::w & 1u) ? hn.y : hn.x, hi = (B::w & 1u) ? hn.w : hn.z;
		const uint32_t child = (B::w & 2u) ? hi : lo;
		return child ? B::descend(child, leaf_dimf, st) : B::advance(miss_t, st);
	}
};

}  // namespace ort
