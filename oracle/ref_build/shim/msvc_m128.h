/* TEST INFRASTRUCTURE ONLY -- used to compile the UNMODIFIED reference headers in
 * place under /root/reference with g++.
 *
 * MSVC's __m128 is a union with named lane members (.m128_f32 / .m128_u32); the
 * reference reads lanes that way (och_h_octree.h:352, :384-386).  g++'s __m128 is
 * a bare vector type.  This shim re-defines the *name* __m128, after <immintrin.h>
 * has been included, to a union that converts implicitly to and from the real
 * vector type, so every intrinsic call still sees the real type and the lane
 * accessors read the same bits MSVC would.  No arithmetic is changed.
 */
#pragma once
#include <immintrin.h>
#include <cmath>      /* INFINITY (och_h_octree.h:429 relies on MSVC pulling it in) */
#include <cstdint>

typedef __m128 och_real_m128;

union och_msvc_m128
{
	och_real_m128 v;
	float         m128_f32[4];
	uint32_t      m128_u32[4];

	och_msvc_m128() {}
	och_msvc_m128(och_real_m128 x) : v(x) {}
	operator och_real_m128() const { return v; }
};

#define __m128 och_msvc_m128
