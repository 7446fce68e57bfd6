// cuda_shim.h -- TEST INFRASTRUCTURE.  Lets g++ compile csrc/ort_trace.cuh (the device-side per-ray traversal) for the
// host, so that the CPU-only test suite can run the very code the kernels run -- FastWalker's float bookkeeping, the
// multi-level POP, the brick walk -- against the oracle on millions of rays.  Nothing here is linked into
// libort_b200.so; the product still has no CPU path.
//
// Every CUDA intrinsic the traversal uses is given its IEEE meaning: the *_rn arithmetic is one correctly rounded
// operation (compiled with -ffp-contract=off, fmaf() is the fused one), the bit casts are memcpy, __ldg is a load.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

#define ORT_HOST_EMU 1

// stand-ins for <cuda_runtime.h>'s decorations
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict

static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
// NaN results: x86 propagates an operand's sign and payload, the GPU's FP32 pipes return the canonical 0x7FFFFFFF --
// the one place where the two machines' IEEE arithmetic differs visibly (the traversal orders t values by their bit
// patterns), so the stand-ins reproduce the device's NaN.
static inline float emu_nan(float r) { return r != r ? __uint_as_float(0x7FFFFFFFu) : r; }
static inline float __fmaf_rn(float a, float b, float c) { return emu_nan(std::fmaf(a, b, c)); }
static inline float __fmul_rn(float a, float b) { return emu_nan(a * b); }
static inline float __fadd_rn(float a, float b) { return emu_nan(a + b); }
static inline float __fsub_rn(float a, float b) { return emu_nan(a - b); }
static inline float __fdiv_rn(float a, float b) { return emu_nan(a / b); }
static inline float __fsqrt_rn(float a) { return emu_nan(std::sqrt(a)); }
// Every global load of the traversal goes through __ldg: the emulator checks each address against the two arrays a
// kernel may read (the node array and the reciprocal table) -- the job compute-sanitizer's memcheck would do on the
// device.  An out-of-range load is counted and yields 0 instead of faulting.
struct EmuBounds
{
	const char* lo[2];
	const char* hi[2];
	unsigned long long violations;
};
inline thread_local EmuBounds g_emu_bounds = { { nullptr, nullptr }, { nullptr, nullptr }, 0 };
template<class T> static inline T __ldg(const T* p)
{
	const char* a = reinterpret_cast<const char*>(p);
	const EmuBounds& b = g_emu_bounds;
	if ((a >= b.lo[0] && a + sizeof(T) <= b.hi[0]) || (a >= b.lo[1] && a + sizeof(T) <= b.hi[1]))
		return *p;
	++g_emu_bounds.violations;
	return T{};
}
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz(static_cast<unsigned>(x)) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
struct uint4 { uint32_t x, y, z, w; };
using std::fminf;
using std::min;
using std::max;
