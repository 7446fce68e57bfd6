// ort_noise.h -- och::simplex_n (och_noise.h:18-367): Gustavson simplex noise in float with int truncation,
// restated once for the host fixture builder (ort_fixture.cpp) and the CUDA fixture kernels (ort_device.cu).
// Every operation is a separately rounded IEEE float operation in the reference's order; both translation units
// are compiled without contraction (nvcc -fmad=false, g++ -ffp-contract=off), so host and device agree bit for bit.
#pragma once

#include <cstdint>

#ifdef __CUDACC__
#define ORT_HD __host__ __device__ __forceinline__
#else
#define ORT_HD inline
#endif

namespace ort_noise {

#define ORT_NOISE_PERM_VALUES \
	151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10, 23, 190, 6, 148, \
	247, 120, 234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87, 174, 20, 125, 136, 171, 168, 68, 175, \
	74, 165, 71, 134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230, 220, 105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54, \
	65, 25, 63, 161, 1, 216, 80, 73, 209, 76, 132, 187, 208, 89, 18, 169, 200, 196, 135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64, \
	52, 217, 226, 250, 124, 123, 5, 202, 38, 147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213, \
	119, 248, 152, 2, 44, 154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232, 178, 185, 112, 104, \
	218, 246, 97, 228, 251, 34, 242, 193, 238, 210, 144, 12, 191, 179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106, 157, \
	184, 84, 204, 176, 115, 121, 50, 45, 127, 4, 150, 254, 138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180

struct G3 { float x, y, z; };

#define ORT_NOISE_GRAD_VALUES \
	{ 1, 1, 0 }, { -1, 1, 0 }, { 1, -1, 0 }, { -1, -1, 0 }, { 1, 0, 1 }, { -1, 0, 1 }, \
	{ 1, 0, -1 }, { -1, 0, -1 }, { 0, 1, 1 }, { 0, -1, 1 }, { 0, 1, -1 }, { 0, -1, -1 }

static const uint8_t kPerm[256] = { ORT_NOISE_PERM_VALUES };
static const G3 kGrad[12] = { ORT_NOISE_GRAD_VALUES };
#ifdef __CUDACC__
static __device__ const uint8_t kPermDev[256] = { ORT_NOISE_PERM_VALUES };
static __device__ const G3 kGradDev[12] = { ORT_NOISE_GRAD_VALUES };
#endif
#ifdef __CUDA_ARCH__
#define ORT_NOISE_PERM kPermDev
#define ORT_NOISE_GRAD kGradDev
#else
#define ORT_NOISE_PERM kPerm
#define ORT_NOISE_GRAD kGrad
#endif

ORT_HD int P(int i) { return ORT_NOISE_PERM[i & 255]; }

ORT_HD float falloff2(float x, float y, int g)
{
	float t = 0.5F - x * x - y * y;
	if (t < 0) return 0.0F;
	t *= t;
	return t * t * (ORT_NOISE_GRAD[g].x * x + ORT_NOISE_GRAD[g].y * y);
}

ORT_HD float simplex2(float freq, float x, float y)                     // och_noise.h:73-179
{
	x *= freq;
	y *= freq;
	const float F2 = 0.5F * (0.73205078F);
	const float G2 = (3.0F - 1.73205078F) / 6.0F;
	const float s = (x + y) * F2;
	const int i = static_cast<int>(x + s), j = static_cast<int>(y + s);
	const float t = static_cast<float>(i + j) * G2;
	const float x0 = x - (static_cast<float>(i) - t), y0 = y - (static_cast<float>(j) - t);
	const int di = x0 > y0 ? 1 : 0, dj = 1 - di;
	const float x1 = x0 - static_cast<float>(di) + G2, y1 = y0 - static_cast<float>(dj) + G2;
	const float x2 = x0 - 1.0F + 2.0F * G2, y2 = y0 - 1.0F + 2.0F * G2;
	const int ii = i & 255, jj = j & 255;
	const float n0 = falloff2(x0, y0, P(ii + P(jj)) % 12);
	const float n1 = falloff2(x1, y1, P(ii + di + P(jj + dj)) % 12);
	const float n2 = falloff2(x2, y2, P(ii + 1 + P(jj + 1)) % 12);
	return 70.0F * (n0 + n1 + n2);
}

ORT_HD float falloff3(float x, float y, float z, int g)
{
	float t = 0.6F - x * x - y * y - z * z;
	if (t < 0) return 0.0F;
	t *= t;
	return t * t * (ORT_NOISE_GRAD[g].x * x + ORT_NOISE_GRAD[g].y * y + ORT_NOISE_GRAD[g].z * z);
}

ORT_HD float simplex3(float freq, float x, float y, float z)            // och_noise.h:181-366
{
	x *= freq; y *= freq; z *= freq;
	const float F3 = 1.0F / 3.0F, G3c = 1.0F / 6.0F;
	const float s = (x + y + z) * F3;
	const int i = static_cast<int>(x + s), j = static_cast<int>(y + s), k = static_cast<int>(z + s);
	const float t = static_cast<float>(i + j + k) * G3c;
	const float x0 = x - (static_cast<float>(i) - t), y0 = y - (static_cast<float>(j) - t), z0 = z - (static_cast<float>(k) - t);

	// rank the three offsets; the second corner steps along the largest, the third along the two largest,
	// with the reference's tie rules (:224-281)
	int a1, b1, c1, a2, b2, c2;
	if (x0 >= y0)
	{
		if (y0 >= z0)      { a1 = 1; b1 = 0; c1 = 0; a2 = 1; b2 = 1; c2 = 0; }
		else if (x0 >= z0) { a1 = 1; b1 = 0; c1 = 0; a2 = 1; b2 = 0; c2 = 1; }
		else               { a1 = 0; b1 = 0; c1 = 1; a2 = 1; b2 = 0; c2 = 1; }
	}
	else
	{
		if (y0 < z0)       { a1 = 0; b1 = 0; c1 = 1; a2 = 0; b2 = 1; c2 = 1; }
		else if (x0 < z0)  { a1 = 0; b1 = 1; c1 = 0; a2 = 0; b2 = 1; c2 = 1; }
		else               { a1 = 0; b1 = 1; c1 = 0; a2 = 1; b2 = 1; c2 = 0; }
	}

	const float x1 = x0 - static_cast<float>(a1) + G3c, y1 = y0 - static_cast<float>(b1) + G3c, z1 = z0 - static_cast<float>(c1) + G3c;
	const float x2 = x0 - static_cast<float>(a2) + G3c * 2.0F, y2 = y0 - static_cast<float>(b2) + G3c * 2.0F, z2 = z0 - static_cast<float>(c2) + G3c * 2.0F;
	const float x3 = x0 - 1.0F + G3c * 3.0F, y3 = y0 - 1.0F + G3c * 3.0F, z3 = z0 - 1.0F + G3c * 3.0F;
	const int ii = i & 255, jj = j & 255, kk = k & 255;
	const float n0 = falloff3(x0, y0, z0, P(ii + P(jj + P(kk))) % 12);
	const float n1 = falloff3(x1, y1, z1, P(ii + a1 + P(jj + b1 + P(kk + c1))) % 12);
	const float n2 = falloff3(x2, y2, z2, P(ii + a2 + P(jj + b2 + P(kk + c2))) % 12);
	const float n3 = falloff3(x3, y3, z3, P(ii + 1 + P(jj + 1 + P(kk + 1))) % 12);
	return 32.0F * (n0 + n1 + n2 + n3);
}

// get_terrain_heigth (test_och_h_octree.cpp:561-566) with noise = simplex_n(0.5F) (:35)
ORT_HD uint16_t terrain_height(int x, int y, int dim)
{
	const float px = static_cast<float>(x * 4) / static_cast<float>(dim);
	const float py = static_cast<float>(y * 4) / static_cast<float>(dim);
	return static_cast<uint16_t>(static_cast<int>(simplex2(0.5F, px, py) * static_cast<float>(dim) / 16 + static_cast<float>(dim / 4)));
}

// splatter_noise(-0.5F, .., 1/16) on the global simplex_n(0.5F) (test_och_h_octree.cpp:755-763, :770): true = carved
ORT_HD bool carve_test(int x, int y, int z)
{
	return !(simplex3(0.5F, static_cast<float>(x) * (1.0F / 16.0F), static_cast<float>(y) * (1.0F / 16.0F), static_cast<float>(z) * (1.0F / 16.0F)) >= -0.5F);
}

}  // namespace ort_noise
