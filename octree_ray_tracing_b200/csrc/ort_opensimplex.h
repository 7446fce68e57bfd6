// ort_opensimplex.h -- 2-D OpenSimplex noise as the reference vendors it (opensimplex.h: OpenSimplexNoise(seed) :222-244,
// the 2-D contribution tables :246-289, Evaluate(x, y) :338-386), restated for the fixture builder: the demo's
// alternative heightmap noise `terrain_noise(8789)` (test_och_h_octree.cpp:33, the commented line :568).  Host only,
// double precision, every operation in the reference's order (the translation units are compiled without contraction),
// so the heights are the reference's bit for bit -- the CPU tests hold it to the real class where the compiled reference
// is present, tests/golden/opensimplex_8789.npz pins it elsewhere.
#pragma once

#include <cstdint>

namespace ort_noise {

class OpenSimplex2
{
public:
	explicit OpenSimplex2(int64_t seed)
	{
		// the permutation: a 64-bit LCG drives a Fisher-Yates shuffle of 0..255 (opensimplex.h:224-243)
		unsigned char source[256];
		for (int i = 0; i < 256; ++i) source[i] = static_cast<unsigned char>(i);
		uint64_t s = static_cast<uint64_t>(seed);
		auto next = [&s]() { s = s * 6364136223846793005ull + 1442695040888963407ull; };
		next(); next(); next();
		for (int i = 255; i >= 0; --i)
		{
			next();
			int r = static_cast<int>((static_cast<int64_t>(s) + 31) % (i + 1));
			if (r < 0) r += i + 1;
			perm_[i] = source[r];
			perm2d_[i] = perm_[i] & 0x0E;
			source[r] = source[i];
		}

		// the six contribution lists (base set + one extra vertex) and the 64-entry hash -> list table (:246-289)
		static const int base[2][3][3] = { { { 1, 1, 0 }, { 1, 0, 1 }, { 0, 0, 0 } }, { { 1, 1, 0 }, { 1, 0, 1 }, { 2, 1, 1 } } };
		static const int extra[6][4] = { { 0, 0, 1, -1 }, { 0, 0, -1, 1 }, { 0, 2, 1, 1 }, { 1, 2, 2, 0 }, { 1, 2, 0, 2 }, { 1, 0, 0, 0 } };
		static const int pairs[12][2] = { { 0, 1 }, { 1, 0 }, { 4, 1 }, { 17, 0 }, { 20, 2 }, { 21, 2 }, { 22, 5 }, { 23, 5 }, { 26, 4 }, { 39, 3 }, { 42, 4 }, { 43, 3 } };
		for (int k = 0; k < 6; ++k)
		{
			for (int j = 0; j < 3; ++j) set_[k][j] = make(base[extra[k][0]][j][0], base[extra[k][0]][j][1], base[extra[k][0]][j][2]);
			set_[k][3] = make(extra[k][1], extra[k][2], extra[k][3]);
		}
		for (int h = 0; h < 64; ++h) lookup_[h] = -1;
		for (const auto& p : pairs) lookup_[p[0]] = p[1];
	}

	// OpenSimplexNoise::Evaluate(x, y) (:338-386)
	double operator()(double x, double y) const
	{
		const double stretch = (x + y) * kStretch;
		const double xs = x + stretch, ys = y + stretch;
		const int xsb = fast_floor(xs), ysb = fast_floor(ys);
		const double squish = (xsb + ysb) * kSquish;
		const double dx0 = x - (xsb + squish), dy0 = y - (ysb + squish);
		const double xins = xs - xsb, yins = ys - ysb;
		const double in_sum = xins + yins;
		const int hash = static_cast<int>(xins - yins + 1) | static_cast<int>(in_sum) << 1 | static_cast<int>(in_sum + yins) << 2 | static_cast<int>(in_sum + xins) << 4;
		double value = 0.0;
		const int k = lookup_[hash & 63];
		if (k >= 0)
			for (int j = 0; j < 4; ++j)
			{
				const Contribution& c = set_[k][j];
				const double dx = dx0 + c.dx, dy = dy0 + c.dy;
				double attn = 2 - dx * dx - dy * dy;
				if (attn > 0)
				{
					const int px = xsb + c.xsb, py = ysb + c.ysb;
					const int i = perm2d_[(perm_[px & 0xFF] + py) & 0xFF];
					const double part = kGrad[i] * dx + kGrad[i + 1] * dy;
					attn *= attn;
					value += attn * attn * part;
				}
			}
		return value * (1.0 / 47.0);
	}

private:
	struct Contribution { double dx, dy; int xsb, ysb; };
	static constexpr double kStretch = -0.211324865405187, kSquish = 0.366025403784439;
	static constexpr double kGrad[16] = { 5, 2, 2, 5, -5, 2, -2, 5, 5, -2, 2, -5, -5, -2, -2, -5 };
	static Contribution make(int multiplier, int xsb, int ysb) { return Contribution{ -xsb - multiplier * kSquish, -ysb - multiplier * kSquish, xsb, ysb }; }
	static int fast_floor(double x) { const int xi = static_cast<int>(x); return x < xi ? xi - 1 : xi; }

	unsigned char perm_[256], perm2d_[256];
	Contribution set_[6][4];
	int lookup_[64];
};

// the commented heightmap line of get_terrain_heigth (test_och_h_octree.cpp:568): a double expression, truncated
inline uint16_t terrain_height_opensimplex(const OpenSimplex2& noise, int x, int y, int dim)
{
	const float px = static_cast<float>(x * 4) / static_cast<float>(dim);
	const float py = static_cast<float>(y * 4) / static_cast<float>(dim);
	return static_cast<uint16_t>(static_cast<int>(noise(px, py) * dim / 16 + dim / 4));
}

}  // namespace ort_noise
