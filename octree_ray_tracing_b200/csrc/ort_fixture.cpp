// ort_fixture.cpp -- headless-harness fixtures: the synthetic terrain the reference's demo builds in
// initialize_h_octree (test_och_h_octree.cpp:767-787), produced by a builder that scales to depth
// 12-14.  Host C++ only.
//
// The reference builds the DAG top-down over the whole solid volume (create_volume, :651-695), then
// patches 3*dim^2 voxels through set() (:776-783) and optionally runs dim^3 noise evaluations to
// carve tunnels (:735-743, :786): minutes at depth 12, hours with tunnels.  Because the DAG is
// content-addressed, the SAME voxel content always yields the same canonical DAG, so this builder
// evaluates the voxel function directly:
//     voxel(x,y,z) = 0            if z > h(x,y)            or carved
//                    2 + grass    if z == h(x,y)
//                    4            if z == h-1 or z == h-2
//                    1            otherwise (stone)
// recursing only where a cell is not trivially empty (above the column maximum) or, without tunnels,
// trivially solid stone (below column minimum - 2; one memoised node per level).  Reference counts
// are then computed top-down as instance counts, which is what the reference's per-instance
// register_node calls add up to (saturating at 2^32-1 instead of wrapping for depth >= 13).
#include "ort_internal.h"

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace {

// ---- och::simplex_n (och_noise.h:18-367): Gustavson simplex noise in float, int truncation ----

const uint8_t kPerm[256] = {
	151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10, 23, 190, 6, 148,
	247, 120, 234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87, 174, 20, 125, 136, 171, 168, 68, 175,
	74, 165, 71, 134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230, 220, 105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54,
	65, 25, 63, 161, 1, 216, 80, 73, 209, 76, 132, 187, 208, 89, 18, 169, 200, 196, 135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64,
	52, 217, 226, 250, 124, 123, 5, 202, 38, 147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213,
	119, 248, 152, 2, 44, 154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232, 178, 185, 112, 104,
	218, 246, 97, 228, 251, 34, 242, 193, 238, 210, 144, 12, 191, 179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106, 157,
	184, 84, 204, 176, 115, 121, 50, 45, 127, 4, 150, 254, 138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180
};

struct G3 { float x, y, z; };
const G3 kGrad[12] = {
	{ 1, 1, 0 }, { -1, 1, 0 }, { 1, -1, 0 }, { -1, -1, 0 }, { 1, 0, 1 }, { -1, 0, 1 },
	{ 1, 0, -1 }, { -1, 0, -1 }, { 0, 1, 1 }, { 0, -1, 1 }, { 0, 1, -1 }, { 0, -1, -1 }
};

inline int P(int i) { return kPerm[i & 255]; }

inline float falloff2(float x, float y, int g)
{
	float t = 0.5F - x * x - y * y;
	if (t < 0) return 0.0F;
	t *= t;
	return t * t * (kGrad[g].x * x + kGrad[g].y * y);
}

float simplex2(float freq, float x, float y)                     // och_noise.h:73-179
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

inline float falloff3(float x, float y, float z, int g)
{
	float t = 0.6F - x * x - y * y - z * z;
	if (t < 0) return 0.0F;
	t *= t;
	return t * t * (kGrad[g].x * x + kGrad[g].y * y + kGrad[g].z * z);
}

float simplex3(float freq, float x, float y, float z)            // och_noise.h:181-366
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

template<class F>
void parallel_rows(int rows, int nthreads, F f)
{
	if (nthreads <= 1) { for (int r = 0; r < rows; ++r) f(r); return; }
	std::atomic<int> next{ 0 };
	std::vector<std::thread> pool;
	for (int w = 0; w < nthreads; ++w)
		pool.emplace_back([&] { for (int r; (r = next.fetch_add(1)) < rows;) f(r); });
	for (auto& t : pool) t.join();
}

// ---- builder -------------------------------------------------------------------------------

struct Builder
{
	ort_tree* tree;
	int depth, dim;
	const uint16_t* h;
	const uint8_t* grass;
	bool tunnels;
	std::vector<std::vector<uint16_t>> hmin, hmax;    // [k] = min/max over 2^k x 2^k column blocks
	uint32_t stone[17];                               // memoised all-stone subtree per cell size 2^k
	bool failed = false;
	std::vector<uint64_t> carved;                     // tunnels: one bit per voxel with z <= zmax, filled in parallel
	int zmax = -1;

	static inline bool carve_test(int x, int y, int z)   // splatter_noise(-0.5F, .., 1/16) on the global simplex_n(0.5F) (:755-763, :770)
	{
		return !(simplex3(0.5F, static_cast<float>(x) * (1.0F / 16.0F), static_cast<float>(y) * (1.0F / 16.0F), static_cast<float>(z) * (1.0F / 16.0F)) >= -0.5F);
	}

	// the dim^3 noise loop of remove() (:735-743) restricted to z <= max height (voxels above are empty anyway),
	// spread over the host cores
	void precompute_carved(int nthreads)
	{
		zmax = hmax[depth][0];
		const size_t words_per_slab = (static_cast<size_t>(dim) * dim + 63) / 64;
		carved.assign(words_per_slab * (zmax + 1), 0);
		parallel_rows(zmax + 1, nthreads, [&](int z) {
			uint64_t* slab = carved.data() + words_per_slab * z;
			for (int y = 0; y < dim; ++y)
				for (int x = 0; x < dim; ++x)
					if (z <= h[static_cast<size_t>(y) * dim + x] && carve_test(x, y, z))
					{
						const size_t b = static_cast<size_t>(y) * dim + x;
						slab[b >> 6] |= 1ull << (b & 63);
					}
		});
	}

	inline bool is_carved(int x, int y, int z) const
	{
		const size_t words_per_slab = (static_cast<size_t>(dim) * dim + 63) / 64;
		const size_t b = static_cast<size_t>(y) * dim + x;
		return (carved[words_per_slab * z + (b >> 6)] >> (b & 63)) & 1u;
	}

	inline uint32_t voxel(int x, int y, int z) const
	{
		const int hh = h[static_cast<size_t>(y) * dim + x];
		uint32_t v;
		if (z > hh) return 0;
		if (z == hh) v = 2u + (grass[static_cast<size_t>(y) * dim + x] ? 1u : 0u);
		else if (z >= hh - 2) v = 4u;
		else v = 1u;
		if (tunnels && is_carved(x, y, z)) return 0;
		return v;
	}

	uint32_t intern(const uint32_t* n)
	{
		const uint32_t id = tree->intern_node(n);
		if (!id) failed = true;
		return id;
	}

	uint32_t stone_node(int k)                         // k = log2(cell size) >= 1
	{
		if (stone[k]) return stone[k];
		uint32_t n[8];
		const uint32_t c = k == 1 ? 1u : stone_node(k - 1);
		for (int i = 0; i < 8; ++i) n[i] = c;
		return stone[k] = intern(n);
	}

	uint32_t build(int x, int y, int z, int k)         // cell [x,x+2^k) x [y,..) x [z,..); returns node id or 0
	{
		if (failed) return 0;
		const int s = 1 << k;
		const size_t bi = static_cast<size_t>(y >> k) * (dim >> k) + (x >> k);
		if (z > hmax[k][bi])
			return 0;                                  // entirely above the terrain
		if (!tunnels && z + s - 1 < static_cast<int>(hmin[k][bi]) - 2)
			return stone_node(k);                      // entirely plain stone

		uint32_t n[8];
		if (k == 1)
			for (int c = 0; c < 8; ++c) n[c] = voxel(x + (c & 1), y + ((c >> 1) & 1), z + (c >> 2));
		else
		{
			const int hs = s >> 1;
			for (int c = 0; c < 8; ++c) n[c] = build(x + (c & 1 ? hs : 0), y + (c & 2 ? hs : 0), z + (c & 4 ? hs : 0), k - 1);
		}
		if (!(n[0] | n[1] | n[2] | n[3] | n[4] | n[5] | n[6] | n[7]))
			return 0;                                  // (only reachable with tunnels: a fully carved cell)
		return intern(n);
	}
};

}  // namespace

// refcount(node) = number of instances of the node in the fully expanded tree = what one
// register_node call per instance (the reference's builders) would have produced.
static void assign_instance_counts(ort_tree* t)
{
	if (!t->root) return;
	std::vector<uint64_t> acc(t->cap, 0);
	std::vector<uint32_t> cur{ t->root - 1 }, next;
	acc[t->root - 1] = 1;
	uint64_t total = 0;
	for (int level = 1; level <= t->depth; ++level)
	{
		next.clear();
		// take this level's counts first: a slot may (rarely) serve on two levels
		std::vector<uint64_t> cnt(cur.size());
		for (size_t i = 0; i < cur.size(); ++i) { cnt[i] = acc[cur[i]]; acc[cur[i]] = 0; }
		for (size_t i = 0; i < cur.size(); ++i)
		{
			const uint32_t slot = cur[i];
			const uint64_t sum = static_cast<uint64_t>(t->refcounts[slot]) + cnt[i];
			t->refcounts[slot] = sum > UINT32_MAX ? UINT32_MAX : static_cast<uint32_t>(sum);
			total += cnt[i];
			if (level == t->depth) continue;
			const uint32_t* n = t->nodes + 8 * static_cast<size_t>(slot);
			for (int c = 0; c < 8; ++c)
				if (n[c])
				{
					if (!acc[n[c] - 1]) next.push_back(n[c] - 1);
					acc[n[c] - 1] += cnt[i];
				}
		}
		cur.swap(next);
	}
	t->nodecnt += static_cast<uint32_t>(total);   // modulo 2^32 like the reference's counter
}

extern "C" {

int ort_parse_voxels(const char* text, size_t len, uint32_t* rgba6, char* names16, int max_voxels)
{
	// och_voxel.cpp:195-305: skip white space, read the name up to ':', then six colours of three hex byte pairs
	if (!text || !rgba6 || max_voxels <= 0)
		return ort_fail(nullptr, ORT_ERR_INVALID, "ort_parse_voxels: bad arguments"), -1;
	size_t i = 0;
	int n = 0;
	auto skip_space = [&] { while (i < len && std::isspace(static_cast<unsigned char>(text[i]))) ++i; };
	auto hexval = [](int c) { return c <= '9' ? c - '0' : (c | 0x20) - 'a' + 10; };
	for (;;)
	{
		skip_space();
		if (i >= len) break;
		if (n == max_voxels) break;
		char name[16] = { 0 };
		int k = 0;
		while (i < len && text[i] != ':')
		{
			if (k == 15)
				return ort_fail(nullptr, ORT_ERR_INVALID, "ort_parse_voxels: voxel-names may not exceed 15 characters (voxel %d)", n + 1), -1;
			name[k++] = text[i++];
		}
		if (i >= len)
			return ort_fail(nullptr, ORT_ERR_INVALID, "ort_parse_voxels: input ended unexpectedly in the name of voxel %d", n + 1), -1;
		if (k == 0)
			return ort_fail(nullptr, ORT_ERR_INVALID, "ort_parse_voxels: voxel-names must contain at least one character"), -1;
		++i;   // ':'
		for (int dir = 0; dir < 6; ++dir)
		{
			skip_space();
			uint32_t px = 0xFF000000u;
			for (int byte = 0; byte < 3; ++byte)
			{
				if (i + 1 >= len + 0 && i >= len)
					return ort_fail(nullptr, ORT_ERR_INVALID, "ort_parse_voxels: input ended unexpectedly (%s, colour %d)", name, dir + 1), -1;
				if (i + 1 >= len || !std::isxdigit(static_cast<unsigned char>(text[i])) || !std::isxdigit(static_cast<unsigned char>(text[i + 1])))
					return ort_fail(nullptr, ORT_ERR_INVALID, "ort_parse_voxels: non-hex character in colour-value (%s at colour no. %d)", name, dir + 1), -1;
				px |= static_cast<uint32_t>((hexval(text[i]) << 4) | hexval(text[i + 1])) << (8 * byte);
				i += 2;
			}
			rgba6[6 * n + dir] = px;
		}
		if (names16) std::memcpy(names16 + 16 * n, name, 16);
		++n;
	}
	return n;
}

void ort_fixture_heightmap(int depth, uint16_t* heights, int nthreads)
{
	const int dim = 1 << depth;
	parallel_rows(dim, nthreads, [&](int y) {
		for (int x = 0; x < dim; ++x)
		{
			// get_terrain_heigth (test_och_h_octree.cpp:561-566) with noise = simplex_n(0.5F) (:35)
			const float px = static_cast<float>(x * 4) / static_cast<float>(dim);
			const float py = static_cast<float>(y * 4) / static_cast<float>(dim);
			heights[static_cast<size_t>(y) * dim + x] = static_cast<uint16_t>(static_cast<int>(simplex2(0.5F, px, py) * static_cast<float>(dim) / 16 + static_cast<float>(dim / 4)));
		}
	});
}

int ort_fixture_build_terrain(ort_tree* tree, const uint16_t* heights, const uint8_t* grass, int tunnels, int nthreads)
{
	if (!tree || !heights || !grass)
		return ort_fail(nullptr, ORT_ERR_INVALID, "ort_fixture_build_terrain: bad arguments");

	Builder b;
	b.tree = tree;
	b.depth = tree->depth;
	b.dim = 1 << tree->depth;
	b.h = heights;
	b.grass = grass;
	b.tunnels = tunnels != 0;
	std::memset(b.stone, 0, sizeof b.stone);

	// min / max pyramids over column blocks
	b.hmin.resize(b.depth + 1);
	b.hmax.resize(b.depth + 1);
	b.hmin[0].assign(heights, heights + static_cast<size_t>(b.dim) * b.dim);
	b.hmax[0] = b.hmin[0];
	for (int k = 1; k <= b.depth; ++k)
	{
		const int w = b.dim >> k, pw = w * 2;
		b.hmin[k].resize(static_cast<size_t>(w) * w);
		b.hmax[k].resize(static_cast<size_t>(w) * w);
		for (int y = 0; y < w; ++y)
			for (int x = 0; x < w; ++x)
			{
				const size_t p = static_cast<size_t>(2 * y) * pw + 2 * x;
				const auto& lo = b.hmin[k - 1];
				const auto& hi = b.hmax[k - 1];
				b.hmin[k][static_cast<size_t>(y) * w + x] = std::min(std::min(lo[p], lo[p + 1]), std::min(lo[p + pw], lo[p + pw + 1]));
				b.hmax[k][static_cast<size_t>(y) * w + x] = std::max(std::max(hi[p], hi[p + 1]), std::max(hi[p + pw], hi[p + pw + 1]));
			}
	}

	if (b.tunnels)
		b.precompute_carved(nthreads > 0 ? nthreads : 1);
	tree->root = b.build(0, 0, 0, b.depth);
	if (b.failed || tree->table_full)
		return ort_fail(nullptr, ORT_ERR_TABLE_FULL, "ort_fixture_build_terrain: node table too full (raise log2_table_capacity)");
	assign_instance_counts(tree);
	tree->invalidate_mirror();
	return ORT_OK;
}

}  // extern "C"
