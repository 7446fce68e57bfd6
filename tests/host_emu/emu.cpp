// emu.cpp -- TEST INFRASTRUCTURE: csrc/ort_trace.cuh compiled for the host (see cuda_shim.h).  Exports the per-ray walks
// of the CUDA kernels as plain C functions over host arrays so that tests/test_host_emu.py can compare them bit for bit
// with the oracle without a GPU, and count what they do (rounds per level).  Built on demand by
// the test into tests/host_emu/_build/; never part of libort_b200.so.
#include "cuda_shim.h"
#include "../../octree_ray_tracing_b200/csrc/ort_trace.cuh"
#include "../../octree_ray_tracing_b200/csrc/ort_trace_experiments.cuh"
#include "../../octree_ray_tracing_b200/csrc/ort_beam.cuh"

#include <vector>

#include <cstddef>
#include <omp.h>

namespace {

struct Stats
{
	unsigned long long rounds_by_level[ort::kMaxDepth + 2];   // child-slot loads issued with the walker at that level
	unsigned long long rays, slow_path_rays;                  // slow path: rays outside FastWalker's preconditions
	unsigned long long oob_loads;                             // loads outside the node array / reciprocal table (cuda_shim.h)
	unsigned long long lean_rays;                             // walker 13: rays that took LeanWalker (the rest: FastWalker / traverse)
	unsigned long long beam_rays, beam_misses, beam_guard;    // beam start: rays re-entered at tau, rays ended as a MISS without a round, guard re-walks
	unsigned long long beam_tile_misses, beam_cert_wrong;     // rays of tiles ended as a whole; of those, rays that were NOT lean-tier rays (must be 0)
};

// walker ids follow ort_set_option("variant"): 0 baseline traverse(), 1 FastWalker, 5 TightWalker, 7 PipeWalker,
// 13 the round-2 tiers (LeanWalker first), 14 FlatWalker, 15 V4Walker (both on LeanWalker's tiers)
template<bool COUNT>
ort::Hit walk(int walker, const uint32_t* nodes_m1, uint32_t root, int depth, float miss_t, float ox, float oy, float oz, const ort::Ray& ray,
              Stats* st, int origin_flags = -1, float tau = 0.0f)
{
	uint32_t stack[ort::kMaxDepth];
	if (walker >= 13 && walker <= 15)
	{
		// the kernels' tier test (ort::trace_ray): origin facts known for the launch (camera frames) or tested per ray;
		// LeanWalker (or an experiment round on its state) where no t can be negative, FastWalker / traverse() otherwise
		const bool origin_ok = origin_flags >= 0 ? (origin_flags & static_cast<int>(ort::kOriginInCube)) != 0 : ort::origin_in_cube(ox, oy, oz, ray);
		if (!(origin_ok && ort::lean_path_ok(ray)))
			return walk<COUNT>(1, nodes_m1, root, depth, miss_t, ox, oy, oz, ray, st);
		if (st) ++st->lean_rays;
		uint32_t lstack[ort::kMaxDepth] = {};
		const ort::LeanStack<0> ls{ lstack };
		const unsigned long long base_biased = reinterpret_cast<unsigned long long>(nodes_m1) - 4ull * ort::kMagicBits;
		const float leaf_dimf = std::ldexp(1.0f, -depth);
		if (walker == 14)
		{
			ort::FlatWalker<COUNT> w;
			w.start(root, ray);
			while (!w.round(base_biased, leaf_dimf, miss_t, ls)) {}
			return w.hit;
		}
		if (walker == 15)
		{
			ort::V4Walker<COUNT> w;
			w.start(root, ray);
			while (!w.round(base_biased, leaf_dimf, miss_t, ls)) {}
			return w.hit;
		}
		// the product's loop shape (ort::walk_ray): beam start or ordinary start, descend while there are children, then
		// one advance; the beam guard re-walks from the start
		ort::LeanWalker<COUNT> w;
		bool beam_used;
		if (ort::lean_start(w, root, ray, tau, miss_t, beam_used))
		{
			if (st) ++st->beam_misses;
			return w.hit;
		}
		if (st && beam_used) ++st->beam_rays;
		for (;;)
		{
			for (;;)
			{
				uint32_t child;
				bool done = false;
				while ((child = w.load_child(base_biased)) != 0u)
					if (w.descend(child, leaf_dimf, ls)) { done = true; break; }
				if (done || w.advance(miss_t, ls)) break;
			}
			if (!(beam_used && w.mti == 8u)) break;
			if (st) ++st->beam_guard;
			beam_used = false;
			w.start(root, ray);
		}
		return w.hit;
	}
	if (walker == 0 || !ort::fast_path_ok(ox, oy, oz, ray))
	{
		if (st) ++st->slow_path_rays;
		ort::Hit h = ort::traverse(nodes_m1, root, depth, miss_t, ray, stack);
		if (!COUNT) h.npush = 0;
		return h;
	}
	if (walker == 1)
	{
		ort::FastWalker<COUNT> w;
		w.start(root, miss_t, ray);
		for (;;)
		{
			if (st) ++st->rounds_by_level[w.level];
			const uint32_t child = w.load_child(nodes_m1);
			if (child ? w.descend(child, depth, stack) : w.advance(stack))
				break;
		}
		return w.hit;
	}
	if (walker == 5)
	{
		ort::TightWalker<COUNT> w;
		w.start(root, miss_t, ray);
		for (;;)
		{
			const uint32_t child = w.load_child(nodes_m1);
			if (child ? w.descend(child, depth, stack) : w.advance(stack))
				break;
		}
		return w.hit;
	}
	if (walker == 7)
	{
		uint32_t stack_n[ort::kMaxDepth];
		float stack_f[ort::kMaxDepth];
		ort::PipeWalker<COUNT> w;
		w.start(root, miss_t, ray);
		const unsigned long long base_biased = reinterpret_cast<unsigned long long>(nodes_m1) - 4ull * ort::kMagicBits;
		for (;;)
		{
			const uint32_t child = w.load_child(base_biased);
			if (child ? w.descend(child, depth, stack_n, stack_f) : w.advance(stack_n, stack_f))
				break;
		}
		return w.hit;
	}
	ort::Hit bad;
	bad.voxel = 0xFFFFFFFFu; bad.face = 0xFF; bad.t = -1.0f; bad.npush = 0;
	return bad;
}

void set_bounds(const uint32_t* nodes8, size_t n_rows, const uint32_t* rcp_tab, int log2n)
{
	g_emu_bounds.lo[0] = reinterpret_cast<const char*>(nodes8);
	g_emu_bounds.hi[0] = reinterpret_cast<const char*>(nodes8 + 8 * n_rows);
	g_emu_bounds.lo[1] = reinterpret_cast<const char*>(rcp_tab);
	g_emu_bounds.hi[1] = reinterpret_cast<const char*>(rcp_tab + (static_cast<size_t>(1) << log2n));
	g_emu_bounds.violations = 0;
}

void merge(Stats* dst, const Stats& s)
{
	for (int i = 0; i < ort::kMaxDepth + 2; ++i) dst->rounds_by_level[i] += s.rounds_by_level[i];
	dst->rays += s.rays; dst->slow_path_rays += s.slow_path_rays; dst->oob_loads += s.oob_loads; dst->lean_rays += s.lean_rays;
	dst->beam_rays += s.beam_rays; dst->beam_misses += s.beam_misses; dst->beam_guard += s.beam_guard;
	dst->beam_tile_misses += s.beam_tile_misses; dst->beam_cert_wrong += s.beam_cert_wrong;
}

}  // namespace

extern "C" {

// nodes8: the compact device array as ort_tree_flatten yields it (node id i at nodes8[8*(i-1)], index_base 1), or the
// och::octree pool (raw rows, root = row 0, index_base 0).  has_root = 0: empty tree, every ray is a MISS.
int emu_trace_rays(const uint32_t* nodes8, size_t n_rows, int index_base, int has_root, uint32_t root, int depth, float miss_t, const uint32_t* rcp_tab, int log2n,
                   const float* o3, int o_stride, const float* d3, size_t n, int walker,
                   uint32_t* voxel, uint8_t* face, float* t, uint16_t* npush, unsigned long long* stats_out, int nthreads, const float* tau)
{
	const uint32_t* nodes_m1 = nodes8 - 8 * static_cast<ptrdiff_t>(index_base);
	const ort::RcpTable rt{rcp_tab, 23 - log2n};
	Stats total{};
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
	{
		Stats st{};
		set_bounds(nodes8, n_rows, rcp_tab, log2n);
#pragma omp for schedule(dynamic, 4096)
		for (long long i = 0; i < static_cast<long long>(n); ++i)
		{
			const float* o = o3 + static_cast<size_t>(i) * o_stride;
			const float* d = d3 + static_cast<size_t>(i) * 3;
			const ort::Ray r = ort::ray_setup(rt, o[0], o[1], o[2], d[0], d[1], d[2], (1u << (23 - depth)) - 1u);
			ort::Hit h;
			if (!has_root) { h.voxel = 0; h.face = 6; h.t = miss_t; h.npush = 0; }
			else h = npush ? walk<true>(walker, nodes_m1, root, depth, miss_t, o[0], o[1], o[2], r, stats_out ? &st : nullptr, -1, tau ? tau[i] : 0.0f)
			               : walk<false>(walker, nodes_m1, root, depth, miss_t, o[0], o[1], o[2], r, stats_out ? &st : nullptr, -1, tau ? tau[i] : 0.0f);
			++st.rays;
			voxel[i] = h.voxel;
			face[i] = static_cast<uint8_t>(h.face);
			t[i] = h.t;
			if (npush) npush[i] = static_cast<uint16_t>(h.npush < 65535u ? h.npush : 65535u);
		}
		st.oob_loads = g_emu_bounds.violations;
#pragma omp critical
		merge(&total, st);
	}
	if (stats_out) std::memcpy(stats_out, &total, sizeof(total));
	return total.oob_loads ? 1 : 0;                              // 1: some load left the node array / the table
}

// camera rays generated like the frame kernels do (ort::camera_ray); rows as ort_trace_frame takes them (ort::frame_row)
int emu_trace_frame(const uint32_t* nodes8, size_t n_rows, int index_base, int has_root, uint32_t root, int depth, float miss_t, const uint32_t* rcp_tab, int log2n,
                    const float pos[3], const float rot[9], float fov, int W, int H, int y0, int rows, int tile_rows, int tile_step, int walker,
                     uint32_t* voxel, uint8_t* face, float* t, uint16_t* npush, unsigned long long* stats_out, int nthreads,
                     const uint8_t* beam_skip, int beam_k, float* tau_out)
{
	const uint32_t* nodes_m1 = nodes8 - 8 * static_cast<ptrdiff_t>(index_base);
	const ort::RcpTable rt{rcp_tab, 23 - log2n};
	ort::Camera cam;
	cam.ox = pos[0]; cam.oy = pos[1]; cam.oz = pos[2];
	for (int i = 0; i < 9; ++i) cam.r[i] = rot[i];
	cam.fov = fov;
	cam.aspect = static_cast<float>(W) / static_cast<float>(H);
	cam.vfx = 2.0F / static_cast<float>(W);
	cam.vfy = 2.0F / static_cast<float>(H);
	cam.origin_flags = ort::camera_origin_flags(cam.ox, cam.oy, cam.oz, (1u << (23 - depth)) - 1u);
	const int oflags = static_cast<int>(cam.origin_flags);
	const float beam_min_comp = beam_skip ? ort::beam_certify_min_comp(cam, ort::beam_tile_radius(cam, ort::rcp_table_rel_error(rcp_tab, log2n)), ort::rcp_table_sig_bits(rcp_tab, log2n)) : 0.0f;
	const ort::FrameRows fr{ W, H, y0, rows, tile_rows, tile_step, 0, 0, ort::tile_shift_of(tile_rows) };
	Stats total{};
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
	{
		Stats st{};
		set_bounds(nodes8, n_rows, rcp_tab, log2n);
#pragma omp for schedule(dynamic, 4)
		for (int r = 0; r < rows; ++r)
			for (int x = 0; x < W; ++x)
			{
				float dx, dy, dz;
				ort::camera_ray(cam, x, ort::frame_row(fr, r), dx, dy, dz);
				const ort::Ray ray = ort::ray_setup_camera(rt, cam.ox, cam.oy, cam.oz, dx, dy, dz, cam.origin_flags);
				// beam start of the pixel's 8 x 4 tile, as beam_start_kernel computes it (the kernels: once per tile)
				float tau = 0.0f;
				if (beam_skip)
					tau = ort::beam_tile_start(ort::BeamGrid{ beam_skip, beam_k }, cam, x & ~7, ort::frame_row(fr, r & ~3), beam_min_comp);
				if (tau_out) tau_out[static_cast<size_t>(r) * W + x] = tau;
				ort::Hit h;
				if (beam_skip && __float_as_uint(tau) == ort::kBeamAllMissBits)
				{
					// the kernels' whole-tile MISS: no ray is set up.  The claim behind it -- every ray of the tile is a lean-tier
					// ray -- is checked here, where the ray exists anyway.
					h.voxel = 0; h.face = 6; h.t = miss_t; h.npush = 0;
					++st.beam_tile_misses;
					if (!((oflags & static_cast<int>(ort::kOriginInCube)) && ort::lean_path_ok(ray) && fmaxf(ray.bx, fmaxf(ray.by, ray.bz)) < __uint_as_float(0x7F800000u)))
						++st.beam_cert_wrong;
				}
				else
				if (!has_root) { h.voxel = 0; h.face = 6; h.t = miss_t; h.npush = 0; }
				else h = npush ? walk<true>(walker, nodes_m1, root, depth, miss_t, cam.ox, cam.oy, cam.oz, ray, stats_out ? &st : nullptr, oflags, tau)
				               : walk<false>(walker, nodes_m1, root, depth, miss_t, cam.ox, cam.oy, cam.oz, ray, stats_out ? &st : nullptr, oflags, tau);
				++st.rays;
				const size_t i = static_cast<size_t>(r) * W + x;
				voxel[i] = h.voxel;
				face[i] = static_cast<uint8_t>(h.face);
				t[i] = h.t;
				if (npush) npush[i] = static_cast<uint16_t>(h.npush < 65535u ? h.npush : 65535u);
			}
		st.oob_loads = g_emu_bounds.violations;
#pragma omp critical
		merge(&total, st);
	}
	if (stats_out) std::memcpy(stats_out, &total, sizeof(total));
	return total.oob_loads ? 1 : 0;                              // 1: some load left the node array / the table
}

int emu_stats_words(void) { return static_cast<int>(sizeof(Stats) / sizeof(unsigned long long)); }

// The beam grid of level k for a DAG (ort_beam.cuh), computed the plain way on the host: occupancy by descending from the
// root per cell, 27-neighbour dilation, OR pyramid, skip level.  skip: (2^k)^3 bytes.  The GPU kernels must produce the
// same bytes (tests/test_gpu_beam.py).
void emu_beam_grid(const uint32_t* nodes8, int index_base, uint32_t root, int k, uint8_t* skip)
{
	const uint32_t* nodes_m1 = nodes8 - 8 * static_cast<ptrdiff_t>(index_base);
	const int N = 1 << k;
	const size_t cells = static_cast<size_t>(N) * N * N;
	std::vector<uint8_t> occ(cells), dil(cells);
#pragma omp parallel for schedule(static)
	for (long long i = 0; i < static_cast<long long>(cells); ++i)
	{
		const int x = static_cast<int>(i % N), y = static_cast<int>((i / N) % N), z = static_cast<int>(i / (static_cast<long long>(N) * N));
		uint32_t node = root, child = 1;
		for (int l = k - 1; l >= 0 && child; --l)
		{
			const uint32_t slot = ((x >> l) & 1) | (((y >> l) & 1) << 1) | (((z >> l) & 1) << 2);
			child = nodes_m1[(static_cast<size_t>(node) << 3) + slot];
			node = child;
		}
		occ[i] = child != 0;
	}
#pragma omp parallel for schedule(static)
	for (long long i = 0; i < static_cast<long long>(cells); ++i)
	{
		const int x = static_cast<int>(i % N), y = static_cast<int>((i / N) % N), z = static_cast<int>(i / (static_cast<long long>(N) * N));
		uint8_t any = 0;
		for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx)
		{
			const int xx = x + dx, yy = y + dy, zz = z + dz;
			if (xx >= 0 && xx < N && yy >= 0 && yy < N && zz >= 0 && zz < N) any |= occ[(static_cast<size_t>(zz) * N + yy) * N + xx];
		}
		dil[i] = any;
	}
	// pyr[j]: level j of the OR pyramid over dil
	std::vector<std::vector<uint8_t>> pyr(k + 1);
	pyr[k] = dil;
	for (int j = k - 1; j >= 1; --j)
	{
		const int M = 1 << j, F = 2 * M;
		pyr[j].assign(static_cast<size_t>(M) * M * M, 0);
		for (int z = 0; z < F; ++z) for (int y = 0; y < F; ++y) for (int x = 0; x < F; ++x)
			pyr[j][(static_cast<size_t>(z / 2) * M + y / 2) * M + x / 2] |= pyr[j + 1][(static_cast<size_t>(z) * F + y) * F + x];
	}
#pragma omp parallel for schedule(static)
	for (long long i = 0; i < static_cast<long long>(cells); ++i)
	{
		const int x = static_cast<int>(i % N), y = static_cast<int>((i / N) % N), z = static_cast<int>(i / (static_cast<long long>(N) * N));
		int s = 0;
		for (int j = 1; j <= k && !s; ++j)
		{
			const int sh = k - j, M = 1 << j;
			if (!pyr[j][(static_cast<size_t>(z >> sh) * M + (y >> sh)) * M + (x >> sh)]) s = j;
		}
		skip[i] = static_cast<uint8_t>(s);
	}
}

// the host-side choice of the grid level for a camera (ort::beam_tile_radius / beam_level_for), 0 = no beam start
// march statistics of a frame (a design aid): steps[tile] = grid lookups of the tile's march, tau[tile] = its start time
void emu_beam_march_stats(const uint8_t* skip, int k, const float pos[3], const float rot[9], float fov, int W, int H, int* steps, float* tau)
{
	ort::Camera cam{};
	cam.ox = pos[0]; cam.oy = pos[1]; cam.oz = pos[2];
	for (int i = 0; i < 9; ++i) cam.r[i] = rot[i];
	cam.fov = fov;
	cam.aspect = static_cast<float>(W) / static_cast<float>(H);
	cam.vfx = 2.0F / static_cast<float>(W);
	cam.vfy = 2.0F / static_cast<float>(H);
	const int tx = (W + 7) / 8, ty = (H + 3) / 4;
#pragma omp parallel for schedule(static)
	for (int i = 0; i < tx * ty; ++i)
	{
		int n = 0;
		tau[i] = ort::beam_tile_start(ort::BeamGrid{ skip, k }, cam, (i % tx) * 8, (i / tx) * 4, 0.0f, &n);
		steps[i] = n;
	}
}

int emu_beam_level(const float pos[3], const float rot[9], float fov, int W, int H, int depth, const uint32_t* rcp_tab, int log2n)
{
	ort::Camera cam{};
	cam.ox = pos[0]; cam.oy = pos[1]; cam.oz = pos[2];
	for (int i = 0; i < 9; ++i) cam.r[i] = rot[i];
	cam.fov = fov;
	cam.aspect = static_cast<float>(W) / static_cast<float>(H);
	cam.vfx = 2.0F / static_cast<float>(W);
	cam.vfy = 2.0F / static_cast<float>(H);
	const double eps = ort::rcp_table_rel_error(rcp_tab, log2n);
	const double radius = ort::beam_tile_radius(cam, eps);
	if (!(ort::camera_origin_flags(cam.ox, cam.oy, cam.oz, 0u) & ort::kOriginInCube)) return 0;
	return radius < 0 ? 0 : ort::beam_level_for(radius, depth, ort::beam_t_max(cam.ox, cam.oy, cam.oz, eps));
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// SIMT cost model of the frame kernel (a design aid, not a test): the rays of an 8 x 4 pixel tile advance in lockstep,
// one FastWalker round per step, like the lanes of a warp.  Per warp-round the model records which of the two
// branches of the round (descend: non-empty child / advance: empty child) have takers -- a warp pays for every
// branch that has at least one lane in it.
// out[0] warp-rounds, [1] with descend lanes only, [2] with advance lanes only, [3] with both,
// [4] sum of active lanes over warp-rounds, [5] sum of descend lanes, [6] sum of advance lanes, [7] warps,
// [8] sum over warps of the longest lane's rounds (= warp-rounds), [9] lane-rounds (sum of all lanes' rounds)
// ------------------------------------------------------------------------------------------------
extern "C" int emu_warp_model(const uint32_t* nodes8, size_t n_rows, uint32_t root, int depth, const uint32_t* rcp_tab, int log2n,
                              const float pos[3], const float rot[9], float fov, int W, int H, unsigned long long* out, int nthreads)
{
	const uint32_t* nodes_m1 = nodes8 - 8;
	const ort::RcpTable rt{rcp_tab, 23 - log2n};
	ort::Camera cam;
	cam.ox = pos[0]; cam.oy = pos[1]; cam.oz = pos[2];
	for (int i = 0; i < 9; ++i) cam.r[i] = rot[i];
	cam.fov = fov;
	cam.aspect = static_cast<float>(W) / static_cast<float>(H);
	cam.vfx = 2.0F / static_cast<float>(W);
	cam.vfy = 2.0F / static_cast<float>(H);
	const int tx = (W + 7) / 8, ty = (H + 3) / 4;
	unsigned long long tot[10] = {};
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
	{
		unsigned long long loc[10] = {};
		set_bounds(nodes8, n_rows, rcp_tab, log2n);
#pragma omp for schedule(dynamic, 64)
		for (int tile = 0; tile < tx * ty; ++tile)
		{
			ort::FastWalker<false> w[32];
			uint32_t stack[32][ort::kMaxDepth];
			bool active[32];
			int n_active = 0;
			for (int l = 0; l < 32; ++l)
			{
				const int x = (tile % tx) * 8 + (l & 7), y = (tile / tx) * 4 + (l >> 3);
				active[l] = false;
				if (x >= W || y >= H) continue;
				float dx, dy, dz;
				ort::camera_ray(cam, x, y, dx, dy, dz);
				const ort::Ray ray = ort::ray_setup(rt, cam.ox, cam.oy, cam.oz, dx, dy, dz);
				if (!ort::fast_path_ok(cam.ox, cam.oy, cam.oz, ray)) continue;      // slow-path rays are outside the model
				w[l].start(root, __uint_as_float(0x7F800000u), ray);
				active[l] = true;
				++n_active;
			}
			if (!n_active) continue;
			++loc[7];
			while (n_active)
			{
				int nd = 0, na = 0;
				for (int l = 0; l < 32; ++l)
				{
					if (!active[l]) continue;
					const uint32_t child = w[l].load_child(nodes_m1);
					bool done;
					if (child) { ++nd; done = w[l].descend(child, depth, stack[l]); }
					else       { ++na; done = w[l].advance(stack[l]); }
					if (done) active[l] = false;
				}
				++loc[0];
				loc[nd && !na ? 1 : (!nd && na ? 2 : 3)] += 1;
				loc[4] += nd + na; loc[5] += nd; loc[6] += na;
				loc[9] += nd + na;
				n_active = 0;
				for (int l = 0; l < 32; ++l) n_active += active[l];
			}
		}
		loc[8] = loc[0];
#pragma omp critical
		for (int i = 0; i < 10; ++i) tot[i] += loc[i];
	}
	std::memcpy(out, tot, sizeof(tot));
	return 0;
}
