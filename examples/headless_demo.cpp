// headless_demo.cpp -- the reference's demo flow (test_och_h_octree.cpp:791-851) without the window:
// build the terrain, trace frames, apply the T/Z edits at the crosshair, report ms per frame like the HUD (:289).
//
//   g++ -std=c++17 -O2 -Iinclude examples/headless_demo.cpp -Loctree_ray_tracing_b200 -lort_b200
//       -Wl,-rpath,$PWD/octree_ray_tracing_b200 -o /tmp/headless_demo && /tmp/headless_demo [frames]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "och_h_octree_b200.hpp"

typedef och::h_octree<19, 8> tree_t;   // the reference's default (test_och_h_octree.cpp:24)

int main(int argc, char** argv)
{
	const int frames = argc > 1 ? std::atoi(argv[1]) : 20;
	const int W = 1280, H = 720;
	try
	{
		tree_t tree;
		const int dim = tree.dim;
		std::vector<uint16_t> heights(static_cast<size_t>(dim) * dim);
		std::vector<uint8_t> grass(static_cast<size_t>(dim) * dim);
		ort_fixture_heightmap(tree.depth, heights.data(), 4);
		uint64_t s = 1;
		for (auto& g : grass) { s = s * 6364136223846793005ull + 1442695040888963407ull; g = (s >> 63) & 1; }
		if (ort_fixture_build_terrain(tree.handle(), heights.data(), grass.data(), 1, 4) != ORT_OK) return 1;
		std::printf("Tree-depth: %d  Tree-dimension: %d  tabled nodes: %u  active nodes: %u\n", tree.depth, dim, tree.get_fillcnt(), tree.get_nodecnt());

		std::vector<uint32_t> vox(static_cast<size_t>(W) * H);
		std::vector<uint8_t> face(vox.size());
		std::vector<float> t(vox.size());
		och::float3 pos{ 1.5F, 1.5F, 1.0F + (heights[static_cast<size_t>(dim / 2) * dim + dim / 2] + 12) / static_cast<float>(dim) };
		float yaw = 0.3F, pitch = -0.9F;

		for (int f = 0; f < frames; ++f)
		{
			och::float3 dir3{ cosf(yaw) * cosf(pitch), sinf(yaw) * cosf(pitch), sinf(pitch) };   // :527
			och::direction hd; uint32_t hv; float ht;
			tree.sse_trace(pos, dir3, hd, hv, ht);                                                // :535-536
			if (hv && ht < 0.5F && f % 4 == 1)                                                    // T / Z held (:397-433)
			{
				const bool place = (f / 4) % 2 == 0;
				float off[3] = { 0, 0, 0 };
				if (static_cast<int>(hd) < 6) off[static_cast<int>(hd) % 3] = (tree.voxel_dim / 2) * (static_cast<int>(hd) < 3 ? 1 : -1);
				const float sgn = place ? 1.0F : -1.0F;
				const uint16_t cx = static_cast<uint16_t>((pos.x + dir3.x * ht + sgn * off[0] - 1.0F) * dim);
				const uint16_t cy = static_cast<uint16_t>((pos.y + dir3.y * ht + sgn * off[1] - 1.0F) * dim);
				const uint16_t cz = static_cast<uint16_t>((pos.z + dir3.z * ht + sgn * off[2] - 1.0F) * dim);
				ort_tree_set_box(tree.handle(), cx, cy, cz, 40, place ? 1 : 0);
			}
			auto beg = std::chrono::steady_clock::now();
			tree.trace_frame(pos, yaw, pitch, W, H, vox.data(), face.data(), t.data());
			auto end = std::chrono::steady_clock::now();
			size_t hits = 0;
			for (uint32_t v : vox) hits += v != 0;
			std::printf("frame %2d: %7.3f ms  hits %zu  tabled nodes %u  looking at voxel %u (dir %d, t %.5f)\n", f,
			            std::chrono::duration<double, std::milli>(end - beg).count(), hits, tree.get_fillcnt(), hv, static_cast<int>(hd), ht);
			yaw += 0.01F;
		}
	}
	catch (const std::exception& e)
	{
		std::fprintf(stderr, "error: %s\n", e.what());
		return 2;
	}
	return 0;
}
