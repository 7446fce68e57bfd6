// octree_demo.cpp -- och::octree (the plain pointer octree, och_octree.h) through the drop-in header: fill a few
// boxes, trace one frame, unset a box, trace again.
#include <cstdio>
#include <vector>

#include "och_octree_b200.hpp"

int main()
{
	try
	{
		och::octree tree(8, 1u << 18);
		for (int b = 0; b < 6; ++b)
			for (int z = 0; z < 12; ++z)
				for (int y = 0; y < 12; ++y)
					for (int x = 0; x < 12; ++x)
						tree.set(static_cast<int16_t>(20 + 35 * b + x), static_cast<int16_t>(100 + y), static_cast<int16_t>(40 + z), 1 + b);
		const int W = 640, H = 360;
		std::vector<uint32_t> vox(W * H);
		std::vector<uint8_t> face(W * H);
		std::vector<float> t(W * H);
		for (int pass = 0; pass < 2; ++pass)
		{
			tree.trace_frame({ 1.5F, 1.1F, 1.6F }, 1.4F, -0.9F, W, H, vox.data(), face.data(), t.data());
			size_t hits = 0;
			for (uint32_t v : vox) hits += v != 0;
			std::printf("pass %d: nodes %d  hits %zu\n", pass, tree.get_node_cnt(), hits);
			for (int z = 0; z < 12; ++z)
				for (int y = 0; y < 12; ++y)
					for (int x = 0; x < 12; ++x)
						tree.unset(static_cast<int16_t>(20 + 35 * 2 + x), static_cast<int16_t>(100 + y), static_cast<int16_t>(40 + z));
		}
		och::direction d; uint32_t v; float tt;
		tree.sse_trace(1.5F, 1.5F, 1.9F, 0.0F, 0.0F, 1.0F, d, v, tt);
		std::printf("miss: dir %d voxel %u t %g\n", static_cast<int>(d), v, tt);
	}
	catch (const std::exception& e)
	{
		std::fprintf(stderr, "error: %s\n", e.what());
		return 2;
	}
	return 0;
}
