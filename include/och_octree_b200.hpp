// och_octree_b200.hpp -- och::octree (och_octree.h:10-69) rebuilt over libort_b200.so: same constructor and
// members (set, unset, at, get_node_cnt, both sse_trace overloads; depth, dim, table_capacity), GPU trace.
// Like the reference, a MISS reports hit_time = 0.0F (och_octree.cpp:302) and set() does no range check.
// Pool exhaustion throws instead of printf + exit(0) (och_octree.cpp:50-54).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>

#include "och_h_octree_b200.hpp"   // och::direction, och::float3, ort_b200.h

namespace och
{
	class octree
	{
	public:
		const uint16_t depth;
		const uint16_t dim;
		const uint32_t table_capacity;

		octree(uint16_t depth, uint32_t table_capacity, int device = 0) : depth(depth), dim(static_cast<uint16_t>(1 << depth)), table_capacity(table_capacity)
		{
			check(ort_octree_create(&tree, depth, table_capacity), "ort_octree_create");
			if (ort_create(&ctx, device, depth, table_capacity < (1u << 16) ? table_capacity : (1u << 16)) != ORT_OK)
			{
				std::string msg = std::string("ort_create: ") + ort_last_error(nullptr);
				ort_octree_destroy(tree);
				throw std::runtime_error(msg);
			}
			ort_octree_attach(tree, ctx);
		}

		~octree()
		{
			ort_octree_destroy(tree);
			ort_destroy(ctx);
		}

		octree(const octree&) = delete;
		octree& operator=(const octree&) = delete;

		void set(int16_t x, int16_t y, int16_t z, uint32_t vx)
		{
			ort_octree_set(tree, x, y, z, vx);
			if (ort_octree_failed(tree)) throw std::runtime_error("octree: Too many allocations");
		}

		void unset(int16_t x, int16_t y, int16_t z) { ort_octree_unset(tree, x, y, z); }
		uint32_t at(int16_t x, int16_t y, int16_t z) const { return ort_octree_at(tree, x, y, z); }
		int get_node_cnt() const { return ort_octree_get_node_cnt(tree); }

		void sse_trace(float ox, float oy, float oz, float dx, float dy, float dz, direction& hit_direction, uint32_t& hit_voxel, float& hit_time) const
		{
			const float o[3] = { ox, oy, oz }, d[3] = { dx, dy, dz };
			uint8_t face = 8;
			check(ort_octree_sync(tree), "ort_octree_sync");
			check(ort_trace_rays(ctx, o, 0, d, 1, &hit_voxel, &face, &hit_time, nullptr), "ort_trace_rays");
			hit_direction = static_cast<direction>(face);
		}

		void sse_trace(float3 o, float3 d, direction& hit_direction, uint32_t& hit_voxel, float& hit_time) const
		{
			sse_trace(o.x, o.y, o.z, d.x, d.y, d.z, hit_direction, hit_voxel, hit_time);
		}

		void trace_frame(float3 pos, float yaw, float pitch, int W, int H, uint32_t* voxel, uint8_t* face, float* t) const
		{
			float rot[9], fov;
			ort_camera_coeffs(yaw, pitch, rot, &fov);
			const float p[3] = { pos.x, pos.y, pos.z };
			check(ort_octree_sync(tree), "ort_octree_sync");
			check(ort_trace_frame(ctx, p, rot, fov, W, H, 0, H, 1, 1, voxel, face, t, nullptr), "ort_trace_frame");
		}

		ort_ctx* context() const { return ctx; }
		ort_octree* handle() const { return tree; }

	private:
		ort_octree* tree = nullptr;
		ort_ctx* ctx = nullptr;

		void check(int rc, const char* what) const
		{
			if (rc != ORT_OK) throw std::runtime_error(std::string(what) + ": " + ort_last_error(ctx));
		}
	};
}
