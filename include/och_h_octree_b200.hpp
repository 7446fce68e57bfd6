// och_h_octree_b200.hpp -- the reference's C++ interface for the trace path, rebuilt over libort_b200.so.
//
// `och::h_octree<Log2_table_capacity, Depth>` below has the public members of the reference class of the
// same name (och_h_octree.h:17-452): constants depth/dim/log2_table_capacity/table_capacity/voxel_dim,
// struct node {children[8]; operator==; is_zero; hash}, register_node, remove_node, set, at, set_root,
// get_root, get_fillcnt, get_nodecnt, get_max_refcnt, clear and both sse_trace overloads with their
// out-reference results.  A translation unit of the reference that includes this header instead of
// "och_h_octree.h"  compiles unchanged and traces on the GPU.
//
// Differences by design: tracing runs on a B200 (no CPU path: construction throws std::runtime_error if no
// sm_100 device is usable); an empty tree traces to direction::exit instead of dereferencing nodes[-1];
// the table-full condition throws instead of exit(0).  Added: trace_rays / trace_frame / trace_frame_rgba (batched), fill_box, sync(), save / load.
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

#include "ort_b200.h"

namespace och
{
#ifndef OCH_B200_HAVE_REFERENCE_TYPES   // define when och_tree_helper.h / och_vec.h of the reference are included too
	enum class direction   // och_tree_helper.h:7-18
	{
		x_pos = 0, y_pos = 1, z_pos = 2, x_neg = 3, y_neg = 4, z_neg = 5, exit = 6, inside = 7, error = 8
	};

	struct float3 { float x, y, z; };   // och_vec.h: vec3<float>
#endif

	template<int Log2_table_capacity, int Depth>
	class h_octree
	{
	public:
		static constexpr int depth = Depth;
		static constexpr int dim = 1 << Depth;
		static constexpr int log2_table_capacity = Log2_table_capacity;
		static constexpr int table_capacity = 1 << Log2_table_capacity;
		static constexpr float voxel_dim = 1.0F / dim;
		static constexpr size_t table_bytes = static_cast<size_t>(table_capacity) * 37;   // cashes + refcounts + nodes per slot

		struct node
		{
			alignas(32) uint32_t children[8];

			bool operator==(const node& n) const { return std::memcmp(children, n.children, 32) == 0; }

			bool is_zero() const
			{
				return !(children[0] | children[1] | children[2] | children[3] | children[4] | children[5] | children[6] | children[7]);
			}

			uint32_t hash() const   // FNV-1a over the bytes as signed char (och_h_octree.h:52-65)
			{
				const signed char* b = reinterpret_cast<const signed char*>(children);
				uint32_t h = 0x811C9DC5u;
				for (int i = 0; i < 32; ++i) h = (static_cast<uint32_t>(static_cast<int>(b[i])) ^ h) * 0x01000193u;
				return h;
			}
		};

		explicit h_octree(int device = 0)
		{
			check(ort_tree_create(&tree, Log2_table_capacity, Depth), "ort_tree_create");
			int rc = ort_create(&ctx, device, Depth, 1u << 16);
			if (rc != ORT_OK)
			{
				std::string msg = std::string("ort_create: ") + ort_last_error(nullptr);
				ort_tree_destroy(tree);
				throw std::runtime_error(msg);
			}
			ort_tree_attach(tree, ctx);
		}

		~h_octree()
		{
			ort_tree_destroy(tree);
			ort_destroy(ctx);
		}

		h_octree(const h_octree&) = delete;
		h_octree& operator=(const h_octree&) = delete;

		uint32_t register_node(const node& n)
		{
			uint32_t idx = ort_tree_register_node(tree, n.children);
			if (!idx) throw std::runtime_error("h_octree: table too full");   // reference: printf + exit(0) (:112-116)
			return idx;
		}

		void remove_node(const uint32_t idx) { ort_tree_remove_node(tree, idx); }

		void set(uint16_t x, uint16_t y, uint16_t z, uint32_t v)
		{
			ort_tree_set(tree, x, y, z, v);
			if (ort_tree_table_full(tree)) throw std::runtime_error("h_octree: table too full");
		}

		// extension: the T/Z box edit in one pass (same result as the set() loop, see ort_tree_fill_box)
		void fill_box(int x0, int y0, int z0, int x1, int y1, int z1, uint32_t v)
		{
			ort_tree_fill_box(tree, x0, y0, z0, x1, y1, z1, v);
			if (ort_tree_table_full(tree)) throw std::runtime_error("h_octree: table too full");
		}

		uint32_t at(int x, int y, int z) { return ort_tree_at(tree, x, y, z); }

		void set_root(uint32_t idx) { ort_tree_set_root(tree, idx); }
		uint32_t get_root() { return ort_tree_get_root(tree); }
		uint32_t get_fillcnt() const { return ort_tree_get_fillcnt(tree); }
		uint32_t get_nodecnt() const { return ort_tree_get_nodecnt(tree); }
		uint32_t get_max_refcnt() const { return ort_tree_get_max_refcnt(tree); }
		void clear() { ort_tree_clear(tree); }

		// TRACING (och_h_octree.h:292-452).  const like the reference's; the device mirror is brought up to date first.
		void sse_trace(float ox, float oy, float oz, float dx, float dy, float dz, direction& hit_direction, uint32_t& hit_voxel, float& hit_time) const
		{
			const float o[3] = { ox, oy, oz }, d[3] = { dx, dy, dz };
			uint8_t face = 8;
			sync();
			check(ort_trace_rays(ctx, o, 0, d, 1, &hit_voxel, &face, &hit_time, nullptr), "ort_trace_rays");
			hit_direction = static_cast<direction>(face);
		}

		void sse_trace(float3 o, float3 d, direction& hit_direction, uint32_t& hit_voxel, float& hit_time) const
		{
			sse_trace(o.x, o.y, o.z, d.x, d.y, d.z, hit_direction, hit_voxel, hit_time);
		}

		// ---- batched forms (what tree_camera::update_position + update_image do per frame) -------------
		void sync() const { check(ort_tree_sync(tree), "ort_tree_sync"); }

		// n rays; o3 has 3 floats per ray (o_stride = 3) or one shared origin (o_stride = 0)
		void trace_rays(const float* o3, int o_stride, const float* d3, size_t n, uint32_t* voxel, uint8_t* face, float* t) const
		{
			sync();
			check(ort_trace_rays(ctx, o3, o_stride, d3, n, voxel, face, t, nullptr), "ort_trace_rays");
		}

		// one W x H frame from a camera at pos looking (yaw, pitch) = tree_camera::{pos, dir}
		void trace_frame(float3 pos, float yaw, float pitch, int W, int H, uint32_t* voxel, uint8_t* face, float* t) const
		{
			float rot[9], fov;
			ort_camera_coeffs(yaw, pitch, rot, &fov);
			const float p[3] = { pos.x, pos.y, pos.z };
			sync();
			check(ort_trace_frame(ctx, p, rot, fov, W, H, 0, H, 1, 1, voxel, face, t, nullptr), "ort_trace_frame");
		}

		// the pixels update_image would Draw() (test_och_h_octree.cpp:64-85, :437-457): colours = voxels.get_colours()
		// (6 olc::Pixel::n values per voxel type), one uint32 per pixel
		void set_palette(const uint32_t* colours6, uint32_t n_voxels, uint32_t exit_rgba = 0xFFFEBF00u, uint32_t inside_rgba = 0xFF07193Fu)
		{
			check(ort_set_palette(ctx, colours6, n_voxels, exit_rgba, inside_rgba), "ort_set_palette");
		}

		void trace_frame_rgba(float3 pos, float yaw, float pitch, int W, int H, uint32_t* rgba) const
		{
			float rot[9], fov;
			ort_camera_coeffs(yaw, pitch, rot, &fov);
			const float p[3] = { pos.x, pos.y, pos.z };
			sync();
			check(ort_trace_frame_rgba(ctx, p, rot, fov, W, H, 0, H, 1, 1, rgba), "ort_trace_frame_rgba");
		}

		// frame loops: with deferred completion the trace_frame* calls above return once their work is queued and
		// wait_frames() collects the results -- frame k+1 is traced while frame k's pixels cross PCIe
		void defer_completion(bool on) { check(ort_set_option(ctx, "defer_sync", on ? 1 : 0), "ort_set_option"); }
		void wait_frames() const { check(ort_sync(ctx), "ort_sync"); }

		// table dump / load (see ort_tree_save)
		void save(const char* path) const { check(ort_tree_save(tree, path), "ort_tree_save"); }
		void load(const char* path) { check(ort_tree_load(tree, path), "ort_tree_load"); }

		ort_ctx* context() const { return ctx; }
		ort_tree* handle() const { return tree; }

	private:
		ort_tree* tree = nullptr;
		ort_ctx* ctx = nullptr;

		void check(int rc, const char* what) const
		{
			if (rc != ORT_OK) throw std::runtime_error(std::string(what) + ": " + ort_last_error(ctx));
		}
	};
}
