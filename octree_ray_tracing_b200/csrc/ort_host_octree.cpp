// ort_host_octree.cpp -- host side of och::octree (och_octree.h:10-69, och_octree.cpp:14-165): the plain pointer
// octree over a fixed node pool with an intrusive free list.  Row 0 is the root, child values are raw row numbers
// (0 = empty), nodes are edited in place.  The device mirror is the pool itself (row for row), so the tracer runs
// with index base 0, root 0 and the reference's MISS time 0.0F (och_octree.cpp:302); edits mark rows dirty and
// sync() ships only those.  Pool contents match the reference's row for row (tests/test_octree.py).
#include "ort_internal.h"

#include <cstdlib>
#include <cstring>
#include <new>

namespace {

inline uint64_t morton3_16(uint16_t x, uint16_t y, uint16_t z)   // och::z_encode_16 (och_z_order.cpp:191-196)
{
	auto spread = [](uint64_t v) {
		v = (v | (v << 32)) & 0x001F00000000FFFFull;
		v = (v | (v << 16)) & 0x001F0000FF0000FFull;
		v = (v | (v << 8)) & 0x100F00F00F00F00Full;
		v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
		v = (v | (v << 2)) & 0x1249249249249249ull;
		return v;
	};
	return spread(x) | (spread(y) << 1) | (spread(z) << 2);
}

}  // namespace

struct ort_octree
{
	int       depth;
	uint32_t  cap;
	uint32_t* nodes = nullptr;        // cap * 8
	uint32_t  head = 1;               // och_octree.h:27
	int       node_cnt = 1;           // och_octree.h:28 (root counted)
	bool      failed = false;         // pool exhausted: the reference prints and exit(0)s (och_octree.cpp:50-54)

	ort_ctx*  ctx = nullptr;
	bool      mirror_valid = false;
	uint32_t  high_water = 1;         // rows [0, high_water) have ever been used
	uint64_t* dirty_bits = nullptr;
	std::vector<uint32_t> dirty;
	std::vector<uint32_t> up_ids, up_nodes;
	uint64_t  last_sync_nodes = 0;
	int       last_sync_full = 0;

	ort_octree(int depth_, uint32_t cap_) : depth(depth_), cap(cap_)
	{
		nodes = static_cast<uint32_t*>(std::aligned_alloc(64, (static_cast<size_t>(cap) * 32 + 63) / 64 * 64));
		dirty_bits = static_cast<uint64_t*>(std::calloc((cap + 63) / 64, 8));
		if (!nodes || !dirty_bits) return;
		std::memset(nodes, 0, static_cast<size_t>(cap) * 32);
		for (uint32_t i = 1; i != cap; ++i) nodes[8 * static_cast<size_t>(i)] = i + 1;   // och_octree.cpp:29
		nodes[8 * static_cast<size_t>(cap - 1)] = 0;                                       // :31
	}
	~ort_octree() { std::free(nodes); std::free(dirty_bits); }

	void touch(uint32_t row)
	{
		uint64_t& w = dirty_bits[row >> 6];
		const uint64_t bit = 1ull << (row & 63);
		if (!(w & bit)) { w |= bit; dirty.push_back(row); }
		if (row >= high_water) high_water = row + 1;
	}

	uint32_t alloc()                                               // och_octree.cpp:46-63
	{
		++node_cnt;
		if (!head) { failed = true; return 0; }
		const uint32_t old = head;
		head = nodes[8 * static_cast<size_t>(old)];
		std::memset(nodes + 8 * static_cast<size_t>(old), 0, 32);
		touch(old);
		return old;
	}

	void dealloc(uint32_t row)                                     // och_octree.cpp:65-72
	{
		--node_cnt;
		nodes[8 * static_cast<size_t>(row)] = head;
		head = row;
		touch(row);
	}

	bool empty(uint32_t row) const
	{
		const uint32_t* n = nodes + 8 * static_cast<size_t>(row);
		return !(n[0] | n[1] | n[2] | n[3] | n[4] | n[5] | n[6] | n[7]);
	}

	void set(int16_t x, int16_t y, int16_t z, uint32_t vx)         // och_octree.cpp:74-91 -- note: no range check there
	{
		const uint64_t key = morton3_16(static_cast<uint16_t>(x), static_cast<uint16_t>(y), static_cast<uint16_t>(z));
		uint32_t cur = 0;
		for (int d = depth - 1; d != 0; --d)
		{
			uint32_t* slot = nodes + 8 * static_cast<size_t>(cur) + ((key >> (3 * d)) & 7);
			if (!*slot)
			{
				const uint32_t a = alloc();
				if (failed) return;
				*slot = a;
				touch(cur);
			}
			cur = *slot;
		}
		nodes[8 * static_cast<size_t>(cur) + (key & 7)] = vx;
		touch(cur);
	}

	void unset(int16_t x, int16_t y, int16_t z)                    // och_octree.cpp:93-139
	{
		const uint64_t key = morton3_16(static_cast<uint16_t>(x), static_cast<uint16_t>(y), static_cast<uint16_t>(z));
		uint32_t cur = 0, path[16];
		int sp = 0;
		for (int d = depth - 1; d != 0; --d)
		{
			const uint32_t child = nodes[8 * static_cast<size_t>(cur) + ((key >> (3 * d)) & 7)];
			if (!child) return;
			path[sp++] = cur;
			cur = child;
		}
		nodes[8 * static_cast<size_t>(cur) + (key & 7)] = 0;
		touch(cur);
		if (!empty(cur)) return;
		dealloc(cur);
		for (int d = 1; d != depth; ++d)                           // may free row 0 (the root) when the tree empties: reference behaviour
		{
			--sp;
			nodes[8 * static_cast<size_t>(path[sp]) + ((key >> (3 * d)) & 7)] = 0;
			touch(path[sp]);
			if (!empty(path[sp])) return;
			dealloc(path[sp]);
		}
	}

	uint32_t at(int16_t x, int16_t y, int16_t z) const             // och_octree.cpp:141-160
	{
		const uint64_t key = morton3_16(static_cast<uint16_t>(x), static_cast<uint16_t>(y), static_cast<uint16_t>(z));
		uint32_t cur = 0;
		for (int d = depth - 1; d > 0; --d)
		{
			cur = nodes[8 * static_cast<size_t>(cur) + ((key >> (3 * d)) & 7)];
			if (!cur) return 0;
		}
		return nodes[8 * static_cast<size_t>(cur) + (key & 7)];
	}

	void clear_dirty()
	{
		for (uint32_t r : dirty) dirty_bits[r >> 6] = 0;
		dirty.clear();
	}

	int sync()
	{
		if (!ctx) return ORT_ERR_NOT_ATTACHED;
		int rc;
		if (!mirror_valid || dirty.size() * 2 > high_water)
		{
			rc = ort_upload_pool(ctx, nodes, high_water);
			last_sync_nodes = high_water;
			last_sync_full = 1;
		}
		else
		{
			up_ids.clear();
			up_nodes.clear();
			for (uint32_t r : dirty)
			{
				up_ids.push_back(r + 1);                                   // delta ids are 1-based rows
				up_nodes.insert(up_nodes.end(), nodes + 8 * static_cast<size_t>(r), nodes + 8 * static_cast<size_t>(r) + 8);
			}
			rc = ort_upload_delta(ctx, up_ids.data(), up_nodes.data(), up_ids.size(), 0);
			if (rc == ORT_ERR_CAPACITY)
			{
				rc = ort_upload_pool(ctx, nodes, high_water);
				last_sync_full = 1;
				last_sync_nodes = high_water;
			}
			else
			{
				last_sync_nodes = up_ids.size();
				last_sync_full = 0;
			}
		}
		if (rc == ORT_OK) { clear_dirty(); mirror_valid = true; }
		else mirror_valid = false;
		return rc;
	}
};

extern "C" {

int ort_octree_create(ort_octree** out, int depth, uint32_t table_capacity)
{
	if (!out || depth < 1 || depth > 16 || table_capacity < 2 || table_capacity > 0x1FFFFFFFu)
		return ort_fail(nullptr, ORT_ERR_INVALID, "ort_octree_create: depth in 1..16 and 2 <= table_capacity < 2^29 required");
	ort_octree* t = new (std::nothrow) ort_octree(depth, table_capacity);
	if (!t || !t->nodes || !t->dirty_bits)
	{
		delete t;
		return ort_fail(nullptr, ORT_ERR_INVALID, "ort_octree_create: out of host memory");
	}
	*out = t;
	return ORT_OK;
}

void     ort_octree_destroy(ort_octree* t) { delete t; }
void     ort_octree_set(ort_octree* t, int16_t x, int16_t y, int16_t z, uint32_t vx) { t->set(x, y, z, vx); }
void     ort_octree_unset(ort_octree* t, int16_t x, int16_t y, int16_t z) { t->unset(x, y, z); }
uint32_t ort_octree_at(const ort_octree* t, int16_t x, int16_t y, int16_t z) { return t->at(x, y, z); }
int      ort_octree_get_node_cnt(const ort_octree* t) { return t->node_cnt; }
int      ort_octree_failed(const ort_octree* t) { return t->failed; }
int      ort_octree_depth(const ort_octree* t) { return t->depth; }
uint32_t ort_octree_table_capacity(const ort_octree* t) { return t->cap; }
const uint32_t* ort_octree_nodes(const ort_octree* t) { return t->nodes; }

void ort_octree_apply(ort_octree* t, const int32_t* ops, size_t n)
{
	for (size_t i = 0; i < n; ++i)
	{
		const int32_t* o = ops + 5 * i;
		if (o[4] == 0) t->set(static_cast<int16_t>(o[0]), static_cast<int16_t>(o[1]), static_cast<int16_t>(o[2]), static_cast<uint32_t>(o[3]));
		else t->unset(static_cast<int16_t>(o[0]), static_cast<int16_t>(o[1]), static_cast<int16_t>(o[2]));
	}
}

int ort_octree_attach(ort_octree* t, ort_ctx* ctx)
{
	t->ctx = ctx;
	t->mirror_valid = false;
	return ORT_OK;
}

int ort_octree_sync(ort_octree* t) { return t->sync(); }

void ort_octree_sync_stats(const ort_octree* t, uint64_t* n, int* full)
{
	if (n) *n = t->last_sync_nodes;
	if (full) *full = t->last_sync_full;
}

}  // extern "C"
