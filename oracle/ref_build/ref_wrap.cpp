// TEST INFRASTRUCTURE ONLY.  Compiles the UNMODIFIED reference sources where they
// lie (/root/reference/Octree_Ray_Tracing, passed with -I) into oracle/_ref/libochref.so
// and exposes them through a tiny C ABI so that
//   * tests/ can pin the oracle restatement (oracle/och_oracle.c) against the real thing,
//   * tests/golden/ vectors can be minted from the real thing,
//   * bench.py --impl reference / cpu_baseline can time the real och::h_octree::sse_trace.
// Nothing in the product (octree_ray_tracing_b200/) may load this library.
//
// How the reference is made to compile with g++ without touching it:
//   <intrin.h>             -> empty shim (shim/intrin.h, found via -I)
//   .m128_f32 / .m128_u32  -> shim/msvc_m128.h (union with implicit conversions)
//   INFINITY               -> <cmath> included up front
//   __forceinline          -> -D__forceinline=  (och_z_order.cpp is compiled as its own TU)
//   och_float.h:111        -> -fpermissive (dead, never instantiated template)
//   private members        -> "#define private public" around the includes, to reach
//                             table->nodes / root_idx for import/export.  Layout is unchanged.
// Built WITHOUT -ffast-math (DAZ/FTZ would break the denormal compare at och_h_octree.h:442).
#include "msvc_m128.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <thread>
#include <atomic>

#define private public
#include "och_h_octree.h"
#undef private
#include "och_octree.h"
#include "och_noise.h"
#include "opensimplex.h"

namespace {

struct tree_iface
{
	virtual ~tree_iface() {}
	virtual int      depth() const = 0;
	virtual int      log2cap() const = 0;
	virtual uint32_t register_node(const uint32_t* c8) = 0;
	virtual void     remove_node(uint32_t idx) = 0;
	virtual void     set(uint16_t x, uint16_t y, uint16_t z, uint32_t v) = 0;
	virtual uint32_t at(int x, int y, int z) = 0;
	virtual void     set_root(uint32_t r) = 0;
	virtual uint32_t get_root() = 0;
	virtual uint32_t fillcnt() const = 0;
	virtual uint32_t nodecnt() const = 0;
	virtual void     clear() = 0;
	virtual uint32_t* nodes() = 0;
	virtual uint8_t*  cashes() = 0;
	virtual uint32_t* refcounts() = 0;
	virtual void trace(const float* o, const float* d, uint32_t* vox, uint8_t* face, float* t) const = 0;
};

template<int L, int D>
struct tree_impl final : tree_iface
{
	typedef och::h_octree<L, D> tree_t;

	tree_t* tree = new tree_t;   // leaked on purpose: the reference's dtor uses a mismatched delete[] (och_h_octree.h:107)

	int depth() const override { return D; }
	int log2cap() const override { return L; }

	uint32_t register_node(const uint32_t* c8) override
	{
		typename tree_t::node n;
		for (int i = 0; i < 8; ++i) n.children[i] = c8[i];
		return tree->register_node(n);
	}
	void     remove_node(uint32_t idx) override { tree->remove_node(idx); }
	void     set(uint16_t x, uint16_t y, uint16_t z, uint32_t v) override { tree->set(x, y, z, v); }
	uint32_t at(int x, int y, int z) override { return tree->at(x, y, z); }
	void     set_root(uint32_t r) override { tree->set_root(r); }
	uint32_t get_root() override { return tree->get_root(); }
	uint32_t fillcnt() const override { return tree->get_fillcnt(); }
	uint32_t nodecnt() const override { return tree->get_nodecnt(); }
	void     clear() override { tree->clear(); }
	uint32_t* nodes() override { return reinterpret_cast<uint32_t*>(tree->table->nodes); }
	uint8_t*  cashes() override { return tree->table->cashes; }
	uint32_t* refcounts() override { return tree->table->refcounts; }

	void trace(const float* o, const float* d, uint32_t* vox, uint8_t* face, float* t) const override
	{
		och::direction dir;
		tree->sse_trace(o[0], o[1], o[2], d[0], d[1], d[2], dir, *vox, *t);
		*face = static_cast<uint8_t>(dir);
	}

	// whole batch in one virtual call so the per-ray cost is the reference's, not the wrapper's
	void trace_range(const float* o3, int o_stride, const float* d3, size_t beg, size_t end, uint32_t* vox, uint8_t* face, float* t) const
	{
		for (size_t i = beg; i < end; ++i)
		{
			const float* o = o3 + i * o_stride;
			const float* d = d3 + i * 3;
			och::direction dir;
			tree->sse_trace(o[0], o[1], o[2], d[0], d[1], d[2], dir, vox[i], t[i]);
			face[i] = static_cast<uint8_t>(dir);
		}
	}
};

template<int L, int D>
void trace_batch_t(const tree_iface* ti, const float* o3, int o_stride, const float* d3, size_t n, uint32_t* vox, uint8_t* face, float* t, int nthreads)
{
	const tree_impl<L, D>* impl = static_cast<const tree_impl<L, D>*>(ti);

	if (nthreads <= 1)
	{
		impl->trace_range(o3, o_stride, d3, 0, n, vox, face, t);
		return;
	}

	// dynamic chunks of consecutive rays: same per-ray code, just more cores
	std::atomic<size_t> next{ 0 };
	const size_t chunk = 4096;
	std::vector<std::thread> pool;
	for (int w = 0; w < nthreads; ++w)
		pool.emplace_back([&]() {
			for (;;)
			{
				size_t b = next.fetch_add(chunk);
				if (b >= n) break;
				size_t e = b + chunk < n ? b + chunk : n;
				impl->trace_range(o3, o_stride, d3, b, e, vox, face, t);
			}
		});
	for (auto& th : pool) th.join();
}

#define OCHREF_CONFIGS(X) X(12, 4) X(16, 6) X(19, 8) X(22, 10) X(24, 12) X(25, 13) X(25, 14)

}

extern "C" {

void* ochref_tree_create(int log2cap, int depth)
{
#define X(L, D) if (log2cap == L && depth == D) return static_cast<tree_iface*>(new tree_impl<L, D>);
	OCHREF_CONFIGS(X)
#undef X
	return nullptr;
}

int      ochref_tree_depth(void* h) { return static_cast<tree_iface*>(h)->depth(); }
int      ochref_tree_log2cap(void* h) { return static_cast<tree_iface*>(h)->log2cap(); }
uint32_t ochref_register_node(void* h, const uint32_t* c8) { return static_cast<tree_iface*>(h)->register_node(c8); }
void     ochref_remove_node(void* h, uint32_t idx) { static_cast<tree_iface*>(h)->remove_node(idx); }
void     ochref_set(void* h, uint16_t x, uint16_t y, uint16_t z, uint32_t v) { static_cast<tree_iface*>(h)->set(x, y, z, v); }
uint32_t ochref_at(void* h, int x, int y, int z) { return static_cast<tree_iface*>(h)->at(x, y, z); }
void     ochref_set_root(void* h, uint32_t r) { static_cast<tree_iface*>(h)->set_root(r); }
uint32_t ochref_get_root(void* h) { return static_cast<tree_iface*>(h)->get_root(); }
uint32_t ochref_get_fillcnt(void* h) { return static_cast<tree_iface*>(h)->fillcnt(); }
uint32_t ochref_get_nodecnt(void* h) { return static_cast<tree_iface*>(h)->nodecnt(); }
void     ochref_clear(void* h) { static_cast<tree_iface*>(h)->clear(); }
uint32_t* ochref_nodes(void* h) { return static_cast<tree_iface*>(h)->nodes(); }
uint8_t*  ochref_cashes(void* h) { return static_cast<tree_iface*>(h)->cashes(); }
uint32_t* ochref_refcounts(void* h) { return static_cast<tree_iface*>(h)->refcounts(); }

// batched set(): xyzv = n x 4 uint32 (x, y, z, v), applied in order
void ochref_set_many(void* h, const uint32_t* xyzv, size_t n)
{
	tree_iface* t = static_cast<tree_iface*>(h);
	for (size_t i = 0; i < n; ++i)
		t->set(static_cast<uint16_t>(xyzv[4 * i]), static_cast<uint16_t>(xyzv[4 * i + 1]), static_cast<uint16_t>(xyzv[4 * i + 2]), xyzv[4 * i + 3]);
}

// Import a compact node array (ids 1..n, id i lives in slot i-1) so the reference's own
// sse_trace can run over a DAG that another builder produced.  Only nodes[]/root_idx are
// written -- the tree is trace-only afterwards (hash tags are not rebuilt).
void ochref_import_compact(void* h, const uint32_t* nodes8, size_t n, uint32_t root)
{
	tree_iface* t = static_cast<tree_iface*>(h);
	std::memcpy(t->nodes(), nodes8, n * 32);
	t->set_root(root);
}

// o_stride = 3 for per-ray origins, 0 for one shared origin (camera)
int ochref_trace_batch(void* h, const float* o3, int o_stride, const float* d3, size_t n, uint32_t* vox, uint8_t* face, float* t, int nthreads)
{
	tree_iface* ti = static_cast<tree_iface*>(h);
	if (!ti->get_root())
	{
		// the reference's callers guard the empty tree (test_och_h_octree.cpp:443, :535)
		for (size_t i = 0; i < n; ++i) { vox[i] = 0; face[i] = 6; t[i] = INFINITY; }
		return 0;
	}
#define X(L, D) if (ti->log2cap() == L && ti->depth() == D) { trace_batch_t<L, D>(ti, o3, o_stride, d3, n, vox, face, t, nthreads); return 0; }
	OCHREF_CONFIGS(X)
#undef X
	return -1;
}

uint64_t ochref_z_encode_16(uint16_t x, uint16_t y, uint16_t z) { return och::z_encode_16(x, y, z); }

// och::simplex_n (och_noise.h:73, :181)
void ochref_simplex2(float freq, const float* xy, size_t n, float* out)
{
	och::simplex_n noise(freq);
	for (size_t i = 0; i < n; ++i) out[i] = noise(xy[2 * i], xy[2 * i + 1]);
}
void ochref_simplex3(float freq, const float* xyz, size_t n, float* out)
{
	och::simplex_n noise(freq);
	for (size_t i = 0; i < n; ++i) out[i] = noise(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}
// OpenSimplexNoise(seed).Evaluate(x, y) (opensimplex.h)
void ochref_opensimplex2(int64_t seed, const double* xy, size_t n, double* out)
{
	OpenSimplexNoise noise(seed);
	for (size_t i = 0; i < n; ++i) out[i] = noise.Evaluate(xy[2 * i], xy[2 * i + 1]);
}

// ---- fixture builder driven through the REAL reference table code ---------------------------
// test_och_h_octree.cpp (the builders' TU) cannot be compiled here (<Windows.h>, olc), so the
// recursion below restates create_volume/initialize_h_octree (:561-566, :651-695, :767-787) but
// every node goes through the reference's own register_node/set, and every noise value comes
// from the reference's own och::simplex_n.

void ochref_heightmap(int depth, uint16_t* heights)
{
	och::simplex_n noise(0.5F);                                            // test_och_h_octree.cpp:35
	const int dim = 1 << depth;
	for (int y = 0; y < dim; ++y)
		for (int x = 0; x < dim; ++x)
		{
			float px = (static_cast<float>(x * 4) / dim);                  // :563-566
			float py = (static_cast<float>(y * 4) / dim);
			heights[static_cast<size_t>(y) * dim + x] = static_cast<uint16_t>(static_cast<int>(noise(px, py) * dim / 16 + dim / 4));
		}
}

static uint32_t ref_create_volume(tree_iface* t, const uint16_t* h, int hdim, int x, int y, int z, int depth)
{
	const int dim = 1 << depth;
	bool active = false;
	for (int _y = 0; _y < dim && !active; ++_y)
		for (int _x = 0; _x < dim; ++_x)
			if (z <= h[static_cast<size_t>(y + _y) * hdim + (x + _x)]) { active = true; break; }
	if (!active)
		return 0;
	uint32_t n[8];
	if (depth != 1)
	{
		const int hd = dim >> 1;
		for (int c = 0; c < 8; ++c)
			n[c] = ref_create_volume(t, h, hdim, x + (c & 1 ? hd : 0), y + (c & 2 ? hd : 0), z + (c & 4 ? hd : 0), depth - 1);
	}
	else
		for (int c = 0; c < 8; ++c)
			n[c] = (z + (c >> 2)) <= h[static_cast<size_t>(y + ((c >> 1) & 1)) * hdim + (x + (c & 1))] ? 1u : 0u;
	return t->register_node(n);
}

void ochref_initialize_terrain(void* hnd, const uint16_t* heights, const uint8_t* grass, int tunnels)
{
	tree_iface* t = static_cast<tree_iface*>(hnd);
	const int dim = 1 << t->depth();
	t->set_root(ref_create_volume(t, heights, dim, 0, 0, 0, t->depth()));
	for (int y = 0; y < dim; ++y)
		for (int x = 0; x < dim; ++x)
		{
			uint16_t z = heights[static_cast<size_t>(y) * dim + x];
			t->set(x, y, z, 2 + (grass[static_cast<size_t>(y) * dim + x] ? 1 : 0));
			t->set(x, y, z - 1, 4);
			t->set(x, y, z - 2, 4);
		}
	if (tunnels)
	{
		och::simplex_n noise(0.5F);
		const float scale = 1.0F / 16.0F;
		for (int z = 0; z < dim; ++z)
			for (int y = 0; y < dim; ++y)
				for (int x = 0; x < dim; ++x)
				{
					float val = static_cast<float>(noise(static_cast<float>(x) * scale, static_cast<float>(y) * scale, static_cast<float>(z) * scale));
					if (!(val >= -0.5F))
						t->set(x, y, z, 0);
				}
	}
}

// ---- och::octree (och_octree.h, och_octree.cpp): the plain pointer octree with the same traversal ----------

void* ochref_octree_create(int depth, uint32_t table_capacity) { return new och::octree(static_cast<uint16_t>(depth), table_capacity); }
void  ochref_octree_set(void* h, int16_t x, int16_t y, int16_t z, uint32_t v) { static_cast<och::octree*>(h)->set(x, y, z, v); }
void  ochref_octree_unset(void* h, int16_t x, int16_t y, int16_t z) { static_cast<och::octree*>(h)->unset(x, y, z); }
uint32_t ochref_octree_at(void* h, int16_t x, int16_t y, int16_t z) { return static_cast<och::octree*>(h)->at(x, y, z); }
int   ochref_octree_node_cnt(void* h) { return static_cast<och::octree*>(h)->get_node_cnt(); }
uint32_t* ochref_octree_nodes(void* h) { return reinterpret_cast<uint32_t*>(static_cast<och::octree*>(h)->_table); }

// ops: n x (x, y, z, v, kind) int32 with kind 0 = set, 1 = unset
void ochref_octree_apply(void* h, const int32_t* ops, size_t n)
{
	och::octree* t = static_cast<och::octree*>(h);
	for (size_t i = 0; i < n; ++i)
	{
		const int32_t* o = ops + 5 * i;
		if (o[4] == 0) t->set(static_cast<int16_t>(o[0]), static_cast<int16_t>(o[1]), static_cast<int16_t>(o[2]), static_cast<uint32_t>(o[3]));
		else t->unset(static_cast<int16_t>(o[0]), static_cast<int16_t>(o[1]), static_cast<int16_t>(o[2]));
	}
}

void ochref_octree_trace_batch(void* h, const float* o3, int o_stride, const float* d3, size_t n, uint32_t* vox, uint8_t* face, float* t)
{
	const och::octree* tr = static_cast<const och::octree*>(h);
	for (size_t i = 0; i < n; ++i)
	{
		const float* o = o3 + i * o_stride;
		const float* d = d3 + i * 3;
		och::direction dir;
		tr->sse_trace(o[0], o[1], o[2], d[0], d[1], d[2], dir, vox[i], t[i]);
		face[i] = static_cast<uint8_t>(dir);
	}
}

uint32_t ochref_node_hash(const uint32_t* c8)
{
	och::h_octree<12, 4>::node n;
	for (int i = 0; i < 8; ++i) n.children[i] = c8[i];
	return n.hash();
}

}
