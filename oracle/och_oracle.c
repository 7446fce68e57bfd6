/* och_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see och_oracle.h).
 *
 * Scalar C restatement of the reference's h_octree path.  Every function cites the reference
 * lines it follows (relative to /root/reference/Octree_Ray_Tracing/).  Build with
 *   gcc -O2 -mfma -msse2 -ffp-contract=off   (NEVER -ffast-math: DAZ/FTZ change RCPSS/FMA corner cases)
 * -mfma makes fmaf() a single VFMADD, which is what _mm_fmadd_ps is in the reference;
 * -ffp-contract=off keeps every other a*b+c as two roundings, as written.
 */
#include "och_oracle.h"

#include <immintrin.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* ============================================================================================
 * reciprocal
 * ========================================================================================== */

uint32_t oc_rcp_hw_bits(uint32_t x_bits)
{
	return f2u(_mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(u2f(x_bits)))));
}

/* Table model of RCPSS as observed on this project's hosts (Intel):
 *   NaN            -> the NaN, quieted
 *   +-0, denormal  -> +-inf
 *   +-inf          -> +-0
 *   normal x = +-2^(e-127) * 1.m :  r = tab[m >> (23-log2n)]  (RCPSS of 1.m, exponent field 126 or 127)
 *                     result exponent field = exp(r) - (e - 127); if that is <= 0 the result is +-0
 *                     (the instruction never returns a denormal), else sign | exponent | mantissa(r).
 */
uint32_t oc_rcp_table_bits(const uint32_t* tab, int log2n, uint32_t x)
{
	const uint32_t sign = x & 0x80000000u;
	const uint32_t e = (x >> 23) & 0xFFu;
	const uint32_t m = x & 0x7FFFFFu;

	if (e == 255u)
		return m ? (x | 0x00400000u) : sign;
	if (e == 0u)
		return sign | 0x7F800000u;

	const uint32_t r = tab[m >> (23 - log2n)];
	const int re = (int)(r >> 23) - ((int)e - 127);

	if (re <= 0)
		return sign;

	return sign | ((uint32_t)re << 23) | (r & 0x7FFFFFu);
}

long oc_rcp_table_from_hw(uint32_t* tab, int log2n)
{
	const int sh = 23 - log2n;
	long bad = 0;

	for (uint32_t k = 0; k < (1u << log2n); ++k)
		tab[k] = oc_rcp_hw_bits(0x3F800000u | (k << sh));

	/* all 2^23 mantissas at exponent 127, both signs */
	for (uint32_t m = 0; m < (1u << 23); ++m)
	{
		const uint32_t x = 0x3F800000u | m;
		bad += oc_rcp_hw_bits(x) != oc_rcp_table_bits(tab, log2n, x);
		bad += oc_rcp_hw_bits(x | 0x80000000u) != oc_rcp_table_bits(tab, log2n, x | 0x80000000u);
	}

	/* every exponent (incl. 0 and 255) with strided mantissas */
	for (uint32_t e = 0; e < 256; ++e)
		for (uint32_t m = 0; m < (1u << 23); m += 2039u)
		{
			const uint32_t x = (e << 23) | m;
			if (e == 255u && m)
				continue; /* NaN payload propagation is not part of the contract */
			bad += oc_rcp_hw_bits(x) != oc_rcp_table_bits(tab, log2n, x);
			bad += oc_rcp_hw_bits(x | 0x80000000u) != oc_rcp_table_bits(tab, log2n, x | 0x80000000u);
		}

	return bad;
}

static inline uint32_t rcp_bits(const uint32_t* tab, int log2n, uint32_t x)
{
	return tab ? oc_rcp_table_bits(tab, log2n, x) : oc_rcp_hw_bits(x);
}

/* ============================================================================================
 * node store -- och_h_octree.h:17-288
 * ========================================================================================== */

oc_tree* oc_tree_create(int log2cap, int depth)
{
	oc_tree* t = (oc_tree*)calloc(1, sizeof(oc_tree));
	t->log2cap = log2cap;
	t->depth = depth;
	t->cap = 1u << log2cap;
	t->idx_mask = ((t->cap - 1u) >> 4) << 4;                               /* :32 */
	t->cashes = (uint8_t*)calloc(t->cap, 1);                               /* :72-76 zero-initialised */
	t->refcounts = (uint32_t*)calloc(t->cap, 4);
	t->nodes = (uint32_t*)aligned_alloc(64, (size_t)t->cap * 32);          /* nodes[] is NOT initialised by the reference either */
	memset(t->nodes, 0, (size_t)t->cap * 32);
	return t;
}

void oc_tree_destroy(oc_tree* t)
{
	if (!t) return;
	free(t->cashes);
	free(t->refcounts);
	free(t->nodes);
	free(t);
}

/* FNV-1a over the 32 node bytes read as SIGNED char (:52-65) */
uint32_t oc_node_hash(const uint32_t c8[8])
{
	const signed char* p = (const signed char*)c8;
	uint32_t h = 0x811C9DC5u;
	for (int i = 0; i < 32; ++i)
		h = ((uint32_t)(int)p[i] ^ h) * 0x01000193u;
	return h;
}

uint32_t oc_register_node(oc_tree* t, const uint32_t c8[8])
{
	if (t->fillcnt > (uint32_t)((float)t->cap * 0.9375F))                  /* :112-116 (reference: printf + exit(0)) */
	{
		t->table_full = 1;
		return 0;
	}

	const uint32_t hash = oc_node_hash(c8);
	uint32_t index = hash & t->idx_mask;                                   /* :120 */
	uint8_t cash = (uint8_t)(hash >> t->log2cap);                          /* :122 */

	if (cash == 0) cash = 1;                                               /* :124-127 */
	else if (cash == 0xFF) cash = 0x7F;

	uint32_t last_grave = 0xFFFFFFFFu;

	while (t->cashes[index])                                               /* :131-146 */
	{
		if (t->cashes[index] == 0xFF)
			last_grave = index;
		if (t->cashes[index] == cash && memcmp(t->nodes + 8 * (size_t)index, c8, 32) == 0)
		{
			++t->nodecnt;
			++t->refcounts[index];
			return index + 1;
		}
		index = (index + 1) & (t->cap - 1);
	}

	++t->nodecnt;                                                          /* :148-159 */
	++t->fillcnt;

	if (last_grave != 0xFFFFFFFFu)
		index = last_grave;

	t->cashes[index] = cash;
	memcpy(t->nodes + 8 * (size_t)index, c8, 32);
	t->refcounts[index] = 1;

	return index + 1;
}

void oc_remove_node(oc_tree* t, uint32_t idx)                              /* :162-174 */
{
	--t->refcounts[idx - 1];
	--t->nodecnt;
	if (!t->refcounts[idx - 1])
	{
		--t->fillcnt;
		t->cashes[idx - 1] = 0xFF;
	}
}

/* 3-D Morton code, x in bits 0,3,6.., y in 1,4,7.., z in 2,5,8.. (och_z_order.cpp:191-196 does it with byte LUTs) */
uint64_t oc_z_encode_16(uint16_t x, uint16_t y, uint16_t z)
{
	uint64_t r = 0;
	for (int b = 0; b < 16; ++b)
		r |= ((uint64_t)((x >> b) & 1u) << (3 * b)) | ((uint64_t)((y >> b) & 1u) << (3 * b + 1)) | ((uint64_t)((z >> b) & 1u) << (3 * b + 2));
	return r;
}

void oc_set(oc_tree* t, uint16_t x, uint16_t y, uint16_t z, uint32_t v)   /* :176-237 */
{
	const int depth = t->depth;

	if ((x | y | z) >= (1 << depth))                                       /* :178 */
		return;

	const uint64_t index = oc_z_encode_16(x, y, z);
	uint32_t stk[16];
	int d = depth - 1;

	for (uint32_t curr = t->root_idx; curr && d >= 0; --d)                /* :188-195 */
	{
		stk[d] = curr;
		curr = t->nodes[8 * (size_t)(curr - 1) + ((index >> (3 * d)) & 7)];
	}

	uint32_t child = v;
	int made = 0;

	if (++d)                                                               /* :202-217 path ended early */
	{
		if (!v)
			return;
		while (made != d)
		{
			uint32_t n[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
			n[(index >> (3 * made++)) & 7] = child;
			child = oc_register_node(t, n);
		}
	}

	for (int i = d; i != depth; ++i)                                       /* :220-234 */
	{
		oc_remove_node(t, stk[i]);

		uint32_t n[8];
		memcpy(n, t->nodes + 8 * (size_t)(stk[i] - 1), 32);
		n[(index >> (3 * i)) & 7] = child;

		int zero = 1;
		for (int k = 0; k < 8; ++k) zero &= (n[k] == 0);

		child = zero ? 0 : oc_register_node(t, n);
	}

	t->root_idx = child;                                                   /* :236 */
}

void oc_set_many(oc_tree* t, const uint32_t* xyzv, size_t n)
{
	for (size_t i = 0; i < n; ++i)
		oc_set(t, (uint16_t)xyzv[4 * i], (uint16_t)xyzv[4 * i + 1], (uint16_t)xyzv[4 * i + 2], xyzv[4 * i + 3]);
}

uint32_t oc_at(const oc_tree* t, int x, int y, int z)                      /* :239-258 (UB on an empty tree there; 0 here) */
{
	if (!t->root_idx)
		return 0;

	const uint64_t index = oc_z_encode_16((uint16_t)x, (uint16_t)y, (uint16_t)z);
	uint32_t curr = t->root_idx;

	for (int i = t->depth - 1; i != 0; --i)
	{
		const uint32_t c = t->nodes[8 * (size_t)(curr - 1) + ((index >> (3 * i)) & 7)];
		if (!c)
			return 0;
		curr = c;
	}

	return t->nodes[8 * (size_t)(curr - 1) + (index & 7)];
}

void oc_clear(oc_tree* t)                                                  /* :285-288: only the tags */
{
	memset(t->cashes, 0, t->cap);
}

/* ============================================================================================
 * trace -- och_h_octree.h:292-447
 * ========================================================================================== */

/* The traversal of och_h_octree.h:292-447 and och_octree.cpp:167-320 -- the two differ only in how a node id
 * maps to a table row (h_octree: id-1, octree: id), where the walk starts (root_idx vs pool row 0) and what a MISS
 * reports as hit_time (INFINITY, och_h_octree.h:429, vs 0.0F, och_octree.cpp:302). */
static void trace_core(const uint32_t* nodes, uint32_t root, uint32_t index_base, float miss_t, int depth,
                       const float o[3], const float d[3],
                       const uint32_t* rcp_tab, int log2n,
                       uint32_t* vox, uint8_t* face, float* t_out, oc_counters* cnt);

void oc_trace(const uint32_t* nodes, uint32_t root, int depth,
              const float o[3], const float d[3],
              const uint32_t* rcp_tab, int log2n,
              uint32_t* vox, uint8_t* face, float* t_out, oc_counters* cnt)
{
	if (!root)                                                             /* callers' guard, test_och_h_octree.cpp:443,:535 */
	{
		*vox = 0; *face = 6; *t_out = INFINITY;
		return;
	}
	trace_core(nodes, root, 1, INFINITY, depth, o, d, rcp_tab, log2n, vox, face, t_out, cnt);
}

void oc_octree_trace(const uint32_t* pool, int depth, const float o[3], const float d[3],
                     const uint32_t* rcp_tab, int log2n, uint32_t* vox, uint8_t* face, float* t_out, oc_counters* cnt)
{
	trace_core(pool, 0, 0, 0.0F, depth, o, d, rcp_tab, log2n, vox, face, t_out, cnt);   /* och_octree.cpp:207, :217, :302 */
}

static void trace_core(const uint32_t* nodes, uint32_t root, uint32_t index_base, float miss_t, int depth,
                       const float o[3], const float d[3],
                       const uint32_t* rcp_tab, int log2n,
                       uint32_t* vox, uint8_t* face, float* t_out, oc_counters* cnt)
{
	uint64_t n_push = 0, n_step = 0, n_pop = 0;

	float    coef[3], bias[3];
	uint32_t pos[3];
	uint32_t inv = 0, idx = 0;

	for (int a = 0; a < 3; ++a)                                            /* :306-324 */
	{
		const int sg = 0.0F < d[a];                                        /* :310  false for -0, +0, NaN */
		inv |= (uint32_t)sg << a;                                          /* :322 */
		const uint32_t dn = f2u(d[a]) | 0x80000000u;                       /* :312 */
		const float oa = u2f(f2u((sg ? 3.0F : 0.0F) - o[a]) & 0x7FFFFFFFu); /* :314 */
		coef[a] = u2f(rcp_bits(rcp_tab, log2n, dn));                       /* :316 */
		bias[a] = u2f(f2u(coef[a] * oa) ^ 0x80000000u);                    /* :318 */
		pos[a] = f2u(oa) & 0x3FC00000u;                                    /* :320 */
		idx |= (uint32_t)(pos[a] == 0x3FC00000u) << a;                     /* :324 */
	}

	uint32_t dim = 1u << 22;                                               /* :326 */
	uint32_t stack[24];                                                    /* :328 parents[depth-1] */
	int      sp = 0;
	uint32_t node = root;                                                  /* :332 */
	int      level = 1;                                                    /* :334 */
	uint32_t mti = 8;                                                      /* :336 */
	float    tmin = 0.0F;                                                  /* :338 */

	for (;;)
	{
		/* PUSH :342-376 */
		++n_push;
		const uint32_t child = nodes[8 * (size_t)(node - index_base) + ((idx ^ inv) & 7u)];   /* :344 / och_octree.cpp:217 */

		if (child)
		{
			if (level++ == depth)                                          /* :346-355 HIT */
			{
				*vox = child;
				*face = (uint8_t)((mti >> 1) + 3u * ((inv & mti) == 0u));
				*t_out = tmin;
				break;
			}

			stack[sp++] = node;                                            /* :357-361 */
			node = child;
			dim >>= 1;

			idx = 0;
			for (int a = 0; a < 3; ++a)                                    /* :363-373 */
			{
				const float tm = fmaf(u2f(pos[a] | dim), coef[a], bias[a]);
				if (tm >= tmin)
				{
					idx |= 1u << a;
					pos[a] |= dim;
				}
			}
			continue;
		}

		for (;;)
		{
			/* STEP :378-419 */
			++n_step;
			const uint32_t tx = f2u(fmaf(u2f(pos[0]), coef[0], bias[0]));  /* :380-386 */
			const uint32_t ty = f2u(fmaf(u2f(pos[1]), coef[1], bias[1]));
			const uint32_t tz = f2u(fmaf(u2f(pos[2]), coef[2], bias[2]));
			int ax;

			if (tx <= ty && tx <= tz)      { ax = 0; tmin = u2f(tx); }     /* :388-406 UNSIGNED compares of the bit patterns */
			else if (ty < tx && ty <= tz)  { ax = 1; tmin = u2f(ty); }
			else                           { ax = 2; tmin = u2f(tz); }

			mti = 1u << ax;

			if (idx & mti)                                                 /* :410-419 */
			{
				pos[ax] &= ~dim;
				idx ^= mti;
				break;                                                     /* goto PUSH */
			}

			/* POP :421-446 */
			++n_pop;
			if (--level == 0)                                              /* :423-432 MISS */
			{
				*vox = 0; *face = 6; *t_out = miss_t;
				goto done;
			}

			node = stack[--sp];                                            /* :434 */
			for (int a = 0; a < 3; ++a) pos[a] &= ~dim;                    /* :436 */
			dim <<= 1;                                                     /* :438 */
			idx = 0;
			for (int a = 0; a < 3; ++a) idx |= (uint32_t)((pos[a] & dim) == dim) << a; /* :440-444 */
		}
	}

done:
	if (cnt)
	{
		cnt->push += n_push;
		cnt->step += n_step;
		cnt->pop += n_pop;
	}
}

typedef struct trace_job
{
	const uint32_t* nodes; uint32_t root; int depth; int pool_mode;
	const float* o3; int o_stride; const float* d3; size_t n;
	const uint32_t* rcp_tab; int log2n;
	uint32_t* vox; uint8_t* face; float* t; uint16_t* npush16;
	atomic_size_t next;
	pthread_mutex_t lock;
	oc_counters total;
} trace_job;

static void trace_range(trace_job* j, size_t b, size_t e, oc_counters* acc)
{
	for (size_t i = b; i < e; ++i)
	{
		oc_counters c = { 0, 0, 0 };
		if (j->pool_mode)
			oc_octree_trace(j->nodes, j->depth, j->o3 + i * (size_t)j->o_stride, j->d3 + 3 * i, j->rcp_tab, j->log2n,
			                j->vox + i, j->face + i, j->t + i, &c);
		else
			oc_trace(j->nodes, j->root, j->depth, j->o3 + i * (size_t)j->o_stride, j->d3 + 3 * i, j->rcp_tab, j->log2n,
			         j->vox + i, j->face + i, j->t + i, &c);
		if (j->npush16)
			j->npush16[i] = (uint16_t)(c.push > 65535u ? 65535u : c.push);
		acc->push += c.push; acc->step += c.step; acc->pop += c.pop;
	}
}

static void* trace_worker(void* arg)
{
	trace_job* j = (trace_job*)arg;
	oc_counters acc = { 0, 0, 0 };
	const size_t chunk = 4096;

	for (;;)
	{
		const size_t b = atomic_fetch_add(&j->next, chunk);
		if (b >= j->n) break;
		trace_range(j, b, b + chunk < j->n ? b + chunk : j->n, &acc);
	}

	pthread_mutex_lock(&j->lock);
	j->total.push += acc.push; j->total.step += acc.step; j->total.pop += acc.pop;
	pthread_mutex_unlock(&j->lock);
	return NULL;
}

static void trace_rays_mode(int pool_mode, const uint32_t* nodes, uint32_t root, int depth,
                   const float* o3, int o_stride, const float* d3, size_t n,
                   const uint32_t* rcp_tab, int log2n,
                   uint32_t* vox, uint8_t* face, float* t,
                   uint16_t* npush16, oc_counters* total, int nthreads);

void oc_trace_rays(const uint32_t* nodes, uint32_t root, int depth,
                   const float* o3, int o_stride, const float* d3, size_t n,
                   const uint32_t* rcp_tab, int log2n,
                   uint32_t* vox, uint8_t* face, float* t,
                   uint16_t* npush16, oc_counters* total, int nthreads)
{
	trace_rays_mode(0, nodes, root, depth, o3, o_stride, d3, n, rcp_tab, log2n, vox, face, t, npush16, total, nthreads);
}

void oc_octree_trace_rays(const uint32_t* pool, int depth,
                   const float* o3, int o_stride, const float* d3, size_t n,
                   const uint32_t* rcp_tab, int log2n,
                   uint32_t* vox, uint8_t* face, float* t,
                   uint16_t* npush16, oc_counters* total, int nthreads)
{
	trace_rays_mode(1, pool, 0, depth, o3, o_stride, d3, n, rcp_tab, log2n, vox, face, t, npush16, total, nthreads);
}

static void trace_rays_mode(int pool_mode, const uint32_t* nodes, uint32_t root, int depth,
                   const float* o3, int o_stride, const float* d3, size_t n,
                   const uint32_t* rcp_tab, int log2n,
                   uint32_t* vox, uint8_t* face, float* t,
                   uint16_t* npush16, oc_counters* total, int nthreads)
{
	trace_job j;
	memset(&j, 0, sizeof j);
	j.nodes = nodes; j.root = root; j.depth = depth; j.pool_mode = pool_mode;
	j.o3 = o3; j.o_stride = o_stride; j.d3 = d3; j.n = n;
	j.rcp_tab = rcp_tab; j.log2n = log2n;
	j.vox = vox; j.face = face; j.t = t; j.npush16 = npush16;
	atomic_init(&j.next, 0);
	pthread_mutex_init(&j.lock, NULL);

	if (nthreads <= 1)
		trace_worker(&j);
	else
	{
		pthread_t th[256];
		if (nthreads > 256) nthreads = 256;
		for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, trace_worker, &j);
		for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
	}

	if (total)
		*total = j.total;
	pthread_mutex_destroy(&j.lock);
}

/* ============================================================================================
 * och::octree -- the plain pointer octree (och_octree.h:10-69, och_octree.cpp:14-165)
 * ========================================================================================== */

oc_octree* oc_octree_create(int depth, uint32_t table_capacity)            /* och_octree.cpp:14, :21-35 */
{
	oc_octree* t = (oc_octree*)calloc(1, sizeof(oc_octree));
	t->depth = depth;
	t->cap = table_capacity;
	t->nodes = (uint32_t*)aligned_alloc(64, (size_t)table_capacity * 32);
	memset(t->nodes, 0, (size_t)table_capacity * 32);
	for (uint32_t i = 1; i != table_capacity; ++i) t->nodes[8 * (size_t)i] = i + 1;   /* free list through children[0] */
	t->nodes[8 * (size_t)(table_capacity - 1)] = 0;
	t->head = 1;                                                           /* och_octree.h:27-28 */
	t->node_cnt = 1;
	return t;
}

void oc_octree_destroy(oc_octree* t)
{
	if (!t) return;
	free(t->nodes);
	free(t);
}

static uint32_t octree_alloc(oc_octree* t)                                 /* och_octree.cpp:46-63 */
{
	++t->node_cnt;
	if (!t->head)
	{
		t->failed = 1;                                                     /* reference: printf + exit(0) */
		return 0;
	}
	const uint32_t old = t->head;
	t->head = t->nodes[8 * (size_t)old];
	memset(t->nodes + 8 * (size_t)old, 0, 32);
	return old;
}

static void octree_dealloc(oc_octree* t, uint32_t idx)                     /* och_octree.cpp:65-72 */
{
	--t->node_cnt;
	t->nodes[8 * (size_t)idx] = t->head;
	t->head = idx;
}

static int octree_empty(const oc_octree* t, uint32_t idx)
{
	const uint32_t* n = t->nodes + 8 * (size_t)idx;
	return !(n[0] | n[1] | n[2] | n[3] | n[4] | n[5] | n[6] | n[7]);
}

void oc_octree_set(oc_octree* t, int16_t x, int16_t y, int16_t z, uint32_t vx)   /* och_octree.cpp:74-91 (no range check) */
{
	const uint64_t index = oc_z_encode_16((uint16_t)x, (uint16_t)y, (uint16_t)z);
	uint32_t curr = 0;
	for (int i = t->depth - 1; i != 0; --i)
	{
		uint32_t* slot = t->nodes + 8 * (size_t)curr + ((index >> (3 * i)) & 7);
		if (!*slot)
		{
			const uint32_t a = octree_alloc(t);
			if (t->failed) return;
			*slot = a;
		}
		curr = *slot;
	}
	t->nodes[8 * (size_t)curr + (index & 7)] = vx;
}

void oc_octree_unset(oc_octree* t, int16_t x, int16_t y, int16_t z)       /* och_octree.cpp:93-139 */
{
	const uint64_t index = oc_z_encode_16((uint16_t)x, (uint16_t)y, (uint16_t)z);
	uint32_t curr = 0;
	int sptr = 0;
	uint32_t stack[16];

	for (int i = t->depth - 1; i != 0; --i)
	{
		const uint32_t child = t->nodes[8 * (size_t)curr + ((index >> (3 * i)) & 7)];
		if (!child)
			return;
		stack[sptr++] = curr;
		curr = child;
	}

	t->nodes[8 * (size_t)curr + (index & 7)] = 0;
	if (!octree_empty(t, curr))
		return;
	octree_dealloc(t, curr);

	for (int i = 1; i != t->depth; ++i)                                    /* (this can free row 0, the root: reference behaviour) */
	{
		--sptr;
		t->nodes[8 * (size_t)stack[sptr] + ((index >> (3 * i)) & 7)] = 0;
		if (!octree_empty(t, stack[sptr]))
			return;
		octree_dealloc(t, stack[sptr]);
	}
}

uint32_t oc_octree_at(const oc_octree* t, int16_t x, int16_t y, int16_t z) /* och_octree.cpp:141-160 */
{
	const uint64_t index = oc_z_encode_16((uint16_t)x, (uint16_t)y, (uint16_t)z);
	uint32_t curr = 0;
	for (int i = t->depth - 1; i > 0; --i)
	{
		const uint32_t c = t->nodes[8 * (size_t)curr + ((index >> (3 * i)) & 7)];
		if (!c)
			return 0;
		curr = c;
	}
	return t->nodes[8 * (size_t)curr + (index & 7)];
}

void oc_octree_apply(oc_octree* t, const int32_t* ops, size_t n)
{
	for (size_t i = 0; i < n; ++i)
	{
		const int32_t* o = ops + 5 * i;
		if (o[4] == 0) oc_octree_set(t, (int16_t)o[0], (int16_t)o[1], (int16_t)o[2], (uint32_t)o[3]);
		else oc_octree_unset(t, (int16_t)o[0], (int16_t)o[1], (int16_t)o[2]);
	}
}

/* ============================================================================================
 * camera rays -- test_och_h_octree.cpp:87-138
 * ========================================================================================== */

void oc_camera_coeffs(float yaw, float pitch, float rot[9], float* fov_factor)
{
	const float fov = 1.25F;                                               /* :95 */
	*fov_factor = 1 / tanf(fov / 2);                                       /* :97 */

	const float sin_a = 0, cos_a = 1;                                      /* :100-105 (roll fixed) */
	const float sin_b = sinf(yaw), cos_b = cosf(yaw);
	const float sin_c = sinf(pitch), cos_c = cosf(pitch);

	rot[0] = cos_a * cos_b;                                                /* :107-115 */
	rot[1] = cos_a * sin_b * sin_c - sin_a * cos_c;
	rot[2] = cos_a * sin_b * cos_c + sin_a * sin_c;
	rot[3] = sin_a * cos_b;
	rot[4] = sin_a * sin_b * sin_c + cos_a * cos_c;
	rot[5] = sin_a * sin_b * cos_c - cos_a * sin_c;
	rot[6] = -sin_b;
	rot[7] = cos_b * sin_c;
	rot[8] = cos_b * cos_c;
}

void oc_gen_rays(const float rot[9], float fov_factor, int W, int H, int y0, int y1, float* d3)
{
	const float aspect = (float)W / (float)H;                              /* :89 */
	const float vfx = 2.0F / (float)W;                                     /* :91 */
	const float vfy = 2.0F / (float)H;                                     /* :93 */

	for (int vert = y0; vert < y1; ++vert)                                 /* :119-137 */
		for (int horiz = 0; horiz < W; ++horiz)
		{
			const float u = aspect * (vfx * (float)horiz - 1.0F);          /* :123 */
			const float v = vfy * (float)vert - 1.0F;                      /* :125 */

			const float ru = u * rot[0] + v * rot[1] + fov_factor * rot[2]; /* :129-131 */
			const float rv = u * rot[3] + v * rot[4] + fov_factor * rot[5];
			const float rw = u * rot[6] + v * rot[7] + fov_factor * rot[8];

			const float rm = 1 / sqrtf(ru * ru + rv * rv + rw * rw);       /* :133 */

			float* r = d3 + 3 * ((size_t)(vert - y0) * (size_t)W + (size_t)horiz);
			r[0] = rw * rm;                                                /* :135 */
			r[1] = ru * rm;
			r[2] = -rv * rm;
		}
}

/* ============================================================================================
 * simplex noise -- och_noise.h:18-367 (Gustavson's classic simplex noise, float, with
 * truncation instead of floor and Perlin's reference permutation)
 * ========================================================================================== */

static const uint8_t k_perm[256] = {                                       /* och_noise.h:20-53 = Ken Perlin's permutation */
	151,160,137,91,90,15,131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,190,6,148,
	247,120,234,75,0,26,197,62,94,252,219,203,117,35,11,32,57,177,33,88,237,149,56,87,174,20,125,136,171,168,68,175,
	74,165,71,134,139,48,27,166,77,146,158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,244,102,143,54,
	65,25,63,161,1,216,80,73,209,76,132,187,208,89,18,169,200,196,135,130,116,188,159,86,164,100,109,198,173,186,3,64,
	52,217,226,250,124,123,5,202,38,147,118,126,255,82,85,212,207,206,59,227,47,16,58,17,182,189,28,42,223,183,170,213,
	119,248,152,2,44,154,163,70,221,153,101,155,167,43,172,9,129,22,39,253,19,98,108,110,79,113,224,232,178,185,112,104,
	218,246,97,228,251,34,242,193,238,210,144,12,191,179,162,241,81,51,145,235,249,14,239,107,49,192,214,31,181,199,106,157,
	184,84,204,176,115,121,50,45,127,4,150,254,138,236,205,93,222,114,67,29,24,72,243,141,128,195,78,66,215,61,156,180
};

static const float k_grad[12][3] = {                                       /* och_noise.h:55-59 */
	{ 1, 1, 0 }, { -1, 1, 0 }, { 1, -1, 0 }, { -1, -1, 0 },
	{ 1, 0, 1 }, { -1, 0, 1 }, { 1, 0, -1 }, { -1, 0, -1 },
	{ 0, 1, 1 }, { 0, -1, 1 }, { 0, 1, -1 }, { 0, -1, -1 }
};

static inline float corner2(float x, float y, int gi)                      /* :146-154 */
{
	float t = 0.5F - x * x - y * y;
	if (t < 0)
		return 0.0F;
	t *= t;
	return t * t * (k_grad[gi][0] * x + k_grad[gi][1] * y);
}

float oc_simplex2(float frequency, float xin, float yin)                   /* :73-179 */
{
	xin *= frequency;
	yin *= frequency;

	const float F2 = 0.5F * (0.73205078F);                                 /* :81 */
	const float G2 = (3.0F - 1.73205078F) / 6.0F;                          /* :89 */

	const float s = (xin + yin) * F2;
	const int i = (int)(xin + s);                                          /* truncation, :85-87 */
	const int j = (int)(yin + s);
	const float t = (float)(i + j) * G2;
	const float x0 = xin - ((float)i - t);
	const float y0 = yin - ((float)j - t);

	const int i1 = x0 > y0, j1 = !i1;                                      /* :106-115 */

	const float x1 = x0 - (float)i1 + G2;                                  /* :121-127 */
	const float y1 = y0 - (float)j1 + G2;
	const float x2 = x0 - 1.0F + 2.0F * G2;
	const float y2 = y0 - 1.0F + 2.0F * G2;

	const int ii = i & 255, jj = j & 255;
	const int gi0 = k_perm[(ii + k_perm[jj]) & 255] % 12;                  /* :139-143 */
	const int gi1 = k_perm[(ii + i1 + k_perm[(jj + j1) & 255]) & 255] % 12;
	const int gi2 = k_perm[(ii + 1 + k_perm[(jj + 1) & 255]) & 255] % 12;

	const float n0 = corner2(x0, y0, gi0);
	const float n1 = corner2(x1, y1, gi1);
	const float n2 = corner2(x2, y2, gi2);

	return 70.0F * (n0 + n1 + n2);                                         /* :178 */
}

static inline float corner3(float x, float y, float z, int gi)            /* :323-331 */
{
	float t = 0.6F - x * x - y * y - z * z;
	if (t < 0)
		return 0.0F;
	t *= t;
	return t * t * (k_grad[gi][0] * x + k_grad[gi][1] * y + k_grad[gi][2] * z);
}

float oc_simplex3(float frequency, float xin, float yin, float zin)        /* :181-366 */
{
	xin *= frequency;
	yin *= frequency;
	zin *= frequency;

	const float F3 = 1.0F / 3.0F;
	const float G3 = 1.0F / 6.0F;

	const float s = (xin + yin + zin) * F3;
	const int i = (int)(xin + s), j = (int)(yin + s), k = (int)(zin + s);
	const float t = (float)(i + j + k) * G3;
	const float x0 = xin - ((float)i - t);
	const float y0 = yin - ((float)j - t);
	const float z0 = zin - ((float)k - t);

	int i1, j1, k1, i2, j2, k2;                                            /* :224-281 */

	if (x0 >= y0)
	{
		if (y0 >= z0)      { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
		else if (x0 >= z0) { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 0; k2 = 1; }
		else               { i1 = 0; j1 = 0; k1 = 1; i2 = 1; j2 = 0; k2 = 1; }
	}
	else
	{
		if (y0 < z0)       { i1 = 0; j1 = 0; k1 = 1; i2 = 0; j2 = 1; k2 = 1; }
		else if (x0 < z0)  { i1 = 0; j1 = 1; k1 = 0; i2 = 0; j2 = 1; k2 = 1; }
		else               { i1 = 0; j1 = 1; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
	}

	const float x1 = x0 - (float)i1 + G3, y1 = y0 - (float)j1 + G3, z1 = z0 - (float)k1 + G3;           /* :288-304 */
	const float x2 = x0 - (float)i2 + G3 * 2.0F, y2 = y0 - (float)j2 + G3 * 2.0F, z2 = z0 - (float)k2 + G3 * 2.0F;
	const float x3 = x0 - 1.0F + G3 * 3.0F, y3 = y0 - 1.0F + G3 * 3.0F, z3 = z0 - 1.0F + G3 * 3.0F;

	const int ii = i & 255, jj = j & 255, kk = k & 255;
	const int gi0 = k_perm[(ii + k_perm[(jj + k_perm[kk]) & 255]) & 255] % 12;                           /* :313-319 */
	const int gi1 = k_perm[(ii + i1 + k_perm[(jj + j1 + k_perm[(kk + k1) & 255]) & 255]) & 255] % 12;
	const int gi2 = k_perm[(ii + i2 + k_perm[(jj + j2 + k_perm[(kk + k2) & 255]) & 255]) & 255] % 12;
	const int gi3 = k_perm[(ii + 1 + k_perm[(jj + 1 + k_perm[(kk + 1) & 255]) & 255]) & 255] % 12;

	const float n0 = corner3(x0, y0, z0, gi0);
	const float n1 = corner3(x1, y1, z1, gi1);
	const float n2 = corner3(x2, y2, z2, gi2);
	const float n3 = corner3(x3, y3, z3, gi3);

	return 32.0F * (n0 + n1 + n2 + n3);                                    /* :365 */
}

void oc_simplex2_many(float frequency, const float* xy, size_t n, float* out)
{
	for (size_t i = 0; i < n; ++i) out[i] = oc_simplex2(frequency, xy[2 * i], xy[2 * i + 1]);
}

void oc_simplex3_many(float frequency, const float* xyz, size_t n, float* out)
{
	for (size_t i = 0; i < n; ++i) out[i] = oc_simplex3(frequency, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

/* ============================================================================================
 * terrain fixture -- test_och_h_octree.cpp:561-598, 651-695, 735-787
 * ========================================================================================== */

void oc_heightmap(int depth, uint16_t* heights)                            /* :561-566, :587-592; noise = simplex_n(0.5F) (:35) */
{
	const int dim = 1 << depth;
	for (int y = 0; y < dim; ++y)
		for (int x = 0; x < dim; ++x)
		{
			const float px = (float)(x * 4) / (float)dim;
			const float py = (float)(y * 4) / (float)dim;
			heights[(size_t)y * dim + x] = (uint16_t)(int)(oc_simplex2(0.5F, px, py) * (float)dim / 16 + (float)(dim / 4));
		}
}

static uint32_t create_volume(oc_tree* t, const uint16_t* h, int hdim, int x, int y, int z, int depth) /* :651-695 */
{
	const int dim = 1 << depth;
	int active = 0;

	for (int _y = 0; _y < dim && !active; ++_y)                            /* :657-662 */
		for (int _x = 0; _x < dim; ++_x)
			if (z <= h[(size_t)(y + _y) * hdim + (x + _x)]) { active = 1; break; }

	if (!active)
		return 0;

	uint32_t n[8];

	if (depth != 1)                                                        /* :666-678 */
	{
		const int hd = dim >> 1;
		for (int c = 0; c < 8; ++c)
			n[c] = create_volume(t, h, hdim, x + (c & 1 ? hd : 0), y + (c & 2 ? hd : 0), z + (c & 4 ? hd : 0), depth - 1);
	}
	else                                                                   /* :679-689, get_leaf_val() == 1 (:600-603) */
		for (int c = 0; c < 8; ++c)
			n[c] = (z + (c >> 2)) <= h[(size_t)(y + ((c >> 1) & 1)) * hdim + (x + (c & 1))] ? 1u : 0u;

	return oc_register_node(t, n);                                         /* :694 */
}

void oc_initialize_terrain(oc_tree* t, const uint16_t* heights, const uint8_t* grass, int tunnels) /* :767-787 */
{
	const int dim = 1 << t->depth;

	t->root_idx = create_volume(t, heights, dim, 0, 0, 0, t->depth);       /* :774 */

	for (int y = 0; y < dim; ++y)                                          /* :776-783 */
		for (int x = 0; x < dim; ++x)
		{
			const uint16_t z = heights[(size_t)y * dim + x];
			oc_set(t, (uint16_t)x, (uint16_t)y, z, 2u + (grass[(size_t)y * dim + x] ? 1u : 0u));
			oc_set(t, (uint16_t)x, (uint16_t)y, (uint16_t)(z - 1), 4);
			oc_set(t, (uint16_t)x, (uint16_t)y, (uint16_t)(z - 2), 4);
		}

	if (tunnels)                                                           /* :735-743, :770, :755-763 via the global simplex_n(0.5F) */
	{
		const float scale = 1.0F / 16.0F;
		for (int z = 0; z < dim; ++z)
			for (int y = 0; y < dim; ++y)
				for (int x = 0; x < dim; ++x)
					if (!(oc_simplex3(0.5F, (float)x * scale, (float)y * scale, (float)z * scale) >= -0.5F))
						oc_set(t, (uint16_t)x, (uint16_t)y, (uint16_t)z, 0);
	}
}
