/* och_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's h_octree hot path (AlexanderRipar/Octree_Ray_Tracing),
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER.
 * Nothing under octree_ray_tracing_b200/ may include, link or load it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here bit-for-bit
 * against the real reference compiled in place from /root/reference (oracle/ref_build ->
 * oracle/_ref/libochref.so), and tests/golden/ holds vectors minted from that build.
 *
 * Reference citations are relative to /root/reference/Octree_Ray_Tracing/.
 */
#ifndef OCH_ORACLE_H
#define OCH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- reciprocal (och_h_octree.h:316 uses the CPU's RCPPS) -------------------------------- */

/* hardware RCPSS of this host */
uint32_t oc_rcp_hw_bits(uint32_t x_bits);

/* Derive tab[1<<log2n] (result bit patterns for inputs 1.0 <= x < 2.0, indexed by the top log2n
 * mantissa bits) from this host's RCPSS.  Returns 0 if the host's instruction really is a pure
 * function of those bits and scales exactly with the exponent (checked over all 2^23 mantissas
 * and sampled over all exponents and both signs), else the number of mismatches. */
long oc_rcp_table_from_hw(uint32_t* tab, int log2n);

/* table model of RCPSS: see oc_rcp_table_bits() in the .c for the exact rule */
uint32_t oc_rcp_table_bits(const uint32_t* tab, int log2n, uint32_t x_bits);

/* ---- node store (och_h_octree.h:17-288) --------------------------------------------------- */

typedef struct oc_tree
{
	int       log2cap;
	int       depth;
	uint32_t  cap;
	uint32_t  idx_mask;
	uint8_t*  cashes;     /* sic: the reference's name for the 1-byte hash tag          (:78) */
	uint32_t* refcounts;  /*                                                             (:80) */
	uint32_t* nodes;      /* cap * 8 children                                            (:82) */
	uint32_t  root_idx;   /*                                                             (:95) */
	uint32_t  fillcnt;    /*                                                             (:97) */
	uint32_t  nodecnt;    /*                                                             (:98) */
	int       table_full; /* set instead of the reference's printf+exit(0)          (:112-116) */
} oc_tree;

oc_tree* oc_tree_create(int log2cap, int depth);
void     oc_tree_destroy(oc_tree* t);
uint32_t oc_node_hash(const uint32_t c8[8]);                       /* :52-65   */
uint32_t oc_register_node(oc_tree* t, const uint32_t c8[8]);       /* :110-160 */
void     oc_remove_node(oc_tree* t, uint32_t idx);                 /* :162-174 */
void     oc_set(oc_tree* t, uint16_t x, uint16_t y, uint16_t z, uint32_t v); /* :176-237 */
void     oc_set_many(oc_tree* t, const uint32_t* xyzv, size_t n);
uint32_t oc_at(const oc_tree* t, int x, int y, int z);             /* :239-258 */
void     oc_clear(oc_tree* t);                                     /* :285-288 */
uint64_t oc_z_encode_16(uint16_t x, uint16_t y, uint16_t z);       /* och_z_order.cpp:191-196 */

/* ---- trace (och_h_octree.h:292-447) -------------------------------------------------------- */

typedef struct oc_counters
{
	uint64_t push;   /* evaluations of label PUSH = child-slot loads (:344) */
	uint64_t step;   /* evaluations of label STEP (:378) */
	uint64_t pop;    /* evaluations of label POP  (:421) */
} oc_counters;

/* One ray.  nodes = table->nodes (1-based ids), root = root_idx.  rcp_tab == NULL uses the
 * host's RCPSS like the reference; otherwise the table model.  root == 0 returns MISS like the
 * reference's callers do (test_och_h_octree.cpp:443, :535).  npush (optional) receives the
 * number of PUSH evaluations of this ray. */
void oc_trace(const uint32_t* nodes, uint32_t root, int depth,
              const float o[3], const float d[3],
              const uint32_t* rcp_tab, int log2n,
              uint32_t* vox, uint8_t* face, float* t, oc_counters* cnt);

/* n rays; o_stride = 3 (per-ray origins) or 0 (shared origin).  npush16 (optional) gets the
 * per-ray PUSH count saturated to 65535; level_hist (optional, depth+1 entries of uint64)
 * accumulates PUSH evaluations by level (1..depth).  nthreads > 1 splits into chunks. */
void oc_trace_rays(const uint32_t* nodes, uint32_t root, int depth,
                   const float* o3, int o_stride, const float* d3, size_t n,
                   const uint32_t* rcp_tab, int log2n,
                   uint32_t* vox, uint8_t* face, float* t,
                   uint16_t* npush16, oc_counters* total, int nthreads);

/* ---- och::octree: plain pointer octree over a node pool (och_octree.h:10-69, och_octree.cpp:14-320) -------- */

typedef struct oc_octree
{
	int       depth;
	uint32_t  cap;
	uint32_t* nodes;      /* cap * 8; row 0 is the root, child values are raw row numbers, 0 = empty */
	uint32_t  head;       /* free list head (och_octree.h:27) */
	int       node_cnt;   /* och_octree.h:28 */
	int       failed;     /* set instead of printf + exit(0) on pool exhaustion (och_octree.cpp:50-54) */
} oc_octree;

oc_octree* oc_octree_create(int depth, uint32_t table_capacity);
void       oc_octree_destroy(oc_octree* t);
void       oc_octree_set(oc_octree* t, int16_t x, int16_t y, int16_t z, uint32_t vx);   /* och_octree.cpp:74-91 */
void       oc_octree_unset(oc_octree* t, int16_t x, int16_t y, int16_t z);              /* :93-139 */
uint32_t   oc_octree_at(const oc_octree* t, int16_t x, int16_t y, int16_t z);           /* :141-160 */
/* ops: n x (x, y, z, v, kind) int32, kind 0 = set, 1 = unset */
void       oc_octree_apply(oc_octree* t, const int32_t* ops, size_t n);
/* och_octree.cpp:167-320: same walk as oc_trace from pool row 0 with raw child ids; MISS reports t = 0.0F */
void oc_octree_trace(const uint32_t* pool, int depth, const float o[3], const float d[3],
                     const uint32_t* rcp_tab, int log2n, uint32_t* vox, uint8_t* face, float* t, oc_counters* cnt);
void oc_octree_trace_rays(const uint32_t* pool, int depth,
                   const float* o3, int o_stride, const float* d3, size_t n,
                   const uint32_t* rcp_tab, int log2n,
                   uint32_t* vox, uint8_t* face, float* t,
                   uint16_t* npush16, oc_counters* total, int nthreads);

/* ---- camera rays (test_och_h_octree.cpp:87-138) ------------------------------------------- */

/* rot[9] = t_x_fx, t_x_fy, t_x_fz, t_y_fx, ... t_z_fz (:107-115); fov_factor = 1/tanf(1.25/2) (:97) */
void oc_camera_coeffs(float yaw, float pitch, float rot[9], float* fov_factor);
/* rays for rows [y0, y1) of a W x H frame, row-major, 3 floats each (:119-137) */
void oc_gen_rays(const float rot[9], float fov_factor, int W, int H, int y0, int y1, float* d3);

/* ---- noise + terrain fixture (och_noise.h:73-366; test_och_h_octree.cpp:561-598, 651-695, 767-787) */

float oc_simplex2(float frequency, float x, float y);
float oc_simplex3(float frequency, float x, float y, float z);
void  oc_simplex2_many(float frequency, const float* xy, size_t n, float* out);
void  oc_simplex3_many(float frequency, const float* xyz, size_t n, float* out);

/* heights[y*dim+x] (:561-566, :587-592), simplex_n(0.5) heightmap of the live code */
void oc_heightmap(int depth, uint16_t* heights);

/* initialize_h_octree (:767-787): create_volume + surface layers + (optionally) tunnels.
 * grass[y*dim+x] in {0,1} replaces the reference's unseeded std::rand() > RAND_MAX/2 (:780).
 * Straight restatement -- O(volume), meant for depth <= 10. */
void oc_initialize_terrain(oc_tree* t, const uint16_t* heights, const uint8_t* grass, int tunnels);

#ifdef __cplusplus
}
#endif

#endif
