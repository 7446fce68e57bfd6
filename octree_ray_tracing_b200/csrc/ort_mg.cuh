// ort_mg.cuh -- multi-GPU side of libort_b200.so: one communicator per context (= per GPU, one process each), used for
// exactly the two exchanges the path has (BASELINE.json north_star): broadcasting the DAG / its edit deltas from the
// rank that owns the host table, and gathering the finished tile strips of a frame on the rank that consumes it
// (replaces the single-consumer frame of test_och_h_octree.cpp:437-457).  The trace itself never communicates.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy the process already holds -- e.g. torch's -- or the
// system one), so single-GPU users need no NCCL at all.  Included at the end of ort_device.cu.
//
// Gather data path, per frame and rank r (all on the device, nothing touches the host):
//   trace stream : ort_trace_frame_async -> strip slot s = [voxel u32 | t f32 | face u8] of r's cyclic tiles
//   comm stream  : waits for the trace (event); ONE ncclSend of the whole slot to the consumer rank; the consumer posts
//                  world-1 ncclRecv into its staging ring in the same group, then three unpack kernels move every rank's
//                  tiles to their final rows of the caller's frame buffers (its own strip straight from its trace slot).
//   The slots form a ring (kSlots): the trace of frame k+1 runs while frame k is on the wire; a slot is reused only after
//   its send (and, on the consumer, its unpack) has completed -- ordered by events, no host synchronisation.
#pragma once

#include <dlfcn.h>

namespace {

// ---- the few NCCL declarations this file needs (ABI-stable since NCCL 2.x) ------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                       // 0 = ncclSuccess
enum { ort_ncclUint8 = 1, ort_ncclUint32 = 3 };

struct NcclApi
{
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	const char*  (*GetErrorString)(ncclResult_t) = nullptr;
	ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi* nccl_api(std::string* why)
{
	static NcclApi api;
	static bool tried = false;
	static std::string err;
	if (!tried)
	{
		tried = true;
		const char* names[] = { "libnccl.so.2", "libnccl.so" };
		for (const char* n : names)
			if ((api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
		if (!api.handle)
			err = std::string("cannot load libnccl.so.2: ") + dlerror();
		else
		{
#define ORT_NCCL_SYM(field, sym) \
			if (err.empty() && !(*reinterpret_cast<void**>(&api.field) = dlsym(api.handle, sym))) err = std::string("libnccl lacks ") + sym
			ORT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
			ORT_NCCL_SYM(CommInitRank, "ncclCommInitRank");
			ORT_NCCL_SYM(CommDestroy, "ncclCommDestroy");
			ORT_NCCL_SYM(GroupStart, "ncclGroupStart");
			ORT_NCCL_SYM(GroupEnd, "ncclGroupEnd");
			ORT_NCCL_SYM(Send, "ncclSend");
			ORT_NCCL_SYM(Recv, "ncclRecv");
			ORT_NCCL_SYM(Broadcast, "ncclBroadcast");
			ORT_NCCL_SYM(GetErrorString, "ncclGetErrorString");
			ORT_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef ORT_NCCL_SYM
		}
	}
	if (!err.empty()) { if (why) *why = err; return nullptr; }
	return &api;
}

constexpr int kSlots = 3;                        // frames in flight between trace and wire

}  // namespace

struct ort_mg
{
	ort_ctx* ctx = nullptr;
	NcclApi* api = nullptr;
	ncclComm_t comm = nullptr;
	int rank = 0, world = 1;
	cudaStream_t comm_stream = nullptr;
	cudaEvent_t  ev_traced[kSlots] = {};         // slot's strip has been traced (trace stream)
	cudaEvent_t  ev_free[kSlots] = {};           // slot's strip has left / has been unpacked (comm stream)
	bool         slot_used[kSlots] = {};
	int          next_slot = 0;
	size_t pitch = 0;                                          // bytes of one strip block (the longest strip's)
	char*  d_strips = nullptr;  size_t strip_bytes = 0;        // sender: kSlots strip blocks of this rank
	char*  d_stage = nullptr;   size_t stage_bytes = 0;        // consumer: kSlots x world strip blocks, rank order (its own among them)
	uint32_t* d_update = nullptr; size_t update_words = 0;     // broadcast payload of ort_mg_broadcast_update
	uint64_t frames = 0;
	double   wire_bytes = 0;                                   // bytes this rank sent + received for gathers
};

#define ORT_NCCL(m, call)                                                                                          \
	do {                                                                                                           \
		ncclResult_t r_ = (call);                                                                                  \
		if (r_ != 0)                                                                                               \
			return ort_fail((m)->ctx, ORT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, (m)->api->GetErrorString(r_), __FILE__, __LINE__); \
	} while (0)

namespace {

// rows of rank's cyclic strip: tiles rank, rank + world, ... of ceil(H / tile_rows); the frame's last tile may be short
int mg_strip_rows(int rank, int world, int H, int tile_rows)
{
	const int n_tiles = (H + tile_rows - 1) / tile_rows;
	if (rank >= n_tiles) return 0;
	const int mine = (n_tiles - 1 - rank) / world + 1;
	int rows = mine * tile_rows;
	const int last_tile = rank + (mine - 1) * world;
	if (last_tile == n_tiles - 1) rows -= n_tiles * tile_rows - H;        // the short last tile is mine
	return rows;
}

// a strip block of n pixels: voxel u32 [n] | t f32 [n] | face u8 [n]  (n is a multiple of 4: W % 4 == 0)
inline size_t mg_block_bytes(size_t n) { return n * 9; }

}  // namespace

extern "C" {

int ort_mg_unique_id(void* id128)
{
	enter(nullptr);
	std::string why;
	NcclApi* api = nccl_api(&why);
	if (!api) return ort_fail(nullptr, ORT_ERR_NOT_ATTACHED, "ort_mg_unique_id: %s", why.c_str());
	if (!id128) return ort_fail(nullptr, ORT_ERR_INVALID, "ort_mg_unique_id: null buffer");
	ncclUniqueId id;
	const ncclResult_t r = api->GetUniqueId(&id);
	if (r != 0) return ort_fail(nullptr, ORT_ERR_CUDA, "ncclGetUniqueId failed: %s", api->GetErrorString(r));
	std::memcpy(id128, &id, sizeof id);
	return ORT_OK;
}

int ort_mg_create(ort_mg** out, ort_ctx* ctx, int rank, int world, const void* id128)
{
	enter(ctx);
	if (!out || !ctx || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128))
		return ort_fail(ctx, ORT_ERR_INVALID, "ort_mg_create: bad arguments");
	*out = nullptr;
	ort_mg* m = new (std::nothrow) ort_mg;
	if (!m) return ort_fail(ctx, ORT_ERR_INVALID, "ort_mg_create: out of host memory");
	m->ctx = ctx; m->rank = rank; m->world = world;
	DeviceGuard g(ctx->device);
	const int rc = [&]() -> int {
		ORT_CUDA(ctx, cudaStreamCreateWithFlags(&m->comm_stream, cudaStreamNonBlocking));
		for (int i = 0; i < kSlots; ++i)
		{
			ORT_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_traced[i], cudaEventDisableTiming));
			ORT_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_free[i], cudaEventDisableTiming));
		}
		if (world > 1)
		{
			std::string why;
			m->api = nccl_api(&why);
			if (!m->api) return ort_fail(ctx, ORT_ERR_NOT_ATTACHED, "ort_mg_create: %s", why.c_str());
			ncclUniqueId id;
			std::memcpy(&id, id128, sizeof id);
			ORT_NCCL(m, m->api->CommInitRank(&m->comm, world, id, rank));
		}
		return ORT_OK;
	}();
	if (rc != ORT_OK)
	{
		const std::string keep = g_last_error;
		ort_mg_destroy(m);
		g_last_error = keep;
		ctx->last_error = keep;
		return rc;
	}
	*out = m;
	return ORT_OK;
}

int ort_mg_destroy(ort_mg* m)
{
	if (!m) return ORT_OK;
	DeviceGuard g(m->ctx->device);
	if (m->comm_stream) cudaStreamSynchronize(m->comm_stream);
	if (m->comm && m->api) m->api->CommDestroy(m->comm);
	cudaFree(m->d_strips);
	cudaFree(m->d_stage);
	cudaFree(m->d_update);
	for (int i = 0; i < kSlots; ++i)
	{
		if (m->ev_traced[i]) cudaEventDestroy(m->ev_traced[i]);
		if (m->ev_free[i]) cudaEventDestroy(m->ev_free[i]);
	}
	if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
	delete m;
	return ORT_OK;
}

int ort_mg_rank(const ort_mg* m) { return m ? m->rank : -1; }
int ort_mg_world(const ort_mg* m) { return m ? m->world : 0; }
void* ort_mg_stream(ort_mg* m) { return m ? m->comm_stream : nullptr; }

int ort_mg_nccl_version(void)
{
	NcclApi* api = nccl_api(nullptr);
	int v = 0;
	if (!api || api->GetVersion(&v) != 0) return 0;
	return v;
}

int ort_mg_strip_rows(int rank, int world, int H, int tile_rows)
{
	if (rank < 0 || world < 1 || rank >= world || H < 0 || tile_rows <= 0) return -1;
	return mg_strip_rows(rank, world, H, tile_rows);
}

// Ship one device update from `src` to every rank and apply it to every rank's context: a full flatten (is_full) or a
// delta (ids + rows).  The arguments are read on `src` only (host pointers); the other ranks pass whatever.
int ort_mg_broadcast_update(ort_mg* m, const uint32_t* ids, const uint32_t* nodes8, size_t n, uint32_t root, int is_full, int src)
{
	if (!m) return ort_fail(nullptr, ORT_ERR_INVALID, "ort_mg_broadcast_update: null communicator");
	ort_ctx* c = m->ctx;
	enter(c);
	if (src < 0 || src >= m->world || (m->rank == src && n && (!nodes8 || (!is_full && !ids))))
		return ort_fail(c, ORT_ERR_INVALID, "ort_mg_broadcast_update: bad arguments");
	if (m->world == 1)
		return is_full ? ort_upload_full(c, nodes8, n, root) : ort_upload_delta(c, ids, nodes8, n, root);
	DeviceGuard g(c->device);
	// header: n, root, is_full (through the same payload buffer: 4 words)
	auto ensure = [&](size_t words) -> int {
		if (words <= m->update_words) return ORT_OK;
		ORT_CUDA(c, cudaStreamSynchronize(m->comm_stream));
		cudaFree(m->d_update); m->d_update = nullptr; m->update_words = 0;
		ORT_CUDA(c, cudaMalloc(&m->d_update, words * 4));
		m->update_words = words;
		return ORT_OK;
	};
	int rc = ensure(4);
	if (rc != ORT_OK) return rc;
	uint32_t hdr[4] = { static_cast<uint32_t>(n), static_cast<uint32_t>(n >> 32), root, static_cast<uint32_t>(is_full != 0) };
	if (m->rank == src)
		ORT_CUDA(c, cudaMemcpyAsync(m->d_update, hdr, sizeof hdr, cudaMemcpyHostToDevice, m->comm_stream));
	ORT_NCCL(m, m->api->Broadcast(m->d_update, m->d_update, 4, ort_ncclUint32, src, m->comm, m->comm_stream));
	ORT_CUDA(c, cudaMemcpyAsync(hdr, m->d_update, sizeof hdr, cudaMemcpyDeviceToHost, m->comm_stream));
	ORT_CUDA(c, cudaStreamSynchronize(m->comm_stream));
	const size_t nn = static_cast<size_t>(hdr[0]) | (static_cast<size_t>(hdr[1]) << 32);
	const uint32_t rt = hdr[2];
	const bool full = hdr[3] != 0;
	// payload: [ids (delta only) padded to 4 words | rows]
	const size_t id_words = full ? 0 : (nn + 3) / 4 * 4;
	const size_t words = id_words + nn * 8;
	if (words)
	{
		rc = ensure(words);
		if (rc != ORT_OK) return rc;
		if (m->rank == src)
		{
			if (!full) ORT_CUDA(c, cudaMemcpyAsync(m->d_update, ids, nn * 4, cudaMemcpyHostToDevice, m->comm_stream));
			ORT_CUDA(c, cudaMemcpyAsync(m->d_update + id_words, nodes8, nn * 32, cudaMemcpyHostToDevice, m->comm_stream));
		}
		ORT_NCCL(m, m->api->Broadcast(m->d_update, m->d_update, words, ort_ncclUint32, src, m->comm, m->comm_stream));
		ORT_CUDA(c, cudaStreamSynchronize(m->comm_stream));       // the context reads the payload on ITS stream
	}
	return full ? ort_upload_full(c, m->d_update, nn, rt) : ort_upload_delta(c, nn ? m->d_update : nullptr, nn ? m->d_update + id_words : nullptr, nn, rt);
}

// Trace this rank's cyclic strips of a W x H frame and gather the frame on rank `dst`: voxel / face / t are device
// buffers of W * H entries on dst (ignored elsewhere).  Collective: every rank of the communicator calls it for every
// frame, with the same camera and geometry.  Enqueue only; ort_mg_sync() (or the comm stream) tells when the frame is
// complete.  W must be a multiple of 4.
int ort_mg_trace_frame_gather(ort_mg* m, const float pos[3], const float rot[9], float fov_factor, int W, int H, int tile_rows, int dst,
                              uint32_t* voxel, uint8_t* face, float* t)
{
	if (!m) return ort_fail(nullptr, ORT_ERR_INVALID, "ort_mg_trace_frame_gather: null communicator");
	ort_ctx* c = m->ctx;
	enter(c);
	if (!pos || !rot || W <= 0 || H <= 0 || (W & 3) || tile_rows <= 0 || dst < 0 || dst >= m->world || (m->rank == dst && (!voxel || !face || !t)))
		return ort_fail(c, ORT_ERR_INVALID, "ort_mg_trace_frame_gather: bad arguments (W must be a multiple of 4)");
	DeviceGuard g(c->device);
	const int world = m->world, rank = m->rank;
	const int my_rows = mg_strip_rows(rank, world, H, tile_rows);
	int max_rows = 0;
	for (int r = 0; r < world; ++r) max_rows = std::max(max_rows, mg_strip_rows(r, world, H, tile_rows));
	const size_t max_n = static_cast<size_t>(max_rows) * W;                // sections of a block: voxel | t | face, max_n entries each
	const size_t pitch = align_up(mg_block_bytes(max_n), 256);
	const bool consumer = rank == dst;

	// a rank that sends keeps kSlots blocks; the consumer keeps kSlots x world (one per rank, its own among them)
	const size_t need = pitch * kSlots * (consumer ? world : 1);
	char*& pool = consumer ? m->d_stage : m->d_strips;
	size_t& pool_bytes = consumer ? m->stage_bytes : m->strip_bytes;
	if (need > pool_bytes || pitch != m->pitch)
	{
		ORT_CUDA(c, cudaDeviceSynchronize());
		if (pitch != m->pitch)
		{
			cudaFree(m->d_strips); m->d_strips = nullptr; m->strip_bytes = 0;
			cudaFree(m->d_stage); m->d_stage = nullptr; m->stage_bytes = 0;
			m->pitch = pitch;
		}
		cudaFree(pool); pool = nullptr; pool_bytes = 0;
		ORT_CUDA(c, cudaMalloc(&pool, need));
		pool_bytes = need;
		for (int i = 0; i < kSlots; ++i) m->slot_used[i] = false;
	}

	const int slot = m->next_slot;
	m->next_slot = (m->next_slot + 1) % kSlots;
	char* blocks = consumer ? m->d_stage + static_cast<size_t>(slot) * world * pitch : nullptr;      // consumer: this frame's blocks, rank order
	char* sb = consumer ? blocks + static_cast<size_t>(rank) * pitch : m->d_strips + static_cast<size_t>(slot) * pitch;
	uint32_t* sv = reinterpret_cast<uint32_t*>(sb);
	float*    st = reinterpret_cast<float*>(sb + max_n * 4);
	uint8_t*  sf = reinterpret_cast<uint8_t*>(sb + max_n * 8);

	// trace into the slot once its previous contents have left (a sender's block) / have been unpacked (the consumer's)
	cudaStream_t ts = c->stream;
	if (m->slot_used[slot]) ORT_CUDA(c, cudaStreamWaitEvent(ts, m->ev_free[slot], 0));
	if (my_rows)
	{
		const int rc = ort_trace_frame_async(c, pos, rot, fov_factor, W, H, rank * tile_rows, my_rows, tile_rows, world, sv, sf, st, nullptr);
		if (rc != ORT_OK) return rc;
	}
	ORT_CUDA(c, cudaEventRecord(m->ev_traced[slot], ts));
	ORT_CUDA(c, cudaStreamWaitEvent(m->comm_stream, m->ev_traced[slot], 0));

	// the wire: one message per (rank -> dst) pair -- the whole block, sections at fixed offsets
	if (world > 1)
	{
		ORT_NCCL(m, m->api->GroupStart());
		if (!consumer)
		{
			ORT_NCCL(m, m->api->Send(sb, mg_block_bytes(max_n), ort_ncclUint8, dst, m->comm, m->comm_stream));
			m->wire_bytes += static_cast<double>(mg_block_bytes(max_n));
		}
		else
			for (int r = 0; r < world; ++r)
			{
				if (r == dst) continue;
				ORT_NCCL(m, m->api->Recv(blocks + static_cast<size_t>(r) * pitch, mg_block_bytes(max_n), ort_ncclUint8, r, m->comm, m->comm_stream));
				m->wire_bytes += static_cast<double>(mg_block_bytes(max_n));
			}
		ORT_NCCL(m, m->api->GroupEnd());
	}

	// consumer: every rank's tiles -> their rows of the frame, one launch
	if (consumer)
	{
		const ort::StripMap map{ W, H, tile_rows, world, max_n, pitch };
		const dim3 grid(static_cast<unsigned>((max_n / 4 + 255) / 256), static_cast<unsigned>(world));
		ort::unpack_strips_kernel<<<grid, 256, 0, m->comm_stream>>>(reinterpret_cast<uint4*>(voxel), reinterpret_cast<uint4*>(t), reinterpret_cast<uint32_t*>(face), blocks, map);
		++c->launches;
		ORT_CUDA(c, cudaGetLastError());
	}
	ORT_CUDA(c, cudaEventRecord(m->ev_free[slot], m->comm_stream));
	m->slot_used[slot] = true;
	++m->frames;
	return ORT_OK;
}

// all frames queued so far are complete on their consumers (and this rank's strips have left)
int ort_mg_sync(ort_mg* m)
{
	if (!m) return ort_fail(nullptr, ORT_ERR_INVALID, "ort_mg_sync: null communicator");
	enter(m->ctx);
	DeviceGuard g(m->ctx->device);
	ORT_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
	ORT_CUDA(m->ctx, cudaStreamSynchronize(m->comm_stream));
	return ORT_OK;
}

double ort_mg_wire_bytes(const ort_mg* m) { return m ? m->wire_bytes : 0.0; }

}  // extern "C"
