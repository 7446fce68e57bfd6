// ort_mg.cuh -- multi-GPU side of libort_b200.so: one communicator per context (= per GPU, one process each), used for
// exactly the two exchanges the path has (BASELINE.json north_star): broadcasting the DAG / its edit deltas from the
// rank that owns the host table, and gathering the finished tile strips of a frame on the rank that consumes it
// (replaces the single-consumer frame of test_och_h_octree.cpp:437-457).  The trace itself never communicates.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy the process already holds -- e.g. torch's -- or the
// system one), so single-GPU users need no NCCL at all.  Included at the end of ort_device.cu.
//
// Gather data path, per frame and rank r (all on the device, nothing touches the host):
//   trace stream : ort_trace_frame_async -> a strip block [voxel u32 | t f32 | face u8] of r's cyclic tiles, taken from a
//                  ring of blocks (on the consumer: the block of its own rank inside the frame's staging area)
//   comm stream  : every `group` frames (ort_mg_set_group, default 1) ONE NCCL group moves all pending strips -- a sender
//                  posts one ncclSend per frame, a consumer world-1 ncclRecv per frame -- and one unpack kernel per
//                  consumed frame writes every rank's tiles to their final rows of the caller's frame buffers.
//   The rings hold 2 x group blocks: the traces of the next group run while this group is on the wire; a block is reused
//   only after its send (on the consumer: its unpack) has completed -- ordered by events, no host synchronisation.
//   group = 1 gives the lowest latency per frame; a larger group turns many small point-to-point messages into one
//   all-to-all-shaped exchange that NCCL spreads over all peers and channels at once.
//
// Two transports for the strips (ort_mg_set_transport; the choice is collective):
//   0  NCCL point-to-point: ncclSend / ncclRecv as described above.  NCCL's copy kernels share the SMs with the trace
//      kernels of the next frames -- which are bound by instruction issue, so the wire runs at a fraction of its idle
//      rate and the traces slow down (2 GPUs: 0.82 of the trace-only throughput, 8 GPUs: 0.80).
//   1  peer copies (default where CUDA IPC works): every rank maps every other rank's receive ring (cudaIpc*MemHandle,
//      exchanged once per geometry through NCCL) and a sender's strip block travels with ONE cudaMemcpyAsync on the copy
//      engines over NVLink straight into its place in the consumer's staging area -- no SM is involved.  NCCL carries only
//      the synchronisation: one 4-byte ncclAllReduce per wire operation after the copies.  It is a rendezvous of all
//      ranks' comm streams, which gives both orderings the rings need: a consumer unpacks after every sender's copies of
//      this operation have landed (they precede the sender's all-reduce in stream order), and a sender overwrites a
//      staging slot two operations later, after the consumer's unpack of the slot's previous frame (which precedes the
//      consumer's NEXT all-reduce, which the sender's own next all-reduce cannot complete without).
#pragma once

#include <dlfcn.h>

namespace {

// ---- the few NCCL declarations this file needs (ABI-stable since NCCL 2.x) ------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                       // 0 = ncclSuccess
enum { ort_ncclUint8 = 1, ort_ncclInt32 = 2, ort_ncclUint32 = 3 };
enum { ort_ncclSum = 0, ort_ncclMin = 3 };

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
	ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
			ORT_NCCL_SYM(AllReduce, "ncclAllReduce");
			ORT_NCCL_SYM(AllGather, "ncclAllGather");
			ORT_NCCL_SYM(GetErrorString, "ncclGetErrorString");
			ORT_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef ORT_NCCL_SYM
		}
	}
	if (!err.empty()) { if (why) *why = err; return nullptr; }
	return &api;
}

struct MgFrame                                     // a traced frame whose strips have not been put on the wire yet
{
	int W, H, tile_rows, dst;
	size_t max_n;
	char* sb;                                      // this rank's strip block
	char* blocks;                                  // consumer: the frame's world blocks (rank order), else null
	uint32_t* voxel; uint8_t* face; float* t;      // consumer: the caller's frame buffers
	int slot;                                      // ring slot to release (sender ring or consumer ring)
	int rslot;                                     // the frame's slot in its consumer's receive ring (the same number on every rank)
};

struct MgRing                                      // blocks handed out round robin, each guarded by an event
{
	char* base = nullptr;
	size_t slot_bytes = 0;
	int n = 0, next = 0;
	std::vector<int> op_of_slot;                   // the wire operation that releases the slot (index into ort_mg::ev_op), -1: never used
};

}  // namespace

struct ort_mg
{
	ort_ctx* ctx = nullptr;
	NcclApi* api = nullptr;
	ncclComm_t comm = nullptr;
	int rank = 0, world = 1;
	int group = 1;                                 // frames per wire operation
	cudaStream_t comm_stream = nullptr;
	// Strips are traced on a few streams in turn: a strip launch ends with the latency tail of its longest rays, and only
	// launches on different streams overlap that tail with the bulk of the next one (one stream: 28.7, four: 35.2 Grays/s
	// on two GPUs).  Every trace is forked from / joined to the context's stream, which stays the ordering timeline for
	// uploads.
	static constexpr int kMaxTraceStreams = 8;
	int          n_trace_streams = 4, next_trace_stream = 0;
	cudaStream_t trace_stream[kMaxTraceStreams] = {};
	cudaEvent_t  ev_traced[kMaxTraceStreams] = {};     // the stream's pending strips have been traced
	bool         trace_stream_dirty[kMaxTraceStreams] = {};
	cudaEvent_t  ev_fork = nullptr;
	// peer copies of one wire operation are spread over a few streams so that several copy engines work at once
	static constexpr int kCopyStreams = 4;
	cudaStream_t copy_stream[kCopyStreams] = {};
	cudaEvent_t  ev_copied[kCopyStreams] = {};
	cudaEvent_t  ev_ready = nullptr;                   // the comm stream has seen the traces of the pending frames
	// one completion event per wire operation, a ring of them: a block is reused after the operation that carried its
	// previous contents (and, on a consumer, unpacked them) -- at most two operations back, the rings hold 2 x group blocks
	static constexpr int kOpEvents = 8;
	cudaEvent_t  ev_op[kOpEvents] = {};
	int          next_op = 0;
	bool         forked = false;                       // the trace streams have seen the context's stream since the last wire operation
	size_t pitch = 0;                              // bytes of one strip block (the longest strip's)
	MgRing send_ring, recv_ring;                   // sender: blocks of this rank; consumer: world blocks per frame
	std::vector<MgFrame> pending;
	int  transport_pref = 1;                       // 1: peer copies over CUDA IPC if every rank can map every ring, 0: NCCL send / recv
	bool peer_copies = false;                      // what the current rings use (agreed by all ranks)
	std::vector<char*> peer_recv_base;             // [world] the other ranks' receive rings, mapped into this process
	std::vector<uint64_t> consumed;                // [world] frames consumed on each rank so far (the same count on every rank)
	int* d_flag = nullptr;                         // the 4 bytes of the all-reduce
	uint32_t* d_update = nullptr; size_t update_words = 0;     // broadcast payload of ort_mg_broadcast_update
	uint64_t frames = 0, wire_ops = 0;
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

void mg_ring_free(MgRing& r)
{
	cudaFree(r.base);
	r = MgRing{};
}

// (re)build a ring of n slots of slot_bytes; the caller has made sure nothing is in flight
int mg_ring_alloc(ort_mg* m, MgRing& r, int n, size_t slot_bytes)
{
	mg_ring_free(r);
	ORT_CUDA(m->ctx, cudaMalloc(&r.base, slot_bytes * n));
	r.slot_bytes = slot_bytes;
	r.n = n;
	r.op_of_slot.assign(n, -1);
	return ORT_OK;
}

void mg_unmap_peers(ort_mg* m)
{
	for (int r = 0; r < static_cast<int>(m->peer_recv_base.size()); ++r)
		if (m->peer_recv_base[r] && r != m->rank) cudaIpcCloseMemHandle(m->peer_recv_base[r]);
	m->peer_recv_base.assign(m->world, nullptr);
	m->peer_copies = false;
}

// Collective: every rank exports its receive ring and maps everybody else's.  All ranks end up with the same answer
// (peer_copies on or off): a rank that cannot export or map makes everybody fall back to NCCL send / recv.
int mg_map_peers(ort_mg* m)
{
	ort_ctx* c = m->ctx;
	const int world = m->world, rank = m->rank;
	mg_unmap_peers(m);
	if (world == 1 || !m->transport_pref) return ORT_OK;
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
	char* d_handles = nullptr;
	ORT_CUDA(c, cudaMalloc(&d_handles, 64 * static_cast<size_t>(world)));
	cudaIpcMemHandle_t mine;
	int ok = cudaIpcGetMemHandle(&mine, m->recv_ring.base) == cudaSuccess;
	if (!ok) { cudaGetLastError(); std::memset(&mine, 0, sizeof mine); }
	std::vector<cudaIpcMemHandle_t> all(world);
	const int rc = [&]() -> int {
		ORT_CUDA(c, cudaMemcpyAsync(d_handles + 64 * static_cast<size_t>(rank), &mine, 64, cudaMemcpyHostToDevice, m->comm_stream));
		ORT_NCCL(m, m->api->AllGather(d_handles + 64 * static_cast<size_t>(rank), d_handles, 64, ort_ncclUint8, m->comm, m->comm_stream));
		ORT_CUDA(c, cudaMemcpyAsync(all.data(), d_handles, 64 * static_cast<size_t>(world), cudaMemcpyDeviceToHost, m->comm_stream));
		ORT_CUDA(c, cudaStreamSynchronize(m->comm_stream));
		for (int r = 0; r < world && ok; ++r)
		{
			if (r == rank) { m->peer_recv_base[r] = m->recv_ring.base; continue; }
			void* p = nullptr;
			if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
			m->peer_recv_base[r] = static_cast<char*>(p);
		}
		// agree: peer copies only if every rank mapped every ring
		ORT_CUDA(c, cudaMemcpyAsync(m->d_flag, &ok, 4, cudaMemcpyHostToDevice, m->comm_stream));
		ORT_NCCL(m, m->api->AllReduce(m->d_flag, m->d_flag, 1, ort_ncclInt32, ort_ncclMin, m->comm, m->comm_stream));
		int all_ok = 0;
		ORT_CUDA(c, cudaMemcpyAsync(&all_ok, m->d_flag, 4, cudaMemcpyDeviceToHost, m->comm_stream));
		ORT_CUDA(c, cudaStreamSynchronize(m->comm_stream));
		if (all_ok) m->peer_copies = true;
		else
		{
			mg_unmap_peers(m);
			static bool said = false;
			if (!said && rank == 0) { said = true; std::fprintf(stderr, "ort_b200: note: CUDA IPC is not available between the ranks; the strip gather uses NCCL send / recv\n"); }
		}
		return ORT_OK;
	}();
	cudaFree(d_handles);
	return rc;
}

// put the pending frames on the wire: one NCCL group, then the consumers' unpack kernels, then the blocks are released
int mg_flush(ort_mg* m)
{
	ort_ctx* c = m->ctx;
	if (m->pending.empty()) return ORT_OK;
	for (int i = 0; i < ort_mg::kMaxTraceStreams; ++i)
		if (m->trace_stream_dirty[i])
		{
			ORT_CUDA(c, cudaEventRecord(m->ev_traced[i], m->trace_stream[i]));
			ORT_CUDA(c, cudaStreamWaitEvent(m->comm_stream, m->ev_traced[i], 0));
			m->trace_stream_dirty[i] = false;
		}
	m->forked = false;
	const int world = m->world, rank = m->rank;
	if (world > 1 && m->peer_copies)
	{
		// transport 1: the copy engines move the blocks, one 4-byte all-reduce orders everything (see the file header).
		// The copies of this operation fan out over a few streams (several engines, several destinations at once) and
		// join the comm stream again before the all-reduce.
		int n_copies = 0;
		for (const MgFrame& f : m->pending) n_copies += f.blocks ? 0 : 1;
		const int n_cs = std::min<int>(ort_mg::kCopyStreams, n_copies);
		if (n_cs > 1)
		{
			ORT_CUDA(c, cudaEventRecord(m->ev_ready, m->comm_stream));
			for (int i = 0; i < n_cs; ++i) ORT_CUDA(c, cudaStreamWaitEvent(m->copy_stream[i], m->ev_ready, 0));
		}
		int k = 0;
		for (const MgFrame& f : m->pending)
		{
			if (f.blocks) continue;
			const size_t bytes = mg_block_bytes(f.max_n);
			char* dst = m->peer_recv_base[f.dst] + static_cast<size_t>(f.rslot) * m->recv_ring.slot_bytes + static_cast<size_t>(rank) * m->pitch;
			ORT_CUDA(c, cudaMemcpyAsync(dst, f.sb, bytes, cudaMemcpyDeviceToDevice, n_cs > 1 ? m->copy_stream[k++ % n_cs] : m->comm_stream));
			m->wire_bytes += static_cast<double>(bytes);
		}
		if (n_cs > 1)
			for (int i = 0; i < n_cs; ++i)
			{
				ORT_CUDA(c, cudaEventRecord(m->ev_copied[i], m->copy_stream[i]));
				ORT_CUDA(c, cudaStreamWaitEvent(m->comm_stream, m->ev_copied[i], 0));
			}
		for (const MgFrame& f : m->pending)
			if (f.blocks) m->wire_bytes += static_cast<double>(mg_block_bytes(f.max_n)) * (world - 1);
		ORT_NCCL(m, m->api->AllReduce(m->d_flag, m->d_flag, 1, ort_ncclInt32, ort_ncclSum, m->comm, m->comm_stream));
		++m->wire_ops;
	}
	else if (world > 1)
	{
		ORT_NCCL(m, m->api->GroupStart());
		for (const MgFrame& f : m->pending)
		{
			const size_t bytes = mg_block_bytes(f.max_n);
			if (!f.blocks)
			{
				ORT_NCCL(m, m->api->Send(f.sb, bytes, ort_ncclUint8, f.dst, m->comm, m->comm_stream));
				m->wire_bytes += static_cast<double>(bytes);
			}
			else
				for (int r = 0; r < world; ++r)
				{
					if (r == rank) continue;
					ORT_NCCL(m, m->api->Recv(f.blocks + static_cast<size_t>(r) * m->pitch, bytes, ort_ncclUint8, r, m->comm, m->comm_stream));
					m->wire_bytes += static_cast<double>(bytes);
				}
		}
		ORT_NCCL(m, m->api->GroupEnd());
		++m->wire_ops;
	}
	const int op = m->next_op;
	m->next_op = (m->next_op + 1) % ort_mg::kOpEvents;
	for (const MgFrame& f : m->pending)
	{
		MgRing& ring = f.blocks ? m->recv_ring : m->send_ring;
		if (f.blocks)
		{
			// every rank's tiles -> their rows of the frame, one launch
			const ort::StripMap map{ f.W, f.H, f.tile_rows, world, f.max_n, m->pitch };
			const dim3 grid(static_cast<unsigned>((f.max_n / 4 + 255) / 256), static_cast<unsigned>(world));
			ort::unpack_strips_kernel<<<grid, 256, 0, m->comm_stream>>>(reinterpret_cast<uint4*>(f.voxel), reinterpret_cast<uint4*>(f.t), reinterpret_cast<uint32_t*>(f.face), f.blocks, map);
			++c->launches;
		}
		ring.op_of_slot[f.slot] = op;
	}
	ORT_CUDA(c, cudaEventRecord(m->ev_op[op], m->comm_stream));        // releases every block of this operation
	ORT_CUDA(c, cudaGetLastError());
	m->pending.clear();
	return ORT_OK;
}

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
		ORT_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
		ORT_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_ready, cudaEventDisableTiming));
		for (int i = 0; i < ort_mg::kOpEvents; ++i) ORT_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_op[i], cudaEventDisableTiming));
		for (int i = 0; i < ort_mg::kCopyStreams; ++i)
		{
			ORT_CUDA(ctx, cudaStreamCreateWithFlags(&m->copy_stream[i], cudaStreamNonBlocking));
			ORT_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_copied[i], cudaEventDisableTiming));
		}
		ORT_CUDA(ctx, cudaMalloc(&m->d_flag, 4));
		ORT_CUDA(ctx, cudaMemset(m->d_flag, 0, 4));
		m->peer_recv_base.assign(world, nullptr);
		m->consumed.assign(world, 0);
		if (const char* e = std::getenv("ORT_MG_TRANSPORT")) m->transport_pref = std::atoi(e) != 0;
		m->n_trace_streams = world > 2 ? 8 : 4;      // (4 GPUs: 8 streams 84.2, 4 streams 81.2 Grays/s with the gather)
		for (int i = 0; i < ort_mg::kMaxTraceStreams; ++i)
		{
			ORT_CUDA(ctx, cudaStreamCreateWithFlags(&m->trace_stream[i], cudaStreamNonBlocking));
			ORT_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_traced[i], cudaEventDisableTiming));
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
	const bool had_peers = m->peer_copies;
	mg_unmap_peers(m);
	if (had_peers && m->comm && m->api && m->d_flag)
	{
		// (collective, like ncclCommDestroy: a receive ring is freed only after every rank has dropped its mapping of it)
		if (m->api->AllReduce(m->d_flag, m->d_flag, 1, ort_ncclInt32, ort_ncclSum, m->comm, m->comm_stream) == 0)
			cudaStreamSynchronize(m->comm_stream);
	}
	if (m->comm && m->api) m->api->CommDestroy(m->comm);
	mg_unmap_peers(m);
	mg_ring_free(m->send_ring);
	mg_ring_free(m->recv_ring);
	cudaFree(m->d_update);
	cudaFree(m->d_flag);
	for (int i = 0; i < ort_mg::kMaxTraceStreams; ++i)
	{
		if (m->trace_stream[i]) { cudaStreamSynchronize(m->trace_stream[i]); cudaStreamDestroy(m->trace_stream[i]); }
		if (m->ev_traced[i]) cudaEventDestroy(m->ev_traced[i]);
	}
	if (m->ev_fork) cudaEventDestroy(m->ev_fork);
	if (m->ev_ready) cudaEventDestroy(m->ev_ready);
	for (int i = 0; i < ort_mg::kOpEvents; ++i) if (m->ev_op[i]) cudaEventDestroy(m->ev_op[i]);
	for (int i = 0; i < ort_mg::kCopyStreams; ++i)
	{
		if (m->copy_stream[i]) { cudaStreamSynchronize(m->copy_stream[i]); cudaStreamDestroy(m->copy_stream[i]); }
		if (m->ev_copied[i]) cudaEventDestroy(m->ev_copied[i]);
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
	DeviceGuard g(c->device);
	{
		// the update changes the array every strip trace in flight reads: they finish first
		const int rc = mg_flush(m);
		if (rc != ORT_OK) return rc;
		for (int i = 0; i < ort_mg::kMaxTraceStreams; ++i) ORT_CUDA(c, cudaStreamSynchronize(m->trace_stream[i]));
	}
	if (m->world == 1)
		return is_full ? ort_upload_full(c, nodes8, n, root) : ort_upload_delta(c, ids, nodes8, n, root);
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

	// Rings of 2 x group slots (at least 3): the send ring holds one strip block per slot, the receive ring world blocks
	// per slot (one per rank, the consumer's own among them).  Every rank keeps both -- with peer copies a rank's
	// receive ring is written by the others -- and every rank hands out the receive slots of every consumer in the same
	// order (consumed[dst]), so a sender knows where its block goes.  Geometry changes are collective: all ranks see
	// the same frames.
	const int want_slots = std::max(3, 2 * m->group);
	if (pitch != m->pitch || m->send_ring.n != want_slots || m->recv_ring.n != want_slots)
	{
		int rc = mg_flush(m);
		if (rc != ORT_OK) return rc;
		rc = ort_mg_sync(m);                                               // nothing in flight on any stream of this rank ...
		if (rc != ORT_OK) return rc;
		enter(c);
		if (world > 1)
		{
			ORT_NCCL(m, m->api->AllReduce(m->d_flag, m->d_flag, 1, ort_ncclInt32, ort_ncclSum, m->comm, m->comm_stream));   // ... nor on any other rank's
			ORT_CUDA(c, cudaStreamSynchronize(m->comm_stream));
		}
		mg_unmap_peers(m);
		if (world > 1)
		{
			// a ring is freed only after every rank has dropped its mapping of it
			ORT_NCCL(m, m->api->AllReduce(m->d_flag, m->d_flag, 1, ort_ncclInt32, ort_ncclSum, m->comm, m->comm_stream));
			ORT_CUDA(c, cudaStreamSynchronize(m->comm_stream));
		}
		m->pitch = pitch;
		rc = mg_ring_alloc(m, m->send_ring, want_slots, pitch);
		if (rc == ORT_OK) rc = mg_ring_alloc(m, m->recv_ring, want_slots, pitch * world);
		if (rc == ORT_OK) rc = mg_map_peers(m);
		if (rc != ORT_OK) return rc;
		std::fill(m->consumed.begin(), m->consumed.end(), 0);
	}

	const int rslot = static_cast<int>(m->consumed[dst] % static_cast<uint64_t>(m->recv_ring.n));
	++m->consumed[dst];
	MgRing& ring = consumer ? m->recv_ring : m->send_ring;
	const size_t slot_bytes = ring.slot_bytes;
	int slot = rslot;
	if (!consumer)
	{
		slot = ring.next;
		ring.next = (ring.next + 1) % ring.n;
	}
	char* blocks = consumer ? ring.base + static_cast<size_t>(slot) * slot_bytes : nullptr;          // consumer: this frame's blocks, rank order
	char* sb = consumer ? blocks + static_cast<size_t>(rank) * pitch : ring.base + static_cast<size_t>(slot) * slot_bytes;
	uint32_t* sv = reinterpret_cast<uint32_t*>(sb);
	float*    st = reinterpret_cast<float*>(sb + max_n * 4);
	uint8_t*  sf = reinterpret_cast<uint8_t*>(sb + max_n * 8);

	// trace into the block once its previous contents have left (a sender's block) / have been unpacked (a consumer's), on
	// the next trace stream; the context's stream is the timeline: the trace starts after what is queued there (an
	// upload)
	cudaStream_t user = c->stream;
	if (!m->forked)
	{
		// once per wire operation: the trace streams see what is queued on the context's stream (an upload)
		ORT_CUDA(c, cudaEventRecord(m->ev_fork, user));
		for (int i = 0; i < m->n_trace_streams; ++i) ORT_CUDA(c, cudaStreamWaitEvent(m->trace_stream[i], m->ev_fork, 0));
		m->forked = true;
	}
	const int tsi = m->next_trace_stream;
	m->next_trace_stream = (m->next_trace_stream + 1) % m->n_trace_streams;
	cudaStream_t ts = m->trace_stream[tsi];
	if (ring.op_of_slot[slot] >= 0) ORT_CUDA(c, cudaStreamWaitEvent(ts, m->ev_op[ring.op_of_slot[slot]], 0));
	if (my_rows)
	{
		c->stream = ts;
		const int rc = ort_trace_frame_async(c, pos, rot, fov_factor, W, H, rank * tile_rows, my_rows, tile_rows, world, sv, sf, st, nullptr);
		c->stream = user;
		if (rc != ORT_OK) return rc;
	}
	m->trace_stream_dirty[tsi] = true;
	// (no join back into the context's stream: the next fork would then wait for this trace and the strips would run one
	// after the other; ort_mg_sync() and ort_mg_broadcast_update() wait for the trace streams instead)
	m->pending.push_back(MgFrame{ W, H, tile_rows, dst, max_n, sb, blocks, voxel, face, t, slot, rslot });
	++m->frames;
	if (static_cast<int>(m->pending.size()) >= m->group)
		return mg_flush(m);
	return ORT_OK;
}

// A sequence of frames in one call (a render loop's step: several cameras, or the frames of a benchmark step).
int ort_mg_trace_frames_gather(ort_mg* m, const ort_mg_frame_job* jobs, int n_jobs)
{
	if (!m || n_jobs < 0 || (n_jobs && !jobs)) return ort_fail(m ? m->ctx : nullptr, ORT_ERR_INVALID, "ort_mg_trace_frames_gather: bad arguments");
	for (int i = 0; i < n_jobs; ++i)
	{
		const ort_mg_frame_job& j = jobs[i];
		const int rc = ort_mg_trace_frame_gather(m, j.pos, j.rot, j.fov_factor, j.W, j.H, j.tile_rows, j.dst, j.voxel, j.face, j.t);
		if (rc != ORT_OK) return rc;
	}
	return ORT_OK;
}

// Frames per wire operation (collective setting: the same on every rank; takes effect at the next frame).  1 (default):
// every frame's strips leave as soon as they are traced.  n > 1: the strips of n consecutive frames leave in ONE NCCL
// group -- fewer, larger, all-to-all-shaped exchanges; ort_mg_flush() / ort_mg_sync() send a partial group.
int ort_mg_set_group(ort_mg* m, int frames)
{
	if (!m || frames < 1 || frames > 256) return ort_fail(m ? m->ctx : nullptr, ORT_ERR_INVALID, "ort_mg_set_group: 1..256 frames");
	enter(m->ctx);
	DeviceGuard g(m->ctx->device);
	const int rc = mg_flush(m);
	m->group = frames;
	return rc;
}

// Streams the strips are traced on in turn (1..8; default 4, 8 from five ranks on): more of them hide the latency tails of
// small strip launches, too many dilute each launch's cache locality.
int ort_mg_set_trace_streams(ort_mg* m, int n)
{
	if (!m || n < 1 || n > ort_mg::kMaxTraceStreams) return ort_fail(m ? m->ctx : nullptr, ORT_ERR_INVALID, "ort_mg_set_trace_streams: 1..8 streams");
	enter(m->ctx);
	DeviceGuard g(m->ctx->device);
	const int rc = mg_flush(m);
	m->n_trace_streams = n;
	m->next_trace_stream = 0;
	return rc;
}

// Collective setting: how the strips travel.  1 (default): peer copies on the copy engines through CUDA IPC mappings of
// the receive rings (NCCL carries a 4-byte all-reduce per wire operation); falls back to 0 when some rank cannot map a
// ring.  0: NCCL ncclSend / ncclRecv.  Takes effect at the next frame (the rings are rebuilt).  ort_mg_transport()
// tells what is in use.
int ort_mg_set_transport(ort_mg* m, int transport)
{
	if (!m || (transport != 0 && transport != 1)) return ort_fail(m ? m->ctx : nullptr, ORT_ERR_INVALID, "ort_mg_set_transport: 0 (NCCL send / recv) or 1 (peer copies)");
	enter(m->ctx);
	DeviceGuard g(m->ctx->device);
	const int rc = mg_flush(m);
	m->transport_pref = transport;
	m->pitch = 0;                                  // forces the (collective) ring set-up at the next frame
	return rc;
}
int ort_mg_transport(const ort_mg* m) { return m ? (m->peer_copies ? 1 : 0) : -1; }

int ort_mg_flush(ort_mg* m)
{
	if (!m) return ort_fail(nullptr, ORT_ERR_INVALID, "ort_mg_flush: null communicator");
	enter(m->ctx);
	DeviceGuard g(m->ctx->device);
	return mg_flush(m);
}

// all frames queued so far are complete on their consumers (and this rank's strips have left)
int ort_mg_sync(ort_mg* m)
{
	if (!m) return ort_fail(nullptr, ORT_ERR_INVALID, "ort_mg_sync: null communicator");
	enter(m->ctx);
	DeviceGuard g(m->ctx->device);
	const int rc = mg_flush(m);
	if (rc != ORT_OK) return rc;
	ORT_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
	for (int i = 0; i < ort_mg::kMaxTraceStreams; ++i) ORT_CUDA(m->ctx, cudaStreamSynchronize(m->trace_stream[i]));
	ORT_CUDA(m->ctx, cudaStreamSynchronize(m->comm_stream));
	return ORT_OK;
}

double ort_mg_wire_bytes(const ort_mg* m) { return m ? m->wire_bytes : 0.0; }
uint64_t ort_mg_wire_ops(const ort_mg* m) { return m ? m->wire_ops : 0; }

}  // extern "C"
