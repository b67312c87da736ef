// libgcgpu: C ABI + CUDA kernels (sm_100a) of the GraphChainer alignment hot path.
// Interface and the reference seams each entry point replaces: include/gcgpu.h.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/gcgpu.h"
#include "gc_common.cuh"
#include "gc_k1.cuh"
#include "gc_k1s.cuh"
#include "gc_k2.cuh"
#include "gc_k3.cuh"
#include "gc_k3w.cuh"
#include "gc_seed.cuh"
#include "gc_post.cuh"
#include "gc_gam.cuh"
#include "gc_host_graph.h"
#include "gc_post_host.h"

#define GCGPU_VERSION 1
#define GC_K1_LONG_ITEM 96   // sequence length from which a K1 item gets a warp of its own

static thread_local std::string g_lastError;
static const bool g_trace = getenv("GCGPU_TRACE") != nullptr;
#define GC_TRACE_MS(name, count) do { if (g_trace) fprintf(stderr, "[gcgpu] %-28s n=%-8u %.3f ms\n", name, (unsigned)(count), ms); } while (0)
static int setError(int code, const std::string& msg) { g_lastError = msg; return code; }

#define CUDA_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return setError(_e == cudaErrorMemoryAllocation ? GCGPU_ERR_NOMEM : GCGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)

// growable device buffer
static const bool g_traceMem = getenv("GCGPU_TRACE_MEM") != nullptr;
struct DevBuf
{
	void* p = nullptr;
	size_t cap = 0;
	cudaError_t ensure(size_t bytes, int line = __builtin_LINE())
	{
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = bytes + std::min<size_t>(bytes / 4, (size_t)2 << 30) + 4096; // a quarter of headroom (batches differ by a few per cent), at most 2 GB
		if (g_traceMem && want >= ((size_t)1 << 30)) { size_t fr = 0, tot = 0; cudaMemGetInfo(&fr, &tot); fprintf(stderr, "[gcgpu] device buffer grows to %.2f GB (%.1f of %.1f GB free), libgcgpu source line %d\n", want / 1e9, fr / 1e9, tot / 1e9, line); }
		cudaError_t e = cudaMalloc(&p, want);
		if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; } // the failed attempt must not stay behind as the "last error" of the next launch check
		if (e == cudaSuccess) cap = want; else { cudaGetLastError(); p = nullptr; }
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T>
static cudaError_t uploadArray(const T* host, size_t count, T** dev)
{
	*dev = nullptr;
	if (count == 0 || host == nullptr) return cudaSuccess;
	cudaError_t e = cudaMalloc((void**)dev, count * sizeof(T));
	if (e != cudaSuccess) return e;
	return cudaMemcpy(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice);
}

struct GcResident;
struct gcgpu_ctx
{
	int device = 0;
	int numSMs = 148;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	cudaStream_t stream2 = nullptr; cudaEvent_t evFork = nullptr, evJoin = nullptr; // side stream for kernels that run beside the main one
	cudaEvent_t evSync = nullptr; // cudaEventBlockingSync: a host thread waiting for its batch sleeps instead of spinning on a core the other batches need
	gcgpu_params params;
	uint32_t numNodes = 0;
	// device copies of the graph
	uint8_t* d_nodeLength = nullptr; uint64_t* d_nodeSeq = nullptr;
	uint32_t* d_inStart = nullptr; uint32_t* d_inNbr = nullptr; uint32_t* d_outStart = nullptr; uint32_t* d_outNbr = nullptr;
	uint32_t* d_componentNumber = nullptr; uint8_t* d_linearizable = nullptr;
	GcNodeRec* d_nodeRec = nullptr; uint64_t* d_outKey = nullptr;
	GcViterbiTables* d_vt = nullptr;
	GcGraphView view;
	// MPC index (K2)
	uint32_t* d_compMap = nullptr; uint32_t* d_compIdx = nullptr; uint32_t* d_compStart = nullptr; uint32_t* d_topoIds = nullptr;
	uint32_t* d_pathsStart = nullptr; uint32_t* d_pathsK = nullptr; uint32_t* d_backStart = nullptr; uint32_t* d_backNode = nullptr; uint32_t* d_backK = nullptr;
	uint32_t* d_pathBase = nullptr; uint32_t totalPaths = 0, maxCompNodes = 0; // first global path id of every component (per-path structures of K2)
	bool haveMpc = false;
	GcMpcView mpc;
	DevBuf k2Work, k2Sort;
	DevBuf seqBuf, nwSeqBuf, descBuf, resBuf, arena, traceArena, compact, copyDesc, itemsBuf;
	DevBuf planes; // bit planes of seqBuf (gc_planes_kernel): the Eq masks of any 64 rows in eight loads
	uint64_t h2dBytes = 0, d2hBytes = 0; // bytes this ctx copied across PCIe (gcgpu_transfer_bytes)
	GcResident* resident = nullptr;      // state of the resident batch entry points (gcgpu_resident.inl)
	// minimizer index (S0)
	GcMzSlot* d_mzSlots = nullptr;
	GcMzView mz;
	bool haveMz = false;
	DevBuf seedBuf, seedMatches;
	uint64_t denseMatches = 0;
	uint64_t seqResident = ~0ULL; // bytes of the K1 sequence buffer currently on the device
	uint64_t nwResident = ~0ULL;  // bytes of the K3 sequence buffer currently on the device (already encoded)
	float lastKernelMs = 0;
	uint64_t launches = 0;
	uint64_t denseTraces = 0; // entries of the last gcgpu_extend call still in `compact`
};

static cudaError_t gcSyncStream(gcgpu_ctx* ctx)
{
	cudaError_t e = cudaEventRecord(ctx->evSync, ctx->stream);
	if (e != cudaSuccess) return e;
	return cudaEventSynchronize(ctx->evSync);
}

// every per-call copy across PCIe goes through here (gcgpu_transfer_bytes)
static cudaError_t gcCopy(gcgpu_ctx* ctx, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream)
{
	if (bytes == 0) return cudaSuccess;
	if (kind == cudaMemcpyHostToDevice) ctx->h2dBytes += bytes; else if (kind == cudaMemcpyDeviceToHost) ctx->d2hBytes += bytes;
	return cudaMemcpyAsync(dst, src, bytes, kind, stream);
}
// grow a device buffer to `bytes`, keeping its first `keep` bytes
static cudaError_t growKeep(gcgpu_ctx* ctx, DevBuf& b, size_t bytes, size_t keep, int line = __builtin_LINE())
{
	if (bytes <= b.cap) return cudaSuccess;
	if (keep == 0 || !b.p) return b.ensure(bytes, line);
	DevBuf nb;
	cudaError_t e = nb.ensure(bytes, line); // ensure() adds its own quarter of headroom; old and new buffer exist side by side until the copy is done
	if (e != cudaSuccess) return e;
	e = cudaMemcpyAsync(nb.p, b.p, keep, cudaMemcpyDeviceToDevice, ctx->stream);
	if (e != cudaSuccess) { nb.release(); return e; }
	e = cudaStreamSynchronize(ctx->stream);
	b.release();
	b = nb;
	return e;
}

// ------------------------------------------------------------------ K1 kernels
struct GcK1Desc
{
	uint64_t seqOff;
	uint64_t wsOff;
	uint64_t traceOff;
	int32_t seqLen;
	uint32_t node;
	uint32_t offset;
	uint32_t itemCap;
	uint32_t heapCap;
	uint32_t traceCap;
	uint32_t numSlices;
	uint32_t resultIndex;
};

static inline size_t alignUp(size_t x, size_t a) { return (x + a - 1) / a * a; }
static size_t k1WorkspaceBytes(uint32_t numSlices, uint32_t itemCap, uint32_t heapCap)
{
	return alignUp((size_t)(numSlices + 2) * sizeof(GcSliceMeta), 16) + (size_t)itemCap * sizeof(GcNodeItem) + (size_t)heapCap * 8 + (size_t)itemCap * 16 + (size_t)itemCap * 4; // slices | items | heap | item aux | slice keys
}

// Short work items (35-bp fragments: one or two slices): one thread = one item; the millions of
// independent items of a batch supply the parallelism.  Workspaces and trace slots are uniform, so
// the kernel derives them from the item index -- the host uploads nothing but the items themselves.
#define GC_K1_SHORT_SLOTS (2u << 20)
struct GcK1ShortLayout
{
	uint64_t wsBase, wsStride;      // bytes
	uint32_t first;                 // first item of this launch: a batch's items go through the same wsSlots slabs chunk after chunk
	uint64_t traceBase;             // entries
	uint32_t traceStride;           // entries
	uint32_t itemCap, heapCap, numSlices;
};
__device__ __forceinline__ GcK1Desc gc_k1_short_desc(const gcgpu_ext_item* __restrict__ items, const uint32_t* __restrict__ shortIdx, uint32_t t, const GcK1ShortLayout& lay)
{
	uint32_t idx = shortIdx ? shortIdx[t] : t;
	gcgpu_ext_item it = items[idx];
	GcK1Desc d;
	d.seqOff = it.seq_offset; d.seqLen = it.seq_len; d.node = it.node; d.offset = it.offset;
	d.wsOff = lay.wsBase + (uint64_t)(t - lay.first) * lay.wsStride;
	d.traceOff = lay.traceBase + (uint64_t)t * lay.traceStride;
	d.itemCap = lay.itemCap; d.heapCap = lay.heapCap; d.traceCap = lay.traceStride; d.numSlices = lay.numSlices; d.resultIndex = idx;
	return d;
}
__device__ __forceinline__ void gc_k1_workspace(const GcK1Desc& d, uint8_t* arena, GcColVV* cols, GcK1Workspace& ws);
// forward pass and backtrace are separate launches, as for the long items (instruction footprint, see below)
__global__ void __launch_bounds__(128) gc_k1_kernel(GcGraphView g, const GcViterbiTables* __restrict__ vt, GcK1Params prm, const uint8_t* __restrict__ seq,
	const gcgpu_ext_item* __restrict__ items, const uint32_t* __restrict__ shortIdx, uint32_t n, GcK1ShortLayout lay, uint8_t* arena, uint64_t* traceArena, GcK1Result* results,
	uint64_t* traceOffOfItem, uint32_t* overflow, int32_t* lastSlice)
{
	uint32_t t = lay.first + blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	GcK1Desc d = gc_k1_short_desc(items, shortIdx, t, lay);
	GcK1Workspace ws;
	gc_k1_workspace(d, arena, nullptr, ws); // the forward pass stores no columns (the last slice is flattened on the fly)
	GcK1Result res;
	res.score = GC_INT_MAX; res.traceLen = 0; res.itemsUsed = 0; res.columns = 0;
	if (d.seqLen < 0)
	{
		// this direction of the seed does not exist (seed at the first / last position of its sequence)
		res.status = GC_FAILED;
		results[d.resultIndex] = res; traceOffOfItem[d.resultIndex] = d.traceOff; lastSlice[t] = 0;
		return;
	}
	int32_t last = gc_k1_forward(g, *vt, prm, seq + d.seqOff, d.seqLen, d.node, d.offset, ws, res);
	if (res.status == GC_OK && last < 1) res.status = GC_FAILED;
	results[d.resultIndex] = res;
	traceOffOfItem[d.resultIndex] = d.traceOff;
	lastSlice[t] = last;
	if (res.status == GC_OVERFLOW_ITEMS || res.status == GC_OVERFLOW_HEAP) atomicAdd(overflow, 1u);
}
__global__ void __launch_bounds__(128) gc_k1_bt_kernel(GcGraphView g, const uint8_t* __restrict__ seq,
	const gcgpu_ext_item* __restrict__ items, const uint32_t* __restrict__ shortIdx, uint32_t n, GcK1ShortLayout lay, uint8_t* arena, uint64_t* traceArena, GcK1Result* results,
	const int32_t* __restrict__ lastSlice)
{
	uint32_t t = lay.first + blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	GcK1Desc d = gc_k1_short_desc(items, shortIdx, t, lay);
	GcK1Result res = results[d.resultIndex];
	if (res.status != GC_OK) return;
	GcColVV cols[64];
	GcK1Workspace ws;
	gc_k1_workspace(d, arena, cols, ws);
	gc_k1_backtrace(g, seq + d.seqOff, d.seqLen, ws, lastSlice[t], traceArena + d.traceOff, d.traceCap, res);
	results[d.resultIndex] = res;
}

// Long work items (whole-read extensions: ~80 slices, tens of thousands of dependent column steps):
// one WARP per item.  All 32 lanes execute the item in lockstep with identical values (one
// instruction stream, no divergence), which (a) keeps the walk at the single-warp issue rate instead
// of serialising 32 divergent walks, (b) lets loop-free lookups use the lanes -- the slice hash
// lookup becomes one ballot over the slice's items, the Eq masks of a slice eight ballots -- and
// (c) moves the recomputed node columns from per-thread local memory to shared memory.
// The walk is split into two launches, forward pass (slices, Viterbi cut) and backtrace: at any time all warps of the
// GPU run the same half of the code (the whole item is ~7 k instructions; with both halves resident `no_instruction` was
// 16 % of the stall cycles of the single kernel, profiles/r01h) and each half needs fewer registers.
__device__ __forceinline__ void gc_k1_workspace(const GcK1Desc& d, uint8_t* arena, GcColVV* cols, GcK1Workspace& ws)
{
	ws.cols = cols;
	uint8_t* base = arena + d.wsOff;
	ws.slices = (GcSliceMeta*)base;
	size_t slicesBytes = ((size_t)(d.numSlices + 2) * sizeof(GcSliceMeta) + 15) / 16 * 16;
	ws.items = (GcNodeItem*)(base + slicesBytes);
	ws.heap = (uint64_t*)(base + slicesBytes + (size_t)d.itemCap * sizeof(GcNodeItem));
	ws.itemCap = d.itemCap;
	ws.heapCap = d.heapCap;
}
// the lanes of the warp execute ONE work item in lock-step
__device__ __forceinline__ uint32_t gc_k1_group_setup(GcGraphView& g)
{
	g.coopLane = (int32_t)(threadIdx.x & 31);
	g.coopWidth = 32;
	g.coopShift = 0;
	g.coopMask = 0xFFFFFFFFu;
	return (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
}
// sequence offset and start cell of a long item come from the device-resident item (the host lays the slabs out from the lengths alone)
__device__ __forceinline__ GcK1Desc gc_k1_long_desc(const GcK1Desc* __restrict__ descs, const gcgpu_ext_item* __restrict__ items, uint32_t t)
{
	GcK1Desc d = descs[t];
	gcgpu_ext_item it = items[d.resultIndex];
	d.seqOff = it.seq_offset; d.node = it.node; d.offset = it.offset;
	return d;
}
template <int MIN_BLOCKS, int W>
__global__ void __launch_bounds__(128, MIN_BLOCKS) gc_k1_long_kernel(GcGraphView g, const GcViterbiTables* __restrict__ vt, GcK1Params prm, const uint8_t* __restrict__ seq, const gcgpu_ext_item* __restrict__ items,
	const GcK1Desc* __restrict__ descs, uint32_t n, uint8_t* arena, uint64_t* traceArena, GcK1Result* results, uint64_t* traceOffOfItem, uint32_t* overflow, int32_t* lastSlice)
{
	__shared__ uint64_t heapShared[128 / W][64]; // the node queue of the slice being filled: a dozen dependent accesses per node visit
	uint32_t t = gc_k1_group_setup(g);
	if (t >= n) return;
	GcK1Desc d = gc_k1_long_desc(descs, items, t);
	GcK1Workspace ws;
	gc_k1_workspace(d, arena, nullptr, ws); // the forward pass stores no columns (the last slice is flattened on the fly)
	if (ws.heapCap <= 64) ws.heap = heapShared[threadIdx.x / W];
	__shared__ uint32_t keyShared[128 / W][2][32];
	ws.keysA = keyShared[threadIdx.x / W][0]; ws.keysB = keyShared[threadIdx.x / W][1];
	GcK1Result res;
	res.score = GC_INT_MAX; res.traceLen = 0; res.itemsUsed = 0;
	int32_t last = gc_k1_forward(g, *vt, prm, seq + d.seqOff, d.seqLen, d.node, d.offset, ws, res);
	if (res.status == GC_OK && last < 1) res.status = GC_FAILED;
	results[d.resultIndex] = res;
	traceOffOfItem[d.resultIndex] = d.traceOff;
	lastSlice[t] = last;
	if (res.status == GC_OVERFLOW_ITEMS || res.status == GC_OVERFLOW_HEAP) atomicAdd(overflow, 1u);
}
template <int MIN_BLOCKS, int W>
__global__ void __launch_bounds__(128, MIN_BLOCKS) gc_k1_long_bt_kernel(GcGraphView g, const uint8_t* __restrict__ seq, const gcgpu_ext_item* __restrict__ items,
	const GcK1Desc* __restrict__ descs, uint32_t n, uint8_t* arena, uint64_t* traceArena, GcK1Result* results, const int32_t* __restrict__ lastSlice)
{
	__shared__ GcColVV colsShared[128 / W][64];
	uint32_t t = gc_k1_group_setup(g);
	if (t >= n) return;
	GcK1Desc d = gc_k1_long_desc(descs, items, t);
	GcK1Result res = results[d.resultIndex];
	if (res.status != GC_OK) return;
	GcK1Workspace ws;
	gc_k1_workspace(d, arena, colsShared[threadIdx.x / W], ws);
	gc_k1_backtrace(g, seq + d.seqOff, d.seqLen, ws, lastSlice[t], traceArena + d.traceOff, d.traceCap, res);
	results[d.resultIndex] = res;
}

// Long work items, lane-per-item form (gc_k1s.cuh): the 32 lanes of a warp walk 32 different items and meet at the shared
// column loop.  Items arrive sorted by length (longest first), so the lanes of a warp finish together.  The node queue of
// every lane sits in shared memory, entry-major (entry e of thread t at heapShared[e * GC_K1S_THREADS + t]).
#define GC_K1S_THREADS 64
#define GC_K1S_HEAP 32
__device__ __forceinline__ void gc_k1s_workspace(const GcK1Desc& d, uint8_t* arena, uint64_t* heapShared, GcK1SWorkspace& ws)
{
	uint8_t* base = arena + d.wsOff;
	ws.slices = (GcSliceMeta*)base;
	size_t slicesBytes = ((size_t)(d.numSlices + 2) * sizeof(GcSliceMeta) + 15) / 16 * 16;
	ws.items = (GcNodeItem*)(base + slicesBytes);
	ws.scratch = (uint32_t*)(base + slicesBytes + (size_t)d.itemCap * sizeof(GcNodeItem));
	ws.scratchCap = d.heapCap * 2;
	ws.aux = (GcItemAux*)(base + slicesBytes + (size_t)d.itemCap * sizeof(GcNodeItem) + (size_t)d.heapCap * 8);
	ws.keys = (uint32_t*)(base + slicesBytes + (size_t)d.itemCap * sizeof(GcNodeItem) + (size_t)d.heapCap * 8 + (size_t)d.itemCap * sizeof(GcItemAux));
	ws.itemCap = d.itemCap;
	ws.heap.base = heapShared + threadIdx.x; ws.heap.stride = GC_K1S_THREADS; ws.heap.cap = GC_K1S_HEAP;
}
// The launch's items are handed out one at a time (longest first, `next` = a counter in global memory): a lane whose item
// ends takes the next one.  `lanes` = lanes per warp that take items at all: a launch with fewer items than resident lanes
// spreads them over all the warps it can have resident instead of filling a few warps to 32.
struct GcK1SForwardQueue
{
	const GcK1Desc* descs; const gcgpu_ext_item* items; uint32_t n; uint32_t* next_; uint32_t lanes;
	const uint8_t* seq; uint8_t* arena; uint64_t* heapShared;
	GcK1Result* results; uint64_t* traceOffOfItem; uint32_t* overflow; int32_t* lastSlice;
	uint32_t t; GcK1Desc d;
	__device__ __forceinline__ bool next(GcK1SItem& it, GcK1SWorkspace& ws)
	{
		if ((threadIdx.x & 31) >= lanes) return false;
		t = atomicAdd(next_, 1u);
		if (t >= n) return false;
		d = gc_k1_long_desc(descs, items, t);
		gc_k1s_workspace(d, arena, heapShared, ws);
		it.seq = seq + d.seqOff; it.seqLen = d.seqLen; it.startNode = d.node; it.startOffset = d.offset; it.planeBit = d.seqOff;
		return true;
	}
	__device__ __forceinline__ void done(GcK1Result res, int32_t last)
	{
		if (res.status == GC_OK && last < 1) res.status = GC_FAILED;
		results[d.resultIndex] = res;
		traceOffOfItem[d.resultIndex] = d.traceOff;
		lastSlice[t] = last;
		if (res.status == GC_OVERFLOW_ITEMS || res.status == GC_OVERFLOW_HEAP) atomicAdd(overflow, 1u);
	}
};
// 6 resident blocks per SM: 168 registers (the compiler takes 188 unbounded; same speed alone -- r03o -- and one more block for the launches of other batches)
__global__ void __launch_bounds__(GC_K1S_THREADS, 6) gc_k1s_forward_kernel(GcGraphView g, const GcViterbiTables* __restrict__ vt, GcK1Params prm, const uint8_t* __restrict__ seq, const uint64_t* __restrict__ planes,
	const gcgpu_ext_item* __restrict__ items, const GcK1Desc* __restrict__ descs, uint32_t n, uint32_t* next, uint32_t lanes, uint8_t* arena, GcK1Result* results, uint64_t* traceOffOfItem, uint32_t* overflow, int32_t* lastSlice)
{
	__shared__ uint64_t heapShared[GC_K1S_HEAP * GC_K1S_THREADS];
	g.coopLane = -1;
	GcK1SForwardQueue q;
	q.descs = descs; q.items = items; q.n = n; q.next_ = next; q.lanes = lanes; q.seq = seq; q.arena = arena; q.heapShared = heapShared;
	q.results = results; q.traceOffOfItem = traceOffOfItem; q.overflow = overflow; q.lastSlice = lastSlice; q.t = 0;
	GcK1SWorkspace ws;
	ws.slices = nullptr; ws.items = nullptr; ws.keys = nullptr; ws.aux = nullptr; ws.scratch = nullptr; ws.scratchCap = 0; ws.itemCap = 0;
	ws.heap.base = heapShared + threadIdx.x; ws.heap.stride = GC_K1S_THREADS; ws.heap.cap = GC_K1S_HEAP;
	gc_k1s_forward_items(g, *vt, prm, planes, ws, q);
}
struct GcK1SBacktraceQueue
{
	const GcK1Desc* descs; const gcgpu_ext_item* items; uint32_t n; uint32_t* next_; uint32_t lanes;
	const uint8_t* seq; uint8_t* arena; uint64_t* traceArena;
	GcK1Result* results; const int32_t* lastSlice;
	GcK1Desc d;
	__device__ __forceinline__ bool next(GcK1SItem& it, GcK1SWorkspace& ws, int32_t& last, uint64_t*& traceOut, uint32_t& traceCap, GcK1Result& res)
	{
		if ((threadIdx.x & 31) >= lanes) return false;
		while (true)
		{
			uint32_t t = atomicAdd(next_, 1u);
			if (t >= n) return false;
			d = gc_k1_long_desc(descs, items, t);
			res = results[d.resultIndex];
			if (res.status != GC_OK) continue; // the forward pass failed: no trace
			gc_k1s_workspace(d, arena, nullptr, ws);
			it.seq = seq + d.seqOff; it.seqLen = d.seqLen; it.startNode = d.node; it.startOffset = d.offset; it.planeBit = d.seqOff;
			last = lastSlice[t];
			traceOut = traceArena + d.traceOff; traceCap = d.traceCap;
			return true;
		}
	}
	__device__ __forceinline__ void done(const GcK1Result& res) { results[d.resultIndex] = res; }
};
__global__ void __launch_bounds__(GC_K1S_THREADS, 8) gc_k1s_backtrace_kernel(GcGraphView g, const uint8_t* __restrict__ seq, const uint64_t* __restrict__ planes, const gcgpu_ext_item* __restrict__ items,
	const GcK1Desc* __restrict__ descs, uint32_t n, uint32_t* next, uint32_t lanes, uint8_t* arena, uint64_t* traceArena, GcK1Result* results, const int32_t* __restrict__ lastSlice)
{
	g.coopLane = -1;
	GcK1SBacktraceQueue q;
	q.descs = descs; q.items = items; q.n = n; q.next_ = next; q.lanes = lanes; q.seq = seq; q.arena = arena; q.traceArena = traceArena; q.results = results; q.lastSlice = lastSlice;
	GcK1SWorkspace ws;
	ws.slices = nullptr; ws.items = nullptr; ws.keys = nullptr; ws.aux = nullptr; ws.scratch = nullptr; ws.scratchCap = 0; ws.itemCap = 0;
	ws.heap.base = nullptr; ws.heap.stride = 0; ws.heap.cap = 0;
	GcColVV cols[64];
	gc_k1s_backtrace_items(g, planes, ws, cols, q);
}
// bit planes of the sequence buffer: for every block of 64 codes, four words (bit i of word b = code i has bit b: A C G T)
__global__ void gc_planes_kernel(const uint8_t* __restrict__ seq, uint64_t bytes, uint64_t blocks, uint64_t* __restrict__ planes)
{
	uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= blocks) return;
	uint64_t w[4] = { 0, 0, 0, 0 };
	uint64_t first = k * 64;
	for (uint32_t i = 0; i < 64 && first + i < bytes; i++)
	{
		uint64_t c = seq[first + i];
		w[0] |= (c & 1) << i; w[1] |= ((c >> 1) & 1) << i; w[2] |= ((c >> 2) & 1) << i; w[3] |= ((c >> 3) & 1) << i;
	}
	planes[4 * k] = w[0]; planes[4 * k + 1] = w[1]; planes[4 * k + 2] = w[2]; planes[4 * k + 3] = w[3];
}

// trace lengths of the finished items (input of the exclusive scan that places them in the dense buffer)
__global__ void gc_k1_lengths_kernel(const GcK1Result* __restrict__ results, uint32_t n, uint64_t* __restrict__ lens, uint64_t* total)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	GcK1Result r = results[i];
	lens[i] = r.status == GC_OK ? r.traceLen : 0;
}
__global__ void gc_k1_total_kernel(const uint64_t* __restrict__ lens, const uint64_t* __restrict__ offs, uint32_t n, uint64_t* total)
{
	*total = offs[n - 1] + lens[n - 1];
}
// one warp per item: copy its trace to its place in the dense buffer, write the public result record
__global__ void gc_k1_gather_kernel(const GcK1Result* __restrict__ results, const uint64_t* __restrict__ traceOffOfItem, const uint64_t* __restrict__ offs, uint32_t n,
	const uint64_t* __restrict__ traceArena, uint64_t* __restrict__ dense, uint64_t denseStart, gcgpu_ext_result* __restrict__ pub)
{
	uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	uint32_t lane = threadIdx.x & 31;
	if (warp >= n) return;
	GcK1Result r = results[warp];
	uint32_t len = r.status == GC_OK ? r.traceLen : 0;
	uint64_t dst = denseStart + offs[warp], src = traceOffOfItem[warp];
	for (uint32_t i = lane; i < len; i += 32) dense[dst + i] = traceArena[src + i];
	if (lane == 0)
	{
		gcgpu_ext_result o;
		o.status = r.status == GC_OK ? GCGPU_ITEM_OK : (r.status == GC_FAILED ? GCGPU_ITEM_FAILED : GCGPU_ITEM_INTERNAL);
		o.score = r.score; o.trace_len = len; o.reserved = 0; o.trace_offset = dst; o.columns = r.columns;
		pub[warp] = o;
	}
}

static int residentCreate(gcgpu_ctx* ctx, const gcgpu_graph* graph);
static void residentDestroy(gcgpu_ctx* ctx);

// ------------------------------------------------------------------ C ABI
extern "C" int gcgpu_version(void) { return GCGPU_VERSION; }
extern "C" const char* gcgpu_last_error(void) { return g_lastError.c_str(); }

extern "C" void gcgpu_destroy(gcgpu_ctx* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaFree(ctx->d_nodeLength); cudaFree(ctx->d_nodeSeq); cudaFree(ctx->d_inStart); cudaFree(ctx->d_inNbr); cudaFree(ctx->d_outStart); cudaFree(ctx->d_outNbr);
	cudaFree(ctx->d_componentNumber); cudaFree(ctx->d_linearizable); cudaFree(ctx->d_vt); cudaFree(ctx->d_nodeRec); cudaFree(ctx->d_outKey);
	cudaFree(ctx->d_compMap); cudaFree(ctx->d_compIdx); cudaFree(ctx->d_compStart); cudaFree(ctx->d_topoIds);
	cudaFree(ctx->d_mzSlots); ctx->seedBuf.release(); ctx->seedMatches.release();
	cudaFree(ctx->d_pathsStart); cudaFree(ctx->d_pathsK); cudaFree(ctx->d_backStart); cudaFree(ctx->d_backNode); cudaFree(ctx->d_backK); cudaFree(ctx->d_pathBase);
	ctx->seqBuf.release(); ctx->nwSeqBuf.release(); ctx->descBuf.release(); ctx->resBuf.release(); ctx->arena.release(); ctx->traceArena.release(); ctx->compact.release(); ctx->copyDesc.release(); ctx->itemsBuf.release(); ctx->planes.release(); ctx->k2Work.release(); ctx->k2Sort.release();
	residentDestroy(ctx);
	if (ctx->ev0) cudaEventDestroy(ctx->ev0);
	if (ctx->ev1) cudaEventDestroy(ctx->ev1);
	if (ctx->evSync) cudaEventDestroy(ctx->evSync);
	if (ctx->evFork) cudaEventDestroy(ctx->evFork);
	if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
	if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
	if (ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

extern "C" int gcgpu_create(int device, const gcgpu_graph* graph, const gcgpu_params* params, gcgpu_ctx** out)
{
	if (!graph || !out || graph->num_nodes == 0 || !graph->node_length || !graph->node_seq || !graph->in_start || !graph->out_start || !graph->component_number || !graph->linearizable)
		return setError(GCGPU_ERR_ARG, "gcgpu_create: missing graph arrays");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) return setError(GCGPU_ERR_CUDA, std::string("gcgpu_create: no CUDA device (") + cudaGetErrorString(e) + "); libgcgpu has no CPU fallback");
	if (device < 0 || device >= count) return setError(GCGPU_ERR_ARG, "gcgpu_create: bad device index");
	CUDA_TRY(cudaSetDevice(device));
	gcgpu_ctx* ctx = new gcgpu_ctx();
	ctx->device = device;
	ctx->params.initial_bandwidth = params ? params->initial_bandwidth : 10;
	if (ctx->params.initial_bandwidth < 1) { delete ctx; return setError(GCGPU_ERR_ARG, "gcgpu_create: bandwidth must be >= 1"); }
	uint32_t N = graph->num_nodes;
	ctx->numNodes = N;
	cudaError_t err = cudaSuccess;
	auto chk = [&err](cudaError_t x) { if (err == cudaSuccess) err = x; };
	chk(cudaDeviceGetAttribute(&ctx->numSMs, cudaDevAttrMultiProcessorCount, device));
	chk(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	chk(cudaEventCreate(&ctx->ev0));
	chk(cudaEventCreate(&ctx->ev1));
	chk(cudaEventCreateWithFlags(&ctx->evSync, cudaEventBlockingSync | cudaEventDisableTiming));
	chk(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
	chk(cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming));
	chk(cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming));
	chk(uploadArray(graph->node_length, N, &ctx->d_nodeLength));
	chk(uploadArray(graph->node_seq, 2 * (size_t)N, &ctx->d_nodeSeq));
	chk(uploadArray(graph->in_start, (size_t)N + 1, &ctx->d_inStart));
	chk(uploadArray(graph->in_nbr, graph->in_start[N], &ctx->d_inNbr));
	chk(uploadArray(graph->out_start, (size_t)N + 1, &ctx->d_outStart));
	chk(uploadArray(graph->out_nbr, graph->out_start[N], &ctx->d_outNbr));
	chk(uploadArray(graph->component_number, N, &ctx->d_componentNumber));
	chk(uploadArray(graph->linearizable, N, &ctx->d_linearizable));
	{
		// derived arrays of the lane-per-item K1 kernels: one 32-byte record per node, one queue key per out-edge
		GcGraphView hv;
		hv.numNodes = N; hv.nodeLength = graph->node_length; hv.nodeSeq = graph->node_seq; hv.inStart = graph->in_start; hv.inNbr = graph->in_nbr; hv.outStart = graph->out_start; hv.outNbr = graph->out_nbr;
		hv.componentNumber = graph->component_number; hv.linearizable = graph->linearizable; hv.nodeRec = nullptr; hv.outKey = nullptr;
		std::vector<GcNodeRec> recs; std::vector<uint64_t> outKeys;
		gcBuildNodeRecs(hv, recs, outKeys);
		chk(uploadArray(recs.data(), recs.size(), &ctx->d_nodeRec));
		chk(uploadArray(outKeys.data(), outKeys.size(), &ctx->d_outKey));
	}
	GcViterbiTables vt = gcMakeViterbiTables();
	chk(uploadArray(&vt, 1, &ctx->d_vt));
	if (graph->comp_map && graph->comp_idx && graph->comp_start && graph->topo_ids && graph->paths_start && graph->back_start)
	{
		chk(uploadArray(graph->comp_map, N, &ctx->d_compMap));
		chk(uploadArray(graph->comp_idx, N, &ctx->d_compIdx));
		chk(uploadArray(graph->comp_start, (size_t)graph->num_components + 1, &ctx->d_compStart));
		chk(uploadArray(graph->topo_ids, N, &ctx->d_topoIds));
		chk(uploadArray(graph->paths_start, (size_t)N + 1, &ctx->d_pathsStart));
		chk(uploadArray(graph->paths_k, graph->paths_start[N], &ctx->d_pathsK));
		chk(uploadArray(graph->back_start, (size_t)N + 1, &ctx->d_backStart));
		chk(uploadArray(graph->back_node, graph->back_start[N], &ctx->d_backNode));
		chk(uploadArray(graph->back_k, graph->back_start[N], &ctx->d_backK));
		{
			// global path ids: component c owns [pathBase[c], pathBase[c + 1]) (its MPC width = 1 + the largest path id on its nodes)
			std::vector<uint32_t> pathBase((size_t)graph->num_components + 1, 0);
			for (uint32_t c = 0; c < graph->num_components; c++)
			{
				uint32_t width = 0;
				for (uint32_t gidx = graph->comp_start[c]; gidx < graph->comp_start[c + 1]; gidx++)
					for (uint32_t q = graph->paths_start[gidx]; q < graph->paths_start[gidx + 1]; q++) if (graph->paths_k[q] + 1 > width) width = graph->paths_k[q] + 1;
				pathBase[c + 1] = pathBase[c] + width;
				if (graph->comp_start[c + 1] - graph->comp_start[c] > ctx->maxCompNodes) ctx->maxCompNodes = graph->comp_start[c + 1] - graph->comp_start[c];
			}
			ctx->totalPaths = pathBase[graph->num_components];
			chk(uploadArray(pathBase.data(), pathBase.size(), &ctx->d_pathBase));
		}
		ctx->haveMpc = true;
	}
	if (err != cudaSuccess)
	{
		std::string msg = std::string("gcgpu_create: ") + cudaGetErrorString(err);
		gcgpu_destroy(ctx);
		return setError(err == cudaErrorMemoryAllocation ? GCGPU_ERR_NOMEM : GCGPU_ERR_CUDA, msg);
	}
	ctx->view.numNodes = N;
	ctx->view.nodeLength = ctx->d_nodeLength; ctx->view.nodeSeq = ctx->d_nodeSeq;
	ctx->view.inStart = ctx->d_inStart; ctx->view.inNbr = ctx->d_inNbr; ctx->view.outStart = ctx->d_outStart; ctx->view.outNbr = ctx->d_outNbr;
	ctx->view.nodeRec = ctx->d_nodeRec; ctx->view.outKey = ctx->d_outKey;
	ctx->view.componentNumber = ctx->d_componentNumber; ctx->view.linearizable = ctx->d_linearizable; ctx->view.coopLane = -1; ctx->view.coopWidth = 32; ctx->view.coopMask = 0xFFFFFFFFu; ctx->view.coopShift = 0;
	ctx->mpc.compMap = ctx->d_compMap; ctx->mpc.compIdx = ctx->d_compIdx; ctx->mpc.compStart = ctx->d_compStart; ctx->mpc.topoIds = ctx->d_topoIds;
	ctx->mpc.pathsStart = ctx->d_pathsStart; ctx->mpc.pathsK = ctx->d_pathsK; ctx->mpc.backStart = ctx->d_backStart; ctx->mpc.backNode = ctx->d_backNode; ctx->mpc.backK = ctx->d_backK; ctx->mpc.pathBase = ctx->d_pathBase;
	int rrc = residentCreate(ctx, graph);
	if (rrc != GCGPU_OK) { gcgpu_destroy(ctx); return rrc; }
	*out = ctx;
	return GCGPU_OK;
}

extern "C" void* gcgpu_host_alloc(size_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}
extern "C" void gcgpu_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" float gcgpu_last_kernel_ms(gcgpu_ctx* ctx) { return ctx ? ctx->lastKernelMs : 0.f; }
extern "C" uint64_t gcgpu_launch_count(gcgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- K1 execution on items that already sit in device memory (uploaded by gcgpu_extend, or generated on the device from
// seed cells by gcgpu_extend_seeds / gcgpu_fragment_anchors).  The host only needs the items' sequence LENGTHS (slab layout).
// lens[i] < 0: item i does not exist (a seed at the first / last position of its sequence has one direction only); its
// result is GC_FAILED.  lens == nullptr: every item is shorter than GC_K1_LONG_ITEM, at most uniformMax long, and the
// kernel itself skips the absent ones (seq_len < 0 in the device item).
struct GcK1Run
{
	uint32_t n = 0;
	GcK1Result* dRes = nullptr; uint64_t* dSlot = nullptr; uint64_t* dLens = nullptr; uint64_t* dOffs = nullptr;
	gcgpu_ext_result* dPub = nullptr; uint64_t* dTotal = nullptr;
	uint64_t traceSlots = 0; // entries of the per-item trace slots (upper bound of the dense size)
};

__global__ void gc_k1_init_results_kernel(GcK1Result* results, uint64_t* slot, uint32_t n)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	GcK1Result r; r.status = GC_FAILED; r.score = GC_INT_MAX; r.traceLen = 0; r.itemsUsed = 0; r.columns = 0;
	results[i] = r;
	slot[i] = 0;
}

// bit planes of the resident K1 sequence buffer (after every upload / re-encoding of ctx->seqBuf)
static int buildPlanes(gcgpu_ctx* ctx, uint64_t bytes)
{
	uint64_t blocks = bytes / 64 + 2; // one block of padding: gc_eq_from_planes reads the block after the one it starts in
	CUDA_TRY(ctx->planes.ensure(blocks * 32));
	gc_planes_kernel<<<(unsigned)((blocks + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t*)ctx->seqBuf.p, bytes, blocks, (uint64_t*)ctx->planes.p);
	ctx->launches++;
	CUDA_TRY(cudaGetLastError());
	return GCGPU_OK;
}
// Long items run either lane-per-item (gc_k1s_*: 32 items per warp, ~1/12 of the issue slots per column step, but a walk takes
// ~270 us per 64-row slice because every lane waits for the slowest lane's bookkeeping -- a launch lasts as long as its longest
// item) or warp-per-item in lock-step (gc_k1_long_*: ~22 us per slice, 32 redundant lanes: a launch of more than a few thousand
// items saturates the integer pipe for ~22 ns per item-slice and starves the kernels of the other batches in flight).  Per launch,
// from the items' slice counts:
//   lock-step time ~ max(longest item x 22 us, sum of slices x 22 ns);  lane-per-item time ~ longest item x 270 us
// and the lane-per-item form is taken unless it is estimated more than four times slower (measurements: profiles/r03d, r03f, r03h,
// r03o: c2 round 2 goes lane-per-item, the one-seed-per-read first round and the tail rounds lock-step, ultra-long reads lock-step).
// Splitting ONE launch between the forms -- its longest items lock-step on a side stream, the rest lane-per-item beside them --
// does not pay: the two groups slow each other down (r03r: 49-50 ms against 43 ms lane-per-item only, e2e 171 vs 208 Mbp/s).
// GCGPU_K1_FORM=lane | lockstep forces one form.
static bool k1UseLaneForm(uint32_t nLong, uint64_t sumSlices, uint32_t maxSlices)
{
	const char* force = getenv("GCGPU_K1_FORM");
	if (force && !strcmp(force, "lane")) return true;
	if (force && !strcmp(force, "lockstep")) return false;
	if (nLong < 1024) return false;
	double tLock = std::max((double)maxSlices * 0.022, (double)sumSlices * 2.2e-5);               // ms
	double tLane = (double)maxSlices * 0.27 * std::max(1.0, (double)nLong / (32.0 * 148 * 12));   // ms
	return tLane <= 4.0 * tLock;
}
// Fewest lanes per warp that take items in a lane-per-item launch (GCGPU_K1_LANES overrides).  32: a launch lasts as long as
// its longest item whatever the number of warps (r03m: 47.6 / 49.7 / 49.9 / 49.4 ms at 32 / 16 / 8 / 4 lanes), and a launch
// spread over more warps leaves fewer SM slots to the launches of the other batches in flight (e2e 199.6 vs 174.4 Mbp/s at 32 vs 16)
static uint32_t k1MinLanes()
{
	static const uint32_t v = [] { const char* e = getenv("GCGPU_K1_LANES"); int x = e ? atoi(e) : 32; return (uint32_t)(x < 1 ? 1 : (x > 32 ? 32 : x)); }();
	return v;
}

static int k1Run(gcgpu_ctx* ctx, const gcgpu_ext_item* dItems, const int32_t* lens, uint32_t n, int32_t uniformMax, GcK1Run& run)
{
	// ---- split into long (warp per item, longest first) and short (thread per item) work
	int32_t maxShort = lens ? 0 : uniformMax;
	std::vector<uint32_t> longIdx, shortIdx;
	uint32_t nShort = n;
	bool needInit = false;
	if (lens)
	{
		#pragma omp parallel
		{
			std::vector<uint32_t> mine; int32_t myMax = 0;
			#pragma omp for schedule(static) nowait
			for (uint32_t i = 0; i < n; i++)
			{
				if (lens[i] >= GC_K1_LONG_ITEM) mine.push_back(i); else if (lens[i] > myMax) myMax = lens[i];
			}
			#pragma omp critical
			{ longIdx.insert(longIdx.end(), mine.begin(), mine.end()); if (myMax > maxShort) maxShort = myMax; }
		}
		std::sort(longIdx.begin(), longIdx.end(), [lens](uint32_t a, uint32_t b) { return lens[a] != lens[b] ? lens[a] > lens[b] : a < b; });
		shortIdx.reserve(n - longIdx.size());
		for (uint32_t i = 0; i < n; i++) if (lens[i] >= 0 && lens[i] < GC_K1_LONG_ITEM) shortIdx.push_back(i);
		nShort = (uint32_t)shortIdx.size();
		needInit = longIdx.size() + shortIdx.size() != n;
		if (shortIdx.size() == n) shortIdx.clear(); // identity
	}
	uint32_t nLong = (uint32_t)longIdx.size();
	// ---- layout: long items get individual slabs, short items uniform ones
	std::vector<GcK1Desc> descs(nLong);
	size_t wsTotal = 0; uint64_t traceTotal = 0;
	for (uint32_t k = 0; k < nLong; k++)
	{
		GcK1Desc& d = descs[k];
		d.seqOff = 0; d.seqLen = lens[longIdx[k]]; d.node = 0; d.offset = 0; // sequence offset and start cell: from the device item
		d.numSlices = (uint32_t)((d.seqLen + 63) / 64);
		d.itemCap = 24 + 8 * d.numSlices;
		d.heapCap = 64;
		d.traceCap = (uint32_t)(2 * (uint64_t)d.seqLen + 72);
		d.traceOff = traceTotal; traceTotal += d.traceCap;
		d.resultIndex = longIdx[k];
		d.wsOff = wsTotal;
		wsTotal += alignUp(k1WorkspaceBytes(d.numSlices, d.itemCap, d.heapCap), 128);
	}
	GcK1ShortLayout lay;
	lay.numSlices = (uint32_t)((maxShort + 63) / 64); if (lay.numSlices < 1) lay.numSlices = 1;
	lay.itemCap = 24 + 8 * lay.numSlices; lay.heapCap = 64;
	// slices | items | heap: the thread-per-item kernels keep no per-item seeding records or slice keys (gc_k1_workspace) -- a batch of
	// HiFi reads is ~10 M fragment items (c5: 25 GB of slabs per 16.8 Mbp batch even so)
	lay.wsBase = wsTotal; lay.wsStride = alignUp(alignUp((size_t)(lay.numSlices + 2) * sizeof(GcSliceMeta), 16) + (size_t)lay.itemCap * sizeof(GcNodeItem) + (size_t)lay.heapCap * 8, 128);
	lay.traceBase = traceTotal; lay.traceStride = (uint32_t)(2 * maxShort + 72);
	// the slabs are needed from an item's forward pass to its backtrace only: at most GC_K1_SHORT_SLOTS of them, shared by successive
	// chunks of the launch (c5: ~9 M fragment items per 16.8 Mbp batch were 20 GB of slabs per batch in flight)
	const char* slotEnv = getenv("GCGPU_K1_SHORT_SLOTS"); // the override is for tests
	const uint32_t slotLimit = slotEnv && atol(slotEnv) >= 128 ? (uint32_t)atol(slotEnv) : GC_K1_SHORT_SLOTS;
	const uint32_t shortSlots = std::min<uint32_t>(nShort, slotLimit);
	lay.first = 0;
	wsTotal += (size_t)shortSlots * lay.wsStride;
	traceTotal += (uint64_t)nShort * lay.traceStride;
	// small per-call device arrays: internal results | trace slot of every item | lengths | offsets | public results | scalars
	size_t offRes = 0, offSlot = alignUp(offRes + (size_t)n * sizeof(GcK1Result), 128), offLens = alignUp(offSlot + (size_t)n * 8, 128), offOffs = alignUp(offLens + (size_t)n * 8, 128);
	size_t offPub = alignUp(offOffs + (size_t)n * 8, 128), offScalars = alignUp(offPub + (size_t)n * sizeof(gcgpu_ext_result), 128);
	size_t offShortIdx = offScalars + 128, offLast = alignUp(offShortIdx + (size_t)shortIdx.size() * 4, 128), resEnd = offLast + (size_t)n * 4;
	CUDA_TRY(ctx->resBuf.ensure(resEnd));
	CUDA_TRY(ctx->arena.ensure(wsTotal));
	CUDA_TRY(ctx->traceArena.ensure(traceTotal * 8));
	uint8_t* R = (uint8_t*)ctx->resBuf.p;
	GcK1Result* dRes = (GcK1Result*)(R + offRes); uint64_t* dSlot = (uint64_t*)(R + offSlot); uint64_t* dLens = (uint64_t*)(R + offLens); uint64_t* dOffs = (uint64_t*)(R + offOffs);
	gcgpu_ext_result* dPub = (gcgpu_ext_result*)(R + offPub); uint32_t* dOverflow = (uint32_t*)(R + offScalars); uint64_t* dTotal = (uint64_t*)(R + offScalars + 8);
	int32_t* dLast = (int32_t*)(R + offLast);
	uint32_t* dShortIdx = shortIdx.empty() ? nullptr : (uint32_t*)(R + offShortIdx);
	run.n = n; run.dRes = dRes; run.dSlot = dSlot; run.dLens = dLens; run.dOffs = dOffs; run.dPub = dPub; run.dTotal = dTotal; run.traceSlots = traceTotal;
	CUDA_TRY(cudaMemsetAsync(R + offScalars, 0, 128, ctx->stream));
	if (!shortIdx.empty()) CUDA_TRY(gcCopy(ctx, dShortIdx, shortIdx.data(), shortIdx.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (nLong)
	{
		CUDA_TRY(ctx->descBuf.ensure(descs.size() * sizeof(GcK1Desc)));
		CUDA_TRY(gcCopy(ctx, ctx->descBuf.p, descs.data(), descs.size() * sizeof(GcK1Desc), cudaMemcpyHostToDevice, ctx->stream));
	}
	GcK1Params prm; prm.bandwidth = ctx->params.initial_bandwidth;
	float ms = 0;
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
	if (needInit) { gc_k1_init_results_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(dRes, dSlot, n); ctx->launches++; }
	if (nLong)
	{
		uint64_t sumSlices = 0; uint32_t maxSlices = 0;
		for (const GcK1Desc& d : descs) { sumSlices += d.numSlices; if (d.numSlices > maxSlices) maxSlices = d.numSlices; }
		const GcK1Desc* dDescs = (const GcK1Desc*)ctx->descBuf.p;
		if (k1UseLaneForm(nLong, sumSlices, maxSlices))
		{
			// lane per item: persistent warps whose lanes take items (sorted by length, longest first) from a counter until none is left
			struct Resident { int fwd = 1, bt = 1; };
			static const Resident resident = [] { // initialised once, by whichever host thread gets here first
				Resident r;
				if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r.fwd, gc_k1s_forward_kernel, GC_K1S_THREADS, 0) != cudaSuccess || r.fwd < 1) r.fwd = 1;
				if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r.bt, gc_k1s_backtrace_kernel, GC_K1S_THREADS, 0) != cudaSuccess || r.bt < 1) r.bt = 1;
				return r;
			}();
			uint32_t* dNext = (uint32_t*)(R + offScalars + 16); // two counters, zeroed with the scalars above
			const uint32_t warpsPerBlock = GC_K1S_THREADS / 32;
			auto shape = [&](int perSM, uint32_t& blocks, uint32_t& lanes)
			{
				uint32_t maxBlocks = (uint32_t)ctx->numSMs * (uint32_t)perSM;
				const char* cap = getenv("GCGPU_K1_BLOCKS"); // tests: a handful of blocks, so that every lane works through many items
				if (cap && atoi(cap) > 0) maxBlocks = std::min<uint32_t>(maxBlocks, (uint32_t)atoi(cap));
				lanes = (nLong + maxBlocks * warpsPerBlock - 1) / (maxBlocks * warpsPerBlock);
				if (lanes < k1MinLanes()) lanes = k1MinLanes();
				if (lanes > 32) lanes = 32;
				blocks = (nLong + lanes * warpsPerBlock - 1) / (lanes * warpsPerBlock);
				if (blocks > maxBlocks) blocks = maxBlocks;
			};
			uint32_t blocksF, lanesF, blocksB, lanesB;
			shape(resident.fwd, blocksF, lanesF);
			shape(resident.bt, blocksB, lanesB);
			gc_k1s_forward_kernel<<<blocksF, GC_K1S_THREADS, 0, ctx->stream>>>(ctx->view, ctx->d_vt, prm, (const uint8_t*)ctx->seqBuf.p, (const uint64_t*)ctx->planes.p, dItems, dDescs, nLong, dNext, lanesF, (uint8_t*)ctx->arena.p, dRes, dSlot, dOverflow, dLast);
			gc_k1s_backtrace_kernel<<<blocksB, GC_K1S_THREADS, 0, ctx->stream>>>(ctx->view, (const uint8_t*)ctx->seqBuf.p, (const uint64_t*)ctx->planes.p, dItems, dDescs, nLong, dNext + 1, lanesB, (uint8_t*)ctx->arena.p, (uint64_t*)ctx->traceArena.p, dRes, dLast);
		}
		else
		{
			// one warp per item in lock-step; resident blocks per SM: 5 (96 registers)
			gc_k1_long_kernel<5, 32><<<(nLong + 3) / 4, 128, 0, ctx->stream>>>(ctx->view, ctx->d_vt, prm, (const uint8_t*)ctx->seqBuf.p, dItems, dDescs, nLong, (uint8_t*)ctx->arena.p, (uint64_t*)ctx->traceArena.p, dRes, dSlot, dOverflow, dLast);
			gc_k1_long_bt_kernel<5, 32><<<(nLong + 3) / 4, 128, 0, ctx->stream>>>(ctx->view, (const uint8_t*)ctx->seqBuf.p, dItems, dDescs, nLong, (uint8_t*)ctx->arena.p, (uint64_t*)ctx->traceArena.p, dRes, dLast);
		}
		ctx->launches += 2;
	}
	if (nShort)
	{
		// lastSlice entries of the short items follow those of the long items
		for (uint32_t first = 0; first < nShort; first += shortSlots)
		{
			const uint32_t upTo = std::min(nShort, first + shortSlots);
			lay.first = first;
			gc_k1_kernel<<<(upTo - first + 127) / 128, 128, 0, ctx->stream>>>(ctx->view, ctx->d_vt, prm, (const uint8_t*)ctx->seqBuf.p, dItems, dShortIdx, upTo, lay, (uint8_t*)ctx->arena.p, (uint64_t*)ctx->traceArena.p, dRes, dSlot, dOverflow, dLast + nLong);
			gc_k1_bt_kernel<<<(upTo - first + 127) / 128, 128, 0, ctx->stream>>>(ctx->view, (const uint8_t*)ctx->seqBuf.p, dItems, dShortIdx, upTo, lay, (uint8_t*)ctx->arena.p, (uint64_t*)ctx->traceArena.p, dRes, dLast + nLong);
			ctx->launches += 2;
		}
	}
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	uint32_t overflow = 0;
	CUDA_TRY(gcCopy(ctx, &overflow, dOverflow, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->lastKernelMs += ms;
	GC_TRACE_MS("k1 (long+short)", n);
	if (overflow)
	{
		// rare: some item outgrew its slab (very wide band).  Re-run those with 4x larger slabs until they fit.
		std::vector<GcK1Result> hres(n);
		CUDA_TRY(gcCopy(ctx, hres.data(), dRes, (size_t)n * sizeof(GcK1Result), cudaMemcpyDeviceToHost, ctx->stream));
		std::vector<uint64_t> hslot(n);
		CUDA_TRY(gcCopy(ctx, hslot.data(), dSlot, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
		std::vector<gcgpu_ext_item> hitems(n);
		CUDA_TRY(gcCopy(ctx, hitems.data(), dItems, (size_t)n * sizeof(gcgpu_ext_item), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(gcSyncStream(ctx));
		uint32_t itemScale = 1, heapScale = 1;
		std::vector<uint32_t> todo;
		for (uint32_t i = 0; i < n; i++) if (hres[i].status == GC_OVERFLOW_ITEMS || hres[i].status == GC_OVERFLOW_HEAP) todo.push_back(i);
		DevBuf retryArena;
		for (int attempt = 0; attempt < 8 && !todo.empty(); attempt++)
		{
			bool needItems = false, needHeap = false;
			for (uint32_t i : todo) { if (hres[i].status == GC_OVERFLOW_ITEMS) needItems = true; else needHeap = true; }
			if (needItems) itemScale *= 4;
			if (needHeap) heapScale *= 4;
			std::vector<GcK1Desc> rd(todo.size());
			size_t ws = 0;
			for (size_t k = 0; k < todo.size(); k++)
			{
				const gcgpu_ext_item& it = hitems[todo[k]];
				GcK1Desc& d = rd[k];
				d.seqOff = it.seq_offset; d.seqLen = it.seq_len; d.node = it.node; d.offset = it.offset;
				d.numSlices = (uint32_t)((it.seq_len + 63) / 64);
				d.itemCap = (24 + 8 * d.numSlices) * itemScale; d.heapCap = 64 * heapScale;
				d.traceCap = (uint32_t)(2 * (uint64_t)it.seq_len + 72);
				d.traceOff = hslot[todo[k]]; d.resultIndex = todo[k];
				d.wsOff = ws; ws += alignUp(k1WorkspaceBytes(d.numSlices, d.itemCap, d.heapCap), 128);
			}
			cudaError_t e = retryArena.ensure(ws);
			if (e != cudaSuccess) { retryArena.release(); return setError(GCGPU_ERR_NOMEM, "gcgpu_extend: retry workspace"); }
			CUDA_TRY(ctx->descBuf.ensure(rd.size() * sizeof(GcK1Desc)));
			CUDA_TRY(gcCopy(ctx, ctx->descBuf.p, rd.data(), rd.size() * sizeof(GcK1Desc), cudaMemcpyHostToDevice, ctx->stream));
			CUDA_TRY(cudaMemsetAsync(dOverflow, 0, 4, ctx->stream));
			uint32_t m = (uint32_t)rd.size();
			CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
			gc_k1_long_kernel<5, 32><<<(m + 3) / 4, 128, 0, ctx->stream>>>(ctx->view, ctx->d_vt, prm, (const uint8_t*)ctx->seqBuf.p, dItems, (const GcK1Desc*)ctx->descBuf.p, m, (uint8_t*)retryArena.p, (uint64_t*)ctx->traceArena.p, dRes, dSlot, dOverflow, dLast);
			gc_k1_long_bt_kernel<5, 32><<<(m + 3) / 4, 128, 0, ctx->stream>>>(ctx->view, (const uint8_t*)ctx->seqBuf.p, dItems, (const GcK1Desc*)ctx->descBuf.p, m, (uint8_t*)retryArena.p, (uint64_t*)ctx->traceArena.p, dRes, dLast);
			ctx->launches += 2;
			CUDA_TRY(cudaGetLastError());
			CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
			CUDA_TRY(gcSyncStream(ctx));
			CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
			ctx->lastKernelMs += ms;
			CUDA_TRY(gcCopy(ctx, hres.data(), dRes, (size_t)n * sizeof(GcK1Result), cudaMemcpyDeviceToHost, ctx->stream));
			CUDA_TRY(gcSyncStream(ctx));
			std::vector<uint32_t> again;
			for (uint32_t i : todo) if (hres[i].status == GC_OVERFLOW_ITEMS || hres[i].status == GC_OVERFLOW_HEAP) again.push_back(i);
			todo.swap(again);
		}
		retryArena.release();
		if (!todo.empty()) return setError(GCGPU_ERR_NOMEM, "gcgpu_extend: " + std::to_string(todo.size()) + " work items still overflow their workspace after 8 attempts");
	}
	return GCGPU_OK;
}

// dense traces + public results, all on the device: lengths -> exclusive scan -> gather into `dense` (grown to fit,
// contents below denseStart preserved) from entry denseStart on.  *total = entries gathered.
static int k1Gather(gcgpu_ctx* ctx, const GcK1Run& run, DevBuf& dense, uint64_t denseStart, uint64_t* total)
{
	uint32_t n = run.n;
	float ms = 0;
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
	gc_k1_lengths_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(run.dRes, n, run.dLens, run.dTotal);
	size_t scanBytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, run.dLens, run.dOffs, (int)n, ctx->stream);
	CUDA_TRY(ctx->copyDesc.ensure(scanBytes + 16));
	cub::DeviceScan::ExclusiveSum(ctx->copyDesc.p, scanBytes, run.dLens, run.dOffs, (int)n, ctx->stream);
	gc_k1_total_kernel<<<1, 1, 0, ctx->stream>>>(run.dLens, run.dOffs, n, run.dTotal);
	uint64_t used = 0;
	CUDA_TRY(gcCopy(ctx, &used, run.dTotal, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	// growing the dense buffer holds its old and new copy for a moment: the slabs of the launch that just ended are dead by now, and
	// when they are tens of GB (ultra-long reads, the whole read set as one batch) they are given back first
	if ((denseStart + used) * 8 + 16 > dense.cap && ctx->arena.cap > ((size_t)8 << 30)) ctx->arena.release();
	CUDA_TRY(growKeep(ctx, dense, (denseStart + used) * 8 + 16, denseStart * 8));
	gc_k1_gather_kernel<<<(n + 3) / 4, 128, 0, ctx->stream>>>(run.dRes, run.dSlot, run.dOffs, n, (const uint64_t*)ctx->traceArena.p, (uint64_t*)dense.p, denseStart, run.dPub);
	ctx->launches += 4;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->lastKernelMs += ms;
	GC_TRACE_MS("k1 scan+gather", n);
	*total = used;
	return GCGPU_OK;
}

extern "C" int gcgpu_extend(gcgpu_ctx* ctx, const uint8_t* seq, uint64_t seq_bytes, const gcgpu_ext_item* items, uint32_t n,
	gcgpu_ext_result* results, uint64_t* traces, uint64_t trace_capacity, uint64_t* trace_used)
{
	if (!ctx || (!items && n) || (!results && n) || !trace_used) return setError(GCGPU_ERR_ARG, "gcgpu_extend: null argument");
	*trace_used = 0;
	ctx->lastKernelMs = 0;
	if (n == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	int bad = -1;
	std::vector<int32_t> lens(n);
	{
		uint32_t numNodes = ctx->numNodes;
		#pragma omp parallel for schedule(static)
		for (uint32_t i = 0; i < n; i++)
		{
			const gcgpu_ext_item& it = items[i];
			if (it.seq_len < 0 || it.seq_offset + (uint64_t)it.seq_len > seq_bytes || it.node >= numNodes || it.seq_len >= (1 << 24)) { _Pragma("omp critical") bad = (int)i; }
			lens[i] = it.seq_len;
		}
	}
	if (bad >= 0) return setError(GCGPU_ERR_ARG, "gcgpu_extend: item " + std::to_string(bad) + " out of range");
	if (seq)
	{
		CUDA_TRY(ctx->seqBuf.ensure(seq_bytes + 16));
		if (seq_bytes) CUDA_TRY(gcCopy(ctx, ctx->seqBuf.p, seq, seq_bytes, cudaMemcpyHostToDevice, ctx->stream));
		ctx->seqResident = seq_bytes;
		{ int prc = buildPlanes(ctx, seq_bytes); if (prc != GCGPU_OK) return prc; }
	}
	else if (ctx->seqResident != seq_bytes) return setError(GCGPU_ERR_ARG, "gcgpu_extend: seq == NULL but no sequence buffer of this size is resident");
	CUDA_TRY(ctx->itemsBuf.ensure((size_t)n * sizeof(gcgpu_ext_item)));
	CUDA_TRY(gcCopy(ctx, ctx->itemsBuf.p, items, (size_t)n * sizeof(gcgpu_ext_item), cudaMemcpyHostToDevice, ctx->stream));
	GcK1Run run;
	int rc = k1Run(ctx, (const gcgpu_ext_item*)ctx->itemsBuf.p, lens.data(), n, 0, run);
	if (rc != GCGPU_OK) return rc;
	uint64_t used = 0;
	rc = k1Gather(ctx, run, ctx->compact, 0, &used);
	if (rc != GCGPU_OK) return rc;
	CUDA_TRY(gcCopy(ctx, results, run.dPub, (size_t)n * sizeof(gcgpu_ext_result), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	*trace_used = used;
	ctx->denseTraces = used;
	if (!traces && trace_capacity == 0) used = 0; // two-phase use: the caller sizes its buffer from *trace_used and calls gcgpu_fetch_traces
	if (used > trace_capacity) return setError(GCGPU_ERR_ARG, "gcgpu_extend: trace buffer too small, need " + std::to_string(used) + " entries");
	if (used)
	{
		if (!traces) return setError(GCGPU_ERR_ARG, "gcgpu_extend: null trace buffer");
		CUDA_TRY(gcCopy(ctx, traces, ctx->compact.p, used * 8, cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(gcSyncStream(ctx));
	}
	bool internal = false;
	#pragma omp parallel for schedule(static) reduction(||: internal)
	for (uint32_t i = 0; i < n; i++) if (results[i].status == GCGPU_ITEM_INTERNAL) internal = true;
	if (internal) return setError(GCGPU_ERR_INTERNAL, "gcgpu_extend: a work item reached a state the reference asserts on (see per-item status)");
	return GCGPU_OK;
}

extern "C" int gcgpu_fetch_traces(gcgpu_ctx* ctx, uint64_t* traces, uint64_t first, uint64_t count)
{
	if (!ctx || (!traces && count)) return setError(GCGPU_ERR_ARG, "gcgpu_fetch_traces: null argument");
	if (first + count > ctx->denseTraces) return setError(GCGPU_ERR_ARG, "gcgpu_fetch_traces: range beyond the traces of the last gcgpu_extend call");
	if (count == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	CUDA_TRY(gcCopy(ctx, traces, (const uint64_t*)ctx->compact.p + first, count * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	return GCGPU_OK;
}

// ------------------------------------------------------------------ S0 kernels
// One block per read; the block walks the read in tiles of blockDim positions, one position per
// thread (gc_seed.cuh).  WRITE == false counts the read's matches, WRITE == true writes them in
// position order at offs[read] (ballot rank inside the warp + running prefix over warps and tiles).
// HBM-bound gather: 16 code bytes per position from L1 + one 16-byte table slot per emitted k-mer.
template <bool WRITE>
__global__ void __launch_bounds__(256) gc_seed_kernel(GcMzView mz, const uint8_t* __restrict__ seq, const gcgpu_seed_read* __restrict__ reads, uint32_t n,
	uint64_t* __restrict__ counts, const uint64_t* __restrict__ offs, gcgpu_seed_match* __restrict__ out)
{
	__shared__ uint32_t warpTotal[8];
	uint32_t r = blockIdx.x;
	if (r >= n) return;
	gcgpu_seed_read rd = reads[r];
	const uint8_t* codes = seq + rd.seq_offset;
	uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t running = 0;
	uint64_t base = WRITE ? offs[r] : 0;
	for (int32_t tile = 0; tile < rd.seq_len; tile += (int32_t)blockDim.x)
	{
		int32_t i = tile + (int32_t)threadIdx.x;
		uint32_t start = 0, count = 0;
		bool hit = gc_seed_position(mz, codes, rd.seq_len, i, start, count);
		uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
		if (lane == 0) warpTotal[warp] = __popc(m);
		__syncthreads();
		uint32_t before = 0, total = 0;
		for (uint32_t w = 0; w < (blockDim.x >> 5); w++) { uint32_t t = warpTotal[w]; if (w < warp) before += t; total += t; }
		if (WRITE && hit)
		{
			gcgpu_seed_match o; o.pos = (uint32_t)i; o.start = start; o.count = count;
			out[base + running + before + __popc(m & ((1u << lane) - 1))] = o;
		}
		running += total;
		__syncthreads();
	}
	if (!WRITE && threadIdx.x == 0) { counts[r] = running; if (r == 0) counts[n] = 0; }
}

extern "C" int gcgpu_set_minimizer_index(gcgpu_ctx* ctx, const gcgpu_minimizer_index* idx)
{
	if (!ctx || !idx || (idx->num_kmers && (!idx->kmers || !idx->kmer_start))) return setError(GCGPU_ERR_ARG, "gcgpu_set_minimizer_index: null argument");
	if (idx->k < 1 || idx->k > 31 || idx->window < idx->k) return setError(GCGPU_ERR_ARG, "gcgpu_set_minimizer_index: need 1 <= k <= 31 and window >= k");
	CUDA_TRY(cudaSetDevice(ctx->device));
	uint64_t cap = 16;
	while (cap < idx->num_kmers * 2 + 2) cap <<= 1;
	std::vector<GcMzSlot> slots(cap);
	for (auto& s : slots) { s.key = 0; s.start = 0; s.count = 0xFFFFFFFFu; }
	for (uint64_t i = 0; i < idx->num_kmers; i++)
	{
		uint64_t h = gc_mz_hash(idx->kmers[i]) & (cap - 1);
		while (slots[h].count != 0xFFFFFFFFu && slots[h].key != idx->kmers[i]) h = (h + 1) & (cap - 1);
		slots[h].key = idx->kmers[i]; slots[h].start = idx->kmer_start[i]; slots[h].count = idx->kmer_start[i + 1] - idx->kmer_start[i];
	}
	if (ctx->d_mzSlots) { cudaFree(ctx->d_mzSlots); ctx->d_mzSlots = nullptr; }
	CUDA_TRY(uploadArray(slots.data(), slots.size(), &ctx->d_mzSlots));
	ctx->mz.slots = ctx->d_mzSlots; ctx->mz.mask = cap - 1; ctx->mz.k = idx->k; ctx->mz.realWindow = idx->window - idx->k + 1; ctx->mz.maxCount = idx->max_count;
	ctx->haveMz = true;
	return GCGPU_OK;
}

extern "C" int gcgpu_seed(gcgpu_ctx* ctx, const uint8_t* seq, uint64_t seq_bytes, const gcgpu_seed_read* reads, uint32_t n,
	uint64_t* match_offsets, gcgpu_seed_match* matches, uint64_t capacity, uint64_t* used_out)
{
	if (!ctx || (!reads && n) || !match_offsets || !used_out) return setError(GCGPU_ERR_ARG, "gcgpu_seed: null argument");
	if (!ctx->haveMz) return setError(GCGPU_ERR_ARG, "gcgpu_seed: no minimizer index (gcgpu_set_minimizer_index)");
	*used_out = 0;
	ctx->lastKernelMs = 0;
	ctx->denseMatches = 0;
	match_offsets[0] = 0;
	if (n == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	for (uint32_t i = 0; i < n; i++)
		if (reads[i].seq_len < 0 || reads[i].seq_offset + (uint64_t)reads[i].seq_len > seq_bytes) return setError(GCGPU_ERR_ARG, "gcgpu_seed: read " + std::to_string(i) + " out of range");
	if (seq)
	{
		CUDA_TRY(ctx->seqBuf.ensure(seq_bytes + 16));
		if (seq_bytes) CUDA_TRY(gcCopy(ctx, ctx->seqBuf.p, seq, seq_bytes, cudaMemcpyHostToDevice, ctx->stream));
		ctx->seqResident = seq_bytes;
		{ int prc = buildPlanes(ctx, seq_bytes); if (prc != GCGPU_OK) return prc; }
	}
	else if (ctx->seqResident != seq_bytes) return setError(GCGPU_ERR_ARG, "gcgpu_seed: seq == NULL but no sequence buffer of this size is resident");
	// device arrays: reads | counts[n+1] | offsets[n+1] | scan scratch
	size_t offReads = 0, offCounts = alignUp((size_t)n * sizeof(gcgpu_seed_read), 128), offOffs = alignUp(offCounts + ((size_t)n + 1) * 8, 128), offScan = alignUp(offOffs + ((size_t)n + 1) * 8, 128);
	size_t scanBytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (int)n + 1, ctx->stream);
	CUDA_TRY(ctx->seedBuf.ensure(offScan + scanBytes + 16));
	uint8_t* B = (uint8_t*)ctx->seedBuf.p;
	gcgpu_seed_read* dReads = (gcgpu_seed_read*)(B + offReads); uint64_t* dCounts = (uint64_t*)(B + offCounts); uint64_t* dOffs = (uint64_t*)(B + offOffs);
	CUDA_TRY(gcCopy(ctx, dReads, reads, (size_t)n * sizeof(gcgpu_seed_read), cudaMemcpyHostToDevice, ctx->stream));
	float ms = 0;
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
	gc_seed_kernel<false><<<n, 256, 0, ctx->stream>>>(ctx->mz, (const uint8_t*)ctx->seqBuf.p, dReads, n, dCounts, nullptr, nullptr);
	cub::DeviceScan::ExclusiveSum(B + offScan, scanBytes, dCounts, dOffs, (int)n + 1, ctx->stream);
	ctx->launches += 2;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(gcCopy(ctx, match_offsets, dOffs, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	uint64_t used = match_offsets[n];
	CUDA_TRY(ctx->seedMatches.ensure(used * sizeof(gcgpu_seed_match) + 16));
	gc_seed_kernel<true><<<n, 256, 0, ctx->stream>>>(ctx->mz, (const uint8_t*)ctx->seqBuf.p, dReads, n, nullptr, dOffs, (gcgpu_seed_match*)ctx->seedMatches.p);
	ctx->launches++;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->lastKernelMs = ms;
	GC_TRACE_MS("s0 seed lookups", n);
	ctx->denseMatches = used;
	*used_out = used;
	if (!matches && capacity == 0) return GCGPU_OK;
	if (used > capacity) return setError(GCGPU_ERR_ARG, "gcgpu_seed: match buffer too small, need " + std::to_string(used) + " entries");
	if (used)
	{
		if (!matches) return setError(GCGPU_ERR_ARG, "gcgpu_seed: null match buffer");
		CUDA_TRY(gcCopy(ctx, matches, ctx->seedMatches.p, used * sizeof(gcgpu_seed_match), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(gcSyncStream(ctx));
	}
	return GCGPU_OK;
}

extern "C" int gcgpu_fetch_seed_matches(gcgpu_ctx* ctx, gcgpu_seed_match* matches, uint64_t first, uint64_t count)
{
	if (!ctx || (!matches && count)) return setError(GCGPU_ERR_ARG, "gcgpu_fetch_seed_matches: null argument");
	if (first + count > ctx->denseMatches) return setError(GCGPU_ERR_ARG, "gcgpu_fetch_seed_matches: range beyond the matches of the last gcgpu_seed call");
	if (count == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	CUDA_TRY(gcCopy(ctx, matches, (const gcgpu_seed_match*)ctx->seedMatches.p + first, count * sizeof(gcgpu_seed_match), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	return GCGPU_OK;
}

// ------------------------------------------------------------------ integer-pipe peak (measurement aid)
// K1/K3 are bound by the 32-bit integer pipes, not by HBM: this micro-benchmark measures the
// sustained rate of independent 32-bit logic/add instructions (the mix of a Myers column step: LOP3 and IADD3,
// 16 independent chains per thread) so that bench.py can state the kernels' int32 op rate as a fraction of it.
__global__ void __launch_bounds__(256) gc_int_peak_kernel(uint32_t* out, uint32_t iters, uint32_t seed)
{
	uint32_t r[16];
	#pragma unroll
	for (int i = 0; i < 16; i++) r[i] = seed + threadIdx.x * 16 + i;
	uint32_t a = seed | 1, b = blockIdx.x + 3;
	for (uint32_t it = 0; it < iters; it++)
	{
		#pragma unroll
		for (int i = 0; i < 16; i++)
		{
			r[i] = (r[i] ^ a) & (r[i] | b); // LOP3
			r[i] = r[i] + a + b;              // IADD3
		}
	}
	uint32_t x = 0;
	#pragma unroll
	for (int i = 0; i < 16; i++) x ^= r[i];
	if (x == 0x12345678u) out[0] = x; // keeps the chains alive
}

extern "C" int gcgpu_int_peak(gcgpu_ctx* ctx, double* int32_ops_per_s)
{
	if (!ctx || !int32_ops_per_s) return setError(GCGPU_ERR_ARG, "gcgpu_int_peak: null argument");
	CUDA_TRY(cudaSetDevice(ctx->device));
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
	CUDA_TRY(ctx->seedBuf.ensure(256));
	const uint32_t iters = 1 << 14;
	const int blocks = prop.multiProcessorCount * 8, threads = 256;
	float best = 1e30f;
	for (int rep = 0; rep < 5; rep++)
	{
		float ms = 0;
		CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
		gc_int_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->seedBuf.p, iters, 12345u + rep);
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
		CUDA_TRY(gcSyncStream(ctx));
		CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		if (rep > 0 && ms < best) best = ms;
	}
	*int32_ops_per_s = (double)blocks * threads * (double)iters * 32.0 / (best / 1e3);
	return GCGPU_OK;
}

// ------------------------------------------------------------------ K3 kernels
__global__ void gc_k3_encode_kernel(uint8_t* seq, uint64_t n)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride)
	{
		uint8_t c = seq[i];
		seq[i] = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
	}
}

struct GcK3Desc
{
	uint64_t qOff, tOff, wsOff, opsOff;
	int32_t q, t, kHint, best;
	uint32_t resultIndex, opsCap;
};
struct GcK3Out { int32_t status; int32_t distance; uint32_t opsLen; uint32_t pad; uint64_t blocks; };

static size_t k3DistWorkspaceBytes(int32_t q) { size_t nb = (size_t)(q + 63) / 64 + 1; return alignUp(nb * 4 * 8 + nb * sizeof(GcK3Block), 128); }
static const uint32_t K3_STORE_CAP = 52432, K3_COL_CAP = 37456, K3_STACK_CAP = 96;
static size_t k3PathWorkspaceBytes(int32_t q)
{
	size_t nb = (size_t)(q + 63) / 64 + 1;
	return alignUp(2 * nb * 4 * 8 + 2 * nb * sizeof(GcK3Block) + (size_t)K3_STORE_CAP * sizeof(GcK3Block) + (size_t)K3_COL_CAP * 4 + (size_t)K3_STACK_CAP * sizeof(GcK3Frame) + 64, 128);
}

// one thread = one edlib NW distance
__global__ void __launch_bounds__(64) gc_k3_distance_kernel(const uint8_t* __restrict__ seq, const GcK3Desc* __restrict__ descs, uint32_t n, uint8_t* arena, GcK3Out* out)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	GcK3Desc d = descs[t];
	GcK3Out o; o.status = GC_OK; o.opsLen = 0; o.pad = 0; o.blocks = 0;
	int32_t nb = (d.q + 63) / 64; if (nb < 1) nb = 1;
	uint64_t* peq = (uint64_t*)(arena + d.wsOff);
	GcK3Block* blocks = (GcK3Block*)(peq + 4 * (size_t)nb);
	gc_k3_build_peq(seq + d.qOff, d.q, peq, nb);
	uint64_t work = 0;
	o.distance = gc_k3_distance(peq, nb, d.q, seq + d.tOff, d.t, blocks, d.kHint, work);
	o.blocks = work;
	out[d.resultIndex] = o;
}


// ---- warp form (gc_k3w.cuh): one warp = one NW distance, lanes = groups of NB 64-row blocks on a
// skewed wavefront, block state in registers, one shuffle per column step.
template <int NB, bool STORE>
__device__ __forceinline__ uint32_t gc_k3w_run_pass_impl(const GcK3wPass& p, GcK3Block* blocksOut)
{
	int lane = threadIdx.x & 31;
	int src = (lane + 31) & 31;
	GcK3wLane<NB> s;
	gc_k3w_lane_init(p, s, lane);
	uint32_t send = 0;
	int32_t tau = 0;
	while (tau <= p.tauEnd)
	{
		int32_t ev = __reduce_min_sync(0xFFFFFFFFu, gc_k3w_next_event(p, s, tau));
		if (ev > tau)
		{
			// no lane has an event before `ev`: a segment of plain steps
			int32_t end = ev <= p.tauEnd ? ev : p.tauEnd + 1;
			GcK3wSegment seg = gc_k3w_segment(p, s, tau);
			for (; tau < end; tau++)
			{
				uint32_t recv = __shfl_sync(0xFFFFFFFFu, send, src);
				send = gc_k3w_lane_fast_step<NB, STORE>(p, s, seg, tau, recv);
			}
		}
		else
		{
			uint32_t recv = __shfl_sync(0xFFFFFFFFu, send, src);
			send = gc_k3w_lane_step(p, s, tau, recv, blocksOut);
			tau++;
		}
	}
	__syncwarp();
	return s.work;
}
template <int NB>
__device__ __forceinline__ uint32_t gc_k3w_run_pass(const GcK3wPass& p, GcK3Block* blocksOut)
{
	return p.store ? gc_k3w_run_pass_impl<NB, true>(p, blocksOut) : gc_k3w_run_pass_impl<NB, false>(p, blocksOut);
}

#define GC_K3W_NEED_LARGER (-2)
// the warp form holds one or two blocks per lane (cutoff bands up to ~4000 diagonals); wider bands go to the block form (gc_k3b_*)
__device__ __forceinline__ int gc_k3w_round_nb(int nb) { return nb <= 1 ? 1 : (nb <= 2 ? 2 : 0); }
__device__ __forceinline__ uint32_t gc_k3w_dispatch(const GcK3wPass& p, int NB, GcK3Block* blocksOut)
{
	return NB == 1 ? gc_k3w_run_pass<1>(p, blocksOut) : gc_k3w_run_pass<2>(p, blocksOut);
}

// one warp = one edlib NW distance (k doubling of edlib.cpp:193-212 from the item's first cutoff)
__global__ void __launch_bounds__(128) gc_k3w_distance_kernel(const uint8_t* __restrict__ seq, const GcK3Desc* __restrict__ descs, uint32_t n, uint8_t* arena, GcK3Out* out)
{
	uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int lane = threadIdx.x & 31;
	if (w >= n) return;
	GcK3Desc d = descs[w];
	int32_t q = d.q, t = d.t;
	int32_t nb = (q + 63) / 64; if (nb < 1) nb = 1;
	uint64_t* peq = (uint64_t*)(arena + d.wsOff);
	GcK3Block* blocks = (GcK3Block*)(peq + 4 * (size_t)nb);
	const uint8_t* query = seq + d.qOff;
	const uint8_t* target = seq + d.tOff;
	for (int32_t b = lane; b < nb; b += 32)
	{
		uint64_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
		int32_t lim = q - b * 64; if (lim > 64) lim = 64;
		for (int32_t i = 0; i < lim; i++)
		{
			uint8_t c = query[b * 64 + i];
			uint64_t bit = 1ULL << i;
			e0 |= c == 0 ? bit : 0; e1 |= c == 1 ? bit : 0; e2 |= c == 2 ? bit : 0; e3 |= c == 3 ? bit : 0;
		}
		peq[b] = e0; peq[nb + b] = e1; peq[2 * (size_t)nb + b] = e2; peq[3 * (size_t)nb + b] = e3;
	}
	__syncwarp();
	GcK3Out o; o.status = GC_OK; o.opsLen = 0; o.pad = 0; o.blocks = 0; o.distance = -1;
	uint32_t work = 0;
	if (q == 0 || t == 0) o.distance = q > t ? q : t;
	else
	{
		int32_t k = d.kHint < 64 ? 64 : d.kHint;
		int32_t diff = q > t ? q - t : t - q, mx = q > t ? q : t;
		while (true)
		{
			if (k >= diff)
			{
				int32_t kk = k > mx ? mx : k;
				int NB = gc_k3w_round_nb(gc_k3w_blocks_per_lane(q, t, kk));
				if (NB == 0) { o.distance = GC_K3W_NEED_LARGER; o.pad = (uint32_t)k; break; }
				GcK3wPass p = gc_k3w_make_pass(peq, nb, 0, q, target, 0, 1, t, kk, t - 1, NB);
				work += gc_k3w_dispatch(p, NB, blocks);
				int32_t v = gc_k3_cell(blocks[(q - 1) >> 6], q - 1);
				if (v <= kk) { o.distance = v; break; }
			}
			k *= 2;
		}
	}
	for (int off = 16; off > 0; off >>= 1) work += __shfl_down_sync(0xFFFFFFFFu, work, off);
	if (lane == 0) { o.blocks = work; out[d.resultIndex] = o; }
}

// ---- block form of the distance pass: the same skewed wavefront over the lanes of GC_K3B_WARPS warps.  A single warp holds a
// band of at most 64 * NB * 31 diagonals in its registers (NB <= 16: ~31 k diagonals, and the whole pass runs at one warp's issue
// rate); the cutoff bands of ultra-long reads (75 kb at 12 % error: k = 16384, 32 k diagonals) do not fit, and the widest
// alignments of a 10-kb batch set its latency.  Here lane 31 of a warp hands its horizontal delta to lane 0 of the next warp
// through shared memory (double-buffered, one __syncthreads per wavefront step), so 256 lanes share the band: NB <= 4 covers
// 65 k diagonals, and the pass is issued by eight warps.
#define GC_K3B_WARPS 8
#define GC_K3B_LANES (32 * GC_K3B_WARPS)
template <int NB>
__device__ __forceinline__ uint32_t gc_k3b_run_pass(const GcK3wPass& p, GcK3Block* blocksOut, uint32_t* xfer /* [2][GC_K3B_WARPS] */, int32_t* evShared)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int src = (lane + 31) & 31, prevWarp = (warp + GC_K3B_WARPS - 1) % GC_K3B_WARPS;
	GcK3wLane<NB> s;
	gc_k3w_lane_init(p, s, (int32_t)threadIdx.x);
	uint32_t send = 0;
	int32_t tau = 0;
	if (threadIdx.x < 2 * GC_K3B_WARPS) xfer[threadIdx.x] = 0;
	__syncthreads();
	while (tau <= p.tauEnd)
	{
		// next step at which some lane of the block changes its control state
		if (threadIdx.x == 0) *evShared = GC_K3W_NO_EVENT;
		__syncthreads();
		int32_t evWarp = __reduce_min_sync(0xFFFFFFFFu, gc_k3w_next_event(p, s, tau));
		if (lane == 0) atomicMin(evShared, evWarp);
		__syncthreads();
		const int32_t ev = *evShared;
		if (ev > tau)
		{
			int32_t end = ev <= p.tauEnd ? ev : p.tauEnd + 1;
			GcK3wSegment seg = gc_k3w_segment(p, s, tau);
			for (; tau < end; tau++)
			{
				uint32_t recv = __shfl_sync(0xFFFFFFFFu, send, src);
				if (lane == 0) recv = xfer[((tau + 1) & 1) * GC_K3B_WARPS + prevWarp]; // what the previous warp's last lane sent in step tau - 1
				send = gc_k3w_lane_fast_step<NB, false>(p, s, seg, tau, recv);
				if (lane == 31) xfer[(tau & 1) * GC_K3B_WARPS + warp] = send;
				__syncthreads();
			}
		}
		else
		{
			uint32_t recv = __shfl_sync(0xFFFFFFFFu, send, src);
			if (lane == 0) recv = xfer[((tau + 1) & 1) * GC_K3B_WARPS + prevWarp];
			send = gc_k3w_lane_step(p, s, tau, recv, blocksOut);
			if (lane == 31) xfer[(tau & 1) * GC_K3B_WARPS + warp] = send;
			__syncthreads();
			tau++;
		}
	}
	__syncthreads();
	return s.work;
}
__device__ __forceinline__ int gc_k3b_round_nb(int nb) { return nb <= 1 ? 1 : nb <= 2 ? 2 : nb <= 3 ? 3 : nb <= 4 ? 4 : 0; }

// one block = one edlib NW distance
__global__ void __launch_bounds__(GC_K3B_LANES, 2) gc_k3b_distance_kernel(const uint8_t* __restrict__ seq, const GcK3Desc* __restrict__ descs, uint32_t n, uint8_t* arena, GcK3Out* out)
{
	__shared__ uint32_t xfer[2 * GC_K3B_WARPS];
	__shared__ int32_t evShared;
	__shared__ unsigned long long workShared;
	const uint32_t w = blockIdx.x;
	if (w >= n) return;
	GcK3Desc d = descs[w];
	int32_t q = d.q, t = d.t;
	int32_t nb = (q + 63) / 64; if (nb < 1) nb = 1;
	uint64_t* peq = (uint64_t*)(arena + d.wsOff);
	GcK3Block* blocks = (GcK3Block*)(peq + 4 * (size_t)nb);
	const uint8_t* query = seq + d.qOff;
	const uint8_t* target = seq + d.tOff;
	for (int32_t b = threadIdx.x; b < nb; b += GC_K3B_LANES)
	{
		uint64_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
		int32_t lim = q - b * 64; if (lim > 64) lim = 64;
		for (int32_t i = 0; i < lim; i++)
		{
			uint8_t c = query[b * 64 + i];
			uint64_t bit = 1ULL << i;
			e0 |= c == 0 ? bit : 0; e1 |= c == 1 ? bit : 0; e2 |= c == 2 ? bit : 0; e3 |= c == 3 ? bit : 0;
		}
		peq[b] = e0; peq[nb + b] = e1; peq[2 * (size_t)nb + b] = e2; peq[3 * (size_t)nb + b] = e3;
	}
	if (threadIdx.x == 0) workShared = 0;
	__syncthreads();
	GcK3Out o; o.status = GC_OK; o.opsLen = 0; o.pad = 0; o.blocks = 0; o.distance = -1;
	uint32_t work = 0;
	if (q == 0 || t == 0) o.distance = q > t ? q : t;
	else
	{
		int32_t k = d.kHint < 64 ? 64 : d.kHint;
		int32_t diff = q > t ? q - t : t - q, mx = q > t ? q : t;
		while (true)
		{
			if (k >= diff)
			{
				int32_t kk = k > mx ? mx : k;
				int NB = gc_k3b_round_nb(gc_k3w_blocks_per_lane(q, t, kk, GC_K3B_LANES));
				if (NB == 0) { o.distance = GC_K3W_NEED_LARGER; o.pad = (uint32_t)k; break; }
				GcK3wPass p = gc_k3w_make_pass(peq, nb, 0, q, target, 0, 1, t, kk, t - 1, NB, GC_K3B_LANES);
				switch (NB)
				{
					case 1: work += gc_k3b_run_pass<1>(p, blocks, xfer, &evShared); break;
					case 2: work += gc_k3b_run_pass<2>(p, blocks, xfer, &evShared); break;
					case 3: work += gc_k3b_run_pass<3>(p, blocks, xfer, &evShared); break;
					default: work += gc_k3b_run_pass<4>(p, blocks, xfer, &evShared); break;
				}
				__syncthreads();
				int32_t v = gc_k3_cell(blocks[(q - 1) >> 6], q - 1);
				__syncthreads(); // every thread has read the stop column before the next pass overwrites it
				if (v <= kk) { o.distance = v; break; }
			}
			k *= 2;
		}
	}
	for (int off = 16; off > 0; off >>= 1) work += __shfl_down_sync(0xFFFFFFFFu, work, off);
	if ((threadIdx.x & 31) == 0) atomicAdd(&workShared, (unsigned long long)work);
	__syncthreads();
	if (threadIdx.x == 0) { o.blocks = workShared; out[d.resultIndex] = o; }
}

// one thread = one edlib NW path (Hirschberg + leaf tracebacks) for a known distance
__global__ void __launch_bounds__(64) gc_k3_path_kernel(const uint8_t* __restrict__ seq, const GcK3Desc* __restrict__ descs, uint32_t n, uint8_t* arena, uint8_t* opsArena, GcK3Out* out)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	GcK3Desc d = descs[t];
	GcK3Out o = out[d.resultIndex];
	int32_t nb = (d.q + 63) / 64; if (nb < 1) nb = 1;
	uint8_t* base = arena + d.wsOff;
	uint64_t* peq = (uint64_t*)base;
	uint64_t* rpeq = peq + 4 * (size_t)nb;
	GcK3Block* blocksA = (GcK3Block*)(rpeq + 4 * (size_t)nb);
	GcK3Block* blocksB = blocksA + nb;
	GcK3Block* store = blocksB + nb;
	uint32_t* colStart = (uint32_t*)(store + K3_STORE_CAP);
	GcK3Frame* stack = (GcK3Frame*)(colStart + K3_COL_CAP);
	const uint8_t* query = seq + d.qOff;
	gc_k3_build_peq(query, d.q, peq, nb);
	for (int32_t b = 0; b < 4 * nb; b++) rpeq[b] = 0;
	for (int32_t i = 0; i < d.q; i++)
	{
		uint8_t c = query[d.q - 1 - i];
		if (c < 4) rpeq[(int32_t)c * nb + (i >> 6)] |= 1ULL << (i & 63);
	}
	GcK3PathWorkspace w;
	w.peq = peq; w.rpeq = rpeq; w.nbTotal = nb; w.qTotal = d.q; w.tTotal = d.t;
	w.blocksA = blocksA; w.blocksB = blocksB; w.store = store; w.colStart = colStart; w.storeCap = K3_STORE_CAP; w.colCap = K3_COL_CAP; w.stack = stack; w.stackCap = K3_STACK_CAP;
	uint64_t work = 0;
	uint32_t nOps = 0;
	bool ok = gc_k3_path(w, seq + d.tOff, d.best, opsArena + d.opsOff, nOps, d.opsCap, work);
	o.opsLen = ok ? nOps : 0;
	if (!ok) o.status = GC_INTERNAL;
	o.blocks += work;
	out[d.resultIndex] = o;
}


// ---- warp form of the edit path (gc_k3w_path): one warp = one alignment, fetched from a shared counter
struct GcK3wDeviceExec
{
	int lane;
	__device__ __forceinline__ bool leader() const { return lane == 0; }
	__device__ __forceinline__ void sync() const { __syncwarp(); }
	__device__ __forceinline__ uint32_t fromLeader(uint32_t v) const { return __shfl_sync(0xFFFFFFFFu, v, 0); }
	__device__ __noinline__ uint64_t pass(const GcK3wPass& p, int NB, GcK3Block* out)
	{
		uint32_t w;
		switch (NB)
		{
			case 1: w = gc_k3w_run_pass<1>(p, out); break;
			case 2: w = gc_k3w_run_pass<2>(p, out); break;
			case 4: w = gc_k3w_run_pass<4>(p, out); break;
			default: w = gc_k3w_run_pass<8>(p, out); break;
		}
		for (int off = 16; off > 0; off >>= 1) w += __shfl_xor_sync(0xFFFFFFFFu, w, off);
		return w;
	}
	// first row r in [0, q-2] with left[r] + right[r+1] == best (edlib.cpp:1330-1337), -1 if none
	__device__ __forceinline__ int32_t firstSplitRow(const GcK3Block* A, int32_t lfb, int32_t llb, const GcK3Block* B, int32_t rfb, int32_t rlb, int32_t q, int32_t best) const
	{
		const int32_t INF = 1 << 29;
		for (int32_t base = 0; base <= q - 2; base += 32)
		{
			int32_t r = base + lane;
			bool ok = false;
			if (r <= q - 2)
			{
				int32_t b = r >> 6; int32_t ls = (b < lfb || b > llb) ? INF : gc_k3_cell(A[b], r);
				int32_t rr = q - 1 - (r + 1); int32_t rb = rr >> 6; int32_t rs = (rb < rfb || rb > rlb) ? INF : gc_k3_cell(B[rb], rr);
				ok = ls + rs == best;
			}
			uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
			if (m) return base + __ffs(m) - 1;
		}
		return -1;
	}
};

static const uint32_t K3W_STORE_CAP = 52432, K3W_STACK_CAP = 96;
static size_t k3wPathSlotBytes(int32_t maxQ)
{
	size_t nb = (size_t)(maxQ + 63) / 64 + 1;
	return alignUp(2 * nb * 4 * 8 + 2 * nb * sizeof(GcK3Block) + (size_t)K3W_STORE_CAP * sizeof(GcK3Block) + (size_t)K3W_STACK_CAP * sizeof(GcK3Frame) + 64, 128);
}

__global__ void __launch_bounds__(128) gc_k3w_path_kernel(const uint8_t* __restrict__ seq, const GcK3Desc* __restrict__ descs, uint32_t n, uint8_t* slots, size_t slotBytes, int32_t maxQ,
	uint8_t* opsArena, GcK3Out* out, uint32_t* counter)
{
	int lane = threadIdx.x & 31;
	uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	uint8_t* base = slots + (size_t)slot * slotBytes;
	size_t nbMax = (size_t)(maxQ + 63) / 64 + 1;
	uint64_t* peq = (uint64_t*)base;
	uint64_t* rpeq = peq + 4 * nbMax;
	GcK3Block* blocksA = (GcK3Block*)(rpeq + 4 * nbMax);
	GcK3Block* blocksB = blocksA + nbMax;
	GcK3Block* store = blocksB + nbMax;
	GcK3Frame* stack = (GcK3Frame*)(store + K3W_STORE_CAP);
	while (true)
	{
		uint32_t w = 0;
		if (lane == 0) w = atomicAdd(counter, 1u);
		w = __shfl_sync(0xFFFFFFFFu, w, 0);
		if (w >= n) return;
		GcK3Desc d = descs[w];
		int32_t q = d.q, t = d.t;
		int32_t nb = (q + 63) / 64; if (nb < 1) nb = 1;
		const uint8_t* query = seq + d.qOff;
		for (int32_t b = lane; b < nb; b += 32)
		{
			uint64_t e0 = 0, e1 = 0, e2 = 0, e3 = 0, f0 = 0, f1 = 0, f2 = 0, f3 = 0;
			int32_t lim = q - b * 64; if (lim > 64) lim = 64;
			for (int32_t i = 0; i < lim; i++)
			{
				uint8_t c = query[b * 64 + i];
				uint8_t rc = query[q - 1 - (b * 64 + i)];
				uint64_t bit = 1ULL << i;
				e0 |= c == 0 ? bit : 0; e1 |= c == 1 ? bit : 0; e2 |= c == 2 ? bit : 0; e3 |= c == 3 ? bit : 0;
				f0 |= rc == 0 ? bit : 0; f1 |= rc == 1 ? bit : 0; f2 |= rc == 2 ? bit : 0; f3 |= rc == 3 ? bit : 0;
			}
			peq[b] = e0; peq[nb + b] = e1; peq[2 * (size_t)nb + b] = e2; peq[3 * (size_t)nb + b] = e3;
			rpeq[b] = f0; rpeq[nb + b] = f1; rpeq[2 * (size_t)nb + b] = f2; rpeq[3 * (size_t)nb + b] = f3;
		}
		__syncwarp();
		GcK3wPathWorkspace ws;
		ws.peq = peq; ws.rpeq = rpeq; ws.nbTotal = nb; ws.qTotal = q; ws.tTotal = t; ws.blocksA = blocksA; ws.blocksB = blocksB;
		ws.store = store; ws.storeCap = K3W_STORE_CAP; ws.stack = stack; ws.stackCap = K3W_STACK_CAP; ws.maxNB = 8;
		GcK3wDeviceExec ex; ex.lane = lane;
		uint64_t work = 0; uint32_t nOps = 0;
		bool ok = gc_k3w_path(ex, ws, seq + d.tOff, d.best, opsArena + d.opsOff, nOps, d.opsCap, work);
		if (lane == 0)
		{
			GcK3Out o = out[d.resultIndex];
			o.opsLen = ok ? nOps : 0;
			o.pad = ok ? 0u : 1u; // 1 = the warp form gave up (band wider than its register budget): the host re-runs the item in the thread form
			o.blocks += work;
			out[d.resultIndex] = o;
		}
		__syncwarp();
	}
}

// ---- edit paths, level-parallel form.  The Hirschberg recursion of one alignment is a tree whose frames of one depth are
// independent: instead of one warp walking the tree depth-first (gc_k3w_path_kernel: 20 ms for 48 alignments, 6 % of the SMs
// busy), every recursion LEVEL is one launch with one warp per frame of every alignment.  A frame owns the slice
// [opsOff, opsOff + q + t) of the operation buffer (the slices of its two children tile it exactly), a leaf writes its
// operations at the start of its slice, and a final pass squeezes the unused bytes (0xFF) out of every alignment's slice.
struct GcK3LFrame { GcK3Frame f; uint32_t item; uint32_t pad; uint64_t opsOff; };

// Peq of the query and of the reversed query, once per alignment (warp per alignment)
__global__ void gc_k3l_peq_kernel(const uint8_t* __restrict__ seq, const GcK3Desc* __restrict__ descs, uint32_t n, uint8_t* itemWs)
{
	uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int lane = threadIdx.x & 31;
	if (w >= n) return;
	GcK3Desc d = descs[w];
	int32_t q = d.q;
	int32_t nb = (q + 63) / 64; if (nb < 1) nb = 1;
	uint64_t* peq = (uint64_t*)(itemWs + d.wsOff);
	uint64_t* rpeq = peq + 4 * (size_t)nb;
	const uint8_t* query = seq + d.qOff;
	for (int32_t b = lane; b < nb; b += 32)
	{
		uint64_t e0 = 0, e1 = 0, e2 = 0, e3 = 0, f0 = 0, f1 = 0, f2 = 0, f3 = 0;
		int32_t lim = q - b * 64; if (lim > 64) lim = 64;
		for (int32_t i = 0; i < lim; i++)
		{
			uint8_t c = query[b * 64 + i];
			uint8_t rc = query[q - 1 - (b * 64 + i)];
			uint64_t bit = 1ULL << i;
			e0 |= c == 0 ? bit : 0; e1 |= c == 1 ? bit : 0; e2 |= c == 2 ? bit : 0; e3 |= c == 3 ? bit : 0;
			f0 |= rc == 0 ? bit : 0; f1 |= rc == 1 ? bit : 0; f2 |= rc == 2 ? bit : 0; f3 |= rc == 3 ? bit : 0;
		}
		peq[b] = e0; peq[nb + b] = e1; peq[2 * (size_t)nb + b] = e2; peq[3 * (size_t)nb + b] = e3;
		rpeq[b] = f0; rpeq[nb + b] = f1; rpeq[2 * (size_t)nb + b] = f2; rpeq[3 * (size_t)nb + b] = f3;
	}
}

// one recursion level: persistent warps take frames from a counter; children go to the next level's frame list
__global__ void __launch_bounds__(128) gc_k3l_level_kernel(const uint8_t* __restrict__ seq, const GcK3Desc* __restrict__ descs, const GcK3LFrame* __restrict__ in, uint32_t nIn,
	GcK3LFrame* outFrames, uint32_t outCap, uint32_t* counters, uint8_t* slots, size_t slotBytes, int32_t maxQ, const uint8_t* itemWs, uint8_t* opsArena, GcK3Out* out)
{
	int lane = threadIdx.x & 31;
	uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	uint8_t* base = slots + (size_t)slot * slotBytes;
	size_t nbMax = (size_t)(maxQ + 63) / 64 + 1;
	GcK3Block* blocksA = (GcK3Block*)base;
	GcK3Block* blocksB = blocksA + nbMax;
	GcK3Block* store = blocksB + nbMax;
	while (true)
	{
		uint32_t w = 0;
		if (lane == 0) w = atomicAdd(&counters[0], 1u);
		w = __shfl_sync(0xFFFFFFFFu, w, 0);
		if (w >= nIn) return;
		GcK3LFrame fr = in[w];
		GcK3Desc d = descs[fr.item];
		int32_t nb = (d.q + 63) / 64; if (nb < 1) nb = 1;
		GcK3wPathWorkspace ws;
		ws.peq = (const uint64_t*)(itemWs + d.wsOff); ws.rpeq = ws.peq + 4 * (size_t)nb; ws.nbTotal = nb; ws.qTotal = d.q; ws.tTotal = d.t;
		ws.blocksA = blocksA; ws.blocksB = blocksB; ws.store = store; ws.storeCap = K3W_STORE_CAP; ws.stack = nullptr; ws.stackCap = 0; ws.maxNB = 8;
		GcK3wDeviceExec ex; ex.lane = lane;
		uint64_t work = 0; uint32_t nOps = 0;
		GcK3Frame ch[2]; int32_t nch = 0;
		bool ok = gc_k3w_path_frame(ex, ws, seq + d.tOff, fr.f, opsArena + fr.opsOff, nOps, (uint32_t)(fr.f.q + fr.f.t), work, ch, nch);
		if (lane == 0)
		{
			atomicAdd((unsigned long long*)&out[d.resultIndex].blocks, (unsigned long long)work);
			if (!ok) out[d.resultIndex].pad = 1u; // the alignment is redone by the depth-first kernel / the thread form
			else if (nch)
			{
				uint32_t idx = atomicAdd(&counters[1], 2u);
				if (idx + 2 > outCap) out[d.resultIndex].pad = 1u;
				else
				{
					GcK3LFrame a; a.f = ch[0]; a.item = fr.item; a.pad = 0; a.opsOff = fr.opsOff;
					GcK3LFrame b; b.f = ch[1]; b.item = fr.item; b.pad = 0; b.opsOff = fr.opsOff + (uint64_t)(ch[0].q + ch[0].t);
					outFrames[idx] = a; outFrames[idx + 1] = b;
				}
			}
		}
		__syncwarp();
	}
}

// squeeze the 0xFF filler out of every alignment's slice (warp per alignment, in place: the write position never passes the read position)
__global__ void gc_k3l_compact_kernel(const GcK3Desc* __restrict__ descs, uint32_t n, uint8_t* opsArena, GcK3Out* out)
{
	uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	uint32_t lane = threadIdx.x & 31;
	if (w >= n) return;
	GcK3Desc d = descs[w];
	if (out[d.resultIndex].pad == 1u) return;
	uint8_t* base = opsArena + d.opsOff;
	uint32_t cap = (uint32_t)(d.q + d.t), dst = 0;
	for (uint32_t tile = 0; tile < cap; tile += 32)
	{
		uint32_t i = tile + lane;
		uint8_t c = i < cap ? base[i] : (uint8_t)0xFF;
		bool valid = c != 0xFF;
		uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
		if (valid) base[dst + __popc(m & ((1u << lane) - 1u))] = c;
		dst += __popc(m);
		__syncwarp();
	}
	if (lane == 0) out[d.resultIndex].opsLen = dst;
}

struct GcByteCopyDesc { uint64_t src; uint64_t dst; uint32_t len; uint32_t pad; };
__global__ void gc_bytes_gather_kernel(const GcByteCopyDesc* __restrict__ descs, uint32_t n, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst)
{
	uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	uint32_t lane = threadIdx.x & 31;
	if (warp >= n) return;
	GcByteCopyDesc d = descs[warp];
	for (uint32_t i = lane; i < d.len; i += 32) dst[d.dst + i] = src[d.src + i];
}

extern "C" int gcgpu_nw(gcgpu_ctx* ctx, const char* seqs, uint64_t seq_bytes, const gcgpu_nw_item* items, uint32_t n,
	gcgpu_nw_result* results, uint8_t* ops, uint64_t ops_capacity, uint64_t* ops_used)
{
	if (!ctx || (!items && n) || (!results && n) || !ops_used) return setError(GCGPU_ERR_ARG, "gcgpu_nw: null argument");
	*ops_used = 0;
	ctx->lastKernelMs = 0;
	if (n == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	for (uint32_t i = 0; i < n; i++)
	{
		const gcgpu_nw_item& it = items[i];
		if (it.query_len < 0 || it.target_len < 0 || it.query_offset + (uint64_t)it.query_len > seq_bytes || it.target_offset + (uint64_t)it.target_len > seq_bytes)
			return setError(GCGPU_ERR_ARG, "gcgpu_nw: item " + std::to_string(i) + " out of range");
	}
	float ms = 0;
	// GCGPU_K3_FORCE=thread | dfs: run every alignment through a FALLBACK form (the thread-per-alignment kernels that take over
	// when a band outgrows the warp kernels' register budget; the depth-first warp kernel that takes over from the level-parallel
	// edit-path kernels) -- how the parity tests reach those kernels
	const char* k3Force = getenv("GCGPU_K3_FORCE");
	const bool forceThread = k3Force && !strcmp(k3Force, "thread"), forceDfs = k3Force && !strcmp(k3Force, "dfs");
	if (seqs)
	{
		CUDA_TRY(ctx->nwSeqBuf.ensure(seq_bytes + 16));
		if (seq_bytes) CUDA_TRY(gcCopy(ctx, ctx->nwSeqBuf.p, seqs, seq_bytes, cudaMemcpyHostToDevice, ctx->stream));
		ctx->nwResident = seq_bytes;
	}
	else if (ctx->nwResident != seq_bytes) return setError(GCGPU_ERR_ARG, "gcgpu_nw: seqs == NULL but no sequence buffer of this size is resident");
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream)); // kernel time: the host-to-device copy above is not part of it
	if (seqs && seq_bytes)
	{
		gc_k3_encode_kernel<<<1184, 256, 0, ctx->stream>>>((uint8_t*)ctx->nwSeqBuf.p, seq_bytes);
		ctx->launches++;
	}
	// ---- distance pass, longest first
	std::vector<uint32_t> order(n);
	for (uint32_t i = 0; i < n; i++) order[i] = i;
	std::sort(order.begin(), order.end(), [items](uint32_t a, uint32_t b) { int64_t wa = (int64_t)items[a].query_len + items[a].target_len, wb = (int64_t)items[b].query_len + items[b].target_len; return wa != wb ? wa > wb : a < b; });
	std::vector<GcK3Desc> descs(n);
	size_t wsTotal = 0;
	for (uint32_t k = 0; k < n; k++)
	{
		const gcgpu_nw_item& it = items[order[k]];
		GcK3Desc& d = descs[k];
		d.qOff = it.query_offset; d.tOff = it.target_offset; d.q = it.query_len; d.t = it.target_len; d.kHint = it.k_hint; d.best = 0;
		d.resultIndex = order[k]; d.opsCap = 0; d.opsOff = 0;
		d.wsOff = wsTotal;
		wsTotal += k3DistWorkspaceBytes(it.query_len);
	}
	CUDA_TRY(ctx->arena.ensure(wsTotal));
	CUDA_TRY(ctx->descBuf.ensure((size_t)n * sizeof(GcK3Desc)));
	CUDA_TRY(ctx->resBuf.ensure((size_t)n * sizeof(GcK3Out)));
	// want_path == 2 on every item: the caller already holds the exact distances (an earlier call on the same pairs), the
	// distance pass would only recompute them
	bool distancesKnown = true;
	for (uint32_t i = 0; i < n; i++) if (items[i].want_path != 2 || items[i].k_hint < 0) { distancesKnown = false; break; }
	std::vector<GcK3Out> hout(n);
	if (distancesKnown)
	{
		for (uint32_t i = 0; i < n; i++) { hout[i].status = GC_OK; hout[i].distance = items[i].k_hint; hout[i].opsLen = 0; hout[i].pad = 0; hout[i].blocks = 0; }
		CUDA_TRY(gcCopy(ctx, ctx->resBuf.p, hout.data(), (size_t)n * sizeof(GcK3Out), cudaMemcpyHostToDevice, ctx->stream));
		CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
		CUDA_TRY(gcSyncStream(ctx));
		CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		ctx->lastKernelMs += ms;
	}
	else
	{
	// items whose first cutoff already needs more than two 64-row blocks per lane belong to the wide class: they are few and each
	// is one long dependent pass, so their kernel runs beside the narrow class on a second stream instead of after it
	uint32_t nWide = 0;
	{
		std::vector<GcK3Desc> wide, narrow;
		for (const GcK3Desc& d : descs)
		{
			bool isWide = false;
			if (d.q > 0 && d.t > 0)
			{
				int32_t k = d.kHint < 64 ? 64 : d.kHint;
				int32_t diff = d.q > d.t ? d.q - d.t : d.t - d.q, mx = d.q > d.t ? d.q : d.t;
				while (k < diff) k *= 2;
				int32_t kk = k > mx ? mx : k;
				int32_t nbl = gc_k3w_blocks_per_lane(d.q, d.t, kk);
				isWide = nbl > 2; // more than two blocks per lane of a single warp: the block form (eight warps share the band)
			}
			(isWide ? wide : narrow).push_back(d);
		}
		nWide = (uint32_t)wide.size();
		descs.clear();
		descs.insert(descs.end(), wide.begin(), wide.end());
		descs.insert(descs.end(), narrow.begin(), narrow.end());
	}
	CUDA_TRY(gcCopy(ctx, ctx->descBuf.p, descs.data(), (size_t)n * sizeof(GcK3Desc), cudaMemcpyHostToDevice, ctx->stream));
	if (forceThread)
	{
		gc_k3_distance_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>((const uint8_t*)ctx->nwSeqBuf.p, (const GcK3Desc*)ctx->descBuf.p, n, (uint8_t*)ctx->arena.p, (GcK3Out*)ctx->resBuf.p);
		ctx->launches++;
		nWide = 0;
	}
	else
	{
	if (nWide)
	{
		CUDA_TRY(cudaEventRecord(ctx->evFork, ctx->stream));
		CUDA_TRY(cudaStreamWaitEvent(ctx->stream2, ctx->evFork, 0));
		gc_k3b_distance_kernel<<<nWide, GC_K3B_LANES, 0, ctx->stream2>>>((const uint8_t*)ctx->nwSeqBuf.p, (const GcK3Desc*)ctx->descBuf.p, nWide, (uint8_t*)ctx->arena.p, (GcK3Out*)ctx->resBuf.p);
		CUDA_TRY(cudaEventRecord(ctx->evJoin, ctx->stream2));
		ctx->launches++;
	}
	if (n > nWide)
	{
		gc_k3w_distance_kernel<<<(n - nWide + 3) / 4, 128, 0, ctx->stream>>>((const uint8_t*)ctx->nwSeqBuf.p, (const GcK3Desc*)ctx->descBuf.p + nWide, n - nWide, (uint8_t*)ctx->arena.p, (GcK3Out*)ctx->resBuf.p);
		ctx->launches++;
	}
	if (nWide) CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->evJoin, 0));
	}
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	CUDA_TRY(gcCopy(ctx, hout.data(), ctx->resBuf.p, (size_t)n * sizeof(GcK3Out), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->lastKernelMs += ms;
	GC_TRACE_MS("k3w distance (warp | block form)", n);
	}
	// items whose cutoff band outgrew the register budget of the launched class: wider class, then the thread form
	for (int cls = 1; cls <= 2; cls++)
	{
		std::vector<GcK3Desc> rd;
		for (uint32_t k = 0; k < n; k++)
		{
			const GcK3Out& o = hout[descs[k].resultIndex];
			if (o.distance == GC_K3W_NEED_LARGER) { GcK3Desc d = descs[k]; d.kHint = (int32_t)o.pad; rd.push_back(d); }
		}
		if (rd.empty()) break;
		uint32_t m = (uint32_t)rd.size();
		uint64_t blocksBefore = 0; // work of the aborted attempts is kept in the item's counter
		(void)blocksBefore;
		CUDA_TRY(gcCopy(ctx, ctx->descBuf.p, rd.data(), (size_t)m * sizeof(GcK3Desc), cudaMemcpyHostToDevice, ctx->stream));
		CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
		if (cls == 1) gc_k3b_distance_kernel<<<m, GC_K3B_LANES, 0, ctx->stream>>>((const uint8_t*)ctx->nwSeqBuf.p, (const GcK3Desc*)ctx->descBuf.p, m, (uint8_t*)ctx->arena.p, (GcK3Out*)ctx->resBuf.p);
		else gc_k3_distance_kernel<<<(m + 63) / 64, 64, 0, ctx->stream>>>((const uint8_t*)ctx->nwSeqBuf.p, (const GcK3Desc*)ctx->descBuf.p, m, (uint8_t*)ctx->arena.p, (GcK3Out*)ctx->resBuf.p);
		ctx->launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
		std::vector<GcK3Out> prev = hout;
		CUDA_TRY(gcCopy(ctx, hout.data(), ctx->resBuf.p, (size_t)n * sizeof(GcK3Out), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(gcSyncStream(ctx));
		CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		ctx->lastKernelMs += ms;
		GC_TRACE_MS(cls == 1 ? "k3b distance (block form)" : "k3 distance (thread)", m);
		for (const GcK3Desc& d : rd) hout[d.resultIndex].blocks += prev[d.resultIndex].blocks;
	}
	// ---- path pass for the items that asked for it
	std::vector<uint32_t> want;
	for (uint32_t k = 0; k < n; k++) if (items[order[k]].want_path && items[order[k]].query_len > 0 && items[order[k]].target_len > 0) want.push_back(order[k]);
	std::vector<uint64_t> opsOffOfItem(n, 0);
	if (!want.empty())
	{
		std::vector<GcK3Desc> pd(want.size());
		uint64_t opsTotal = 0;
		int32_t maxQ = 1;
		for (size_t k = 0; k < want.size(); k++)
		{
			const gcgpu_nw_item& it = items[want[k]];
			GcK3Desc& d = pd[k];
			d.qOff = it.query_offset; d.tOff = it.target_offset; d.q = it.query_len; d.t = it.target_len; d.kHint = 0; d.best = hout[want[k]].distance;
			d.resultIndex = want[k];
			d.opsCap = (uint32_t)(it.query_len + it.target_len + 8);
			d.opsOff = opsTotal; opsOffOfItem[want[k]] = opsTotal; opsTotal += d.opsCap;
			d.wsOff = 0;
			if (it.query_len > maxQ) maxQ = it.query_len;
		}
		uint32_t m = (uint32_t)pd.size();
		CUDA_TRY(ctx->traceArena.ensure(opsTotal + 16));
		const bool levelForm = !forceThread && !forceDfs;
		std::vector<GcK3Desc> dfs; // alignments for the depth-first kernel: all of them, or the ones the level form gave up on
		if (levelForm)
		{
			// per alignment: Peq + reversed Peq; per resident warp: two block columns + the leaf store
			size_t itemWsTotal = 0; uint64_t frameCap = 4096;
			for (GcK3Desc& d : pd)
			{
				size_t nb = (size_t)(d.q + 63) / 64 + 1;
				d.wsOff = itemWsTotal; itemWsTotal += alignUp(2 * 4 * nb * 8, 128);
				uint64_t leaves = ((uint64_t)nb * (uint64_t)d.t) / (K3W_STORE_CAP / 4) + 2, p2 = 1;
				while (p2 < leaves) p2 <<= 1;
				frameCap += 4 * p2 + 8;
			}
			size_t nbMax = (size_t)(maxQ + 63) / 64 + 1;
			size_t slotBytes = alignUp(2 * nbMax * sizeof(GcK3Block) + (size_t)K3W_STORE_CAP * sizeof(GcK3Block) + 64, 128);
			const uint32_t ctasMax = 148 * 2;
			size_t slotsTotal = slotBytes * ctasMax * 4;
			CUDA_TRY(ctx->arena.ensure(slotsTotal + 256 + itemWsTotal));
			uint32_t* counters = (uint32_t*)((uint8_t*)ctx->arena.p + slotsTotal);
			uint8_t* itemWs = (uint8_t*)ctx->arena.p + slotsTotal + 256;
			CUDA_TRY(ctx->descBuf.ensure(pd.size() * sizeof(GcK3Desc)));
			CUDA_TRY(ctx->copyDesc.ensure(2 * frameCap * sizeof(GcK3LFrame)));
			GcK3LFrame* frames[2] = { (GcK3LFrame*)ctx->copyDesc.p, (GcK3LFrame*)ctx->copyDesc.p + frameCap };
			std::vector<GcK3LFrame> f0(m);
			for (uint32_t k = 0; k < m; k++) { f0[k].f.qOff = 0; f0[k].f.q = pd[k].q; f0[k].f.tOff = 0; f0[k].f.t = pd[k].t; f0[k].f.best = pd[k].best; f0[k].item = k; f0[k].pad = 0; f0[k].opsOff = pd[k].opsOff; }
			CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
			CUDA_TRY(gcCopy(ctx, ctx->descBuf.p, pd.data(), pd.size() * sizeof(GcK3Desc), cudaMemcpyHostToDevice, ctx->stream));
			CUDA_TRY(gcCopy(ctx, frames[0], f0.data(), (size_t)m * sizeof(GcK3LFrame), cudaMemcpyHostToDevice, ctx->stream));
			CUDA_TRY(cudaMemsetAsync(ctx->traceArena.p, 0xFF, opsTotal, ctx->stream));
			gc_k3l_peq_kernel<<<(m + 3) / 4, 128, 0, ctx->stream>>>((const uint8_t*)ctx->nwSeqBuf.p, (const GcK3Desc*)ctx->descBuf.p, m, itemWs);
			ctx->launches++;
			uint32_t nIn = m; int cur = 0, levels = 0;
			while (nIn > 0)
			{
				if (++levels > 64) return setError(GCGPU_ERR_INTERNAL, "gcgpu_nw: edit path recursion deeper than 64 levels");
				CUDA_TRY(cudaMemsetAsync(counters, 0, 8, ctx->stream));
				uint32_t ctas = std::min<uint32_t>((nIn + 3) / 4, ctasMax);
				gc_k3l_level_kernel<<<ctas, 128, 0, ctx->stream>>>((const uint8_t*)ctx->nwSeqBuf.p, (const GcK3Desc*)ctx->descBuf.p, frames[cur], nIn, frames[cur ^ 1], (uint32_t)frameCap, counters,
					(uint8_t*)ctx->arena.p, slotBytes, maxQ, itemWs, (uint8_t*)ctx->traceArena.p, (GcK3Out*)ctx->resBuf.p);
				ctx->launches++;
				CUDA_TRY(cudaGetLastError());
				uint32_t cnt[2] = { 0, 0 };
				CUDA_TRY(gcCopy(ctx, cnt, counters, 8, cudaMemcpyDeviceToHost, ctx->stream));
				CUDA_TRY(gcSyncStream(ctx));
				nIn = std::min<uint64_t>(cnt[1], frameCap);
				cur ^= 1;
			}
			gc_k3l_compact_kernel<<<(m + 3) / 4, 128, 0, ctx->stream>>>((const GcK3Desc*)ctx->descBuf.p, m, (uint8_t*)ctx->traceArena.p, (GcK3Out*)ctx->resBuf.p);
			ctx->launches++;
			CUDA_TRY(cudaGetLastError());
			CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
			CUDA_TRY(gcCopy(ctx, hout.data(), ctx->resBuf.p, (size_t)n * sizeof(GcK3Out), cudaMemcpyDeviceToHost, ctx->stream));
			CUDA_TRY(gcSyncStream(ctx));
			CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
			ctx->lastKernelMs += ms;
			if (g_trace) fprintf(stderr, "[gcgpu] %-28s n=%-8u %.3f ms (%d levels)\n", "k3 path, level-parallel", (unsigned)m, ms, levels);
			for (const GcK3Desc& d : pd) if (hout[d.resultIndex].pad == 1) { GcK3Desc e = d; e.wsOff = 0; dfs.push_back(e); }
		}
		else dfs = pd;
		if (forceThread) for (const GcK3Desc& d : dfs) hout[d.resultIndex].pad = 1; // straight to the thread form below
		if (!dfs.empty() && !forceThread)
		{
		m = (uint32_t)dfs.size();
		for (const GcK3Desc& d : dfs) { hout[d.resultIndex].pad = 0; hout[d.resultIndex].opsLen = 0; }
		CUDA_TRY(gcCopy(ctx, ctx->resBuf.p, hout.data(), (size_t)n * sizeof(GcK3Out), cudaMemcpyHostToDevice, ctx->stream));
		// persistent warps: one workspace slot per resident warp, items fetched from a counter (longest first)
		uint32_t ctas = std::min<uint32_t>((m + 3) / 4, 148 * 4);
		size_t slotBytes = k3wPathSlotBytes(maxQ);
		size_t slotsTotal = slotBytes * ctas * 4;
		CUDA_TRY(ctx->arena.ensure(slotsTotal + 256));
		CUDA_TRY(ctx->descBuf.ensure(dfs.size() * sizeof(GcK3Desc)));
		CUDA_TRY(gcCopy(ctx, ctx->descBuf.p, dfs.data(), dfs.size() * sizeof(GcK3Desc), cudaMemcpyHostToDevice, ctx->stream));
		uint32_t* counter = (uint32_t*)((uint8_t*)ctx->arena.p + slotsTotal);
		CUDA_TRY(cudaMemsetAsync(counter, 0, 4, ctx->stream));
		CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
		gc_k3w_path_kernel<<<ctas, 128, 0, ctx->stream>>>((const uint8_t*)ctx->nwSeqBuf.p, (const GcK3Desc*)ctx->descBuf.p, m, (uint8_t*)ctx->arena.p, slotBytes, maxQ, (uint8_t*)ctx->traceArena.p, (GcK3Out*)ctx->resBuf.p, counter);
		ctx->launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
		CUDA_TRY(gcCopy(ctx, hout.data(), ctx->resBuf.p, (size_t)n * sizeof(GcK3Out), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(gcSyncStream(ctx));
		CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		ctx->lastKernelMs += ms;
		GC_TRACE_MS("k3w path", m);
		}
		// items the warp form gave up on (band beyond its register budget): thread form
		std::vector<GcK3Desc> rd;
		for (const GcK3Desc& d : dfs) if (hout[d.resultIndex].pad == 1) rd.push_back(d);
		if (!rd.empty())
		{
			size_t ws = 0;
			for (GcK3Desc& d : rd) { d.wsOff = ws; ws += k3PathWorkspaceBytes(d.q); }
			CUDA_TRY(ctx->arena.ensure(ws));
			CUDA_TRY(gcCopy(ctx, ctx->descBuf.p, rd.data(), rd.size() * sizeof(GcK3Desc), cudaMemcpyHostToDevice, ctx->stream));
			uint32_t m2 = (uint32_t)rd.size();
			CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
			gc_k3_path_kernel<<<(m2 + 63) / 64, 64, 0, ctx->stream>>>((const uint8_t*)ctx->nwSeqBuf.p, (const GcK3Desc*)ctx->descBuf.p, m2, (uint8_t*)ctx->arena.p, (uint8_t*)ctx->traceArena.p, (GcK3Out*)ctx->resBuf.p);
			ctx->launches++;
			CUDA_TRY(cudaGetLastError());
			CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
			CUDA_TRY(gcCopy(ctx, hout.data(), ctx->resBuf.p, (size_t)n * sizeof(GcK3Out), cudaMemcpyDeviceToHost, ctx->stream));
			CUDA_TRY(gcSyncStream(ctx));
			CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
			ctx->lastKernelMs += ms;
		}
	}
	uint64_t used = 0;
	bool internal = false;
	std::vector<GcByteCopyDesc> copies;
	for (uint32_t i = 0; i < n; i++)
	{
		results[i].status = hout[i].status == GC_OK ? 0 : GCGPU_ITEM_INTERNAL;
		if (results[i].status) internal = true;
		results[i].distance = hout[i].distance;
		results[i].ops_len = hout[i].opsLen;
		results[i].reserved = 0;
		results[i].ops_offset = used;
		results[i].blocks = hout[i].blocks;
		if (hout[i].opsLen)
		{
			GcByteCopyDesc c; c.src = opsOffOfItem[i]; c.dst = used; c.len = hout[i].opsLen; c.pad = 0;
			copies.push_back(c);
			used += hout[i].opsLen;
		}
	}
	*ops_used = used;
	if (used > ops_capacity) return setError(GCGPU_ERR_ARG, "gcgpu_nw: ops buffer too small, need " + std::to_string(used) + " bytes");
	if (used)
	{
		if (!ops) return setError(GCGPU_ERR_ARG, "gcgpu_nw: null ops buffer");
		CUDA_TRY(ctx->compact.ensure(used + 16));
		CUDA_TRY(ctx->copyDesc.ensure(copies.size() * sizeof(GcByteCopyDesc)));
		CUDA_TRY(gcCopy(ctx, ctx->copyDesc.p, copies.data(), copies.size() * sizeof(GcByteCopyDesc), cudaMemcpyHostToDevice, ctx->stream));
		uint32_t m = (uint32_t)copies.size();
		gc_bytes_gather_kernel<<<(m + 3) / 4, 128, 0, ctx->stream>>>((const GcByteCopyDesc*)ctx->copyDesc.p, m, (const uint8_t*)ctx->traceArena.p, (uint8_t*)ctx->compact.p);
		ctx->launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(gcCopy(ctx, ops, ctx->compact.p, used, cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(gcSyncStream(ctx));
	}
	if (internal) return setError(GCGPU_ERR_INTERNAL, "gcgpu_nw: an alignment path could not be reconstructed");
	return GCGPU_OK;
}

// ------------------------------------------------------------------ K2 kernels
// One block = one read.  Two forms of the same closed form (gc_k2.cuh):
//
// (1) sweep form (the default).  The reference keeps, per MPC path k, search trees over the anchors that END on path k, and asks
//     them, for every backward link (v, k) of an anchor's start node, for the best anchor ending on k at or before v
//     (AlignmentGraph.cpp:1777-1846).  Here every (anchor, path through its end node) pair of the whole batch becomes one entry
//     keyed (read, global path id, topological index of the end node); ONE device radix sort (CUB) puts the entries of a
//     (read, path) next to each other in path order, and a Fenwick max-tree lies over each such segment.  The block sweeps the
//     read's anchors by increasing read end y: an anchor is inserted into the trees of its paths as soon as it can no longer
//     overlap the anchors being evaluated (y_i <= x_j - 1: fragments have one length, so x_j grows with the sweep); an anchor
//     being evaluated asks, per link (v, k), for the prefix maximum of segment (read, k) up to topo(v) -- two binary searches
//     and O(log n) tree steps -- and, per path through its own start node, for the prefix up to that node itself (anchors that
//     end on the start node).  The few anchors still overlapping it (x_j <= y_i < y_j: the I-type term) are tested directly.
//     Work per read: O(N (K log N + overlap)) instead of the O(N^2 K^2) of form (2).
// (2) pairwise form: every earlier anchor is tested against every later one.  Used for reads whose anchors do not all have the
//     same length (a caller of gcgpu_chain may pass anything) and when the key fields would overflow; GCGPU_K2_FORCE=pairwise.
#define GC_K2_THREADS 128
#define GC_K2_PAIRWISE_MAX 384   // c2: ~145 anchors per 10-kb read -> pairwise (1.3 ms per 839 reads vs 6.4 ms swept); c4 / c5: 1-3 k anchors -> sweep
#define GC_K2_READ_BITS 18
#define GC_K2_PATH_BITS 18
#define GC_K2_TOPO_BITS 28
__device__ __forceinline__ uint64_t gc_k2_entry_key(uint32_t read, uint32_t gpath, uint32_t topo) { return ((uint64_t)read << (GC_K2_PATH_BITS + GC_K2_TOPO_BITS)) | ((uint64_t)gpath << GC_K2_TOPO_BITS) | topo; }

struct GcK2Sweep
{
	const uint64_t* pairStart;   // [total + 1] entries of anchor a (batch-global index): [pairStart[a], pairStart[a + 1])
	const uint32_t* posOf;       // [entries] sorted position of entry q
	const uint64_t* keys;        // [entries] sorted keys
	const uint32_t* segLo;       // [entries] first / one-past-last sorted position of the segment (read, path) an entry lies in
	const uint32_t* segHi;
	long long* tree;             // [entries] Fenwick max-trees over the segments, keys gc_k2_key(score, anchor)
	uint64_t entries;
};

__global__ void gc_k2_pair_count_kernel(GcMpcView m, const GcAnchor* __restrict__ a, uint64_t total, uint64_t* __restrict__ cnt)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i > total) return;
	if (i == total) { cnt[i] = 0; return; }
	uint32_t e = a[i].endNode;
	uint32_t gidx = m.compStart[m.compMap[e]] + m.compIdx[e];
	cnt[i] = m.pathsStart[gidx + 1] - m.pathsStart[gidx];
}
__global__ void gc_k2_pair_emit_kernel(GcMpcView m, const GcAnchor* __restrict__ a, const uint64_t* __restrict__ readOffsets, uint32_t numReads, uint64_t total, const uint64_t* __restrict__ pairStart,
	uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= total) return;
	uint32_t lo = 0, hi = numReads; // read of anchor i: last r with readOffsets[r] <= i
	while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (readOffsets[mid] <= i) lo = mid; else hi = mid; }
	uint32_t e = a[i].endNode, c = m.compMap[e];
	uint32_t gidx = m.compStart[c] + m.compIdx[e];
	uint32_t topo = m.topoIds[gidx];
	uint64_t q = pairStart[i];
	for (uint32_t p = m.pathsStart[gidx]; p < m.pathsStart[gidx + 1]; p++, q++) { keys[q] = gc_k2_entry_key(lo, m.pathBase[c] + m.pathsK[p], topo); vals[q] = (uint32_t)q; }
}
__global__ void gc_k2_segments_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ sortedQ, uint64_t entries, uint32_t* __restrict__ posOf, uint32_t* __restrict__ segLo, uint32_t* __restrict__ segHi)
{
	uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= entries) return;
	posOf[sortedQ[e]] = (uint32_t)e;
	const uint64_t seg = keys[e] >> GC_K2_TOPO_BITS;
	uint64_t lo = 0, hi = e; // first position whose (read, path) equals this one's
	while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if ((keys[mid] >> GC_K2_TOPO_BITS) < seg) lo = mid + 1; else hi = mid; }
	segLo[e] = (uint32_t)lo;
	lo = e + 1; hi = entries;
	while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if ((keys[mid] >> GC_K2_TOPO_BITS) <= seg) lo = mid + 1; else hi = mid; }
	segHi[e] = (uint32_t)lo;
}

// block-wide maximum of `best`, returned to every thread
__device__ __forceinline__ long long gc_k2_block_max(long long best, long long* warpBest)
{
	for (int off = 16; off > 0; off >>= 1) { long long o = __shfl_down_sync(0xFFFFFFFFu, best, off); if (o > best) best = o; }
	__syncthreads(); // the previous round's readers are done with warpBest
	if ((threadIdx.x & 31) == 0) warpBest[threadIdx.x >> 5] = best;
	__syncthreads();
	long long b = warpBest[0];
	for (int w = 1; w < GC_K2_THREADS / 32; w++) if (warpBest[w] > b) b = warpBest[w];
	return b;
}

template <bool SWEEP>
__global__ void __launch_bounds__(GC_K2_THREADS) gc_k2_chain_kernel(GcMpcView m, const GcAnchor* __restrict__ anchors, const uint64_t* __restrict__ readOffsets, uint32_t numReads, GcK2Sweep sw,
	uint32_t* order, int32_t* score, int32_t* pred, uint32_t* chain, uint32_t* chainLen, int64_t* chainScore)
{
	uint32_t r = blockIdx.x;
	if (r >= numReads) return;
	uint64_t base = readOffsets[r];
	uint32_t n = (uint32_t)(readOffsets[r + 1] - base);
	const GcAnchor* a = anchors + base;
	uint32_t* ord = order + base;
	int32_t* sc = score + base;
	int32_t* pr = pred + base;
	__shared__ long long warpBest[GC_K2_THREADS / 32];
	__shared__ int sUniform;
	// rank sort by (y, index)
	if (threadIdx.x == 0) sUniform = 1;
	__syncthreads();
	for (uint32_t j = threadIdx.x; j < n; j += blockDim.x)
	{
		uint32_t rank = 0;
		int32_t yj = a[j].y;
		for (uint32_t i = 0; i < n; i++) { int32_t yi = a[i].y; rank += (yi < yj || (yi == yj && i < j)) ? 1 : 0; }
		ord[rank] = j;
		if (a[j].y - a[j].x != a[0].y - a[0].x) sUniform = 0;
	}
	__syncthreads();
	// few anchors: testing every earlier anchor directly is cheaper than the per-anchor tree queries and barriers of the sweep
	if (SWEEP && sUniform && n > GC_K2_PAIRWISE_MAX)
	{
		const int32_t len = a[0].y - a[0].x + 1;
		// entries of this read in the sorted array
		uint64_t eLo, eHi;
		{
			uint64_t lo = 0, hi = sw.entries, want = (uint64_t)r << (GC_K2_PATH_BITS + GC_K2_TOPO_BITS);
			while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (sw.keys[mid] < want) lo = mid + 1; else hi = mid; }
			eLo = lo; hi = sw.entries; want = (uint64_t)(r + 1) << (GC_K2_PATH_BITS + GC_K2_TOPO_BITS);
			while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (sw.keys[mid] < want) lo = mid + 1; else hi = mid; }
			eHi = lo;
		}
		// prefix maximum of segment (r, gpath) over the entries with topological index <= topo
		auto query = [&](uint32_t gpath, uint32_t topo) -> long long
		{
			uint64_t lo = eLo, hi = eHi, want = gc_k2_entry_key(r, gpath, 0);
			while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (sw.keys[mid] < want) lo = mid + 1; else hi = mid; }
			const uint64_t s0 = lo;
			hi = eHi; want = gc_k2_entry_key(r, gpath, topo);
			while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (sw.keys[mid] <= want) lo = mid + 1; else hi = mid; }
			long long best = (long long)0x8000000000000000LL;
			for (uint64_t idx = lo - s0; idx > 0; idx -= idx & (~idx + 1)) { long long v = sw.tree[s0 + idx - 1]; if (v > best) best = v; }
			return best;
		};
		uint32_t ins = 0; // ord[0 .. ins) are in the trees
		for (uint32_t oj = 0; oj < n; oj++)
		{
			const uint32_t j = ord[oj];
			const GcAnchor aj = a[j];
			// ---- anchors that can no longer overlap the one being evaluated go into the trees of their paths
			uint32_t insEnd = ins;
			while (insEnd < oj && a[ord[insEnd]].y <= aj.x - 1) insEnd++;
			if (insEnd > ins)
			{
				for (uint32_t oi = ins + threadIdx.x; oi < insEnd; oi += blockDim.x)
				{
					const uint32_t i = ord[oi];
					const long long key = gc_k2_key(sc[i], (int32_t)i);
					for (uint64_t q = sw.pairStart[base + i]; q < sw.pairStart[base + i + 1]; q++)
					{
						const uint32_t e = sw.posOf[q], s0 = sw.segLo[e], segLen = sw.segHi[e] - s0;
						for (uint32_t idx = e - s0 + 1; idx <= segLen; idx += idx & (~idx + 1)) atomicMax(&sw.tree[s0 + idx - 1], key);
					}
				}
				ins = insEnd;
				__syncthreads();
			}
			// ---- candidates: through the links of the start node, on the paths through the start node itself, and the overlapping anchors
			const uint32_t cs = m.compMap[aj.startNode], cj = m.compMap[aj.endNode];
			const uint32_t gs = m.compStart[cs] + m.compIdx[aj.startNode];
			long long best = (long long)0x8000000000000000LL;
			if (cs == cj)
			{
				const uint32_t nLinks = m.backStart[gs + 1] - m.backStart[gs], nOwn = m.pathsStart[gs + 1] - m.pathsStart[gs];
				for (uint32_t t = threadIdx.x; t < nLinks + nOwn; t += blockDim.x)
				{
					long long v;
					if (t < nLinks) { uint32_t b = m.backStart[gs] + t; v = query(m.pathBase[cs] + m.backK[b], m.topoIds[m.compStart[cs] + m.backNode[b]]); }
					else v = query(m.pathBase[cs] + m.pathsK[m.pathsStart[gs] + (t - nLinks)], m.topoIds[gs]);
					if (v > best) best = v;
				}
			}
			if (best != (long long)0x8000000000000000LL) best += (long long)len << 32; // val = len_j + C[i]  (y_i <= x_j - 1)
			for (uint32_t oi = ins + threadIdx.x; oi < oj; oi += blockDim.x)
			{
				const uint32_t i = ord[oi];
				const GcAnchor ai = a[i];
				if (ai.y >= aj.y) continue;
				if (m.compMap[ai.endNode] != cj) continue;
				long long key = gc_k2_candidate(m, ai, aj, i, sc[i]);
				if (key > best) best = key;
			}
			long long own = gc_k2_key(len, -1);
			if (own > best) best = own;
			long long b = gc_k2_block_max(best, warpBest);
			if (threadIdx.x == 0) { sc[j] = (int32_t)(b >> 32); pr[j] = (int32_t)(uint32_t)(b & 0xFFFFFFFFu) - 1; }
			__syncthreads();
		}
	}
	else
	{
		for (uint32_t oj = 0; oj < n; oj++)
		{
			uint32_t j = ord[oj];
			GcAnchor aj = a[j];
			uint32_t cj = m.compMap[aj.endNode];
			long long best = gc_k2_key(aj.y - aj.x + 1, -1);
			for (uint32_t oi = threadIdx.x; oi < oj; oi += blockDim.x)
			{
				uint32_t i = ord[oi];
				GcAnchor ai = a[i];
				if (ai.y >= aj.y) continue;
				if (m.compMap[ai.endNode] != cj) continue;
				long long key = gc_k2_candidate(m, ai, aj, i, sc[i]);
				if (key > best) best = key;
			}
			long long b = gc_k2_block_max(best, warpBest);
			if (threadIdx.x == 0) { sc[j] = (int32_t)(b >> 32); pr[j] = (int32_t)(uint32_t)(b & 0xFFFFFFFFu) - 1; }
			__syncthreads();
		}
	}
	if (threadIdx.x == 0)
	{
		long long bs = 0;
		chainLen[r] = gc_k2_select(m, a, n, sc, pr, chain + base, (int64_t*)&bs);
		chainScore[r] = bs;
	}
}

// chaining of `total` anchors of `numReads` reads that sit in device memory (dAnchors, dReadOff[numReads + 1])
static int k2Run(gcgpu_ctx* ctx, const GcAnchor* dAnchors, const uint64_t* dReadOff, uint32_t numReads, uint64_t total, uint64_t maxPerRead,
	uint32_t* dOrder, int32_t* dScore, int32_t* dPred, uint32_t* dChain, uint32_t* dChainLen, int64_t* dChainScore)
{
	const char* force = getenv("GCGPU_K2_FORCE");
	bool sweep = !(force && !strcmp(force, "pairwise")) && total > 0 && maxPerRead > GC_K2_PAIRWISE_MAX && numReads < (1u << GC_K2_READ_BITS) && ctx->totalPaths < (1u << GC_K2_PATH_BITS) && ctx->maxCompNodes < (1u << GC_K2_TOPO_BITS);
	GcK2Sweep sw; memset(&sw, 0, sizeof(sw));
	if (sweep)
	{
		// (anchor, path) entries: count, scan, emit, sort, segment bounds
		size_t offCnt = 0, offStart = alignUp(offCnt + (total + 1) * 8, 128);
		CUDA_TRY(ctx->k2Work.ensure(offStart + (total + 1) * 8 + 128));
		uint8_t* W = (uint8_t*)ctx->k2Work.p;
		uint64_t* dCnt = (uint64_t*)(W + offCnt); uint64_t* dStart = (uint64_t*)(W + offStart);
		gc_k2_pair_count_kernel<<<(unsigned)((total + 1 + 255) / 256), 256, 0, ctx->stream>>>(ctx->mpc, dAnchors, total, dCnt);
		ctx->launches++;
		size_t scanBytes = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, dCnt, dStart, (int)(total + 1), ctx->stream);
		CUDA_TRY(ctx->copyDesc.ensure(scanBytes + 16));
		cub::DeviceScan::ExclusiveSum(ctx->copyDesc.p, scanBytes, dCnt, dStart, (int)(total + 1), ctx->stream);
		ctx->launches++;
		uint64_t entries = 0;
		CUDA_TRY(gcCopy(ctx, &entries, dStart + total, 8, cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(gcSyncStream(ctx));
		if (entries == 0 || entries >= 0xFFFFFFFFull) sweep = false;
		else
		{
			size_t oKeysA = 0, oKeysB = alignUp(oKeysA + entries * 8, 128), oValsA = alignUp(oKeysB + entries * 8, 128), oValsB = alignUp(oValsA + entries * 4, 128);
			size_t oPos = alignUp(oValsB + entries * 4, 128), oLo = alignUp(oPos + entries * 4, 128), oHi = alignUp(oLo + entries * 4, 128), oTree = alignUp(oHi + entries * 4, 128), end = oTree + entries * 8;
			CUDA_TRY(ctx->k2Sort.ensure(end));
			uint8_t* S = (uint8_t*)ctx->k2Sort.p;
			uint64_t* keysA = (uint64_t*)(S + oKeysA); uint64_t* keysB = (uint64_t*)(S + oKeysB); uint32_t* valsA = (uint32_t*)(S + oValsA); uint32_t* valsB = (uint32_t*)(S + oValsB);
			gc_k2_pair_emit_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ctx->mpc, dAnchors, dReadOff, numReads, total, dStart, keysA, valsA);
			size_t sortBytes = 0;
			cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, keysA, keysB, valsA, valsB, (int)entries, 0, 64, ctx->stream);
			CUDA_TRY(ctx->copyDesc.ensure(sortBytes + 16));
			cub::DeviceRadixSort::SortPairs(ctx->copyDesc.p, sortBytes, keysA, keysB, valsA, valsB, (int)entries, 0, 64, ctx->stream);
			gc_k2_segments_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, ctx->stream>>>(keysB, valsB, entries, (uint32_t*)(S + oPos), (uint32_t*)(S + oLo), (uint32_t*)(S + oHi));
			CUDA_TRY(cudaMemsetAsync(S + oTree, 0x80, entries * 8, ctx->stream)); // every tree node "minus infinity"
			ctx->launches += 3;
			sw.pairStart = dStart; sw.posOf = (const uint32_t*)(S + oPos); sw.keys = keysB; sw.segLo = (const uint32_t*)(S + oLo); sw.segHi = (const uint32_t*)(S + oHi);
			sw.tree = (long long*)(S + oTree); sw.entries = entries;
		}
	}
	if (sweep) gc_k2_chain_kernel<true><<<numReads, GC_K2_THREADS, 0, ctx->stream>>>(ctx->mpc, dAnchors, dReadOff, numReads, sw, dOrder, dScore, dPred, dChain, dChainLen, dChainScore);
	else gc_k2_chain_kernel<false><<<numReads, GC_K2_THREADS, 0, ctx->stream>>>(ctx->mpc, dAnchors, dReadOff, numReads, sw, dOrder, dScore, dPred, dChain, dChainLen, dChainScore);
	ctx->launches++;
	CUDA_TRY(cudaGetLastError());
	return GCGPU_OK;
}

extern "C" int gcgpu_chain(gcgpu_ctx* ctx, const gcgpu_anchor* anchors, const uint64_t* read_offsets, uint32_t num_reads, uint32_t* chain, uint32_t* chain_len, int64_t* chain_score)
{
	if (!ctx || !read_offsets || (!chain_len && num_reads) || (!chain_score && num_reads)) return setError(GCGPU_ERR_ARG, "gcgpu_chain: null argument");
	if (!ctx->haveMpc) return setError(GCGPU_ERR_ARG, "gcgpu_chain: the context was created without an MPC index");
	ctx->lastKernelMs = 0;
	if (num_reads == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	uint64_t total = read_offsets[num_reads];
	if (total && (!anchors || !chain)) return setError(GCGPU_ERR_ARG, "gcgpu_chain: null anchor/chain buffer");
	for (uint64_t i = 0; i < total; i++) if (anchors[i].start_node >= ctx->numNodes || anchors[i].end_node >= ctx->numNodes) return setError(GCGPU_ERR_ARG, "gcgpu_chain: anchor node out of range");
	static_assert(sizeof(gcgpu_anchor) == sizeof(GcAnchor), "anchor layout");
	// layout in arena: anchors | offsets | order | score | pred | chain | chainLen | chainScore
	size_t offA = 0, offO = alignUp(offA + total * sizeof(GcAnchor), 128), offOrd = alignUp(offO + ((size_t)num_reads + 1) * 8, 128), offSc = alignUp(offOrd + total * 4, 128), offPr = alignUp(offSc + total * 4, 128);
	size_t offCh = alignUp(offPr + total * 4, 128), offLen = alignUp(offCh + total * 4, 128), offScore = alignUp(offLen + (size_t)num_reads * 4, 128), end = offScore + (size_t)num_reads * 8;
	CUDA_TRY(ctx->arena.ensure(end));
	uint8_t* A = (uint8_t*)ctx->arena.p;
	if (total) CUDA_TRY(gcCopy(ctx, A + offA, anchors, total * sizeof(GcAnchor), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(gcCopy(ctx, A + offO, read_offsets, ((size_t)num_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
	{
		uint64_t maxPerRead = 0;
		for (uint32_t r = 0; r < num_reads; r++) maxPerRead = std::max<uint64_t>(maxPerRead, read_offsets[r + 1] - read_offsets[r]);
		int krc = k2Run(ctx, (const GcAnchor*)(A + offA), (const uint64_t*)(A + offO), num_reads, total, maxPerRead,
			(uint32_t*)(A + offOrd), (int32_t*)(A + offSc), (int32_t*)(A + offPr), (uint32_t*)(A + offCh), (uint32_t*)(A + offLen), (int64_t*)(A + offScore));
		if (krc != GCGPU_OK) return krc;
	}
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	if (total) CUDA_TRY(gcCopy(ctx, chain, A + offCh, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcCopy(ctx, chain_len, A + offLen, (size_t)num_reads * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcCopy(ctx, chain_score, A + offScore, (size_t)num_reads * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	float ms = 0;
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->lastKernelMs = ms;
	return GCGPU_OK;
}

#include "gcgpu_resident.inl"
