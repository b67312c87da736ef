// libgcgpu: C ABI + CUDA kernels (sm_100a) of the GraphChainer alignment hot path.
// Interface and the reference seams each entry point replaces: include/gcgpu.h.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/gcgpu.h"
#include "gc_common.cuh"
#include "gc_k1.cuh"
#include "gc_host_graph.h"

#define GCGPU_VERSION 1

static thread_local std::string g_lastError;
static int setError(int code, const std::string& msg) { g_lastError = msg; return code; }

#define CUDA_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return setError(_e == cudaErrorMemoryAllocation ? GCGPU_ERR_NOMEM : GCGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)

// growable device buffer
struct DevBuf
{
	void* p = nullptr;
	size_t cap = 0;
	cudaError_t ensure(size_t bytes)
	{
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 4 + 4096;
		cudaError_t e = cudaMalloc(&p, want);
		if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
		if (e == cudaSuccess) cap = want;
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

struct gcgpu_ctx
{
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	gcgpu_params params;
	uint32_t numNodes = 0;
	// device copies of the graph
	uint8_t* d_nodeLength = nullptr; uint64_t* d_nodeSeq = nullptr;
	uint32_t* d_inStart = nullptr; uint32_t* d_inNbr = nullptr; uint32_t* d_outStart = nullptr; uint32_t* d_outNbr = nullptr;
	uint32_t* d_componentNumber = nullptr; uint8_t* d_linearizable = nullptr;
	GcViterbiTables* d_vt = nullptr;
	GcGraphView view;
	DevBuf seqBuf, descBuf, resBuf, arena, traceArena, compact, copyDesc;
	float lastKernelMs = 0;
	uint64_t launches = 0;
};

// ------------------------------------------------------------------ K1 kernel
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
	return alignUp((size_t)(numSlices + 2) * sizeof(GcSliceMeta), 16) + (size_t)itemCap * sizeof(GcNodeItem) + (size_t)heapCap * 8;
}

// One thread = one extension work item (see gc_k1.cuh for the design rationale).
__global__ void __launch_bounds__(128) gc_k1_kernel(GcGraphView g, const GcViterbiTables* __restrict__ vt, GcK1Params prm, const uint8_t* __restrict__ seq,
	const GcK1Desc* __restrict__ descs, uint32_t n, uint8_t* arena, uint64_t* traceArena, GcK1Result* results)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	GcK1Desc d = descs[t];
	GcK1Workspace ws;
	uint8_t* base = arena + d.wsOff;
	ws.slices = (GcSliceMeta*)base;
	size_t slicesBytes = ((size_t)(d.numSlices + 2) * sizeof(GcSliceMeta) + 15) / 16 * 16;
	ws.items = (GcNodeItem*)(base + slicesBytes);
	ws.heap = (uint64_t*)(base + slicesBytes + (size_t)d.itemCap * sizeof(GcNodeItem));
	ws.itemCap = d.itemCap;
	ws.heapCap = d.heapCap;
	GcK1Result res;
	gc_k1_extend(g, *vt, prm, seq + d.seqOff, d.seqLen, d.node, d.offset, ws, traceArena + d.traceOff, d.traceCap, res);
	results[d.resultIndex] = res;
}

// gather the per-item traces into one dense buffer: one warp per item
struct GcCopyDesc { uint64_t src; uint64_t dst; uint32_t len; uint32_t pad; };
__global__ void gc_trace_gather_kernel(const GcCopyDesc* __restrict__ descs, uint32_t n, const uint64_t* __restrict__ traceArena, uint64_t* __restrict__ out)
{
	uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	uint32_t lane = threadIdx.x & 31;
	if (warp >= n) return;
	GcCopyDesc d = descs[warp];
	for (uint32_t i = lane; i < d.len; i += 32) out[d.dst + i] = traceArena[d.src + i];
}

// ------------------------------------------------------------------ C ABI
extern "C" int gcgpu_version(void) { return GCGPU_VERSION; }
extern "C" const char* gcgpu_last_error(void) { return g_lastError.c_str(); }

extern "C" void gcgpu_destroy(gcgpu_ctx* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaFree(ctx->d_nodeLength); cudaFree(ctx->d_nodeSeq); cudaFree(ctx->d_inStart); cudaFree(ctx->d_inNbr); cudaFree(ctx->d_outStart); cudaFree(ctx->d_outNbr);
	cudaFree(ctx->d_componentNumber); cudaFree(ctx->d_linearizable); cudaFree(ctx->d_vt);
	ctx->seqBuf.release(); ctx->descBuf.release(); ctx->resBuf.release(); ctx->arena.release(); ctx->traceArena.release(); ctx->compact.release(); ctx->copyDesc.release();
	if (ctx->ev0) cudaEventDestroy(ctx->ev0);
	if (ctx->ev1) cudaEventDestroy(ctx->ev1);
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
	chk(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	chk(cudaEventCreate(&ctx->ev0));
	chk(cudaEventCreate(&ctx->ev1));
	chk(uploadArray(graph->node_length, N, &ctx->d_nodeLength));
	chk(uploadArray(graph->node_seq, 2 * (size_t)N, &ctx->d_nodeSeq));
	chk(uploadArray(graph->in_start, (size_t)N + 1, &ctx->d_inStart));
	chk(uploadArray(graph->in_nbr, graph->in_start[N], &ctx->d_inNbr));
	chk(uploadArray(graph->out_start, (size_t)N + 1, &ctx->d_outStart));
	chk(uploadArray(graph->out_nbr, graph->out_start[N], &ctx->d_outNbr));
	chk(uploadArray(graph->component_number, N, &ctx->d_componentNumber));
	chk(uploadArray(graph->linearizable, N, &ctx->d_linearizable));
	GcViterbiTables vt = gcMakeViterbiTables();
	chk(uploadArray(&vt, 1, &ctx->d_vt));
	if (err != cudaSuccess)
	{
		std::string msg = std::string("gcgpu_create: ") + cudaGetErrorString(err);
		gcgpu_destroy(ctx);
		return setError(err == cudaErrorMemoryAllocation ? GCGPU_ERR_NOMEM : GCGPU_ERR_CUDA, msg);
	}
	ctx->view.numNodes = N;
	ctx->view.nodeLength = ctx->d_nodeLength; ctx->view.nodeSeq = ctx->d_nodeSeq;
	ctx->view.inStart = ctx->d_inStart; ctx->view.inNbr = ctx->d_inNbr; ctx->view.outStart = ctx->d_outStart; ctx->view.outNbr = ctx->d_outNbr;
	ctx->view.componentNumber = ctx->d_componentNumber; ctx->view.linearizable = ctx->d_linearizable;
	*out = ctx;
	return GCGPU_OK;
}

extern "C" float gcgpu_last_kernel_ms(gcgpu_ctx* ctx) { return ctx ? ctx->lastKernelMs : 0.f; }
extern "C" uint64_t gcgpu_launch_count(gcgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int gcgpu_extend(gcgpu_ctx* ctx, const uint8_t* seq, uint64_t seq_bytes, const gcgpu_ext_item* items, uint32_t n,
	gcgpu_ext_result* results, uint64_t* traces, uint64_t trace_capacity, uint64_t* trace_used)
{
	if (!ctx || (!items && n) || (!results && n) || !trace_used) return setError(GCGPU_ERR_ARG, "gcgpu_extend: null argument");
	*trace_used = 0;
	ctx->lastKernelMs = 0;
	if (n == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	for (uint32_t i = 0; i < n; i++)
	{
		if (items[i].seq_len < 0 || items[i].seq_offset + (uint64_t)items[i].seq_len > seq_bytes || items[i].node >= ctx->numNodes || items[i].seq_len >= (1 << 24))
			return setError(GCGPU_ERR_ARG, "gcgpu_extend: item " + std::to_string(i) + " out of range");
	}
	CUDA_TRY(ctx->seqBuf.ensure(seq_bytes + 16));
	if (seq_bytes) CUDA_TRY(cudaMemcpyAsync(ctx->seqBuf.p, seq, seq_bytes, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx->resBuf.ensure((size_t)n * sizeof(GcK1Result)));

	// work list, longest first so that the threads of a warp carry similar loads
	std::vector<uint32_t> todo(n);
	for (uint32_t i = 0; i < n; i++) todo[i] = i;
	std::sort(todo.begin(), todo.end(), [items](uint32_t a, uint32_t b) { return items[a].seq_len != items[b].seq_len ? items[a].seq_len > items[b].seq_len : a < b; });
	std::vector<GcK1Result> hres(n);
	std::vector<uint64_t> traceOffOfItem(n, 0);
	std::vector<GcK1Desc> descs;
	// trace arena: every item keeps a fixed slot across retries
	uint64_t traceTotal = 0;
	for (uint32_t i = 0; i < n; i++) { traceOffOfItem[i] = traceTotal; traceTotal += 2 * (uint64_t)items[i].seq_len + 72; }
	CUDA_TRY(ctx->traceArena.ensure(traceTotal * 8));
	uint32_t itemScale = 1, heapScale = 1;
	GcK1Params prm; prm.bandwidth = ctx->params.initial_bandwidth;
	for (int attempt = 0; attempt < 8 && !todo.empty(); attempt++)
	{
		descs.resize(todo.size());
		size_t wsTotal = 0;
		for (size_t k = 0; k < todo.size(); k++)
		{
			const gcgpu_ext_item& it = items[todo[k]];
			GcK1Desc& d = descs[k];
			d.seqOff = it.seq_offset; d.seqLen = it.seq_len; d.node = it.node; d.offset = it.offset;
			d.numSlices = (uint32_t)((it.seq_len + 63) / 64);
			d.itemCap = (24 + 8 * d.numSlices) * itemScale;
			d.heapCap = 64 * heapScale;
			d.traceCap = (uint32_t)(2 * (uint64_t)it.seq_len + 72);
			d.traceOff = traceOffOfItem[todo[k]];
			d.resultIndex = todo[k];
			d.wsOff = wsTotal;
			wsTotal += alignUp(k1WorkspaceBytes(d.numSlices, d.itemCap, d.heapCap), 128);
		}
		CUDA_TRY(ctx->arena.ensure(wsTotal));
		CUDA_TRY(ctx->descBuf.ensure(descs.size() * sizeof(GcK1Desc)));
		CUDA_TRY(cudaMemcpyAsync(ctx->descBuf.p, descs.data(), descs.size() * sizeof(GcK1Desc), cudaMemcpyHostToDevice, ctx->stream));
		uint32_t m = (uint32_t)descs.size();
		CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
		gc_k1_kernel<<<(m + 127) / 128, 128, 0, ctx->stream>>>(ctx->view, ctx->d_vt, prm, (const uint8_t*)ctx->seqBuf.p, (const GcK1Desc*)ctx->descBuf.p, m, (uint8_t*)ctx->arena.p, (uint64_t*)ctx->traceArena.p, (GcK1Result*)ctx->resBuf.p);
		ctx->launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
		CUDA_TRY(cudaMemcpyAsync(hres.data(), ctx->resBuf.p, (size_t)n * sizeof(GcK1Result), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(cudaStreamSynchronize(ctx->stream));
		float ms = 0;
		CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		ctx->lastKernelMs += ms;
		std::vector<uint32_t> again;
		bool needItems = false, needHeap = false;
		for (uint32_t idx : todo)
		{
			int st = hres[idx].status;
			if (st == GC_OVERFLOW_ITEMS) { again.push_back(idx); needItems = true; }
			else if (st == GC_OVERFLOW_HEAP) { again.push_back(idx); needHeap = true; }
		}
		if (needItems) itemScale *= 4;
		if (needHeap) heapScale *= 4;
		todo.swap(again);
	}
	if (!todo.empty()) return setError(GCGPU_ERR_NOMEM, "gcgpu_extend: " + std::to_string(todo.size()) + " work items still overflow their workspace after 8 attempts");

	// results + dense traces
	uint64_t used = 0;
	std::vector<GcCopyDesc> copies;
	copies.reserve(n);
	bool internal = false;
	for (uint32_t i = 0; i < n; i++)
	{
		const GcK1Result& r = hres[i];
		results[i].status = r.status == GC_OK ? GCGPU_ITEM_OK : (r.status == GC_FAILED ? GCGPU_ITEM_FAILED : GCGPU_ITEM_INTERNAL);
		if (results[i].status == GCGPU_ITEM_INTERNAL) internal = true;
		results[i].score = r.score;
		results[i].trace_len = r.status == GC_OK ? r.traceLen : 0;
		results[i].reserved = 0;
		results[i].trace_offset = used;
		results[i].columns = r.columns;
		if (results[i].trace_len)
		{
			GcCopyDesc c; c.src = traceOffOfItem[i]; c.dst = used; c.len = results[i].trace_len; c.pad = 0;
			copies.push_back(c);
			used += results[i].trace_len;
		}
	}
	*trace_used = used;
	if (used > trace_capacity) return setError(GCGPU_ERR_ARG, "gcgpu_extend: trace buffer too small, need " + std::to_string(used) + " entries");
	if (used)
	{
		if (!traces) return setError(GCGPU_ERR_ARG, "gcgpu_extend: null trace buffer");
		CUDA_TRY(ctx->compact.ensure(used * 8));
		CUDA_TRY(ctx->copyDesc.ensure(copies.size() * sizeof(GcCopyDesc)));
		CUDA_TRY(cudaMemcpyAsync(ctx->copyDesc.p, copies.data(), copies.size() * sizeof(GcCopyDesc), cudaMemcpyHostToDevice, ctx->stream));
		uint32_t m = (uint32_t)copies.size();
		CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
		gc_trace_gather_kernel<<<(m + 3) / 4, 128, 0, ctx->stream>>>((const GcCopyDesc*)ctx->copyDesc.p, m, (const uint64_t*)ctx->traceArena.p, (uint64_t*)ctx->compact.p);
		ctx->launches++;
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
		CUDA_TRY(cudaMemcpyAsync(traces, ctx->compact.p, used * 8, cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(cudaStreamSynchronize(ctx->stream));
		float ms = 0;
		CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		ctx->lastKernelMs += ms;
	}
	if (internal) return setError(GCGPU_ERR_INTERNAL, "gcgpu_extend: a work item reached a state the reference asserts on (see per-item status)");
	return GCGPU_OK;
}
