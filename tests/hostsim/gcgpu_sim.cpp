// TEST DOUBLE of libgcgpu's C ABI (include/gcgpu.h) for the GPU-less build box.
//
// It runs the SAME GC_HD work-item functions that gcgpu.cu launches as CUDA kernels, but
// compiled for the host, so that `-m "not gpu"` tests can exercise the host driver logic
// (seed loops, anchor building, chain connection, output encoding) end to end against the
// reference's golden output.  It lives under tests/, is never built into the package and
// never shipped: the product (libgcgpu.so) has no CPU path.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/gcgpu.h"
#include "../../graphchainer_b200/csrc/gc_host_graph.h"
#include "../../graphchainer_b200/csrc/gc_k1.cuh"
#include "../../graphchainer_b200/csrc/gc_k2.cuh"
#include "../../graphchainer_b200/csrc/gc_k3.cuh"
#include "../../graphchainer_b200/csrc/gc_seed.cuh"

struct gcgpu_ctx
{
	GcGraphView view;
	GcMpcView mpc;
	GcViterbiTables vt;
	int bandwidth;
	uint32_t numNodes;
	uint64_t launches = 0;
	std::vector<uint8_t> seqCopy, nwCodes;
	std::vector<uint64_t> dense; // traces of the last gcgpu_extend call
	std::vector<GcMzSlot> mzSlots; GcMzView mz; bool haveMz = false;
	std::vector<gcgpu_seed_match> denseMatches;
};
static std::string g_err;

extern "C" int gcgpu_version(void) { return 1; }
extern "C" const char* gcgpu_last_error(void) { return g_err.c_str(); }
extern "C" int gcgpu_create(int, const gcgpu_graph* g, const gcgpu_params* p, gcgpu_ctx** out)
{
	gcgpu_ctx* c = new gcgpu_ctx();
	c->view.numNodes = g->num_nodes; c->view.nodeLength = g->node_length; c->view.nodeSeq = g->node_seq;
	c->view.inStart = g->in_start; c->view.inNbr = g->in_nbr; c->view.outStart = g->out_start; c->view.outNbr = g->out_nbr;
	c->view.componentNumber = g->component_number; c->view.linearizable = g->linearizable; c->view.coopLane = -1; c->view.coopWidth = 32; c->view.coopMask = 0xFFFFFFFFu; c->view.coopShift = 0;
	c->mpc.compMap = g->comp_map; c->mpc.compIdx = g->comp_idx; c->mpc.compStart = g->comp_start; c->mpc.topoIds = g->topo_ids;
	c->mpc.pathsStart = g->paths_start; c->mpc.pathsK = g->paths_k; c->mpc.backStart = g->back_start; c->mpc.backNode = g->back_node; c->mpc.backK = g->back_k;
	c->vt = gcMakeViterbiTables();
	c->bandwidth = p ? p->initial_bandwidth : 10;
	c->numNodes = g->num_nodes;
	*out = c;
	return 0;
}
extern "C" void gcgpu_destroy(gcgpu_ctx* c) { delete c; }
extern "C" float gcgpu_last_kernel_ms(gcgpu_ctx*) { return 0; }
extern "C" uint64_t gcgpu_launch_count(gcgpu_ctx* c) { return c->launches; }

extern "C" void* gcgpu_host_alloc(size_t bytes) { return malloc(bytes); }
extern "C" void gcgpu_host_free(void* p) { free(p); }

extern "C" int gcgpu_extend(gcgpu_ctx* ctx, const uint8_t* seqIn, uint64_t seqBytes, const gcgpu_ext_item* items, uint32_t n, gcgpu_ext_result* results, uint64_t* traces, uint64_t trace_capacity, uint64_t* trace_used)
{
	if (seqIn) ctx->seqCopy.assign(seqIn, seqIn + seqBytes); // "resident" buffer of the real library
	const uint8_t* seq = ctx->seqCopy.data();
	std::vector<std::vector<uint64_t>> tr(n);
	bool internal = false;
	#pragma omp parallel for schedule(dynamic, 64)
	for (uint32_t i = 0; i < n; i++)
	{
		int32_t seqLen = items[i].seq_len;
		int32_t numSlices = (seqLen + 63) / 64;
		uint32_t itemCap = 24 + 8 * numSlices, heapCap = 64;
		GcK1Result res;
		std::vector<uint64_t> trace(2 * (size_t)seqLen + 72);
		for (int attempt = 0; attempt < 8; attempt++)
		{
			std::vector<GcSliceMeta> slices(numSlices + 2);
			std::vector<GcNodeItem> nodeItems(itemCap);
			std::vector<uint64_t> heap(heapCap);
			GcWord colsBuf[64];
			GcK1Workspace ws { slices.data(), nodeItems.data(), heap.data(), colsBuf, itemCap, heapCap };
			GcK1Params prm { ctx->bandwidth };
			gc_k1_extend(ctx->view, ctx->vt, prm, seq + items[i].seq_offset, seqLen, items[i].node, items[i].offset, ws, trace.data(), (uint32_t)trace.size(), res);
			if (res.status == GC_OVERFLOW_ITEMS) { itemCap *= 4; continue; }
			if (res.status == GC_OVERFLOW_HEAP) { heapCap *= 4; continue; }
			break;
		}
		results[i].status = res.status == GC_OK ? GCGPU_ITEM_OK : (res.status == GC_FAILED ? GCGPU_ITEM_FAILED : GCGPU_ITEM_INTERNAL);
		results[i].score = res.score;
		results[i].trace_len = res.status == GC_OK ? res.traceLen : 0;
		results[i].columns = res.columns;
		results[i].reserved = 0;
		trace.resize(results[i].trace_len);
		tr[i].swap(trace);
	}
	uint64_t used = 0;
	for (uint32_t i = 0; i < n; i++) { results[i].trace_offset = used; used += results[i].trace_len; if (results[i].status == GCGPU_ITEM_INTERNAL) internal = true; }
	*trace_used = used;
	ctx->dense.resize(used);
	for (uint32_t i = 0; i < n; i++) if (results[i].trace_len) memcpy(ctx->dense.data() + results[i].trace_offset, tr[i].data(), results[i].trace_len * 8);
	if (!traces && trace_capacity == 0) used = 0; // two-phase form: gcgpu_fetch_traces follows
	if (used > trace_capacity) { g_err = "trace buffer too small"; return GCGPU_ERR_ARG; }
	if (used) memcpy(traces, ctx->dense.data(), used * 8);
	ctx->launches++;
	return internal ? GCGPU_ERR_INTERNAL : GCGPU_OK;
}

static uint8_t k3code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4; }
extern "C" int gcgpu_fetch_traces(gcgpu_ctx* ctx, uint64_t* traces, uint64_t first, uint64_t count)
{
	if (first + count > ctx->dense.size()) { g_err = "range beyond the traces of the last call"; return GCGPU_ERR_ARG; }
	if (count) memcpy(traces, ctx->dense.data() + first, count * 8);
	return GCGPU_OK;
}

extern "C" int gcgpu_nw(gcgpu_ctx* ctx, const char* seqs, uint64_t seq_bytes, const gcgpu_nw_item* items, uint32_t n, gcgpu_nw_result* results, uint8_t* ops, uint64_t ops_capacity, uint64_t* ops_used)
{
	if (seqs) { ctx->nwCodes.resize(seq_bytes); for (uint64_t i = 0; i < seq_bytes; i++) ctx->nwCodes[i] = k3code(seqs[i]); } // "resident" buffer of the real library
	else if (ctx->nwCodes.size() != seq_bytes) { g_err = "gcgpu_nw: seqs == NULL but no sequence buffer of this size is resident"; return GCGPU_ERR_ARG; }
	const std::vector<uint8_t>& codes = ctx->nwCodes;
	std::vector<std::vector<uint8_t>> allOps(n);
	bool internal = false;
	#pragma omp parallel for schedule(dynamic, 1)
	for (uint32_t i = 0; i < n; i++)
	{
		const gcgpu_nw_item& it = items[i];
		int32_t Q = it.query_len, T = it.target_len;
		int32_t nb = (Q + 63) / 64; if (nb < 1) nb = 1;
		const uint8_t* q = codes.data() + it.query_offset; const uint8_t* t = codes.data() + it.target_offset;
		std::vector<uint64_t> peq(4 * (size_t)nb), rpeq(4 * (size_t)nb);
		gc_k3_build_peq(q, Q, peq.data(), nb);
		std::vector<GcK3Block> ba(nb + 1), bb(nb + 1);
		uint64_t work = 0;
		int32_t d = gc_k3_distance(peq.data(), nb, Q, t, T, ba.data(), it.k_hint, work);
		results[i].status = 0; results[i].distance = d; results[i].ops_len = 0; results[i].reserved = 0;
		if (it.want_path && Q > 0 && T > 0)
		{
			std::vector<uint8_t> rq(q, q + Q); std::reverse(rq.begin(), rq.end());
			gc_k3_build_peq(rq.data(), Q, rpeq.data(), nb);
			GcK3PathWorkspace w;
			w.peq = peq.data(); w.rpeq = rpeq.data(); w.nbTotal = nb; w.qTotal = Q; w.tTotal = T; w.blocksA = ba.data(); w.blocksB = bb.data();
			std::vector<GcK3Block> store(52432); std::vector<uint32_t> colStart(37456); std::vector<GcK3Frame> stack(96);
			w.store = store.data(); w.storeCap = (uint32_t)store.size(); w.colStart = colStart.data(); w.colCap = (uint32_t)colStart.size(); w.stack = stack.data(); w.stackCap = (uint32_t)stack.size();
			allOps[i].resize((size_t)Q + T + 8);
			uint32_t nOps = 0;
			if (!gc_k3_path(w, t, d, allOps[i].data(), nOps, (uint32_t)allOps[i].size(), work)) { results[i].status = GCGPU_ITEM_INTERNAL; nOps = 0; }
			allOps[i].resize(nOps);
			results[i].ops_len = nOps;
		}
		results[i].blocks = work;
	}
	uint64_t used = 0;
	for (uint32_t i = 0; i < n; i++) { results[i].ops_offset = used; used += results[i].ops_len; if (results[i].status) internal = true; }
	*ops_used = used;
	if (used > ops_capacity) { g_err = "ops buffer too small"; return GCGPU_ERR_ARG; }
	for (uint32_t i = 0; i < n; i++) if (results[i].ops_len) memcpy(ops + results[i].ops_offset, allOps[i].data(), results[i].ops_len);
	ctx->launches++;
	return internal ? GCGPU_ERR_INTERNAL : GCGPU_OK;
}

extern "C" int gcgpu_chain(gcgpu_ctx* ctx, const gcgpu_anchor* anchors, const uint64_t* read_offsets, uint32_t num_reads, uint32_t* chain, uint32_t* chain_len, int64_t* chain_score)
{
	#pragma omp parallel for schedule(dynamic, 8)
	for (uint32_t r = 0; r < num_reads; r++)
	{
		uint64_t base = read_offsets[r];
		uint32_t n = (uint32_t)(read_offsets[r + 1] - base);
		const GcAnchor* a = (const GcAnchor*)(anchors + base);
		std::vector<uint32_t> order(n);
		for (uint32_t i = 0; i < n; i++) order[i] = i;
		std::stable_sort(order.begin(), order.end(), [a](uint32_t l, uint32_t rr) { return a[l].y < a[rr].y; });
		std::vector<int32_t> score(n), pred(n);
		int64_t best = 0;
		chain_len[r] = gc_k2_chain_seq(ctx->mpc, a, n, order.data(), score.data(), pred.data(), chain + base, &best);
		chain_score[r] = best;
	}
	ctx->launches++;
	return GCGPU_OK;
}

extern "C" int gcgpu_set_minimizer_index(gcgpu_ctx* ctx, const gcgpu_minimizer_index* idx)
{
	uint64_t cap = 16;
	while (cap < idx->num_kmers * 2 + 2) cap <<= 1;
	ctx->mzSlots.assign(cap, GcMzSlot { 0, 0, 0xFFFFFFFFu });
	for (uint64_t i = 0; i < idx->num_kmers; i++)
	{
		uint64_t h = gc_mz_hash(idx->kmers[i]) & (cap - 1);
		while (ctx->mzSlots[h].count != 0xFFFFFFFFu && ctx->mzSlots[h].key != idx->kmers[i]) h = (h + 1) & (cap - 1);
		ctx->mzSlots[h] = GcMzSlot { idx->kmers[i], idx->kmer_start[i], idx->kmer_start[i + 1] - idx->kmer_start[i] };
	}
	ctx->mz.slots = ctx->mzSlots.data(); ctx->mz.mask = cap - 1; ctx->mz.k = idx->k; ctx->mz.realWindow = idx->window - idx->k + 1; ctx->mz.maxCount = idx->max_count;
	ctx->haveMz = true;
	return 0;
}
extern "C" int gcgpu_seed(gcgpu_ctx* ctx, const uint8_t* seqIn, uint64_t seqBytes, const gcgpu_seed_read* reads, uint32_t n, uint64_t* match_offsets, gcgpu_seed_match* matches, uint64_t capacity, uint64_t* used)
{
	if (!ctx->haveMz) { g_err = "gcgpu_seed: no minimizer index"; return GCGPU_ERR_ARG; }
	if (seqIn) ctx->seqCopy.assign(seqIn, seqIn + seqBytes);
	const uint8_t* seq = ctx->seqCopy.data();
	std::vector<std::vector<gcgpu_seed_match>> per(n);
	#pragma omp parallel for schedule(dynamic, 4)
	for (uint32_t r = 0; r < n; r++)
		for (int32_t i = 0; i < reads[r].seq_len; i++)
		{
			uint32_t start, count;
			if (gc_seed_position(ctx->mz, seq + reads[r].seq_offset, reads[r].seq_len, i, start, count)) per[r].push_back(gcgpu_seed_match { (uint32_t)i, start, count });
		}
	ctx->denseMatches.clear();
	match_offsets[0] = 0;
	for (uint32_t r = 0; r < n; r++) { ctx->denseMatches.insert(ctx->denseMatches.end(), per[r].begin(), per[r].end()); match_offsets[r + 1] = ctx->denseMatches.size(); }
	ctx->launches += 3;
	*used = ctx->denseMatches.size();
	if (!matches && capacity == 0) return 0;
	if (*used > capacity) { g_err = "gcgpu_seed: match buffer too small"; return GCGPU_ERR_ARG; }
	if (*used) memcpy(matches, ctx->denseMatches.data(), *used * sizeof(gcgpu_seed_match));
	return 0;
}
extern "C" int gcgpu_fetch_seed_matches(gcgpu_ctx* ctx, gcgpu_seed_match* matches, uint64_t first, uint64_t count)
{
	if (first + count > ctx->denseMatches.size()) { g_err = "gcgpu_fetch_seed_matches: range"; return GCGPU_ERR_ARG; }
	if (count) memcpy(matches, ctx->denseMatches.data() + first, count * sizeof(gcgpu_seed_match));
	return 0;
}
extern "C" int gcgpu_int_peak(gcgpu_ctx*, double* v) { *v = 0; return 0; }
