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
#include "../../graphchainer_b200/csrc/gc_k1s.cuh"
#include "../../graphchainer_b200/csrc/gc_k2.cuh"
#include "../../graphchainer_b200/csrc/gc_k3.cuh"
#include "../../graphchainer_b200/csrc/gc_seed.cuh"
#include "../../graphchainer_b200/csrc/gc_post.cuh"
#include "../../graphchainer_b200/csrc/gc_post_host.h"
#include "../../graphchainer_b200/csrc/gc_gam.cuh"

struct gcgpu_ctx
{
	GcGraphView view;
	std::vector<GcNodeRec> nodeRecs; std::vector<uint64_t> outKeys;
	GcMpcView mpc;
	GcViterbiTables vt;
	int bandwidth;
	uint32_t numNodes;
	uint64_t launches = 0;
	std::vector<uint8_t> seqCopy, nwCodes;
	std::vector<uint64_t> dense; // traces of the last gcgpu_extend call
	std::vector<GcMzSlot> mzSlots; GcMzView mz; bool haveMz = false;
	std::vector<gcgpu_seed_match> denseMatches;
	// resident batch
	std::vector<int32_t> nodeIDs; std::vector<uint32_t> nodeOffset, origNodes; GcRevTables rev; GcPostGraph pg; bool havePost = false;
	uint8_t codeTable[256];
	std::vector<GcReadDesc> reads; std::vector<GcSeedCell> cells;
	struct Set { std::vector<uint64_t> traces; std::vector<GcPair> pairs; } sets[GCGPU_TRACE_SETS];
	std::vector<GcAnchor> anchors; std::vector<gcgpu_chained_anchor> anchorMeta; std::vector<uint32_t> anchorPaths; std::vector<uint64_t> readAnchorOff;
	std::vector<gcgpu_chained_anchor> chainedMeta; std::vector<uint32_t> chainedPaths;
	std::vector<uint32_t> tokens;
	std::vector<char> charsCopy; std::vector<int32_t> origIds, origIndexOfId; std::vector<uint32_t> nameOff; std::vector<uint8_t> nameChars; bool haveNames = false; GcDeflateTables gamTables;
	std::vector<uint8_t> gamOut;
	uint64_t h2d = 0, d2h = 0;
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
	c->view.nodeRec = nullptr; c->view.outKey = nullptr;
	gcBuildNodeRecs(c->view, c->nodeRecs, c->outKeys);
	c->view.nodeRec = c->nodeRecs.data(); c->view.outKey = c->outKeys.data();
	c->mpc.compMap = g->comp_map; c->mpc.compIdx = g->comp_idx; c->mpc.compStart = g->comp_start; c->mpc.topoIds = g->topo_ids;
	c->mpc.pathsStart = g->paths_start; c->mpc.pathsK = g->paths_k; c->mpc.backStart = g->back_start; c->mpc.backNode = g->back_node; c->mpc.backK = g->back_k;
	c->vt = gcMakeViterbiTables();
	c->bandwidth = p ? p->initial_bandwidth : 10;
	c->numNodes = g->num_nodes;
	if (g->node_ids && g->node_offset && g->orig_ids && g->orig_start && g->orig_nodes && g->orig_size && g->num_orig)
	{
		c->nodeIDs.assign(g->node_ids, g->node_ids + g->num_nodes); c->nodeOffset.assign(g->node_offset, g->node_offset + g->num_nodes);
		c->origNodes.assign(g->orig_nodes, g->orig_nodes + g->orig_start[g->num_orig]);
		if (!gcBuildRevTables(g, c->rev)) { g_err = "gcgpu_create: orig_nodes entry out of range"; delete c; return GCGPU_ERR_ARG; }
		c->pg.nodeIDs = c->nodeIDs.data(); c->pg.nodeOffset = c->nodeOffset.data(); c->pg.nodeLength = g->node_length; c->pg.nodeSeq = g->node_seq;
		c->pg.revFirst = c->rev.revFirst.data(); c->pg.revCount = c->rev.revCount.data(); c->pg.revLast = c->rev.revLast.data(); c->pg.origNodes = c->origNodes.data();
		gcBuildCodeTable(c->codeTable);
		c->origIds.assign(g->orig_ids, g->orig_ids + g->num_orig);
		c->havePost = true;
	}
	*out = c;
	return 0;
}
extern "C" void gcgpu_destroy(gcgpu_ctx* c) { delete c; }
extern "C" float gcgpu_last_kernel_ms(gcgpu_ctx*) { return 0; }
extern "C" uint64_t gcgpu_launch_count(gcgpu_ctx* c) { return c->launches; }

extern "C" void* gcgpu_host_alloc(size_t bytes) { return malloc(bytes); }
extern "C" void gcgpu_host_free(void* p) { free(p); }

// K1 for a list of items; an item with seq_len < 0 does not exist (status FAILED).  Dense traces appended to `dense`.
static bool simK1(gcgpu_ctx* ctx, const gcgpu_ext_item* items, uint32_t n, gcgpu_ext_result* results, std::vector<uint64_t>& dense)
{
	const uint8_t* seq = ctx->seqCopy.data();
	std::vector<std::vector<uint64_t>> tr(n);
	bool internal = false;
	#pragma omp parallel for schedule(dynamic, 64)
	for (uint32_t i = 0; i < n; i++)
	{
		int32_t seqLen = items[i].seq_len;
		if (seqLen < 0) { results[i].status = GCGPU_ITEM_FAILED; results[i].score = 0; results[i].trace_len = 0; results[i].columns = 0; results[i].reserved = 0; continue; }
		int32_t numSlices = (seqLen + 63) / 64;
		uint32_t itemCap = 24 + 8 * numSlices, heapCap = 64;
		GcK1Result res;
		std::vector<uint64_t> trace(2 * (size_t)seqLen + 72);
		for (int attempt = 0; attempt < 8; attempt++)
		{
			std::vector<GcSliceMeta> slices(numSlices + 2);
			std::vector<GcNodeItem> nodeItems(itemCap);
			std::vector<uint64_t> heap(heapCap);
			GcColVV colsBuf[64];
			GcK1Workspace ws { slices.data(), nodeItems.data(), heap.data(), colsBuf, itemCap, heapCap };
			GcK1Params prm { ctx->bandwidth };
			if (seqLen >= 96)
			{
				// whole-read extensions: the lane-per-item form the library launches for them (gc_k1s.cuh), as a warp of one lane
				std::vector<uint32_t> keys(itemCap), scratch(2 * (size_t)heapCap);
				std::vector<GcItemAux> aux(itemCap);
				GcK1SWorkspace sw; sw.slices = slices.data(); sw.items = nodeItems.data(); sw.keys = keys.data(); sw.aux = aux.data(); sw.scratch = scratch.data(); sw.scratchCap = (uint32_t)scratch.size(); sw.itemCap = itemCap;
				sw.heap.base = heap.data(); sw.heap.stride = 1; sw.heap.cap = heapCap;
				res.score = GC_INT_MAX; res.traceLen = 0; res.itemsUsed = 0;
				int32_t last = gc_k1s_forward(ctx->view, ctx->vt, prm, true, seq + items[i].seq_offset, seqLen, items[i].node, items[i].offset, nullptr, 0, sw, res);
				if (res.status == GC_OK && last < 1) res.status = GC_FAILED;
				if (res.status == GC_OK) gc_k1s_backtrace(ctx->view, true, seq + items[i].seq_offset, seqLen, nullptr, 0, sw, last, colsBuf, trace.data(), (uint32_t)trace.size(), res);
			}
			else
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
	uint64_t used = dense.size();
	for (uint32_t i = 0; i < n; i++) { results[i].trace_offset = used; used += results[i].trace_len; if (results[i].status == GCGPU_ITEM_INTERNAL) internal = true; }
	dense.resize(used);
	for (uint32_t i = 0; i < n; i++) if (results[i].trace_len) memcpy(dense.data() + results[i].trace_offset, tr[i].data(), results[i].trace_len * 8);
	return internal;
}

extern "C" int gcgpu_extend(gcgpu_ctx* ctx, const uint8_t* seqIn, uint64_t seqBytes, const gcgpu_ext_item* items, uint32_t n, gcgpu_ext_result* results, uint64_t* traces, uint64_t trace_capacity, uint64_t* trace_used)
{
	if (seqIn) ctx->seqCopy.assign(seqIn, seqIn + seqBytes); // "resident" buffer of the real library
	ctx->dense.clear();
	bool internal = simK1(ctx, items, n, results, ctx->dense);
	uint64_t used = ctx->dense.size();
	*trace_used = used;
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

// ---- resident batch entry points: the same GC_HD functions the kernels of gcgpu_resident.inl run
#define SIM_NEED_POST(name) do { if (!ctx->havePost) { g_err = name ": the context was created without the original-node arrays"; return GCGPU_ERR_ARG; } } while (0)
extern "C" int gcgpu_load_reads(gcgpu_ctx* ctx, const char* chars, uint64_t char_bytes, const gcgpu_read* reads, uint32_t n)
{
	SIM_NEED_POST("gcgpu_load_reads");
	ctx->reads.assign((const GcReadDesc*)reads, (const GcReadDesc*)reads + n);
	ctx->cells.clear();
	for (auto& s : ctx->sets) { s.traces.clear(); s.pairs.clear(); }
	ctx->seqCopy.assign(2 * char_bytes + 16, 0);
	ctx->charsCopy.assign(chars, chars + char_bytes);
	for (uint32_t r = 0; r < n; r++)
		for (int32_t i = 0; i < reads[r].len; i++)
		{
			uint8_t m = ctx->codeTable[(uint8_t)chars[reads[r].char_offset + i]];
			ctx->seqCopy[2 * reads[r].char_offset + i] = m;
			ctx->seqCopy[2 * reads[r].char_offset + reads[r].len + (reads[r].len - 1 - i)] = gcComplementCode(m);
		}
	ctx->seqCopy.resize(2 * char_bytes);
	ctx->h2d += char_bytes + (uint64_t)n * sizeof(gcgpu_read);
	return 0;
}
extern "C" int gcgpu_set_seed_cells(gcgpu_ctx* ctx, const gcgpu_seed_cell* cells, uint64_t num_cells, const gcgpu_read* reads, uint32_t n)
{
	SIM_NEED_POST("gcgpu_set_seed_cells");
	if (n != ctx->reads.size()) { g_err = "gcgpu_set_seed_cells: read count"; return GCGPU_ERR_ARG; }
	ctx->reads.assign((const GcReadDesc*)reads, (const GcReadDesc*)reads + n);
	ctx->cells.assign((const GcSeedCell*)cells, (const GcSeedCell*)cells + num_cells);
	ctx->h2d += num_cells * sizeof(gcgpu_seed_cell) + (uint64_t)n * sizeof(gcgpu_read);
	return 0;
}
static int simExtend(gcgpu_ctx* ctx, int set, int append, int32_t fragLen, const gcgpu_seed_ext* exts, uint32_t n, gcgpu_pair_brief* brief, uint32_t* firstPair, uint64_t* columns)
{
	gcgpu_ctx::Set& S = ctx->sets[set];
	if (!append) { S.traces.clear(); S.pairs.clear(); }
	uint32_t first = (uint32_t)S.pairs.size();
	if (firstPair) *firstPair = first;
	std::vector<gcgpu_ext_item> items(2 * (size_t)n);
	for (uint32_t i = 0; i < n; i++)
	{
		if (exts[i].cell >= ctx->cells.size()) { g_err = "seed extension out of range"; return GCGPU_ERR_ARG; }
		const GcSeedCell& c = ctx->cells[exts[i].cell];
		const GcReadDesc& rd = ctx->reads[c.read];
		int32_t bl, fl, lp;
		gc_ext_lengths(c.seqPos, exts[i].frag_start, rd.len, fragLen, bl, fl, lp);
		int32_t seqLen = exts[i].frag_start < 0 ? rd.len : fragLen;
		if (lp < 0 || lp >= seqLen) { g_err = "seed extension out of range"; return GCGPU_ERR_ARG; }
		gcgpu_ext_item& b = items[2 * (size_t)i]; gcgpu_ext_item& f = items[2 * (size_t)i + 1];
		b.seq_offset = 2 * rd.charOffset + (uint64_t)rd.len + (uint64_t)(rd.len - c.seqPos); b.seq_len = bl; b.reserved = 0;
		gc_reverse_cell(ctx->pg, c.node, c.offset, b.node, b.offset);
		f.seq_offset = 2 * rd.charOffset + (uint64_t)c.seqPos + 1; f.seq_len = fl; f.node = c.node; f.offset = c.offset; f.reserved = 0;
	}
	std::vector<gcgpu_ext_result> res(2 * (size_t)n);
	simK1(ctx, items.data(), 2 * n, res.data(), S.traces);
	// test hook (this file is the TEST DOUBLE): GCGPU_SIM_FAIL=s1:<read> or s2:<read>:<position> turns the forward extension
	// of that read's whole-read seeds / of its fragments starting at or after <position> into "a state the reference asserts on" (GCGPU_ITEM_INTERNAL)
	if (const char* inj = getenv("GCGPU_SIM_FAIL"))
	{
		unsigned rd = 0; int fs = -1;
		bool s1 = sscanf(inj, "s1:%u", &rd) == 1, s2 = !s1 && sscanf(inj, "s2:%u:%d", &rd, &fs) == 2;
		for (uint32_t i = 0; i < n && (s1 || s2); i++)
		{
			const GcSeedCell& c = ctx->cells[exts[i].cell];
			if (c.read == rd && ((s1 && exts[i].frag_start < 0) || (s2 && exts[i].frag_start >= fs))) res[2 * (size_t)i + 1].status = GCGPU_ITEM_INTERNAL;
		}
	}
	uint64_t cols = 0;
	S.pairs.resize(first + n);
	for (uint32_t i = 0; i < n; i++)
	{
		const GcSeedCell& c = ctx->cells[exts[i].cell];
		const gcgpu_ext_result& rb = res[2 * (size_t)i]; const gcgpu_ext_result& rf = res[2 * (size_t)i + 1];
		GcPair p;
		p.bwdOff = rb.trace_offset; p.fwdOff = rf.trace_offset; p.bwdLen = rb.trace_len; p.fwdLen = rf.trace_len;
		p.seedPos = c.seqPos - (exts[i].frag_start < 0 ? 0 : exts[i].frag_start);
		p.cell = exts[i].cell; p.fragStart = exts[i].frag_start; p.read = c.read;
		gc_pair_finish(S.traces.data(), p, rb.status == GCGPU_ITEM_OK, rb.score, rb.status == GCGPU_ITEM_INTERNAL, rf.status == GCGPU_ITEM_OK, rf.score, rf.status == GCGPU_ITEM_INTERNAL);
		S.pairs[first + i] = p;
		if (brief) { brief[i].start = p.start; brief[i].end = p.end; brief[i].score = p.score; brief[i].flags = p.flags; }
		cols += rb.columns + rf.columns;
	}
	if (columns) *columns = cols;
	ctx->launches += 4;
	ctx->h2d += (uint64_t)n * sizeof(gcgpu_seed_ext);
	if (brief) ctx->d2h += (uint64_t)n * sizeof(gcgpu_pair_brief);
	return 0;
}
extern "C" int gcgpu_extend_seeds(gcgpu_ctx* ctx, int set, int append, int32_t frag_len, const gcgpu_seed_ext* exts, uint32_t n, gcgpu_pair_brief* brief,
	uint32_t* cover_bits, const uint64_t* cover_word_offsets, uint32_t* first_pair, uint64_t* columns)
{
	SIM_NEED_POST("gcgpu_extend_seeds");
	uint32_t first = 0;
	int rc = simExtend(ctx, set, append, frag_len, exts, n, brief, &first, columns);
	if (first_pair) *first_pair = first;
	if (rc || !cover_bits) return rc;
	gcgpu_ctx::Set& S = ctx->sets[set];
	#pragma omp parallel for schedule(dynamic, 4)
	for (uint32_t i = 0; i < n; i++)
	{
		const GcPair& p = S.pairs[first + i];
		const GcReadDesc& rd = ctx->reads[p.read];
		uint32_t* out = cover_bits + cover_word_offsets[i];
		for (uint32_t w = 0; w < (rd.numCells + 31) / 32; w++) out[w] = 0;
		if (!(p.flags & (GC_PAIR_BWD | GC_PAIR_FWD))) continue;
		int32_t shift = p.fragStart < 0 ? 0 : p.fragStart;
		for (uint32_t c = 0; c < rd.numCells; c++)
		{
			const GcSeedCell& cell = ctx->cells[rd.firstCell + c];
			if (gc_pair_has_cell(ctx->pg, S.traces.data(), p, cell.node, cell.offset, cell.seqPos - shift)) out[c >> 5] |= 1u << (c & 31);
		}
	}
	ctx->d2h += cover_word_offsets[n] * 4;
	return 0;
}
extern "C" int gcgpu_fragment_anchors(gcgpu_ctx* ctx, int set, int32_t frag_len, const gcgpu_seed_ext* exts, uint32_t num_exts, const gcgpu_frag* frags, uint32_t num_frags,
	uint32_t num_reads, gcgpu_read_anchors* per_read, uint64_t* columns)
{
	SIM_NEED_POST("gcgpu_fragment_anchors");
	int rc = simExtend(ctx, set, 0, frag_len, exts, num_exts, nullptr, nullptr, columns);
	if (rc) return rc;
	gcgpu_ctx::Set& S = ctx->sets[set];
	std::vector<uint8_t> kept(num_exts + 1, 0);
	std::vector<uint32_t> extended(num_frags, 0); std::vector<uint8_t> ok(num_frags, 1);
	#pragma omp parallel for schedule(dynamic, 64)
	for (uint32_t f = 0; f < num_frags; f++)
	{
		const gcgpu_frag& fr = frags[f];
		ok[f] = gc_fragment_filter(ctx->pg, S.traces.data(), ctx->cells.data(), (const GcSeedExt*)exts + fr.first_ext, S.pairs.data() + fr.first_ext, fr.num_exts, kept.data() + fr.first_ext, extended[f]) ? 1 : 0;
	}
	for (uint32_t r = 0; r < num_reads; r++) { per_read[r].anchors = 0; per_read[r].seeds_extended = 0; per_read[r].last_frag_extended = 0; per_read[r].dropped = 0; }
	for (uint32_t f = 0; f < num_frags; f++)
	{
		const gcgpu_frag& fr = frags[f];
		gcgpu_read_anchors& o = per_read[fr.read];
		if (!o.dropped && !ok[f]) o.dropped = 1;
		if (o.dropped) { for (uint32_t k = 0; k < fr.num_exts; k++) kept[fr.first_ext + k] = 0; continue; }
		o.seeds_extended += extended[f]; o.last_frag_extended = extended[f];
		for (uint32_t k = 0; k < fr.num_exts; k++) o.anchors += kept[fr.first_ext + k];
	}
	ctx->anchors.clear(); ctx->anchorMeta.clear(); ctx->anchorPaths.clear();
	ctx->readAnchorOff.assign((size_t)num_reads + 1, 0);
	for (uint32_t r = 0; r < num_reads; r++) ctx->readAnchorOff[r + 1] = ctx->readAnchorOff[r] + per_read[r].anchors;
	for (uint32_t i = 0; i < num_exts; i++)
	{
		if (!kept[i]) continue;
		const GcPair& p = S.pairs[i];
		uint32_t fo = 0, lo = 0;
		uint32_t len = gc_anchor_path(ctx->pg, S.traces.data(), p, nullptr, fo, lo);
		uint64_t po = ctx->anchorPaths.size();
		ctx->anchorPaths.resize(po + len);
		gc_anchor_path(ctx->pg, S.traces.data(), p, ctx->anchorPaths.data() + po, fo, lo);
		GcAnchor an; an.startNode = ctx->anchorPaths[po]; an.endNode = ctx->anchorPaths[po + len - 1]; an.x = p.fragStart; an.y = p.fragStart + frag_len - 1;
		ctx->anchors.push_back(an);
		gcgpu_chained_anchor m; m.first_offset = fo; m.last_offset = lo; m.path_first = po; m.path_len = len; m.reserved = 0;
		ctx->anchorMeta.push_back(m);
	}
	ctx->h2d += (uint64_t)num_frags * sizeof(gcgpu_frag);
	ctx->d2h += (uint64_t)num_reads * sizeof(gcgpu_read_anchors);
	ctx->launches += 5;
	return 0;
}
extern "C" int gcgpu_chain_resident(gcgpu_ctx* ctx, uint32_t num_reads, uint32_t* chain_len, int64_t* chain_score, uint64_t* chained_total, uint64_t* path_nodes_total)
{
	SIM_NEED_POST("gcgpu_chain_resident");
	if (num_reads + 1 != ctx->readAnchorOff.size()) { g_err = "gcgpu_chain_resident: read count"; return GCGPU_ERR_ARG; }
	std::vector<uint32_t> chain(ctx->anchors.size() + 1);
	int rc = gcgpu_chain(ctx, (const gcgpu_anchor*)ctx->anchors.data(), ctx->readAnchorOff.data(), num_reads, chain.data(), chain_len, chain_score);
	if (rc) return rc;
	ctx->chainedMeta.clear(); ctx->chainedPaths.clear();
	for (uint32_t r = 0; r < num_reads; r++)
		for (uint32_t i = 0; i < chain_len[r]; i++)
		{
			gcgpu_chained_anchor m = ctx->anchorMeta[ctx->readAnchorOff[r] + chain[ctx->readAnchorOff[r] + i]];
			uint64_t po = ctx->chainedPaths.size();
			ctx->chainedPaths.insert(ctx->chainedPaths.end(), ctx->anchorPaths.begin() + m.path_first, ctx->anchorPaths.begin() + m.path_first + m.path_len);
			m.path_first = po;
			ctx->chainedMeta.push_back(m);
		}
	*chained_total = ctx->chainedMeta.size(); *path_nodes_total = ctx->chainedPaths.size();
	ctx->d2h += (uint64_t)num_reads * 12;
	return 0;
}
extern "C" int gcgpu_fetch_chained(gcgpu_ctx* ctx, gcgpu_chained_anchor* anchors, uint32_t* path_nodes)
{
	if (!ctx->chainedMeta.empty()) memcpy(anchors, ctx->chainedMeta.data(), ctx->chainedMeta.size() * sizeof(gcgpu_chained_anchor));
	if (!ctx->chainedPaths.empty()) memcpy(path_nodes, ctx->chainedPaths.data(), ctx->chainedPaths.size() * 4);
	ctx->d2h += ctx->chainedMeta.size() * sizeof(gcgpu_chained_anchor) + ctx->chainedPaths.size() * 4;
	return 0;
}
extern "C" int gcgpu_nw_compose(gcgpu_ctx* ctx, const gcgpu_nw_piece* pieces, uint32_t n, const uint32_t* path_nodes, uint64_t num_path_nodes, uint64_t* piece_offsets)
{
	SIM_NEED_POST("gcgpu_nw_compose");
	ctx->nwCodes.clear();
	piece_offsets[0] = 0;
	for (uint32_t i = 0; i < n; i++)
	{
		const gcgpu_nw_piece& pc = pieces[i];
		size_t at = ctx->nwCodes.size();
		if (pc.kind == GCGPU_PIECE_READ)
		{
			const GcReadDesc& rd = ctx->reads[pc.index];
			ctx->nwCodes.resize(at + rd.len);
			for (int32_t k = 0; k < rd.len; k++) ctx->nwCodes[at + k] = gcK3CodeOf(ctx->seqCopy[2 * rd.charOffset + k]);
		}
		else if (pc.kind == GCGPU_PIECE_PAIR_PATH)
		{
			const gcgpu_ctx::Set& S = ctx->sets[pc.set];
			uint32_t len = gc_pair_path_string(ctx->pg, S.traces.data(), S.pairs[pc.index], nullptr);
			ctx->nwCodes.resize(at + len);
			gc_pair_path_string(ctx->pg, S.traces.data(), S.pairs[pc.index], ctx->nwCodes.data() + at);
		}
		else
		{
			uint32_t len = gc_node_path_string(ctx->pg, path_nodes + pc.first_node, pc.num_nodes, pc.first_offset, pc.last_offset, nullptr);
			ctx->nwCodes.resize(at + len);
			gc_node_path_string(ctx->pg, path_nodes + pc.first_node, pc.num_nodes, pc.first_offset, pc.last_offset, ctx->nwCodes.data() + at);
		}
		piece_offsets[i + 1] = ctx->nwCodes.size();
	}
	ctx->h2d += (uint64_t)n * sizeof(gcgpu_nw_piece) + num_path_nodes * 4;
	ctx->d2h += ((uint64_t)n + 1) * 8;
	return 0;
}
extern "C" int gcgpu_encode_alignments(gcgpu_ctx* ctx, int set, const uint32_t* pairs, uint32_t n, gcgpu_aln_tokens* out, uint64_t* tokens_used)
{
	SIM_NEED_POST("gcgpu_encode_alignments");
	const gcgpu_ctx::Set& S = ctx->sets[set];
	ctx->tokens.clear();
	for (uint32_t i = 0; i < n; i++)
	{
		if (pairs[i] >= S.pairs.size()) { g_err = "gcgpu_encode_alignments: pair out of range"; return GCGPU_ERR_ARG; }
		const GcPair& p = S.pairs[pairs[i]];
		GcPairTokenSrc src; src.pg = &ctx->pg; src.tr = S.traces.data(); src.p = &p; src.codes = ctx->seqCopy.data() + 2 * ctx->reads[p.read].charOffset;
		GcTokenCounts c = gc_tokenize(src, gc_pair_size(p), (uint32_t*)nullptr);
		size_t at = ctx->tokens.size();
		ctx->tokens.resize(at + c.tokens);
		gc_tokenize(src, gc_pair_size(p), ctx->tokens.data() + at);
		out[i].token_offset = at; out[i].num_tokens = c.tokens; out[i].matches = c.matches; out[i].steps = c.matches + c.mismatches + c.insertions + c.deletions; out[i].reserved = 0;
	}
	*tokens_used = ctx->tokens.size();
	ctx->d2h += (uint64_t)n * sizeof(gcgpu_aln_tokens);
	return 0;
}
extern "C" int gcgpu_fetch_tokens(gcgpu_ctx* ctx, uint32_t* tokens, uint64_t first, uint64_t count)
{
	if (first + count > ctx->tokens.size()) { g_err = "gcgpu_fetch_tokens: range"; return GCGPU_ERR_ARG; }
	if (count) memcpy(tokens, ctx->tokens.data() + first, count * 4);
	ctx->d2h += count * 4;
	return 0;
}
extern "C" int gcgpu_set_node_names(gcgpu_ctx* ctx, const uint32_t* name_offsets, const char* names)
{
	SIM_NEED_POST("gcgpu_set_node_names");
	size_t numOrig = ctx->origIds.size();
	int32_t maxId = -1;
	for (int32_t id : ctx->origIds) if (id > maxId) maxId = id;
	ctx->origIndexOfId.assign((size_t)maxId + 2, -1);
	for (size_t o = 0; o < numOrig; o++) if (ctx->origIds[o] >= 0) ctx->origIndexOfId[ctx->origIds[o]] = (int32_t)o;
	ctx->nameOff.assign(name_offsets, name_offsets + numOrig + 1);
	ctx->nameChars.assign((const uint8_t*)names, (const uint8_t*)names + name_offsets[numOrig]);
	ctx->nameChars.push_back(0);
	gcBuildGamTables(ctx->gamTables);
	ctx->haveNames = true;
	return 0;
}
extern "C" int gcgpu_encode_gam(gcgpu_ctx* ctx, int set, const gcgpu_gam_read* reads, uint32_t n, const gcgpu_gam_aln* alns, uint32_t num_alns,
	const char* names, uint64_t name_bytes, uint64_t* member_offsets, uint64_t* bytes_used)
{
	SIM_NEED_POST("gcgpu_encode_gam");
	if (!ctx->haveNames) { g_err = "gcgpu_encode_gam: gcgpu_set_node_names was not called"; return GCGPU_ERR_ARG; }
	const gcgpu_ctx::Set& S = ctx->sets[set];
	// tokens of every alignment (as gcgpu_encode_alignments), then one record per read with the device functions of gc_gam.cuh
	std::vector<uint32_t> tokens; std::vector<GcGamAln> ga(num_alns);
	for (uint32_t k = 0; k < num_alns; k++)
	{
		if (alns[k].pair >= S.pairs.size()) { g_err = "gcgpu_encode_gam: pair out of range"; return GCGPU_ERR_ARG; }
		const GcPair& p = S.pairs[alns[k].pair];
		GcPairTokenSrc src; src.pg = &ctx->pg; src.tr = S.traces.data(); src.p = &p; src.codes = ctx->seqCopy.data() + 2 * ctx->reads[p.read].charOffset;
		GcTokenCounts c = gc_tokenize(src, gc_pair_size(p), (uint32_t*)nullptr);
		size_t at = tokens.size();
		tokens.resize(at + c.tokens);
		gc_tokenize(src, gc_pair_size(p), tokens.data() + at);
		ga[k].tokenOff = at; ga[k].numTokens = c.tokens; ga[k].start = alns[k].start; ga[k].end = alns[k].end; ga[k].traceScore = alns[k].trace_score;
		ga[k].matches = c.matches; ga[k].steps = c.matches + c.mismatches + c.insertions + c.deletions;
	}
	GcNameTable nt; nt.origIndexOfId = ctx->origIndexOfId.data(); nt.nameOff = ctx->nameOff.data(); nt.nameChars = ctx->nameChars.data();
	ctx->gamOut.clear();
	member_offsets[0] = 0;
	for (uint32_t i = 0; i < n; i++)
	{
		const gcgpu_gam_read& rd = reads[i];
		uint32_t len = gc_gam_record_size(nt, ga.data() + rd.first_aln, rd.num_alns, tokens.data(), rd.name_len);
		std::vector<uint8_t> raw(len + 16), gz((size_t)len * 2 + 1024), wsBuf(gc_deflate_ws_bytes(len));
		GcDeflateWs ws; ws.tokens = (uint32_t*)wsBuf.data(); ws.tokenCap = len + 16;
		uint32_t written = gc_gam_write_record(nt, (const uint8_t*)ctx->charsCopy.data() + ctx->reads[rd.read].charOffset, (const uint8_t*)names + rd.name_offset, rd.name_len, ga.data() + rd.first_aln, rd.num_alns, tokens.data(), raw.data());
		uint32_t size = written == len ? gc_gzip_member(ctx->gamTables, raw.data(), len, ws, gz.data(), (uint32_t)gz.size()) : 0;
		ctx->gamOut.insert(ctx->gamOut.end(), gz.begin(), gz.begin() + size);
		member_offsets[i + 1] = ctx->gamOut.size();
	}
	*bytes_used = ctx->gamOut.size();
	ctx->launches += 6;
	ctx->h2d += (uint64_t)n * sizeof(gcgpu_gam_read) + (uint64_t)num_alns * sizeof(gcgpu_gam_aln) + name_bytes;
	ctx->d2h += ((uint64_t)n + 1) * 8;
	return 0;
}
extern "C" int gcgpu_fetch_gam(gcgpu_ctx* ctx, uint8_t* out, uint64_t first, uint64_t count)
{
	if (first + count > ctx->gamOut.size()) { g_err = "gcgpu_fetch_gam: range"; return GCGPU_ERR_ARG; }
	if (count) memcpy(out, ctx->gamOut.data() + first, count);
	ctx->d2h += count;
	return 0;
}
extern "C" void gcgpu_transfer_bytes(gcgpu_ctx* ctx, uint64_t* h2d, uint64_t* d2h) { if (h2d) *h2d = ctx->h2d; if (d2h) *d2h = ctx->d2h; }
