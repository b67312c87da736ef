// Resident batch entry points of libgcgpu (include/gcgpu.h, "resident batch"): the K1 traces of a batch stay in HBM
// and the reference's per-read logic between the DP kernels runs on them there (gc_post.cuh).  Included by gcgpu.cu.

struct GcTraceSet
{
	DevBuf traces; uint64_t used = 0;   // dense packed traces of every extension of the set
	DevBuf pairs; uint32_t numPairs = 0; // GcPair per seed extension
};

struct GcResident
{
	// per split node coordinate tables (GcPostGraph)
	int32_t* d_nodeIDs = nullptr; uint32_t* d_nodeOffset = nullptr; uint32_t* d_revFirst = nullptr; uint32_t* d_revCount = nullptr; uint32_t* d_revLast = nullptr; uint32_t* d_origNodes = nullptr;
	uint8_t* d_codeTable = nullptr;
	GcPostGraph pg;
	bool havePost = false;
	// the batch
	DevBuf chars, reads, cells;
	std::vector<GcReadDesc> hostReads; std::vector<GcSeedCell> hostCells;
	uint32_t numReads = 0; uint64_t numCells = 0;
	GcTraceSet sets[GCGPU_TRACE_SETS];
	DevBuf exts, brief, cover, coverOff, frags, kept, fragOut, readFrag, perRead, counts, offsets;
	// anchors of the batch (gcgpu_fragment_anchors) and their chains
	DevBuf anchors, anchorMeta, anchorPaths, readAnchorOff, chainWork, chainOut, chainedMeta, chainedPaths;
	uint64_t numAnchors = 0, numAnchorPathNodes = 0, numChained = 0, numChainedPathNodes = 0, maxAnchorsPerRead = 0;
	uint32_t anchorReads = 0;
	// K3 composition / tokens
	DevBuf pieces, pathNodes, tokenSlots, tokens, tokenMeta;
	uint64_t tokensUsed = 0;
	// GAM records on the device
	std::vector<int32_t> hostOrigIds;
	int32_t* d_origIndexOfId = nullptr; uint32_t* d_nameOff = nullptr; uint8_t* d_nameChars = nullptr; GcDeflateTables* d_gamTables = nullptr;
	bool haveNames = false;
	DevBuf gamIn, gamWork, gamArena, gamOut;
	uint64_t gamBytes = 0;
	void release()
	{
		cudaFree(d_nodeIDs); cudaFree(d_nodeOffset); cudaFree(d_revFirst); cudaFree(d_revCount); cudaFree(d_revLast); cudaFree(d_origNodes); cudaFree(d_codeTable); cudaFree(d_origIndexOfId); cudaFree(d_nameOff); cudaFree(d_nameChars); cudaFree(d_gamTables);
		DevBuf* all[] = { &chars, &reads, &cells, &exts, &brief, &cover, &coverOff, &frags, &kept, &fragOut, &readFrag, &perRead, &counts, &offsets, &anchors, &anchorMeta, &anchorPaths, &readAnchorOff,
			&chainWork, &chainOut, &chainedMeta, &chainedPaths, &pieces, &pathNodes, &tokenSlots, &tokens, &tokenMeta, &gamIn, &gamWork, &gamArena, &gamOut };
		for (DevBuf* b : all) b->release();
		for (auto& s : sets) { s.traces.release(); s.pairs.release(); }
	}
};

static_assert(sizeof(GcSeedCell) == sizeof(gcgpu_seed_cell) && sizeof(GcReadDesc) == sizeof(gcgpu_read) && sizeof(GcSeedExt) == sizeof(gcgpu_seed_ext), "resident batch struct layouts");

static int residentInitGraph(gcgpu_ctx* ctx, const gcgpu_graph* g)
{
	GcResident* R = ctx->resident;
	uint32_t N = g->num_nodes;
	if (!g->node_ids || !g->node_offset || !g->orig_ids || !g->orig_start || !g->orig_nodes || !g->orig_size || g->num_orig == 0) return GCGPU_OK; // not provided: the resident entry points stay unavailable
	GcRevTables rev;
	if (!gcBuildRevTables(g, rev)) return setError(GCGPU_ERR_ARG, "gcgpu_create: orig_nodes entry out of range");
	std::vector<uint32_t>& revFirst = rev.revFirst; std::vector<uint32_t>& revCount = rev.revCount; std::vector<uint32_t>& revLast = rev.revLast;
	uint8_t table[256];
	gcBuildCodeTable(table);
	cudaError_t err = cudaSuccess;
	auto chk = [&err](cudaError_t x) { if (err == cudaSuccess) err = x; };
	chk(uploadArray(g->node_ids, N, &R->d_nodeIDs));
	chk(uploadArray(g->node_offset, N, &R->d_nodeOffset));
	chk(uploadArray(revFirst.data(), N, &R->d_revFirst));
	chk(uploadArray(revCount.data(), N, &R->d_revCount));
	chk(uploadArray(revLast.data(), N, &R->d_revLast));
	chk(uploadArray(g->orig_nodes, g->orig_start[g->num_orig], &R->d_origNodes));
	chk(uploadArray(table, 256, &R->d_codeTable));
	if (err != cudaSuccess) return setError(err == cudaErrorMemoryAllocation ? GCGPU_ERR_NOMEM : GCGPU_ERR_CUDA, std::string("gcgpu_create: ") + cudaGetErrorString(err));
	R->pg.nodeIDs = R->d_nodeIDs; R->pg.nodeOffset = R->d_nodeOffset; R->pg.nodeLength = ctx->d_nodeLength; R->pg.nodeSeq = ctx->d_nodeSeq;
	R->pg.revFirst = R->d_revFirst; R->pg.revCount = R->d_revCount; R->pg.revLast = R->d_revLast; R->pg.origNodes = R->d_origNodes;
	R->havePost = true;
	R->hostOrigIds.assign(g->orig_ids, g->orig_ids + g->num_orig);
	return GCGPU_OK;
}

static int residentCreate(gcgpu_ctx* ctx, const gcgpu_graph* graph)
{
	ctx->resident = new GcResident();
	return residentInitGraph(ctx, graph);
}
static void residentDestroy(gcgpu_ctx* ctx)
{
	if (!ctx->resident) return;
	ctx->resident->release();
	delete ctx->resident;
	ctx->resident = nullptr;
}

#define GC_NEED_RESIDENT(name) do { if (!ctx) return setError(GCGPU_ERR_ARG, name ": null context"); if (!ctx->resident || !ctx->resident->havePost) return setError(GCGPU_ERR_ARG, name ": the context was created without the original-node arrays (gcgpu_graph.node_ids ...)"); } while (0)

// exclusive sum of count + 1 values (in[count] must be 0): out[count] = total
static int scanU64(gcgpu_ctx* ctx, const uint64_t* in, uint64_t* out, uint32_t count)
{
	size_t bytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)count + 1, ctx->stream);
	CUDA_TRY(ctx->copyDesc.ensure(bytes + 16));
	cub::DeviceScan::ExclusiveSum(ctx->copyDesc.p, bytes, in, out, (int)count + 1, ctx->stream);
	ctx->launches++;
	return GCGPU_OK;
}

// ------------------------------------------------------------------ reads
// one block per read: characters -> codes, forward and reverse complement
__global__ void __launch_bounds__(256) gc_reads_encode_kernel(const uint8_t* __restrict__ table, const uint8_t* __restrict__ chars, const GcReadDesc* __restrict__ reads, uint32_t n, uint8_t* __restrict__ codes)
{
	uint32_t r = blockIdx.x;
	if (r >= n) return;
	GcReadDesc rd = reads[r];
	const uint8_t* src = chars + rd.charOffset;
	uint8_t* fwd = codes + 2 * rd.charOffset;
	uint8_t* rc = fwd + rd.len;
	for (int32_t i = threadIdx.x; i < rd.len; i += blockDim.x)
	{
		uint8_t m = table[src[i]];
		fwd[i] = m;
		rc[rd.len - 1 - i] = (uint8_t)(((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3) | (m & 0x30));
	}
}

extern "C" int gcgpu_load_reads(gcgpu_ctx* ctx, const char* chars, uint64_t char_bytes, const gcgpu_read* reads, uint32_t n)
{
	GC_NEED_RESIDENT("gcgpu_load_reads");
	if ((!chars && char_bytes) || (!reads && n)) return setError(GCGPU_ERR_ARG, "gcgpu_load_reads: null argument");
	GcResident* R = ctx->resident;
	CUDA_TRY(cudaSetDevice(ctx->device));
	for (uint32_t i = 0; i < n; i++) if (reads[i].len < 0 || reads[i].char_offset + (uint64_t)reads[i].len > char_bytes || reads[i].len >= (1 << 24)) return setError(GCGPU_ERR_ARG, "gcgpu_load_reads: read " + std::to_string(i) + " out of range");
	R->hostReads.assign((const GcReadDesc*)reads, (const GcReadDesc*)reads + n);
	R->numReads = n;
	R->numCells = 0; R->hostCells.clear();
	for (auto& s : R->sets) { s.used = 0; s.numPairs = 0; }
	ctx->lastKernelMs = 0;
	if (n == 0) return GCGPU_OK;
	CUDA_TRY(R->chars.ensure(char_bytes + 16));
	CUDA_TRY(R->reads.ensure((size_t)n * sizeof(GcReadDesc)));
	CUDA_TRY(ctx->seqBuf.ensure(2 * char_bytes + 16));
	CUDA_TRY(gcCopy(ctx, R->chars.p, chars, char_bytes, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(gcCopy(ctx, R->reads.p, reads, (size_t)n * sizeof(GcReadDesc), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
	gc_reads_encode_kernel<<<n, 256, 0, ctx->stream>>>(R->d_codeTable, (const uint8_t*)R->chars.p, (const GcReadDesc*)R->reads.p, n, (uint8_t*)ctx->seqBuf.p);
	ctx->launches++;
	CUDA_TRY(cudaGetLastError());
	{ int prc = buildPlanes(ctx, 2 * char_bytes); if (prc != GCGPU_OK) return prc; }
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	float ms = 0;
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->lastKernelMs = ms;
	ctx->seqResident = 2 * char_bytes;
	return GCGPU_OK;
}

extern "C" int gcgpu_set_seed_cells(gcgpu_ctx* ctx, const gcgpu_seed_cell* cells, uint64_t num_cells, const gcgpu_read* reads, uint32_t n)
{
	GC_NEED_RESIDENT("gcgpu_set_seed_cells");
	GcResident* R = ctx->resident;
	if ((!cells && num_cells) || (!reads && n)) return setError(GCGPU_ERR_ARG, "gcgpu_set_seed_cells: null argument");
	if (n != R->numReads) return setError(GCGPU_ERR_ARG, "gcgpu_set_seed_cells: read count differs from gcgpu_load_reads");
	CUDA_TRY(cudaSetDevice(ctx->device));
	for (uint32_t r = 0; r < n; r++)
	{
		if (reads[r].char_offset != R->hostReads[r].charOffset || reads[r].len != R->hostReads[r].len) return setError(GCGPU_ERR_ARG, "gcgpu_set_seed_cells: read " + std::to_string(r) + " differs from gcgpu_load_reads");
		if ((uint64_t)reads[r].first_cell + reads[r].num_cells > num_cells) return setError(GCGPU_ERR_ARG, "gcgpu_set_seed_cells: cells of read " + std::to_string(r) + " out of range");
	}
	int bad = -1;
	#pragma omp parallel for schedule(static)
	for (uint64_t i = 0; i < num_cells; i++)
	{
		const gcgpu_seed_cell& c = cells[i];
		if (c.node >= ctx->numNodes || c.read >= n || c.seq_pos < 0 || c.seq_pos >= reads[c.read].len) { _Pragma("omp critical") bad = (int)i; }
	}
	if (bad >= 0) return setError(GCGPU_ERR_ARG, "gcgpu_set_seed_cells: cell " + std::to_string(bad) + " out of range");
	R->hostReads.assign((const GcReadDesc*)reads, (const GcReadDesc*)reads + n);
	R->hostCells.assign((const GcSeedCell*)cells, (const GcSeedCell*)cells + num_cells);
	R->numCells = num_cells;
	CUDA_TRY(R->cells.ensure(num_cells * sizeof(GcSeedCell) + 16));
	CUDA_TRY(gcCopy(ctx, R->cells.p, cells, num_cells * sizeof(GcSeedCell), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(gcCopy(ctx, R->reads.p, reads, (size_t)n * sizeof(GcReadDesc), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	return GCGPU_OK;
}

// ------------------------------------------------------------------ seed extensions
// the two K1 work items of every seed extension (getTwoDirectionalTrace, GraphAligner.h:480-525): items[2i] backward, items[2i+1] forward
__global__ void gc_ext_items_kernel(GcPostGraph pg, const GcSeedCell* __restrict__ cells, const GcReadDesc* __restrict__ reads, const GcSeedExt* __restrict__ exts, uint32_t n, int32_t fragLen, gcgpu_ext_item* __restrict__ items)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	GcSeedExt e = exts[i];
	GcSeedCell c = cells[e.cell];
	GcReadDesc rd = reads[c.read];
	int32_t bl, fl, lp;
	gc_ext_lengths(c.seqPos, e.fragStart, rd.len, fragLen, bl, fl, lp);
	uint64_t fwdCodes = 2 * rd.charOffset, rcCodes = fwdCodes + (uint64_t)rd.len;
	gcgpu_ext_item b, f;
	// backward: revcomp(sequence)'s last localPos characters = rc(read)[len - seqPos, len - fragStart)
	b.seq_offset = rcCodes + (uint64_t)(rd.len - c.seqPos); b.seq_len = bl; b.reserved = 0;
	gc_reverse_cell(pg, c.node, c.offset, b.node, b.offset);
	f.seq_offset = fwdCodes + (uint64_t)c.seqPos + 1; f.seq_len = fl; f.node = c.node; f.offset = c.offset; f.reserved = 0;
	items[2 * (size_t)i] = b;
	items[2 * (size_t)i + 1] = f;
}

// getAlignmentFromSeed (GraphAligner.h:567-626): the pair record of every seed extension from its two K1 results
__global__ void gc_pairs_kernel(const GcSeedCell* __restrict__ cells, const GcSeedExt* __restrict__ exts, uint32_t n, const gcgpu_ext_result* __restrict__ pub, const uint64_t* __restrict__ tr,
	GcPair* __restrict__ pairs, gcgpu_pair_brief* __restrict__ brief, unsigned long long* columns)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long cols = 0;
	if (i < n)
	{
		GcSeedExt e = exts[i];
		GcSeedCell c = cells[e.cell];
		gcgpu_ext_result rb = pub[2 * (size_t)i], rf = pub[2 * (size_t)i + 1];
		GcPair p;
		p.bwdOff = rb.trace_offset; p.fwdOff = rf.trace_offset; p.bwdLen = rb.trace_len; p.fwdLen = rf.trace_len;
		p.seedPos = c.seqPos - (e.fragStart < 0 ? 0 : e.fragStart);
		p.cell = e.cell; p.fragStart = e.fragStart; p.read = c.read;
		gc_pair_finish(tr, p, rb.status == GCGPU_ITEM_OK, rb.score, rb.status == GCGPU_ITEM_INTERNAL, rf.status == GCGPU_ITEM_OK, rf.score, rf.status == GCGPU_ITEM_INTERNAL);
		pairs[i] = p;
		if (brief) { gcgpu_pair_brief b; b.start = p.start; b.end = p.end; b.score = p.score; b.flags = p.flags; brief[i] = b; }
		cols = rb.columns + rf.columns;
	}
	for (int off = 16; off > 0; off >>= 1) cols += __shfl_down_sync(0xFFFFFFFFu, cols, off);
	if ((threadIdx.x & 31) == 0 && cols) atomicAdd(columns, cols);
}

// exactAlignmentPart of every cell of the read against one alignment: warp per seed extension, lanes over the read's cells
__global__ void __launch_bounds__(128) gc_cover_kernel(GcPostGraph pg, const GcSeedCell* __restrict__ cells, const GcReadDesc* __restrict__ reads, const GcPair* __restrict__ pairs, uint32_t n,
	const uint64_t* __restrict__ tr, const uint64_t* __restrict__ wordOff, uint32_t* __restrict__ bits)
{
	uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (w >= n) return;
	GcPair p = pairs[w];
	GcReadDesc rd = reads[p.read];
	uint32_t* out = bits + wordOff[w];
	bool live = (p.flags & (GC_PAIR_BWD | GC_PAIR_FWD)) != 0;
	int32_t shift = p.fragStart < 0 ? 0 : p.fragStart;
	for (uint32_t base = 0; base < rd.numCells; base += 32)
	{
		uint32_t ci = base + lane;
		bool hit = false;
		if (live && ci < rd.numCells)
		{
			GcSeedCell c = cells[rd.firstCell + ci];
			hit = gc_pair_has_cell(pg, tr, p, c.node, c.offset, c.seqPos - shift);
		}
		uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
		if (lane == 0) out[base >> 5] = m;
	}
}

// common part of gcgpu_extend_seeds and gcgpu_fragment_anchors: items on the device, K1, traces appended to the set, pair records
static int residentExtend(gcgpu_ctx* ctx, int set, int append, int32_t fragLen, const gcgpu_seed_ext* exts, uint32_t n, bool uniformShort, gcgpu_pair_brief* brief, uint32_t* firstPair, uint64_t* columns)
{
	GcResident* R = ctx->resident;
	GcTraceSet& S = R->sets[set];
	if (!append) { S.used = 0; S.numPairs = 0; }
	if (firstPair) *firstPair = S.numPairs;
	if (columns) *columns = 0;
	if (n == 0) return GCGPU_OK;
	// lengths of the work items (host: slab layout).  Fragment batches are uniform: every item is shorter than the fragment.
	std::vector<int32_t> lens;
	{
		int bad = -1;
		if (!uniformShort) lens.resize(2 * (size_t)n);
		#pragma omp parallel for schedule(static)
		for (uint32_t i = 0; i < n; i++)
		{
			if (exts[i].cell >= R->numCells) { _Pragma("omp critical") bad = (int)i; continue; }
			const GcSeedCell& c = R->hostCells[exts[i].cell];
			const GcReadDesc& rd = R->hostReads[c.read];
			int32_t seqLen = exts[i].frag_start < 0 ? rd.len : fragLen;
			int32_t lp = c.seqPos - (exts[i].frag_start < 0 ? 0 : exts[i].frag_start);
			if (lp < 0 || lp >= seqLen || (exts[i].frag_start >= 0 && exts[i].frag_start + fragLen > rd.len)) { _Pragma("omp critical") bad = (int)i; continue; }
			if (!uniformShort) { int32_t bl, fl, l2; gc_ext_lengths(c.seqPos, exts[i].frag_start, rd.len, fragLen, bl, fl, l2); lens[2 * (size_t)i] = bl; lens[2 * (size_t)i + 1] = fl; }
		}
		if (bad >= 0) return setError(GCGPU_ERR_ARG, "seed extension " + std::to_string(bad) + " out of range");
	}
	CUDA_TRY(R->exts.ensure((size_t)n * sizeof(GcSeedExt)));
	CUDA_TRY(ctx->itemsBuf.ensure(2 * (size_t)n * sizeof(gcgpu_ext_item)));
	CUDA_TRY(gcCopy(ctx, R->exts.p, exts, (size_t)n * sizeof(GcSeedExt), cudaMemcpyHostToDevice, ctx->stream));
	gc_ext_items_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(R->pg, (const GcSeedCell*)R->cells.p, (const GcReadDesc*)R->reads.p, (const GcSeedExt*)R->exts.p, n, fragLen, (gcgpu_ext_item*)ctx->itemsBuf.p);
	ctx->launches++;
	GcK1Run run;
	int rc = k1Run(ctx, (const gcgpu_ext_item*)ctx->itemsBuf.p, uniformShort ? nullptr : lens.data(), 2 * n, fragLen - 1, run);
	if (rc != GCGPU_OK) return rc;
	uint64_t used = 0;
	rc = k1Gather(ctx, run, S.traces, S.used, &used);
	if (rc != GCGPU_OK) return rc;
	S.used += used;
	CUDA_TRY(growKeep(ctx, S.pairs, ((size_t)S.numPairs + n) * sizeof(GcPair), (size_t)S.numPairs * sizeof(GcPair)));
	CUDA_TRY(R->brief.ensure((size_t)n * sizeof(gcgpu_pair_brief) + 16));
	unsigned long long* dCols = (unsigned long long*)((uint8_t*)R->brief.p + (size_t)n * sizeof(gcgpu_pair_brief));
	CUDA_TRY(cudaMemsetAsync(dCols, 0, 8, ctx->stream));
	gc_pairs_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>((const GcSeedCell*)R->cells.p, (const GcSeedExt*)R->exts.p, n, run.dPub, (const uint64_t*)S.traces.p,
		(GcPair*)S.pairs.p + S.numPairs, brief ? (gcgpu_pair_brief*)R->brief.p : nullptr, dCols);
	ctx->launches++;
	CUDA_TRY(cudaGetLastError());
	if (brief) CUDA_TRY(gcCopy(ctx, brief, R->brief.p, (size_t)n * sizeof(gcgpu_pair_brief), cudaMemcpyDeviceToHost, ctx->stream));
	unsigned long long cols = 0;
	CUDA_TRY(gcCopy(ctx, &cols, dCols, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	if (columns) *columns = cols;
	S.numPairs += n;
	return GCGPU_OK;
}

extern "C" int gcgpu_extend_seeds(gcgpu_ctx* ctx, int set, int append, int32_t frag_len, const gcgpu_seed_ext* exts, uint32_t n, gcgpu_pair_brief* brief,
	uint32_t* cover_bits, const uint64_t* cover_word_offsets, uint32_t* first_pair, uint64_t* columns)
{
	GC_NEED_RESIDENT("gcgpu_extend_seeds");
	if (set < 0 || set >= GCGPU_TRACE_SETS || (!exts && n) || (cover_bits && !cover_word_offsets)) return setError(GCGPU_ERR_ARG, "gcgpu_extend_seeds: bad argument");
	CUDA_TRY(cudaSetDevice(ctx->device));
	ctx->lastKernelMs = 0;
	GcResident* R = ctx->resident;
	uint32_t first = 0;
	int rc = residentExtend(ctx, set, append, frag_len, exts, n, false, brief, &first, columns);
	if (first_pair) *first_pair = first;
	if (rc != GCGPU_OK || n == 0 || !cover_bits) return rc;
	uint64_t words = cover_word_offsets[n];
	if (words == 0) return GCGPU_OK;
	CUDA_TRY(R->coverOff.ensure(((size_t)n + 1) * 8));
	CUDA_TRY(R->cover.ensure(words * 4 + 16));
	CUDA_TRY(gcCopy(ctx, R->coverOff.p, cover_word_offsets, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
	GcTraceSet& S = R->sets[set];
	gc_cover_kernel<<<(n + 3) / 4, 128, 0, ctx->stream>>>(R->pg, (const GcSeedCell*)R->cells.p, (const GcReadDesc*)R->reads.p, (const GcPair*)S.pairs.p + first, n, (const uint64_t*)S.traces.p, (const uint64_t*)R->coverOff.p, (uint32_t*)R->cover.p);
	ctx->launches++;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	CUDA_TRY(gcCopy(ctx, cover_bits, R->cover.p, words * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	float ms = 0;
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->lastKernelMs += ms;
	GC_TRACE_MS("s1 seed coverage", n);
	return GCGPU_OK;
}

// ------------------------------------------------------------------ S2: fragment seed loops + anchors
struct GcFragOut { uint32_t extended; uint32_t ok; };
__global__ void gc_frag_filter_kernel(GcPostGraph pg, const uint64_t* __restrict__ tr, const GcSeedCell* __restrict__ cells, const GcSeedExt* __restrict__ exts, const GcPair* __restrict__ pairs,
	const gcgpu_frag* __restrict__ frags, uint32_t n, uint8_t* __restrict__ kept, GcFragOut* __restrict__ out)
{
	uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n) return;
	gcgpu_frag fr = frags[f];
	GcFragOut o;
	o.ok = gc_fragment_filter(pg, tr, cells, exts + fr.first_ext, pairs + fr.first_ext, fr.num_exts, kept + fr.first_ext, o.extended) ? 1u : 0u;
	out[f] = o;
}
// per read, fragments in order: the first fragment that hit an assertion-class state ends the read's fragment loop (cont = true)
__global__ void gc_read_anchors_kernel(const gcgpu_frag* __restrict__ frags, const uint32_t* __restrict__ readFrag, uint32_t numReads, const GcFragOut* __restrict__ fragOut, uint8_t* __restrict__ kept,
	gcgpu_read_anchors* __restrict__ perRead, uint64_t* __restrict__ readAnchorCount)
{
	uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r > numReads) return;
	if (r == numReads) { readAnchorCount[r] = 0; return; }
	gcgpu_read_anchors o; o.anchors = 0; o.seeds_extended = 0; o.last_frag_extended = 0; o.dropped = 0;
	bool dead = false;
	for (uint32_t f = readFrag[r]; f < readFrag[r + 1]; f++)
	{
		gcgpu_frag fr = frags[f];
		if (!dead && !fragOut[f].ok) { dead = true; o.dropped = 1; }
		if (dead) { for (uint32_t k = 0; k < fr.num_exts; k++) kept[fr.first_ext + k] = 0; continue; }
		o.seeds_extended += fragOut[f].extended; o.last_frag_extended = fragOut[f].extended;
		for (uint32_t k = 0; k < fr.num_exts; k++) o.anchors += kept[fr.first_ext + k];
	}
	perRead[r] = o;
	readAnchorCount[r] = o.anchors;
}
__global__ void gc_anchor_count_kernel(GcPostGraph pg, const uint64_t* __restrict__ tr, const GcPair* __restrict__ pairs, const uint8_t* __restrict__ kept, uint32_t n, uint64_t* __restrict__ isAnchor, uint64_t* __restrict__ pathLen)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n) return;
	if (i == n) { isAnchor[i] = 0; pathLen[i] = 0; return; }
	uint32_t fo, lo;
	bool k = kept[i] != 0;
	isAnchor[i] = k ? 1 : 0;
	pathLen[i] = k ? gc_anchor_path(pg, tr, pairs[i], nullptr, fo, lo) : 0;
}
__global__ void gc_anchor_write_kernel(GcPostGraph pg, const uint64_t* __restrict__ tr, const GcPair* __restrict__ pairs, const uint8_t* __restrict__ kept, uint32_t n, int32_t fragLen,
	const uint64_t* __restrict__ anchorIdx, const uint64_t* __restrict__ pathOff, GcAnchor* __restrict__ anchors, gcgpu_chained_anchor* __restrict__ meta, uint32_t* __restrict__ paths)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !kept[i]) return;
	GcPair p = pairs[i];
	uint64_t a = anchorIdx[i], po = pathOff[i];
	uint32_t fo = 0, lo = 0;
	uint32_t len = gc_anchor_path(pg, tr, p, paths + po, fo, lo);
	GcAnchor an; an.startNode = paths[po]; an.endNode = paths[po + len - 1]; an.x = p.fragStart; an.y = p.fragStart + fragLen - 1;
	anchors[a] = an;
	gcgpu_chained_anchor m; m.first_offset = fo; m.last_offset = lo; m.path_first = po; m.path_len = len; m.reserved = 0;
	meta[a] = m;
}

extern "C" int gcgpu_fragment_anchors(gcgpu_ctx* ctx, int set, int32_t frag_len, const gcgpu_seed_ext* exts, uint32_t num_exts, const gcgpu_frag* frags, uint32_t num_frags,
	uint32_t num_reads, gcgpu_read_anchors* per_read, uint64_t* columns)
{
	GC_NEED_RESIDENT("gcgpu_fragment_anchors");
	GcResident* R = ctx->resident;
	if (set < 0 || set >= GCGPU_TRACE_SETS || (!exts && num_exts) || (!frags && num_frags) || (!per_read && num_reads) || frag_len < 2 || frag_len > GC_K1_LONG_ITEM) return setError(GCGPU_ERR_ARG, "gcgpu_fragment_anchors: bad argument");
	if (num_reads != R->numReads) return setError(GCGPU_ERR_ARG, "gcgpu_fragment_anchors: read count differs from gcgpu_load_reads");
	CUDA_TRY(cudaSetDevice(ctx->device));
	ctx->lastKernelMs = 0;
	R->numAnchors = 0; R->numAnchorPathNodes = 0; R->anchorReads = num_reads;
	if (columns) *columns = 0;
	// fragments of read r: [readFrag[r], readFrag[r+1])
	std::vector<uint32_t> readFrag((size_t)num_reads + 1, 0);
	{
		uint32_t prevRead = 0; int32_t prevStart = -1; uint64_t nextExt = 0;
		for (uint32_t f = 0; f < num_frags; f++)
		{
			const gcgpu_frag& fr = frags[f];
			if (fr.read >= num_reads || fr.read < prevRead || (fr.read == prevRead && f > 0 && fr.start <= prevStart) || fr.first_ext != nextExt || (uint64_t)fr.first_ext + fr.num_exts > num_exts)
				return setError(GCGPU_ERR_ARG, "gcgpu_fragment_anchors: fragment " + std::to_string(f) + " out of order or out of range");
			for (uint32_t k = 0; k < fr.num_exts; k++) if (exts[fr.first_ext + k].frag_start != fr.start) return setError(GCGPU_ERR_ARG, "gcgpu_fragment_anchors: seed extension of fragment " + std::to_string(f) + " has another frag_start");
			prevRead = fr.read; prevStart = fr.start; nextExt += fr.num_exts;
			readFrag[fr.read + 1]++;
		}
		if (nextExt != num_exts) return setError(GCGPU_ERR_ARG, "gcgpu_fragment_anchors: seed extensions outside the fragments");
		for (uint32_t r = 0; r < num_reads; r++) readFrag[r + 1] += readFrag[r];
	}
	CUDA_TRY(R->perRead.ensure((size_t)num_reads * sizeof(gcgpu_read_anchors) + 16));
	CUDA_TRY(R->readAnchorOff.ensure(2 * ((size_t)num_reads + 1) * 8));
	uint64_t* dReadCount = (uint64_t*)R->readAnchorOff.p + ((size_t)num_reads + 1);
	uint64_t* dReadOff = (uint64_t*)R->readAnchorOff.p;
	float kernelMs = 0;
	if (num_exts)
	{
		int rc = residentExtend(ctx, set, 0, frag_len, exts, num_exts, true, nullptr, nullptr, columns);
		if (rc != GCGPU_OK) return rc;
		kernelMs = ctx->lastKernelMs;
	}
	GcTraceSet& S = R->sets[set];
	CUDA_TRY(R->frags.ensure((size_t)num_frags * sizeof(gcgpu_frag) + 16));
	CUDA_TRY(R->readFrag.ensure(((size_t)num_reads + 1) * 4));
	CUDA_TRY(R->kept.ensure((size_t)num_exts + 16));
	CUDA_TRY(R->fragOut.ensure((size_t)num_frags * sizeof(GcFragOut) + 16));
	CUDA_TRY(R->counts.ensure(2 * ((size_t)num_exts + 1) * 8));
	CUDA_TRY(R->offsets.ensure(2 * ((size_t)num_exts + 1) * 8));
	uint64_t* dIsAnchor = (uint64_t*)R->counts.p; uint64_t* dPathLen = dIsAnchor + ((size_t)num_exts + 1);
	uint64_t* dAnchorIdx = (uint64_t*)R->offsets.p; uint64_t* dPathOff = dAnchorIdx + ((size_t)num_exts + 1);
	CUDA_TRY(gcCopy(ctx, R->frags.p, frags, (size_t)num_frags * sizeof(gcgpu_frag), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(gcCopy(ctx, R->readFrag.p, readFrag.data(), ((size_t)num_reads + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
	if (num_frags)
	{
		gc_frag_filter_kernel<<<(num_frags + 127) / 128, 128, 0, ctx->stream>>>(R->pg, (const uint64_t*)S.traces.p, (const GcSeedCell*)R->cells.p, (const GcSeedExt*)R->exts.p, (const GcPair*)S.pairs.p,
			(const gcgpu_frag*)R->frags.p, num_frags, (uint8_t*)R->kept.p, (GcFragOut*)R->fragOut.p);
		ctx->launches++;
	}
	gc_read_anchors_kernel<<<(num_reads + 1 + 127) / 128, 128, 0, ctx->stream>>>((const gcgpu_frag*)R->frags.p, (const uint32_t*)R->readFrag.p, num_reads, (const GcFragOut*)R->fragOut.p, (uint8_t*)R->kept.p,
		(gcgpu_read_anchors*)R->perRead.p, dReadCount);
	gc_anchor_count_kernel<<<(num_exts + 1 + 127) / 128, 128, 0, ctx->stream>>>(R->pg, (const uint64_t*)S.traces.p, (const GcPair*)S.pairs.p, (const uint8_t*)R->kept.p, num_exts, dIsAnchor, dPathLen);
	ctx->launches += 2;
	int rc = scanU64(ctx, dReadCount, dReadOff, num_reads); if (rc != GCGPU_OK) return rc;
	rc = scanU64(ctx, dIsAnchor, dAnchorIdx, num_exts); if (rc != GCGPU_OK) return rc;
	rc = scanU64(ctx, dPathLen, dPathOff, num_exts); if (rc != GCGPU_OK) return rc;
	CUDA_TRY(cudaGetLastError());
	uint64_t totals[2] = { 0, 0 };
	CUDA_TRY(gcCopy(ctx, &totals[0], dAnchorIdx + num_exts, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcCopy(ctx, &totals[1], dPathOff + num_exts, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcCopy(ctx, per_read, R->perRead.p, (size_t)num_reads * sizeof(gcgpu_read_anchors), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	R->numAnchors = totals[0]; R->numAnchorPathNodes = totals[1];
	R->maxAnchorsPerRead = 0;
	for (uint32_t r = 0; r < num_reads; r++) if (per_read[r].anchors > R->maxAnchorsPerRead) R->maxAnchorsPerRead = per_read[r].anchors;
	CUDA_TRY(R->anchors.ensure(R->numAnchors * sizeof(GcAnchor) + 16));
	CUDA_TRY(R->anchorMeta.ensure(R->numAnchors * sizeof(gcgpu_chained_anchor) + 16));
	CUDA_TRY(R->anchorPaths.ensure(R->numAnchorPathNodes * 4 + 16));
	if (num_exts)
	{
		gc_anchor_write_kernel<<<(num_exts + 127) / 128, 128, 0, ctx->stream>>>(R->pg, (const uint64_t*)S.traces.p, (const GcPair*)S.pairs.p, (const uint8_t*)R->kept.p, num_exts, frag_len,
			dAnchorIdx, dPathOff, (GcAnchor*)R->anchors.p, (gcgpu_chained_anchor*)R->anchorMeta.p, (uint32_t*)R->anchorPaths.p);
		ctx->launches++;
	}
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	float ms = 0;
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	GC_TRACE_MS("s2 filter + anchors", num_exts);
	ctx->lastKernelMs = kernelMs + ms;
	return GCGPU_OK;
}

// ------------------------------------------------------------------ K2 on the resident anchors
__global__ void gc_chained_count_kernel(const uint64_t* __restrict__ readOff, uint32_t numReads, const uint32_t* __restrict__ chain, const uint32_t* __restrict__ chainLen, const gcgpu_chained_anchor* __restrict__ meta,
	uint64_t* __restrict__ nChained, uint64_t* __restrict__ nPath)
{
	uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r > numReads) return;
	if (r == numReads) { nChained[r] = 0; nPath[r] = 0; return; }
	uint64_t base = readOff[r], total = 0;
	uint32_t n = chainLen[r];
	for (uint32_t i = 0; i < n; i++) total += meta[base + chain[base + i]].path_len;
	nChained[r] = n; nPath[r] = total;
}
__global__ void gc_chained_write_kernel(const uint64_t* __restrict__ readOff, uint32_t numReads, const uint32_t* __restrict__ chain, const uint32_t* __restrict__ chainLen, const gcgpu_chained_anchor* __restrict__ meta,
	const uint32_t* __restrict__ paths, const uint64_t* __restrict__ outIdx, const uint64_t* __restrict__ outPath, gcgpu_chained_anchor* __restrict__ outMeta, uint32_t* __restrict__ outPaths)
{
	uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (w >= numReads) return;
	uint64_t base = readOff[w], oi = outIdx[w], op = outPath[w];
	uint32_t n = chainLen[w];
	for (uint32_t i = 0; i < n; i++)
	{
		gcgpu_chained_anchor m = meta[base + chain[base + i]];
		for (uint32_t k = lane; k < m.path_len; k += 32) outPaths[op + k] = paths[m.path_first + k];
		if (lane == 0) { gcgpu_chained_anchor o = m; o.path_first = op; outMeta[oi + i] = o; }
		op += m.path_len;
	}
}

extern "C" int gcgpu_chain_resident(gcgpu_ctx* ctx, uint32_t num_reads, uint32_t* chain_len, int64_t* chain_score, uint64_t* chained_total, uint64_t* path_nodes_total)
{
	GC_NEED_RESIDENT("gcgpu_chain_resident");
	GcResident* R = ctx->resident;
	if ((!chain_len && num_reads) || (!chain_score && num_reads) || !chained_total || !path_nodes_total) return setError(GCGPU_ERR_ARG, "gcgpu_chain_resident: null argument");
	if (!ctx->haveMpc) return setError(GCGPU_ERR_ARG, "gcgpu_chain_resident: the context was created without an MPC index");
	if (num_reads != R->anchorReads) return setError(GCGPU_ERR_ARG, "gcgpu_chain_resident: read count differs from gcgpu_fragment_anchors");
	*chained_total = 0; *path_nodes_total = 0;
	R->numChained = 0; R->numChainedPathNodes = 0;
	ctx->lastKernelMs = 0;
	if (num_reads == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	uint64_t total = R->numAnchors;
	// order | score | pred | chain | chainLen | chainScore | counts (2 x (R+1)) | offsets (2 x (R+1))
	size_t offOrd = 0, offSc = alignUp(offOrd + total * 4, 128), offPr = alignUp(offSc + total * 4, 128), offCh = alignUp(offPr + total * 4, 128), offLen = alignUp(offCh + total * 4, 128);
	size_t offScore = alignUp(offLen + (size_t)num_reads * 4, 128), offCnt = alignUp(offScore + (size_t)num_reads * 8, 128), offOff = alignUp(offCnt + 2 * ((size_t)num_reads + 1) * 8, 128), end = offOff + 2 * ((size_t)num_reads + 1) * 8;
	CUDA_TRY(R->chainWork.ensure(end));
	uint8_t* A = (uint8_t*)R->chainWork.p;
	uint64_t* dCntC = (uint64_t*)(A + offCnt); uint64_t* dCntP = dCntC + ((size_t)num_reads + 1);
	uint64_t* dOffC = (uint64_t*)(A + offOff); uint64_t* dOffP = dOffC + ((size_t)num_reads + 1);
	const uint64_t* dReadOff = (const uint64_t*)R->readAnchorOff.p;
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
	{
		int krc = k2Run(ctx, (const GcAnchor*)R->anchors.p, dReadOff, num_reads, total, R->maxAnchorsPerRead,
			(uint32_t*)(A + offOrd), (int32_t*)(A + offSc), (int32_t*)(A + offPr), (uint32_t*)(A + offCh), (uint32_t*)(A + offLen), (int64_t*)(A + offScore));
		if (krc != GCGPU_OK) return krc;
	}
	gc_chained_count_kernel<<<(num_reads + 1 + 127) / 128, 128, 0, ctx->stream>>>(dReadOff, num_reads, (const uint32_t*)(A + offCh), (const uint32_t*)(A + offLen), (const gcgpu_chained_anchor*)R->anchorMeta.p, dCntC, dCntP);
	ctx->launches += 1;
	int rc = scanU64(ctx, dCntC, dOffC, num_reads); if (rc != GCGPU_OK) return rc;
	rc = scanU64(ctx, dCntP, dOffP, num_reads); if (rc != GCGPU_OK) return rc;
	CUDA_TRY(cudaGetLastError());
	uint64_t totals[2] = { 0, 0 };
	CUDA_TRY(gcCopy(ctx, &totals[0], dOffC + num_reads, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcCopy(ctx, &totals[1], dOffP + num_reads, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcCopy(ctx, chain_len, A + offLen, (size_t)num_reads * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcCopy(ctx, chain_score, A + offScore, (size_t)num_reads * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	R->numChained = totals[0]; R->numChainedPathNodes = totals[1];
	CUDA_TRY(R->chainedMeta.ensure(R->numChained * sizeof(gcgpu_chained_anchor) + 16));
	CUDA_TRY(R->chainedPaths.ensure(R->numChainedPathNodes * 4 + 16));
	gc_chained_write_kernel<<<(num_reads + 3) / 4, 128, 0, ctx->stream>>>(dReadOff, num_reads, (const uint32_t*)(A + offCh), (const uint32_t*)(A + offLen), (const gcgpu_chained_anchor*)R->anchorMeta.p,
		(const uint32_t*)R->anchorPaths.p, dOffC, dOffP, (gcgpu_chained_anchor*)R->chainedMeta.p, (uint32_t*)R->chainedPaths.p);
	ctx->launches++;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	float ms = 0;
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->lastKernelMs = ms;
	GC_TRACE_MS("k2 chain (resident anchors)", num_reads);
	*chained_total = R->numChained; *path_nodes_total = R->numChainedPathNodes;
	return GCGPU_OK;
}

extern "C" int gcgpu_fetch_chained(gcgpu_ctx* ctx, gcgpu_chained_anchor* anchors, uint32_t* path_nodes)
{
	GC_NEED_RESIDENT("gcgpu_fetch_chained");
	GcResident* R = ctx->resident;
	if ((R->numChained && !anchors) || (R->numChainedPathNodes && !path_nodes)) return setError(GCGPU_ERR_ARG, "gcgpu_fetch_chained: null argument");
	CUDA_TRY(cudaSetDevice(ctx->device));
	CUDA_TRY(gcCopy(ctx, anchors, R->chainedMeta.p, R->numChained * sizeof(gcgpu_chained_anchor), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcCopy(ctx, path_nodes, R->chainedPaths.p, R->numChainedPathNodes * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	return GCGPU_OK;
}

// ------------------------------------------------------------------ K3 sequence buffer from pieces
// resident read code -> the code K3 compares (0..3 = the upper-case letters A C G T, 4 = anything else)
__device__ __forceinline__ uint8_t gc_k3_code_of(uint8_t m)
{
	return (m & 0x30) ? (uint8_t)4 : (m == 1 ? (uint8_t)0 : m == 2 ? (uint8_t)1 : m == 4 ? (uint8_t)2 : m == 8 ? (uint8_t)3 : (uint8_t)4);
}
// One block per piece.  WRITE = false: lens[i] = length of the piece (flags[i] = 1 if the lanes cannot split it, see below);
// WRITE = true: its codes at out + offs[i].
//   read        the read's codes, re-coded for K3
//   pair path   traceToPoses + traceToSequence (Aligner.cpp:376-408, 425-428), gc_pair_path_string: entry k of the merged trace
//               emits the bases between the previous entry's cell and its own -- a function of entries k-1 and k alone as long as
//               the offsets inside a node never decrease (true for every trace of a DAG; checked, and a piece that violates it is
//               expanded by one thread with the sequential code) -- so: per-entry counts, block scan, scattered writes
//   node path   pathToTrace (Aligner.cpp:409-424), gc_node_path_string: per node a range of bases, same scan
#define GC_PIECE_THREADS 256
template <bool WRITE>
__global__ void __launch_bounds__(GC_PIECE_THREADS) gc_piece_kernel(GcPostGraph pg, const gcgpu_nw_piece* __restrict__ pieces, uint32_t n, const GcReadDesc* __restrict__ reads, const uint8_t* __restrict__ codes,
	const GcPair* const* __restrict__ setPairs, const uint64_t* const* __restrict__ setTraces, const uint32_t* __restrict__ pathNodes,
	uint64_t* __restrict__ lens, uint8_t* __restrict__ flags, const uint64_t* __restrict__ offs, uint8_t* __restrict__ out)
{
	typedef cub::BlockScan<uint32_t, GC_PIECE_THREADS> Scan;
	__shared__ typename Scan::TempStorage scanTmp;
	__shared__ uint32_t sViolation;
	const uint32_t i = blockIdx.x, tid = threadIdx.x;
	if (i >= n) { if (!WRITE && i == n && tid == 0) lens[i] = 0; return; }
	const gcgpu_nw_piece pc = pieces[i];
	if (pc.kind == GCGPU_PIECE_READ)
	{
		const GcReadDesc rd = reads[pc.index];
		if (!WRITE) { if (tid == 0) { lens[i] = (uint64_t)rd.len; flags[i] = 0; } return; }
		const uint8_t* src = codes + 2 * rd.charOffset;
		uint8_t* dst = out + offs[i];
		for (int32_t k = tid; k < rd.len; k += GC_PIECE_THREADS) dst[k] = gc_k3_code_of(src[k]);
		return;
	}
	if (tid == 0) sViolation = 0;
	__syncthreads();
	uint8_t* dst = WRITE ? out + offs[i] : nullptr;
	if (WRITE && flags[i])
	{
		// sequential form (never taken on a DAG)
		if (tid == 0)
		{
			if (pc.kind == GCGPU_PIECE_PAIR_PATH) gc_pair_path_string(pg, setTraces[pc.set], setPairs[pc.set][pc.index], dst);
			else gc_node_path_string(pg, pathNodes + pc.first_node, pc.num_nodes, pc.first_offset, pc.last_offset, dst);
		}
		return;
	}
	uint32_t base = 0;
	if (pc.kind == GCGPU_PIECE_PAIR_PATH)
	{
		const uint64_t* tr = setTraces[pc.set];
		const GcPair p = setPairs[pc.set][pc.index];
		const uint32_t m = gc_pair_size(p);
		for (uint32_t tile = 0; tile < m; tile += GC_PIECE_THREADS)
		{
			const uint32_t k = tile + tid;
			uint32_t cnt = 0, tailNode = 0, tailFrom = 0, tailCnt = 0, headNode = 0, headFrom = 0, headCnt = 0;
			if (k < m)
			{
				GcMergedEntry e = gc_pair_entry(pg, tr, p, k);
				if (k == 0) { headNode = e.node; headFrom = e.offset; headCnt = 1; }
				else
				{
					GcMergedEntry pe = gc_pair_entry(pg, tr, p, k - 1);
					if (e.node == pe.node)
					{
						if (e.offset < pe.offset) sViolation = 1;
						else { headNode = e.node; headFrom = pe.offset + 1; headCnt = e.offset - pe.offset; }
					}
					else
					{
						tailNode = pe.node; tailFrom = pe.offset + 1; tailCnt = pg.nodeLength[pe.node] - pe.offset - 1; // the rest of the node left behind
						headNode = e.node; headFrom = 0; headCnt = e.offset + 1;                                          // the new node up to the cell entered
					}
				}
				cnt = tailCnt + headCnt;
			}
			uint32_t off, total;
			Scan(scanTmp).ExclusiveSum(cnt, off, total);
			__syncthreads();
			if (WRITE)
			{
				uint8_t* o = dst + base + off;
				for (uint32_t c = 0; c < tailCnt; c++) o[c] = (uint8_t)gc_post_base(pg, tailNode, tailFrom + c);
				o += tailCnt;
				for (uint32_t c = 0; c < headCnt; c++) o[c] = (uint8_t)gc_post_base(pg, headNode, headFrom + c);
			}
			base += total;
		}
	}
	else
	{
		const uint32_t* path = pathNodes + pc.first_node;
		const uint32_t m = pc.num_nodes, first = path[0], lastNode = path[m - 1];
		for (uint32_t tile = 0; tile < m; tile += GC_PIECE_THREADS)
		{
			const uint32_t k = tile + tid;
			uint32_t node = 0, S = 0, cnt = 0;
			if (k < m)
			{
				node = path[k];
				uint32_t L = pg.nodeLength[node];
				if (node == first) S = pc.first_offset; else if (node == lastNode) L = pc.last_offset + 1; // compared by VALUE, as Aligner.cpp:413-416 does
				cnt = L > S ? L - S : 0;
			}
			uint32_t off, total;
			Scan(scanTmp).ExclusiveSum(cnt, off, total);
			__syncthreads();
			if (WRITE) { uint8_t* o = dst + base + off; for (uint32_t c = 0; c < cnt; c++) o[c] = (uint8_t)gc_post_base(pg, node, S + c); }
			base += total;
		}
	}
	if (!WRITE)
	{
		__syncthreads();
		if (tid == 0)
		{
			uint64_t len = base;
			if (sViolation) len = gc_pair_path_string(pg, setTraces[pc.set], setPairs[pc.set][pc.index], nullptr);
			lens[i] = len; flags[i] = (uint8_t)sViolation;
		}
	}
}

extern "C" int gcgpu_nw_compose(gcgpu_ctx* ctx, const gcgpu_nw_piece* pieces, uint32_t n, const uint32_t* path_nodes, uint64_t num_path_nodes, uint64_t* piece_offsets)
{
	GC_NEED_RESIDENT("gcgpu_nw_compose");
	GcResident* R = ctx->resident;
	if ((!pieces && n) || !piece_offsets || (!path_nodes && num_path_nodes)) return setError(GCGPU_ERR_ARG, "gcgpu_nw_compose: null argument");
	piece_offsets[0] = 0;
	ctx->lastKernelMs = 0;
	ctx->nwResident = ~0ULL;
	if (n == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	for (uint32_t i = 0; i < n; i++)
	{
		const gcgpu_nw_piece& pc = pieces[i];
		bool ok = false;
		if (pc.kind == GCGPU_PIECE_READ) ok = pc.index < R->numReads;
		else if (pc.kind == GCGPU_PIECE_PAIR_PATH) ok = pc.set < GCGPU_TRACE_SETS && pc.index < R->sets[pc.set].numPairs;
		else if (pc.kind == GCGPU_PIECE_NODE_PATH) ok = pc.num_nodes > 0 && pc.first_node + (uint64_t)pc.num_nodes <= num_path_nodes && pc.first_offset < 64 && pc.last_offset < 64;
		if (!ok) return setError(GCGPU_ERR_ARG, "gcgpu_nw_compose: piece " + std::to_string(i) + " out of range");
	}
	for (uint64_t i = 0; i < num_path_nodes; i++) if (path_nodes[i] >= ctx->numNodes) return setError(GCGPU_ERR_ARG, "gcgpu_nw_compose: path node out of range");
	struct SetPtrs { const GcPair* pairs[GCGPU_TRACE_SETS]; const uint64_t* traces[GCGPU_TRACE_SETS]; } sp;
	for (int s = 0; s < GCGPU_TRACE_SETS; s++) { sp.pairs[s] = (const GcPair*)R->sets[s].pairs.p; sp.traces[s] = (const uint64_t*)R->sets[s].traces.p; }
	size_t offPieces = 0, offPtrs = alignUp((size_t)n * sizeof(gcgpu_nw_piece), 128), offLens = alignUp(offPtrs + sizeof(SetPtrs), 128), offOffs = alignUp(offLens + ((size_t)n + 1) * 8, 128), offFlags = offOffs + ((size_t)n + 1) * 8, end = offFlags + n + 16;
	CUDA_TRY(R->pieces.ensure(end));
	CUDA_TRY(R->pathNodes.ensure(num_path_nodes * 4 + 16));
	uint8_t* P = (uint8_t*)R->pieces.p;
	const GcPair* const* dPairs = (const GcPair* const*)(P + offPtrs); const uint64_t* const* dTraces = (const uint64_t* const*)(P + offPtrs + sizeof(sp.pairs));
	uint64_t* dLens = (uint64_t*)(P + offLens); uint64_t* dOffs = (uint64_t*)(P + offOffs); uint8_t* dFlags = P + offFlags;
	CUDA_TRY(gcCopy(ctx, P + offPieces, pieces, (size_t)n * sizeof(gcgpu_nw_piece), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(gcCopy(ctx, P + offPtrs, &sp, sizeof(sp), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(gcCopy(ctx, R->pathNodes.p, path_nodes, num_path_nodes * 4, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
	gc_piece_kernel<false><<<n + 1, GC_PIECE_THREADS, 0, ctx->stream>>>(R->pg, (const gcgpu_nw_piece*)(P + offPieces), n, (const GcReadDesc*)R->reads.p, (const uint8_t*)ctx->seqBuf.p, dPairs, dTraces, (const uint32_t*)R->pathNodes.p,
		dLens, dFlags, nullptr, nullptr);
	ctx->launches++;
	int rc = scanU64(ctx, dLens, dOffs, n); if (rc != GCGPU_OK) return rc;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(gcCopy(ctx, piece_offsets, dOffs, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	uint64_t total = piece_offsets[n];
	CUDA_TRY(ctx->nwSeqBuf.ensure(total + 16));
	gc_piece_kernel<true><<<n, GC_PIECE_THREADS, 0, ctx->stream>>>(R->pg, (const gcgpu_nw_piece*)(P + offPieces), n, (const GcReadDesc*)R->reads.p, (const uint8_t*)ctx->seqBuf.p, dPairs, dTraces, (const uint32_t*)R->pathNodes.p,
		dLens, dFlags, dOffs, (uint8_t*)ctx->nwSeqBuf.p);
	ctx->launches++;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	float ms = 0;
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->lastKernelMs = ms;
	GC_TRACE_MS("k3 sequence pieces", n);
	ctx->nwResident = total;
	return GCGPU_OK;
}

// ------------------------------------------------------------------ edit runs of whole-read alignments
// One block per alignment (GraphAlignerVGAlignment::traceToAlignment, GraphAlignerVGAlignment.h:37-165, as gc_tokenize states it).
// Whether a trace step opens a new mapping and which edit type it is are functions of steps pos-1 and pos alone as long as a
// node switch never re-enters the current original node at or before the mapping's first offset (impossible on a DAG;
// checked, and an alignment that violates it is tokenized by one thread with the sequential code).  So: per-step flags,
// one block scan of (mappings opened, runs opened, position of the last run start), scattered token writes.
// WRITE = false: counts[i] = number of token words, flags[i] = sequential fallback; WRITE = true: tokens + meta.
#define GC_TOKEN_THREADS 256
struct GcTokScan { uint32_t b, r; int32_t s; };
struct GcTokScanOp { __device__ __forceinline__ GcTokScan operator()(const GcTokScan& x, const GcTokScan& y) const { GcTokScan o; o.b = x.b + y.b; o.r = x.r + y.r; o.s = x.s > y.s ? x.s : y.s; return o; } };
template <bool WRITE>
__global__ void __launch_bounds__(GC_TOKEN_THREADS) gc_tokens_kernel(GcPostGraph pg, const GcPair* __restrict__ pairs, const uint64_t* __restrict__ tr, const uint32_t* __restrict__ which, uint32_t n, const GcReadDesc* __restrict__ reads, const uint8_t* __restrict__ codes,
	const uint64_t* __restrict__ offs, uint32_t* __restrict__ tokens, uint64_t* __restrict__ counts, uint8_t* __restrict__ flags, gcgpu_aln_tokens* __restrict__ meta)
{
	typedef cub::BlockScan<GcTokScan, GC_TOKEN_THREADS> Scan;
	__shared__ typename Scan::TempStorage scanTmp;
	__shared__ uint8_t tileType[GC_TOKEN_THREADS];
	__shared__ uint32_t sViolation, sCount[4];
	__shared__ uint8_t sLastType;
	const uint32_t i = blockIdx.x, tid = threadIdx.x;
	if (i >= n) { if (!WRITE && i == n && tid == 0) counts[i] = 0; return; }
	const GcPair p = pairs[which[i]];
	GcPairTokenSrc src; src.pg = &pg; src.tr = tr; src.p = &p; src.codes = codes + 2 * reads[p.read].charOffset;
	const uint32_t m = gc_pair_size(p);
	uint32_t* out = WRITE ? tokens + offs[i] : nullptr;
	if (WRITE && flags[i])
	{
		if (tid == 0)
		{
			GcTokenCounts c = gc_tokenize(src, m, out);
			gcgpu_aln_tokens mt; mt.token_offset = offs[i]; mt.num_tokens = c.tokens; mt.matches = c.matches; mt.steps = c.matches + c.mismatches + c.insertions + c.deletions; mt.reserved = 0;
			meta[i] = mt;
		}
		return;
	}
	if (tid == 0) { sViolation = 0; sCount[0] = sCount[1] = sCount[2] = sCount[3] = 0; sLastType = 0; }
	__syncthreads();
	GcTokScan carry; carry.b = 0; carry.r = 0; carry.s = 0;
	for (uint32_t tile = 0; tile < m; tile += GC_TOKEN_THREADS)
	{
		const uint32_t pos = tile + tid;
		bool boundary = false;
		uint32_t t = 0;
		GcTokenStep e; e.node = 0; e.nodeOffset = 0; e.seqPos = 0; e.nodeSwitch = false; e.match = false;
		if (pos < m)
		{
			e = src(pos);
			if (pos == 0) { boundary = true; t = e.match ? GC_EDIT_MATCH : GC_EDIT_MISMATCH; }
			else
			{
				GcTokenStep prev = src(pos - 1);
				boundary = prev.nodeSwitch && e.node != prev.node;
				if (prev.nodeSwitch && e.node == prev.node && e.nodeOffset <= prev.nodeOffset) sViolation = 1; // would need the mapping's first offset
				if (!boundary && e.nodeOffset < prev.nodeOffset) sViolation = 1;
				if (prev.seqPos == e.seqPos) t = GC_EDIT_DELETION;
				else if (!boundary && prev.nodeOffset == e.nodeOffset) t = GC_EDIT_INSERTION;
				else t = e.match ? GC_EDIT_MATCH : GC_EDIT_MISMATCH;
			}
			atomicAdd(&sCount[t], 1u);
		}
		tileType[tid] = (uint8_t)t;
		__syncthreads();
		const uint32_t prevType = tid > 0 ? tileType[tid - 1] : sLastType;
		const bool runStart = pos < m && (pos == 0 || boundary || t != prevType);
		GcTokScan mine; mine.b = (pos < m && boundary) ? 1u : 0u; mine.r = runStart ? 1u : 0u; mine.s = runStart ? (int32_t)pos : 0;
		GcTokScan ex, agg, ident; ident.b = 0; ident.r = 0; ident.s = 0;
		Scan(scanTmp).ExclusiveScan(mine, ex, ident, GcTokScanOp(), agg);
		ex = GcTokScanOp()(carry, ex);
		if (WRITE && runStart)
		{
			uint32_t w = 3 * ex.b + ex.r; // words before this step's tokens, plus one for the end word of the run that ends here
			if (pos > 0) out[w - 1] = (prevType << 30) | (pos - (uint32_t)ex.s);
			if (boundary) { out[w] = 0; out[w + 1] = (uint32_t)e.node; out[w + 2] = e.nodeOffset; }
		}
		__syncthreads();
		carry = GcTokScanOp()(carry, agg);
		if (tid == GC_TOKEN_THREADS - 1) sLastType = (uint8_t)t;
		__syncthreads();
	}
	if (tid == 0)
	{
		uint32_t total = m ? 3 * carry.b + carry.r : 0;
		if (!WRITE)
		{
			uint8_t fl = (uint8_t)sViolation;
			if (fl) { GcTokenCounts c = gc_tokenize(src, m, (uint32_t*)nullptr); total = c.tokens; }
			counts[i] = total; flags[i] = fl;
		}
		else
		{
			if (m) out[total - 1] = ((uint32_t)tileType[(m - 1) % GC_TOKEN_THREADS] << 30) | (m - (uint32_t)carry.s);
			gcgpu_aln_tokens mt; mt.token_offset = offs[i]; mt.num_tokens = total; mt.matches = sCount[GC_EDIT_MATCH]; mt.steps = sCount[0] + sCount[1] + sCount[2] + sCount[3]; mt.reserved = 0;
			meta[i] = mt;
		}
	}
}

extern "C" int gcgpu_encode_alignments(gcgpu_ctx* ctx, int set, const uint32_t* pairs, uint32_t n, gcgpu_aln_tokens* out, uint64_t* tokens_used)
{
	GC_NEED_RESIDENT("gcgpu_encode_alignments");
	GcResident* R = ctx->resident;
	if (set < 0 || set >= GCGPU_TRACE_SETS || (!pairs && n) || (!out && n) || !tokens_used) return setError(GCGPU_ERR_ARG, "gcgpu_encode_alignments: bad argument");
	*tokens_used = 0; R->tokensUsed = 0;
	ctx->lastKernelMs = 0;
	if (n == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	GcTraceSet& S = R->sets[set];
	for (uint32_t i = 0; i < n; i++) if (pairs[i] >= S.numPairs) return setError(GCGPU_ERR_ARG, "gcgpu_encode_alignments: pair " + std::to_string(i) + " out of range");
	size_t offWhich = 0, offCnt = alignUp((size_t)n * 4, 128), offOffs = alignUp(offCnt + ((size_t)n + 1) * 8, 128), offMeta = alignUp(offOffs + ((size_t)n + 1) * 8, 128), offFlags = offMeta + (size_t)n * sizeof(gcgpu_aln_tokens), end = offFlags + n + 16;
	CUDA_TRY(R->tokenMeta.ensure(end));
	uint8_t* T = (uint8_t*)R->tokenMeta.p;
	uint64_t* dCnt = (uint64_t*)(T + offCnt); uint64_t* dOffs = (uint64_t*)(T + offOffs); uint8_t* dFlags = T + offFlags;
	CUDA_TRY(gcCopy(ctx, T + offWhich, pairs, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
	gc_tokens_kernel<false><<<n + 1, GC_TOKEN_THREADS, 0, ctx->stream>>>(R->pg, (const GcPair*)S.pairs.p, (const uint64_t*)S.traces.p, (const uint32_t*)(T + offWhich), n, (const GcReadDesc*)R->reads.p, (const uint8_t*)ctx->seqBuf.p,
		nullptr, nullptr, dCnt, dFlags, nullptr);
	ctx->launches++;
	int rc = scanU64(ctx, dCnt, dOffs, n); if (rc != GCGPU_OK) return rc;
	CUDA_TRY(cudaGetLastError());
	uint64_t total = 0;
	CUDA_TRY(gcCopy(ctx, &total, dOffs + n, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	CUDA_TRY(R->tokens.ensure(total * 4 + 16));
	gc_tokens_kernel<true><<<n, GC_TOKEN_THREADS, 0, ctx->stream>>>(R->pg, (const GcPair*)S.pairs.p, (const uint64_t*)S.traces.p, (const uint32_t*)(T + offWhich), n, (const GcReadDesc*)R->reads.p, (const uint8_t*)ctx->seqBuf.p,
		dOffs, (uint32_t*)R->tokens.p, nullptr, dFlags, (gcgpu_aln_tokens*)(T + offMeta));
	ctx->launches++;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	CUDA_TRY(gcCopy(ctx, out, T + offMeta, (size_t)n * sizeof(gcgpu_aln_tokens), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	float ms = 0;
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->lastKernelMs = ms;
	GC_TRACE_MS("s7 edit runs", n);
	R->tokensUsed = total;
	*tokens_used = total;
	return GCGPU_OK;
}

extern "C" int gcgpu_fetch_tokens(gcgpu_ctx* ctx, uint32_t* tokens, uint64_t first, uint64_t count)
{
	GC_NEED_RESIDENT("gcgpu_fetch_tokens");
	GcResident* R = ctx->resident;
	if (!tokens && count) return setError(GCGPU_ERR_ARG, "gcgpu_fetch_tokens: null argument");
	if (first + count > R->tokensUsed) return setError(GCGPU_ERR_ARG, "gcgpu_fetch_tokens: range beyond the tokens of the last gcgpu_encode_alignments call");
	if (count == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	CUDA_TRY(gcCopy(ctx, tokens, (const uint32_t*)R->tokens.p + first, count * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	return GCGPU_OK;
}

// ------------------------------------------------------------------ GAM records on the device (gc_gam.cuh)
extern "C" int gcgpu_set_node_names(gcgpu_ctx* ctx, const uint32_t* name_offsets, const char* names)
{
	GC_NEED_RESIDENT("gcgpu_set_node_names");
	GcResident* R = ctx->resident;
	if (!name_offsets || (!names && name_offsets[R->hostOrigIds.size()] > 0)) return setError(GCGPU_ERR_ARG, "gcgpu_set_node_names: null argument");
	CUDA_TRY(cudaSetDevice(ctx->device));
	size_t numOrig = R->hostOrigIds.size();
	int32_t maxId = -1;
	for (int32_t id : R->hostOrigIds) if (id > maxId) maxId = id;
	std::vector<int32_t> indexOfId((size_t)maxId + 2, -1);
	for (size_t o = 0; o < numOrig; o++) if (R->hostOrigIds[o] >= 0) indexOfId[R->hostOrigIds[o]] = (int32_t)o;
	GcDeflateTables tables; gcBuildGamTables(tables);
	cudaFree(R->d_origIndexOfId); cudaFree(R->d_nameOff); cudaFree(R->d_nameChars); cudaFree(R->d_gamTables);
	R->d_origIndexOfId = nullptr; R->d_nameOff = nullptr; R->d_nameChars = nullptr; R->d_gamTables = nullptr;
	cudaError_t err = cudaSuccess;
	auto chk = [&err](cudaError_t x) { if (err == cudaSuccess) err = x; };
	chk(uploadArray(indexOfId.data(), indexOfId.size(), &R->d_origIndexOfId));
	chk(uploadArray(name_offsets, numOrig + 1, &R->d_nameOff));
	chk(uploadArray((const uint8_t*)names, (size_t)name_offsets[numOrig] + 1, &R->d_nameChars));
	chk(uploadArray(&tables, 1, &R->d_gamTables));
	if (err != cudaSuccess) return setError(err == cudaErrorMemoryAllocation ? GCGPU_ERR_NOMEM : GCGPU_ERR_CUDA, std::string("gcgpu_set_node_names: ") + cudaGetErrorString(err));
	R->haveNames = true;
	return GCGPU_OK;
}

struct GcGamSlot { uint64_t rawOff, gzOff, wsOff; uint32_t rawLen, gzCap; };
__global__ void gc_gam_alns_kernel(const gcgpu_gam_aln* __restrict__ in, const gcgpu_aln_tokens* __restrict__ meta, uint32_t n, GcGamAln* __restrict__ out)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	GcGamAln a; a.tokenOff = meta[i].token_offset; a.numTokens = meta[i].num_tokens; a.start = in[i].start; a.end = in[i].end; a.traceScore = in[i].trace_score; a.matches = meta[i].matches; a.steps = meta[i].steps;
	out[i] = a;
}
// bytes every record needs: the raw record, the gzip member (worst case) and the encoder's workspace
// record sizes: one record per warp, its alignments on the lanes (a thread per record serialises 32 different token walks in a
// warp: 4.9 ms per 1678 records, r03t)
__global__ void __launch_bounds__(32) gc_gam_size_kernel(GcNameTable nt, const gcgpu_gam_read* __restrict__ reads, uint32_t n, const GcGamAln* __restrict__ alns, const uint32_t* __restrict__ tokens, uint32_t* __restrict__ rawLen, uint64_t* __restrict__ slotBytes)
{
	const uint32_t i = blockIdx.x, lane = threadIdx.x;
	if (i > n) return;
	if (i == n) { if (lane == 0) slotBytes[i] = 0; return; }
	gcgpu_gam_read rd = reads[i];
	uint32_t len = 0;
	for (uint32_t k = lane; k < rd.num_alns; k += 32)
	{
		uint32_t ps;
		uint32_t m = gc_gam_message_size(nt, alns[rd.first_aln + k], tokens + alns[rd.first_aln + k].tokenOff, rd.name_len, ps);
		len += gc_gam_vsize(m) + m;
	}
	for (int d = 16; d > 0; d >>= 1) len += __shfl_xor_sync(0xFFFFFFFFu, len, d);
	len += gc_gam_vsize(rd.num_alns);
	if (lane == 0)
	{
		rawLen[i] = len;
		slotBytes[i] = (((uint64_t)len + 16 + 127) / 128 + ((uint64_t)len * 2 + 1024 + 127) / 128 + (gc_deflate_ws_bytes(len) + 127) / 128) * 128;
	}
}
// One record = one warp (a block of 32 threads).  Lane 0 writes the record -- a chain of data-dependent varint fields -- and the 32
// lanes then make its gzip member together (gc_gam.cuh: chunk-parallel LZ77 parse, one Huffman code, every lane writes its chunk's
// codes at its bit offset).  r03j, lane 0 doing everything and the other 63 threads of the block idle: ~15 ms per 1678 records and
// a block's registers held for all of it.
__global__ void __launch_bounds__(32) gc_gam_kernel(GcNameTable nt, const GcDeflateTables* __restrict__ tables, const gcgpu_gam_read* __restrict__ reads, uint32_t n, const GcGamAln* __restrict__ alns, const uint32_t* __restrict__ tokens,
	const GcReadDesc* __restrict__ readDescs, const uint8_t* __restrict__ chars, const uint8_t* __restrict__ names, const uint32_t* __restrict__ rawLen, const uint64_t* __restrict__ slotOff, uint8_t* arena, uint64_t* __restrict__ memberLen)
{
	__shared__ GcGzShared sh;
	const uint32_t lane = threadIdx.x, i = blockIdx.x;
	if (i > n) return;
	if (i == n) { if (lane == 0) memberLen[i] = 0; return; }
	gcgpu_gam_read rd = reads[i];
	const uint32_t len = rawLen[i];
	uint8_t* raw = arena + slotOff[i];
	uint8_t* gz = raw + ((uint64_t)len + 16 + 127) / 128 * 128;
	const uint32_t gzCap = len * 2 + 1024;
	uint32_t* lzTokens = (uint32_t*)(gz + ((uint64_t)gzCap + 127) / 128 * 128);
	uint32_t written = 0;
	if (lane == 0) written = gc_gam_write_record(nt, chars + readDescs[rd.read].charOffset, names + rd.name_offset, rd.name_len, alns + rd.first_aln, rd.num_alns, tokens, raw);
	written = __shfl_sync(0xFFFFFFFFu, written, 0);
	if (written != len || !gc_gz_fits(len, gzCap)) { if (lane == 0) memberLen[i] = 0; return; } // the host encodes this record
	gc_gz_clear(sh, lane);
	__syncwarp();
	gc_gz_tokenize(*tables, raw, len, lane, lzTokens, sh);
	__syncwarp();
	if (lane == 0) gc_gz_header(sh);
	__syncwarp();
	if (!sh.ok) { if (lane == 0) memberLen[i] = 0; return; }
	gc_gz_bits(*tables, len, lane, lzTokens, sh);
	__syncwarp();
	{
		const uint32_t bits = sh.laneBits[lane];
		uint32_t incl = bits;
		for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += v; }
		sh.laneStart[lane] = incl - bits;
	}
	__syncwarp();
	gc_gz_prepare(lane, sh, (uint32_t*)gz);
	__syncwarp();
	gc_gz_emit(*tables, len, lane, lzTokens, sh, (uint32_t*)gz);
	__syncwarp();
	if (lane == 0) memberLen[i] = gc_gz_trailer(len, sh, gz);
}
__global__ void gc_gam_gather_kernel(const uint64_t* __restrict__ slotOff, const uint32_t* __restrict__ rawLen, const uint64_t* __restrict__ memberLen, const uint64_t* __restrict__ memberOff, uint32_t n, const uint8_t* __restrict__ arena, uint8_t* __restrict__ out)
{
	uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (w >= n) return;
	const uint8_t* gz = arena + slotOff[w] + ((uint64_t)rawLen[w] + 16 + 127) / 128 * 128;
	uint8_t* dst = out + memberOff[w];
	for (uint64_t k = lane; k < memberLen[w]; k += 32) dst[k] = gz[k];
}

extern "C" int gcgpu_encode_gam(gcgpu_ctx* ctx, int set, const gcgpu_gam_read* reads, uint32_t n, const gcgpu_gam_aln* alns, uint32_t num_alns,
	const char* names, uint64_t name_bytes, uint64_t* member_offsets, uint64_t* bytes_used)
{
	GC_NEED_RESIDENT("gcgpu_encode_gam");
	GcResident* R = ctx->resident;
	if (!R->haveNames) return setError(GCGPU_ERR_ARG, "gcgpu_encode_gam: gcgpu_set_node_names was not called");
	if (set < 0 || set >= GCGPU_TRACE_SETS || (!reads && n) || (!alns && num_alns) || !member_offsets || !bytes_used || (!names && name_bytes)) return setError(GCGPU_ERR_ARG, "gcgpu_encode_gam: bad argument");
	*bytes_used = 0; R->gamBytes = 0; member_offsets[0] = 0;
	ctx->lastKernelMs = 0;
	if (n == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	GcTraceSet& S = R->sets[set];
	uint32_t nextAln = 0;
	for (uint32_t i = 0; i < n; i++)
	{
		const gcgpu_gam_read& rd = reads[i];
		if (rd.read >= R->numReads || rd.first_aln != nextAln || rd.num_alns == 0 || (uint64_t)rd.first_aln + rd.num_alns > num_alns || rd.name_offset + rd.name_len > name_bytes) return setError(GCGPU_ERR_ARG, "gcgpu_encode_gam: read " + std::to_string(i) + " out of range");
		nextAln += rd.num_alns;
	}
	if (nextAln != num_alns) return setError(GCGPU_ERR_ARG, "gcgpu_encode_gam: alignments outside the reads");
	std::vector<uint32_t> pairs(num_alns);
	for (uint32_t k = 0; k < num_alns; k++)
	{
		if (alns[k].pair >= S.numPairs || alns[k].start < 0 || alns[k].end <= alns[k].start) return setError(GCGPU_ERR_ARG, "gcgpu_encode_gam: alignment " + std::to_string(k) + " out of range");
		pairs[k] = alns[k].pair;
	}
	for (uint32_t i = 0; i < n; i++) for (uint32_t k = 0; k < reads[i].num_alns; k++) if (alns[reads[i].first_aln + k].end > R->hostReads[reads[i].read].len) return setError(GCGPU_ERR_ARG, "gcgpu_encode_gam: alignment beyond its read");
	// ---- inputs
	size_t oReads = 0, oAlns = alignUp((size_t)n * sizeof(gcgpu_gam_read), 128), oPairs = alignUp(oAlns + (size_t)num_alns * sizeof(gcgpu_gam_aln), 128), oNames = alignUp(oPairs + (size_t)num_alns * 4, 128), inEnd = oNames + name_bytes + 16;
	CUDA_TRY(R->gamIn.ensure(inEnd));
	uint8_t* I = (uint8_t*)R->gamIn.p;
	CUDA_TRY(gcCopy(ctx, I + oReads, reads, (size_t)n * sizeof(gcgpu_gam_read), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(gcCopy(ctx, I + oAlns, alns, (size_t)num_alns * sizeof(gcgpu_gam_aln), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(gcCopy(ctx, I + oPairs, pairs.data(), (size_t)num_alns * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (name_bytes) CUDA_TRY(gcCopy(ctx, I + oNames, names, name_bytes, cudaMemcpyHostToDevice, ctx->stream));
	// ---- work arrays: token counts / offsets / meta | GcGamAln | rawLen | slot bytes / offsets | member lengths / offsets
	size_t oCnt = 0, oOffs = alignUp(oCnt + ((size_t)num_alns + 1) * 8, 128), oMeta = alignUp(oOffs + ((size_t)num_alns + 1) * 8, 128), oFlags = alignUp(oMeta + (size_t)num_alns * sizeof(gcgpu_aln_tokens), 128);
	size_t oGAln = alignUp(oFlags + num_alns + 16, 128), oRaw = alignUp(oGAln + (size_t)num_alns * sizeof(GcGamAln), 128), oSlotB = alignUp(oRaw + (size_t)n * 4, 128), oSlotO = alignUp(oSlotB + ((size_t)n + 1) * 8, 128);
	size_t oMemL = alignUp(oSlotO + ((size_t)n + 1) * 8, 128), oMemO = alignUp(oMemL + ((size_t)n + 1) * 8, 128), workEnd = oMemO + ((size_t)n + 1) * 8;
	CUDA_TRY(R->gamWork.ensure(workEnd));
	uint8_t* W = (uint8_t*)R->gamWork.p;
	uint64_t* dCnt = (uint64_t*)(W + oCnt); uint64_t* dOffs = (uint64_t*)(W + oOffs); gcgpu_aln_tokens* dMeta = (gcgpu_aln_tokens*)(W + oMeta); uint8_t* dFlags = W + oFlags;
	GcGamAln* dGAln = (GcGamAln*)(W + oGAln); uint32_t* dRawLen = (uint32_t*)(W + oRaw); uint64_t* dSlotB = (uint64_t*)(W + oSlotB); uint64_t* dSlotO = (uint64_t*)(W + oSlotO);
	uint64_t* dMemL = (uint64_t*)(W + oMemL); uint64_t* dMemO = (uint64_t*)(W + oMemO);
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
	// ---- edit-run tokens of the alignments (as gcgpu_encode_alignments)
	gc_tokens_kernel<false><<<num_alns + 1, GC_TOKEN_THREADS, 0, ctx->stream>>>(R->pg, (const GcPair*)S.pairs.p, (const uint64_t*)S.traces.p, (const uint32_t*)(I + oPairs), num_alns, (const GcReadDesc*)R->reads.p, (const uint8_t*)ctx->seqBuf.p,
		nullptr, nullptr, dCnt, dFlags, nullptr);
	ctx->launches++;
	int rc = scanU64(ctx, dCnt, dOffs, num_alns); if (rc != GCGPU_OK) return rc;
	CUDA_TRY(cudaGetLastError());
	uint64_t totalTokens = 0;
	CUDA_TRY(gcCopy(ctx, &totalTokens, dOffs + num_alns, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	CUDA_TRY(R->tokens.ensure(totalTokens * 4 + 16));
	gc_tokens_kernel<true><<<num_alns, GC_TOKEN_THREADS, 0, ctx->stream>>>(R->pg, (const GcPair*)S.pairs.p, (const uint64_t*)S.traces.p, (const uint32_t*)(I + oPairs), num_alns, (const GcReadDesc*)R->reads.p, (const uint8_t*)ctx->seqBuf.p,
		dOffs, (uint32_t*)R->tokens.p, nullptr, dFlags, dMeta);
	gc_gam_alns_kernel<<<(num_alns + 127) / 128, 128, 0, ctx->stream>>>((const gcgpu_gam_aln*)(I + oAlns), dMeta, num_alns, dGAln);
	GcNameTable nt; nt.origIndexOfId = R->d_origIndexOfId; nt.nameOff = R->d_nameOff; nt.nameChars = R->d_nameChars;
	gc_gam_size_kernel<<<n + 1, 32, 0, ctx->stream>>>(nt, (const gcgpu_gam_read*)(I + oReads), n, dGAln, (const uint32_t*)R->tokens.p, dRawLen, dSlotB);
	ctx->launches += 3;
	rc = scanU64(ctx, dSlotB, dSlotO, n); if (rc != GCGPU_OK) return rc;
	CUDA_TRY(cudaGetLastError());
	uint64_t arenaBytes = 0;
	CUDA_TRY(gcCopy(ctx, &arenaBytes, dSlotO + n, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	CUDA_TRY(R->gamArena.ensure(arenaBytes + 256));
	gc_gam_kernel<<<n + 1, 32, 0, ctx->stream>>>(nt, R->d_gamTables, (const gcgpu_gam_read*)(I + oReads), n, dGAln, (const uint32_t*)R->tokens.p, (const GcReadDesc*)R->reads.p, (const uint8_t*)R->chars.p,
		I + oNames, dRawLen, dSlotO, (uint8_t*)R->gamArena.p, dMemL);
	ctx->launches++;
	rc = scanU64(ctx, dMemL, dMemO, n); if (rc != GCGPU_OK) return rc;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(gcCopy(ctx, member_offsets, dMemO, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	uint64_t total = member_offsets[n];
	if (total > arenaBytes)
	{
		// a member cannot be larger than its slot: something wrote a wrong length
		std::vector<uint32_t> hRaw(n); std::vector<uint64_t> hLen((size_t)n + 1);
		CUDA_TRY(gcCopy(ctx, hRaw.data(), dRawLen, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(gcCopy(ctx, hLen.data(), dMemL, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(gcSyncStream(ctx));
		std::string msg = "gcgpu_encode_gam: member lengths exceed the record slots (" + std::to_string(total) + " > " + std::to_string(arenaBytes) + " bytes, " + std::to_string(n) + " records):";
		int shown = 0;
		for (uint32_t i = 0; i < n && shown < 4; i++) if (hLen[i] > (uint64_t)hRaw[i] * 2 + 1024) { msg += " record " + std::to_string(i) + " raw " + std::to_string(hRaw[i]) + " member " + std::to_string(hLen[i]) + " alns " + std::to_string(reads[i].num_alns) + ";"; shown++; }
		return setError(GCGPU_ERR_INTERNAL, msg);
	}
	CUDA_TRY(R->gamOut.ensure(total + 16));
	gc_gam_gather_kernel<<<(n + 3) / 4, 128, 0, ctx->stream>>>(dSlotO, dRawLen, dMemL, dMemO, n, (const uint8_t*)R->gamArena.p, (uint8_t*)R->gamOut.p);
	ctx->launches++;
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	float ms = 0;
	CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	ctx->lastKernelMs = ms;
	GC_TRACE_MS("s7 gam records", n);
	R->gamBytes = total;
	*bytes_used = total;
	return GCGPU_OK;
}

extern "C" int gcgpu_fetch_gam(gcgpu_ctx* ctx, uint8_t* out, uint64_t first, uint64_t count)
{
	GC_NEED_RESIDENT("gcgpu_fetch_gam");
	GcResident* R = ctx->resident;
	if (!out && count) return setError(GCGPU_ERR_ARG, "gcgpu_fetch_gam: null argument");
	if (first + count > R->gamBytes) return setError(GCGPU_ERR_ARG, "gcgpu_fetch_gam: range beyond the members of the last gcgpu_encode_gam call");
	if (count == 0) return GCGPU_OK;
	CUDA_TRY(cudaSetDevice(ctx->device));
	CUDA_TRY(gcCopy(ctx, out, (const uint8_t*)R->gamOut.p + first, count, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(gcSyncStream(ctx));
	return GCGPU_OK;
}

extern "C" void gcgpu_transfer_bytes(gcgpu_ctx* ctx, uint64_t* h2d, uint64_t* d2h)
{
	if (h2d) *h2d = ctx ? ctx->h2dBytes : 0;
	if (d2h) *d2h = ctx ? ctx->d2hBytes : 0;
}
