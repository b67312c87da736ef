// Work between the DP kernels, on the packed K1 traces where they lie in HBM: the traces of a
// batch never cross PCIe, the host only sees alignment extents, anchor records and edit runs.
//
//   pair            one seed extension = the backward + forward K1 traces of getTwoDirectionalTrace
//                   (src/GraphAligner.h:480-525, 567-626), merged lazily (gc_pair_entry)
//   has_cell        GraphAligner::exactAlignmentPart            (src/GraphAligner.h:407-461)
//   fragment filter the seed loop of AlignOneWay for one 35-bp fragment (src/GraphAligner.h:114-203
//                   with sloppyOptimizations == false, as Aligner.cpp:690 calls it)
//   anchor path     Aligner.cpp:706-729
//   path string     traceToPoses + traceToSequence               (src/Aligner.cpp:376-408, 425-428)
//   node path bases pathToTrace + the pathseq loop               (src/Aligner.cpp:409-424, 832-836)
//   edit runs       GraphAlignerVGAlignment::traceToAlignment   (src/GraphAlignerVGAlignment.h:37-165)
//
// All functions are GC_HD: the kernels in gcgpu_resident.inl run them on the device, the C-ABI test
// double (tests/hostsim) runs the same code on the CPU-only build box.
#pragma once
#include "gc_common.cuh"

// trace entry fields (include/gcgpu.h GCGPU_TRACE_*)
GC_HD uint32_t gc_te_node(uint64_t t) { return (uint32_t)(t & 0xFFFFFFFFu); }
GC_HD uint32_t gc_te_offset(uint64_t t) { return (uint32_t)((t >> 32) & 63); }
GC_HD bool gc_te_switch(uint64_t t) { return ((t >> 38) & 1) != 0; }
GC_HD int32_t gc_te_seqpos(uint64_t t) { return (int32_t)((t >> 39) & 0x1FFFFFF) - 1; }

// per split node: where it sits in its original (bigraph-doubled) node, and where the reverse-complement
// strand of that original node lives (AlignmentGraph::GetReversePosition / GetUnitigNode, AlignmentGraph.cpp:832-868)
struct GcPostGraph
{
	const int32_t* nodeIDs;      // [N] digraph node id (2 * id + strand)
	const uint32_t* nodeOffset;  // [N] offset of the split node inside its original node
	const uint8_t* nodeLength;   // [N]
	const uint64_t* nodeSeq;     // [2N]
	const uint32_t* revFirst;    // [N] index into origNodes of the first split node of the reverse-strand original node
	const uint32_t* revCount;    // [N] number of split nodes of it
	const uint32_t* revLast;     // [N] originalNodeSize - 1 - nodeOffset[n]: reverse-strand offset of the split node's first base
	const uint32_t* origNodes;   // split nodes of every original node, in offset order
};

GC_HD int gc_post_base(const GcPostGraph& pg, uint32_t node, uint32_t pos)
{
	return (int)((pg.nodeSeq[2 * (uint64_t)node + (pos >> 5)] >> ((pos & 31) * 2)) & 3);
}

// GetUnitigNode(GetReversePosition(cell)): the same base seen from the other strand
GC_HD void gc_reverse_cell(const GcPostGraph& pg, uint32_t node, uint32_t off, uint32_t& rnode, uint32_t& roff)
{
	uint32_t ro = pg.revLast[node] - off;
	const uint32_t* nodes = pg.origNodes + pg.revFirst[node];
	uint32_t n = pg.revCount[node];
	uint32_t index = ro >> 6;
	if (index >= n) index = n - 1;
	while (index < n - 1 && pg.nodeOffset[nodes[index]] + pg.nodeLength[nodes[index]] <= ro) index++;
	while (index > 0 && pg.nodeOffset[nodes[index]] > ro) index--;
	rnode = nodes[index];
	roff = ro - pg.nodeOffset[rnode];
}

// consecutive trace entries mostly stay inside one split node: remember the reverse-strand split node the last entry mapped to
struct GcRevCache { uint32_t rawNode = 0xFFFFFFFFu; uint32_t revLast = 0; uint32_t rnode = 0; uint32_t lo = 1, hi = 0; };
GC_HD void gc_reverse_cell_cached(const GcPostGraph& pg, uint32_t node, uint32_t off, uint32_t& rnode, uint32_t& roff, GcRevCache& c)
{
	if (node == c.rawNode)
	{
		uint32_t ro = c.revLast - off;
		if (ro >= c.lo && ro < c.hi) { rnode = c.rnode; roff = ro - c.lo; return; }
	}
	gc_reverse_cell(pg, node, off, rnode, roff);
	c.rawNode = node; c.revLast = pg.revLast[node]; c.rnode = rnode; c.lo = pg.nodeOffset[rnode]; c.hi = c.lo + pg.nodeLength[rnode];
}

// one seed of a read (GraphAlignerWrapper.h:11-37 SeedHit, the fields the extension and the skip rules read)
struct GcSeedCell
{
	int32_t seqPos;    // SeedHit::seqPos: k-mer END position in the read
	uint32_t node;     // SeedHit::alignmentGraphNodeId (split node)
	uint32_t read;     // index of the read in the batch
	uint8_t offset;    // SeedHit::alignmentGraphNodeOffset
	uint8_t flags;     // bit 0: seedClusterSize >= seedClusterMinSize
	uint16_t reserved;
};
struct GcReadDesc
{
	uint64_t charOffset;  // characters of the read in the batch buffer; codes: forward at 2 * charOffset, reverse complement at 2 * charOffset + len
	int32_t len;
	uint32_t firstCell;
	uint32_t numCells;
	uint32_t reserved;
};
struct GcSeedExt { uint32_t cell; int32_t fragStart; }; // fragStart < 0: the whole read is the aligned sequence

#define GC_PAIR_BWD 1u
#define GC_PAIR_FWD 2u
#define GC_PAIR_INTERNAL 4u

// one seed extension with its traces inside a trace set
struct GcPair
{
	uint64_t bwdOff, fwdOff;   // first entry of each K1 trace in the set's dense trace array
	uint32_t bwdLen, fwdLen;   // 0 = that direction is absent or failed
	int32_t seedPos;           // seed position in the coordinates of the aligned sequence (read or fragment)
	int32_t start, end;        // AlignmentItem::alignmentStart / alignmentEnd (same coordinates)
	int32_t score;             // OnewayTrace::score summed over the directions that succeeded
	uint32_t flags;            // GC_PAIR_*
	uint32_t cell;
	int32_t fragStart;
	uint32_t read;
};

// the two K1 work items of one seed (getTwoDirectionalTrace, GraphAligner.h:480-525): lengths only
GC_HD void gc_ext_lengths(int32_t seqPosGlobal, int32_t fragStart, int32_t readLen, int32_t fragLen, int32_t& bwdLen, int32_t& fwdLen, int32_t& localPos)
{
	int32_t seqLen = fragStart < 0 ? readLen : fragLen;
	localPos = seqPosGlobal - (fragStart < 0 ? 0 : fragStart);
	bwdLen = localPos > 0 ? localPos : -1;
	fwdLen = localPos < seqLen - 1 ? seqLen - localPos - 1 : -1;
}

GC_HD uint32_t gc_pair_nb(const GcPair& p) { return p.bwdLen ? (p.fwdLen ? p.bwdLen - 1 : p.bwdLen) : 0; }
GC_HD uint32_t gc_pair_size(const GcPair& p) { return gc_pair_nb(p) + p.fwdLen; }

// entry k of the merged trace (backward part in kernel order without its last entry -- the duplicated seed cell,
// GraphAligner.h:599 -- then the forward part reversed), in forward-strand split-node coordinates
struct GcMergedEntry { uint32_t node, offset; int32_t seqPos; bool nodeSwitch; bool fromBwd; uint32_t rawNode; uint32_t rawOffset; };
GC_HD GcMergedEntry gc_pair_entry(const GcPostGraph& pg, const uint64_t* tr, const GcPair& p, uint32_t k, GcRevCache* cache = nullptr)
{
	GcMergedEntry e;
	uint32_t nb = gc_pair_nb(p);
	if (k < nb)
	{
		uint64_t t = tr[p.bwdOff + k];
		e.rawNode = gc_te_node(t); e.rawOffset = gc_te_offset(t);
		if (cache) gc_reverse_cell_cached(pg, e.rawNode, e.rawOffset, e.node, e.offset, *cache);
		else gc_reverse_cell(pg, e.rawNode, e.rawOffset, e.node, e.offset);
		e.seqPos = (p.seedPos - 1) - gc_te_seqpos(t);
		e.nodeSwitch = (k + 1 < p.bwdLen) ? gc_te_switch(tr[p.bwdOff + k + 1]) : false; // fixReverseTraceSeqPosAndOrder, GraphAligner.h:543-565
		e.fromBwd = true;
	}
	else
	{
		uint64_t t = tr[p.fwdOff + (p.fwdLen - 1 - (k - nb))];
		e.rawNode = e.node = gc_te_node(t); e.rawOffset = e.offset = gc_te_offset(t);
		e.seqPos = gc_te_seqpos(t) + p.seedPos + 1;
		e.nodeSwitch = gc_te_switch(t);
		e.fromBwd = false;
	}
	return e;
}

// getAlignmentFromSeed (GraphAligner.h:567-626) from the two K1 results: extent and score
GC_HD void gc_pair_finish(const uint64_t* tr, GcPair& p, bool haveB, int32_t scoreB, bool internalB, bool haveF, int32_t scoreF, bool internalF)
{
	p.flags = (haveB ? GC_PAIR_BWD : 0u) | (haveF ? GC_PAIR_FWD : 0u) | ((internalB || internalF) ? GC_PAIR_INTERNAL : 0u);
	if (!haveB) p.bwdLen = 0;
	if (!haveF) p.fwdLen = 0;
	p.score = (haveB ? scoreB : 0) + (haveF ? scoreF : 0);
	p.start = p.end = 0;
	if (!haveB && !haveF) return;
	uint32_t nb = gc_pair_nb(p);
	p.start = nb ? (p.seedPos - 1) - gc_te_seqpos(tr[p.bwdOff]) : gc_te_seqpos(tr[p.fwdOff + p.fwdLen - 1]) + p.seedPos + 1;
	p.end = (p.fwdLen ? gc_te_seqpos(tr[p.fwdOff]) + p.seedPos + 1 : (p.seedPos - 1) - gc_te_seqpos(tr[p.bwdOff + nb - 1])) + 1;
}

// first index of a kernel-order trace (seqPos non-increasing) whose seqPos is <= want
GC_HD uint32_t gc_trace_lower(const uint64_t* t, uint32_t n, int32_t want)
{
	uint32_t lo = 0, hi = n;
	while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (gc_te_seqpos(t[mid]) > want) lo = mid + 1; else hi = mid; }
	return lo;
}

// exactAlignmentPart (GraphAligner.h:407-461) without its assertion: is the cell (forward-strand split node `node`, `offset`,
// sequence position sp in the pair's coordinates) on the merged trace of the pair?
GC_HD bool gc_pair_has_cell(const GcPostGraph& pg, const uint64_t* tr, const GcPair& p, uint32_t node, uint32_t offset, int32_t sp)
{
	if (!(p.flags & (GC_PAIR_BWD | GC_PAIR_FWD))) return false;
	if (sp < p.start || sp > p.end - 1) return false;
	uint32_t nb = gc_pair_nb(p);
	if (nb)
	{
		int32_t want = (p.seedPos - 1) - sp; // seqPos of such an entry inside the backward trace
		if (want >= -1)
		{
			uint32_t rnode, roff;
			gc_reverse_cell(pg, node, offset, rnode, roff);
			const uint64_t* t = tr + p.bwdOff;
			for (uint32_t i = gc_trace_lower(t, nb, want); i < nb && gc_te_seqpos(t[i]) == want; i++)
				if (gc_te_node(t[i]) == rnode && gc_te_offset(t[i]) == roff) return true;
		}
	}
	if (p.fwdLen)
	{
		int32_t want = sp - p.seedPos - 1;
		if (want >= -1)
		{
			const uint64_t* t = tr + p.fwdOff;
			for (uint32_t i = gc_trace_lower(t, p.fwdLen, want); i < p.fwdLen && gc_te_seqpos(t[i]) == want; i++)
				if (gc_te_node(t[i]) == node && gc_te_offset(t[i]) == offset) return true;
		}
	}
	return false;
}

// ---- S2: the seed loop of one fragment.  exts[0..n) are the fragment's window seeds in seed order, pairs[0..n) their
// extensions (all extended speculatively); kept[k] = 1 for the alignments the reference keeps (they become anchors).
// Returns false on an assertion-class condition (the reference leaves the fragment loop with cont = true, Aligner.cpp:695-703).
GC_HD bool gc_fragment_filter(const GcPostGraph& pg, const uint64_t* tr, const GcSeedCell* cells, const GcSeedExt* exts, const GcPair* pairs, uint32_t n, uint8_t* kept, uint32_t& seedsExtended)
{
	bool ok = true;
	seedsExtended = 0;
	for (uint32_t k = 0; k < n; k++) kept[k] = 0;
	for (uint32_t k = 0; k < n; k++)
	{
		const GcSeedCell c = cells[exts[k].cell];
		if (!(c.flags & 1)) continue; // seed.seedClusterSize < minClusterSize (GraphAligner.h:141)
		int32_t sp = c.seqPos - exts[k].fragStart;
		bool found = false, assertion = false;
		for (uint32_t j = 0; j < k; j++)
		{
			if (!kept[j]) continue;
			const GcPair& a = pairs[j];
			if (!(a.end - 1 > a.start)) { assertion = true; continue; } // the reference asserts trace.back().seqPos > trace[0].seqPos (GraphAligner.h:410)
			if (gc_pair_has_cell(pg, tr, a, c.node, c.offset, sp)) { found = true; break; }
		}
		if (assertion) { ok = false; break; }
		if (found) continue;
		seedsExtended++;
		const GcPair& p = pairs[k];
		if (p.flags & GC_PAIR_INTERNAL) ok = false;
		if (!(p.flags & (GC_PAIR_BWD | GC_PAIR_FWD))) continue;
		if (p.end == p.start) continue;
		kept[k] = 1;
	}
	return ok;
}

// anchor of a kept fragment alignment (Aligner.cpp:706-729): run-length-deduplicated split nodes of the trace
// + the in-node offsets of its first and last entry.  pathOut == nullptr: count only.
GC_HD uint32_t gc_anchor_path(const GcPostGraph& pg, const uint64_t* tr, const GcPair& p, uint32_t* pathOut, uint32_t& firstOffset, uint32_t& lastOffset)
{
	uint32_t n = gc_pair_size(p), len = 0, last = 0xFFFFFFFFu;
	GcRevCache cache;
	for (uint32_t k = 0; k < n; k++)
	{
		GcMergedEntry e = gc_pair_entry(pg, tr, p, k, &cache);
		if (len == 0 || e.node != last) { if (pathOut) pathOut[len] = e.node; len++; last = e.node; }
		if (k == 0) firstOffset = e.offset;
		if (k == n - 1) lastOffset = e.offset;
	}
	return len;
}

// traceToPoses + traceToSequence (Aligner.cpp:376-408, 425-428): the padded graph path of a whole-read alignment as
// base codes 0..3.  out == nullptr: length only.
GC_HD uint32_t gc_pair_path_string(const GcPostGraph& pg, const uint64_t* tr, const GcPair& p, uint8_t* out)
{
	uint32_t n = gc_pair_size(p), w = 0;
	uint32_t lastNode = 0, lastOffset = 0, lastLength = 0;
	GcRevCache cache;
	for (uint32_t j = 0; j < n; j++)
	{
		GcMergedEntry e = gc_pair_entry(pg, tr, p, j, &cache);
		if (j == 0)
		{
			lastNode = e.node; lastOffset = e.offset; lastLength = pg.nodeLength[e.node];
			if (out) out[w] = (uint8_t)gc_post_base(pg, lastNode, lastOffset);
			w++; lastOffset++;
		}
		else
		{
			if (e.node != lastNode)
			{
				while (lastOffset < lastLength) { if (out) out[w] = (uint8_t)gc_post_base(pg, lastNode, lastOffset); w++; lastOffset++; }
				lastNode = e.node; lastLength = pg.nodeLength[e.node]; lastOffset = 0;
			}
			while (lastOffset <= e.offset) { if (out) out[w] = (uint8_t)gc_post_base(pg, lastNode, lastOffset); w++; lastOffset++; }
		}
	}
	return w;
}

// pathToTrace (Aligner.cpp:409-424, with its compare-by-value tests) -> bases of a chained node path
GC_HD uint32_t gc_node_path_string(const GcPostGraph& pg, const uint32_t* path, uint32_t n, uint32_t firstOffset, uint32_t lastOffset, uint8_t* out)
{
	uint32_t w = 0;
	for (uint32_t i = 0; i < n; i++)
	{
		uint32_t node = path[i];
		uint32_t S = 0, L = pg.nodeLength[node];
		if (node == path[0]) S = firstOffset;
		else if (node == path[n - 1]) L = lastOffset + 1;
		for (uint32_t o = S; o < L; o++) { if (out) out[w] = (uint8_t)gc_post_base(pg, node, o); w++; }
	}
	return w;
}

// ---- edit runs of an alignment (GraphAlignerVGAlignment::traceToAlignment, GraphAlignerVGAlignment.h:37-165).
// The vg::Alignment of a trace is a list of mappings (one per visit of an original node) each with runs of
// match / mismatch / insertion / deletion steps.  Token stream (uint32 words):
//   mapping : 0, digraph node id, offset in the original node
//   edit    : type << 30 | run length (>= 1)            type: 0 match, 1 mismatch, 2 insertion, 3 deletion
// The characters of mismatch and insertion runs are consecutive read characters, so the host that frames the
// protobuf message takes them from the read (first-entry quirk of :75 included).
#define GC_EDIT_MATCH 0u
#define GC_EDIT_MISMATCH 1u
#define GC_EDIT_INSERTION 2u
#define GC_EDIT_DELETION 3u
struct GcTokenStep { int32_t node; uint32_t nodeOffset; int32_t seqPos; bool nodeSwitch; bool match; }; // node = digraph id, nodeOffset in the original node
struct GcTokenCounts { uint32_t matches, mismatches, insertions, deletions, tokens; };

// Src: GcTokenStep operator()(uint32_t pos).  out == nullptr: count only.
template <typename Src>
GC_HD GcTokenCounts gc_tokenize(Src& src, uint32_t n, uint32_t* out)
{
	GcTokenCounts c; c.matches = c.mismatches = c.insertions = c.deletions = c.tokens = 0;
	if (n == 0) return c;
	const uint32_t EMPTY = 4;
	GcTokenStep prev = src(0);
	int32_t curNode = prev.node; uint32_t curOffset = prev.nodeOffset;
	uint32_t w = 0;
	if (out) { out[w] = 0; out[w + 1] = (uint32_t)curNode; out[w + 2] = curOffset; }
	w += 3;
	uint32_t type = prev.match ? GC_EDIT_MATCH : GC_EDIT_MISMATCH, len = 1;
	if (prev.match) c.matches++; else c.mismatches++;
	for (uint32_t pos = 1; pos < n; pos++)
	{
		GcTokenStep e = src(pos);
		bool inside = !prev.nodeSwitch || (e.node == curNode && e.nodeOffset > curOffset);
		if (!inside)
		{
			if (out) out[w] = (type << 30) | len;
			w++;
			curNode = e.node; curOffset = e.nodeOffset;
			if (out) { out[w] = 0; out[w + 1] = (uint32_t)curNode; out[w + 2] = curOffset; }
			w += 3;
			type = EMPTY; len = 0;
		}
		uint32_t t;
		if (prev.seqPos == e.seqPos) { t = GC_EDIT_DELETION; c.deletions++; }
		else if (inside && prev.nodeOffset == e.nodeOffset) { t = GC_EDIT_INSERTION; c.insertions++; }
		else if (e.match) { t = GC_EDIT_MATCH; c.matches++; }
		else { t = GC_EDIT_MISMATCH; c.mismatches++; }
		if (type == EMPTY) type = t;
		else if (type != t) { if (out) out[w] = (type << 30) | len; w++; type = t; len = 0; }
		len++;
		prev = e;
	}
	if (out) out[w] = (type << 30) | len;
	w++;
	c.tokens = w;
	return c;
}

// token source over a pair's merged trace: materialize (fixReverseTraceSeqPosAndOrder / fixForwardTraceSeqPos, GraphAligner.h:527-565)
struct GcPairTokenSrc
{
	const GcPostGraph* pg; const uint64_t* tr; const GcPair* p; const uint8_t* codes; // forward IUPAC codes of the aligned sequence
	GC_HD GcTokenStep operator()(uint32_t k) const
	{
		GcTokenStep s;
		uint32_t nb = gc_pair_nb(*p);
		int graphBase;
		if (k < nb)
		{
			uint64_t t = tr[p->bwdOff + k];
			uint32_t node = gc_te_node(t), off = gc_te_offset(t);
			s.node = pg->nodeIDs[node] ^ 1;
			s.nodeOffset = pg->revLast[node] - off;
			s.seqPos = (p->seedPos - 1) - gc_te_seqpos(t);
			s.nodeSwitch = (k + 1 < p->bwdLen) ? gc_te_switch(tr[p->bwdOff + k + 1]) : false;
			graphBase = 3 - gc_post_base(*pg, node, off);
		}
		else
		{
			uint64_t t = tr[p->fwdOff + (p->fwdLen - 1 - (k - nb))];
			uint32_t node = gc_te_node(t), off = gc_te_offset(t);
			s.node = pg->nodeIDs[node];
			s.nodeOffset = pg->nodeOffset[node] + off;
			s.seqPos = gc_te_seqpos(t) + p->seedPos + 1;
			s.nodeSwitch = gc_te_switch(t);
			graphBase = gc_post_base(*pg, node, off);
		}
		s.match = ((codes[s.seqPos] >> graphBase) & 1) != 0; // Common::characterMatch on the encoded forms
		return s;
	}
};
