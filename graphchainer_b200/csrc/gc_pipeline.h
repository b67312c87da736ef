// Host driver of the per-read pipeline: the body of runComponentMappings
// (src/Aligner.cpp:492-1062, colinear mode) re-organised for batches of reads, with the
// three dynamic-programming stages delegated to libgcgpu through its C ABI:
//   S0  seeding + clustering               host   (gc_seeder.h)
//   S1  whole-read seed-and-extend         host seed loop (GraphAligner.h:114-203) in ROUNDS,
//                                          extensions = gcgpu_extend (K1)
//   S1b distance(GA path, read)            gcgpu_nw (K3)                     Aligner.cpp:642-654
//   S2  fragment anchoring                 all window seeds extended speculatively by one
//                                          gcgpu_extend call, then the reference's in-order
//                                          exactAlignmentPart filter      Aligner.cpp:668-730
//   S3  co-linear chaining                 gcgpu_chain (K2)                  Aligner.cpp:735
//   S4  chain -> node path                 host BFS getChainPath             Aligner.cpp:748-822
//   S5  NW(path, read)                     gcgpu_nw (K3): distance for every read, the edit
//                                          path only when the chained alignment wins (S6)
//   S6  decision, S7 vg::Alignment         host                              Aligner.cpp:880-1013
// There is no CPU implementation of K1/K2/K3 here: without libgcgpu nothing aligns.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>
#ifdef GC_PROF
#include <x86intrin.h>
static double g_prof[32]; static const char* g_profName[32];
struct GcProfScope { int id; unsigned long long t0; GcProfScope(int id, const char* n) : id(id), t0(__rdtsc()) { g_profName[id] = n; } ~GcProfScope() { g_prof[id] += (double)(__rdtsc() - t0); } };
#define GC_PROF_SCOPE(id, name) GcProfScope _prof##id(id, name)
#else
#define GC_PROF_SCOPE(id, name)
#endif

#include "../../include/gcgpu.h"
#include "gc_host_graph.h"
#include "gc_seeder.h"

struct GcRead
{
	std::string name;     // full FASTA/FASTQ header after '>' / '@'
	std::string sequence;
};

// GraphAlignerCommon::TraceItem (GraphAlignerCommon.h:127-160) after the seqPos/node fix-ups
struct GcTraceItem
{
	int32_t node;          // digraph node id (2*id + strand)
	uint32_t nodeOffset;   // offset in the original node
	int64_t seqPos;
	bool nodeSwitch;
	char sequenceCharacter;
	char graphCharacter;
};

struct GcAlnItem
{
	std::vector<GcTraceItem> trace;
	int32_t traceScore = 0;       // OnewayTrace::score (what AddAlignment serialises)
	size_t alignmentStart = 0, alignmentEnd = 0;
	size_t alignmentScore = 0;
	size_t seedGoodness = 0;
};


// One seed extension (getAlignmentFromSeed, GraphAligner.h:567-626) as the two K1 traces the kernel
// wrote, still packed: the merged GraphAligner trace is only materialised for alignments that
// survive selection.  Merged order = backward part (kernel order; its last entry, the seed cell,
// is dropped when a forward part exists, GraphAligner.h:599) then the forward part reversed.
struct GcPackedAln
{
	const uint64_t* bwd = nullptr; uint32_t bwdLen = 0;
	const uint64_t* fwd = nullptr; uint32_t fwdLen = 0;
	int32_t bwdScore = 0, fwdScore = 0;
	int64_t seedPos = 0;          // seed.seqPos in the coordinates of the aligned sequence
	int32_t traceScore = 0;
	size_t alignmentStart = 0, alignmentEnd = 0, alignmentScore = 0, seedGoodness = 0;
	uint32_t bwdUsed() const { return bwdLen ? (fwdLen ? bwdLen - 1 : bwdLen) : 0; }
	uint32_t size() const { return bwdUsed() + fwdLen; }
	int64_t seqPosAt(uint32_t k) const
	{
		uint32_t nb = bwdUsed();
		if (k < nb) return (seedPos - 1) - (int64_t)GCGPU_TRACE_SEQPOS(bwd[k]);
		return (int64_t)GCGPU_TRACE_SEQPOS(fwd[fwdLen - 1 - (k - nb)]) + seedPos + 1;
	}
};

struct GcReadResult
{
	std::vector<GcAlnItem> alignments; // final, sorted by alignmentStart
	bool usedChain = false;            // S6: the chained (CLC) alignment was strictly better
	bool dropped = false;              // assertion-class failure: the reference drops the read
	// the fields of the reference's --short-verbose line (Aligner.cpp:909-915)
	size_t anchors = 0, chained = 0, pathBp = 0, clcScore = 0, longEditDistance = 0;
	bool hasLong = false;
	size_t seedsFound = 0, seedsExtended = 0;
};

struct GcPipelineParams
{
	double minimizerSeedDensity = 10;
	size_t seedClusterMinSize = 1;
	long long colinearGap = 10000;
	long long colinearSplitLen = 35;
	long long colinearSplitGap = 35;
	bool tryAllSeeds = true;
	size_t s1FirstRoundSeeds = 1;   // S1 speculation: seeds extended per read in the first round ...
	size_t s1LaterRoundSeeds = 8;   // ... in the second round ...
	size_t s1TailRoundSeeds = 32;  // ... and from the third round on: few reads get that far, their rounds cost one extension latency each whatever the item count
};

struct GcPipelineStats
{
	uint64_t k1Items = 0, k1Columns = 0, k1Launches = 0;
	uint64_t k3Items = 0, k3Blocks = 0;
	uint64_t k2Reads = 0, k2Anchors = 0;
	double k1Ms = 0, k2Ms = 0, k3Ms = 0, s0Ms = 0;
	double hostSeedMs = 0, hostS1Ms = 0, hostS2Ms = 0, hostConnectMs = 0;
	uint64_t s1Rounds = 0;
	uint64_t s1Wasted = 0; // speculative S1 seed extensions whose result was discarded
};

// hands the reference's minimizer index (MinimizerSeeder.h:17-29 as flat arrays) to a libgcgpu context
inline int gcUploadMinimizerIndex(gcgpu_ctx* ctx, const GcHostGraph& g)
{
	gcgpu_minimizer_index mi;
	mi.k = (uint32_t)g.mzLength; mi.window = (uint32_t)g.mzWindow; mi.max_count = g.mzMaxCount;
	mi.num_kmers = g.mzKmers.size(); mi.kmers = g.mzKmers.data(); mi.kmer_start = g.mzKmerStart.data();
	return gcgpu_set_minimizer_index(ctx, &mi);
}

namespace gcpipe {

inline char complementChar(char c)
{
	switch (c)
	{
		case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'T': case 't': return 'A'; case 'G': case 'g': return 'C';
		case 'N': case 'n': return 'N'; case 'U': case 'u': return 'A'; case 'R': case 'r': return 'Y'; case 'Y': case 'y': return 'R';
		case 'K': case 'k': return 'M'; case 'M': case 'm': return 'K'; case 'S': case 's': return 'S'; case 'W': case 'w': return 'W';
		case 'B': case 'b': return 'V'; case 'V': case 'v': return 'B'; case 'D': case 'd': return 'H'; case 'H': case 'h': return 'D';
	}
	return 0; // the reference asserts (CommonUtils.cpp:131)
}
inline uint8_t complementMask(uint8_t m) { return (uint8_t)(((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3)); }

// digraph node id + offset in the original node of merged-trace entry k
inline void packedNodePos(const GcHostGraph& g, const GcPackedAln& a, uint32_t k, int& node, size_t& nodeOffset)
{
	uint32_t nb = a.bwdUsed();
	if (k < nb)
	{
		uint64_t e = a.bwd[k];
		uint32_t sn = GCGPU_TRACE_NODE(e);
		auto rp = g.reversePosition(g.nodeIDs[sn], (size_t)g.nodeOffset[sn] + GCGPU_TRACE_OFFSET(e));
		node = rp.first; nodeOffset = rp.second;
	}
	else
	{
		uint64_t e = a.fwd[a.fwdLen - 1 - (k - nb)];
		uint32_t sn = GCGPU_TRACE_NODE(e);
		node = g.nodeIDs[sn]; nodeOffset = (size_t)g.nodeOffset[sn] + GCGPU_TRACE_OFFSET(e);
	}
}
// split node of merged-trace entry k (= GetUnitigNode(node, nodeOffset))
inline size_t packedSplitNode(const GcHostGraph& g, const GcPackedAln& a, uint32_t k, GcHostGraph::UnitigCache& cache)
{
	uint32_t nb = a.bwdUsed();
	if (k >= nb) return GCGPU_TRACE_NODE(a.fwd[a.fwdLen - 1 - (k - nb)]);
	int node; size_t off;
	packedNodePos(g, a, k, node, off);
	return g.unitigNode(node, off, cache);
}

// exactAlignmentPart (GraphAligner.h:407-461): is the seed cell on the trace of `aln`?
inline bool exactAlignmentPart(const GcHostGraph& g, const GcPackedAln& aln, const GcSeedHit& seed, bool& assertion)
{
	uint32_t n = aln.size();
	if (n == 0 || !(aln.seqPosAt(n - 1) > aln.seqPosAt(0))) { assertion = true; return false; }
	int64_t sp = (int64_t)seed.seqPos;
	if (aln.seqPosAt(n - 1) < sp) return false;
	if (aln.seqPosAt(0) > sp) return false;
	// seqPos is non-decreasing with unit steps: find the run of items at seqPos == sp
	uint32_t lo = 0, hi = n;
	while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (aln.seqPosAt(mid) < sp) lo = mid + 1; else hi = mid; }
	int compareNode = seed.nodeID * 2 + (seed.reverse ? 1 : 0);
	for (uint32_t i = lo; i < n && aln.seqPosAt(i) == sp; i++)
	{
		int node; size_t off;
		packedNodePos(g, aln, i, node, off);
		if (node == compareNode && off == seed.nodeOffset) return true;
	}
	return false;
}

// AlignmentSelection::alignmentIncompatible (AlignmentSelection.cpp:13-31)
template <typename Aln>
inline bool alignmentIncompatible(const Aln& left, const Aln& right)
{
	const float OverlapIncompatibleFractionCutoff = 0.05;
	auto minOverlapLen = std::min((left.alignmentEnd - left.alignmentStart), (right.alignmentEnd - right.alignmentStart)) * OverlapIncompatibleFractionCutoff;
	size_t leftStart = left.alignmentStart, leftEnd = left.alignmentEnd, rightStart = right.alignmentStart, rightEnd = right.alignmentEnd;
	if (leftStart > rightStart) { std::swap(leftStart, rightStart); std::swap(leftEnd, rightEnd); }
	int overlap = 0;
	if (leftEnd > rightStart) overlap = leftEnd - rightStart;
	return overlap > minOverlapLen;
}
// GreedySelectAlignments with alignmentLengthCompare (AlignmentSelection.h:36-55, .cpp:45-51)
template <typename Aln>
inline std::vector<Aln> selectGreedyLength(const std::vector<Aln>& alignments)
{
	std::vector<size_t> items;
	for (size_t i = 0; i < alignments.size(); i++) items.push_back(i);
	std::sort(items.begin(), items.end(), [&alignments](size_t l, size_t r)
	{
		const Aln& left = alignments[l]; const Aln& right = alignments[r];
		if ((left.alignmentEnd - left.alignmentStart) > (right.alignmentEnd - right.alignmentStart)) return true;
		if ((right.alignmentEnd - right.alignmentStart) > (left.alignmentEnd - left.alignmentStart)) return false;
		if (left.alignmentScore < right.alignmentScore) return true;
		return false;
	});
	std::vector<Aln> result;
	for (auto i : items)
	{
		if (!std::any_of(result.begin(), result.end(), [&alignments, i](const Aln& existing) { return alignmentIncompatible(existing, alignments[i]); }))
			result.push_back(alignments[i]);
	}
	return result;
}

// traceToPoses + traceToSequence (Aligner.cpp:376-408, 425-428): the padded graph path of a GA alignment
inline std::string traceToSequence(const GcHostGraph& g, const GcAlnItem& aln)
{
	std::string ret;
	size_t lastNode = 0, lastOffset = 0, lastLength = 0;
	GcHostGraph::UnitigCache ucache;
	ret.reserve(aln.trace.size() + 64);
	for (size_t j = 0; j < aln.trace.size(); j++)
	{
		size_t node = g.unitigNode(aln.trace[j].node, aln.trace[j].nodeOffset, ucache);
		size_t nodeOffset = aln.trace[j].nodeOffset - g.nodeOffset[node];
		if (j == 0)
		{
			lastNode = node; lastOffset = nodeOffset; lastLength = g.nodeLength[node];
			ret.push_back(g.nodeChar((uint32_t)lastNode, (uint32_t)lastOffset));
			lastOffset++;
		}
		else
		{
			if (node != lastNode)
			{
				while (lastOffset < lastLength) { ret.push_back(g.nodeChar((uint32_t)lastNode, (uint32_t)lastOffset)); lastOffset++; }
				lastNode = node; lastLength = g.nodeLength[node]; lastOffset = 0;
			}
			while (lastOffset <= nodeOffset) { ret.push_back(g.nodeChar((uint32_t)lastNode, (uint32_t)lastOffset)); lastOffset++; }
		}
	}
	return ret;
}

struct MatrixPos { size_t node; size_t nodeOffset; size_t seqPos; };

// pathToTrace (Aligner.cpp:409-424), including its single-node quirk
inline std::vector<MatrixPos> pathToTrace(const GcHostGraph& g, const std::vector<size_t>& path, size_t firstNodeOffset, size_t lastNodeOffset)
{
	std::vector<MatrixPos> ret;
	for (size_t node : path)
	{
		size_t S = 0, L = g.nodeLength[node];
		if (node == path[0]) S = firstNodeOffset;
		else if (node == path.back()) L = lastNodeOffset + 1;
		MatrixPos p { node, S, 0 };
		while (p.nodeOffset < L) { ret.push_back(p); p.nodeOffset++; }
	}
	return ret;
}

// AlignmentGraph::getChainPath (AlignmentGraph.cpp:1866-1916): FIFO BFS with the unsigned distance prune
struct ChainPathScratch { std::vector<size_t> vis, dis, Q, pre; size_t flag = 1; };
inline std::vector<size_t> getChainPath(const GcHostGraph& g, ChainPathScratch& s, size_t S, size_t T, long long sep_limit)
{
	size_t N = g.numNodes();
	if (s.vis.size() < N) { s.vis.resize(N, 0); s.pre.resize(N); s.dis.resize(N); s.Q.reserve(1024); }
	s.Q.clear();
	s.Q.push_back(S);
	s.vis[S] = ++s.flag;
	s.dis[S] = 0;
	for (size_t i = 0; s.vis[T] != s.flag && i < s.Q.size(); )
	{
		size_t v = s.Q[i++];
		if (s.dis[v] > (size_t)sep_limit) continue; // size_t vs long long comparison, AlignmentGraph.cpp:1897
		for (uint32_t e = g.outStart[v]; e < g.outStart[v + 1]; e++)
		{
			size_t t = g.outNbr[e];
			if (s.vis[t] != s.flag)
			{
				s.Q.push_back(t);
				s.vis[t] = s.flag;
				s.dis[t] = s.dis[v] + g.nodeLength[t];
				s.pre[t] = v;
			}
		}
	}
	std::vector<size_t> tmp;
	if (s.vis[T] != s.flag) return tmp;
	for (size_t i = T; i != S; i = s.pre[i]) tmp.push_back(i);
	tmp.push_back(S);
	std::reverse(tmp.begin(), tmp.end());
	return tmp;
}

}

class GcPipeline
{
public:
	GcPipeline(const GcHostGraph& graph, gcgpu_ctx* ctx, const GcPipelineParams& params) : g(graph), ctx(ctx), params(params) {}
	GcPipelineStats stats;

	void alignBatch(const std::vector<GcRead>& reads, std::vector<GcReadResult>& out);

private:
	const GcHostGraph& g;
	gcgpu_ctx* ctx;
	GcPipelineParams params;

	struct ExtRef { int32_t item[2]; }; // indices of the backward / forward work items, -1 if absent
	struct Batch
	{
		uint8_t* codes = nullptr;            // per read: forward masks then reverse-complement masks (page-locked)
		size_t codesBytes = 0;
		std::vector<uint64_t> fwdOff, rcOff; // offsets into codes
	};

	// page-locked host buffers for the kernel outputs, grow-only, reused across batches:
	// tracePool[k] receives the packed traces of the k-th gcgpu_extend call of a batch and stays
	// valid until the packed alignments that point into it have been consumed
	struct Pinned
	{
		void* p = nullptr; size_t cap = 0;
		void ensure(size_t bytes)
		{
			if (bytes <= cap) return;
			if (p) gcgpu_host_free(p);
			cap = bytes + bytes / 4 + (1 << 20);
			p = gcgpu_host_alloc(cap);
			if (!p) { cap = 0; throw std::runtime_error("gcgpu_host_alloc failed"); }
		}
		~Pinned() { if (p) gcgpu_host_free(p); }
		Pinned() = default;
		Pinned(const Pinned&) = delete;
		Pinned& operator=(const Pinned&) = delete;
	};
	std::vector<std::unique_ptr<Pinned>> tracePool;
	Pinned codesBuf, seedBuf, nwPinned;

	void stats_s1Wasted_add(size_t n) { if (n) { _Pragma("omp atomic") stats.s1Wasted += n; } }
	void check(int rc, const char* what)
	{
		if (rc != GCGPU_OK) throw std::runtime_error(std::string(what) + " failed: " + gcgpu_last_error());
	}

	// the two K1 work items of one seed (getTwoDirectionalTrace, GraphAligner.h:480-525).
	// seqStart/seqLen delimit `sequence` inside the read (whole read, or one fragment).
	ExtRef makeItems(const Batch& b, size_t r, size_t readLen, size_t seqStart, size_t seqLen, const GcSeedHit& seed, std::vector<gcgpu_ext_item>& items) const
	{
		ExtRef ref; ref.item[0] = ref.item[1] = -1;
		int forwardNodeId = seed.nodeID * 2 + (seed.reverse ? 1 : 0);
		if (seed.seqPos > 0)
		{
			auto reversePos = g.reversePosition(forwardNodeId, seed.nodeOffset);
			uint32_t node = g.unitigNode(reversePos.first, reversePos.second);
			gcgpu_ext_item it;
			// revcomp(sequence) = rc(read)[readLen - seqStart - seqLen, readLen - seqStart); its last seqPos characters
			it.seq_offset = b.rcOff[r] + (readLen - seqStart - seed.seqPos);
			it.seq_len = (int32_t)seed.seqPos;
			it.node = node;
			it.offset = (uint32_t)(reversePos.second - g.nodeOffset[node]);
			it.reserved = 0;
			ref.item[0] = (int32_t)items.size();
			items.push_back(it);
		}
		if (seed.seqPos < seqLen - 1)
		{
			uint32_t node = g.unitigNode(forwardNodeId, seed.nodeOffset);
			gcgpu_ext_item it;
			it.seq_offset = b.fwdOff[r] + seqStart + seed.seqPos + 1;
			it.seq_len = (int32_t)(seqLen - seed.seqPos - 1);
			it.node = node;
			it.offset = (uint32_t)(seed.nodeOffset - g.nodeOffset[node]);
			it.reserved = 0;
			ref.item[1] = (int32_t)items.size();
			items.push_back(it);
		}
		return ref;
	}

	// getAlignmentFromSeed (GraphAligner.h:567-626) from the two K1 results, kept packed.  `seed.seqPos` is in
	// the coordinates of the aligned sequence (whole read or fragment).  Returns false if both failed.
	bool buildAlignment(const GcSeedHit& seed, const ExtRef& ref, const gcgpu_ext_result* results, const uint64_t* traces, GcPackedAln& out) const
	{
		bool haveB = ref.item[0] >= 0 && results[ref.item[0]].status == GCGPU_ITEM_OK;
		bool haveF = ref.item[1] >= 0 && results[ref.item[1]].status == GCGPU_ITEM_OK;
		if (!haveB && !haveF) return false;
		out = GcPackedAln();
		out.seedPos = (int64_t)seed.seqPos;
		if (haveB) { const gcgpu_ext_result& rb = results[ref.item[0]]; out.bwd = traces + rb.trace_offset; out.bwdLen = rb.trace_len; out.bwdScore = rb.score; }
		if (haveF) { const gcgpu_ext_result& rf = results[ref.item[1]]; out.fwd = traces + rf.trace_offset; out.fwdLen = rf.trace_len; out.fwdScore = rf.score; }
		out.traceScore = (haveB ? out.bwdScore : 0) + (haveF ? out.fwdScore : 0);
		out.alignmentScore = (size_t)out.traceScore;
		out.alignmentStart = (size_t)out.seqPosAt(0);
		out.alignmentEnd = (size_t)out.seqPosAt(out.size() - 1) + 1;
		out.seedGoodness = seed.seedGoodness;
		return true;
	}

	// the merged GraphAligner trace of a packed alignment: fixReverseTraceSeqPosAndOrder (GraphAligner.h:543-565),
	// getTwoDirectionalTrace (:522), fixForwardTraceSeqPos (:527-540)
	void materialize(const char* sequence, const GcPackedAln& a, GcAlnItem& out) const
	{
		uint32_t nb = a.bwdUsed(), n = a.size();
		out.trace.resize(n);
		uint32_t lastNode = 0xFFFFFFFFu; int lastRevId = 0; size_t lastRevEnd = 0; // reversePosition(id, o) = (id ^ 1, origSize - 1 - o): constant per kernel node up to the offset
		for (uint32_t i = 0; i < nb; i++)
		{
			uint64_t e = a.bwd[i];
			uint32_t node = GCGPU_TRACE_NODE(e), off = GCGPU_TRACE_OFFSET(e);
			GcTraceItem& it = out.trace[i];
			it.seqPos = (a.seedPos - 1) - (int64_t)GCGPU_TRACE_SEQPOS(e);
			if (node != lastNode)
			{
				auto reversePos = g.reversePosition(g.nodeIDs[node], (size_t)g.nodeOffset[node]);
				lastNode = node; lastRevId = reversePos.first; lastRevEnd = reversePos.second;
			}
			it.node = lastRevId;
			it.nodeOffset = (uint32_t)(lastRevEnd - off);
			it.sequenceCharacter = sequence[it.seqPos];
			it.graphCharacter = gcpipe::complementChar(g.nodeChar(node, off));
			it.nodeSwitch = (i + 1 < a.bwdLen) ? GCGPU_TRACE_SWITCH(a.bwd[i + 1]) != 0 : false;
		}
		for (uint32_t i = 0; i < a.fwdLen; i++)
		{
			uint64_t e = a.fwd[a.fwdLen - 1 - i];
			uint32_t node = GCGPU_TRACE_NODE(e), off = GCGPU_TRACE_OFFSET(e);
			GcTraceItem& it = out.trace[nb + i];
			it.seqPos = (int64_t)GCGPU_TRACE_SEQPOS(e) + a.seedPos + 1;
			it.node = g.nodeIDs[node];
			it.nodeOffset = g.nodeOffset[node] + off;
			it.nodeSwitch = GCGPU_TRACE_SWITCH(e) != 0;
			it.sequenceCharacter = sequence[it.seqPos];
			it.graphCharacter = g.nodeChar(node, off);
		}
		out.traceScore = a.traceScore;
		out.alignmentScore = a.alignmentScore;
		out.alignmentStart = a.alignmentStart;
		out.alignmentEnd = a.alignmentEnd;
		out.seedGoodness = a.seedGoodness;
	}
};

// ------------------------------------------------------------------------------------------
inline void GcPipeline::alignBatch(const std::vector<GcRead>& reads, std::vector<GcReadResult>& out)
{
	size_t R = reads.size();
	out.assign(R, GcReadResult());
	if (R == 0) return;
	const bool traceOn = getenv("GC_TRACE") != nullptr;
	auto wallNow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double tPhase = wallNow();
	double devMs = 0; // wall time spent inside libgcgpu calls during the current phase
	auto phase = [&](const char* name) { if (traceOn) { double n = wallNow(); fprintf(stderr, "[gc] phase %-10s %.2f ms (libgcgpu calls %.2f ms, host %.2f ms)\n", name, n - tPhase, devMs, n - tPhase - devMs); tPhase = n; devMs = 0; } };
	// ---- encode reads (forward + reverse complement IUPAC masks)
	Batch b;
	b.fwdOff.resize(R); b.rcOff.resize(R);
	{
		size_t total = 0;
		for (size_t r = 0; r < R; r++) { b.fwdOff[r] = total; total += reads[r].sequence.size(); b.rcOff[r] = total; total += reads[r].sequence.size(); }
		codesBuf.ensure(total + 8);
		b.codes = (uint8_t*)codesBuf.p; b.codesBytes = total + 8;
		memset(b.codes + total, 0, 8);
		#pragma omp parallel for schedule(dynamic, 16)
		for (size_t r = 0; r < R; r++)
		{
			const std::string& s = reads[r].sequence;
			size_t L = s.size();
			for (size_t i = 0; i < L; i++)
			{
				uint8_t m = gcEncodeSeedBase(s[i]);
				b.codes[b.fwdOff[r] + i] = m;
				b.codes[b.rcOff[r] + (L - 1 - i)] = gcpipe::complementMask(m);
			}
		}
	}
	// ---- S0: seeds (the reference calls getSeeds + OrderSeeds twice per read with identical results).
	// k-mer walk + index probes on the device (gcgpu_seed uploads the read codes, which stay resident for K1);
	// the count sort, density cut, seed-hit expansion and clustering per read on the host
	std::vector<std::vector<GcSeedHit>> seedsOrdered(R);
	{
		std::vector<gcgpu_seed_read> sr(R);
		for (size_t r = 0; r < R; r++) { sr[r].seq_offset = b.fwdOff[r]; sr[r].seq_len = (int32_t)reads[r].sequence.size(); sr[r].reserved = 0; }
		std::vector<uint64_t> matchOff(R + 1, 0);
		uint64_t used = 0;
		double tDev = wallNow();
		check(gcgpu_seed(ctx, b.codes, b.codesBytes, sr.data(), (uint32_t)R, matchOff.data(), nullptr, 0, &used), "gcgpu_seed");
		seedBuf.ensure((used + 1) * sizeof(gcgpu_seed_match));
		check(gcgpu_fetch_seed_matches(ctx, (gcgpu_seed_match*)seedBuf.p, 0, used), "gcgpu_fetch_seed_matches");
		devMs += wallNow() - tDev;
		stats.s0Ms += gcgpu_last_kernel_ms(ctx);
		const gcgpu_seed_match* matches = (const gcgpu_seed_match*)seedBuf.p;
		#pragma omp parallel for schedule(dynamic, 4)
		for (size_t r = 0; r < R; r++)
		{
			std::vector<std::tuple<size_t, size_t, size_t, size_t>> matchIndices;
			matchIndices.reserve(matchOff[r + 1] - matchOff[r]);
			for (uint64_t i = matchOff[r]; i < matchOff[r + 1]; i++) matchIndices.emplace_back((size_t)matches[i].pos, (size_t)0, (size_t)matches[i].start, (size_t)matches[i].count);
			{ GC_PROF_SCOPE(0, "seed.seedsFromMatches"); seedsOrdered[r] = gcseed::seedsFromMatches(g, matchIndices, reads[r].sequence.size(), params.minimizerSeedDensity); }
			out[r].seedsFound = 2 * seedsOrdered[r].size(); // counted in align_fn and again for the split pass (Aligner.cpp:553,663)
			GC_PROF_SCOPE(1, "seed.orderSeeds");
			if (!seedsOrdered[r].empty()) gcseed::orderSeeds(g, seedsOrdered[r]);
		}
	}
	phase("seed");
	std::vector<gcgpu_ext_item> items;
	std::vector<gcgpu_ext_result> results;
	const uint64_t* traces = nullptr;
	size_t extendCalls = 0;
	auto runExtend = [&]()
	{
		results.resize(items.size());
		if (tracePool.size() <= extendCalls) tracePool.emplace_back(new Pinned());
		Pinned& buf = *tracePool[extendCalls];
		uint64_t used = 0;
		double tDev = wallNow();
		// the read codes were uploaded by gcgpu_seed and stay resident (seq == NULL);
		// two-phase call: the page-locked trace buffer is sized from what the extensions really produced
		int rc = gcgpu_extend(ctx, nullptr, b.codesBytes, items.data(), (uint32_t)items.size(), results.data(), nullptr, 0, &used);
		if (rc != GCGPU_OK && rc != GCGPU_ERR_INTERNAL) check(rc, "gcgpu_extend");
		buf.ensure((used + 1) * 8);
		check(gcgpu_fetch_traces(ctx, (uint64_t*)buf.p, 0, used), "gcgpu_fetch_traces");
		devMs += wallNow() - tDev;
		traces = (const uint64_t*)buf.p;
		extendCalls++;
		stats.k1Items += items.size();
		stats.k1Ms += gcgpu_last_kernel_ms(ctx);
		stats.k1Launches++;
		uint64_t cols = 0;
		for (const auto& r : results) cols += r.columns;
		stats.k1Columns += cols;
		if (traceOn) fprintf(stderr, "[gc] extend items=%zu kernel_ms=%.3f columns=%llu trace_MB=%.1f\n", items.size(), (double)gcgpu_last_kernel_ms(ctx), (unsigned long long)cols, used * 8 / 1e6);
	};

	// ---- S1: whole-read alignment, AlignOneWay(seeds, sloppy=true) (GraphAligner.h:114-203).
	// The reference walks the seeds in goodness order and decides, from the alignments kept so far,
	// whether to extend each one.  An extension is a pure function of (read, seed), so every ROUND
	// extends, for every read, the next few seeds that pass the skip rules under the current state;
	// the results are then consumed strictly in seed order with the rules re-evaluated exactly as
	// the reference does -- a speculative result whose seed turns out to be skipped is discarded.
	struct S1Cand { size_t seedIdx; ExtRef ref; };
	struct S1State
	{
		size_t i = 0; std::vector<GcPackedAln> alns; size_t seedsExtended = 0; size_t seedScoreForEndToEndAln = 0; bool done = false; std::vector<S1Cand> cands; size_t round = 0;
		std::vector<gcgpu_ext_item> localItems; size_t itemBase = 0;
		// memo of the skip rules: alignments are only ever added and never change, so "this seed is skipped" is permanent and
		// "no alignment so far contains this seed cell" only needs the alignments added since it was last asked
		std::vector<GcPackedAln> alnsAdded;   // insertion order (alns is kept sorted by alignmentStart like the reference's vector)
		std::vector<uint8_t> skip;            // per seed: 1 = a skip rule fired
		std::vector<uint32_t> checked;        // per seed: alnsAdded[0..checked) do not contain its cell
		bool degenerate = false;              // an alignment on which exactAlignmentPart asserts exists: evaluate in reference order
	};
	std::vector<S1State> s1(R);
	for (size_t r = 0; r < R; r++) { if (seedsOrdered[r].empty()) s1[r].done = true; else { s1[r].skip.assign(seedsOrdered[r].size(), 0); s1[r].checked.assign(seedsOrdered[r].size(), 0); } }
	// 0 = extend, 1 = skip, 2 = stop the seed loop, 3 = assertion (read dropped)
	auto seedRule = [&](S1State& st, const GcSeedHit& seed, size_t idx) -> int
	{
		if (seed.seedGoodness < st.seedScoreForEndToEndAln) return 2;
		if (seed.seedClusterSize < params.seedClusterMinSize) return 1;
		if (!st.degenerate)
		{
			if (st.skip[idx]) return 1;
			for (const auto& aln : st.alns)
				if (aln.alignmentStart <= seed.seqPos && aln.alignmentEnd >= seed.seqPos && aln.seedGoodness > seed.seedGoodness) { st.skip[idx] = 1; return 1; }
			bool assertion = false;
			for (uint32_t k = st.checked[idx]; k < st.alnsAdded.size(); k++)
				if (gcpipe::exactAlignmentPart(g, st.alnsAdded[k], seed, assertion)) { st.skip[idx] = 1; return 1; }
			st.checked[idx] = (uint32_t)st.alnsAdded.size();
			return 0;
		}
		for (const auto& aln : st.alns)
			if (aln.alignmentStart <= seed.seqPos && aln.alignmentEnd >= seed.seqPos && aln.seedGoodness > seed.seedGoodness) return 1;
		bool assertion = false;
		for (const auto& aln : st.alns) { if (gcpipe::exactAlignmentPart(g, aln, seed, assertion)) return 1; if (assertion) return 3; }
		return 0;
	};
	while (true)
	{
		items.clear();
		std::vector<size_t> active;
		// candidates are collected per read in parallel (work items local to the read), then concatenated
		#pragma omp parallel for schedule(dynamic, 8)
		for (size_t r = 0; r < R; r++)
		{
			S1State& st = s1[r];
			if (st.done) continue;
			GC_PROF_SCOPE(2, "s1.collect");
			const std::vector<GcSeedHit>& seedHits = seedsOrdered[r];
			size_t want = st.round == 0 ? params.s1FirstRoundSeeds : (st.round == 1 ? params.s1LaterRoundSeeds : params.s1TailRoundSeeds);
			st.cands.clear();
			st.localItems.clear();
			for (size_t i = st.i; i < seedHits.size() && st.cands.size() < want; i++)
			{
				int rule = seedRule(st, seedHits[i], i);
				if (rule >= 2) break; // decided again, in order, when the results are consumed
				if (rule == 1) continue;
				S1Cand c; c.seedIdx = i;
				c.ref = makeItems(b, r, reads[r].sequence.size(), 0, reads[r].sequence.size(), seedHits[i], st.localItems);
				st.cands.push_back(c);
			}
			st.round++;
			if (st.cands.empty())
			{
				// nothing left to extend: finish the seed walk for the assertion-class exit only
				for (; st.i < seedHits.size(); st.i++) { int rule = seedRule(st, seedHits[st.i], st.i); if (rule == 3) out[r].dropped = true; if (rule >= 2) break; }
				st.done = true;
			}
		}
		for (size_t r = 0; r < R; r++)
		{
			S1State& st = s1[r];
			if (st.done || st.cands.empty()) continue;
			st.itemBase = items.size();
			items.insert(items.end(), st.localItems.begin(), st.localItems.end());
			active.push_back(r);
		}
		if (active.empty()) break;
		stats.s1Rounds++;
		runExtend();
		#pragma omp parallel for schedule(dynamic, 4)
		for (size_t k = 0; k < active.size(); k++)
		{
			size_t r = active[k];
			GC_PROF_SCOPE(3, "s1.consume");
			S1State& st = s1[r];
			const std::vector<GcSeedHit>& seedHits = seedsOrdered[r];
			size_t next = 0; // next unconsumed speculative result
			for (; st.i < seedHits.size(); st.i++)
			{
				const GcSeedHit& seed = seedHits[st.i];
				int rule = seedRule(st, seed, st.i);
				if (rule == 3) { out[r].dropped = true; st.i = seedHits.size(); break; }
				if (rule == 2) { st.i = seedHits.size(); break; }
				if (rule == 1) continue;
				while (next < st.cands.size() && st.cands[next].seedIdx < st.i) next++;
				if (next >= st.cands.size() || st.cands[next].seedIdx != st.i) break; // not extended yet: first seed of the next round
				const S1Cand& c = st.cands[next++];
				st.seedsExtended += 1;
				GcPackedAln item;
				const gcgpu_ext_result* readResults = results.data() + st.itemBase;
				bool ok = buildAlignment(seed, c.ref, readResults, traces, item);
				for (int d = 0; d < 2; d++) if (c.ref.item[d] >= 0 && readResults[c.ref.item[d]].status == GCGPU_ITEM_INTERNAL) out[r].dropped = true;
				if (!ok || item.alignmentEnd == item.alignmentStart) continue;
				st.alns.emplace_back(item);
				st.alnsAdded.emplace_back(item);
				{ uint32_t n = item.size(); if (n == 0 || !(item.seqPosAt(n - 1) > item.seqPosAt(0))) st.degenerate = true; }
				std::sort(st.alns.begin(), st.alns.end(), [](const GcPackedAln& left, const GcPackedAln& right) { return left.alignmentStart < right.alignmentStart; });
				if (st.alns[0].alignmentStart == 0)
				{
					size_t minSeedGoodness = st.alns[0].seedGoodness;
					size_t contiguousEnd = st.alns[0].alignmentEnd;
					for (size_t i = 1; i < st.alns.size(); i++)
					{
						if (st.alns[i].alignmentStart <= contiguousEnd)
						{
							minSeedGoodness = std::min(minSeedGoodness, st.alns[i].seedGoodness);
							contiguousEnd = std::max(contiguousEnd, st.alns[i].alignmentEnd);
						}
					}
					if (contiguousEnd == reads[r].sequence.size()) st.seedScoreForEndToEndAln = minSeedGoodness;
				}
			}
			if (st.i >= seedHits.size()) st.done = true;
			stats_s1Wasted_add(st.cands.size() - next);
		}
	}
	phase("s1");
	// GreedyLength selection of the GA alignments + their path strings (Aligner.cpp:637-654)
	std::vector<std::vector<GcAlnItem>> longAlns(R);
	std::vector<std::string> longPathSeq(R);
	std::vector<size_t> longSeedsExtended(R, 0);
	#pragma omp parallel for schedule(dynamic, 4)
	for (size_t r = 0; r < R; r++)
	{
		GC_PROF_SCOPE(4, "s1.select+materialize+pathseq");
		longSeedsExtended[r] = s1[r].seedsExtended;
		if (out[r].dropped) { s1[r].alns.clear(); continue; }
		if (!s1[r].alns.empty())
		{
			std::vector<GcPackedAln> sel = gcpipe::selectGreedyLength(s1[r].alns);
			longAlns[r].resize(sel.size());
			for (size_t k = 0; k < sel.size(); k++) materialize(reads[r].sequence.data(), sel[k], longAlns[r][k]);
		}
		s1[r].alns.clear();
		if (!longAlns[r].empty()) longPathSeq[r] = gcpipe::traceToSequence(g, longAlns[r][0]);
	}

	// ---- S2: fragment anchoring (Aligner.cpp:656-730); all window seeds extended speculatively
	struct FragSeed { uint32_t seedIdx; ExtRef ref; };
	struct Frag { size_t l; size_t firstSeed, numSeeds; };
	std::vector<std::vector<GcSeedHit>> seedsByPos(R);
	std::vector<std::vector<Frag>> frags(R);
	std::vector<std::vector<FragSeed>> fragSeeds(R);
	items.clear();
	const size_t len = (size_t)params.colinearSplitLen, sep = (size_t)params.colinearSplitGap;
	// work items are generated per read in parallel (indices local to the read), then concatenated
	std::vector<std::vector<gcgpu_ext_item>> localItems(R);
	std::vector<size_t> itemBase(R + 1, 0);
	#pragma omp parallel for schedule(dynamic, 4)
	for (size_t r = 0; r < R; r++)
	{
		if (seedsOrdered[r].empty()) continue;
		GC_PROF_SCOPE(5, "s2.items");
		std::vector<GcSeedHit>& seeds = seedsByPos[r];
		seeds = seedsOrdered[r];
		std::sort(seeds.begin(), seeds.end(), [](const GcSeedHit& left, const GcSeedHit& right) { return left.seqPos < right.seqPos; });
		if (out[r].dropped) continue; // `cont` stays true after an assertion: every fragment is skipped (Aligner.cpp:700-703)
		const std::string& sequence = reads[r].sequence;
		size_t sl = 0, sr = 0;
		for (size_t l = 0; l + len <= sequence.length(); l += sep)
		{
			while (sr < seeds.size() && seeds[sr].seqPos + seeds[sr].matchLen <= l + len) sr++;
			while (sl < sr && seeds[sl].seqPos < l) sl++;
			if (sl >= sr) continue;
			Frag f; f.l = l; f.firstSeed = fragSeeds[r].size(); f.numSeeds = sr - sl;
			for (size_t i = sl; i < sr; i++)
			{
				GcSeedHit seed = seeds[i];
				seed.seqPos -= l;
				FragSeed fs; fs.seedIdx = (uint32_t)i;
				fs.ref = makeItems(b, r, sequence.size(), l, len, seed, localItems[r]);
				fragSeeds[r].push_back(fs);
			}
			frags[r].push_back(f);
		}
	}
	for (size_t r = 0; r < R; r++) itemBase[r + 1] = itemBase[r] + localItems[r].size();
	items.resize(itemBase[R]);
	#pragma omp parallel for schedule(dynamic, 16)
	for (size_t r = 0; r < R; r++)
	{
		if (!localItems[r].empty()) memcpy(items.data() + itemBase[r], localItems[r].data(), localItems[r].size() * sizeof(gcgpu_ext_item));
		std::vector<gcgpu_ext_item>().swap(localItems[r]);
	}
	extendCalls = 0; // S1's trace buffers are free again
	if (!items.empty()) runExtend();
	// in-order filter + anchors
	struct AnchorRec { std::vector<size_t> path; size_t x, y; size_t firstNode, firstOffset, lastNode, lastOffset; };
	std::vector<std::vector<AnchorRec>> anchors(R);
	std::vector<size_t> s2SeedsExtended(R, 0), lastFragExtended(R, 0);
	#pragma omp parallel for schedule(dynamic, 4)
	for (size_t r = 0; r < R; r++)
	{
		const std::string& sequence = reads[r].sequence;
		GC_PROF_SCOPE(6, "s2.filter+anchors");
		std::vector<GcPackedAln> kept;
		for (const Frag& f : frags[r])
		{
			kept.clear();
			size_t before = s2SeedsExtended[r];
			for (size_t k = 0; k < f.numSeeds; k++)
			{
				const FragSeed& fs = fragSeeds[r][f.firstSeed + k];
				GcSeedHit seed = seedsByPos[r][fs.seedIdx];
				seed.seqPos -= f.l;
				if (seed.seedClusterSize < params.seedClusterMinSize) continue;
				bool found = false, assertion = false;
				for (const auto& aln : kept) if (gcpipe::exactAlignmentPart(g, aln, seed, assertion)) { found = true; break; }
				if (assertion) { out[r].dropped = true; break; }
				if (found) continue;
				s2SeedsExtended[r] += 1;
				const gcgpu_ext_result* readResults = results.data() + itemBase[r];
				for (int d = 0; d < 2; d++) if (fs.ref.item[d] >= 0 && readResults[fs.ref.item[d]].status == GCGPU_ITEM_INTERNAL) out[r].dropped = true;
				GcPackedAln item;
				if (!buildAlignment(seed, fs.ref, readResults, traces, item)) continue;
				if (item.alignmentEnd == item.alignmentStart) continue;
				kept.emplace_back(item);
			}
			if (out[r].dropped) break;
			lastFragExtended[r] = s2SeedsExtended[r] - before;
			GC_PROF_SCOPE(7, "s2.anchors");
			for (const GcPackedAln& alignment : kept)
			{
				AnchorRec a; a.x = f.l; a.y = f.l + len - 1;
				uint32_t n = alignment.size();
				GcHostGraph::UnitigCache ucache;
				for (uint32_t k = 0; k < n; k++)
				{
					size_t node = gcpipe::packedSplitNode(g, alignment, k, ucache);
					if (a.path.empty() || node != a.path.back()) a.path.push_back(node);
				}
				int n0, n1; size_t o0, o1;
				gcpipe::packedNodePos(g, alignment, 0, n0, o0);
				gcpipe::packedNodePos(g, alignment, n - 1, n1, o1);
				a.firstNode = a.path[0]; a.firstOffset = o0 - g.nodeOffset[a.firstNode];
				a.lastNode = a.path.back(); a.lastOffset = o1 - g.nodeOffset[a.lastNode];
				anchors[r].push_back(std::move(a));
			}
		}
		if (out[r].dropped) anchors[r].clear();
	}

	phase("s2");
	// ---- S3: chaining (K2)
	std::vector<gcgpu_anchor> flatAnchors;
	std::vector<uint64_t> anchorOff(R + 1, 0);
	for (size_t r = 0; r < R; r++)
	{
		for (const AnchorRec& a : anchors[r])
		{
			gcgpu_anchor ga; ga.start_node = (uint32_t)a.path[0]; ga.end_node = (uint32_t)a.path.back(); ga.x = (int32_t)a.x; ga.y = (int32_t)a.y;
			flatAnchors.push_back(ga);
		}
		anchorOff[r + 1] = flatAnchors.size();
	}
	std::vector<uint32_t> chain(std::max<size_t>(1, flatAnchors.size())), chainLen(R);
	std::vector<int64_t> chainScore(R);
	{ double tDev = wallNow();
	check(gcgpu_chain(ctx, flatAnchors.data(), anchorOff.data(), (uint32_t)R, chain.data(), chainLen.data(), chainScore.data()), "gcgpu_chain");
	devMs += wallNow() - tDev; }
	stats.k2Reads += R; stats.k2Anchors += flatAnchors.size(); stats.k2Ms += gcgpu_last_kernel_ms(ctx);

	phase("s3");
	// ---- S4: chain -> node path (Aligner.cpp:738-831).  A candidate segment is kept as (node path, first offset,
	// last offset); its per-base position list (pathToTrace) is only materialised for the reads where the chain wins --
	// here only its LENGTH (the `longest` comparison, `<` => first wins ties) and its base string are needed.
	struct PathSeg { std::vector<size_t> path; size_t firstOffset = 0, lastOffset = 0, size = 0; };
	std::vector<PathSeg> longest(R);
	std::vector<std::string> pathSeq(R);
	// number of positions pathToTrace emits (same per-node tests by VALUE as Aligner.cpp:409-424)
	auto pathTraceSize = [&](const std::vector<size_t>& path, size_t firstNodeOffset, size_t lastNodeOffset)
	{
		size_t total = 0;
		for (size_t node : path)
		{
			size_t S = 0, L = g.nodeLength[node];
			if (node == path[0]) S = firstNodeOffset;
			else if (node == path.back()) L = lastNodeOffset + 1;
			if (L > S) total += L - S;
		}
		return total;
	};
	#pragma omp parallel
	{
		gcpipe::ChainPathScratch scratch;
		// the reference's std::unordered_set<size_t> nodes, as an epoch-stamped array that lives as long as the thread
		static thread_local std::vector<uint32_t> inPath;
		static thread_local uint32_t epoch = 0;
		if (inPath.size() != g.numNodes() || epoch > 0xF0000000u) { inPath.assign(g.numNodes(), 0); epoch = 0; }
		#pragma omp for schedule(dynamic, 4)
		for (size_t r = 0; r < R; r++)
		{
			const std::vector<AnchorRec>& A = anchors[r];
			GC_PROF_SCOPE(8, "s4.all");
			std::vector<size_t> pos_path;
			size_t firstNodeOffset = 0, lastNodeOffset = 0;
			epoch++;
			auto closeSegment = [&]()
			{
				size_t sz = pathTraceSize(pos_path, firstNodeOffset, lastNodeOffset);
				if (longest[r].size < sz) { longest[r].path = pos_path; longest[r].firstOffset = firstNodeOffset; longest[r].lastOffset = lastNodeOffset; longest[r].size = sz; }
			};
			for (uint32_t ci = 0; ci < chainLen[r]; ci++)
			{
				const AnchorRec& anchor = A[chain[anchorOff[r] + ci]];
				if (pos_path.empty())
				{
					pos_path = anchor.path;
					firstNodeOffset = anchor.firstOffset;
					lastNodeOffset = anchor.lastOffset;
					for (size_t j : pos_path) inPath[j] = epoch;
				}
				else
				{
					bool gap = anchor.path[0] == pos_path.back() && params.colinearGap != -1 && (long long)anchor.firstOffset - (long long)lastNodeOffset > params.colinearGap + 1;
					std::vector<size_t> path;
					if (inPath[anchor.path[0]] != epoch && pos_path.back() != anchor.firstNode)
					{
						long long gapLimit = params.colinearGap;
						if (gapLimit != -1) gapLimit -= (long long)anchor.firstOffset + (long long)((long long)g.nodeLength[pos_path.back()] - (long long)lastNodeOffset - 1);
						path = gcpipe::getChainPath(g, scratch, pos_path.back(), anchor.firstNode, gapLimit);
						if (path.empty()) gap = true;
					}
					if (gap)
					{
						closeSegment();
						epoch++; // nodes.clear()
						pos_path.clear();
						firstNodeOffset = anchor.firstOffset;
					}
					else
						for (size_t j : path) if (inPath[j] != epoch) { inPath[j] = epoch; pos_path.push_back(j); }
					for (size_t j : anchor.path) if (inPath[j] != epoch) { inPath[j] = epoch; pos_path.push_back(j); }
					lastNodeOffset = anchor.lastOffset;
				}
			}
			if (!pos_path.empty()) closeSegment();
			GC_PROF_SCOPE(9, "s4.pathSeq");
			const PathSeg& lg = longest[r];
			pathSeq[r].reserve(lg.size);
			for (size_t node : lg.path)
			{
				size_t S = 0, L = g.nodeLength[node];
				if (node == lg.path[0]) S = lg.firstOffset;
				else if (node == lg.path.back()) L = lg.lastOffset + 1;
				for (size_t o = S; o < L; o++) pathSeq[r].push_back(g.nodeChar((uint32_t)node, (uint32_t)o));
			}
		}
	}

	phase("s4");
	// ---- S1b + S5: NW distances (K3), then the edit path only where the chain wins (S6)
	// characters of the reads and path strings of the batch, in page-locked memory (the copy to the device runs at PCIe rate)
	char* nwBuf = nullptr; size_t nwBufSize = 0;
	std::vector<gcgpu_nw_item> nwItems;
	std::vector<int> gaItem(R, -1), clcItem(R, -1);
	std::vector<uint64_t> readOffInBuf(R);
	{
		// layout first (serial, a few integers per read), then the character copies in parallel
		uint64_t total = 0;
		for (size_t r = 0; r < R; r++)
		{
			bool needGa = !longAlns[r].empty();
			bool needClc = !seedsOrdered[r].empty() && !out[r].dropped;
			if (!needGa && !needClc) continue;
			readOffInBuf[r] = total;
			total += reads[r].sequence.size();
			// first cutoff of the NW passes: the whole-read alignment bounds its own distance from above
			// (score + unaligned read ends); the chained path normally lies at the same locus.  Only a
			// starting point -- the kernel doubles the cutoff until the pass succeeds, like edlib does from 64.
			int32_t kHint = 0, clcHint = (int32_t)(reads[r].sequence.size() / 5);
			if (needGa)
			{
				// the whole-read alignment IS a global alignment of (its padded path string, read): its score + the unaligned read
				// ends + the padding of the first and last node (< 2 x 64 graph characters, traceToPoses) bounds the distance from
				// above, so one pass at that cutoff always succeeds
				const GcAlnItem& a0 = longAlns[r][0];
				size_t ub = a0.alignmentScore + a0.alignmentStart + (reads[r].sequence.size() - a0.alignmentEnd);
				kHint = (int32_t)std::min<size_t>(ub + 130, (size_t)1 << 30);
				clcHint = (int32_t)std::min<size_t>(ub + ub * 3 / 10, reads[r].sequence.size() / 5);
			}
			if (needGa)
			{
				gcgpu_nw_item it; it.query_offset = total; it.target_offset = readOffInBuf[r]; it.query_len = (int32_t)longPathSeq[r].size(); it.target_len = (int32_t)reads[r].sequence.size();
				it.k_hint = kHint; it.want_path = 0;
				total += longPathSeq[r].size();
				gaItem[r] = (int)nwItems.size(); nwItems.push_back(it);
			}
			if (needClc)
			{
				gcgpu_nw_item it; it.query_offset = total; it.target_offset = readOffInBuf[r]; it.query_len = (int32_t)pathSeq[r].size(); it.target_len = (int32_t)reads[r].sequence.size();
				// the chained path usually costs 5-40 % more than the whole-read alignment; the first guess is capped at a fifth of
				// the read (the widest band the two-blocks-per-lane class holds for 10 kb reads) and is also what a read without a
				// whole-read alignment starts from (edlib's 64 would cost five doubling passes before the answer fits)
				it.k_hint = clcHint; it.want_path = 0;
				total += pathSeq[r].size();
				clcItem[r] = (int)nwItems.size(); nwItems.push_back(it);
			}
		}
		nwPinned.ensure(total + 16);
		nwBuf = (char*)nwPinned.p; nwBufSize = total;
		#pragma omp parallel for schedule(dynamic, 16)
		for (size_t r = 0; r < R; r++)
		{
			if (gaItem[r] < 0 && clcItem[r] < 0) continue;
			memcpy(&nwBuf[readOffInBuf[r]], reads[r].sequence.data(), reads[r].sequence.size());
			if (gaItem[r] >= 0) memcpy(&nwBuf[nwItems[gaItem[r]].query_offset], longPathSeq[r].data(), longPathSeq[r].size());
			if (clcItem[r] >= 0) memcpy(&nwBuf[nwItems[clcItem[r]].query_offset], pathSeq[r].data(), pathSeq[r].size());
		}
	}
	std::vector<gcgpu_nw_result> nwRes(nwItems.size());
	uint64_t opsUsed = 0;
	if (!nwItems.empty())
	{
		double tDev = wallNow();
		check(gcgpu_nw(ctx, nwBuf, nwBufSize, nwItems.data(), (uint32_t)nwItems.size(), nwRes.data(), nullptr, 0, &opsUsed), "gcgpu_nw");
		devMs += wallNow() - tDev;
		stats.k3Items += nwItems.size(); stats.k3Ms += gcgpu_last_kernel_ms(ctx);
		for (const auto& x : nwRes) stats.k3Blocks += x.blocks;
		if (traceOn)
		{
			size_t zero = 0, ok = 0, fail = 0; double sumHint = 0, sumD = 0;
			for (size_t i = 0; i < nwItems.size(); i++)
			{
				if (nwItems[i].k_hint <= 0) zero++;
				else if (nwRes[i].distance <= nwItems[i].k_hint) { ok++; sumHint += nwItems[i].k_hint; sumD += nwRes[i].distance; }
				else { fail++; if (fail <= 12) fprintf(stderr, "[gc]   hint %d distance %d\n", nwItems[i].k_hint, nwRes[i].distance); }
			}
			fprintf(stderr, "[gc] nw hints: none=%zu ok=%zu (mean hint %.0f, mean distance %.0f) too_small=%zu\n", zero, ok, ok ? sumHint / ok : 0.0, ok ? sumD / ok : 0.0, fail);
		}
	}
	// decision (Aligner.cpp:901-920): better = no GA alignment, or long_edit_distance > CLC score (strict)
	std::vector<gcgpu_nw_item> pathItems;
	std::vector<size_t> pathRead;
	for (size_t r = 0; r < R; r++)
	{
		GcReadResult& res = out[r];
		res.anchors = anchors[r].size();
		res.chained = chainLen[r];
		res.pathBp = longest[r].size;
		res.hasLong = gaItem[r] >= 0;
		if (gaItem[r] >= 0) res.longEditDistance = (size_t)nwRes[gaItem[r]].distance;
		if (clcItem[r] >= 0) res.clcScore = (size_t)nwRes[clcItem[r]].distance;
		res.seedsExtended = 0;
		bool haveClc = clcItem[r] >= 0 && longest[r].size != 0;
		bool better = haveClc && (longAlns[r].empty() || res.longEditDistance > res.clcScore);
		res.usedChain = better;
		if (better)
		{
			gcgpu_nw_item it = nwItems[clcItem[r]];
			it.want_path = 2; it.k_hint = nwRes[clcItem[r]].distance; // the distance is known: only the edit path is computed
			pathItems.push_back(it); pathRead.push_back(r);
		}
	}
	std::vector<gcgpu_nw_result> pathRes(pathItems.size());
	std::vector<uint8_t> ops;
	if (!pathItems.empty())
	{
		uint64_t cap = 0;
		for (const auto& it : pathItems) cap += (uint64_t)it.query_len + it.target_len + 8;
		ops.resize(cap);
		double tDev = wallNow();
		check(gcgpu_nw(ctx, nullptr, nwBufSize, pathItems.data(), (uint32_t)pathItems.size(), pathRes.data(), ops.data(), ops.size(), &opsUsed), "gcgpu_nw(path)");
		devMs += wallNow() - tDev;
		stats.k3Items += pathItems.size(); stats.k3Ms += gcgpu_last_kernel_ms(ctx);
		for (const auto& x : pathRes) stats.k3Blocks += x.blocks;
	}
	phase("nw");
	// ---- S5 trace conversion (Aligner.cpp:851-897) and final ordering (:1004)
	#pragma omp parallel for schedule(dynamic, 4)
	for (size_t k = 0; k < pathRead.size(); k++)
	{
		size_t r = pathRead[k];
		const std::string& sequence = reads[r].sequence;
		const std::vector<gcpipe::MatrixPos> lg = gcpipe::pathToTrace(g, longest[r].path, longest[r].firstOffset, longest[r].lastOffset);
		const uint8_t* op = ops.data() + pathRes[k].ops_offset;
		size_t n = pathRes[k].ops_len;
		GcAlnItem item;
		item.trace.resize(n);
		std::vector<size_t> splitNode(n);
		size_t pos_i = 0, seq_i = 0;
		for (size_t j = 0; j < n; j++)
		{
			GcTraceItem& t = item.trace[j];
			size_t node = lg[pos_i].node, off = lg[pos_i].nodeOffset;
			splitNode[j] = node;
			t.seqPos = (int64_t)seq_i;
			t.sequenceCharacter = seq_i < sequence.size() ? sequence[seq_i] : '-';
			t.graphCharacter = g.nodeChar((uint32_t)node, (uint32_t)off);
			t.node = g.nodeIDs[node];
			t.nodeOffset = (uint32_t)(off + g.nodeOffset[node]);
			t.nodeSwitch = false;
			unsigned char c = op[j];
			if (c == 0 || c == 3) { pos_i++; seq_i++; }
			else if (c == 1) pos_i++;
			else if (c == 2) seq_i++;
			seq_i = std::min(seq_i, sequence.length() - 1);
			pos_i = std::min(pos_i, lg.size() - 1);
		}
		// nodeSwitch compares the SPLIT nodes of consecutive entries (Aligner.cpp:880-884)
		for (size_t j = 0; j + 1 < n; j++) item.trace[j].nodeSwitch = splitNode[j] != splitNode[j + 1];
		if (n > 0)
		{
			item.traceScore = 0; // the reference leaves trace.score at 0 for the chained alignment
			item.alignmentScore = out[r].clcScore;
			item.alignmentStart = (size_t)item.trace[0].seqPos;
			item.alignmentEnd = (size_t)item.trace.back().seqPos + 1;
			out[r].alignments.clear();
			out[r].alignments.push_back(std::move(item));
		}
		else out[r].usedChain = false;
	}
	for (size_t r = 0; r < R; r++)
	{
		GcReadResult& res = out[r];
		if (!res.usedChain)
		{
			res.alignments = std::move(longAlns[r]);
			res.seedsExtended = res.alignments.empty() ? 0 : longSeedsExtended[r]; // stats.seedsExtended += alignments.seedsExtended (Aligner.cpp:995)
		}
		else res.seedsExtended = lastFragExtended[r]; // `alignments` still carries the last fragment's seedsExtended (Aligner.cpp:691,995)
		res.seedsExtended += s2SeedsExtended[r];
		std::sort(res.alignments.begin(), res.alignments.end(), [](const GcAlnItem& left, const GcAlnItem& right) { return left.alignmentStart < right.alignmentStart; });
	}
	phase("final");
#ifdef GC_PROF
	if (traceOn) for (int i = 0; i < 32; i++) if (g_profName[i]) fprintf(stderr, "[prof] %-32s %.2f ms\n", g_profName[i], g_prof[i] / 2.0e6);
#endif
	if (traceOn) fprintf(stderr, "[gc] batch reads=%zu s1_rounds=%llu s1_wasted=%llu\n", R, (unsigned long long)stats.s1Rounds, (unsigned long long)stats.s1Wasted);
}
