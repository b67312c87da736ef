// Host-side view of the alignment graph + MPC + minimizer index (flat arrays in
// reference numbering, see gc_index.h) with the accessors of the reference's
// AlignmentGraph that the per-read pipeline calls.
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>
#include "gc_common.cuh"
#include "gc_index.h"
#include "gc_k1.cuh"

// derived per-node records and out-edge queue keys of a graph view (GcNodeRec, GcGraphView::outKey)
inline void gcBuildNodeRecs(const GcGraphView& v, std::vector<GcNodeRec>& recs, std::vector<uint64_t>& outKey)
{
	recs.resize(v.numNodes);
	outKey.resize(v.outStart[v.numNodes]);
	for (uint32_t n = 0; n < v.numNodes; n++)
	{
		GcNodeRec& r = recs[n];
		r.seq0 = v.nodeSeq[2 * (size_t)n]; r.seq1 = v.nodeSeq[2 * (size_t)n + 1];
		r.inStart = v.inStart[n]; r.outStart = v.outStart[n];
		uint32_t ic = v.inStart[n + 1] - v.inStart[n], oc = v.outStart[n + 1] - v.outStart[n];
		r.firstIn = ic ? v.inNbr[v.inStart[n]] : 0;
		r.len = v.nodeLength[n]; r.linearizable = v.linearizable[n];
		r.inCount = (uint8_t)(ic < 255 ? ic : 255); r.outCount = (uint8_t)(oc < 255 ? oc : 255);
		for (uint32_t e = v.outStart[n]; e < v.outStart[n + 1]; e++) outKey[e] = ((uint64_t)v.componentNumber[v.outNbr[e]] << 32) | v.outNbr[e];
	}
}

struct GcHostGraph
{
	// split-node graph (AlignmentGraph.h:145-164)
	std::vector<uint8_t> nodeLength;
	std::vector<uint32_t> nodeOffset;
	std::vector<int32_t> nodeIDs;
	std::vector<uint8_t> reverse;
	std::vector<uint8_t> linearizable;
	std::vector<uint32_t> componentNumber;
	std::vector<uint32_t> chainNumber;
	std::vector<uint64_t> chainApproxPos;
	std::vector<uint64_t> nodeSeq;
	std::vector<uint32_t> inStart, inNbr, outStart, outNbr;
	// original (bigraph-doubled) nodes: id -> split nodes, in the reference's nodeLookup iteration order
	std::vector<int32_t> origIds;
	std::vector<uint32_t> origStart, origNodes, origSize;
	std::vector<std::string> origNames;
	std::vector<int32_t> origIndexOfId; // dense: digraph node id -> index into origIds (-1 if absent)
	// MPC index (AlignmentGraph.cpp:1430-1463, 1328-1391)
	std::vector<uint32_t> compMap, compIdx, compStart, compIds, topoIds, mpcWidth;
	std::vector<uint32_t> pathsStart, pathsK, backStart, backNode, backK;
	// minimizer index
	std::vector<uint64_t> mzKmers, mzPositions;
	std::vector<uint32_t> mzKmerStart;
	uint64_t mzLength = 15, mzWindow = 20, mzMaxCount = 0, mzBuckets = 1;
	uint64_t bpSize = 0;

	size_t numNodes() const { return nodeLength.size(); }

	template <typename T> static std::vector<T> vec(const GcIndexArray& a) { const T* p = (const T*)a.bytes.data(); return std::vector<T>(p, p + a.count); }

	void fromIndex(const GcIndexFile& f)
	{
		{
			auto nl = vec<uint32_t>(f.get("nodeLength"));
			nodeLength.assign(nl.begin(), nl.end());
		}
		nodeOffset = vec<uint32_t>(f.get("nodeOffset"));
		nodeIDs = vec<int32_t>(f.get("nodeIDs"));
		reverse = vec<uint8_t>(f.get("reverse"));
		linearizable = vec<uint8_t>(f.get("linearizable"));
		componentNumber = vec<uint32_t>(f.get("componentNumber"));
		chainNumber = vec<uint32_t>(f.get("chainNumber"));
		chainApproxPos = vec<uint64_t>(f.get("chainApproxPos"));
		nodeSeq = vec<uint64_t>(f.get("nodeSeq"));
		inStart = vec<uint32_t>(f.get("inStart")); inNbr = vec<uint32_t>(f.get("inNbr"));
		outStart = vec<uint32_t>(f.get("outStart")); outNbr = vec<uint32_t>(f.get("outNbr"));
		origIds = vec<int32_t>(f.get("origIds"));
		origStart = vec<uint32_t>(f.get("origStart"));
		origNodes = vec<uint32_t>(f.get("origNodes"));
		origSize = vec<uint32_t>(f.get("origSize"));
		{
			auto off = vec<uint32_t>(f.get("origNameOff"));
			auto names = vec<uint8_t>(f.get("origNames"));
			origNames.clear();
			for (size_t i = 0; i + 1 < off.size(); i++) origNames.emplace_back((const char*)names.data() + off[i], off[i + 1] - off[i]);
		}
		compMap = vec<uint32_t>(f.get("compMap")); compIdx = vec<uint32_t>(f.get("compIdx"));
		compStart = vec<uint32_t>(f.get("compStart")); compIds = vec<uint32_t>(f.get("compIds"));
		topoIds = vec<uint32_t>(f.get("topoIds")); mpcWidth = vec<uint32_t>(f.get("mpcWidth"));
		pathsStart = vec<uint32_t>(f.get("pathsStart")); pathsK = vec<uint32_t>(f.get("pathsK"));
		backStart = vec<uint32_t>(f.get("backStart")); backNode = vec<uint32_t>(f.get("backNode")); backK = vec<uint32_t>(f.get("backK"));
		mzKmers = vec<uint64_t>(f.get("mzKmers")); mzPositions = vec<uint64_t>(f.get("mzPositions")); mzKmerStart = vec<uint32_t>(f.get("mzKmerStart"));
		{
			auto p = vec<uint64_t>(f.get("mzParams"));
			mzLength = p[0]; mzWindow = p[1]; mzMaxCount = p[2]; mzBuckets = p[3];
		}
		bpSize = f.get("bpSize").u64()[0];
		finish();
	}

	// per split node, what seed clustering reads (GraphAligner.h:236-245), in one cache line fetch instead of two
	struct SeedAttr { uint64_t chainApproxPos; uint32_t chainNumber; uint32_t pad; };
	std::vector<SeedAttr> seedAttr;

	void finish()
	{
		seedAttr.resize(chainNumber.size());
		for (size_t i = 0; i < chainNumber.size(); i++) { seedAttr[i].chainApproxPos = chainApproxPos[i]; seedAttr[i].chainNumber = chainNumber[i]; seedAttr[i].pad = 0; }
		int32_t maxId = -1;
		for (auto id : origIds) if (id > maxId) maxId = id;
		origIndexOfId.assign((size_t)maxId + 1, -1);
		for (size_t i = 0; i < origIds.size(); i++) origIndexOfId[origIds[i]] = (int32_t)i;
	}

	GcGraphView view() const
	{
		GcGraphView v;
		v.numNodes = (uint32_t)nodeLength.size();
		v.nodeLength = nodeLength.data();
		v.nodeSeq = nodeSeq.data();
		v.inStart = inStart.data(); v.inNbr = inNbr.data();
		v.outStart = outStart.data(); v.outNbr = outNbr.data();
		v.componentNumber = componentNumber.data();
		v.linearizable = linearizable.data();
		v.coopLane = -1; v.coopWidth = 32; v.coopMask = 0xFFFFFFFFu; v.coopShift = 0;
		v.nodeRec = nullptr; v.outKey = nullptr; // gcBuildNodeRecs, by the caller that runs the lane-per-item form
		return v;
	}

	// AlignmentGraph::GetUnitigNode (AlignmentGraph.cpp:832-848): split node holding `offset` of digraph node `nodeId`.
	// The reference starts from a proportional guess and walks to the split node whose range contains the offset;
	// that node is unique, so any starting guess gives the same answer -- offset/64 is exact for the 64-bp splits.
	uint32_t unitigNode(int nodeId, size_t offset) const
	{
		int32_t oi = origIndexOfId[nodeId];
		const uint32_t* nodes = origNodes.data() + origStart[oi];
		size_t n = origStart[oi + 1] - origStart[oi];
		size_t index = offset >> 6;
		if (index >= n) index = n - 1;
		while (index < n - 1 && (nodeOffset[nodes[index]] + nodeLength[nodes[index]] <= offset)) index++;
		while (index > 0 && (nodeOffset[nodes[index]] > offset)) index--;
		return nodes[index];
	}
	// consecutive trace entries mostly stay inside one split node: remember its range
	struct UnitigCache { int nodeId = -1; uint32_t node = 0; size_t lo = 1, hi = 0; };
	uint32_t unitigNode(int nodeId, size_t offset, UnitigCache& c) const
	{
		if (nodeId == c.nodeId && offset >= c.lo && offset < c.hi) return c.node;
		uint32_t node = unitigNode(nodeId, offset);
		c.nodeId = nodeId; c.node = node; c.lo = nodeOffset[node]; c.hi = c.lo + nodeLength[node];
		return node;
	}
	// AlignmentGraph::GetReversePosition (AlignmentGraph.cpp:850-868)
	std::pair<int, size_t> reversePosition(int nodeId, size_t offset) const
	{
		size_t originalSize = origSize[origIndexOfId[nodeId]];
		size_t newOffset = originalSize - offset - 1;
		int reverseNodeId = (nodeId % 2 == 0) ? (nodeId / 2) * 2 + 1 : (nodeId / 2) * 2;
		return std::make_pair(reverseNodeId, newOffset);
	}
	char nodeChar(uint32_t node, uint32_t pos) const { return "ACGT"[(nodeSeq[2 * (size_t)node + (pos >> 5)] >> ((pos & 31) * 2)) & 3]; }
	const std::string& originalNodeName(int nodeId) const { return origNames[origIndexOfId[nodeId]]; }
};

// AlignmentCorrectnessEstimation.cpp:6-70 -- same expressions, same libm
inline GcViterbiTables gcMakeViterbiTables()
{
	const double correctMean = 0.1875, correctStddev = 0.0955, wrongMean = 0.5, wrondStddev = 0.0291;
	const int wordSize = 64;
	GcViterbiTables t;
	auto stddistlog = [](double val, double mean, double stddev) { return -(val - mean) * (val - mean) / (2 * stddev * stddev); };
	auto make = [&](double mean, double stddev, double* out)
	{
		std::vector<double> result;
		for (int i = 0; i <= wordSize / 2; i++) result.push_back(stddistlog(i, mean * wordSize, stddev * wordSize));
		double sum = 0;
		for (auto x : result) sum += exp(x);
		double add = log(1.0 / sum);
		for (auto& x : result) x += add;
		for (int i = wordSize / 2; i < wordSize; i++) result.push_back(result.back());
		// the reference indexes mismatches < size() (=96) else back(); entries 32.. are all equal, so 64 suffice
		for (int i = 0; i < 64; i++) out[i] = result[i];
	};
	make(correctMean, correctStddev, t.correctLogOdds);
	make(wrongMean, wrondStddev, t.wrongLogOdds);
	t.falseToCorrect = log(0.00001);
	t.falseToFalse = log(1.0 - 0.00001);
	t.correctToFalse = log(0.0000000001);
	t.correctToCorrect = log(1.0 - 0.0000000001);
	t.initialCorrect = log(0.8);
	t.initialFalse = log(0.2);
	return t;
}

// read character -> code of the resident sequence buffer: the IUPAC mask, plus bit 4 on characters that match
// in the DP but are not seeding bases (U/u: Common::ambiguousMatch maps them to T, MinimizerSeeder.cpp:24-43 does not)
inline uint8_t gcEncodeBase(char c);
inline uint8_t gcEncodeSeedBase(char c) { uint8_t m = gcEncodeBase(c); return (c == 'U' || c == 'u') ? (uint8_t)(m | 16) : m; }

// read characters -> IUPAC bit masks (bit0 A, bit1 C, bit2 G, bit3 T), Common::ambiguousMatch (GraphAlignerCommon.h:219-296)
inline uint8_t gcEncodeBase(char c)
{
	switch (c)
	{
		case 'A': case 'a': return 1;
		case 'C': case 'c': return 2;
		case 'G': case 'g': return 4;
		case 'T': case 't': case 'U': case 'u': return 8;
		case 'R': case 'r': return 1 | 4;
		case 'Y': case 'y': return 2 | 8;
		case 'K': case 'k': return 4 | 8;
		case 'M': case 'm': return 2 | 1;
		case 'S': case 's': return 2 | 4;
		case 'W': case 'w': return 1 | 8;
		case 'B': case 'b': return 2 | 4 | 8;
		case 'D': case 'd': return 1 | 4 | 8;
		case 'H': case 'h': return 1 | 2 | 8;
		case 'V': case 'v': return 1 | 2 | 4;
		case 'N': case 'n': return 15;
		default: return 0;
	}
}
