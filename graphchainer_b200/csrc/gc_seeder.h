// Host side of S0: from the index answers of a read (gcgpu_seed: which k-mers the reference looks up and what the
// minimizer index holds for them) to its seeds in the two orders the pipeline walks them.
//   MinimizerSeeder::getSeeds, second half + matchToSeedHit   (src/MinimizerSeeder.cpp:533-555)
//   GraphAligner::orderSeedsByChaining                        (src/GraphAligner.h:233-295)
//   the split pass's sort by position                         (src/Aligner.cpp:667)
//
// What of the reference is observable here, and what is not:
//   * Three std::sort calls decide the ORDER of seeds with equal keys (matches of equal count, seeds of equal goodness,
//     seeds at the same read position), and that order reaches the output (which seed is extended first).  libstdc++'s
//     introsort is driven by comparison results and the element count only, so sorting 8-byte (key, index) proxies with the
//     same comparison on the same initial order yields the reference's permutation without moving 90-byte seed records.
//   * Everything the clustering computes (seedGoodness, seedClusterSize) is a function of the cluster a seed falls in and
//     of the SET of read positions in it -- not of the order in which equal diagonals or equal positions are visited -- so
//     it is computed here on flat arrays: one sort of (chain, diagonal) keys, a scan for the gaps > 100, and per cluster
//     the sorted positions (the reference: a hash map of per-chain vectors, two sorts of index pairs per cluster).
//   * The per-hit graph lookups (position list, chain number, approximate chain position) are cache misses into arrays of
//     tens to hundreds of MB: they are issued as prefetches for all hits of the read before the first one is consumed.
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <string>
#include <vector>
#include "../../include/gcgpu.h"
#include "gc_host_graph.h"

// the fields of the reference's SeedHit (src/GraphAlignerWrapper.h:11-37) that the per-read pipeline reads
struct GcSeedHit
{
	uint32_t seqPos;                    // k-mer END position in the read
	uint32_t alignmentGraphNodeId;      // split node
	uint32_t seedGoodness;
	uint32_t seedClusterSize;
	uint32_t orderedIdx = 0;            // position after OrderSeeds (goodness order)
	uint32_t byPosIdx = 0;              // position after the split pass's sort by seqPos = index of the seed's cell on the device
	uint8_t alignmentGraphNodeOffset;
	uint8_t matchLen;
};

namespace gcseed {

// scratch of one host thread, reused from read to read
struct Scratch
{
	struct Key32 { uint32_t key; uint32_t idx; };
	struct DiagKey { uint64_t diag; uint32_t chain; uint32_t idx; };
	std::vector<Key32> byCount, byGoodness, byPos;
	std::vector<DiagKey> diag;
	std::vector<uint32_t> raw, positions;     // rawSeedGoodness per hit; read positions of one cluster
	std::vector<GcSeedHit> hits;              // expansion order
};

// Seeds of one read.  matches[0..n) = the device's answers in ascending read position.  ordered = the reference's seed
// vector after OrderSeeds (goodness order), byPos = the same seeds after the split pass's sort by position; the two carry
// each other's indices.
inline void seedRead(const GcHostGraph& g, const gcgpu_seed_match* matches, size_t n, size_t sequenceSize, double density, Scratch& s, std::vector<GcSeedHit>& ordered, std::vector<GcSeedHit>& byPos)
{
	ordered.clear(); byPos.clear();
	// ---- matches by count (MinimizerSeeder.cpp:533-536); ties keep introsort's order from ascending position
	s.byCount.resize(n);
	for (size_t i = 0; i < n; i++) { s.byCount[i].key = matches[i].count; s.byCount[i].idx = (uint32_t)i; }
	std::sort(s.byCount.begin(), s.byCount.end(), [](const Scratch::Key32& left, const Scratch::Key32& right) { return left.key < right.key; });
	// ---- density cut (:537-544): stop at the first match with more positions than the last accepted one once the budget is spent
	size_t maxHits = density == -1 ? std::numeric_limits<size_t>::max() : (size_t)(sequenceSize * density);
	size_t kept = 0, numHits = 0, allowedCount = 0;
	for (; kept < n; kept++)
	{
		size_t count = s.byCount[kept].key;
		if (numHits >= maxHits && count > allowedCount) break;
		allowedCount = count;
		numHits += count;
		__builtin_prefetch(&g.mzPositions[matches[s.byCount[kept].idx].start]);
	}
	if (numHits == 0) return;
	// ---- expansion (matchToSeedHit, :546-555): one seed per indexed position of every kept k-mer, in that order
	s.hits.resize(numHits); s.raw.resize(numHits);
	const uint32_t maxCount = (uint32_t)g.mzMaxCount;
	{
		size_t h = 0;
		for (size_t k = 0; k < kept; k++)
		{
			const gcgpu_seed_match& m = matches[s.byCount[k].idx];
			for (uint32_t i = m.start; i < m.start + m.count; i++, h++)
			{
				uint64_t mergepos = g.mzPositions[i];
				GcSeedHit& hit = s.hits[h];
				hit.seqPos = m.pos;
				hit.alignmentGraphNodeId = (uint32_t)(mergepos >> 6);
				hit.alignmentGraphNodeOffset = (uint8_t)(mergepos & 63);
				hit.matchLen = (uint8_t)g.mzLength;
				s.raw[h] = maxCount - m.count; // rawSeedGoodness
				__builtin_prefetch(&g.seedAttr[hit.alignmentGraphNodeId]);
			}
		}
	}
	// ---- clusters (GraphAligner.h:236-283): per chain, runs of diagonals no more than 100 apart
	s.diag.resize(numHits);
	for (size_t h = 0; h < numHits; h++)
	{
		const GcHostGraph::SeedAttr& a = g.seedAttr[s.hits[h].alignmentGraphNodeId];
		s.diag[h].chain = a.chainNumber;
		s.diag[h].diag = a.chainApproxPos + s.hits[h].alignmentGraphNodeOffset - s.hits[h].seqPos; // size_t arithmetic, as the reference's
		s.diag[h].idx = (uint32_t)h;
	}
	std::sort(s.diag.begin(), s.diag.end(), [](const Scratch::DiagKey& left, const Scratch::DiagKey& right) { return left.chain != right.chain ? left.chain < right.chain : left.diag < right.diag; });
	const int matchLen = (int)g.mzLength;
	for (size_t first = 0; first < numHits; )
	{
		size_t last = first + 1;
		while (last < numHits && s.diag[last].chain == s.diag[first].chain && s.diag[last].diag <= s.diag[last - 1].diag + 100) last++;
		// bases of the read covered by the cluster's k-mers (:265-272)
		s.positions.resize(last - first);
		for (size_t k = first; k < last; k++) s.positions[k - first] = s.hits[s.diag[k].idx].seqPos;
		std::sort(s.positions.begin(), s.positions.end());
		size_t matchingBps = 0;
		int lastEnd = std::numeric_limits<int>::min();
		for (uint32_t p : s.positions)
		{
			int thisStart = (int)p - matchLen + 1, thisEnd = (int)p;
			matchingBps += (size_t)(thisEnd - std::max(thisStart, lastEnd));
			lastEnd = thisEnd;
		}
		for (size_t k = first; k < last; k++)
		{
			GcSeedHit& hit = s.hits[s.diag[k].idx];
			hit.seedGoodness = (uint32_t)(matchingBps + s.raw[s.diag[k].idx]);
			hit.seedClusterSize = (uint32_t)(last - first);
		}
		first = last;
	}
	// ---- goodness order (:284-285: ascending sort, then reversed)
	s.byGoodness.resize(numHits);
	for (size_t h = 0; h < numHits; h++) { s.byGoodness[h].key = s.hits[h].seedGoodness; s.byGoodness[h].idx = (uint32_t)h; }
	std::sort(s.byGoodness.begin(), s.byGoodness.end(), [](const Scratch::Key32& left, const Scratch::Key32& right) { return left.key < right.key; });
	ordered.resize(numHits);
	for (size_t h = 0; h < numHits; h++) { ordered[h] = s.hits[s.byGoodness[numHits - 1 - h].idx]; ordered[h].orderedIdx = (uint32_t)h; }
	// ---- position order of the split pass (Aligner.cpp:667): std::sort of the ordered vector by seqPos
	s.byPos.resize(numHits);
	for (size_t h = 0; h < numHits; h++) { s.byPos[h].key = ordered[h].seqPos; s.byPos[h].idx = (uint32_t)h; }
	std::sort(s.byPos.begin(), s.byPos.end(), [](const Scratch::Key32& left, const Scratch::Key32& right) { return left.key < right.key; });
	byPos.resize(numHits);
	for (size_t h = 0; h < numHits; h++)
	{
		ordered[s.byPos[h].idx].byPosIdx = (uint32_t)h;
		byPos[h] = ordered[s.byPos[h].idx];
	}
}

}
