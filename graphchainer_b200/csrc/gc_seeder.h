// Host side of S0: minimizer seeding of a read and seed clustering.
//   MinimizerSeeder::getSeeds / iterateKmers / addMinimizers / matchToSeedHit
//       (src/MinimizerSeeder.cpp:60-102, 494-555)
//   GraphAligner::orderSeedsByChaining  (src/GraphAligner.h:233-295)
// The k-mer walk and the index probes (iterateKmers + addMinimizers) run on the device
// (gcgpu_seed, gc_seed.cuh); what stays here is sequential per read and full of libstdc++
// std::sort calls with partial keys whose permutation must match the reference's:
// the same std::sort calls on the same element order are used.  iterateKmers is kept as the
// restatement the device form is tested against (tests/hostsim/seed_ref.h).
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>
#include "gc_host_graph.h"
#ifndef GC_PROF_SCOPE
#define GC_PROF_SCOPE(id, name)
#endif

// src/GraphAlignerWrapper.h:11-37
struct GcSeedHit
{
	int nodeID;
	size_t nodeOffset;
	size_t seqPos;
	size_t matchLen;
	bool reverse;
	size_t alignmentGraphNodeId;
	size_t alignmentGraphNodeOffset;
	size_t rawSeedGoodness;
	size_t seedGoodness;
	size_t seedClusterSize;
	uint32_t orderedIdx = 0; // position after OrderSeeds (goodness order)
	uint32_t byPosIdx = 0;   // position after the split pass's sort by seqPos (Aligner.cpp:667) = index of the seed's cell on the device
};

namespace gcseed {

inline int charToInt(char c)
{
	switch (c) { case 'a': case 'A': return 0; case 'c': case 'C': return 1; case 'g': case 'G': return 2; case 't': case 'T': return 3; }
	return -1;
}

// iterateKmers (MinimizerSeeder.cpp:60-102): every k-mer of the read, re-emitted when it
// changed or the last emission is a whole window back
template <typename F>
void iterateKmers(const std::string& str, size_t kmerLength, size_t windowSize, F callback)
{
	const size_t realWindow = windowSize - kmerLength + 1;
	if (str.size() < kmerLength) return;
	const size_t mask = ~(0xFFFFFFFFFFFFFFFFull << (kmerLength * 2));
	size_t offset = 0;
	while (true)
	{
		while (offset < str.size() && charToInt(str[offset]) < 0) offset++;
		if (offset + kmerLength > str.size()) return;
		size_t kmer = 0;
		bool restart = false;
		for (size_t i = 0; i < kmerLength; i++)
		{
			int v = charToInt(str[offset + i]);
			if (v < 0) { offset += i; restart = true; break; }
			kmer <<= 2;
			kmer |= (size_t)v;
		}
		if (restart) continue;
		callback(offset + kmerLength - 1, kmer);
		size_t lastKmer = kmer;
		size_t lastPos = offset + kmerLength - 1;
		size_t i = kmerLength;
		for (; offset + i < str.size(); i++)
		{
			int v = charToInt(str[offset + i]);
			if (v < 0) { offset += i; restart = true; break; }
			kmer <<= 2;
			kmer &= mask;
			kmer |= (size_t)v;
			if (lastKmer != kmer || lastPos <= offset + i - realWindow)
			{
				callback(offset + i, kmer);
				lastKmer = kmer;
				lastPos = offset + i;
			}
		}
		if (!restart) return;
	}
}

// the second half of MinimizerSeeder::getSeeds (MinimizerSeeder.cpp:533-544) + matchToSeedHit (:546-555):
// sort the matches by count, apply the density cut, expand every kept k-mer into its positions.
// matchIndices = (k-mer END position, 0, first position index, count) in ascending position order.
inline std::vector<GcSeedHit> seedsFromMatches(const GcHostGraph& g, std::vector<std::tuple<size_t, size_t, size_t, size_t>>& matchIndices, size_t sequenceSize, double density)
{
	const size_t maxCount = g.mzMaxCount;
	std::vector<GcSeedHit> result;
	size_t maxHits = (size_t)(sequenceSize * density);
	if (density == -1) maxHits = std::numeric_limits<size_t>::max();
	std::sort(matchIndices.begin(), matchIndices.end(), [](const std::tuple<size_t, size_t, size_t, size_t>& left, const std::tuple<size_t, size_t, size_t, size_t>& right)
	{
		return std::get<3>(left) < std::get<3>(right);
	});
	size_t seedsHere = 0;
	size_t allowedCount = 0;
	for (auto match : matchIndices)
	{
		size_t start = std::get<2>(match);
		size_t end = start + std::get<3>(match);
		if (seedsHere >= maxHits && end - start > allowedCount) break;
		allowedCount = end - start;
		for (size_t i = start; i < end; i++)
		{
			size_t mergepos = g.mzPositions[i];
			size_t node = mergepos >> 6;
			size_t offset = mergepos & 63;
			// matchToSeedHit (:546-555)
			GcSeedHit s;
			s.nodeID = g.nodeIDs[node] / 2;
			s.nodeOffset = offset + g.nodeOffset[node];
			s.seqPos = std::get<0>(match);
			s.matchLen = g.mzLength;
			s.rawSeedGoodness = maxCount - (size_t)(int)std::get<3>(match);
			s.reverse = g.reverse[node] != 0;
			s.alignmentGraphNodeId = node;
			s.alignmentGraphNodeOffset = offset;
			s.seedGoodness = 0;
			s.seedClusterSize = 0;
			result.push_back(s);
		}
		seedsHere += end - start;
	}
	return result;
}


// GraphAligner::orderSeedsByChaining (GraphAligner.h:233-295)
inline void orderSeeds(const GcHostGraph& g, std::vector<GcSeedHit>& seedHits)
{
	std::unordered_map<size_t, std::vector<std::pair<size_t, size_t>>> seedPoses;
	for (size_t i = 0; i < seedHits.size(); i++)
	{
		size_t nodeIndex = seedHits[i].alignmentGraphNodeId;
		size_t realOffset = seedHits[i].alignmentGraphNodeOffset;
		seedPoses[g.chainNumber[nodeIndex]].emplace_back(i, g.chainApproxPos[nodeIndex] + realOffset - seedHits[i].seqPos);
	}
	for (auto& pair : seedPoses)
	{
		std::sort(pair.second.begin(), pair.second.end(), [](std::pair<size_t, size_t> left, std::pair<size_t, size_t> right) { return left.second < right.second; });
		size_t clusterStart = 0;
		for (size_t i = 1; i <= pair.second.size(); i++)
		{
			if (i < pair.second.size() && pair.second[i].second <= pair.second[i - 1].second + 100) continue;
			std::sort(pair.second.begin() + clusterStart, pair.second.begin() + i, [&seedHits](std::pair<size_t, size_t> left, std::pair<size_t, size_t> right) { return seedHits[left.first].seqPos < seedHits[right.first].seqPos; });
			size_t matchingBps = 0;
			int lastEnd = std::numeric_limits<int>::min();
			for (size_t j = clusterStart; j < i; j++)
			{
				int thisStart = (int)seedHits[pair.second[j].first].seqPos - (int)seedHits[pair.second[j].first].matchLen + 1;
				int thisEnd = (int)seedHits[pair.second[j].first].seqPos;
				matchingBps += (thisEnd - std::max(thisStart, lastEnd));
				lastEnd = thisEnd;
			}
			for (size_t j = clusterStart; j < i; j++)
			{
				seedHits[pair.second[j].first].seedGoodness = matchingBps + seedHits[pair.second[j].first].rawSeedGoodness;
				seedHits[pair.second[j].first].seedClusterSize = i - clusterStart;
			}
			clusterStart = i;
		}
	}
	std::sort(seedHits.begin(), seedHits.end(), [](const GcSeedHit& left, const GcSeedHit& right) { return left.seedGoodness < right.seedGoodness; });
	std::reverse(seedHits.begin(), seedHits.end());
}

}
