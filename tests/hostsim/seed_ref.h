// TEST-ONLY restatement of MinimizerSeeder::getSeeds (src/MinimizerSeeder.cpp:522-544) with the
// reference's own control flow: sequential iterateKmers walk (:60-102) and one index probe per
// emitted k-mer (addMinimizers, :494-520; the BBHash + kmerCheck lookup has "exact k-mer or nothing"
// semantics, here an std::unordered_map).  The product does these lookups on the device
// (gcgpu_seed); tests/test_seed.py compares the two on reads with N/U characters and homopolymers.
#pragma once
#include <tuple>
#include <unordered_map>
#include "../../graphchainer_b200/csrc/gc_seeder.h"

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
}

struct SeedRefIndex
{
	std::unordered_map<uint64_t, uint32_t> kmerIndex; // a repeated key keeps the LAST index
	explicit SeedRefIndex(const GcHostGraph& g) { kmerIndex.reserve(g.mzKmers.size() * 2); for (size_t i = 0; i < g.mzKmers.size(); i++) kmerIndex[g.mzKmers[i]] = (uint32_t)i; }
};

inline std::vector<std::tuple<size_t, size_t, size_t, size_t>> seedRefMatches(const GcHostGraph& g, const SeedRefIndex& idx, const std::string& sequence)
{
	std::vector<std::tuple<size_t, size_t, size_t, size_t>> matchIndices;
	const size_t maxCount = g.mzMaxCount;
	gcseed::iterateKmers(sequence, g.mzLength, g.mzWindow, [&](size_t pos, size_t kmer)
	{
		auto found = idx.kmerIndex.find(kmer);
		if (found == idx.kmerIndex.end()) return;
		size_t index = found->second;
		size_t start = g.mzKmerStart[index];
		size_t end = g.mzKmerStart[index + 1];
		size_t count = end - start;
		if (count >= maxCount) return;
		matchIndices.emplace_back(pos, (size_t)0, start, count);
	});
	return matchIndices;
}
