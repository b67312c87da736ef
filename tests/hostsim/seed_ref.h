// TEST-ONLY restatement of MinimizerSeeder::getSeeds (src/MinimizerSeeder.cpp:522-544) with the
// reference's own control flow: sequential iterateKmers walk (:60-102) and one index probe per
// emitted k-mer (addMinimizers, :494-520; the BBHash + kmerCheck lookup has "exact k-mer or nothing"
// semantics, here an std::unordered_map).  The product does these lookups on the device
// (gcgpu_seed); tests/test_seed.py compares the two on reads with N/U characters and homopolymers.
#pragma once
#include <unordered_map>
#include "../../graphchainer_b200/csrc/gc_seeder.h"

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
