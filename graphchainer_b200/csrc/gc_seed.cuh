// S0 -- minimizer seeding of a read: which k-mers does the reference look up, and what does the
// index answer.
//
// Reference: MinimizerSeeder::getSeeds -> iterateKmers -> addMinimizers
// (src/MinimizerSeeder.cpp:522-544, 60-102, 494-520).  iterateKmers walks the read with a rolling
// 2-bit k-mer, restarting after every non-ACGT character, and calls back for the first k-mer of a
// segment and then whenever `lastKmer != kmer || lastPos <= pos - realWindow` (:95), realWindow =
// windowSize - k + 1.  addMinimizers keeps a k-mer iff it is in the index with fewer than maxCount
// positions and records (pos, first position index, count).
//
// Per-position form (B200: one thread per read position).  Inside a segment lastKmer always equals
// the k-mer of the previous position (it is either emitted there, or not emitted because equal), so
//     emitted(i)  <=>  first of segment  ||  kmer(i) != kmer(i-1)  ||  (i - p0) % realWindow == 0
// where p0 = first position of the run of equal k-mers containing i.  kmer(i) == kmer(i-1) iff the
// k+1 characters ending at i are one base b, so p0 = (start of that homopolymer run) + k - 1.
// The position needs nothing from its neighbours' results: every emitted k-mer is probed
// independently in an open-addressing table (16-byte slots: key, first position index, count).
#pragma once
#include "gc_common.cuh"

struct GcMzSlot
{
	uint64_t key;
	uint32_t start;
	uint32_t count; // 0xFFFFFFFF = empty slot
};

struct GcMzView
{
	const GcMzSlot* slots;
	uint64_t mask;       // capacity - 1 (capacity is a power of two)
	uint32_t k;          // k-mer length (<= 31)
	uint32_t realWindow; // windowSize - k + 1
	uint64_t maxCount;
};

GC_HD uint64_t gc_mz_hash(uint64_t k)
{
	k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
	return k;
}

// read code (IUPAC mask, bit 4 = "the character was U/u": matches like T in the DP but is not a
// seeding base, MinimizerSeeder.cpp:24-43) -> 0..3, or -1
GC_HD int gc_seed_base(uint8_t code)
{
	code &= 0xDF; // bit 5 (not an upper-case A C G T: only K3's byte comparison cares) does not stop seeding, charToInt takes both cases
	return code == 1 ? 0 : code == 2 ? 1 : code == 4 ? 2 : code == 8 ? 3 : -1;
}

// Is the k-mer ENDING at position i looked up by the reference and kept by addMinimizers?
GC_HD bool gc_seed_position(const GcMzView& mz, const uint8_t* codes, int32_t len, int32_t i, uint32_t& start, uint32_t& count)
{
	const int32_t k = (int32_t)mz.k;
	if (i < k - 1 || i >= len) return false;
	uint64_t kmer = 0;
	for (int32_t j = i - k + 1; j <= i; j++)
	{
		int b = gc_seed_base(codes[j]);
		if (b < 0) return false;
		kmer = (kmer << 2) | (uint64_t)b;
	}
	int prev = (i - k >= 0) ? gc_seed_base(codes[i - k]) : -1;
	if (prev >= 0)
	{
		uint64_t prevKmer = (kmer >> 2) | ((uint64_t)prev << (2 * (k - 1)));
		if (prevKmer == kmer)
		{
			// inside a run of equal k-mers: re-emitted every realWindow positions from the run's first k-mer
			int b = gc_seed_base(codes[i]);
			int32_t j = i - k - 1;
			while (j >= 0 && gc_seed_base(codes[j]) == b) j--;
			int32_t p0 = (j + 1) + k - 1;
			if ((uint32_t)(i - p0) % mz.realWindow != 0) return false;
		}
	}
	for (uint64_t h = gc_mz_hash(kmer) & mz.mask; ; h = (h + 1) & mz.mask)
	{
		GcMzSlot s = mz.slots[h];
		if (s.count == 0xFFFFFFFFu) return false;
		if (s.key == kmer)
		{
			if ((uint64_t)s.count >= mz.maxCount) return false;
			start = s.start; count = s.count;
			return true;
		}
	}
}
