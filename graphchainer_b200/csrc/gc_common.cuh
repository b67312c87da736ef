// Shared definitions for the libgcgpu kernels (graph view, bit-parallel column type).
//
// All per-work-item algorithms in this directory are written as GC_HD functions:
// nvcc compiles them as device code for the kernels in gcgpu.cu; tests/host_sim.cpp
// compiles the same functions with g++ to check the logic on the CPU-only build
// box.  The shipped library never calls the host instantiation.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GC_HD __host__ __device__ __forceinline__
#define GC_HD_NOINLINE __host__ __device__ __noinline__
#else
#define GC_HD inline
#define GC_HD_NOINLINE inline __attribute__((noinline))
#endif

#define GC_INT_MAX 2147483647

// Warp-wide votes of the lane-per-item kernels (gc_k1s.cuh): every lane of a warp owns one work item and the lanes
// meet at loop heads whose conditions are warp-uniform.  The CPU instantiation (tests/hostsim) is a warp of one lane.
#if defined(__CUDA_ARCH__)
#define GC_WARP_ANY(p) (__any_sync(0xFFFFFFFFu, (p)) != 0)
#define GC_WARP_MAX(v) __reduce_max_sync(0xFFFFFFFFu, (uint32_t)(v))
#else
#define GC_WARP_ANY(p) (p)
#define GC_WARP_MAX(v) ((uint32_t)(v))
#endif

// status codes of a work item
enum GcStatus : int32_t
{
	GC_OK = 0,
	GC_FAILED = 1,          // the reference's OnewayTrace::TraceFailed()
	GC_OVERFLOW_ITEMS = 2,  // workspace too small: host retries with a bigger slab
	GC_OVERFLOW_HEAP = 3,
	GC_OVERFLOW_TRACE = 4,
	GC_INTERNAL = 5         // a condition the reference guards with assert() was hit
};

// Split-node alignment graph, flat arrays in reference node numbering
// (AlignmentGraph.h:145-172).  Sequences: 2 bits/base, A0 C1 G2 T3, 32 bases per
// u64 chunk, LSB first (AlignmentGraph.cpp:114-142).
// Everything a node visit of the lane-per-item K1 kernels reads about the node itself, as ONE 32-byte record (one sector):
// the nine scattered loads of the flat arrays were two dependent memory round trips per visit.  Derived from the arrays
// below at gcgpu_create (gcBuildNodeRecs); counts saturate at 255 (then the CSR arrays are consulted).
struct __attribute__((aligned(32))) GcNodeRec
{
	uint64_t seq0, seq1;
	uint32_t inStart, outStart;
	uint32_t firstIn;      // inNbr[inStart] (0 if the node has no in-neighbour)
	uint8_t len, linearizable, inCount, outCount;
};

struct GcGraphView
{
	uint32_t numNodes;
	const uint8_t* nodeLength;        // [N] 1..64
	const uint64_t* nodeSeq;          // [2N]
	const uint32_t* inStart;          // [N+1] CSR, reference insertion order
	const uint32_t* inNbr;
	const uint32_t* outStart;         // [N+1]
	const uint32_t* outNbr;
	const uint32_t* componentNumber;  // [N] topological rank (unique per node on a DAG)
	const uint8_t* linearizable;      // [N]
	const GcNodeRec* nodeRec;         // [N] packed per-node record (lane-per-item kernels)
	const uint64_t* outKey;           // [out edges] componentNumber[outNbr[e]] << 32 | outNbr[e]: the queue key of the out-neighbour, one load instead of two dependent ones
	// >= 0: a GROUP of coopWidth (32, 16 or 8) adjacent lanes of a warp executes ONE work item in lockstep (same control
	// flow, same values) and this is the lane's index inside its group; helpers may then split loop-free lookups across the
	// group.  coopMask = the group's lanes inside the warp, coopShift = its first lane.  Groups of one warp run different
	// items and diverge freely (independent thread scheduling); they share the warp's registers-per-lane, which is the point:
	// an item's state is per lane, so a warp of four groups keeps four items resident for the registers of one.
	int32_t coopLane;
	int32_t coopWidth;
	uint32_t coopMask;
	uint32_t coopShift;
};

GC_HD int gc_popc(uint64_t x)
{
#if defined(__CUDA_ARCH__)
	return __popcll(x);
#else
	return __builtin_popcountll(x);
#endif
}

GC_HD int gc_node_base(const GcGraphView& g, uint32_t node, uint32_t pos)
{
	return (int)((g.nodeSeq[2 * (uint64_t)node + (pos >> 5)] >> ((pos & 31) * 2)) & 3);
}

// One DP column over 64 read rows in Myers' difference encoding: VP/VN bit r = the
// vertical delta into row r is +1/-1; scoreEnd = value at row 63.
// (reference: WordSlice.h:150-166)
struct GcWord
{
	uint64_t VP;
	uint64_t VN;
	int32_t scoreEnd;
};

// WordSlice::getScoreBeforeStart (WordSlice.h:244-247)
GC_HD int32_t gc_sbs(const GcWord& w) { return w.scoreEnd - gc_popc(w.VP) + gc_popc(w.VN); }

// WordSlice::getValue (WordSlice.h:177-186)
GC_HD int32_t gc_value(const GcWord& w, int row)
{
	uint64_t mask = 0;
	if (row < 63) mask = ~0ULL << (row + 1);
	return w.scoreEnd + gc_popc(w.VN & mask) - gc_popc(w.VP & mask);
}

// One Myers column step with horizontal input (GraphAlignerBitvectorCommon.h:243-263)
GC_HD GcWord gc_next_column(uint64_t Eq, GcWord s, uint64_t hinP, uint64_t hinN, uint64_t& houtP, uint64_t& houtN)
{
	uint64_t Xv = Eq | s.VN;
	Eq |= hinN;
	uint64_t Xh = (((Eq & s.VP) + s.VP) ^ s.VP) | Eq;
	uint64_t Ph = s.VN | ~(Xh | s.VP);
	uint64_t Mh = s.VP & Xh;
	uint64_t tempMh = (Mh << 1) | hinN;
	houtN = Mh >> 63;
	uint64_t tempPh = (Ph << 1) | hinP;
	s.VP = tempMh | ~(Xv | tempPh);
	houtP = Ph >> 63;
	s.VN = tempPh & Xv;
	s.scoreEnd -= (int32_t)houtN;
	s.scoreEnd += (int32_t)houtP;
	return s;
}

// WordSlice::differenceMasksBitTwiddle (WordSlice.h:555-653): bit r of `leftSmaller`
// iff left's value at row r < right's, given scoreDifference = right.sbs - left.sbs.
GC_HD void gc_difference_masks(uint64_t leftVP, uint64_t leftVN, uint64_t rightVP, uint64_t rightVN, int scoreDifference, uint64_t& leftSmaller, uint64_t& rightSmaller)
{
	leftSmaller = 0;
	rightSmaller = 0;
	uint64_t VPcommon = ~(leftVP & rightVP);
	uint64_t VNcommon = ~(leftVN & rightVN);
	leftVP &= VPcommon;
	leftVN &= VNcommon;
	rightVP &= VPcommon;
	rightVN &= VNcommon;
	uint64_t twosmaller = leftVN & rightVP;
	uint64_t onesmaller = (rightVP & ~leftVN) | (leftVN & ~rightVP);
	uint64_t onebigger = (leftVP & ~rightVN) | (rightVN & ~leftVP);
	uint64_t twobigger = rightVN & leftVP;
	onebigger |= twobigger;
	onesmaller |= twosmaller;
	if (scoreDifference > 0)
	{
		for (int i = 1; i < scoreDifference; i++)
		{
			uint64_t ls = onebigger & ~(onebigger - 1);
			onebigger ^= (~twobigger & ls);
			twobigger &= ~ls;
			if (onebigger == 0) { leftSmaller = ~0ULL; rightSmaller = 0; return; }
		}
		uint64_t ls = onebigger & ~(onebigger - 1);
		leftSmaller |= ls - 1;
		onebigger ^= (~twobigger & ls);
		twobigger &= ~ls;
	}
	else if (scoreDifference < 0)
	{
		for (int i = 1; i < -scoreDifference; i++)
		{
			uint64_t ls = onesmaller & ~(onesmaller - 1);
			onesmaller ^= (~twosmaller & ls);
			twosmaller &= ~ls;
			if (onesmaller == 0) { leftSmaller = 0; rightSmaller = ~0ULL; return; }
		}
		uint64_t ls = onesmaller & ~(onesmaller - 1);
		rightSmaller |= ls - 1;
		onesmaller ^= (~twosmaller & ls);
		twosmaller &= ~ls;
	}
	for (int i = 0; i < 64; i++)
	{
		if (onesmaller == 0)
		{
			if (onebigger == 0) break;
			uint64_t ls = onebigger & ~(onebigger - 1);
			rightSmaller |= (uint64_t)(-(int64_t)ls);
			break;
		}
		if (onebigger == 0)
		{
			uint64_t ls = onesmaller & ~(onesmaller - 1);
			leftSmaller |= (uint64_t)(-(int64_t)ls);
			break;
		}
		uint64_t lsBigger = onebigger & ~(onebigger - 1);
		uint64_t lsSmaller = onesmaller & ~(onesmaller - 1);
		if (lsBigger > lsSmaller) leftSmaller |= lsBigger - lsSmaller;
		else rightSmaller |= lsSmaller - lsBigger;
		onebigger ^= (~twobigger & lsBigger);
		twobigger &= ~lsBigger;
		onesmaller ^= (~twosmaller & lsSmaller);
		twosmaller &= ~lsSmaller;
	}
}

// WordSlice::mergeTwoSlices (WordSlice.h:491-530): element-wise minimum of two columns
GC_HD GcWord gc_merge(GcWord left, GcWord right)
{
	int32_t lsbs = gc_sbs(left), rsbs = gc_sbs(right);
	if (lsbs > rsbs) { GcWord t = left; left = right; right = t; int32_t ti = lsbs; lsbs = rsbs; rsbs = ti; }
	uint64_t leftSmaller, rightSmaller;
	gc_difference_masks(left.VP, left.VN, right.VP, right.VN, rsbs - lsbs, leftSmaller, rightSmaller);
	uint64_t mask = (rightSmaller | ((leftSmaller | rightSmaller) - (rightSmaller << 1))) & ~leftSmaller;
	uint64_t leftReduction = leftSmaller & (rightSmaller << 1);
	uint64_t rightReduction = rightSmaller & (leftSmaller << 1);
	if ((rightSmaller & 1) && lsbs < rsbs) rightReduction |= 1;
	left.VN &= ~leftReduction;
	right.VN &= ~rightReduction;
	GcWord result;
	result.VN = (left.VN & ~mask) | (right.VN & mask);
	result.VP = (left.VP & ~mask) | (right.VP & mask);
	result.scoreEnd = left.scoreEnd < right.scoreEnd ? left.scoreEnd : right.scoreEnd;
	return result;
}

// WordSlice::changedMinScoreLocalMinima (WordSlice.h:427-461): minimum over the rows
// (and the cell before row 0) where `w` is strictly below `old`; INT_MAX if none.
GC_HD int32_t gc_changed_min_score(const GcWord& w, const GcWord& old)
{
	int32_t scoreBeforeStart = gc_sbs(w);
	int32_t otherScoreBeforeStart = gc_sbs(old);
	uint64_t VP = w.VP, VN = w.VN;
	uint64_t possibleLocalMinima = (VP & (VN - VP));
	possibleLocalMinima >>= 1;
	possibleLocalMinima |= 0x8000000000000000ULL & (VN | ~(VN - VP)) & ~VP;
	if (w.scoreEnd + gc_popc(VN) >= old.scoreEnd - gc_popc(old.VP))
	{
		uint64_t smaller, dummy;
		gc_difference_masks(VP, VN, old.VP, old.VN, otherScoreBeforeStart - scoreBeforeStart, smaller, dummy);
		if (smaller != ~0ULL)
		{
			possibleLocalMinima |= (~smaller) >> 1;
			possibleLocalMinima |= (~smaller) << 1;
			possibleLocalMinima |= 1;
			possibleLocalMinima |= 0x8000000000000000ULL;
			possibleLocalMinima &= smaller;
		}
	}
	int32_t result = (scoreBeforeStart < otherScoreBeforeStart) ? scoreBeforeStart : GC_INT_MAX;
	while (possibleLocalMinima != 0)
	{
		uint64_t currentMinimumMask = possibleLocalMinima ^ (possibleLocalMinima - 1);
		int32_t scoreHere = scoreBeforeStart + gc_popc(VP & currentMinimumMask) - gc_popc(VN & currentMinimumMask);
		if (scoreHere < result) result = scoreHere;
		possibleLocalMinima &= ~currentMinimumMask;
	}
	return result;
}

// Eq masks of one 64-row slice of the read (GraphAlignerBitvectorCommon.h:280-319).
// `seq` holds one IUPAC bit mask per read base: bit0 A, bit1 C, bit2 G, bit3 T
// (so 'N' = 15 matches everything, as Common::characterMatch does).
GC_HD void gc_eq_vector(const uint8_t* seq, int32_t seqLen, int32_t j, uint64_t eq[4], int32_t coopLane = -1, int32_t coopWidth = 32, uint32_t coopMask = 0xFFFFFFFFu, uint32_t coopShift = 0)
{
#if defined(__CUDA_ARCH__)
	if (coopLane >= 0)
	{
		// lane l of a group of W looks at rows l, W+l, 2W+l ...; one ballot per base and W-row stripe
		const uint32_t laneBits = coopWidth >= 32 ? 0xFFFFFFFFu : ((1u << coopWidth) - 1u);
		eq[0] = eq[1] = eq[2] = eq[3] = 0;
		for (int32_t r = 0; r < 64; r += coopWidth)
		{
			uint32_t c = (j + r + coopLane < seqLen) ? seq[j + r + coopLane] : 0u;
			for (int b = 0; b < 4; b++)
			{
				uint32_t m = (__ballot_sync(coopMask, (c >> b) & 1u) >> coopShift) & laneBits;
				eq[b] |= (uint64_t)m << r;
			}
		}
		return;
	}
#endif
	eq[0] = eq[1] = eq[2] = eq[3] = 0;
	for (int i = 0; i < 64 && j + i < seqLen; i++)
	{
		uint64_t c = seq[j + i];
		uint64_t bit = 1ULL << i;
		eq[0] |= (c & 1) ? bit : 0;
		eq[1] |= (c & 2) ? bit : 0;
		eq[2] |= (c & 4) ? bit : 0;
		eq[3] |= (c & 8) ? bit : 0;
	}
}

// Common::characterMatch(read char, graph base) on the encoded forms
GC_HD bool gc_char_match(uint8_t seqMask, int graphBase) { return (seqMask >> graphBase) & 1; }
