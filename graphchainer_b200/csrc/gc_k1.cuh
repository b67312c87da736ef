// K1 -- banded bit-parallel graph extension from one seed cell, one direction.
//
// One work item = the reference's
//   GraphAlignerBitvectorBanded::getReverseTraceFromSeed(sequence, node, offset)
// (GraphAlignerBitvectorBanded.h:46-71) with the default parameters of the colinear
// pipeline (bandwidth 10, ramp 0, no precise clipping, no forced global, low-memory
// slices, component-ordered queue -- AlignerMain.cpp:145-209):
//   forward pass   getViterbiSlices/fillDPSlice/calculateSlice   (Banded.h:206-701)
//   node sweep     calculateNodeInner / getNextSlice             (BVCommon.h:243-263,885-1168)
//   last slice     flattenLastSliceEnd + phmap slot order        (BVCommon.h:1171-1229)
//   Viterbi cut    NextState / removeWronglyAlignedEnd           (AlignmentCorrectnessEstimation.cpp:105-129, BVCommon.h:1231-1241)
//   backtrace      getReverseTraceFromTable + pickBacktrace*     (BVCommon.h:392-804)
//
// Design (B200): one thread owns one work item.  Everything a work item touches
// lives in its own slab of HBM (slice metadata, 64-byte node items, a small
// binary heap) so no inter-thread communication is needed; parallelism comes from
// the millions of independent (fragment, seed, direction) items of a read batch.
// The reference's hash maps / priority queue with per-target lists become:
//   * a slice = node items stored in processing order, which on a DAG is strictly
//     ascending componentNumber => membership tests are binary searches;
//   * the queue = a binary min-heap of (componentNumber,node) keys, duplicates
//     dropped at pop; a node's incoming columns are re-derived from its
//     in-neighbours' stored end columns + a `pushed` flag (merge is an element-wise
//     minimum, so the order of merging does not matter).
#pragma once
#include "gc_common.cuh"

// one node of one 64-row slice (reference: NodeSliceMapItemStruct, NodeSlice.h:15-47)
struct __attribute__((aligned(16))) GcNodeItem
{
	uint64_t startVP, startVN;
	uint64_t endVP, endVN;
	uint64_t HP, HN;
	int32_t startScore, endScore;
	int32_t minScore;
	uint32_t nodeAndFlag; // bit31 = end column was pushed to the out-neighbours
};

struct GcSliceMeta
{
	double correctLogOdds, falseLogOdds;
	int32_t minScore;
	uint32_t minScoreNode;
	uint32_t minScoreNodeOffset;
	int32_t bandwidth;
	uint32_t firstItem;
	uint32_t numItems;
	uint8_t correctFromCorrect, falseFromCorrect;
	uint8_t pad[6];
};

// constants of AlignmentCorrectnessEstimation.cpp:39-70, computed on the host with
// libm (exactly as the reference's static initialisers do) and uploaded
struct GcViterbiTables
{
	double correctLogOdds[64];
	double wrongLogOdds[64];
	double falseToCorrect, falseToFalse, correctToFalse, correctToCorrect;
	double initialCorrect, initialFalse; // log(0.8), log(0.2)
};

// Columns of the node a backtrace is walking: only the bit vectors are stored (16 B per column); the value at row 63 of
// column h follows from the node's horizontal deltas along that row, which the node item holds anyway:
// scoreEnd(h) = scoreEnd(0) + sum over columns 1..h of (HP - HN)  (getNextSlice updates scoreEnd by exactly hout, BVCommon.h:258-259)
struct __attribute__((aligned(16))) GcColVV { uint64_t VP, VN; };
struct GcCols { const GcColVV* c; uint64_t HP, HN; int32_t score0; };
GC_HD GcWord gc_cols_get(const GcCols& cv, uint32_t h)
{
	const uint64_t m = (2ULL << h) - 2; // bits 1..h
	GcWord w; w.VP = cv.c[h].VP; w.VN = cv.c[h].VN;
	w.scoreEnd = cv.score0 + gc_popc(cv.HP & m) - gc_popc(cv.HN & m);
	return w;
}

// per work item memory, carved from one slab by the host (see gcgpu.cu)
struct GcK1Workspace
{
	GcSliceMeta* slices;   // [numSlices + 1]
	GcNodeItem* items;     // [itemCap]
	uint64_t* heap;        // [heapCap]
	GcColVV* cols;         // [64] columns of the node being recomputed (backtrace)
	uint32_t itemCap;
	uint32_t heapCap;
	// optional (lock-step kernels): node|flag of the first 32 items of the slice being filled and of the previous slice, in
	// shared memory -- the slice lookups are the most frequent dependent loads of a node visit (9 % of a lone warp's stall samples)
	uint32_t* keysA = nullptr;
	uint32_t* keysB = nullptr;
};

struct GcK1Result
{
	int32_t status;      // GcStatus
	int32_t score;
	uint32_t traceLen;   // entries written to the trace buffer (reverse order, as the reference builds it)
	uint32_t itemsUsed;
	uint64_t columns;    // work counter: Myers column steps (cellsProcessed semantics, BVCommon.h:1162)
};

// trace entry: node(32) | nodeOffset(6) << 32 | nodeSwitch(1) << 38 | (seqPos+1)(25) << 39
GC_HD uint64_t gc_pack_trace(uint32_t node, uint32_t offset, int32_t seqPos, bool nodeSwitch)
{
	return (uint64_t)node | ((uint64_t)offset << 32) | ((uint64_t)(nodeSwitch ? 1 : 0) << 38) | ((uint64_t)(uint32_t)(seqPos + 1) << 39);
}

GC_HD GcWord gc_item_start(const GcNodeItem& it) { GcWord w; w.VP = it.startVP; w.VN = it.startVN; w.scoreEnd = it.startScore; return w; }
GC_HD GcWord gc_item_end(const GcNodeItem& it) { GcWord w; w.VP = it.endVP; w.VN = it.endVN; w.scoreEnd = it.endScore; return w; }
GC_HD uint32_t gc_item_node(const GcNodeItem& it) { return it.nodeAndFlag & 0x7FFFFFFFu; }

// binary search of `node` in a slice (items sorted by componentNumber)
GC_HD const GcNodeItem* gc_find_item(const GcGraphView& g, const GcNodeItem* items, uint32_t n, uint32_t node, const uint32_t* keys = nullptr)
{
#if defined(__CUDA_ARCH__)
	if (g.coopLane >= 0 && n <= (uint32_t)g.coopWidth)
	{
		// one probe: lane k of the group holds the node of item k (from the shared-memory copy of the keys if there is one)
		uint32_t mine = ((uint32_t)g.coopLane < n) ? (keys ? (keys[g.coopLane] & 0x7FFFFFFFu) : gc_item_node(items[g.coopLane])) : 0xFFFFFFFFu;
		uint32_t m = __ballot_sync(g.coopMask, mine == node) >> g.coopShift;
		return m ? &items[__ffs(m) - 1] : nullptr;
	}
#endif
	uint32_t key = g.componentNumber[node];
	uint32_t lo = 0, hi = n;
	while (lo < hi)
	{
		uint32_t mid = (lo + hi) >> 1;
		uint32_t k = g.componentNumber[gc_item_node(items[mid])];
		if (k < key) lo = mid + 1; else hi = mid;
	}
	if (lo < n && gc_item_node(items[lo]) == node) return &items[lo];
	return nullptr;
}

// ---- heap of (componentNumber << 32 | node)
GC_HD bool gc_heap_push(uint64_t* heap, uint32_t& size, uint32_t cap, uint64_t key)
{
	if (size >= cap) return false;
	uint32_t i = size++;
	while (i > 0)
	{
		uint32_t p = (i - 1) >> 1;
		if (heap[p] <= key) break;
		heap[i] = heap[p];
		i = p;
	}
	heap[i] = key;
	return true;
}
GC_HD uint64_t gc_heap_pop(uint64_t* heap, uint32_t& size)
{
	uint64_t top = heap[0];
	uint64_t last = heap[--size];
	uint32_t i = 0;
	while (true)
	{
		uint32_t c = 2 * i + 1;
		if (c >= size) break;
		if (c + 1 < size && heap[c + 1] < heap[c]) c++;
		if (heap[c] >= last) break;
		heap[i] = heap[c];
		i = c;
	}
	if (size > 0) heap[i] = last;
	return top;
}

// state of the column loop of one node (registers)
struct GcColumnRun
{
	GcWord ws;
	uint64_t eq[4];          // Eq masks of the slice, already combined with the first-row rule
	uint64_t prevHP, prevHN; // horizontal deltas along the last row of the previous slice (bit = column)
	uint64_t HP, HN;         // this node's horizontal deltas along row 63, shifted in from the top
	uint64_t chunk0, chunk1; // node sequence, 2 bits per base
	int32_t minScore; uint32_t minOffset;
	uint64_t flatMask;       // FLAT runs: rows above the last row of the read (flattenWordSlice subtracts their deltas)
};
// flattenWordSlice (BVCommon.h:265-273): the column's value at the last row of the read
GC_HD int32_t gc_flat_score(const GcWord& w, uint64_t flatMask) { return w.scoreEnd - gc_popc(w.VP & flatMask) + gc_popc(w.VN & flatMask); }
// columns [begin, end) of the node: one getNextSlice step each (BVCommon.h:1118-1161)
// FORCE: 1 = every column of the range has its first row forced, 0 = none, 2 = columns up to forceUntil (decided per column:
// the thread-per-item kernel keeps ONE loop so that the lanes of a warp, whose forced ranges differ, stay in one loop)
// FLAT: the minimum tracked is that of the FLATTENED column values (last slice of a read that does not fill 64 rows); used
// by flattenLastSliceEnd, which only needs that minimum -- the columns themselves are not stored
// eq[base] as selects: a dynamically indexed array would live in local memory, and its load sits at the head of the column step's dependency chain
GC_HD uint64_t gc_sel4(const uint64_t eq[4], int base)
{
	uint64_t lo = (base & 1) ? eq[1] : eq[0];
	uint64_t hi = (base & 1) ? eq[3] : eq[2];
	return (base & 2) ? hi : lo;
}
// one column of the run: getNextSlice + the bookkeeping of the column loop of calculateNodeInner (BVCommon.h:1118-1161)
// FLAT: the minimum tracked is that of the FLATTENED column values (last slice of a read that does not fill 64 rows); used
// by flattenLastSliceEnd, which only needs that minimum -- the columns themselves are not stored
template <bool FLAT>
GC_HD void gc_col_step(GcColumnRun& r, uint32_t pos, int base, bool forced)
{
	uint64_t hP, hN;
	r.ws = gc_next_column(gc_sel4(r.eq, base), r.ws, (r.prevHP >> pos) & 1, (r.prevHN >> pos) & 1, hP, hN);
	if (forced)
	{
		r.ws.VP &= ~1ULL;
		r.ws.VN |= 1;
	}
	const int32_t tracked = FLAT ? gc_flat_score(r.ws, r.flatMask) : r.ws.scoreEnd;
	if (tracked < r.minScore)
	{
		r.minScore = tracked;
		r.minOffset = pos;
	}
	r.HP = (r.HP >> 1) | (hP << 63);
	r.HN = (r.HN >> 1) | (hN << 63);
}
// bases of columns pos.. of a node as a rolling word: two bits per column, refilled every 16 columns
GC_HD uint32_t gc_col_bases(const GcColumnRun& r, uint32_t pos) { return (uint32_t)((pos < 32 ? r.chunk0 : r.chunk1) >> ((pos & 31) * 2)); }

// columns [begin, end) of the node
// FORCE: 1 = every column of the range has its first row forced, 0 = none, 2 = columns up to forceUntil (decided per column:
// the thread-per-item kernel keeps ONE loop so that the lanes of a warp, whose forced ranges differ, stay in one loop)
template <int FORCE, bool FLAT = false>
GC_HD void gc_columns_range(GcColumnRun& r, uint32_t begin, uint32_t end, GcColVV* cols, uint32_t forceUntil = 0)
{
	// 16 columns at a time: their bases are one 32-bit word that is shifted down two bits per column (the
	// per-column "which chunk, which bit pair" arithmetic of the obvious form was a fifth of the loop's instructions)
	uint32_t pos = begin;
	while (pos < end)
	{
		uint32_t segEnd = (pos | 15u) + 1;
		if (segEnd > end) segEnd = end;
		uint32_t bases = gc_col_bases(r, pos);
		for (; pos < segEnd; pos++)
		{
			gc_col_step<FLAT>(r, pos, (int)(bases & 3), FORCE == 1 || (FORCE == 2 && forceUntil >= pos));
			bases >>= 2;
			if (cols) { cols[pos].VP = r.ws.VP; cols[pos].VN = r.ws.VN; }
		}
	}
}

// Start of the column run of a node from its start column (calculateNodeInner, BVCommon.h:1060-1117): the forceUntil
// fix-up of the previous slice's horizontal deltas, the first-row rule, the run state.  Returns forceUntil: columns
// 1..forceUntil have their first row forced (a column cannot start below the previous slice's row), the rest not.
GC_HD uint32_t gc_cols_prepare_seq(uint64_t chunk0, uint64_t chunk1, uint32_t len, const uint64_t eq[4], const GcWord& ws, bool prevExists, int32_t prevStartScore, uint64_t prevHP, uint64_t prevHN,
	uint64_t flatMask, GcColumnRun& run)
{
	uint32_t forceUntil = 0;
	if (prevExists)
	{
		int32_t scoreBefore = gc_sbs(ws);
		int32_t scoreComparison = prevStartScore;
		if (scoreBefore < scoreComparison)
		{
			for (uint32_t fixoffset = 1; fixoffset < 64; fixoffset++)
			{
				int32_t newScoreComparison = scoreComparison;
				newScoreComparison += (int32_t)((prevHP >> fixoffset) & 1);
				newScoreComparison -= (int32_t)((prevHN >> fixoffset) & 1);
				uint64_t mask = 1ULL << fixoffset;
				if (scoreBefore < newScoreComparison)
				{
					prevHP |= mask;
					prevHN &= ~mask;
					forceUntil = fixoffset;
				}
				if (scoreBefore == newScoreComparison)
				{
					prevHP &= ~mask;
					prevHN &= ~mask;
				}
				scoreBefore++;
				scoreComparison = newScoreComparison;
				if (scoreBefore >= scoreComparison) break;
			}
		}
	}
	else
	{
		forceUntil = len;
	}
	uint64_t forceEq = ~0ULL;
	if (!prevExists) forceEq ^= 1;
	run.ws = ws; run.minScore = flatMask ? gc_flat_score(ws, flatMask) : ws.scoreEnd; run.minOffset = 0; run.HP = 0; run.HN = 0;
	run.flatMask = flatMask;
	run.prevHP = prevHP; run.prevHN = prevHN;
	run.eq[0] = eq[0] & forceEq; run.eq[1] = eq[1] & forceEq; run.eq[2] = eq[2] & forceEq; run.eq[3] = eq[3] & forceEq;
	run.chunk0 = chunk0; run.chunk1 = chunk1;
	return forceUntil;
}
GC_HD uint32_t gc_cols_prepare(const GcGraphView& g, uint32_t node, uint32_t len, const uint64_t eq[4], const GcWord& ws, bool prevExists, int32_t prevStartScore, uint64_t prevHP, uint64_t prevHN,
	uint64_t flatMask, GcColumnRun& run)
{
	return gc_cols_prepare_seq(g.nodeSeq[2 * (uint64_t)node], g.nodeSeq[2 * (uint64_t)node + 1], len, eq, ws, prevExists, prevStartScore, prevHP, prevHN, flatMask, run);
}
// the horizontal bits were shifted in from the top, one per column: bring the bit of column p to bit p
GC_HD void gc_cols_finish(GcColumnRun& run, uint32_t len) { if (len > 1) { run.HP >>= (64 - len); run.HN >>= (64 - len); } }

// Columns 1..len-1 of a node from its start column (the tail of calculateNodeInner,
// BVCommon.h:1060-1167).  If `cols` is non-null every column is also stored there
// (recalcNodeWordslice, BVCommon.h:828-852).
GC_HD void gc_node_columns(const GcGraphView& g, uint32_t node, const uint64_t eq[4], GcWord ws, bool prevExists, int32_t prevStartScore, uint64_t prevHP, uint64_t prevHN,
	GcWord& endOut, uint64_t& HPout, uint64_t& HNout, int32_t& minScore, uint32_t& minOffset, GcColVV* cols, uint64_t flatMask = 0)
{
	uint32_t len = g.nodeLength[node];
	if (cols) { cols[0].VP = ws.VP; cols[0].VN = ws.VN; }
	GcColumnRun run;
	uint32_t forceUntil = gc_cols_prepare(g, node, len, eq, ws, prevExists, prevStartScore, prevHP, prevHN, flatMask, run);
	if (flatMask) gc_columns_range<2, true>(run, 1, len, nullptr, forceUntil);
	else if (g.coopLane >= 0)
	{
		uint32_t forcedEnd = forceUntil + 1 < len ? forceUntil + 1 : len;
		if (forcedEnd > 1) gc_columns_range<1>(run, 1, forcedEnd, cols);
		if (forcedEnd < len) gc_columns_range<0>(run, forcedEnd < 1 ? 1 : forcedEnd, len, cols);
	}
	else gc_columns_range<2>(run, 1, len, cols, forceUntil);
	gc_cols_finish(run, len);
	endOut = run.ws;
	HPout = run.HP;
	HNout = run.HN;
	minScore = run.minScore;
	minOffset = run.minOffset;
}

// recalcNodeWordslice (BVCommon.h:828-852): all columns of a stored node
// flatMask != 0: nothing is stored; *flatMin / *flatOffset receive the minimum of the flattened column values and its first column
GC_HD void gc_recalc_node(const GcGraphView& g, const GcNodeItem& item, const uint64_t eq[4], const GcNodeItem* prev, GcColVV* cols, uint64_t flatMask = 0, int32_t* flatMin = nullptr, uint32_t* flatOffset = nullptr, int32_t* score0 = nullptr)
{
	GcWord ws = gc_item_start(item);
	bool prevExists = prev != nullptr;
	int32_t prevStart = prevExists ? prev->startScore : 0;
	if (prevExists && gc_sbs(ws) > prevStart)
	{
		GcWord src; src.VP = ~0ULL; src.VN = 0; src.scoreEnd = prevStart + 64;
		ws = gc_merge(ws, src);
	}
	if (score0) *score0 = ws.scoreEnd;
	GcWord endOut; uint64_t hp, hn; int32_t ms; uint32_t mo;
	gc_node_columns(g, gc_item_node(item), eq, ws, prevExists, prevStart, prevExists ? prev->HP : ~0ULL, prevExists ? prev->HN : 0ULL, endOut, hp, hn, ms, mo, cols, flatMask);
	if (flatMin) { *flatMin = ms; *flatOffset = mo; }
}

// phmap::flat_hash_map<size_t,...> slot assignment (phmap.h:487-517,1869-1896,
// 2008-2020, phmap_utils.h:68-79): returns in `order[]` the indices 0..n-1 of the keys
// in the iteration order of a table created with reserve(reserveN) and filled by
// inserting keys[0..n-1] in that order.  `slots` is scratch of >= 2*max(n,reserveN)+2 entries.
GC_HD uint64_t gc_phmap_hash(uint64_t key)
{
#if defined(__CUDA_ARCH__)
	const uint64_t k = 0xde5fb9d2630458e9ULL;
	return __umul64hi(key, k) + key * k;
#else
	unsigned __int128 p = (unsigned __int128)key * 0xde5fb9d2630458e9ULL;
	return (uint64_t)(p >> 64) + (uint64_t)p;
#endif
}
GC_HD uint32_t gc_phmap_find_slot(const uint32_t* slots, uint32_t capacity, uint64_t hash)
{
	// probe_seq<16>: groups of 16 control bytes starting at H1 & capacity, triangular steps.
	// Control index `capacity` is the sentinel; indices above it mirror the first 15 slots,
	// bytes beyond the mirror are always empty (they can only be reached when the table has
	// a free real slot earlier in the same group, see tests/test_phmap_order).
	uint64_t offset = (hash >> 7) & capacity;
	uint64_t index = 0;
	while (true)
	{
		for (uint32_t i = 0; i < 16; i++)
		{
			uint64_t ctrl = offset + i;
			if (ctrl == capacity) continue; // sentinel
			uint64_t slot = ctrl > capacity ? ctrl - capacity - 1 : ctrl;
			if (slot >= capacity) return (uint32_t)((offset + i) & capacity);
			if (slots[slot] == 0xFFFFFFFFu) return (uint32_t)((offset + i) & capacity);
		}
		index += 16;
		offset = (offset + index) & capacity;
	}
}
GC_HD uint32_t gc_phmap_normalize_capacity(uint64_t n)
{
	if (n == 0) return 1;
	uint32_t c = 1;
	while (c < n) c = c * 2 + 1;
	return c;
}
// keys = key(0..n-1), inserted in that order.  Returns false if scratch is too small.
template <typename KeyFn>
GC_HD bool gc_phmap_order_keys(const KeyFn& key, uint32_t n, uint32_t reserveN, uint32_t* slots, uint32_t scratchCap, uint32_t* capacityOut)
{
	// reserve(n): rehash(GrowthToLowerboundCapacity(n)) -> capacity NormalizeCapacity(n + (n-1)/7)
	uint32_t capacity = 0;
	if (reserveN > 0)
	{
		uint64_t lower = (uint64_t)reserveN + (uint64_t)(((int64_t)reserveN - 1) / 7);
		capacity = gc_phmap_normalize_capacity(lower);
	}
	uint32_t size = 0;
	uint32_t growthLeft = capacity - capacity / 8;
	if (capacity > scratchCap) return false;
	for (uint32_t s = 0; s < capacity; s++) slots[s] = 0xFFFFFFFFu;
	for (uint32_t i = 0; i < n; i++)
	{
		if (growthLeft == 0)
		{
			// rehash_and_grow_if_necessary: no deletions ever happen here, so it always grows
			uint32_t newCap = capacity == 0 ? 1 : capacity * 2 + 1;
			if ((uint64_t)newCap + capacity > scratchCap) return false;
			// re-insert in old slot order into a fresh table placed after the old one
			uint32_t* fresh = slots + capacity;
			for (uint32_t s = 0; s < newCap; s++) fresh[s] = 0xFFFFFFFFu;
			for (uint32_t s = 0; s < capacity; s++)
			{
				if (slots[s] == 0xFFFFFFFFu) continue;
				uint32_t t = gc_phmap_find_slot(fresh, newCap, gc_phmap_hash(key(slots[s])));
				fresh[t] = slots[s];
			}
			for (uint32_t s = 0; s < newCap; s++) slots[s] = fresh[s];
			capacity = newCap;
			growthLeft = (capacity - capacity / 8) - size;
		}
		uint32_t t = gc_phmap_find_slot(slots, capacity, gc_phmap_hash(key(i)));
		slots[t] = i;
		size++;
		growthLeft--;
	}
	*capacityOut = capacity;
	return true;
}
struct GcItemNodeKey
{
	const GcNodeItem* items;
	GC_HD uint64_t operator()(uint32_t i) const { return (uint64_t)gc_item_node(items[i]); }
};
// keys = node ids of items[0..n-1]
GC_HD bool gc_phmap_order(const GcNodeItem* items, uint32_t n, uint32_t reserveN, uint32_t* slots, uint32_t scratchCap, uint32_t* capacityOut)
{
	GcItemNodeKey key; key.items = items;
	return gc_phmap_order_keys(key, n, reserveN, slots, scratchCap, capacityOut);
}

// AlignmentCorrectnessEstimationState::NextState (AlignmentCorrectnessEstimation.cpp:105-129)
GC_HD void gc_viterbi_next(const GcViterbiTables& t, const GcSliceMeta& prev, int32_t mismatches, GcSliceMeta& out)
{
	double cc = prev.correctLogOdds + t.correctToCorrect;
	double fc = prev.falseLogOdds + t.falseToCorrect;
	double cf = prev.correctLogOdds + t.correctToFalse;
	double ff = prev.falseLogOdds + t.falseToFalse;
	out.correctFromCorrect = cc >= fc;
	out.falseFromCorrect = cf >= ff;
	double newCorrect = (cc < fc) ? fc : cc; // std::max(a,b) returns b only if a < b
	double newFalse = (cf < ff) ? ff : cf;
	int idx = mismatches < 64 ? mismatches : 63;
	newCorrect += t.correctLogOdds[idx];
	newFalse += t.wrongLogOdds[idx];
	out.correctLogOdds = newCorrect;
	out.falseLogOdds = newFalse;
}

struct GcK1Params
{
	int32_t bandwidth; // params.initialBandwidth (10)
};

// ------------------------------------------------------------------------------------
// Forward pass.  Returns the number of slices kept (index of the last slice in ws.slices),
// after removeWronglyAlignedEnd; <= 0 means the extension failed.
//
// Shape of the loop (B200): one thread owns one work item and the 32 items of a warp run in SIMT.
// The reference's nest "for slice { while queue { node: merge incoming; <=63 column steps; push } }"
// is flattened into ONE loop whose body is
//     stage 1  finish the previous node / close and open slices / pop the next node and merge its
//              incoming columns -- repeated while the popped node is a single column (no column steps)
//     stage 2  the column steps of that node (gc_node_columns)
// so that every lane of a warp meets the others at the single column loop once per node, no matter
// which slice each of them is in.  (In the nested form the lanes wait for each other at every slice
// end and the one-column SNP nodes occupy a whole column-loop slot: measured 6 of 32 lanes active.)
GC_HD int32_t gc_k1_forward(const GcGraphView& g, const GcViterbiTables& vt, const GcK1Params& prm, const uint8_t* seq, int32_t seqLen, uint32_t startNode, uint32_t startOffset,
	GcK1Workspace& ws, GcK1Result& res)
{
	res.status = GC_OK;
	res.columns = 0;
	uint32_t itemsUsed = 0;
	int32_t numSlices = (seqLen + 63) / 64;
	// ---- getInitialSliceExactPosition (BVCommon.h:1243-1279)
	{
		if (ws.itemCap < 1) { res.status = GC_OVERFLOW_ITEMS; return 0; }
		GcSliceMeta& m = ws.slices[0];
		m.correctLogOdds = vt.initialCorrect;
		m.falseLogOdds = vt.initialFalse;
		m.correctFromCorrect = 0;
		m.falseFromCorrect = 0;
		m.minScore = 0;
		m.minScoreNode = startNode;
		m.minScoreNodeOffset = startOffset;
		m.bandwidth = 1;
		m.firstItem = 0;
		m.numItems = 1;
		GcNodeItem& it = ws.items[0];
		uint32_t len = g.nodeLength[startNode];
		it.startVP = 0; it.startVN = 0; it.startScore = (int32_t)startOffset;
		it.endVP = 0; it.endVN = 0; it.endScore = (int32_t)len - 1 - (int32_t)startOffset;
		it.minScore = 0;
		it.nodeAndFlag = startNode;
		// HN bits 1..startOffset, HP bits startOffset+1..len-1
		uint64_t upTo = startOffset >= 63 ? ~0ULL : ((2ULL << startOffset) - 1);
		uint64_t lenMask = len >= 64 ? ~0ULL : ((1ULL << len) - 1);
		it.HN = upTo & ~1ULL;
		it.HP = lenMask & ~upTo;
		itemsUsed = 1;
	}
	const int32_t bandwidth = prm.bandwidth;
	uint32_t* keysCur = ws.keysA;   // keys of the slice being filled (the initial slice first)
	uint32_t* keysPrev = ws.keysB;
	if (keysCur) keysCur[0] = startNode;
	int32_t lastSlice = 0;
	int32_t status = GC_OK;
	uint64_t columns = 0;
	// ---- state of the slice being filled
	int32_t slice = -1, j = 0;
	const GcNodeItem* prevItems = nullptr;
	uint32_t prevN = 0;
	int32_t previousMinScore = 0, previousQuitScore = 0;
	uint64_t eq[4] = { 0, 0, 0, 0 };
	uint32_t heapSize = 0;
	uint32_t firstItem = 0, curN = 0;
	GcNodeItem* curItems = ws.items;
	int32_t sliceMinScore = 0, currentMinScoreAtEndRow = 0;
	uint32_t sliceMinNode = 0xFFFFFFFFu, sliceMinOffset = 0xFFFFFFFFu;
	uint64_t lastKey = ~0ULL;
	bool needFlatten = false;
	// ---- state of the node in flight
	bool pending = false;
	uint32_t node = 0;
	GcWord w; w.VP = 0; w.VN = 0; w.scoreEnd = 0;
	GcWord endW = w;
	uint64_t HP = 0, HN = 0, prevHP = 0, prevHN = 0;
	int32_t nodeMin = 0, prevStart = 0;
	uint32_t nodeMinOffset = 0;
	bool prevExists = false;
	// out-neighbours of the node in flight with their queue keys: requested when the node is popped, consumed when its end
	// column is pushed -- the three dependent loads (CSR range, neighbour, component number) complete under the column loop
	uint32_t outBegin = 0, outEnd = 0;
	uint64_t outKey[3] = { 0, 0, 0 };
	bool running = numSlices > 0;
	while (running)
	{
		// ================= stage 1 =================
		while (true)
		{
			if (pending)
			{
				// ---- store the node, push its end column to the out-neighbours (Banded.h:363-387); old end = {0,0,INT_MAX}
				pending = false;
				GcNodeItem& item = ws.items[itemsUsed++];
				curN++;
				item.startVP = w.VP; item.startVN = w.VN; item.startScore = w.scoreEnd;
				item.endVP = endW.VP; item.endVN = endW.VN; item.endScore = endW.scoreEnd;
				item.HP = HP; item.HN = HN;
				item.minScore = nodeMin;
				if (nodeMin < currentMinScoreAtEndRow) currentMinScoreAtEndRow = nodeMin;
				int32_t sbsEnd = gc_sbs(endW);
				uint64_t VP = endW.VP, VN = endW.VN;
				uint64_t plm = (VP & (VN - VP));
				plm >>= 1;
				plm |= 0x8000000000000000ULL & (VN | ~(VN - VP)) & ~VP;
				int32_t newEndMinScore = sbsEnd;
				while (plm != 0)
				{
					uint64_t cm = plm ^ (plm - 1);
					int32_t sh = sbsEnd + gc_popc(VP & cm) - gc_popc(VN & cm);
					if (sh < newEndMinScore) newEndMinScore = sh;
					plm &= ~cm;
				}
				uint32_t flag = 0;
				if (newEndMinScore <= currentMinScoreAtEndRow + bandwidth)
				{
					flag = 0x80000000u;
					for (uint32_t e = outBegin; e < outEnd; e++)
					{
						uint64_t pushKey;
						if (e - outBegin < 3) pushKey = outKey[e - outBegin];
						else { uint32_t nb = g.outNbr[e]; pushKey = ((uint64_t)g.componentNumber[nb] << 32) | nb; }
						if (!gc_heap_push(ws.heap, heapSize, ws.heapCap, pushKey)) { status = GC_OVERFLOW_HEAP; running = false; break; }
					}
				}
				item.nodeAndFlag = node | flag;
				if (keysCur && curN <= 32) keysCur[curN - 1] = node | flag;
				if (nodeMin < sliceMinScore)
				{
					sliceMinScore = nodeMin;
					sliceMinNode = node;
					sliceMinOffset = nodeMinOffset;
				}
				if (!running) break;
			}
			if (heapSize == 0)
			{
				if (slice >= 0)
				{
					// ---- close the slice (Banded.h:589-607)
					if (sliceMinNode == 0xFFFFFFFFu) { status = GC_INTERNAL; running = false; break; }
					if (j + 64 > seqLen) { needFlatten = true; running = false; break; } // partial last slice: flattened below
					const GcSliceMeta& pm = ws.slices[lastSlice];
					GcSliceMeta& nm = ws.slices[lastSlice + 1];
					nm.minScore = sliceMinScore;
					nm.minScoreNode = sliceMinNode;
					nm.minScoreNodeOffset = sliceMinOffset;
					nm.bandwidth = bandwidth;
					nm.firstItem = firstItem;
					nm.numItems = curN;
					gc_viterbi_next(vt, pm, sliceMinScore - previousMinScore, nm);
					if (!nm.correctFromCorrect) { running = false; break; } // the new slice is dropped
					lastSlice++;
					if (slice + 1 >= numSlices) { running = false; break; }
				}
				// ---- open the next slice: seed the queue from the previous one (Banded.h:235-277)
				slice++;
				j = slice * 64;
				const GcSliceMeta& pm = ws.slices[lastSlice];
				prevItems = ws.items + pm.firstItem;
				prevN = pm.numItems;
				{ uint32_t* tmp = keysPrev; keysPrev = keysCur; keysCur = tmp; } // the slice just closed is the previous one now
				previousMinScore = pm.minScore;
				previousQuitScore = pm.minScore + pm.bandwidth;
				gc_eq_vector(seq, seqLen, j, eq, g.coopLane, g.coopWidth, g.coopMask, g.coopShift);
				for (uint32_t k = 0; k < prevN; k++)
				{
					const GcNodeItem& pn = prevItems[k];
					uint32_t nd = gc_item_node(pn);
					if (j > 0)
					{
						if (pn.minScore > previousQuitScore) continue;
						if (g.linearizable[nd])
						{
							uint32_t nb = g.inNbr[g.inStart[nd]];
							const GcNodeItem* nbItem = gc_find_item(g, prevItems, prevN, nb, keysPrev);
							if (nbItem && nbItem->endScore < previousQuitScore && nbItem->minScore < previousQuitScore) continue;
						}
					}
					if (!gc_heap_push(ws.heap, heapSize, ws.heapCap, ((uint64_t)g.componentNumber[nd] << 32) | nd)) { status = GC_OVERFLOW_HEAP; running = false; break; }
				}
				if (!running) break;
				firstItem = itemsUsed;
				curN = 0;
				curItems = ws.items + firstItem;
				sliceMinScore = GC_INT_MAX - bandwidth - 1;
				sliceMinNode = 0xFFFFFFFFu; sliceMinOffset = 0xFFFFFFFFu;
				currentMinScoreAtEndRow = sliceMinScore;
				lastKey = ~0ULL;
				if (heapSize == 0) continue; // nothing seeded: closed (as an internal error) at the top
			}
			// ---- pop the next node, merge its incoming columns (BVCommon.h:903-964)
			uint64_t key = gc_heap_pop(ws.heap, heapSize);
			if (key == lastKey) continue;
			lastKey = key;
			node = (uint32_t)key;
			outBegin = g.outStart[node]; outEnd = g.outStart[node + 1];
			#pragma unroll
			for (uint32_t i = 0; i < 3; i++)
				if (outBegin + i < outEnd) { uint32_t nb = g.outNbr[outBegin + i]; outKey[i] = ((uint64_t)g.componentNumber[nb] << 32) | nb; }
			// everything the visit reads about the node itself, requested together (one memory round trip instead of a chain)
			const uint32_t inBegin = g.inStart[node], inEnd = g.inStart[node + 1];
			const uint32_t nodeLen = g.nodeLength[node];
			const bool nodeLinearizable = g.linearizable[node] != 0;
			const int firstBase = (int)(g.nodeSeq[2 * (uint64_t)node] & 3);
			const uint32_t firstIn = inBegin < inEnd ? g.inNbr[inBegin] : 0;
			const GcNodeItem* prevItem = gc_find_item(g, prevItems, prevN, node, keysPrev);
			prevExists = prevItem != nullptr;
			prevStart = prevExists ? prevItem->startScore : 0;
			prevHP = prevExists ? prevItem->HP : ~0ULL;
			prevHN = prevExists ? prevItem->HN : 0ULL;
			bool hasWs = false;
			w.VP = 0; w.VN = 0; w.scoreEnd = 0;
			bool seeded = false;
			if (prevExists)
			{
				seeded = true;
				if (j > 0)
				{
					if (prevItem->minScore > previousQuitScore) seeded = false;
					else if (nodeLinearizable)
					{
						uint32_t nb = firstIn;
						const GcNodeItem* nbItem = gc_find_item(g, prevItems, prevN, nb, keysPrev);
						if (nbItem && nbItem->endScore < previousQuitScore && nbItem->minScore < previousQuitScore) seeded = false;
					}
				}
			}
			if (seeded)
			{
				w.VP = ~0ULL; w.VN = 0; w.scoreEnd = prevStart + 64; // getSourceSliceFromScore
				hasWs = true;
			}
			uint64_t Eq0 = eq[firstBase];
			for (uint32_t e = inBegin; e < inEnd; e++)
			{
				uint32_t p = e == inBegin ? firstIn : g.inNbr[e];
				const GcNodeItem* pit = gc_find_item(g, curItems, curN, p, keysCur);
				if (!pit || !(pit->nodeAndFlag & 0x80000000u)) continue;
				GcWord inc = gc_item_end(*pit);
				uint64_t hinP, hinN;
				if (prevExists)
				{
					int32_t incSbs = gc_sbs(inc);
					if (prevStart < incSbs) { hinP = 0; hinN = 1; }
					else if (prevStart > incSbs) { hinP = 1; hinN = 0; }
					else { hinP = 0; hinN = 0; }
				}
				else { hinP = 1; hinN = 0; }
				uint64_t oP, oN;
				GcWord nw = gc_next_column(Eq0, inc, hinP, hinN, oP, oN);
				if (!prevExists || gc_sbs(nw) < prevStart)
				{
					nw.VP &= ~1ULL;
					nw.VN |= 1;
				}
				if (!hasWs) { w = nw; hasWs = true; }
				else w = gc_merge(w, nw);
				columns++;
			}
			if (!hasWs) { status = GC_INTERNAL; running = false; break; }
			if (prevExists && gc_sbs(w) > prevStart)
			{
				GcWord src; src.VP = ~0ULL; src.VN = 0; src.scoreEnd = prevStart + 64;
				w = gc_merge(w, src);
			}
			if (itemsUsed >= ws.itemCap) { status = GC_OVERFLOW_ITEMS; running = false; break; }
			uint32_t len = nodeLen;
			columns += len;
			pending = true;
			if (len > 1) break; // to the column steps
			// a single-column node has no column steps (the tail of calculateNodeInner does nothing)
			endW = w; HP = 0; HN = 0; nodeMin = w.scoreEnd; nodeMinOffset = 0;
		}
		if (!running) break;
		// ================= stage 2 =================
		gc_node_columns(g, node, eq, w, prevExists, prevStart, prevHP, prevHN, endW, HP, HN, nodeMin, nodeMinOffset, nullptr);
	}
	res.columns = columns;
	res.status = status;
	if (status != GC_OK) return 0;
	// ---- flattenLastSliceEnd (BVCommon.h:1171-1229) for a partial last slice, then close it
	if (needFlatten)
	{
		uint32_t rows = (uint32_t)(seqLen - j);
		uint64_t rowMask = ~(~0ULL << rows);
		// phmap iteration order of the slice's node map: reserve(previous size), insert in processing order
		uint32_t* slots = (uint32_t*)ws.heap;
		uint32_t capacity = 0;
		if (!gc_phmap_order(curItems, curN, prevN, slots, ws.heapCap * 2, &capacity)) { res.status = GC_OVERFLOW_HEAP; return 0; }
		sliceMinScore = GC_INT_MAX;
		sliceMinNode = 0xFFFFFFFFu;
		sliceMinOffset = 0xFFFFFFFFu;
		for (uint32_t s = 0; s < capacity; s++)
		{
			uint32_t k = slots[s];
			if (k == 0xFFFFFFFFu) continue;
			const GcNodeItem& it = curItems[k];
			uint32_t nd = gc_item_node(it);
			const GcNodeItem* old = gc_find_item(g, prevItems, prevN, nd, keysPrev);
			// the reference recomputes the node's columns, flattens each (flattenWordSlice, BVCommon.h:265-273) and keeps the first
			// strict minimum in map order; the column run tracks exactly that minimum, so the columns are never stored
			int32_t nodeFlatMin; uint32_t nodeFlatOffset;
			gc_recalc_node(g, it, eq, old, nullptr, ~rowMask, &nodeFlatMin, &nodeFlatOffset);
			res.columns += g.nodeLength[nd];
			if (nodeFlatMin < sliceMinScore)
			{
				sliceMinScore = nodeFlatMin;
				sliceMinNode = nd;
				sliceMinOffset = nodeFlatOffset;
			}
		}
		const GcSliceMeta& pm = ws.slices[lastSlice];
		GcSliceMeta& nm = ws.slices[lastSlice + 1];
		nm.minScore = sliceMinScore;
		nm.minScoreNode = sliceMinNode;
		nm.minScoreNodeOffset = sliceMinOffset;
		nm.bandwidth = bandwidth;
		nm.firstItem = firstItem;
		nm.numItems = curN;
		gc_viterbi_next(vt, pm, sliceMinScore - previousMinScore, nm);
		if (nm.correctFromCorrect) lastSlice++;
	}
	res.itemsUsed = itemsUsed;
	// ---- removeWronglyAlignedEnd (BVCommon.h:1231-1241)
	int32_t count = lastSlice + 1;
	{
		bool currentlyCorrect = ws.slices[count - 1].correctLogOdds > ws.slices[count - 1].falseLogOdds;
		while (!currentlyCorrect)
		{
			currentlyCorrect = ws.slices[count - 1].falseFromCorrect;
			count--;
			if (count == 0) break;
		}
	}
	return count - 1;
}

// ------------------------------------------------------------------------------------
// Backtrace (BVCommon.h:392-544).  `last` = index of the last kept slice (>= 1).
struct GcTraceWriter
{
	uint64_t* out;
	uint32_t cap;
	uint32_t n;
	uint32_t node; uint32_t offset; int32_t seqPos; // trace.back()
	bool overflow;
	GC_HD void push(uint32_t nd, uint32_t off, int32_t sp, bool sw)
	{
		if (n < cap) out[n] = gc_pack_trace(nd, off, sp, sw); else overflow = true;
		n++;
		node = nd; offset = off; seqPos = sp;
	}
};

struct GcBtPos { uint32_t node; uint32_t offset; int32_t seqPos; bool nodeSwitch; };

// pickBacktraceCorner (BVCommon.h:710-804); scoresNotValid is always false here
// findCur / findPrev: node -> const GcNodeItem* in the current / previous slice (nullptr if absent)
template <typename FindCur, typename FindPrev>
GC_HD bool gc_bt_corner_with(const GcGraphView& g, const FindCur& findCur, const FindPrev& findPrev, uint32_t node, int32_t j, const uint8_t* seq, int32_t quitScore, int32_t previousQuitScore, GcBtPos& out)
{
	const GcNodeItem* me = findCur(node);
	int32_t scoreHere = gc_value(gc_item_start(*me), 0);
	const GcNodeItem* prevMe = findPrev(node);
	if (scoreHere > quitScore)
	{
		int32_t smallestFound = scoreHere + 1;
		out.node = 0; out.offset = 0; out.seqPos = 0; out.nodeSwitch = false;
		if (prevMe)
		{
			smallestFound = prevMe->startScore;
			out.node = node; out.offset = 0; out.seqPos = j - 1; out.nodeSwitch = false;
		}
		for (uint32_t e = g.inStart[node]; e < g.inStart[node + 1]; e++)
		{
			uint32_t nb = g.inNbr[e];
			const GcNodeItem* pn = findPrev(nb);
			if (pn)
			{
				if (pn->endScore <= smallestFound)
				{
					smallestFound = pn->endScore;
					out.node = nb; out.offset = g.nodeLength[nb] - 1; out.seqPos = j - 1; out.nodeSwitch = true;
				}
			}
			const GcNodeItem* cn = findCur(nb);
			if (cn && nb != node)
			{
				int32_t v = gc_value(gc_item_end(*cn), 0);
				if (v < smallestFound)
				{
					smallestFound = v;
					out.node = nb; out.offset = g.nodeLength[nb] - 1; out.seqPos = j; out.nodeSwitch = true;
				}
			}
		}
		return true;
	}
	bool eq = gc_char_match(seq[j], gc_node_base(g, node, 0));
	if (prevMe && prevMe->startScore == scoreHere - 1)
	{
		out.node = node; out.offset = 0; out.seqPos = j - 1; out.nodeSwitch = false;
		return true;
	}
	bool haveInvalid = false;
	GcBtPos bestInvalid; bestInvalid.node = 0; bestInvalid.offset = 0; bestInvalid.seqPos = 0; bestInvalid.nodeSwitch = true;
	int32_t bestInvalidScore = scoreHere + 1;
	for (uint32_t e = g.inStart[node]; e < g.inStart[node + 1]; e++)
	{
		uint32_t nb = g.inNbr[e];
		const GcNodeItem* cn = findCur(nb);
		if (cn && gc_value(gc_item_end(*cn), 0) == scoreHere - 1)
		{
			out.node = nb; out.offset = g.nodeLength[nb] - 1; out.seqPos = j; out.nodeSwitch = true;
			return true;
		}
		const GcNodeItem* pn = findPrev(nb);
		if (pn)
		{
			int32_t cornerScore = pn->endScore;
			if (cornerScore > previousQuitScore)
			{
				if (cornerScore < bestInvalidScore)
				{
					bestInvalidScore = cornerScore;
					bestInvalid.node = nb; bestInvalid.offset = g.nodeLength[nb] - 1; bestInvalid.seqPos = j - 1;
					haveInvalid = true;
				}
			}
			else if (cornerScore == scoreHere - (eq ? 0 : 1))
			{
				out.node = nb; out.offset = g.nodeLength[nb] - 1; out.seqPos = j - 1; out.nodeSwitch = true;
				return true;
			}
		}
	}
	if (bestInvalidScore < scoreHere + 1 && haveInvalid)
	{
		out = bestInvalid;
		return true;
	}
	return false; // the reference asserts here
}

GC_HD bool gc_bt_corner(const GcGraphView& g, const GcNodeItem* cur, uint32_t curN, const GcNodeItem* prev, uint32_t prevN, uint32_t node, int32_t j, const uint8_t* seq, int32_t quitScore, int32_t previousQuitScore, GcBtPos& out)
{
	auto findCur = [&](uint32_t nd) { return gc_find_item(g, cur, curN, nd); };
	auto findPrev = [&](uint32_t nd) { return gc_find_item(g, prev, prevN, nd); };
	return gc_bt_corner_with(g, findCur, findPrev, node, j, seq, quitScore, previousQuitScore, out);
}

GC_HD void gc_k1_backtrace(const GcGraphView& g, const uint8_t* seq, int32_t seqLen, GcK1Workspace& ws, int32_t last, uint64_t* traceOut, uint32_t traceCap, GcK1Result& res)
{
	GcTraceWriter tw;
	tw.out = traceOut; tw.cap = traceCap; tw.n = 0; tw.overflow = false;
	const GcSliceMeta& lm = ws.slices[last];
	res.score = lm.minScore;
	{
		int32_t sp = (last - 1) * 64 + 63;
		if (sp > seqLen - 1) sp = seqLen - 1;
		tw.push(lm.minScoreNode, lm.minScoreNodeOffset, sp, false);
	}
	uint32_t currentNode = 0xFFFFFFFFu;
	int32_t currentSlice = -1;
	GcColVV* cols = ws.cols;
	GcCols cv; cv.c = cols; cv.HP = 0; cv.HN = 0; cv.score0 = 0;
	uint64_t eq[4];
	uint32_t guard = 0;
	uint32_t guardMax = (uint32_t)seqLen * 4 + 1024 + traceCap;
	while (tw.seqPos != -1)
	{
		if (++guard > guardMax) { res.status = GC_INTERNAL; return; }
		int32_t newSlice = tw.seqPos / 64 + 1;
		uint32_t newNode = tw.node;
		const GcSliceMeta& cm = ws.slices[newSlice];
		const GcSliceMeta& pmeta = ws.slices[newSlice - 1];
		const GcNodeItem* cur = ws.items + cm.firstItem;
		const GcNodeItem* prev = ws.items + pmeta.firstItem;
		int32_t j = (newSlice - 1) * 64;
		if (newSlice != currentSlice || newNode != currentNode)
		{
			if (newSlice != currentSlice) gc_eq_vector(seq, seqLen, j, eq, g.coopLane, g.coopWidth, g.coopMask, g.coopShift);
			currentSlice = newSlice;
			currentNode = newNode;
			const GcNodeItem* me = gc_find_item(g, cur, cm.numItems, currentNode);
			if (!me) { res.status = GC_INTERNAL; return; }
			const GcNodeItem* pme = gc_find_item(g, prev, pmeta.numItems, currentNode);
			gc_recalc_node(g, *me, eq, pme, cols, 0, nullptr, nullptr, &cv.score0);
			cv.HP = me->HP; cv.HN = me->HN;
			res.columns += g.nodeLength[currentNode];
		}
		// inside the node (BVCommon.h:556-597): walk to the node's first column or the slice's first row, then cross below
		// (one loop iteration = one node visit: recompute, walk, cross -- the lanes of a warp stay in step)
		if ((tw.seqPos & 63) != 0 && tw.offset != 0)
		{
			// The reference reads three cell values per step (here, above, diagonal: getValue = two 64-bit popcounts each).
			// Same values, incrementally: along a column value(row-1) = value(row) + VN[row] - VP[row], so the cell above is a
			// bit test, and the value of the left neighbour column at the current row is carried along and only recomputed
			// (one getValue) when the walk moves a column to the left.
			uint32_t hori = tw.offset;
			int32_t vert = tw.seqPos - j;
			const uint64_t chunk0 = g.nodeSeq[2 * (uint64_t)currentNode], chunk1 = g.nodeSeq[2 * (uint64_t)currentNode + 1];
			GcWord cur = gc_cols_get(cv, hori), left = gc_cols_get(cv, hori - 1);
			int32_t scoreHere = gc_value(cur, vert);
			int32_t leftHere = gc_value(left, vert);
			while (hori > 0 && vert > 0)
			{
				int32_t dv = (int32_t)((cur.VN >> vert) & 1) - (int32_t)((cur.VP >> vert) & 1);    // value(cur, vert-1) - value(cur, vert)
				int32_t dl = (int32_t)((left.VN >> vert) & 1) - (int32_t)((left.VP >> vert) & 1); // the same in the left column
				int32_t diagonalScore = leftHere + dl;
				int base = (int)(((hori < 32 ? chunk0 : chunk1) >> ((hori & 31) * 2)) & 3);
				bool eqc = gc_char_match(seq[vert + j], base);
				if (dv == -1)
				{
					vert--;
					scoreHere -= 1;
					leftHere = diagonalScore;
				}
				else
				{
					if (diagonalScore == scoreHere - (eqc ? 0 : 1)) { vert--; scoreHere = diagonalScore; }
					else scoreHere = leftHere;
					hori--;
					cur = left;
					if (hori > 0) { left = gc_cols_get(cv, hori - 1); leftHere = gc_value(left, vert); }
				}
				tw.push(currentNode, hori, vert + j, false);
			}
		}
		int32_t quitScore = cm.minScore + cm.bandwidth;
		int32_t previousQuitScore = pmeta.minScore + pmeta.bandwidth;
		if ((tw.seqPos & 63) == 0 && tw.offset == 0)
		{
			GcBtPos bt;
			if (!gc_bt_corner(g, cur, cm.numItems, prev, pmeta.numItems, currentNode, j, seq, quitScore, previousQuitScore, bt)) { res.status = GC_INTERNAL; return; }
			tw.push(bt.node, bt.offset, bt.seqPos, bt.nodeSwitch);
			continue;
		}
		if ((tw.seqPos & 63) == 0)
		{
			// vertical crossing (BVCommon.h:451-477, 665-708)
			const GcNodeItem* pme = gc_find_item(g, prev, pmeta.numItems, currentNode);
			if (!pme)
			{
				tw.push(currentNode, 0, tw.seqPos, false);
				continue;
			}
			uint32_t off = tw.offset;
			int32_t sp = tw.seqPos;
			uint32_t origOff = off;
			while (off > 0 && gc_value(gc_cols_get(cv, off - 1), 0) == gc_value(gc_cols_get(cv, off), 0) - 1) off--;
			GcBtPos second;
			if (off == 0)
			{
				if (!gc_bt_corner(g, cur, cm.numItems, prev, pmeta.numItems, currentNode, j, seq, quitScore, previousQuitScore, second)) { res.status = GC_INTERNAL; return; }
			}
			else
			{
				bool eqc = gc_char_match(seq[sp], gc_node_base(g, currentNode, off));
				int32_t scoreHere = gc_value(gc_cols_get(cv, off), 0);
				int32_t scoreDiagonal = pme->startScore;
				for (uint32_t i = 1; i + 1 <= off; i++)
				{
					scoreDiagonal += (int32_t)((pme->HP >> i) & 1);
					scoreDiagonal -= (int32_t)((pme->HN >> i) & 1);
				}
				int32_t scoreUp = scoreDiagonal;
				scoreUp += (int32_t)((pme->HP >> off) & 1);
				scoreUp -= (int32_t)((pme->HN >> off) & 1);
				second.node = currentNode; second.seqPos = sp - 1; second.nodeSwitch = false;
				if (scoreHere > quitScore || scoreDiagonal > previousQuitScore || scoreUp > previousQuitScore)
				{
					second.offset = (scoreDiagonal < scoreUp) ? off - 1 : off;
				}
				else if (scoreUp == scoreHere - 1) second.offset = off;
				else
				{
					if (scoreDiagonal != scoreHere - (eqc ? 0 : 1)) { res.status = GC_INTERNAL; return; }
					second.offset = off - 1;
				}
			}
			if (off != origOff)
			{
				for (uint32_t o = origOff - 1; o != off; o--) tw.push(currentNode, o, sp, false);
			}
			if (off != tw.offset || sp != tw.seqPos) tw.push(currentNode, off, sp, false);
			tw.push(second.node, second.offset, second.seqPos, second.nodeSwitch);
			continue;
		}
		if (tw.offset == 0)
		{
			// horizontal crossing (BVCommon.h:478-499, 599-663)
			const GcNodeItem* me = gc_find_item(g, cur, cm.numItems, currentNode);
			GcWord startSlice = gc_item_start(*me);
			int32_t sp = tw.seqPos;
			int32_t origSp = sp;
			while ((sp & 63) != 0 && (startSlice.VP & (1ULL << (sp & 63)))) sp--;
			int32_t offset = sp & 63;
			GcBtPos second;
			if (offset == 0)
			{
				if (!gc_bt_corner(g, cur, cm.numItems, prev, pmeta.numItems, currentNode, j, seq, quitScore, previousQuitScore, second)) { res.status = GC_INTERNAL; return; }
			}
			else
			{
				bool eqc = gc_char_match(seq[sp], gc_node_base(g, currentNode, 0));
				int32_t scoreHere = gc_value(startSlice, offset);
				bool found = false;
				if (scoreHere > quitScore)
				{
					int32_t smallestFound = gc_value(startSlice, offset - 1);
					second.node = currentNode; second.offset = 0; second.seqPos = sp - 1; second.nodeSwitch = false;
					for (uint32_t e = g.inStart[currentNode]; e < g.inStart[currentNode + 1]; e++)
					{
						uint32_t nb = g.inNbr[e];
						const GcNodeItem* cn = gc_find_item(g, cur, cm.numItems, nb);
						if (!cn) continue;
						GcWord ns = gc_item_end(*cn);
						int32_t v1 = gc_value(ns, offset - 1);
						if (v1 <= smallestFound)
						{
							smallestFound = v1;
							second.node = nb; second.offset = g.nodeLength[nb] - 1; second.seqPos = sp - 1; second.nodeSwitch = true;
						}
						int32_t v0 = gc_value(ns, offset);
						if (v0 < smallestFound && nb != currentNode)
						{
							smallestFound = v0;
							second.node = nb; second.offset = g.nodeLength[nb] - 1; second.seqPos = sp; second.nodeSwitch = true;
						}
					}
					found = true;
				}
				else
				{
					for (uint32_t e = g.inStart[currentNode]; e < g.inStart[currentNode + 1] && !found; e++)
					{
						uint32_t nb = g.inNbr[e];
						const GcNodeItem* cn = gc_find_item(g, cur, cm.numItems, nb);
						if (!cn) continue;
						GcWord ns = gc_item_end(*cn);
						if (gc_value(ns, offset) == scoreHere - 1)
						{
							second.node = nb; second.offset = g.nodeLength[nb] - 1; second.seqPos = sp; second.nodeSwitch = true;
							found = true;
						}
						else if (gc_value(ns, offset - 1) == scoreHere - (eqc ? 0 : 1))
						{
							second.node = nb; second.offset = g.nodeLength[nb] - 1; second.seqPos = sp - 1; second.nodeSwitch = true;
							found = true;
						}
					}
				}
				if (!found) { res.status = GC_INTERNAL; return; }
			}
			if (sp != origSp)
			{
				for (int32_t s = origSp - 1; s != sp; s--) tw.push(currentNode, 0, s, false);
			}
			if (sp != tw.seqPos) tw.push(currentNode, 0, sp, false);
			tw.push(second.node, second.offset, second.seqPos, second.nodeSwitch);
			continue;
		}
	}
	// ---- slide left in row -1 (BVCommon.h:508-542); the do/while(false) runs once
	{
		const GcSliceMeta& m0 = ws.slices[0];
		const GcNodeItem* s0 = ws.items + m0.firstItem;
		const GcNodeItem* it = gc_find_item(g, s0, m0.numItems, tw.node);
		if (!it) { res.status = GC_INTERNAL; return; }
		// beforeSliceScores[i] = startScore + sum_{k=1..i} (HP_k - HN_k)
		uint32_t off = tw.offset;
		int32_t here = it->startScore + gc_popc(it->HP & ((2ULL << off) - 2)) - gc_popc(it->HN & ((2ULL << off) - 2));
		while (here != 0 && off > 0)
		{
			int32_t before = here - (int32_t)((it->HP >> off) & 1) + (int32_t)((it->HN >> off) & 1);
			if (before != here - 1) break;
			off--;
			here = before;
			tw.push(tw.node, off, tw.seqPos, false);
		}
		if (off == 0 && here != 0)
		{
			for (uint32_t e = g.inStart[tw.node]; e < g.inStart[tw.node + 1]; e++)
			{
				uint32_t nb = g.inNbr[e];
				const GcNodeItem* nit = gc_find_item(g, s0, m0.numItems, nb);
				if (nit && gc_sbs(gc_item_end(*nit)) == here - 1)
				{
					tw.push(nb, g.nodeLength[nb] - 1, tw.seqPos, true);
					break;
				}
			}
		}
	}
	res.traceLen = tw.n;
	if (tw.overflow) res.status = GC_OVERFLOW_TRACE;
}

// one complete work item
GC_HD void gc_k1_extend(const GcGraphView& g, const GcViterbiTables& vt, const GcK1Params& prm, const uint8_t* seq, int32_t seqLen, uint32_t startNode, uint32_t startOffset,
	GcK1Workspace& ws, uint64_t* traceOut, uint32_t traceCap, GcK1Result& res)
{
	res.score = GC_INT_MAX;
	res.traceLen = 0;
	res.itemsUsed = 0;
	int32_t last = gc_k1_forward(g, vt, prm, seq, seqLen, startNode, startOffset, ws, res);
	if (res.status != GC_OK) return;
	if (last < 1) { res.status = GC_FAILED; return; }
	gc_k1_backtrace(g, seq, seqLen, ws, last, traceOut, traceCap, res);
}
