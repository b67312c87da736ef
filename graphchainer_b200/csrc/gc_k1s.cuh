// K1, lane-per-item form -- the same work item as gc_k1.cuh (GraphAlignerBitvectorBanded::getReverseTraceFromSeed,
// GraphAlignerBitvectorBanded.h:46-71), organised so that the 32 lanes of a warp carry 32 DIFFERENT items and still
// execute mostly the same instructions.
//
// Why: the lock-step form (one warp per item, gcgpu.cu gc_k1_long_kernel) saturates the integer pipe with 32 lanes
// computing identical values (ncu r02h: pipe_alu 82 %, 1 useful lane in 32).  A plain thread-per-item launch of the nested
// reference loop serialises instead: the 32 walks of a warp are in different slices / nodes / phases at any time.
//
// Shape here: every lane walks its own item, but the walk is cut so that the warp meets at two loop heads with
// warp-uniform conditions (GC_WARP_ANY / GC_WARP_MAX):
//   stage 1   node bookkeeping, repeated while some lane has no multi-column node in hand: store the finished node and push
//             its end column (Banded.h:363-387), close / open slices (Banded.h:589-607, 235-277), pop the next node and
//             merge its incoming columns (BVCommon.h:903-964).  One-column nodes (SNP alleles: most node visits of a
//             variation graph) finish inside this stage.
//   stage 2   ONE column loop for the whole warp (BVCommon.h:1118-1161): trip count = the longest node among the lanes,
//             a lane whose node is shorter idles for the rest.  This loop is where the column steps -- the work -- are.
//             (Cutting the loop into rounds of 16 columns, so that lanes with short nodes go on while a 64-column node is
//             finished over four rounds, raises the active lanes per instruction but makes every launch slower -- r03n,
//             68-74 ms instead of 48-50: a launch lasts as long as its longest item, and that item then pays the
//             bookkeeping stage of the whole warp four times per long node.)
// The partial last slice of every item (flattenLastSliceEnd, BVCommon.h:1171-1229) is handled the same way after the
// main loop: the lanes recompute their slices' nodes in phmap slot order with a shared column loop.
// The backtrace (BVCommon.h:392-544) has the same two stages: per path node a lookup / crossing stage and a shared
// column loop that recomputes the node (recalcNodeWordslice, BVCommon.h:828-852), then a shared cell-walk loop.
//
// Per-lane state lives in registers; slices, node items and slice keys in the item's slab in HBM (L1/L2-cached, a
// few hundred bytes hot per item); the node queue in shared memory (entry-major, conflict-free when lanes agree).
#pragma once
#include "gc_k1.cuh"

// binary min-heap of (componentNumber << 32 | node) keys whose entries are `stride` words apart
struct GcHeapRef { uint64_t* base; uint32_t stride; uint32_t cap; };
GC_HD uint64_t& gc_sheap_at(const GcHeapRef& h, uint32_t i) { return h.base[(size_t)i * h.stride]; }
GC_HD bool gc_sheap_push(const GcHeapRef& h, uint32_t& size, uint64_t key)
{
	if (size >= h.cap) return false;
	uint32_t i = size++;
	while (i > 0)
	{
		uint32_t p = (i - 1) >> 1;
		uint64_t pk = gc_sheap_at(h, p);
		if (pk <= key) break;
		gc_sheap_at(h, i) = pk;
		i = p;
	}
	gc_sheap_at(h, i) = key;
	return true;
}
GC_HD uint64_t gc_sheap_pop(const GcHeapRef& h, uint32_t& size)
{
	uint64_t top = gc_sheap_at(h, 0);
	uint64_t last = gc_sheap_at(h, --size);
	uint32_t i = 0;
	while (true)
	{
		uint32_t c = 2 * i + 1;
		if (c >= size) break;
		uint64_t ck = gc_sheap_at(h, c);
		if (c + 1 < size) { uint64_t ck2 = gc_sheap_at(h, c + 1); if (ck2 < ck) { ck = ck2; c++; } }
		if (ck >= last) break;
		gc_sheap_at(h, i) = ck;
		i = c;
	}
	if (size > 0) gc_sheap_at(h, i) = last;
	return top;
}

// The fields of a stored node that seeding the NEXT slice needs (Banded.h:235-277), 16 bytes beside the 64-byte item:
// the queue key's component number, the two scores of the skip tests, and where the node's only in-neighbour sits in
// the same slice when the node is linearizable (the test of Banded.h:257-266 reads that neighbour's scores).
struct __attribute__((aligned(16))) GcItemAux
{
	uint32_t comp;
	int32_t minScore, endScore;
	int32_t linIdx;        // >= 0: index (within the slice) of the in-neighbour of a linearizable node; -1: linearizable, neighbour absent; -2: not linearizable
};

// per work item memory (slab in HBM carved by the host, see gcgpu.cu) + the lane's queue
struct GcK1SWorkspace
{
	GcSliceMeta* slices;   // [numSlices + 2]
	GcNodeItem* items;     // [itemCap]
	uint32_t* keys;        // [itemCap] node | pushed flag of items[k]: a slice lookup scans 4-byte keys, not 64-byte items
	GcItemAux* aux;        // [itemCap] what seeding the next slice reads about items[k]
	uint32_t* scratch;     // phmap slot emulation of the last slice
	uint32_t scratchCap;   // entries
	uint32_t itemCap;
	GcHeapRef heap;
};

// index of `node` among the n keys of a slice, -1 if absent.  Slices hold a handful of nodes: independent loads, no search tree.
GC_HD int32_t gc_find_key(const uint32_t* keys, uint32_t n, uint32_t node)
{
	for (uint32_t base = 0; base < n; base += 8)
	{
		uint32_t end = base + 8 < n ? base + 8 : n;
		int32_t found = -1;
		for (uint32_t k = base; k < end; k++) if ((keys[k] & 0x7FFFFFFFu) == node) found = (int32_t)k;
		if (found >= 0) return found;
	}
	return -1;
}

// Eq masks of the 64 rows from bit `bit` of the sequence buffer's bit planes (plane b of 64-base block k at planes[4k+b];
// gc_planes_kernel), rows >= `rows` cleared -- the same masks gc_eq_vector derives from the codes, in eight loads
GC_HD void gc_eq_from_planes(const uint64_t* planes, uint64_t bit, int32_t rows, uint64_t eq[4])
{
	const uint64_t* p = planes + 4 * (bit >> 6);
	uint32_t s = (uint32_t)(bit & 63);
	uint64_t keep = rows >= 64 ? ~0ULL : ((1ULL << rows) - 1);
	for (int b = 0; b < 4; b++)
	{
		uint64_t lo = p[b], hi = p[4 + b];
		eq[b] = (s ? ((lo >> s) | (hi << (64 - s))) : lo) & keep;
	}
}

// The same lookup for a node expected at or just before index `hint` (an in-neighbour of the node last stored; the node a
// backtrace steps to): the scan runs from `hint` down to 0 in groups of 8 independent loads, then over the entries after `hint`.
// A node sits at most once in a slice, so the result is that of gc_find_key.
GC_HD int32_t gc_find_key_near(const uint32_t* keys, uint32_t n, uint32_t node, uint32_t hint)
{
	if (n == 0) return -1;
	if (hint >= n) hint = n - 1;
	for (int32_t top = (int32_t)hint; top >= 0; top -= 8)
	{
		const int32_t low = top >= 7 ? top - 7 : 0;
		int32_t found = -1;
		for (int32_t k = top; k >= low; k--) if ((keys[k] & 0x7FFFFFFFu) == node) found = k;
		if (found >= 0) return found;
	}
	for (uint32_t base = hint + 1; base < n; base += 8)
	{
		uint32_t end = base + 8 < n ? base + 8 : n;
		int32_t found = -1;
		for (uint32_t k = base; k < end; k++) if ((keys[k] & 0x7FFFFFFFu) == node) found = (int32_t)k;
		if (found >= 0) return found;
	}
	return -1;
}

// The warp's shared column loop, one window of 16 columns: every lane with `mine` computes columns [pos0, pos0 + 16) of
// its node (pos0 a multiple of 16; column 0 is the start column, the node ends at len), one getNextSlice step each
// (BVCommon.h:1118-1161).  The trip count is the longest window among the lanes.  The 16 bases of a window are one 32-bit
// word, and the unrolled body leaves one uniform exit test and one lane predicate per column (r03f: the loop head and the
// per-column base arithmetic of the rolled form were 21 % of the forward kernel's instructions).
// STORE: the columns are also written to `cols` (recalcNodeWordslice, BVCommon.h:828-852).
template <bool FLAT, bool STORE>
GC_HD void gc_k1s_columns16(GcColumnRun& run, bool mine, uint32_t pos0, uint32_t len, uint32_t forceUntil, GcColVV* cols)
{
	const uint32_t myTrip = mine ? (len - pos0 < 16u ? len - pos0 : 16u) : 0u;
	const uint32_t trip = GC_WARP_MAX(myTrip);
	const uint32_t bases = gc_col_bases(run, pos0 & 48u);
	#pragma unroll
	for (uint32_t u = 0; u < 16; u++)
	{
		if (u >= trip) break;
		const uint32_t pos = pos0 + u;
		if (u < myTrip && pos >= 1)
		{
			gc_col_step<FLAT>(run, pos, (int)((bases >> (2 * u)) & 3), forceUntil >= pos);
			if (STORE) { cols[pos].VP = run.ws.VP; cols[pos].VN = run.ws.VN; }
		}
	}
}
// all the columns of the nodes in hand (trip count = the longest node)
template <bool FLAT, bool STORE>
GC_HD void gc_k1s_columns(GcColumnRun& run, bool mine, uint32_t len, uint32_t forceUntil, GcColVV* cols)
{
	const uint32_t maxLen = GC_WARP_MAX(mine ? len : 0u);
	for (uint32_t pos0 = 0; pos0 < maxLen; pos0 += 16) gc_k1s_columns16<FLAT, STORE>(run, mine && pos0 < len, pos0, len, forceUntil, cols);
}

// ------------------------------------------------------------------------------------
// Forward pass (the lane-per-item twin of gc_k1_forward: same slices, items and Viterbi states, written to the same slab
// layout).  A lane works through items one after the other: when its item ends -- many extensions are cut after a few
// slices -- it takes the next one from `src`, so the lanes of a warp stay busy until the launch runs out of items
// (r03f, fixed 32 items per warp: 18.6 of 32 lanes still had an item on average).
//   src.next(item, ws)   the lane's next work item and its slab (false: none left)
//   src.done(res, last)  its result; last = index of the last kept slice after removeWronglyAlignedEnd, <= 0: the extension failed
struct GcK1SItem
{
	const uint8_t* seq;
	int32_t seqLen;
	uint32_t startNode, startOffset;
	uint64_t planeBit;     // where the item's sequence starts in the bit planes
};
template <class Source>
GC_HD void gc_k1s_forward_items(const GcGraphView& g, const GcViterbiTables& vt, const GcK1Params& prm, const uint64_t* planes, GcK1SWorkspace& ws, Source& src)
{
	const uint32_t NONE = 0xFFFFFFFFu;
	const int32_t bandwidth = prm.bandwidth;
	bool have = false, exhausted = false;
	GcK1SItem it; it.seq = nullptr; it.seqLen = 0; it.startNode = 0; it.startOffset = 0; it.planeBit = 0;
	uint32_t itemsUsed = 0;
	int32_t numSlices = 0;
	int32_t status = GC_OK;
	bool running = false;
	int32_t lastSlice = 0;
	uint64_t columns = 0;
	// ---- the slice being filled and the one before it (the last KEPT slice: its extent and scores stay in registers)
	int32_t slice = -1, j = 0;
	uint32_t prevFirst = 0, prevN = 0;
	int32_t previousMinScore = 0, previousQuitScore = 0;
	uint32_t keptFirst = 0, keptN = 1; int32_t keptMin = 0, keptBandwidth = 1; // slices[lastSlice]
	double keptCorrect = vt.initialCorrect, keptFalse = vt.initialFalse;
	uint64_t eq[4] = { 0, 0, 0, 0 };
	uint32_t heapSize = 0;
	uint32_t firstItem = 0, curN = 0;
	int32_t sliceMinScore = 0, currentMinScoreAtEndRow = 0;
	uint32_t sliceMinNode = NONE, sliceMinOffset = NONE;
	uint64_t lastKey = ~0ULL;
	// The previous slice's items lie in queue-key order (they were stored as they were popped) and the pops of this slice come
	// in the same order: one cursor walks the previous slice once per slice instead of one search per pop
	uint32_t prevCursor = 0; uint64_t prevCursorKey = ~0ULL;
	bool needFlatten = false;
	// ---- the node in hand
	bool pending = false;   // computed, not stored yet
	bool needCols = false;  // its columns 1..len-1 are still to be computed (stage 2)
	uint32_t node = 0, nodeComp = 0, len = 0, forceUntil = 0;
	int32_t nodeLinIdx = -2;
	GcWord w; w.VP = 0; w.VN = 0; w.scoreEnd = 0;
	GcWord endW = w;
	uint64_t HP = 0, HN = 0;
	int32_t nodeMin = 0;
	uint32_t nodeMinOffset = 0;
	// queue keys of the out-neighbours of the node in hand: requested when the node is popped, consumed when its end column is
	// pushed -- the loads complete under the column loop
	uint32_t outBegin = 0, outCount = 0;
	uint64_t outKey0 = 0, outKey1 = 0, outKey2 = 0;
	GcColumnRun run;
	run.ws = w; run.eq[0] = run.eq[1] = run.eq[2] = run.eq[3] = 0; run.prevHP = run.prevHN = run.HP = run.HN = 0; run.chunk0 = run.chunk1 = 0; run.minScore = 0; run.minOffset = 0; run.flatMask = 0;
	while (true)
	{
		// ================= stage 0: a lane without an item takes the next one =================
		if (!have && !exhausted)
		{
			if (!src.next(it, ws)) exhausted = true;
			else
			{
				have = true;
				status = GC_OK;
				columns = 0;
				itemsUsed = 0;
				numSlices = (it.seqLen + 63) / 64;
				running = numSlices > 0;
				// ---- getInitialSliceExactPosition (BVCommon.h:1243-1279)
				if (ws.itemCap < 1) { status = GC_OVERFLOW_ITEMS; running = false; }
				else
				{
					GcSliceMeta& m = ws.slices[0];
					m.correctLogOdds = vt.initialCorrect;
					m.falseLogOdds = vt.initialFalse;
					m.correctFromCorrect = 0;
					m.falseFromCorrect = 0;
					m.minScore = 0;
					m.minScoreNode = it.startNode;
					m.minScoreNodeOffset = it.startOffset;
					m.bandwidth = 1;
					m.firstItem = 0;
					m.numItems = 1;
					GcNodeItem& first = ws.items[0];
					uint32_t len0 = g.nodeLength[it.startNode];
					first.startVP = 0; first.startVN = 0; first.startScore = (int32_t)it.startOffset;
					first.endVP = 0; first.endVN = 0; first.endScore = (int32_t)len0 - 1 - (int32_t)it.startOffset;
					first.minScore = 0;
					first.nodeAndFlag = it.startNode;
					uint64_t upTo = it.startOffset >= 63 ? ~0ULL : ((2ULL << it.startOffset) - 1);
					uint64_t lenMask = len0 >= 64 ? ~0ULL : ((1ULL << len0) - 1);
					first.HN = upTo & ~1ULL;
					first.HP = lenMask & ~upTo;
					ws.keys[0] = it.startNode;
					GcItemAux a; a.comp = g.componentNumber[it.startNode]; a.minScore = 0; a.endScore = first.endScore; a.linIdx = -2; // slice -1 seeds slice 0 without the skip tests (j == 0)
					ws.aux[0] = a;
					itemsUsed = 1;
				}
				lastSlice = 0;
				slice = -1; j = 0;
				keptFirst = 0; keptN = 1; keptMin = 0; keptBandwidth = 1;
				keptCorrect = vt.initialCorrect; keptFalse = vt.initialFalse;
				heapSize = 0;
				needFlatten = false; pending = false; needCols = false;
			}
		}
		if (!GC_WARP_ANY(have)) break; // no lane has an item and none could get one
		// ================= stage 1: one bookkeeping step of every lane that has no multi-column node in hand =================
		if (running && !needCols)
		{
			bool go = true;
			if (pending)
			{
				// ---- store the node, push its end column to the out-neighbours (Banded.h:363-387); old end = {0,0,INT_MAX}
				pending = false;
				GcNodeItem& item = ws.items[itemsUsed];
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
					for (uint32_t e = 0; e < outCount; e++)
					{
						uint64_t pushKey = e == 0 ? outKey0 : (e == 1 ? outKey1 : (e == 2 ? outKey2 : g.outKey[outBegin + e]));
						if (!gc_sheap_push(ws.heap, heapSize, pushKey)) { status = GC_OVERFLOW_HEAP; running = false; go = false; break; }
					}
				}
				item.nodeAndFlag = node | flag;
				ws.keys[itemsUsed] = node | flag;
				GcItemAux a; a.comp = nodeComp; a.minScore = nodeMin; a.endScore = endW.scoreEnd; a.linIdx = nodeLinIdx;
				ws.aux[itemsUsed] = a;
				itemsUsed++;
				curN++;
				if (nodeMin < sliceMinScore)
				{
					sliceMinScore = nodeMin;
					sliceMinNode = node;
					sliceMinOffset = nodeMinOffset;
				}
			}
			if (go && heapSize == 0)
			{
				if (slice >= 0)
				{
					// ---- close the slice (Banded.h:589-607)
					if (sliceMinNode == NONE) { status = GC_INTERNAL; running = false; go = false; }
					else if (j + 64 > it.seqLen) { needFlatten = true; running = false; go = false; } // partial last slice: flattened below
					else
					{
						GcSliceMeta pm; pm.correctLogOdds = keptCorrect; pm.falseLogOdds = keptFalse;
						GcSliceMeta nm;
						nm.minScore = sliceMinScore;
						nm.minScoreNode = sliceMinNode;
						nm.minScoreNodeOffset = sliceMinOffset;
						nm.bandwidth = bandwidth;
						nm.firstItem = firstItem;
						nm.numItems = curN;
						for (int z = 0; z < 6; z++) nm.pad[z] = 0;
						gc_viterbi_next(vt, pm, sliceMinScore - previousMinScore, nm);
						ws.slices[lastSlice + 1] = nm;
						if (!nm.correctFromCorrect) { running = false; go = false; } // the new slice is dropped
						else
						{
							lastSlice++;
							keptFirst = firstItem; keptN = curN; keptMin = sliceMinScore; keptBandwidth = bandwidth; keptCorrect = nm.correctLogOdds; keptFalse = nm.falseLogOdds;
							if (slice + 1 >= numSlices) { running = false; go = false; }
						}
					}
				}
				if (go)
				{
					// ---- open the next slice: seed the queue from the previous one (Banded.h:235-277)
					slice++;
					j = slice * 64;
					prevFirst = keptFirst;
					prevN = keptN;
					previousMinScore = keptMin;
					previousQuitScore = keptMin + keptBandwidth;
					if (planes) gc_eq_from_planes(planes, it.planeBit + (uint64_t)j, it.seqLen - j, eq);
					else gc_eq_vector(it.seq, it.seqLen, j, eq);
					const uint32_t* prevKeys = ws.keys + prevFirst;
					const GcItemAux* prevAux = ws.aux + prevFirst;
					for (uint32_t k = 0; k < prevN; k++)
					{
						const GcItemAux a = prevAux[k];
						if (j > 0)
						{
							if (a.minScore > previousQuitScore) continue;
							if (a.linIdx >= 0) { const GcItemAux b = prevAux[a.linIdx]; if (b.endScore < previousQuitScore && b.minScore < previousQuitScore) continue; }
						}
						if (!gc_sheap_push(ws.heap, heapSize, ((uint64_t)a.comp << 32) | (prevKeys[k] & 0x7FFFFFFFu))) { status = GC_OVERFLOW_HEAP; running = false; go = false; break; }
					}
					firstItem = itemsUsed;
					curN = 0;
					sliceMinScore = GC_INT_MAX - bandwidth - 1;
					sliceMinNode = NONE; sliceMinOffset = NONE;
					currentMinScoreAtEndRow = sliceMinScore;
					lastKey = ~0ULL;
					prevCursor = 0;
					prevCursorKey = prevN > 0 ? (((uint64_t)prevAux[0].comp << 32) | (prevKeys[0] & 0x7FFFFFFFu)) : ~0ULL;
					if (heapSize == 0) go = false; // nothing seeded: closed (as an internal error) in the next step
				}
			}
			if (go)
			{
				// ---- pop the next node, merge its incoming columns (BVCommon.h:903-964)
				uint64_t key = gc_sheap_pop(ws.heap, heapSize);
				while (key == lastKey && heapSize > 0) key = gc_sheap_pop(ws.heap, heapSize); // queued again by another in-neighbour or by the previous slice (a third of the pops, r03f)
				if (key != lastKey)
				{
					lastKey = key;
					node = (uint32_t)key;
					nodeComp = (uint32_t)(key >> 32);
					const GcNodeRec rec = g.nodeRec[node];
					const uint32_t* prevKeys = ws.keys + prevFirst;
					const uint32_t* curKeys = ws.keys + firstItem;
					while (prevCursorKey < key)
					{
						prevCursor++;
						prevCursorKey = prevCursor < prevN ? (((uint64_t)ws.aux[prevFirst + prevCursor].comp << 32) | (prevKeys[prevCursor] & 0x7FFFFFFFu)) : ~0ULL;
					}
					const int32_t pi = prevCursorKey == key ? (int32_t)prevCursor : -1;
					uint32_t inCount = rec.inCount, oc = rec.outCount;
					if (inCount == 255) inCount = g.inStart[node + 1] - g.inStart[node];
					if (oc == 255) oc = g.outStart[node + 1] - g.outStart[node];
					outBegin = rec.outStart; outCount = oc;
					if (oc > 0) outKey0 = g.outKey[outBegin];
					if (oc > 1) outKey1 = g.outKey[outBegin + 1];
					if (oc > 2) outKey2 = g.outKey[outBegin + 2];
					len = rec.len;
					// in-neighbours: where they sit in the slice being filled (the first two looked up and loaded together)
					const uint32_t in1 = inCount > 1 ? g.inNbr[rec.inStart + 1] : 0;
					const int32_t ci0 = inCount > 0 ? gc_find_key_near(curKeys, curN, rec.firstIn, curN) : -1;
					const int32_t ci1 = inCount > 1 ? gc_find_key_near(curKeys, curN, in1, curN) : -1;
					const bool use0 = ci0 >= 0 && (curKeys[ci0] & 0x80000000u), use1 = ci1 >= 0 && (curKeys[ci1] & 0x80000000u);
					GcWord inc0, inc1; inc0.VP = inc0.VN = 0; inc0.scoreEnd = 0; inc1 = inc0;
					if (use0) inc0 = gc_item_end(ws.items[firstItem + ci0]);
					if (use1) inc1 = gc_item_end(ws.items[firstItem + ci1]);
					nodeLinIdx = rec.linearizable ? ci0 : -2;
					const bool prevExists = pi >= 0;
					int32_t prevStart = 0; uint64_t prevHP = ~0ULL, prevHN = 0;
					bool seeded = false;
					if (prevExists)
					{
						const GcNodeItem& p = ws.items[prevFirst + pi];
						prevStart = p.startScore; prevHP = p.HP; prevHN = p.HN;
						seeded = true;
						if (j > 0)
						{
							const GcItemAux pa = ws.aux[prevFirst + pi];
							if (pa.minScore > previousQuitScore) seeded = false;
							else if (pa.linIdx >= 0) { const GcItemAux b = ws.aux[prevFirst + pa.linIdx]; if (b.endScore < previousQuitScore && b.minScore < previousQuitScore) seeded = false; }
						}
					}
					bool hasWs = false;
					w.VP = 0; w.VN = 0; w.scoreEnd = 0;
					if (seeded)
					{
						w.VP = ~0ULL; w.VN = 0; w.scoreEnd = prevStart + 64; // getSourceSliceFromScore
						hasWs = true;
					}
					const uint64_t Eq0 = gc_sel4(eq, (int)(rec.seq0 & 3));
					for (uint32_t e = 0; e < inCount; e++)
					{
						GcWord inc;
						if (e == 0) { if (!use0) continue; inc = inc0; }
						else if (e == 1) { if (!use1) continue; inc = inc1; }
						else
						{
							int32_t ci = gc_find_key_near(curKeys, curN, g.inNbr[rec.inStart + e], curN);
							if (ci < 0 || !(curKeys[ci] & 0x80000000u)) continue;
							inc = gc_item_end(ws.items[firstItem + ci]);
						}
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
					if (!hasWs) { status = GC_INTERNAL; running = false; }
					else
					{
						if (prevExists && gc_sbs(w) > prevStart)
						{
							GcWord src2; src2.VP = ~0ULL; src2.VN = 0; src2.scoreEnd = prevStart + 64;
							w = gc_merge(w, src2);
						}
						if (itemsUsed >= ws.itemCap) { status = GC_OVERFLOW_ITEMS; running = false; }
						else
						{
							columns += len;
							pending = true;
							if (len > 1)
							{
								forceUntil = gc_cols_prepare_seq(rec.seq0, rec.seq1, len, eq, w, prevExists, prevStart, prevHP, prevHN, 0, run);
								needCols = true;
							}
							else { endW = w; HP = 0; HN = 0; nodeMin = w.scoreEnd; nodeMinOffset = 0; } // a single-column node has no column steps
						}
					}
				}
			}
		}
		// ================= stage 2: the column steps of the nodes in hand, one loop for the warp =================
		gc_k1s_columns<false, false>(run, needCols, len, forceUntil, nullptr);
		if (needCols)
		{
			gc_cols_finish(run, len);
			endW = run.ws; HP = run.HP; HN = run.HN; nodeMin = run.minScore; nodeMinOffset = run.minOffset;
			needCols = false;
		}
		// ================= stage 3: lanes whose item ended in a partial last slice =================
		// flattenLastSliceEnd (BVCommon.h:1171-1229), then close the slice.  The reference recomputes every node of the slice,
		// flattens each column (flattenWordSlice, BVCommon.h:265-273) and keeps the first strict minimum in the iteration order
		// of its phmap node map; the column run tracks exactly that minimum, so no column is stored.
		const bool wantFlat = have && !running && needFlatten && status == GC_OK;
		if (GC_WARP_ANY(wantFlat))
		{
			bool flat = wantFlat;
			uint32_t capacity = 0, slot = 0;
			uint64_t flatMask = 0;
			if (flat)
			{
				uint32_t rows = (uint32_t)(it.seqLen - j);
				flatMask = ~0ULL << rows;
				GcItemNodeKey keyFn; keyFn.items = ws.items + firstItem;
				if (!gc_phmap_order_keys(keyFn, curN, prevN, ws.scratch, ws.scratchCap, &capacity)) { status = GC_OVERFLOW_HEAP; flat = false; }
				sliceMinScore = GC_INT_MAX;
				sliceMinNode = NONE;
				sliceMinOffset = NONE;
			}
			bool flatCols = false;
			while (GC_WARP_ANY(flat))
			{
				// next occupied slot; one-column nodes are finished here
				if (flat)
				{
					while (slot < capacity && ws.scratch[slot] == NONE) slot++;
					if (slot >= capacity) flat = false;
				}
				if (flat)
				{
					const GcNodeItem& fi = ws.items[firstItem + ws.scratch[slot]];
					slot++;
					node = gc_item_node(fi);
					len = g.nodeLength[node];
					columns += len;
					const int32_t pi = gc_find_key(ws.keys + prevFirst, prevN, node);
					const bool prevExists = pi >= 0;
					int32_t prevStart = 0; uint64_t prevHP = ~0ULL, prevHN = 0;
					if (prevExists) { const GcNodeItem& p = ws.items[prevFirst + pi]; prevStart = p.startScore; prevHP = p.HP; prevHN = p.HN; }
					GcWord sw = gc_item_start(fi);
					if (prevExists && gc_sbs(sw) > prevStart)
					{
						GcWord src2; src2.VP = ~0ULL; src2.VN = 0; src2.scoreEnd = prevStart + 64;
						sw = gc_merge(sw, src2);
					}
					forceUntil = gc_cols_prepare(g, node, len, eq, sw, prevExists, prevStart, prevHP, prevHN, flatMask, run);
					if (len > 1) flatCols = true;
					else if (run.minScore < sliceMinScore) { sliceMinScore = run.minScore; sliceMinNode = node; sliceMinOffset = 0; }
				}
				gc_k1s_columns<true, false>(run, flatCols, len, forceUntil, nullptr);
				if (flatCols)
				{
					if (run.minScore < sliceMinScore) { sliceMinScore = run.minScore; sliceMinNode = node; sliceMinOffset = run.minOffset; }
					flatCols = false;
				}
			}
			if (wantFlat && status == GC_OK)
			{
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
		}
		// ================= the item is over: removeWronglyAlignedEnd (BVCommon.h:1231-1241), hand the result over =================
		if (have && !running)
		{
			GcK1Result res;
			res.score = GC_INT_MAX; res.traceLen = 0;
			res.columns = columns;
			res.status = status;
			res.itemsUsed = itemsUsed;
			int32_t last = 0;
			if (status == GC_OK)
			{
				int32_t count = lastSlice + 1;
				bool currentlyCorrect = ws.slices[count - 1].correctLogOdds > ws.slices[count - 1].falseLogOdds;
				while (!currentlyCorrect)
				{
					currentlyCorrect = ws.slices[count - 1].falseFromCorrect;
					count--;
					if (count == 0) break;
				}
				last = count - 1;
			}
			src.done(res, last);
			have = false;
		}
	}
}

// one item, a warp of one lane: the form the host-side simulations (tests/hostsim) call
struct GcK1SSingle
{
	GcK1SItem item; GcK1SWorkspace slab; bool taken; GcK1Result res; int32_t last;
	GC_HD bool next(GcK1SItem& it, GcK1SWorkspace& ws) { if (taken) return false; taken = true; it = item; ws = slab; return true; }
	GC_HD void done(const GcK1Result& r, int32_t l) { res = r; last = l; }
};
GC_HD int32_t gc_k1s_forward(const GcGraphView& g, const GcViterbiTables& vt, const GcK1Params& prm, bool have, const uint8_t* seq, int32_t seqLen, uint32_t startNode, uint32_t startOffset,
	const uint64_t* planes, uint64_t planeBit, GcK1SWorkspace& ws, GcK1Result& res)
{
	GcK1SSingle one;
	one.item.seq = seq; one.item.seqLen = seqLen; one.item.startNode = startNode; one.item.startOffset = startOffset; one.item.planeBit = planeBit;
	one.slab = ws; one.taken = !have; one.last = 0;
	one.res.status = GC_OK; one.res.score = GC_INT_MAX; one.res.traceLen = 0; one.res.itemsUsed = 0; one.res.columns = 0;
	GcK1SWorkspace lane = ws;
	gc_k1s_forward_items(g, vt, prm, planes, lane, one);
	res.status = one.res.status; res.columns = one.res.columns; res.itemsUsed = one.res.itemsUsed;
	return one.last;
}

// ------------------------------------------------------------------------------------
// Backtrace (BVCommon.h:392-544), lane-per-item twin of gc_k1_backtrace: same cells, same preferences, same trace.
// Per path node: stage 1 = the crossing out of the previous node (pickBacktrace*, BVCommon.h:556-804) and the lookup of the
// next one, repeated through one-column nodes; stage 2 = ONE column loop for the warp that recomputes the nodes in hand
// (recalcNodeWordslice, BVCommon.h:828-852; columns in per-lane local memory); stage 3 = ONE cell-walk loop for the warp.
// `cols` = 64 columns of per-lane scratch.  `last` = index of the last kept slice (>= 1) of a lane with `have`.
// A lane works through items one after the other (as in the forward pass): src.next() hands out the items whose forward pass
// succeeded -- (item, slab, last kept slice >= 1, trace buffer, the forward result) -- and src.done() takes the finished result.
template <class Source>
GC_HD void gc_k1s_backtrace_items(const GcGraphView& g, const uint64_t* planes, GcK1SWorkspace& ws, GcColVV* cols, Source& src)
{
	bool have = false, exhausted = false, active = false;
	GcK1SItem it; it.seq = nullptr; it.seqLen = 0; it.startNode = 0; it.startOffset = 0; it.planeBit = 0;
	GcK1Result res; res.status = GC_OK; res.score = 0; res.traceLen = 0; res.itemsUsed = 0; res.columns = 0;
	GcTraceWriter tw;
	tw.out = nullptr; tw.cap = 0; tw.n = 0; tw.overflow = false; tw.node = 0; tw.offset = 0; tw.seqPos = -1;
	uint32_t currentNode = 0xFFFFFFFFu;
	int32_t currentSlice = -1;
	uint64_t eq[4] = { 0, 0, 0, 0 };
	uint32_t guard = 0;
	uint32_t guardMax = 0;
	uint64_t columns = 0;
	// the slice pair the walk is in
	uint32_t curFirst = 0, curN = 0, prevFirst = 0, prevN = 0;
	uint32_t curHint = 0, prevHint = 0; // where the last lookups in the two slices hit
	int32_t quitScore = 0, previousQuitScore = 0, j = 0;
	// stage 2 / 3 state
	bool needCols = false, walking = false;
	uint32_t len = 0, forceUntil = 0;
	GcColumnRun run;
	run.ws.VP = run.ws.VN = 0; run.ws.scoreEnd = 0; run.eq[0] = run.eq[1] = run.eq[2] = run.eq[3] = 0; run.prevHP = run.prevHN = run.HP = run.HN = 0; run.chunk0 = run.chunk1 = 0; run.minScore = 0; run.minOffset = 0; run.flatMask = 0;
	uint64_t chunk0 = 0, chunk1 = 0;
	GcCols cv; cv.c = cols; cv.HP = 0; cv.HN = 0; cv.score0 = 0;
	while (true)
	{
		// ================= stage 0: a lane without an item takes the next one =================
		if (!have && !exhausted)
		{
			int32_t last = 0; uint64_t* traceOut = nullptr; uint32_t traceCap = 0;
			if (!src.next(it, ws, last, traceOut, traceCap, res)) exhausted = true;
			else
			{
				have = true; active = true;
				tw.out = traceOut; tw.cap = traceCap; tw.n = 0; tw.overflow = false; tw.node = 0; tw.offset = 0; tw.seqPos = -1;
				const GcSliceMeta& lm = ws.slices[last];
				res.score = lm.minScore;
				int32_t sp = (last - 1) * 64 + 63;
				if (sp > it.seqLen - 1) sp = it.seqLen - 1;
				tw.push(lm.minScoreNode, lm.minScoreNodeOffset, sp, false);
				currentNode = 0xFFFFFFFFu;
				currentSlice = -1;
				guard = 0;
				guardMax = (uint32_t)it.seqLen * 4 + 1024 + traceCap;
				columns = 0;
				needCols = false; walking = false;
			}
		}
		if (!GC_WARP_ANY(have)) break; // no lane has an item and none could get one
		// ================= stage 1: one step of every lane that is neither waiting for its columns nor walking =================
		// (a) the crossing out of the node in hand (its columns are in `cols`), (b) lookup of the node the trace is in now
		if (active && !needCols && !walking)
		{
			// the walk moves against the edges and the items of a slice lie in topological order: what is looked up sits at or just before the last hit
			auto findCur = [&](uint32_t nd) -> const GcNodeItem* { int32_t i = gc_find_key_near(ws.keys + curFirst, curN, nd, curHint); if (i < 0) return nullptr; curHint = (uint32_t)i; return &ws.items[curFirst + i]; };
			auto findPrev = [&](uint32_t nd) -> const GcNodeItem* { int32_t i = gc_find_key_near(ws.keys + prevFirst, prevN, nd, prevHint); if (i < 0) return nullptr; prevHint = (uint32_t)i; return &ws.items[prevFirst + i]; };
			do
			{
				if (++guard > guardMax) { res.status = GC_INTERNAL; active = false; break; }
				if (currentSlice >= 0 && tw.seqPos / 64 + 1 == currentSlice && tw.node == currentNode)
				{
					if ((tw.seqPos & 63) != 0 && tw.offset != 0) { walking = true; break; }
					if ((tw.seqPos & 63) == 0 && tw.offset == 0)
					{
						GcBtPos bt;
						if (!gc_bt_corner_with(g, findCur, findPrev, currentNode, j, it.seq, quitScore, previousQuitScore, bt)) { res.status = GC_INTERNAL; active = false; break; }
						tw.push(bt.node, bt.offset, bt.seqPos, bt.nodeSwitch);
					}
					else if ((tw.seqPos & 63) == 0)
					{
						// vertical crossing (BVCommon.h:451-477, 665-708)
						const GcNodeItem* pme = findPrev(currentNode);
						if (!pme) tw.push(currentNode, 0, tw.seqPos, false);
						else
						{
							uint32_t off = tw.offset;
							int32_t sp = tw.seqPos;
							uint32_t origOff = off;
							while (off > 0 && gc_value(gc_cols_get(cv, off - 1), 0) == gc_value(gc_cols_get(cv, off), 0) - 1) off--;
							GcBtPos second;
							if (off == 0)
							{
								if (!gc_bt_corner_with(g, findCur, findPrev, currentNode, j, it.seq, quitScore, previousQuitScore, second)) { res.status = GC_INTERNAL; active = false; break; }
							}
							else
							{
								int base = (int)(((off < 32 ? chunk0 : chunk1) >> ((off & 31) * 2)) & 3);
								bool eqc = gc_char_match(it.seq[sp], base);
								int32_t scoreHere = gc_value(gc_cols_get(cv, off), 0);
								// previous slice's last row: startScore + horizontal deltas of columns 1..off-1 (diagonal) and 1..off (up)
								uint64_t below = (1ULL << off) - 2; // bits 1..off-1
								int32_t scoreDiagonal = pme->startScore + gc_popc(pme->HP & below) - gc_popc(pme->HN & below);
								int32_t scoreUp = scoreDiagonal + (int32_t)((pme->HP >> off) & 1) - (int32_t)((pme->HN >> off) & 1);
								second.node = currentNode; second.seqPos = sp - 1; second.nodeSwitch = false;
								if (scoreHere > quitScore || scoreDiagonal > previousQuitScore || scoreUp > previousQuitScore)
								{
									second.offset = (scoreDiagonal < scoreUp) ? off - 1 : off;
								}
								else if (scoreUp == scoreHere - 1) second.offset = off;
								else
								{
									if (scoreDiagonal != scoreHere - (eqc ? 0 : 1)) { res.status = GC_INTERNAL; active = false; break; }
									second.offset = off - 1;
								}
							}
							if (off != origOff)
							{
								for (uint32_t o = origOff - 1; o != off; o--) tw.push(currentNode, o, sp, false);
							}
							if (off != tw.offset || sp != tw.seqPos) tw.push(currentNode, off, sp, false);
							tw.push(second.node, second.offset, second.seqPos, second.nodeSwitch);
						}
					}
					else
					{
						// tw.offset == 0: horizontal crossing (BVCommon.h:478-499, 599-663)
						const GcNodeItem* me = findCur(currentNode);
						GcWord startSlice = gc_item_start(*me);
						int32_t sp = tw.seqPos;
						int32_t origSp = sp;
						while ((sp & 63) != 0 && (startSlice.VP & (1ULL << (sp & 63)))) sp--;
						int32_t offset = sp & 63;
						GcBtPos second;
						if (offset == 0)
						{
							if (!gc_bt_corner_with(g, findCur, findPrev, currentNode, j, it.seq, quitScore, previousQuitScore, second)) { res.status = GC_INTERNAL; active = false; break; }
						}
						else
						{
							bool eqc = gc_char_match(it.seq[sp], (int)(chunk0 & 3));
							int32_t scoreHere = gc_value(startSlice, offset);
							bool found = false;
							if (scoreHere > quitScore)
							{
								int32_t smallestFound = gc_value(startSlice, offset - 1);
								second.node = currentNode; second.offset = 0; second.seqPos = sp - 1; second.nodeSwitch = false;
								for (uint32_t e = g.inStart[currentNode]; e < g.inStart[currentNode + 1]; e++)
								{
									uint32_t nb = g.inNbr[e];
									const GcNodeItem* cn = findCur(nb);
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
									const GcNodeItem* cn = findCur(nb);
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
							if (!found) { res.status = GC_INTERNAL; active = false; break; }
						}
						if (sp != origSp)
						{
							for (int32_t s2 = origSp - 1; s2 != sp; s2--) tw.push(currentNode, 0, s2, false);
						}
						if (sp != tw.seqPos) tw.push(currentNode, 0, sp, false);
						tw.push(second.node, second.offset, second.seqPos, second.nodeSwitch);
					}
				}
				// ---- (b) the node the trace is in now
				if (tw.seqPos == -1) { active = false; break; }
				const int32_t newSlice = tw.seqPos / 64 + 1;
				const uint32_t newNode = tw.node;
				if (newSlice == currentSlice && newNode == currentNode) break; // the crossing stayed in the node
				if (newSlice != currentSlice)
				{
					const GcSliceMeta& cm = ws.slices[newSlice];
					const GcSliceMeta& pmeta = ws.slices[newSlice - 1];
					curHint = (newSlice == currentSlice - 1) ? prevHint : cm.numItems; // one slice down: the old previous slice is the current one now
					prevHint = pmeta.numItems;
					curFirst = cm.firstItem; curN = cm.numItems; prevFirst = pmeta.firstItem; prevN = pmeta.numItems;
					quitScore = cm.minScore + cm.bandwidth;
					previousQuitScore = pmeta.minScore + pmeta.bandwidth;
					j = (newSlice - 1) * 64;
					if (planes) gc_eq_from_planes(planes, it.planeBit + (uint64_t)j, it.seqLen - j, eq);
					else gc_eq_vector(it.seq, it.seqLen, j, eq);
				}
				currentSlice = newSlice;
				currentNode = newNode;
				const GcNodeItem* me = findCur(currentNode);
				if (!me) { res.status = GC_INTERNAL; active = false; break; }
				const GcNodeItem* pme = findPrev(currentNode);
				const GcNodeRec rec = g.nodeRec[currentNode];
				len = rec.len;
				columns += len;
				GcWord sw = gc_item_start(*me);
				const bool prevExists = pme != nullptr;
				const int32_t prevStart = prevExists ? pme->startScore : 0;
				if (prevExists && gc_sbs(sw) > prevStart)
				{
					GcWord src; src.VP = ~0ULL; src.VN = 0; src.scoreEnd = prevStart + 64;
					sw = gc_merge(sw, src);
				}
				cols[0].VP = sw.VP; cols[0].VN = sw.VN;
				cv.HP = me->HP; cv.HN = me->HN; cv.score0 = sw.scoreEnd;
				chunk0 = rec.seq0; chunk1 = rec.seq1;
				if (len > 1)
				{
					forceUntil = gc_cols_prepare_seq(chunk0, chunk1, len, eq, sw, prevExists, prevStart, prevExists ? pme->HP : ~0ULL, prevExists ? pme->HN : 0ULL, 0, run);
					needCols = true; // the columns (stage 2), then the walk (stage 3), then the crossing in the next step
				}
			} while (false);
		}
		// ================= stage 2: recompute the nodes in hand =================
		{
			gc_k1s_columns<false, true>(run, needCols, len, forceUntil, cols);
			if (needCols)
			{
				needCols = false;
				if ((tw.seqPos & 63) != 0 && tw.offset != 0) walking = true;
			}
		}
		// ================= stage 3: walk inside the node (BVCommon.h:556-597) to its first column or the slice's first row ==========
		// The reference reads three cell values per step (here, above, diagonal: getValue = two 64-bit popcounts each).
		// Same values, incrementally: along a column value(row-1) = value(row) + VN[row] - VP[row], so the cell above is a
		// bit test, and the value of the left neighbour column at the current row is carried along and only recomputed
		// (one getValue) when the walk moves a column to the left.
		if (GC_WARP_ANY(walking))
		{
			uint32_t hori = 0; int32_t vert = 0;
			GcWord cur, left; cur.VP = cur.VN = 0; cur.scoreEnd = 0; left = cur;
			int32_t scoreHere = 0, leftHere = 0;
			if (walking)
			{
				hori = tw.offset; vert = tw.seqPos - j;
				cur = gc_cols_get(cv, hori); left = gc_cols_get(cv, hori - 1);
				scoreHere = gc_value(cur, vert);
				leftHere = gc_value(left, vert);
			}
			while (GC_WARP_ANY(walking))
			{
				if (!walking) continue;
				int32_t dv = (int32_t)((cur.VN >> vert) & 1) - (int32_t)((cur.VP >> vert) & 1);    // value(cur, vert-1) - value(cur, vert)
				int32_t dl = (int32_t)((left.VN >> vert) & 1) - (int32_t)((left.VP >> vert) & 1); // the same in the left column
				int32_t diagonalScore = leftHere + dl;
				int base = (int)(((hori < 32 ? chunk0 : chunk1) >> ((hori & 31) * 2)) & 3);
				bool eqc = (gc_sel4(eq, base) >> vert) & 1; // gc_char_match(seq[vert + j], base): the slice's Eq masks are in registers
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
				if (!(hori > 0 && vert > 0)) walking = false;
			}
		}
		// ================= the walk is over: slide left in row -1 (BVCommon.h:508-542), hand the result over =================
		if (have && !active)
		{
			res.columns += columns;
			if (res.status == GC_OK)
			{
				const GcSliceMeta& m0 = ws.slices[0];
				int32_t ii = gc_find_key(ws.keys + m0.firstItem, m0.numItems, tw.node);
				if (ii < 0) res.status = GC_INTERNAL;
				else
				{
					const GcNodeItem* first = &ws.items[m0.firstItem + ii];
					uint32_t off = tw.offset;
					int32_t here = first->startScore + gc_popc(first->HP & ((2ULL << off) - 2)) - gc_popc(first->HN & ((2ULL << off) - 2));
					while (here != 0 && off > 0)
					{
						int32_t before = here - (int32_t)((first->HP >> off) & 1) + (int32_t)((first->HN >> off) & 1);
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
							int32_t ni = gc_find_key(ws.keys + m0.firstItem, m0.numItems, nb);
							if (ni >= 0 && gc_sbs(gc_item_end(ws.items[m0.firstItem + ni])) == here - 1)
							{
								tw.push(nb, g.nodeLength[nb] - 1, tw.seqPos, true);
								break;
							}
						}
					}
					res.traceLen = tw.n;
					if (tw.overflow) res.status = GC_OVERFLOW_TRACE;
				}
			}
			src.done(res);
			have = false;
		}
	}
}

// one item, a warp of one lane (tests/hostsim)
struct GcK1SSingleBt
{
	GcK1SItem item; GcK1SWorkspace slab; bool taken; int32_t last; uint64_t* traceOut; uint32_t traceCap; GcK1Result res;
	GC_HD bool next(GcK1SItem& it, GcK1SWorkspace& ws, int32_t& l, uint64_t*& out, uint32_t& cap, GcK1Result& r) { if (taken) return false; taken = true; it = item; ws = slab; l = last; out = traceOut; cap = traceCap; r = res; return true; }
	GC_HD void done(const GcK1Result& r) { res = r; }
};
GC_HD void gc_k1s_backtrace(const GcGraphView& g, bool have, const uint8_t* seq, int32_t seqLen, const uint64_t* planes, uint64_t planeBit, GcK1SWorkspace& ws, int32_t last, GcColVV* cols, uint64_t* traceOut, uint32_t traceCap, GcK1Result& res)
{
	GcK1SSingleBt one;
	one.item.seq = seq; one.item.seqLen = seqLen; one.item.startNode = 0; one.item.startOffset = 0; one.item.planeBit = planeBit;
	one.slab = ws; one.taken = !have; one.last = last; one.traceOut = traceOut; one.traceCap = traceCap; one.res = res;
	GcK1SWorkspace lane = ws;
	gc_k1s_backtrace_items(g, planes, lane, cols, one);
	res = one.res;
}
