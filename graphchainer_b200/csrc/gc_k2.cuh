// K2 -- co-linear chaining of one read's anchors over the minimum-path-cover index:
// AlignmentGraph::colinearChaining / colinearChainingByComponent
// (src/AlignmentGraph.cpp:1712-1735, 1737-1863).
//
// The reference sweeps anchor end-points in topological order and keeps, per MPC path k,
// two treaps keyed by the anchor's read end y (values (C, j) and (C - y, j)); an anchor j
// starting at node u queries, for every backward link (v, k) of u, the anchors already
// inserted under path k, i.e. the anchors whose END node lies on path k at or before v.
// All updates are `max` over (score, anchor index) pairs, so the result is the closed form
//
//   C[j] = max( (len_j, -1),  max over i in pred(j) of (val(i, j), i) )          (lexicographic)
//   pred(j) = { i : y_i < y_j,  end(i) == start(j)  or  end(i) reaches start(j) through a link }
//   val     = len_j + C[i]               if y_i <= x_j - 1      (T[k].RMQ(0, x-1),   :1841-1843)
//           = y_j - y_i + C[i]           if x_j <= y_i          (I[k].RMQ(x, y-1),   :1844-1846)
//   reach   : exists (v, k) in backwards[start(j)] with k in paths[end(i)] and
//             topo(end(i)) <= topo(v)                            (:1828-1832 before :1834-1846)
//   same node: the (y, x)-ordered temporary treaps of :1789-1826
//
// with ties resolved towards the larger anchor index, the best chain end = max (C[j], j)
// (:1848-1850), components visited in ascending id and replaced only by a strictly larger
// score (:1728).  Every pred has a strictly smaller y, so evaluating anchors by increasing
// y is a valid order: one block per read, threads stride over the candidate predecessors
// and a block-wide max-reduction picks the winner.
#pragma once
#include "gc_common.cuh"

struct GcMpcView
{
	const uint32_t* compMap;    // [N] component of a node
	const uint32_t* compIdx;    // [N] index of the node inside its component
	const uint32_t* compStart;  // [C+1]
	const uint32_t* topoIds;    // [N] at compStart[c] + idx
	const uint32_t* pathsStart; // [N+1] CSR at compStart[c] + idx
	const uint32_t* pathsK;
	const uint32_t* backStart;  // [N+1]
	const uint32_t* backNode;   // component-local node index
	const uint32_t* backK;
	const uint32_t* pathBase;   // [C+1] first global path id of every component (device only: derived at gcgpu_create)
};

struct GcAnchor
{
	uint32_t startNode; // Anchor::path[0]
	uint32_t endNode;   // Anchor::path.back()
	int32_t x, y;       // fragment bounds in the read (Aligner.cpp:707)
};

// does an anchor ending at node `e` precede one starting at node `s` (same component)?
GC_HD bool gc_k2_reaches(const GcMpcView& m, uint32_t e, uint32_t s)
{
	if (e == s) return true;
	uint32_t c = m.compMap[s];
	uint32_t base = m.compStart[c];
	uint32_t ge = base + m.compIdx[e], gs = base + m.compIdx[s];
	uint32_t topoE = m.topoIds[ge];
	for (uint32_t b = m.backStart[gs]; b < m.backStart[gs + 1]; b++)
	{
		if (topoE > m.topoIds[base + m.backNode[b]]) continue;
		uint32_t k = m.backK[b];
		for (uint32_t p = m.pathsStart[ge]; p < m.pathsStart[ge + 1]; p++)
			if (m.pathsK[p] == k) return true;
	}
	return false;
}

GC_HD int64_t gc_k2_key(int32_t score, int32_t idx) { return ((int64_t)score << 32) | (uint32_t)(idx + 1); }

// candidate key of predecessor i for anchor j (both in the same component), or INT64_MIN
GC_HD int64_t gc_k2_candidate(const GcMpcView& m, const GcAnchor& ai, const GcAnchor& aj, uint32_t i, int32_t scoreI)
{
	if (!gc_k2_reaches(m, ai.endNode, aj.startNode)) return (int64_t)0x8000000000000000LL;
	int32_t len = aj.y - aj.x + 1;
	int32_t val = (ai.y <= aj.x - 1) ? len + scoreI : aj.y - ai.y + scoreI;
	return gc_k2_key(val, (int32_t)i);
}

GC_HD uint32_t gc_k2_select(const GcMpcView& m, const GcAnchor* a, uint32_t n, const int32_t* score, const int32_t* pred, uint32_t* chainOut, int64_t* bestScoreOut);

// Sequential form (host checks, and the single-thread tail of the kernel).  order[] = anchor
// indices sorted by (y, index); score[]/pred[] are outputs; chainOut receives the chain
// (anchor indices in read order), returns its length.  bestScore = covered read bases.
GC_HD uint32_t gc_k2_chain_seq(const GcMpcView& m, const GcAnchor* a, uint32_t n, const uint32_t* order, int32_t* score, int32_t* pred, uint32_t* chainOut, int64_t* bestScoreOut)
{
	for (uint32_t oj = 0; oj < n; oj++)
	{
		uint32_t j = order[oj];
		int32_t len = a[j].y - a[j].x + 1;
		int64_t best = gc_k2_key(len, -1);
		uint32_t cj = m.compMap[a[j].endNode];
		for (uint32_t oi = 0; oi < oj; oi++)
		{
			uint32_t i = order[oi];
			if (a[i].y >= a[j].y) break;
			if (m.compMap[a[i].endNode] != cj) continue;
			if (!gc_k2_reaches(m, a[i].endNode, a[j].startNode)) continue;
			int32_t val = (a[i].y <= a[j].x - 1) ? len + score[i] : a[j].y - a[i].y + score[i];
			int64_t key = gc_k2_key(val, (int32_t)i);
			if (key > best) best = key;
		}
		score[j] = (int32_t)(best >> 32);
		pred[j] = (int32_t)(uint32_t)(best & 0xFFFFFFFFu) - 1;
	}
	return gc_k2_select(m, a, n, score, pred, chainOut, bestScoreOut);
}

// best chain end + backtrack (AlignmentGraph.cpp:1848-1862 and :1717-1733)
GC_HD uint32_t gc_k2_select(const GcMpcView& m, const GcAnchor* a, uint32_t n, const int32_t* score, const int32_t* pred, uint32_t* chainOut, int64_t* bestScoreOut)
{
	// per component max (C, j); first component (ascending id) with a strictly larger score
	int64_t bestKey = -1; uint32_t bestComp = 0xFFFFFFFFu; bool first = true;
	// components in ascending order: scan for the smallest unseen component id repeatedly (n is small)
	uint32_t lastComp = 0; bool haveLast = false;
	while (true)
	{
		uint32_t c = 0xFFFFFFFFu;
		for (uint32_t j = 0; j < n; j++)
		{
			uint32_t cj = m.compMap[a[j].endNode];
			if ((!haveLast || cj > lastComp) && cj < c) c = cj;
		}
		if (c == 0xFFFFFFFFu) break;
		int64_t key = gc_k2_key(0, -1);
		for (uint32_t j = 0; j < n; j++) if (m.compMap[a[j].endNode] == c) { int64_t kj = gc_k2_key(score[j], (int32_t)j); if (kj > key) key = kj; }
		if (first || (key >> 32) > (bestKey >> 32)) { first = false; bestKey = key; bestComp = c; }
		lastComp = c; haveLast = true;
	}
	(void)bestComp;
	uint32_t len = 0;
	if (first) { *bestScoreOut = 0; return 0; }
	*bestScoreOut = bestKey >> 32;
	int32_t cur = (int32_t)(uint32_t)(bestKey & 0xFFFFFFFFu) - 1;
	while (cur != -1) { chainOut[len++] = (uint32_t)cur; cur = pred[cur]; }
	for (uint32_t x = 0, y2 = len; x + 1 < y2; x++, y2--) { uint32_t t = chainOut[x]; chainOut[x] = chainOut[y2 - 1]; chainOut[y2 - 1] = t; }
	return len;
}
