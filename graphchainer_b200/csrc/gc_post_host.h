// Host-side preparation of the tables gc_post.cuh reads: done once per context by gcgpu_create (and by the C-ABI
// test double under tests/hostsim), from the reference-shaped arrays of gcgpu_graph.
#pragma once
#include <cstdint>
#include <vector>
#include "../../include/gcgpu.h"
#include "gc_host_graph.h"

// character -> code of the resident sequence buffer: IUPAC mask (bits 0-3, Common::ambiguousMatch GraphAlignerCommon.h:219-296),
// bit 4 = U/u (matches T in the DP but is not a seeding base, MinimizerSeeder.cpp:24-43), bit 5 = not one of the upper-case
// letters A C G T (edlib compares bytes, so only those can equal a graph base in K3)
inline void gcBuildCodeTable(uint8_t* table)
{
	for (int c = 0; c < 256; c++)
	{
		uint8_t m = gcEncodeSeedBase((char)c);
		if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) m |= 32;
		table[c] = m;
	}
}
inline uint8_t gcComplementCode(uint8_t m) { return (uint8_t)(((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3) | (m & 0x30)); }
inline uint8_t gcK3CodeOf(uint8_t m) { return (m & 0x30) ? (uint8_t)4 : (m == 1 ? (uint8_t)0 : m == 2 ? (uint8_t)1 : m == 4 ? (uint8_t)2 : m == 8 ? (uint8_t)3 : (uint8_t)4); }

// per split node: the split nodes of the reverse-strand original node and the reverse offset of the node's first base
// (AlignmentGraph::GetReversePosition, AlignmentGraph.cpp:850-868)
struct GcRevTables { std::vector<uint32_t> revFirst, revCount, revLast; };
inline bool gcBuildRevTables(const gcgpu_graph* g, GcRevTables& t)
{
	uint32_t N = g->num_nodes;
	int32_t maxId = -1;
	for (uint32_t o = 0; o < g->num_orig; o++) if (g->orig_ids[o] > maxId) maxId = g->orig_ids[o];
	std::vector<int32_t> indexOfId((size_t)maxId + 2, -1);
	for (uint32_t o = 0; o < g->num_orig; o++) if (g->orig_ids[o] >= 0) indexOfId[g->orig_ids[o]] = (int32_t)o;
	t.revFirst.assign(N, 0); t.revCount.assign(N, 0); t.revLast.assign(N, 0);
	for (uint32_t o = 0; o < g->num_orig; o++)
	{
		int32_t rid = g->orig_ids[o] ^ 1;
		int32_t ro = (rid >= 0 && (size_t)rid < indexOfId.size()) ? indexOfId[rid] : -1;
		for (uint32_t k = g->orig_start[o]; k < g->orig_start[o + 1]; k++)
		{
			uint32_t n = g->orig_nodes[k];
			if (n >= N) return false;
			if (ro >= 0) { t.revFirst[n] = g->orig_start[ro]; t.revCount[n] = g->orig_start[ro + 1] - g->orig_start[ro]; }
			t.revLast[n] = g->orig_size[o] - 1 - g->node_offset[n];
		}
	}
	return true;
}

// fills the original-node members of a gcgpu_graph from the host graph
inline void gcFillOrigArrays(const GcHostGraph& g, gcgpu_graph& gg)
{
	gg.node_ids = g.nodeIDs.data(); gg.node_offset = g.nodeOffset.data();
	gg.num_orig = (uint32_t)g.origIds.size(); gg.orig_ids = g.origIds.data(); gg.orig_start = g.origStart.data(); gg.orig_nodes = g.origNodes.data(); gg.orig_size = g.origSize.data();
}
