// Index build (once per run, host): GFA -> split-node alignment graph -> component order,
// chains, minimum-path-cover index, minimizer index, as flat arrays in REFERENCE numbering.
//
// The per-read path is bit-exact only if node numbers, neighbour orders, component numbers,
// chain positions and minimizer position lists equal the reference's, and those follow from
// libstdc++ container iteration orders.  So this builder uses the same
// standard containers with the same insertion sequences as the reference does:
//   GfaGraph::LoadFromStream          src/GfaGraph.cpp:212-370     (names -> ids by first appearance)
//   DirectedGraph::BuildFromGFA       src/BigraphToDigraph.cpp:215-267
//   AlignmentGraph::AddNode/AddEdge   src/AlignmentGraph.cpp:50-253
//   findLinearizable                  src/AlignmentGraph.cpp:644-736 (see note: always all-false)
//   doComponentOrder                  src/AlignmentGraph.cpp:1008-1115 (Tarjan)
//   findChains & helpers              src/AlignmentGraph.cpp:309-642
//   buildMPC & helpers                src/AlignmentGraph.cpp:1157-1489
//   MinimizerSeeder (one bucket)      src/MinimizerSeeder.cpp:104-200, 286-492, 557-575
// tests/test_builder.py diffs every array against a dump of the reference's own structures.
// Graphs with ambiguous bases or edge overlaps are rejected (outside the hot path's scope).
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <cstring>
#include <zlib.h>
#include "gc_index.h"

namespace gcbuild {

struct NodePos
{
	int id; bool end;
	bool operator==(const NodePos& o) const { return id == o.id && end == o.end; }
};
struct NodePosHash { size_t operator()(const NodePos& x) const { return std::hash<int>()(x.id) ^ std::hash<bool>()(x.end); } }; // GfaGraph.h:23-32

struct SplitGraph
{
	std::vector<size_t> nodeLength, nodeOffset;
	std::vector<int> nodeIDs;
	std::vector<std::vector<size_t>> inNeighbors, outNeighbors;
	std::vector<bool> reverse;
	std::vector<uint64_t> nodeSeq; // 2 per node
	std::unordered_map<int, std::vector<size_t>> nodeLookup;
	std::unordered_map<int, size_t> originalNodeSize;
	std::unordered_map<int, std::string> originalNodeName;
	std::vector<size_t> componentNumber, chainNumber, chainApproxPos;
	size_t bpSize = 0;
	// MPC
	std::vector<size_t> component_map, component_idx;
	std::vector<std::vector<size_t>> component_ids, topo_ids;
	std::vector<std::vector<std::vector<size_t>>> mpc, paths;
	std::vector<std::vector<std::vector<std::pair<size_t, size_t>>>> backwards;
	size_t size() const { return nodeLength.size(); }
	char base(size_t node, size_t pos) const { return "ACGT"[(nodeSeq[2 * node + (pos >> 5)] >> ((pos & 31) * 2)) & 3]; }
	size_t unitigNode(int nodeId, size_t offset) const
	{
		const auto& nodes = nodeLookup.at(nodeId);
		size_t index = (size_t)(nodes.size() * ((double)offset / (double)originalNodeSize.at(nodeId)));
		if (index >= nodes.size()) index = nodes.size() - 1;
		while (index < nodes.size() - 1 && (nodeOffset[nodes[index]] + nodeLength[nodes[index]] <= offset)) index++;
		while (index > 0 && (nodeOffset[nodes[index]] > offset)) index--;
		return nodes[index];
	}
};

inline char complement(char c)
{
	switch (c) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; }
	throw std::runtime_error(std::string("Invalid sequence character: ") + c + " (only A C G T graphs are supported by the B200 path)");
}

inline void addSplitNode(SplitGraph& g, int nodeId, size_t offset, const std::string& seq, bool reverseNode)
{
	g.bpSize += seq.size();
	g.nodeLookup[nodeId].push_back(g.nodeLength.size());
	g.nodeLength.push_back(seq.size());
	g.nodeIDs.push_back(nodeId);
	g.inNeighbors.emplace_back();
	g.outNeighbors.emplace_back();
	g.reverse.push_back(reverseNode);
	g.nodeOffset.push_back(offset);
	uint64_t chunks[2] = { 0, 0 };
	for (size_t i = 0; i < seq.size(); i++)
	{
		uint64_t v;
		switch (seq[i]) { case 'a': case 'A': v = 0; break; case 'c': case 'C': v = 1; break; case 'g': case 'G': v = 2; break; case 't': case 'T': case 'u': case 'U': v = 3; break;
			default: throw std::runtime_error(std::string("Invalid sequence character: ") + seq[i] + " (only A C G T graphs are supported by the B200 path)"); }
		chunks[i >> 5] |= v << ((i & 31) * 2);
	}
	g.nodeSeq.push_back(chunks[0]);
	g.nodeSeq.push_back(chunks[1]);
}

// AlignmentGraph::AddNode (AlignmentGraph.cpp:50-85) with breakpoints {0, size}
inline void addNode(SplitGraph& g, int nodeId, const std::string& sequence, const std::string& name, bool reverseNode)
{
	if (g.nodeLookup.count(nodeId) != 0) return;
	g.originalNodeSize[nodeId] = sequence.size();
	g.originalNodeName[nodeId] = name;
	for (size_t offset = 0; offset < sequence.size(); offset += 64)
	{
		size_t size = 64;
		if (sequence.size() - offset < size) size = sequence.size() - offset;
		addSplitNode(g, nodeId, offset, sequence.substr(offset, size), reverseNode);
		if (offset > 0)
		{
			size_t n = g.outNeighbors.size();
			g.outNeighbors[n - 2].push_back(n - 1);
			g.inNeighbors[n - 1].push_back(n - 2);
		}
	}
}

// AlignmentGraph::AddEdgeNodeId (AlignmentGraph.cpp:233-253), startOffset = overlap = 0
inline void addEdge(SplitGraph& g, int from_id, int to_id)
{
	size_t from = g.nodeLookup.at(from_id).back();
	size_t to = std::numeric_limits<size_t>::max();
	for (auto node : g.nodeLookup[to_id]) if (g.nodeOffset[node] == 0) to = node;
	if (to == std::numeric_limits<size_t>::max()) throw std::runtime_error("edge to a missing node");
	if (std::find(g.inNeighbors[to].begin(), g.inNeighbors[to].end(), from) == g.inNeighbors[to].end()) g.inNeighbors[to].push_back(from);
	if (std::find(g.outNeighbors[from].begin(), g.outNeighbors[from].end(), to) == g.outNeighbors[from].end()) g.outNeighbors[from].push_back(to);
}

inline SplitGraph loadGfa(const std::string& filename)
{
	std::ifstream file(filename);
	if (!file.good()) throw std::runtime_error("No graph file exists");
	std::unordered_map<std::string, int> nameMapping;
	std::unordered_map<int, std::string> nodes;
	std::unordered_map<NodePos, std::vector<NodePos>, NodePosHash> edges;
	auto getNameId = [&nameMapping](const std::string& name) { auto f = nameMapping.find(name); if (f == nameMapping.end()) { int r = (int)nameMapping.size(); nameMapping[name] = r; return r; } return f->second; };
	// GFA 1 records, whitespace-separated fields: "S name sequence ..." and "L from +|- to +|- overlap ..."; everything else is skipped.
	// Segment ids = order of first appearance of a NAME in any S or L record (GfaGraph.cpp:164-174): the numbering of the whole index.
	std::string line;
	std::vector<std::pair<size_t, size_t>> field; // (begin, length) of the first six fields of the line
	auto splitFields = [&field](const std::string& text)
	{
		field.clear();
		size_t i = 0, n = text.size();
		while (i < n && field.size() < 6)
		{
			while (i < n && (text[i] == ' ' || text[i] == '\t' || text[i] == '\r')) i++;
			size_t b = i;
			while (i < n && text[i] != ' ' && text[i] != '\t' && text[i] != '\r') i++;
			if (i > b) field.emplace_back(b, i - b);
		}
	};
	auto fieldText = [&](size_t k) { return k < field.size() ? line.substr(field[k].first, field[k].second) : std::string(); };
	while (std::getline(file, line))
	{
		if (line.empty() || (line[0] != 'S' && line[0] != 'L')) continue;
		splitFields(line);
		if (line[0] == 'S')
		{
			std::string name = fieldText(1);
			int id = getNameId(name);
			std::string seq = fieldText(2);
			if (seq == "*") throw std::runtime_error("Nodes without sequence (*) are not currently supported (nodeid " + name + ")");
			nodes[id] = seq;
		}
		else
		{
			int from = getNameId(fieldText(1));
			bool fromForward = fieldText(2) == "+";
			int to = getNameId(fieldText(3));
			bool toForward = fieldText(4) == "+";
			int overlap = atoi(fieldText(5).c_str()); // "0M" -> 0
			if (overlap != 0) throw std::runtime_error("Edge overlaps are not supported by the B200 path (only 0M links)");
			edges[NodePos { from, fromForward }].push_back(NodePos { to, toForward });
		}
	}
	std::unordered_map<int, std::string> originalNodeName;
	for (auto pair : nameMapping) originalNodeName[pair.second] = pair.first;
	// edges from/to non-existent nodes are dropped (GfaGraph.cpp:321-366)
	std::vector<NodePos> nonexistent;
	for (auto& edge : edges)
	{
		if (nodes.count(edge.first.id) == 0) { nonexistent.push_back(edge.first); continue; }
		for (size_t i = edge.second.size() - 1; i < edge.second.size() + 1; i--)
			if (nodes.count(edge.second[i].id) == 0) edge.second.erase(edge.second.begin() + i);
	}
	for (auto n : nonexistent) edges.erase(edges.find(n));
	// DirectedGraph::BuildFromGFA (BigraphToDigraph.cpp:215-267)
	SplitGraph g;
	for (auto node : nodes)
	{
		std::string name = originalNodeName.count(node.first) ? originalNodeName.at(node.first) : std::string();
		std::string rc;
		rc.reserve(node.second.size());
		for (size_t i = node.second.size(); i-- > 0; ) rc += complement(node.second[i]);
		addNode(g, node.first * 2, node.second, name, false);
		addNode(g, node.first * 2 + 1, rc, name, true);
	}
	for (auto edge : edges)
	{
		for (auto target : edge.second)
		{
			// ConvertGFAEdgeToEdges (BigraphToDigraph.cpp:101-132)
			size_t fromLeft, fromRight, toLeft, toRight;
			if (!edge.first.end) { fromLeft = edge.first.id * 2; fromRight = edge.first.id * 2 + 1; } else { fromLeft = edge.first.id * 2 + 1; fromRight = edge.first.id * 2; }
			if (!target.end) { toLeft = target.id * 2; toRight = target.id * 2 + 1; } else { toLeft = target.id * 2 + 1; toRight = target.id * 2; }
			addEdge(g, (int)fromRight, (int)toRight);
			addEdge(g, (int)toLeft, (int)fromLeft);
		}
	}
	return g;
}

// ---- .vg input (DirectedGraph::StreamVGGraphFromFile, BigraphToDigraph.cpp:134-179): a stream of vg::Graph messages in the framing of
// stream.hpp:82-124 -- gzip members (or plain bytes) holding groups "varint64 count, {varint32 size, message}*".  Nodes are added in
// file order with the digraph ids id*2 / id*2+1 (ConvertVGNodeToNodes, :67-72), then the edges in file order (ConvertVGEdgeToEdges,
// :74-100).  Only the fields the reference reads are decoded: Graph.node (1), Graph.edge (2); Node.sequence (1), .name (2), .id (3);
// Edge.from (1), .to (2), .from_start (3), .to_end (4), .overlap (5).
struct VgReader
{
	const uint8_t* p; const uint8_t* end; bool ok = true;
	VgReader(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
	bool more() const { return ok && p < end; }
	uint64_t varint()
	{
		uint64_t v = 0; int shift = 0;
		while (p < end && shift < 64) { uint8_t b = *p++; v |= (uint64_t)(b & 0x7F) << shift; if (!(b & 0x80)) return v; shift += 7; }
		ok = false; return 0;
	}
	VgReader sub() { uint64_t n = varint(); if (!ok || n > (uint64_t)(end - p)) { ok = false; return VgReader(p, p); } VgReader r(p, p + n); p += n; return r; }
	void skip(uint32_t wire)
	{
		if (wire == 0) varint();
		else if (wire == 1) { if (end - p < 8) ok = false; else p += 8; }
		else if (wire == 2) sub();
		else if (wire == 5) { if (end - p < 4) ok = false; else p += 4; }
		else ok = false;
	}
};
inline std::string vgFileBytes(const std::string& filename)
{
	std::ifstream file(filename, std::ios::in | std::ios::binary);
	if (!file.good()) throw std::runtime_error("No graph file exists");
	std::string raw((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
	if (raw.size() < 2 || (uint8_t)raw[0] != 0x1f || (uint8_t)raw[1] != 0x8b) return raw;
	std::string data;
	size_t at = 0;
	std::vector<char> buf(1 << 20);
	while (at < raw.size())
	{
		z_stream zs; memset(&zs, 0, sizeof(zs));
		if (inflateInit2(&zs, 15 + 32) != Z_OK) throw std::runtime_error("zlib: inflateInit2 failed");
		zs.next_in = (Bytef*)raw.data() + at; zs.avail_in = (uInt)std::min<size_t>(raw.size() - at, 1u << 30);
		int ret;
		do
		{
			zs.next_out = (Bytef*)buf.data(); zs.avail_out = (uInt)buf.size();
			ret = inflate(&zs, Z_NO_FLUSH);
			if (ret != Z_OK && ret != Z_STREAM_END) { inflateEnd(&zs); throw std::runtime_error("the .vg file is not a valid gzip stream"); }
			data.append(buf.data(), buf.size() - zs.avail_out);
			if (ret == Z_OK && zs.avail_in == 0 && zs.avail_out != 0) { inflateEnd(&zs); throw std::runtime_error("the .vg file ends inside a gzip member"); }
		} while (ret != Z_STREAM_END);
		at = (size_t)((const char*)zs.next_in - raw.data());
		inflateEnd(&zs);
	}
	return data;
}
inline SplitGraph loadVg(const std::string& filename)
{
	const std::string data = vgFileBytes(filename);
	struct VgEdge { uint64_t from = 0, to = 0; bool fromStart = false, toEnd = false; int64_t overlap = 0; };
	std::vector<VgEdge> edges;
	SplitGraph g;
	VgReader in((const uint8_t*)data.data(), (const uint8_t*)data.data() + data.size());
	while (in.more())
	{
		uint64_t count = in.varint();
		for (uint64_t m = 0; m < count && in.ok; m++)
		{
			VgReader graph = in.sub();
			while (graph.more())
			{
				uint64_t key = graph.varint();
				if ((key >> 3) == 1 && (key & 7) == 2)
				{
					VgReader node = graph.sub();
					std::string sequence, name; int64_t id = 0;
					while (node.more())
					{
						uint64_t k = node.varint();
						if ((k >> 3) == 1 && (k & 7) == 2) { VgReader f = node.sub(); sequence.assign((const char*)f.p, f.end - f.p); }
						else if ((k >> 3) == 2 && (k & 7) == 2) { VgReader f = node.sub(); name.assign((const char*)f.p, f.end - f.p); }
						else if ((k >> 3) == 3 && (k & 7) == 0) id = (int64_t)node.varint();
						else node.skip((uint32_t)(k & 7));
					}
					if (!node.ok) { graph.ok = false; break; }
					if (id < 0 || id + 1 >= std::numeric_limits<int>::max() / 2) throw std::runtime_error("vg node id out of range: " + std::to_string(id));
					std::string rc;
					rc.reserve(sequence.size());
					for (size_t i = sequence.size(); i-- > 0; ) rc += complement(sequence[i]);
					addNode(g, (int)id * 2, sequence, name, false);
					addNode(g, (int)id * 2 + 1, rc, name, true);
				}
				else if ((key >> 3) == 2 && (key & 7) == 2)
				{
					VgReader edge = graph.sub();
					VgEdge e;
					while (edge.more())
					{
						uint64_t k = edge.varint();
						if ((k & 7) != 0) { edge.skip((uint32_t)(k & 7)); continue; }
						uint64_t v = edge.varint();
						switch (k >> 3) { case 1: e.from = v; break; case 2: e.to = v; break; case 3: e.fromStart = v != 0; break; case 4: e.toEnd = v != 0; break; case 5: e.overlap = (int64_t)v; break; default: break; }
					}
					if (!edge.ok) { graph.ok = false; break; }
					edges.push_back(e);
				}
				else graph.skip((uint32_t)(key & 7));
			}
			if (!graph.ok) in.ok = false;
		}
	}
	if (!in.ok) throw std::runtime_error("the .vg file is not a stream of vg::Graph messages");
	for (const VgEdge& e : edges)
	{
		if (e.overlap != 0) throw std::runtime_error("Edge overlaps are not supported by the B200 path (only overlap 0)");
		size_t fromLeft, fromRight, toLeft, toRight;
		if (e.fromStart) { fromLeft = e.from * 2; fromRight = e.from * 2 + 1; } else { fromLeft = e.from * 2 + 1; fromRight = e.from * 2; }
		if (e.toEnd) { toLeft = e.to * 2; toRight = e.to * 2 + 1; } else { toLeft = e.to * 2 + 1; toRight = e.to * 2; }
		addEdge(g, (int)fromRight, (int)toRight);
		addEdge(g, (int)toLeft, (int)fromLeft);
	}
	return g;
}

// Tarjan SCC in the reference's visiting order (AlignmentGraph.cpp:1008-1105); componentNumber =
// reversed emission index, a topological rank on a DAG
inline void componentOrder(SplitGraph& g)
{
	size_t N = g.size();
	const size_t NONE = std::numeric_limits<size_t>::max();
	std::vector<size_t> index(N, NONE), lowlink(N, NONE), stack;
	std::vector<bool> onStack(N, false);
	std::vector<std::pair<size_t, size_t>> call; // (node, next neighbour)
	g.componentNumber.assign(N, NONE);
	size_t counter = 0, nextComponent = 0;
	for (size_t root = 0; root < N; root++)
	{
		if (index[root] != NONE) continue;
		call.emplace_back(root, 0);
		index[root] = lowlink[root] = counter++;
		stack.push_back(root); onStack[root] = true;
		while (!call.empty())
		{
			size_t v = call.back().first;
			size_t& ni = call.back().second;
			if (ni < g.outNeighbors[v].size())
			{
				size_t w = g.outNeighbors[v][ni];
				if (index[w] == NONE)
				{
					index[w] = lowlink[w] = counter++;
					stack.push_back(w); onStack[w] = true;
					call.emplace_back(w, 0); // `ni` stays: on return lowlink is folded in and ni advanced
					continue;
				}
				if (onStack[w]) lowlink[v] = std::min(lowlink[v], index[w]);
				ni++;
				continue;
			}
			if (lowlink[v] == index[v])
			{
				size_t w;
				do { w = stack.back(); stack.pop_back(); onStack[w] = false; g.componentNumber[w] = nextComponent; } while (w != v);
				nextComponent++;
			}
			call.pop_back();
			if (!call.empty())
			{
				size_t parent = call.back().first;
				lowlink[parent] = std::min(lowlink[parent], lowlink[v]);
				call.back().second++;
			}
		}
	}
	for (size_t i = 0; i < N; i++) g.componentNumber[i] = nextComponent - 1 - g.componentNumber[i];
}

// ---- chains (AlignmentGraph.cpp:309-642): union-find over bubbles / tips; only used by seed clustering
inline size_t ufFind(std::vector<size_t>& parent, size_t item)
{
	if (parent[item] == item) return item;
	std::vector<size_t> st;
	st.push_back(item);
	while (parent[st.back()] != st.back()) st.push_back(parent[st.back()]);
	for (size_t i : st) parent[i] = st.back();
	return st.back();
}
inline void ufMerge(std::vector<size_t>& parent, std::vector<size_t>& rank, size_t left, size_t right)
{
	left = ufFind(parent, left);
	right = ufFind(parent, right);
	if (rank[left] < rank[right]) std::swap(left, right);
	parent[right] = left;
	if (rank[left] == rank[right]) rank[left] += 1;
}

// Is `start` the entrance of a bubble -- do all walks leaving it (tips ignored) come together again in ONE node before any of
// them ends or returns -- and where?  (chainBubble's test, AlignmentGraph.cpp:309-373: the superbubble search of Onodera et al.)
// A node is expanded once all its non-tip parents have been expanded; the bubble closes when exactly one node is discovered
// but unexpanded and it is the only one ready.  Scratch arrays stamped per call replace per-call hash sets: the search runs
// once per original node.
struct BubbleSearch
{
	std::vector<uint32_t> expandedAt, discoveredAt;
	std::vector<size_t> ready;
	uint32_t stamp = 0;
	explicit BubbleSearch(size_t n) : expandedAt(n, 0), discoveredAt(n, 0) {}
	bool run(const SplitGraph& g, size_t start, const std::vector<bool>& ignorableTip, size_t& exitNode)
	{
		stamp++;
		ready.assign(1, start);
		discoveredAt[start] = stamp;
		size_t waiting = 1; // discovered, not expanded
		while (!ready.empty())
		{
			const size_t v = ready.back();
			ready.pop_back();
			expandedAt[v] = stamp;
			waiting--;
			if (g.outNeighbors[v].empty()) return false; // a walk ends inside
			for (const size_t u : g.outNeighbors[v])
			{
				if (ignorableTip[u] || u == v) continue;
				if (u == start) return false;
				if (discoveredAt[u] != stamp) { discoveredAt[u] = stamp; waiting++; }
				bool allParentsExpanded = true;
				for (const size_t w : g.inNeighbors[u])
					if (w != u && !ignorableTip[w] && expandedAt[w] != stamp) { allParentsExpanded = false; break; }
				if (allParentsExpanded) ready.push_back(u);
			}
			if (ready.size() == 1 && waiting == 1 && expandedAt[ready[0]] != stamp)
			{
				exitNode = ready[0];
				for (const size_t u : g.outNeighbors[exitNode]) if (u == start) return false;
				return true;
			}
		}
		return false;
	}
};

inline void findChains(SplitGraph& g)
{
	size_t N = g.size();
	const size_t NONE = std::numeric_limits<size_t>::max();
	std::vector<size_t>& chainNumber = g.chainNumber;
	chainNumber.resize(N);
	for (size_t i = 0; i < N; i++) chainNumber[i] = i;
	std::vector<bool> ignorableTip(N, false);
	std::vector<size_t> rank(N, 0);
	for (const auto& pair : g.nodeLookup)
		for (size_t i = 1; i < pair.second.size(); i++) ufMerge(chainNumber, rank, pair.second[0], pair.second[i]);
	// chainTips (AlignmentGraph.cpp:425-529)
	std::unordered_map<size_t, std::unordered_set<size_t>> tipChainers;
	std::vector<size_t> tipOrder; // first-insertion order of tipChainers keys (merges below commute on the partition)
	{
		std::vector<size_t> order(N);
		for (size_t i = 0; i < N; i++) order[i] = i;
		std::sort(order.begin(), order.end(), [&g](size_t left, size_t right) { return g.componentNumber[left] < g.componentNumber[right]; });
		size_t numComp = g.componentNumber[order.back()] + 1;
		std::vector<bool> fwTip(numComp, true), bwTip(numComp, true);
		for (size_t ind = order.size() - 1; ind < order.size(); ind--)
		{
			size_t i = order[ind];
			if (!fwTip[g.componentNumber[i]]) continue;
			for (auto nb : g.outNeighbors[i])
			{
				if (g.componentNumber[nb] == g.componentNumber[i]) { fwTip[g.componentNumber[i]] = false; break; }
				if (!fwTip[g.componentNumber[nb]]) { fwTip[g.componentNumber[i]] = false; break; }
			}
		}
		for (size_t ind = order.size() - 1; ind < order.size(); ind--)
		{
			size_t i = order[ind];
			if (!fwTip[g.componentNumber[i]]) continue;
			for (auto nb : g.outNeighbors[i]) ufMerge(chainNumber, rank, i, nb);
		}
		for (size_t ind = 0; ind < order.size(); ind++)
		{
			size_t i = order[ind];
			if (!bwTip[g.componentNumber[i]]) continue;
			for (auto nb : g.inNeighbors[i])
			{
				if (g.componentNumber[nb] == g.componentNumber[i]) { bwTip[g.componentNumber[i]] = false; break; }
				if (!bwTip[g.componentNumber[nb]]) { bwTip[g.componentNumber[i]] = false; break; }
			}
		}
		for (size_t ind = 0; ind < order.size(); ind++)
		{
			size_t i = order[ind];
			if (!bwTip[g.componentNumber[i]]) continue;
			for (auto nb : g.inNeighbors[i]) ufMerge(chainNumber, rank, i, nb);
		}
		for (size_t i = 0; i < N; i++)
		{
			if (bwTip[g.componentNumber[i]] || fwTip[g.componentNumber[i]]) ignorableTip[i] = true;
			if (bwTip[g.componentNumber[i]])
				for (auto nb : g.outNeighbors[i]) { if (chainNumber[nb] == chainNumber[i]) continue; if (!tipChainers.count(chainNumber[i])) tipOrder.push_back(chainNumber[i]); tipChainers[chainNumber[i]].insert(nb); }
			if (fwTip[g.componentNumber[i]])
				for (auto nb : g.inNeighbors[i]) { if (chainNumber[nb] == chainNumber[i]) continue; if (!tipChainers.count(chainNumber[i])) tipOrder.push_back(chainNumber[i]); tipChainers[chainNumber[i]].insert(nb); }
		}
	}
	// chainCycles (AlignmentGraph.cpp:531-581)
	for (size_t i = 0; i < N; i++)
	{
		size_t uniqueFw = NONE;
		for (auto u : g.outNeighbors[i])
		{
			if (ignorableTip[u] || u == i) continue;
			if (uniqueFw == NONE) uniqueFw = u; else uniqueFw = NONE - 1;
		}
		size_t uniqueBw = NONE;
		for (auto u : g.inNeighbors[i])
		{
			if (ignorableTip[u] || u == i) continue;
			if (uniqueBw == NONE) uniqueBw = u; else if (u != uniqueBw) uniqueBw = NONE - 1;
		}
		if (uniqueFw != uniqueBw) continue;
		if (uniqueFw == NONE || uniqueFw == NONE - 1) continue;
		ignorableTip[i] = true;
		ufMerge(chainNumber, rank, i, uniqueFw);
	}
	// chainBubble for the last split node of every original node (AlignmentGraph.cpp:375-399, 596-599)
	BubbleSearch bubbles(N);
	for (const auto& pair : g.nodeLookup)
	{
		size_t start = pair.second.back();
		size_t bubbleEnd = 0;
		if (!bubbles.run(g, start, ignorableTip, bubbleEnd)) continue;
		std::unordered_set<size_t> visited;
		std::vector<size_t> stack { start };
		visited.insert(start);
		ufMerge(chainNumber, rank, start, bubbleEnd);
		while (stack.size() > 0)
		{
			const size_t top = stack.back();
			stack.pop_back();
			if (visited.count(top) == 1) continue; // sic: `start` is already visited, so the walk ends at once
			if (ignorableTip[top]) continue;
			visited.insert(top);
			ufMerge(chainNumber, rank, start, top);
			for (const auto nb : g.outNeighbors[top]) { if (visited.count(nb) == 1) continue; if (nb == bubbleEnd) continue; stack.push_back(nb); }
		}
	}
	for (size_t key : tipOrder)
	{
		auto& set = tipChainers[key];
		size_t uniqueNeighbor = NONE;
		for (auto n : set)
		{
			if (uniqueNeighbor == NONE) uniqueNeighbor = chainNumber[n];
			if (uniqueNeighbor != chainNumber[n]) { uniqueNeighbor = NONE - 1; break; }
		}
		if (uniqueNeighbor == NONE - 1) continue;
		ufMerge(chainNumber, rank, key, *set.begin());
	}
	for (size_t i = 0; i < N; i++) ufFind(chainNumber, i);
	// fixChainApproxPos (AlignmentGraph.cpp:401-423)
	g.chainApproxPos.assign(N, NONE);
	for (size_t s = 0; s < N; s++)
	{
		if (g.chainApproxPos[s] != NONE) continue;
		std::vector<std::pair<size_t, size_t>> stack;
		size_t chain = chainNumber[s];
		stack.emplace_back(s, (N + 5) * 64);
		while (stack.size() > 0)
		{
			size_t v = stack.back().first, dist = stack.back().second;
			stack.pop_back();
			if (g.chainApproxPos[v] != NONE) continue;
			g.chainApproxPos[v] = dist;
			for (const size_t u : g.outNeighbors[v]) { if (chainNumber[u] != chain) continue; if (g.chainApproxPos[u] != NONE) continue; stack.emplace_back(u, dist + g.nodeLength[u]); }
			for (const size_t u : g.inNeighbors[v]) { if (chainNumber[u] != chain) continue; if (g.chainApproxPos[u] != NONE) continue; stack.emplace_back(u, dist - g.nodeLength[v]); }
		}
	}
}

// ---- path cover index for co-linear chaining (what AlignmentGraph::buildMPC provides, AlignmentGraph.cpp:1157-1489)
//
// K2 uses the cover for REACHABILITY only (gc_k2.cuh): an anchor ending at node e precedes one starting at s iff some path k
// through e has its last s-reaching node at or after e.  That holds for every set of paths that covers all nodes, so the
// cover need not be the reference's -- only narrow, because K2's work grows with the width.  Built here per weakly connected
// component, components in parallel:
//   1. an initial cover by forward walks: in topological order every still uncovered node starts a path that prefers
//      uncovered successors (two walks cover a chain of SNP bubbles);
//   2. width reduction to the minimum (Dilworth width of the DAG): the cover is a flow with a lower bound of one per node;
//      while the residual network of (flow - lower bound) has a source-sink route, one path is cancelled along it;
//   3. decomposition of the final flow into paths;
//   4. per path, one sweep over the component in topological order gives every node the last position on that path that
//      reaches it (paths in parallel, one int32 array per thread -- not the reference's N x K table).
typedef long long LL;

struct ComponentCover
{
	// local numbering 0..n-1 of one component, CSR adjacency, Kahn order
	std::vector<uint32_t> outStart, outTo, inStart, inFrom, topoOrder, topoIndex;
	size_t n = 0;
};

inline void buildComponentsMap(SplitGraph& g)
{
	// weakly connected components by flooding from every unlabelled node (index inside the component = flood order)
	const size_t N = g.size(), UNSET = N + 1;
	g.component_map.assign(N, UNSET);
	g.component_idx.assign(N, UNSET);
	g.component_ids.clear();
	std::vector<size_t> frontier;
	for (size_t root = 0; root < N; root++)
	{
		if (g.component_map[root] != UNSET) continue;
		const size_t c = g.component_ids.size();
		frontier.assign(1, root);
		g.component_map[root] = c; g.component_idx[root] = 0;
		for (size_t head = 0; head < frontier.size(); head++)
		{
			const size_t v = frontier[head];
			auto visit = [&](size_t w) { if (g.component_map[w] == UNSET) { g.component_map[w] = c; g.component_idx[w] = frontier.size(); frontier.push_back(w); } };
			for (size_t w : g.outNeighbors[v]) visit(w);
			for (size_t w : g.inNeighbors[v]) visit(w);
		}
		g.component_ids.push_back(frontier);
	}
}

inline ComponentCover localComponent(const SplitGraph& g, size_t cid)
{
	const std::vector<size_t>& ids = g.component_ids[cid];
	ComponentCover c;
	c.n = ids.size();
	c.outStart.assign(c.n + 1, 0); c.inStart.assign(c.n + 1, 0);
	for (size_t i = 0; i < c.n; i++) { c.outStart[i + 1] = c.outStart[i] + (uint32_t)g.outNeighbors[ids[i]].size(); c.inStart[i + 1] = c.inStart[i] + (uint32_t)g.inNeighbors[ids[i]].size(); }
	c.outTo.resize(c.outStart[c.n]); c.inFrom.resize(c.inStart[c.n]);
	for (size_t i = 0; i < c.n; i++)
	{
		uint32_t o = c.outStart[i]; for (size_t w : g.outNeighbors[ids[i]]) c.outTo[o++] = (uint32_t)g.component_idx[w];
		uint32_t q = c.inStart[i]; for (size_t w : g.inNeighbors[ids[i]]) c.inFrom[q++] = (uint32_t)g.component_idx[w];
	}
	// Kahn order; a node left over means a directed cycle
	std::vector<uint32_t> pending(c.n);
	c.topoOrder.clear(); c.topoOrder.reserve(c.n);
	for (size_t i = 0; i < c.n; i++) { pending[i] = c.inStart[i + 1] - c.inStart[i]; if (pending[i] == 0) c.topoOrder.push_back((uint32_t)i); }
	for (size_t head = 0; head < c.topoOrder.size(); head++)
	{
		uint32_t v = c.topoOrder[head];
		for (uint32_t e = c.outStart[v]; e < c.outStart[v + 1]; e++) if (--pending[c.outTo[e]] == 0) c.topoOrder.push_back(c.outTo[e]);
	}
	if (c.topoOrder.size() < c.n) throw std::runtime_error("The input sequence graph has a directed cycle.\nThe current version of GraphChainer only supports DAGs.");
	c.topoIndex.resize(c.n);
	for (size_t i = 0; i < c.n; i++) c.topoIndex[c.topoOrder[i]] = (uint32_t)i;
	return c;
}

// the cover as a flow: units through every node and edge, path starts and ends per node
struct CoverFlow { std::vector<uint32_t> node, edge, starts, ends; size_t width = 0; };

inline CoverFlow initialCover(const ComponentCover& c)
{
	CoverFlow f;
	f.node.assign(c.n, 0); f.edge.assign(c.outTo.size(), 0); f.starts.assign(c.n, 0); f.ends.assign(c.n, 0);
	for (uint32_t v0 : c.topoOrder)
	{
		if (f.node[v0]) continue;
		f.starts[v0]++; f.width++;
		uint32_t v = v0;
		while (true)
		{
			f.node[v]++;
			uint32_t pick = 0xFFFFFFFFu;
			for (uint32_t e = c.outStart[v]; e < c.outStart[v + 1]; e++) { if (pick == 0xFFFFFFFFu) pick = e; if (!f.node[c.outTo[e]]) { pick = e; break; } }
			if (pick == 0xFFFFFFFFu) break;
			f.edge[pick]++;
			v = c.outTo[pick];
		}
		f.ends[v]++;
	}
	return f;
}

// One round: find a route source -> sink in the residual network and cancel one unit along it.  Vertices: in(v) = 2v,
// out(v) = 2v + 1, source, sink.  An arc can be traversed forwards if its flow exceeds its lower bound (the unit is removed)
// and backwards always (a unit is added).  Returns false when the width is minimal.
inline bool cancelOnePath(const ComponentCover& c, CoverFlow& f, std::vector<int64_t>& via, std::vector<uint32_t>& queue)
{
	const size_t n = c.n;
	const uint32_t SRC = (uint32_t)(2 * n), SNK = (uint32_t)(2 * n + 1);
	// via[x]: how x was reached -- arc code: (kind << 33) | (index << 1) | backwards
	enum { START = 0, END = 1, NODE = 2, EDGE = 3 };
	auto code = [](int kind, uint64_t index, bool backwards) { return (int64_t)(((uint64_t)kind << 40) | (index << 1) | (backwards ? 1u : 0u)); };
	via.assign(2 * n + 2, -1);
	queue.clear();
	queue.push_back(SRC); via[SRC] = -2;
	bool found = false;
	for (size_t head = 0; head < queue.size() && !found; head++)
	{
		const uint32_t x = queue[head];
		auto reach = [&](uint32_t y, int64_t how) { if (via[y] == -1) { via[y] = how; queue.push_back(y); if (y == SNK) found = true; } };
		if (x == SRC) { for (uint32_t v = 0; v < n; v++) if (f.starts[v] > 0) reach(2 * v, code(START, v, false)); }
		else if (x == SNK) { }
		else if ((x & 1) == 0)
		{
			const uint32_t v = x >> 1;
			if (f.node[v] > 1) reach(2 * v + 1, code(NODE, v, false));                                     // one unit fewer through v (lower bound 1)
			reach(SRC, code(START, v, true));                                                             // a new path could start here
			for (uint32_t q = c.inStart[v]; q < c.inStart[v + 1]; q++)
			{
				// backwards over edge u -> v: find its CSR slot in u's out-list
				const uint32_t u = c.inFrom[q];
				for (uint32_t e = c.outStart[u]; e < c.outStart[u + 1]; e++) if (c.outTo[e] == v) { reach(2 * u + 1, code(EDGE, e, true)); break; }
			}
		}
		else
		{
			const uint32_t v = x >> 1;
			if (f.ends[v] > 0) reach(SNK, code(END, v, false));
			reach(2 * v, code(NODE, v, true));
			for (uint32_t e = c.outStart[v]; e < c.outStart[v + 1]; e++) if (f.edge[e] > 0) reach(2 * c.outTo[e], code(EDGE, e, false));
		}
	}
	if (!found) return false;
	// walk back from the sink, applying the changes
	uint32_t x = SNK;
	while (x != SRC)
	{
		const int64_t how = via[x];
		const int kind = (int)((uint64_t)how >> 40); const uint64_t index = ((uint64_t)how & ((1ull << 40) - 1)) >> 1; const bool backwards = how & 1;
		uint32_t from;
		if (kind == START) { if (!backwards) { f.starts[index]--; from = SRC; } else { f.starts[index]++; from = (uint32_t)(2 * index); } }
		else if (kind == END) { f.ends[index]--; from = (uint32_t)(2 * index + 1); }
		else if (kind == NODE) { if (!backwards) { f.node[index]--; from = (uint32_t)(2 * index); } else { f.node[index]++; from = (uint32_t)(2 * index + 1); } }
		else
		{
			// edge index = CSR slot e of u -> v
			uint32_t e = (uint32_t)index, v = c.outTo[e];
			uint32_t u = (uint32_t)(std::upper_bound(c.outStart.begin(), c.outStart.end(), e) - c.outStart.begin() - 1);
			if (!backwards) { f.edge[e]--; from = 2 * u + 1; } else { f.edge[e]++; from = 2 * v; }
		}
		x = from;
	}
	f.width--;
	return true;
}

// the flow as explicit paths (local node ids)
inline std::vector<std::vector<uint32_t>> decomposeCover(const ComponentCover& c, CoverFlow f)
{
	std::vector<std::vector<uint32_t>> paths;
	for (uint32_t s = 0; s < c.n; s++)
		while (f.starts[s] > 0)
		{
			f.starts[s]--;
			std::vector<uint32_t> path;
			uint32_t v = s;
			while (true)
			{
				path.push_back(v);
				uint32_t next = 0xFFFFFFFFu;
				for (uint32_t e = c.outStart[v]; e < c.outStart[v + 1]; e++) if (f.edge[e] > 0) { f.edge[e]--; next = c.outTo[e]; break; }
				if (next == 0xFFFFFFFFu) { f.ends[v]--; break; }
				v = next;
			}
			paths.push_back(std::move(path));
		}
	return paths;
}

inline void buildComponentIndex(SplitGraph& g, size_t cid, bool verbose)
{
	const ComponentCover c = localComponent(g, cid);
	CoverFlow f = initialCover(c);
	const size_t initialWidth = f.width;
	{
		std::vector<int64_t> via; std::vector<uint32_t> queue;
		while (cancelOnePath(c, f, via, queue)) { }
	}
	const std::vector<std::vector<uint32_t>> cover = decomposeCover(c, f);
	const std::vector<size_t>& ids = g.component_ids[cid];
	const size_t n = c.n, K = cover.size();
	g.mpc[cid].assign(K, {});
	for (size_t k = 0; k < K; k++) for (uint32_t v : cover[k]) g.mpc[cid][k].push_back(ids[v]);
	g.topo_ids[cid].assign(n, 0);
	for (size_t v = 0; v < n; v++) g.topo_ids[cid][v] = c.topoIndex[v];
	g.paths[cid].assign(n, {});
	g.backwards[cid].assign(n, {});
	for (size_t k = 0; k < K; k++) for (uint32_t v : cover[k]) g.paths[cid][v].push_back(k); // k ascending per node
	// per path: last position on it that reaches each node; a node on the path itself links to its predecessor on the path
	std::vector<std::vector<std::pair<size_t, size_t>>> perPath(K); // (node, linked local node) for path k
	#pragma omp parallel for schedule(dynamic, 1)
	for (size_t k = 0; k < K; k++)
	{
		std::vector<int32_t> last(n, -1), own(n, -1);
		for (size_t p = 0; p < cover[k].size(); p++) own[cover[k][p]] = (int32_t)p;
		for (uint32_t v : c.topoOrder)
		{
			int32_t best = -1;
			for (uint32_t q = c.inStart[v]; q < c.inStart[v + 1]; q++) { int32_t r = last[c.inFrom[q]]; if (r > best) best = r; }
			if (best >= 0) perPath[k].emplace_back(v, cover[k][best]);
			last[v] = own[v] > best ? own[v] : best;
		}
	}
	for (size_t k = 0; k < K; k++) for (const auto& link : perPath[k]) g.backwards[cid][link.first].push_back({ link.second, k }); // k ascending per node
	if (verbose)
	{
		#pragma omp critical
		std::cout << "cid = " << cid << " greedy width " << initialWidth << " optimal width " << K << std::endl;
	}
}

inline void buildMpc(SplitGraph& g, bool verbose)
{
	buildComponentsMap(g);
	size_t C = g.component_ids.size();
	g.mpc.resize(C); g.topo_ids.resize(C); g.paths.resize(C); g.backwards.resize(C);
	#pragma omp parallel for schedule(dynamic, 1)
	for (size_t cid = 0; cid < C; cid++) buildComponentIndex(g, cid, verbose);
	size_t tw = 0, mw = 0;
	for (size_t cid = 0; cid < C; cid++) { tw += g.mpc[cid].size(); mw = std::max(mw, g.mpc[cid].size()); }
	if (verbose) std::cout << "MPC building done" << std::endl << "total width " << tw << " and max component width " << mw << std::endl;
}

// ---- minimizer index, single bucket = the reference at -t 1 (MinimizerSeeder.cpp:286-492)
inline uint64_t mzHash(uint64_t key)
{
	key = (~key) + (key << 21);
	key = key ^ (key >> 24);
	key = (key + (key << 3)) + (key << 8);
	key = key ^ (key >> 14);
	key = (key + (key << 2)) + (key << 4);
	key = key ^ (key >> 28);
	key = key + (key << 31);
	return key;
}

// iterateMinimizersReal (MinimizerSeeder.cpp:104-192), sequences here are pure A C G T
template <typename F>
void iterateMinimizers(const std::string& str, size_t k, size_t windowSize, F callback)
{
	auto code = [](char c) -> size_t { switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; default: return 3; } };
	if (str.size() < k) return;
	const size_t realWindow = windowSize - k + 1;
	const size_t mask = ~(0xFFFFFFFFFFFFFFFFull << (k * 2));
	if (windowSize > str.size()) return;
	std::deque<std::tuple<size_t, size_t, size_t>> window;
	size_t kmer = 0;
	for (size_t i = 0; i < k; i++) { kmer <<= 2; kmer |= code(str[i]); }
	window.emplace_back(k - 1, kmer, mzHash(kmer));
	for (size_t i = k; i < k + realWindow; i++)
	{
		kmer <<= 2; kmer &= mask; kmer |= code(str[i]);
		auto hashed = mzHash(kmer);
		while (!window.empty() && std::get<2>(window.back()) > hashed) window.pop_back();
		window.emplace_back(i, kmer, hashed);
	}
	{
		auto iter = window.begin();
		while (iter != window.end() && std::get<2>(*iter) == std::get<2>(window.front())) { callback(std::get<0>(*iter), std::get<1>(*iter)); ++iter; }
	}
	for (size_t i = k + realWindow; i < str.size(); i++)
	{
		kmer <<= 2; kmer &= mask; kmer |= code(str[i]);
		auto hashed = mzHash(kmer);
		size_t oldMinimum = std::get<2>(window.front());
		bool frontPopped = false;
		while (!window.empty() && std::get<0>(window.front()) <= i - realWindow) { frontPopped = true; window.pop_front(); }
		if (frontPopped) while (window.size() >= 2 && std::get<2>(window.front()) == std::get<2>(*(window.begin() + 1))) window.pop_front();
		while (!window.empty() && std::get<2>(window.back()) > hashed) window.pop_back();
		window.emplace_back(i, kmer, hashed);
		if (std::get<2>(window.front()) != oldMinimum)
		{
			auto iter = window.begin();
			while (iter != window.end() && std::get<2>(*iter) == std::get<2>(window.front())) { callback(std::get<0>(*iter), std::get<1>(*iter)); ++iter; }
		}
		else if (std::get<2>(window.back()) == std::get<2>(window.front())) callback(std::get<0>(window.back()), std::get<1>(window.back()));
	}
}

struct MinimizerIndex { std::vector<uint64_t> kmers, positions; std::vector<uint32_t> kmerStart; uint64_t maxCount = 0; };

inline MinimizerIndex buildMinimizers(const SplitGraph& g, size_t k, size_t windowSize, double discardMostNumerousFraction)
{
	std::unordered_map<size_t, size_t> nodeMinimizerStart;
	for (size_t i = 0; i < g.size(); i++)
	{
		nodeMinimizerStart[g.nodeIDs[i]] = std::max(nodeMinimizerStart[g.nodeIDs[i]], (size_t)0);
		bool skipStart = false;
		for (auto n : g.inNeighbors[i]) if (g.nodeIDs[n] != g.nodeIDs[i]) { skipStart = true; break; }
		if (skipStart) nodeMinimizerStart[g.nodeIDs[i]] = std::max(nodeMinimizerStart[g.nodeIDs[i]], g.nodeOffset[i]);
	}
	std::vector<std::pair<uint64_t, uint64_t>> hits; // (kmer, split << 6 | offset) in emission order
	for (const auto& entry : g.nodeLookup) // iteration order = insertion order of positions per k-mer
	{
		int nodeId = entry.first;
		std::string sequence;
		sequence.resize(g.originalNodeSize.at(nodeId));
		for (size_t node : entry.second)
			for (size_t p = 0; p < g.nodeLength[node]; p++) sequence[g.nodeOffset[node] + p] = g.base(node, p);
		size_t minStart = nodeMinimizerStart.at(nodeId);
		iterateMinimizers(sequence, k, windowSize, [&](size_t pos, size_t kmer)
		{
			if (pos < minStart) return;
			size_t splitNode = g.unitigNode(nodeId, pos);
			hits.emplace_back((uint64_t)kmer, ((uint64_t)splitNode << 6) + (pos - g.nodeOffset[splitNode]));
		});
	}
	// group by k-mer; inside a k-mer the reference's counting sort fills from the END of the range
	// while scanning the hits forwards (MinimizerSeeder.cpp:470-482) => reverse emission order
	MinimizerIndex idx;
	std::vector<size_t> order(hits.size());
	for (size_t i = 0; i < hits.size(); i++) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&hits](size_t a, size_t b) { return hits[a].first < hits[b].first; });
	idx.kmerStart.push_back(0);
	for (size_t i = 0; i < order.size(); )
	{
		size_t j = i;
		while (j < order.size() && hits[order[j]].first == hits[order[i]].first) j++;
		idx.kmers.push_back(hits[order[i]].first);
		for (size_t x = j; x-- > i; ) idx.positions.push_back(hits[order[x]].second);
		idx.kmerStart.push_back((uint32_t)idx.positions.size());
		i = j;
	}
	// initMaxCount (MinimizerSeeder.cpp:557-575).  The reference leaves out ONE key per bucket (the last minimal-perfect-hash
	// index; its bucket count equals -t) -- which key depends on BBHash internals.  Here ONE key is left out as well (the last
	// in k-mer order, the single-bucket case): the quantile can only differ from the reference's when the count of the key it
	// drops sits exactly on the cut.  Parity is pinned to the reference's -t 1 index (tests/golden); the
	// bench's parity gate compares against the reference at -t <all cores> on 4000 reads per run.
	std::vector<size_t> counts;
	for (size_t i = 0; i + 1 < idx.kmerStart.size(); i++) counts.push_back(idx.kmerStart[i + 1] - idx.kmerStart[i]);
	if (!counts.empty()) counts.pop_back();
	std::sort(counts.begin(), counts.end());
	if (!counts.empty())
	{
		size_t index = (size_t)(counts.size() * (1.0 - discardMostNumerousFraction));
		if (index == counts.size()) index = counts.size() - 1;
		idx.maxCount = counts[index] + 1;
	}
	return idx;
}

template <typename T, typename V> std::vector<T> conv(const V& v) { std::vector<T> r; r.reserve(v.size()); for (auto x : v) r.push_back((T)x); return r; }

inline GcIndexFile toIndex(const SplitGraph& g, const MinimizerIndex& mz, size_t k, size_t windowSize)
{
	GcIndexFile f;
	size_t N = g.size();
	f.putU32("nodeLength", conv<uint32_t>(g.nodeLength));
	f.putU32("nodeOffset", conv<uint32_t>(g.nodeOffset));
	f.putI32("nodeIDs", conv<int32_t>(g.nodeIDs));
	f.putU8("reverse", conv<uint8_t>(g.reverse));
	// findLinearizable (AlignmentGraph.cpp:644-736) marks its start node `checked` before walking and then
	// immediately takes the "already checked" exit, so the reference's array is all false for every graph.
	f.putU8("linearizable", std::vector<uint8_t>(N, 0));
	f.putU32("componentNumber", conv<uint32_t>(g.componentNumber));
	f.putU32("chainNumber", conv<uint32_t>(g.chainNumber));
	f.putU64("chainApproxPos", conv<uint64_t>(g.chainApproxPos));
	f.putU64("firstAmbiguous", { (uint64_t)N });
	f.putU64("nodeSeq", g.nodeSeq);
	auto csr = [&](const std::string& name, const std::vector<std::vector<size_t>>& adj)
	{
		std::vector<uint32_t> start { 0 }, nbr;
		for (const auto& l : adj) { for (auto x : l) nbr.push_back((uint32_t)x); start.push_back((uint32_t)nbr.size()); }
		f.putU32(name + "Start", start); f.putU32(name + "Nbr", nbr);
	};
	csr("in", g.inNeighbors);
	csr("out", g.outNeighbors);
	{
		std::vector<int32_t> ids; std::vector<uint32_t> start { 0 }, nodes, sizes, nameOff { 0 }; std::vector<uint8_t> names;
		for (const auto& pair : g.nodeLookup)
		{
			ids.push_back(pair.first);
			for (auto x : pair.second) nodes.push_back((uint32_t)x);
			start.push_back((uint32_t)nodes.size());
			sizes.push_back((uint32_t)g.originalNodeSize.at(pair.first));
			const std::string& nm = g.originalNodeName.at(pair.first);
			names.insert(names.end(), nm.begin(), nm.end());
			nameOff.push_back((uint32_t)names.size());
		}
		f.putI32("origIds", ids); f.putU32("origStart", start); f.putU32("origNodes", nodes); f.putU32("origSize", sizes); f.putU32("origNameOff", nameOff); f.putU8("origNames", names);
	}
	{
		f.putU32("compMap", conv<uint32_t>(g.component_map));
		f.putU32("compIdx", conv<uint32_t>(g.component_idx));
		std::vector<uint32_t> compStart { 0 }, compIds, topoIds, width, pathsStart { 0 }, pathsK, backStart { 0 }, backNode, backK, mpcStart { 0 }, mpcPathStart { 0 }, mpcNodes;
		for (size_t c = 0; c < g.component_ids.size(); c++)
		{
			for (auto x : g.component_ids[c]) compIds.push_back((uint32_t)x);
			compStart.push_back((uint32_t)compIds.size());
			for (auto x : g.topo_ids[c]) topoIds.push_back((uint32_t)x);
			width.push_back((uint32_t)g.mpc[c].size());
			for (size_t i = 0; i < g.component_ids[c].size(); i++)
			{
				for (auto kk : g.paths[c][i]) pathsK.push_back((uint32_t)kk);
				pathsStart.push_back((uint32_t)pathsK.size());
				for (auto b : g.backwards[c][i]) { backNode.push_back((uint32_t)b.first); backK.push_back((uint32_t)b.second); }
				backStart.push_back((uint32_t)backNode.size());
			}
			for (const auto& p : g.mpc[c]) { for (auto x : p) mpcNodes.push_back((uint32_t)x); mpcPathStart.push_back((uint32_t)mpcNodes.size()); }
			mpcStart.push_back((uint32_t)(mpcPathStart.size() - 1));
		}
		f.putU32("compStart", compStart); f.putU32("compIds", compIds); f.putU32("topoIds", topoIds); f.putU32("mpcWidth", width);
		f.putU32("pathsStart", pathsStart); f.putU32("pathsK", pathsK); f.putU32("backStart", backStart); f.putU32("backNode", backNode); f.putU32("backK", backK);
		f.putU32("mpcStart", mpcStart); f.putU32("mpcPathStart", mpcPathStart); f.putU32("mpcNodes", mpcNodes);
	}
	f.putU32("mzBucketStart", { 0, (uint32_t)mz.kmers.size() });
	f.putU64("mzKmers", mz.kmers);
	f.putU32("mzKmerStart", mz.kmerStart);
	f.putU64("mzPositions", mz.positions);
	f.putU64("mzParams", { (uint64_t)k, (uint64_t)windowSize, (uint64_t)mz.maxCount, 1 });
	f.putU64("bpSize", { (uint64_t)g.bpSize });
	return f;
}

inline GcIndexFile buildIndexFromGfa(const std::string& gfaPath, size_t k, size_t windowSize, double discardMostNumerousFraction, bool verbose)
{
	const bool vg = gfaPath.size() > 3 && gfaPath.compare(gfaPath.size() - 3, 3, ".vg") == 0; // getGraph, Aligner.cpp:1079-1100
	SplitGraph g = vg ? loadVg(gfaPath) : loadGfa(gfaPath);
	if (verbose) std::cout << "Build alignment graph" << std::endl;
	componentOrder(g);
	findChains(g);
	if (verbose)
	{
		std::cout << g.nodeLookup.size() << " original nodes, " << (g.nodeLookup.size() / 2) << " in one strand" << std::endl;
		std::cout << g.size() << " split nodes, " << (g.size() / 2) << " in one strand" << std::endl;
		std::cout << "Build MPC Index" << std::endl;
	}
	buildMpc(g, verbose);
	if (verbose) std::cout << "Build minimizer seeder from the graph" << std::endl;
	MinimizerIndex mz = buildMinimizers(g, k, windowSize, discardMostNumerousFraction);
	return toIndex(g, mz, k, windowSize);
}

}
