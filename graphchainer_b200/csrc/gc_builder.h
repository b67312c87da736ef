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
	std::string line;
	while (std::getline(file, line))
	{
		if (line.size() == 0) continue;
		if (line[0] == 'S')
		{
			std::stringstream sstr { line };
			std::string dummy, idstr, seq;
			sstr >> dummy >> idstr;
			int id = getNameId(idstr);
			sstr >> seq;
			if (seq == "*") throw std::runtime_error("Nodes without sequence (*) are not currently supported (nodeid " + idstr + ")");
			nodes[id] = seq;
		}
		else if (line[0] == 'L')
		{
			std::stringstream sstr { line };
			std::string dummy, fromstr, tostr, fromstart, toend;
			int overlap = 0;
			sstr >> dummy >> fromstr;
			int from = getNameId(fromstr);
			sstr >> fromstart >> tostr;
			int to = getNameId(tostr);
			sstr >> toend >> overlap;
			if (overlap != 0) throw std::runtime_error("Edge overlaps are not supported by the B200 path (only 0M links)");
			edges[NodePos { from, fromstart == "+" }].push_back(NodePos { to, toend == "+" });
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

inline std::pair<bool, size_t> findBubble(const SplitGraph& g, size_t start, const std::vector<bool>& ignorableTip)
{
	std::vector<size_t> S { start };
	std::unordered_set<size_t> visited, seen;
	seen.insert(start);
	while (S.size() > 0)
	{
		const size_t v = S.back();
		S.pop_back();
		seen.erase(v);
		visited.insert(v);
		if (g.outNeighbors[v].size() == 0) return std::make_pair(false, (size_t)0);
		for (const size_t u : g.outNeighbors[v])
		{
			if (ignorableTip[u]) continue;
			if (u == v) continue;
			if (u == start) return std::make_pair(false, (size_t)0);
			seen.insert(u);
			bool hasNonvisitedParent = false;
			for (const size_t w : g.inNeighbors[u])
			{
				if (w == u) continue;
				if (!ignorableTip[w] && visited.count(w) == 0) { hasNonvisitedParent = true; break; }
			}
			if (!hasNonvisitedParent) S.push_back(u);
		}
		if (S.size() == 1 && seen.size() == 1 && seen.count(S[0]) == 1)
		{
			const size_t t = S.back();
			for (const size_t u : g.outNeighbors[t]) if (u == start) return std::make_pair(false, (size_t)0);
			return std::make_pair(true, t);
		}
	}
	return std::make_pair(false, (size_t)0);
}

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
	for (const auto& pair : g.nodeLookup)
	{
		size_t start = pair.second.back();
		auto bubble = findBubble(g, start, ignorableTip);
		if (!bubble.first) continue;
		size_t bubbleEnd = bubble.second;
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

// ---- minimum path cover (AlignmentGraph.cpp:1157-1489)
typedef long long LL;

inline void buildComponentsMap(SplitGraph& g)
{
	size_t N = g.size();
	g.component_map.assign(N, N + 1);
	g.component_idx.assign(N, N + 1);
	g.component_ids.clear();
	std::vector<size_t> Q;
	for (size_t S = 0; S < N; S++)
	{
		if (g.component_map[S] != N + 1) continue;
		Q.clear();
		Q.push_back(S);
		size_t c = g.component_ids.size();
		g.component_map[S] = c; g.component_idx[S] = 0;
		for (size_t i = 0; i < Q.size(); )
		{
			size_t s = Q[i++];
			for (size_t t : g.outNeighbors[s]) if (g.component_map[t] == N + 1) { g.component_map[t] = c; g.component_idx[t] = Q.size(); Q.push_back(t); }
			for (size_t t : g.inNeighbors[s]) if (g.component_map[t] == N + 1) { g.component_map[t] = c; g.component_idx[t] = Q.size(); Q.push_back(t); }
		}
		g.component_ids.push_back(Q);
	}
}

// greedy cover: repeatedly the path with the most uncovered nodes (AlignmentGraph.cpp:1267-1326)
inline std::vector<std::vector<size_t>> greedyCover(const SplitGraph& g, size_t cid)
{
	const std::vector<size_t>& cids = g.component_ids[cid];
	size_t N = cids.size();
	std::vector<std::vector<size_t>> ret;
	std::vector<size_t> covered(N, 0);
	size_t covered_cnt = 0;
	std::vector<std::pair<size_t, size_t>> d(N);
	std::vector<size_t> incd(N), Q(N);
	while (covered_cnt < covered.size())
	{
		size_t Qsize = 0;
		for (size_t i = 0; i < N; i++)
		{
			d[i] = std::make_pair((size_t)0, i);
			incd[i] = g.inNeighbors[cids[i]].size();
			if (incd[i] == 0) Q[Qsize++] = i;
		}
		std::pair<size_t, size_t> best = { 0, 0 };
		for (size_t i = 0; i < Qsize; )
		{
			size_t s = Q[i++];
			if (covered[s] == 0) d[s].first++;
			best = std::max(best, { d[s].first, s });
			for (size_t tid : g.outNeighbors[cids[s]])
			{
				size_t t = g.component_idx[tid];
				incd[t]--;
				d[t] = std::max(d[t], { d[s].first, s });
				if (incd[t] == 0) Q[Qsize++] = t;
			}
		}
		if (Qsize < N) throw std::runtime_error("The input sequence graph has a directed cycle.\nThe current version of GraphChainer only supports DAGs.");
		std::vector<size_t> tmp, path;
		if (best.second == d[best.second].second) tmp.push_back(best.second);
		else for (size_t i = best.second; d[i].second != i || i != tmp.back(); i = d[i].second) tmp.push_back(i);
		std::reverse(tmp.begin(), tmp.end());
		size_t l = 0, r = tmp.size() - 1;
		while (covered[tmp[l]]) l++;
		while (covered[tmp[r]]) r--;
		size_t new_covered = 0;
		for (size_t i = l; i <= r; i++)
		{
			path.push_back(cids[tmp[i]]);
			if (covered[tmp[i]] == 0) new_covered++;
			covered[tmp[i]]++;
		}
		covered_cnt += new_covered;
		ret.push_back(path);
	}
	return ret;
}

// shrink the cover to minimum width by augmenting a flow with lower bound 1 per node (AlignmentGraph.cpp:1157-1265)
inline std::vector<std::vector<size_t>> shrinkCover(const SplitGraph& g, size_t cid, const std::vector<std::vector<size_t>>& pc)
{
	const std::vector<size_t>& cids = g.component_ids[cid];
	LL N = (LL)cids.size();
	std::vector<std::vector<size_t>> ret;
	LL inf = (LL)pc.size();
	std::vector<LL> covered(N, 0), starts(N, 0), ends(N, 0);
	std::map<std::pair<LL, LL>, LL> edge_covered;
	for (const auto& path : pc)
	{
		for (size_t i = 0; i < path.size(); i++)
		{
			covered[g.component_idx[path[i]]]++;
			if (i > 0) edge_covered[{ (LL)g.component_idx[path[i - 1]], (LL)g.component_idx[path[i]] }]++;
		}
		starts[g.component_idx[path[0]]]++;
		ends[g.component_idx[path.back()]]++;
	}
	// adjacency-list flow network: nodes 0..N-1 = "in" halves, N..2N-1 = "out" halves, S = 2N, T = 2N+1
	LL FN = 2 * N + 2, S = 2 * N, T = 2 * N + 1;
	std::vector<LL> head(FN, 0), to(2), next(2), cap(2);
	auto add_edge = [&](LL i, LL j, LL c) { to.push_back(j); next.push_back(head[i]); cap.push_back(c); head[i] = (LL)next.size() - 1; };
	auto add = [&](LL i, LL j, LL capacity, LL lower, LL flow) { add_edge(i, j, flow - lower); add_edge(j, i, capacity - flow); };
	for (LL i = 0; i < N; i++)
		for (size_t jid : g.outNeighbors[cids[i]])
		{
			LL j = (LL)g.component_idx[jid];
			auto f = edge_covered.find({ i, j });
			add(i + N, j, inf, 0, f == edge_covered.end() ? 0 : f->second);
		}
	for (LL i = 0; i < N; i++)
	{
		add(i, i + N, inf, 1, covered[i]);
		add(S, i, inf, 0, starts[i]);
		add(i + N, T, inf, 0, ends[i]);
	}
	LL total = inf;
	std::vector<LL> Q(FN, 0), pre(FN, -1), d(FN, 0);
	while (true)
	{
		LL Qsize = 0;
		Q[Qsize++] = S;
		for (LL i = 0; i < FN; i++) { pre[i] = -1; d[i] = 0; }
		d[S] = 1;
		for (LL idx = 0; idx < Qsize && d[T] == 0; )
		{
			LL i = Q[idx++];
			for (LL e = head[i]; e; e = next[e])
			{
				LL j = to[e];
				if (cap[e] > 0 && d[j] == 0) { d[j] = 1; pre[j] = e; Q[Qsize++] = j; }
			}
		}
		if (d[T] == 0) break;
		LL flow = cap[pre[T]];
		for (LL i = T; ; ) { LL e = pre[i]; if (e == -1) break; flow = std::min(flow, cap[e]); i = to[e ^ 1]; }
		for (LL i = T; ; ) { LL e = pre[i]; if (e == -1) break; cap[e] -= flow; cap[e ^ 1] += flow; i = to[e ^ 1]; }
		if (flow == 0) throw std::runtime_error("MPC shrink: zero augmenting flow");
		total -= flow;
	}
	for (LL itr = 0; itr < total; itr++)
	{
		std::vector<size_t> tmp;
		for (LL i = S; i != T; )
		{
			if (0 <= i && i < N) tmp.push_back(cids[i]);
			LL nxt = -1;
			for (LL e = head[i]; e; e = next[e])
			{
				LL j = to[e];
				LL ff = cap[e] + ((i < N && i + N == j) ? 1 : 0);
				if ((e & 1) == 0 && ff > 0) { nxt = j; cap[e]--; break; }
			}
			if (nxt == -1) return ret;
			i = nxt;
		}
		ret.push_back(tmp);
	}
	return ret;
}

inline void computeMpcIndex(SplitGraph& g, size_t cid, const std::vector<std::vector<size_t>>& pc)
{
	const std::vector<size_t>& cids = g.component_ids[cid];
	size_t N = cids.size();
	LL K = (LL)pc.size();
	g.backwards[cid].assign(N, {});
	g.paths[cid].assign(N, {});
	// last2reach[x][k]: index on path k of the last node that reaches x; kept flat, N*K
	std::vector<LL> last2reach(N * (size_t)K, -1);
	for (LL i = 0; i < K; i++)
		for (size_t j = 0; j < pc[i].size(); j++)
		{
			size_t x = g.component_idx[pc[i][j]];
			last2reach[x * K + i] = (LL)j;
			g.paths[cid][x].push_back((size_t)i);
		}
	std::vector<LL> incd(N, 0), Q;
	for (size_t i = 0; i < N; i++) { incd[i] = (LL)g.inNeighbors[cids[i]].size(); if (incd[i] == 0) Q.push_back((LL)i); }
	g.topo_ids[cid].assign(N, 0);
	size_t topoCount = 0;
	for (size_t i = 0; i < Q.size(); )
	{
		LL s = Q[i++];
		for (size_t tid : g.outNeighbors[cids[s]])
		{
			size_t t = g.component_idx[tid];
			incd[t]--;
			if (incd[t] == 0) Q.push_back((LL)t);
		}
		g.topo_ids[cid][s] = topoCount++;
	}
	for (LL i : Q)
		for (size_t jid : g.outNeighbors[cids[i]])
		{
			size_t j = g.component_idx[jid];
			for (LL k = 0; k < K; k++) last2reach[j * K + k] = std::max(last2reach[j * K + k], last2reach[i * K + k]);
		}
	for (size_t i = 0; i < N; i++)
		for (LL k = 0; k < K; k++)
		{
			LL idx = last2reach[i * K + k];
			if (idx != -1 && g.component_idx[pc[k][idx]] == i) idx--;
			if (idx != -1) g.backwards[cid][i].push_back({ g.component_idx[pc[k][idx]], (size_t)k });
		}
}

inline void buildMpc(SplitGraph& g, bool verbose)
{
	buildComponentsMap(g);
	size_t C = g.component_ids.size();
	g.mpc.resize(C); g.topo_ids.resize(C); g.paths.resize(C); g.backwards.resize(C);
	size_t tw = 0, mw = 0;
	for (size_t cid = 0; cid < C; cid++)
	{
		g.mpc[cid] = greedyCover(g, cid);
		size_t greedy = g.mpc[cid].size();
		g.mpc[cid] = shrinkCover(g, cid, g.mpc[cid]);
		computeMpcIndex(g, cid, g.mpc[cid]);
		if (verbose) std::cout << "cid = " << cid << " greedy width " << greedy << " optimal width " << g.mpc[cid].size() << std::endl;
		tw += g.mpc[cid].size(); mw = std::max(mw, g.mpc[cid].size());
	}
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
	// initMaxCount (MinimizerSeeder.cpp:557-575).  The reference leaves out ONE key (the last
	// minimal-perfect-hash index); which one depends on BBHash internals and cannot change the
	// quantile unless that key's count sits exactly on the cut, so all keys are counted here.
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
	SplitGraph g = loadGfa(gfaPath);
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
