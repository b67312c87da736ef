// TEST INFRASTRUCTURE ONLY (oracle build) -- not part of the product.
//
// Stage-level dumper driven by the UNMODIFIED reference objects.  It links the same
// reference translation units as GraphChainer_ref and
//   (1) writes the reference's graph / MPC / minimizer index arrays to a flat
//       ".gcidx" file (the arrays a maintainer would hand to libgcgpu, see
//       include/gcgpu.h and INTEGRATION.md), and
//   (2) re-drives the per-read body of runComponentMappings (Aligner.cpp:492-1062)
//       through the reference's own seams -- MinimizerSeeder::getSeeds, OrderSeeds,
//       GraphAlignerBitvectorBanded::getReverseTraceFromSeed, AlignOneWay,
//       AlignmentGraph::colinearChaining / getChainPath, edlibAlign -- recording
//       every stage's inputs and outputs as text records.
// The driver loop here is a restatement (needed to observe intermediate values);
// each read's final alignments are cross-checked against the reference's own
// public AlignOneWay() results inside this program, and the whole-program output
// of GraphChainer_ref (unmodified Aligner.cpp) is the final arbiter in tests/.
//
// Private members of AlignmentGraph / MinimizerSeeder / GraphAligner are read by
// compiling THIS translation unit with `private` re-defined; the reference objects
// themselves are untouched (access specifiers do not change layout).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <string_view>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <phmap.h>
#include <concurrentqueue.h>
#include <zstr.hpp>
#include "vg.pb.h"
#include "google/protobuf/util/json_util.h"

#define private public
#define protected public
#include "AlignmentGraph.h"
#include "MinimizerSeeder.h"
#include "GraphAligner.h"
#undef private
#undef protected

#include "fastqloader.h"
#include "BigraphToDigraph.h"
#include "AlignmentSelection.h"
#include "MummerSeeder.h"
#include "ref_params.h"
#include <edlib.h>

// defined in the reference's Aligner.cpp:1067 (external linkage)
AlignmentGraph getGraph(std::string graphFile, MummerSeeder** mxmSeeder, const AlignerParams& params);

typedef GraphAlignerCommon<size_t, int32_t, uint64_t> Common;
typedef Common::TraceItem TraceItem;
typedef Common::OnewayTrace OnewayTrace;

// ------------------------------------------------------------------ .gcidx writer
struct IdxWriter
{
	std::ofstream f;
	IdxWriter(const std::string& path) : f(path, std::ios::binary) { f.write("GCIDX001", 8); }
	// dtype: 1=u8 4=u32 8=u64 5=i32
	void raw(const std::string& name, uint8_t dtype, uint64_t count, const void* data, size_t bytes)
	{
		uint32_t nl = name.size();
		f.write((const char*)&nl, 4); f.write(name.data(), nl); f.write((const char*)&dtype, 1); f.write((const char*)&count, 8); f.write((const char*)data, bytes);
	}
	void u8(const std::string& n, const std::vector<uint8_t>& v) { raw(n, 1, v.size(), v.data(), v.size()); }
	void u32(const std::string& n, const std::vector<uint32_t>& v) { raw(n, 4, v.size(), v.data(), v.size() * 4); }
	void i32(const std::string& n, const std::vector<int32_t>& v) { raw(n, 5, v.size(), v.data(), v.size() * 4); }
	void u64(const std::string& n, const std::vector<uint64_t>& v) { raw(n, 8, v.size(), v.data(), v.size() * 8); }
};

template <typename T, typename V> std::vector<T> conv(const V& v) { std::vector<T> r; r.reserve(v.size()); for (auto x : v) r.push_back((T)x); return r; }

static void dumpIndex(const std::string& path, const AlignmentGraph& g, const MinimizerSeeder& seeder, const AlignerParams& params)
{
	IdxWriter w(path);
	size_t N = g.nodeLength.size();
	w.u32("nodeLength", conv<uint32_t>(g.nodeLength));
	w.u32("nodeOffset", conv<uint32_t>(g.nodeOffset));
	w.i32("nodeIDs", conv<int32_t>(g.nodeIDs));
	w.u8("reverse", conv<uint8_t>(g.reverse));
	w.u8("linearizable", conv<uint8_t>(g.linearizable));
	w.u32("componentNumber", conv<uint32_t>(g.componentNumber));
	w.u32("chainNumber", conv<uint32_t>(g.chainNumber));
	w.u64("chainApproxPos", conv<uint64_t>(g.chainApproxPos));
	w.u64("firstAmbiguous", { (uint64_t)g.firstAmbiguous });
	if (g.firstAmbiguous != N) { std::cerr << "gc_refdump: graphs with ambiguous bases are not supported by the dump" << std::endl; std::exit(2); }
	std::vector<uint64_t> seq; seq.reserve(N * 2);
	for (size_t i = 0; i < N; i++) { seq.push_back(g.nodeSequences[i][0]); seq.push_back(g.nodeSequences[i][1]); }
	w.u64("nodeSeq", seq);
	auto csr = [&](const std::string& name, const std::vector<std::vector<size_t>>& adj)
	{
		std::vector<uint32_t> start { 0 }, nbr;
		for (const auto& l : adj) { for (auto x : l) nbr.push_back((uint32_t)x); start.push_back((uint32_t)nbr.size()); }
		w.u32(name + "Start", start); w.u32(name + "Nbr", nbr);
	};
	csr("in", g.inNeighbors);
	csr("out", g.outNeighbors);
	{
		// nodeLookup in ITERATION order (that order drives the minimizer index and findChains)
		std::vector<int32_t> ids; std::vector<uint32_t> start { 0 }, nodes, sizes, nameOff { 0 }; std::vector<uint8_t> names;
		for (const auto& pair : g.nodeLookup)
		{
			ids.push_back(pair.first);
			for (auto x : pair.second) nodes.push_back((uint32_t)x);
			start.push_back((uint32_t)nodes.size());
			sizes.push_back((uint32_t)g.originalNodeSize.at(pair.first));
			std::string nm = g.OriginalNodeName(pair.first);
			names.insert(names.end(), nm.begin(), nm.end());
			nameOff.push_back((uint32_t)names.size());
		}
		w.i32("origIds", ids); w.u32("origStart", start); w.u32("origNodes", nodes); w.u32("origSize", sizes); w.u32("origNameOff", nameOff); w.u8("origNames", names);
	}
	{
		w.u32("compMap", conv<uint32_t>(g.component_map));
		w.u32("compIdx", conv<uint32_t>(g.component_idx));
		std::vector<uint32_t> compStart { 0 }, compIds, topoIds, width, pathsStart { 0 }, pathsK, backStart { 0 }, backNode, backK;
		std::vector<uint32_t> mpcStart { 0 }, mpcPathStart { 0 }, mpcNodes;
		for (size_t c = 0; c < g.component_ids.size(); c++)
		{
			for (auto x : g.component_ids[c]) compIds.push_back((uint32_t)x);
			compStart.push_back((uint32_t)compIds.size());
			for (auto x : g.topo_ids[c]) topoIds.push_back((uint32_t)x);
			width.push_back((uint32_t)g.mpc[c].size());
			for (size_t i = 0; i < g.component_ids[c].size(); i++)
			{
				for (auto k : g.paths[c][i]) pathsK.push_back((uint32_t)k);
				pathsStart.push_back((uint32_t)pathsK.size());
				for (auto b : g.backwards[c][i]) { backNode.push_back((uint32_t)b.first); backK.push_back((uint32_t)b.second); }
				backStart.push_back((uint32_t)backNode.size());
			}
			for (const auto& p : g.mpc[c]) { for (auto x : p) mpcNodes.push_back((uint32_t)x); mpcPathStart.push_back((uint32_t)mpcNodes.size()); }
			mpcStart.push_back((uint32_t)(mpcPathStart.size() - 1));
		}
		w.u32("compStart", compStart); w.u32("compIds", compIds); w.u32("topoIds", topoIds); w.u32("mpcWidth", width);
		w.u32("pathsStart", pathsStart); w.u32("pathsK", pathsK); w.u32("backStart", backStart); w.u32("backNode", backNode); w.u32("backK", backK);
		w.u32("mpcStart", mpcStart); w.u32("mpcPathStart", mpcPathStart); w.u32("mpcNodes", mpcNodes);
	}
	{
		// minimizer index: per bucket, keys in MPHF index order with their position lists
		std::vector<uint32_t> bucketStart { 0 }; std::vector<uint64_t> kmers; std::vector<uint32_t> kstart { 0 }; std::vector<uint64_t> positions;
		for (size_t b = 0; b < seeder.buckets.size(); b++)
		{
			size_t nk = seeder.buckets[b].locator->nbKeys();
			for (size_t i = 0; i < nk; i++)
			{
				kmers.push_back(seeder.buckets[b].kmerCheck[i]);
				size_t s = seeder.buckets[b].startPos[i], e = seeder.buckets[b].startPos[i + 1];
				for (size_t p = s; p < e; p++) positions.push_back(seeder.buckets[b].positions[p]);
				kstart.push_back((uint32_t)positions.size());
			}
			bucketStart.push_back((uint32_t)kmers.size());
		}
		w.u32("mzBucketStart", bucketStart); w.u64("mzKmers", kmers); w.u32("mzKmerStart", kstart); w.u64("mzPositions", positions);
		w.u64("mzParams", { (uint64_t)seeder.minimizerLength, (uint64_t)seeder.windowSize, (uint64_t)seeder.maxCount, (uint64_t)seeder.buckets.size() });
	}
	w.u64("bpSize", { (uint64_t)g.bpSize });
}

// ------------------------------------------------------------------ stage dump helpers
static std::ostream* SD = nullptr; // stage dump stream (may be null)

static void dumpTraceItems(std::ostream& o, const std::vector<TraceItem>& t)
{
	o << t.size();
	for (const auto& it : t) o << ' ' << it.DPposition.node << ',' << it.DPposition.nodeOffset << ',' << (long long)it.DPposition.seqPos << ',' << (it.nodeSwitch ? 1 : 0) << ',' << it.sequenceCharacter << it.graphCharacter;
}

static void dumpSeeds(std::ostream& o, const char* tag, const std::vector<SeedHit>& seeds)
{
	o << tag << ' ' << seeds.size() << '\n';
	for (const auto& s : seeds) o << "S " << s.nodeID << ' ' << s.nodeOffset << ' ' << s.seqPos << ' ' << s.matchLen << ' ' << (s.reverse ? 1 : 0) << ' ' << s.alignmentGraphNodeId << ' ' << s.alignmentGraphNodeOffset << ' ' << s.rawSeedGoodness << ' ' << s.seedGoodness << ' ' << s.seedClusterSize << '\n';
}

struct Harness
{
	const AlignmentGraph& graph;
	const MinimizerSeeder& seeder;
	AlignerParams params;
	Common::AlignerGraphsizedState state;
	size_t extCounter = 0;
	Harness(const AlignmentGraph& g, const MinimizerSeeder& s, const AlignerParams& p) : graph(g), seeder(s), params(p), state(g, std::max(p.initialBandwidth, p.rampBandwidth), !p.highMemory) {}

	// one K1 work item through the reference kernel, recorded
	OnewayTrace extend(const Common::Params& gp, const GraphAligner<size_t, int32_t, uint64_t>& aligner, const char* stage, size_t fragL, size_t seedIdx, char dir, const std::string_view& seq, int bigraphNodeId, size_t nodeOffset)
	{
		OnewayTrace t = aligner.bvAligner.getReverseTraceFromSeed(seq, bigraphNodeId, nodeOffset, gp.forceGlobal, gp.Xdropcutoff, state);
		if (SD)
		{
			*SD << "EXT " << stage << ' ' << fragL << ' ' << seedIdx << ' ' << dir << ' ' << bigraphNodeId << ' ' << nodeOffset << ' ' << (seq.size() ? std::string(seq) : std::string("-")) << '\n';
			if (t.failed()) *SD << "RES F\n"; else { *SD << "RES " << t.score << ' '; dumpTraceItems(*SD, t.trace); *SD << '\n'; }
		}
		extCounter++;
		return t;
	}

	// restatement of GraphAligner::AlignOneWay(seeds,l,r,offset) (GraphAligner.h:114-203) calling the
	// reference's own helpers; each extension is recorded.  Checked against the public wrapper below.
	AlignmentResult alignOneWay(const char* stage, const std::string& seq_id, const std::string& sequence, bool sloppy, const std::vector<SeedHit>& seedHits, size_t l, size_t r, size_t offset)
	{
		Common::Params gp { params.initialBandwidth, params.rampBandwidth, graph, params.maxCellsPerSlice, true, sloppy, !params.highMemory, params.forceGlobal, params.preciseClipping, params.seedClusterMinSize, params.seedExtendDensity, params.nondeterministicOptimizations, params.preciseClippingIdentityCutoff, params.Xdropcutoff };
		GraphAligner<size_t, int32_t, uint64_t> aligner { gp };
		AlignmentResult result;
		result.readName = seq_id;
		size_t seedScoreForEndToEndAln = 0;
		std::string revSequence = CommonUtils::ReverseComplement(sequence);
		for (size_t i = l; i < seedHits.size() && i < r; i++)
		{
			if (sloppy && seedHits[i].seedGoodness < seedScoreForEndToEndAln) break;
			SeedHit seed = seedHits[i];
			seed.seqPos -= offset;
			if (seed.seedClusterSize < gp.minSeedClusterSize) continue;
			if (sloppy)
			{
				bool found = false;
				for (const auto& aln : result.alignments)
					if (aln.alignmentStart <= seed.seqPos && aln.alignmentEnd >= seed.seqPos && aln.seedGoodness > seed.seedGoodness) { found = true; break; }
				if (found) continue;
			}
			bool found = false;
			for (const auto& aln : result.alignments) if (aligner.exactAlignmentPart(aln, seed)) { found = true; break; }
			if (found) continue;
			result.seedsExtended += 1;
			// getTwoDirectionalTrace (GraphAligner.h:480-525)
			int forwardNodeId = seed.nodeID * 2 + (seed.reverse ? 1 : 0);
			int backwardNodeId = seed.nodeID * 2 + (seed.reverse ? 0 : 1);
			Common::Trace trace;
			trace.backward.score = std::numeric_limits<int32_t>::max();
			trace.forward.score = std::numeric_limits<int32_t>::max();
			if (seed.seqPos > 0)
			{
				std::string_view backwardPart { revSequence.data() + revSequence.size() - seed.seqPos, seed.seqPos };
				auto reversePos = graph.GetReversePosition(forwardNodeId, seed.nodeOffset);
				trace.backward = extend(gp, aligner, stage, offset, i, 'B', backwardPart, backwardNodeId, reversePos.second);
			}
			if (seed.seqPos < sequence.size() - 1)
			{
				std::string_view forwardPart { sequence.data() + seed.seqPos + 1, sequence.size() - seed.seqPos - 1 };
				trace.forward = extend(gp, aligner, stage, offset, i, 'F', forwardPart, forwardNodeId, seed.nodeOffset);
			}
			if (!trace.backward.failed()) std::reverse(trace.backward.trace.begin(), trace.backward.trace.end());
			if (!trace.forward.failed()) std::reverse(trace.forward.trace.begin(), trace.forward.trace.end());
			// getAlignmentFromSeed (GraphAligner.h:567-626)
			aligner.fixReverseTraceSeqPosAndOrder(trace.backward.trace, seed.seqPos - 1, sequence);
			aligner.fixForwardTraceSeqPos(trace.forward.trace, seed.seqPos + 1, sequence);
			if (trace.forward.failed() && trace.backward.failed()) continue;
			auto mergedTrace = std::move(trace.backward);
			if (mergedTrace.failed()) mergedTrace = std::move(trace.forward);
			else if (!trace.forward.failed())
			{
				mergedTrace.trace.pop_back();
				mergedTrace.trace.insert(mergedTrace.trace.end(), trace.forward.trace.begin(), trace.forward.trace.end());
				mergedTrace.score += trace.forward.score;
			}
			AlignmentResult::AlignmentItem item { std::move(mergedTrace), 0, std::numeric_limits<size_t>::max() };
			item.alignmentScore = item.trace->score;
			item.alignmentStart = item.trace->trace[0].DPposition.seqPos;
			item.alignmentEnd = item.trace->trace.back().DPposition.seqPos + 1;
			if (item.alignmentFailed()) continue;
			item.seedGoodness = seed.seedGoodness;
			result.alignments.emplace_back(std::move(item));
			if (sloppy)
			{
				std::sort(result.alignments.begin(), result.alignments.end(), [](const AlignmentResult::AlignmentItem& left, const AlignmentResult::AlignmentItem& right) { return left.alignmentStart < right.alignmentStart; });
				if (result.alignments[0].alignmentStart == 0)
				{
					size_t minSeedGoodness = result.alignments[0].seedGoodness;
					size_t contiguousEnd = result.alignments[0].alignmentEnd;
					for (size_t k = 1; k < result.alignments.size(); k++)
					{
						if (result.alignments[k].alignmentStart <= contiguousEnd)
						{
							minSeedGoodness = std::min(minSeedGoodness, result.alignments[k].seedGoodness);
							contiguousEnd = std::max(contiguousEnd, result.alignments[k].alignmentEnd);
						}
					}
					if (contiguousEnd == sequence.size()) seedScoreForEndToEndAln = minSeedGoodness;
				}
			}
		}
		return result;
	}

	static bool sameAlignments(const AlignmentResult& a, const AlignmentResult& b)
	{
		if (a.alignments.size() != b.alignments.size() || a.seedsExtended != b.seedsExtended) return false;
		for (size_t i = 0; i < a.alignments.size(); i++)
		{
			const auto& x = a.alignments[i]; const auto& y = b.alignments[i];
			if (x.alignmentStart != y.alignmentStart || x.alignmentEnd != y.alignmentEnd || x.alignmentScore != y.alignmentScore || x.seedGoodness != y.seedGoodness) return false;
			if (x.trace->trace.size() != y.trace->trace.size()) return false;
			for (size_t j = 0; j < x.trace->trace.size(); j++)
			{
				const auto& p = x.trace->trace[j]; const auto& q = y.trace->trace[j];
				if (p.DPposition != q.DPposition || p.nodeSwitch != q.nodeSwitch || p.sequenceCharacter != q.sequenceCharacter || p.graphCharacter != q.graphCharacter) return false;
			}
		}
		return true;
	}

	void dumpAlignments(const char* tag, const AlignmentResult& res)
	{
		if (!SD) return;
		*SD << tag << ' ' << res.alignments.size() << ' ' << res.seedsExtended << '\n';
		for (const auto& a : res.alignments)
		{
			*SD << "A " << a.alignmentStart << ' ' << a.alignmentEnd << ' ' << (long long)a.alignmentScore << ' ' << a.seedGoodness << ' ' << a.trace->score << ' ';
			dumpTraceItems(*SD, a.trace->trace);
			*SD << '\n';
		}
	}

	// per-read body, following Aligner.cpp:492-1062 (colinear mode)
	void processRead(const std::string& seq_id, const std::string& sequence, std::ostream* gamOut, size_t& mismatchCount)
	{
		std::string short_id;
		for (char c : seq_id) { if (isspace(c)) break; short_id += c; }
		if (SD) *SD << "READ " << seq_id << ' ' << sequence << '\n';
		AlignmentSelection::SelectionOptions selectionOptions;
		selectionOptions.method = params.alignmentSelectionMethod;
		selectionOptions.graphSize = graph.SizeInBP();
		selectionOptions.ECutoff = params.selectionECutoff;
		selectionOptions.EValueCalc = EValueCalculator { .7 };
		selectionOptions.readSize = sequence.size();

		AlignmentResult alignments;
		AlignmentResult long_alignments;
		size_t long_edit_distance = 0;
		// ---- S0/S1
		{
			std::vector<SeedHit> seeds = seeder.getSeeds(sequence, params.minimizerSeedDensity);
			if (SD) dumpSeeds(*SD, "SEEDS_RAW", seeds);
			if (seeds.size() > 0)
			{
				OrderSeeds(graph, seeds);
				if (SD) dumpSeeds(*SD, "SEEDS_ORDERED", seeds);
				long_alignments = alignOneWay("S1", seq_id, sequence, true, seeds, 0, seeds.size(), 0);
				AlignmentResult check = AlignOneWay(graph, seq_id, sequence, params.initialBandwidth, params.rampBandwidth, params.maxCellsPerSlice, true, true, seeds, state, !params.highMemory, params.forceGlobal, params.preciseClipping, params.seedClusterMinSize, params.seedExtendDensity, params.nondeterministicOptimizations, params.preciseClippingIdentityCutoff, params.Xdropcutoff);
				if (!sameAlignments(long_alignments, check)) { mismatchCount++; std::cerr << "gc_refdump: S1 restatement differs from reference AlignOneWay for " << seq_id << std::endl; }
			}
		}
		dumpAlignments("GA_ALL", long_alignments);
		AlignmentSelection::SelectionOptions gaSelectionOptions = selectionOptions;
		gaSelectionOptions.method = AlignmentSelection::SelectionMethod::GreedyLength;
		if (long_alignments.alignments.size() > 0) long_alignments.alignments = AlignmentSelection::SelectAlignments(long_alignments.alignments, gaSelectionOptions);
		dumpAlignments("GA_SELECTED", long_alignments);
		if (!long_alignments.alignments.empty())
		{
			// traceToSequence (Aligner.cpp:425) is a free function in Aligner.cpp
			extern std::string traceToSequence(const AlignmentGraph& alignmentGraph, AlignmentResult::AlignmentItem &aln);
			std::string long_pathseq = traceToSequence(graph, long_alignments.alignments[0]);
			EdlibAlignResult result = edlibAlign(long_pathseq.c_str(), long_pathseq.length(), sequence.c_str(), sequence.length(), edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_DISTANCE, NULL, 0));
			long_edit_distance = (result.status != EDLIB_STATUS_OK) ? sequence.length() : result.editDistance;
			edlibFreeAlignResult(result);
			if (SD) *SD << "GA_PATHSEQ " << long_pathseq << ' ' << long_edit_distance << '\n';
		}
		// ---- S2
		std::vector<AlignmentGraph::Anchor> A;
		std::vector<std::vector<TraceItem>> Apos;
		std::vector<SeedHit> seeds = seeder.getSeeds(sequence, params.minimizerSeedDensity);
		if (seeds.size() == 0) { if (SD) *SD << "FINAL NOSEEDS\n"; return; }
		OrderSeeds(graph, seeds);
		std::sort(seeds.begin(), seeds.end(), [](const SeedHit& left, const SeedHit& right) { return left.seqPos < right.seqPos; });
		if (SD) dumpSeeds(*SD, "SEEDS_BYPOS", seeds);
		size_t len = params.colinearSplitLen, sep = params.colinearSplitGap;
		size_t sl = 0, sr = 0;
		for (size_t l = 0; l + len <= sequence.length(); l += sep)
		{
			while (sr < seeds.size() && seeds[sr].seqPos + seeds[sr].matchLen <= l + len) sr++;
			while (sl < sr && seeds[sl].seqPos < l) sl++;
			if (sl >= sr) continue;
			std::string seq = sequence.substr(l, len);
			std::string name = short_id + "_" + std::to_string(l) + "_" + std::to_string(l + len - 1);
			alignments = alignOneWay("S2", name, seq, !params.tryAllSeeds, seeds, sl, sr, l);
			AlignmentResult check = AlignOneWay(graph, name, seq, params.initialBandwidth, params.rampBandwidth, params.maxCellsPerSlice, true, !params.tryAllSeeds, seeds, state, !params.highMemory, params.forceGlobal, params.preciseClipping, params.seedClusterMinSize, params.seedExtendDensity, params.nondeterministicOptimizations, params.preciseClippingIdentityCutoff, params.Xdropcutoff, sl, sr, l);
			if (!sameAlignments(alignments, check)) { mismatchCount++; std::cerr << "gc_refdump: S2 restatement differs from reference AlignOneWay for " << name << std::endl; }
			for (size_t i = 0; i < alignments.alignments.size(); i++)
			{
				AlignmentGraph::Anchor anchor = {{}, l, l + len - 1};
				AlignmentResult::AlignmentItem& alignment = alignments.alignments[i];
				if (alignment.alignmentFailed()) continue;
				auto trace = alignment.trace->trace;
				if (trace.size() == 0) continue;
				for (size_t j = 0; j < trace.size(); j++)
				{
					size_t node = graph.GetUnitigNode(trace[j].DPposition.node, trace[j].DPposition.nodeOffset);
					if (anchor.path.empty() || node != anchor.path.back()) anchor.path.push_back(node);
				}
				A.push_back(anchor);
				Apos.push_back({ trace[0], trace.back() });
				for (size_t j = 0; j < Apos.back().size(); j++)
				{
					AlignmentGraph::MatrixPosition &p = Apos.back()[j].DPposition;
					p.seqPos += l;
					p.node = graph.GetUnitigNode(p.node, p.nodeOffset);
					p.nodeOffset -= graph.NodeOffset(p.node);
				}
			}
		}
		if (SD)
		{
			*SD << "ANCHORS " << A.size() << '\n';
			for (size_t i = 0; i < A.size(); i++)
			{
				*SD << "AN " << A[i].x << ' ' << A[i].y << ' ' << Apos[i][0].DPposition.node << ' ' << Apos[i][0].DPposition.nodeOffset << ' ' << Apos[i][1].DPposition.node << ' ' << Apos[i][1].DPposition.nodeOffset << ' ' << A[i].path.size();
				for (auto n : A[i].path) *SD << ' ' << n;
				*SD << '\n';
			}
		}
		// ---- S3
		std::vector<size_t> ids = graph.colinearChaining(A, params.colinearGap);
		if (SD) { *SD << "CHAIN " << ids.size(); for (auto i : ids) *SD << ' ' << i; *SD << '\n'; }
		// ---- S4 (Aligner.cpp:738-822)
		alignments.alignments.clear();
		OnewayTrace trace;
		std::vector<AlignmentGraph::MatrixPosition> longest, tmp;
		std::vector<size_t> pos_path;
		std::unordered_set<size_t> nodes;
		size_t firstNodeOffset = 0, lastNodeOffset = 0;
		extern std::vector<AlignmentGraph::MatrixPosition> pathToTrace(const AlignmentGraph& alignmentGraph, const std::vector<size_t> &path, size_t firstNodeOffset, size_t lastNodeOffset);
		for (size_t ai : ids)
		{
			const AlignmentGraph::Anchor &anchor = A[ai];
			if (pos_path.empty())
			{
				pos_path = anchor.path;
				firstNodeOffset = Apos[ai][0].DPposition.nodeOffset;
				lastNodeOffset = Apos[ai].back().DPposition.nodeOffset;
				for (size_t j : pos_path) nodes.insert(j);
			}
			else
			{
				bool gap = anchor.path[0] == pos_path.back() && params.colinearGap != -1 && (long long)Apos[ai][0].DPposition.nodeOffset - (long long)lastNodeOffset > params.colinearGap + 1;
				std::vector<size_t> path;
				if (!nodes.count(anchor.path[0]) && pos_path.back() != Apos[ai][0].DPposition.node)
				{
					long long gapLimit = params.colinearGap;
					if (gapLimit != -1) gapLimit -= (long long)Apos[ai][0].DPposition.nodeOffset + (long long)(graph.NodeLength(pos_path.back()) - (long long)lastNodeOffset - 1);
					path = graph.getChainPath(pos_path.back(), Apos[ai][0].DPposition.node, gapLimit);
					if (SD) { *SD << "CHAINPATH " << pos_path.back() << ' ' << Apos[ai][0].DPposition.node << ' ' << gapLimit << ' ' << path.size(); for (auto n : path) *SD << ' ' << n; *SD << '\n'; }
					if (path.empty()) gap = true;
				}
				if (gap)
				{
					tmp = pathToTrace(graph, pos_path, firstNodeOffset, lastNodeOffset);
					if (longest.size() < tmp.size()) longest.swap(tmp);
					nodes.clear();
					pos_path.clear();
					firstNodeOffset = Apos[ai][0].DPposition.nodeOffset;
				}
				else
					for (size_t j : path) if (!nodes.count(j)) { nodes.insert(j); pos_path.push_back(j); }
				for (size_t j : anchor.path) if (!nodes.count(j)) { nodes.insert(j); pos_path.push_back(j); }
				lastNodeOffset = Apos[ai].back().DPposition.nodeOffset;
			}
		}
		if (!pos_path.empty())
		{
			tmp = pathToTrace(graph, pos_path, firstNodeOffset, lastNodeOffset);
			if (longest.size() < tmp.size()) longest.swap(tmp);
		}
		std::string pathseq = "";
		for (AlignmentGraph::MatrixPosition &p : longest) pathseq.push_back(graph.NodeSequences(p.node, p.nodeOffset));
		if (SD)
		{
			*SD << "LONGEST " << longest.size();
			// run-length form: node,firstOffset,count
			for (size_t i = 0; i < longest.size(); )
			{
				size_t j = i;
				while (j + 1 < longest.size() && longest[j+1].node == longest[i].node && longest[j+1].nodeOffset == longest[j].nodeOffset + 1) j++;
				*SD << ' ' << longest[i].node << ',' << longest[i].nodeOffset << ',' << (j - i + 1);
				i = j + 1;
			}
			*SD << '\n';
			*SD << "PATHSEQ " << (pathseq.empty() ? "-" : pathseq) << '\n';
		}
		// ---- S5 (Aligner.cpp:832-878)
		size_t alnScore = 0;
		{
			EdlibAlignResult result = edlibAlign(pathseq.c_str(), pathseq.length(), sequence.c_str(), sequence.length(), edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_PATH, NULL, 0));
			if (result.status != EDLIB_STATUS_OK) { longest.clear(); if (SD) *SD << "EDLIB ERR\n"; }
			else
			{
				alnScore = result.editDistance;
				if (SD)
				{
					*SD << "EDLIB " << result.editDistance << ' ' << result.alignmentLength << ' ' << (result.numLocations > 0 ? result.startLocations[0] : -1) << ' ' << (result.numLocations > 0 ? result.endLocations[0] : -1) << ' ';
					for (int j = 0; j < result.alignmentLength; j++) *SD << (char)('0' + result.alignment[j]);
					*SD << '\n';
				}
				std::vector<AlignmentGraph::MatrixPosition> tr;
				tr.reserve(result.alignmentLength);
				size_t pos_i = 0, seq_i = result.startLocations[0];
				for (size_t j = 0; j < (size_t)result.alignmentLength; j++)
				{
					AlignmentGraph::MatrixPosition p(longest[pos_i].node, longest[pos_i].nodeOffset, seq_i);
					tr.push_back(p);
					unsigned char c = result.alignment[j];
					if (c == 0 || c == 3) { pos_i++; seq_i++; }
					else if (c == 1) pos_i++;
					else if (c == 2) seq_i++;
					seq_i = std::min(seq_i, sequence.length() - 1);
					pos_i = std::min(pos_i, longest.size() - 1);
				}
				longest.swap(tr);
			}
			edlibFreeAlignResult(result);
		}
		for (size_t i = 0; i < longest.size(); i++)
		{
			bool nodeSwitch = false;
			if (i + 1 < longest.size() && longest[i].node != longest[i + 1].node) nodeSwitch = true;
			trace.trace.emplace_back(longest[i], nodeSwitch, sequence, graph);
			AlignmentGraph::MatrixPosition &p = trace.trace.back().DPposition;
			p.nodeOffset += graph.NodeOffset(p.node);
			p.node = graph.NodeID(p.node);
		}
		if (trace.trace.size() > 0)
		{
			AlignmentResult::AlignmentItem result { std::move(trace), 0, std::numeric_limits<size_t>::max() };
			result.alignmentScore = alnScore;
			result.alignmentStart = result.trace->trace[0].DPposition.seqPos;
			result.alignmentEnd = result.trace->trace.back().DPposition.seqPos + 1;
			alignments.alignments.push_back(result);
		}
		bool better = false;
		if (alignments.alignments.size() > 0)
		{
			alignments.alignments = AlignmentSelection::SelectAlignments(alignments.alignments, selectionOptions);
			better = (long_alignments.alignments.empty() || long_edit_distance > alignments.alignments.front().alignmentScore);
		}
		dumpAlignments("CLC", alignments);
		if (!better) alignments = std::move(long_alignments);
		if (SD) *SD << "DECISION " << (better ? "CLC" : "GA") << ' ' << alnScore << ' ' << long_edit_distance << '\n';
		if (alignments.alignments.size() == 0) { if (SD) *SD << "FINAL NONE\n"; return; }
		std::sort(alignments.alignments.begin(), alignments.alignments.end(), [](const AlignmentResult::AlignmentItem& left, const AlignmentResult::AlignmentItem& right) { return left.alignmentStart < right.alignmentStart; });
		for (size_t i = 0; i < alignments.alignments.size(); i++)
		{
			AddAlignment(seq_id, sequence, alignments.alignments[i]);
			replaceDigraphNodeIdsWithOriginalNodeIds(*alignments.alignments[i].alignment, graph);
		}
		if (SD)
		{
			*SD << "FINAL " << alignments.alignments.size() << '\n';
			google::protobuf::util::JsonPrintOptions options;
			for (size_t i = 0; i < alignments.alignments.size(); i++)
			{
				std::string s;
				google::protobuf::util::MessageToJsonString(*alignments.alignments[i].alignment, &s, options);
				*SD << "J " << s << '\n';
			}
		}
	}
};

int main(int argc, char** argv)
{
	std::vector<std::string> extra;
	AlignerParams params = gcParseArgs(argc, argv, extra);
	std::string indexOut, stagesOut;
	size_t maxReads = std::numeric_limits<size_t>::max();
	for (size_t i = 0; i < extra.size(); i++)
	{
		if (extra[i] == "--gc-index" && i + 1 < extra.size()) indexOut = extra[++i];
		else if (extra[i] == "--gc-stages" && i + 1 < extra.size()) stagesOut = extra[++i];
		else if (extra[i] == "--gc-max-reads" && i + 1 < extra.size()) maxReads = std::stoull(extra[++i]);
		else { std::cerr << "unknown option " << extra[i] << std::endl; return 1; }
	}
	if (params.graphFile == "") { std::cerr << "usage: gc_refdump -g graph.gfa [-f reads.fa] [-t N] [--gc-index out.gcidx] [--gc-stages out.txt] [--gc-max-reads N]" << std::endl; return 1; }
	MummerSeeder* mummerseeder = nullptr;
	auto alignmentGraph = getGraph(params.graphFile, &mummerseeder, params);
	alignmentGraph.buildMPC();
	MinimizerSeeder seeder(alignmentGraph, params.minimizerLength, params.minimizerWindowSize, params.numThreads, 1.0 - params.minimizerDiscardMostNumerousFraction);
	if (indexOut != "") dumpIndex(indexOut, alignmentGraph, seeder, params);
	if (params.fastqFiles.empty() || stagesOut == "") return 0;
	std::ofstream stages(stagesOut);
	SD = &stages;
	Harness h(alignmentGraph, seeder, params);
	size_t count = 0, mismatches = 0;
	for (auto filename : params.fastqFiles)
	{
		FastQ::streamFastqFromFile(filename, false, [&](FastQ& read)
		{
			if (count >= maxReads) return;
			count++;
			try { h.processRead(read.seq_id, read.sequence, nullptr, mismatches); }
			catch (const ThreadReadAssertion::AssertionFailure& a) { stages << "FINAL ASSERTION\n"; h.state.clear(); std::cerr << "gc_refdump: assertion in read " << read.seq_id << std::endl; }
		});
	}
	std::cerr << "gc_refdump: " << count << " reads, " << h.extCounter << " extensions, " << mismatches << " restatement mismatches" << std::endl;
	return mismatches == 0 ? 0 : 3;
}
