// Host driver of the per-read pipeline: the body of runComponentMappings
// (src/Aligner.cpp:492-1062, colinear mode) re-organised for batches of reads.  The DP stages and
// everything that reads their traces run on the device through libgcgpu's C ABI; the K1 traces of a
// batch never leave HBM (include/gcgpu.h, "resident batch"):
//   S0  seeding                            k-mer lookups gcgpu_seed; count sort, density cut, clustering and
//                                          the goodness sort on the host (gc_seeder.h: libstdc++ tie orders)
//   S1  whole-read seed-and-extend         host seed loop (GraphAligner.h:114-203) in ROUNDS over alignment
//                                          extents + seed-coverage bit masks; extensions and exactAlignmentPart
//                                          = gcgpu_extend_seeds (K1 + coverage kernel)
//   S1b distance(GA path, read)            path string built on the device (gcgpu_nw_compose), gcgpu_nw (K3)
//   S2  fragment anchoring                 gcgpu_fragment_anchors: all window seeds extended speculatively,
//                                          the in-order seed loop per fragment and the anchors on the device
//   S3  co-linear chaining                 gcgpu_chain_resident (K2) on the device-resident anchors
//   S4  chain -> node path                 host BFS getChainPath over the chained anchors only
//   S5  NW(path, read)                     gcgpu_nw (K3): distance for every read, the edit path only when
//                                          the chained alignment wins (S6)
//   S6  decision                           host                              Aligner.cpp:880-920
//   S7  vg::Alignment                      edit runs from gcgpu_encode_alignments; protobuf framing on the host
// There is no CPU implementation of the device stages here: without libgcgpu nothing aligns.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>
#ifdef GC_PROF
#include <x86intrin.h>
static double g_prof[32]; static const char* g_profName[32];
struct GcProfScope { int id; unsigned long long t0; GcProfScope(int id, const char* n) : id(id), t0(__rdtsc()) { g_profName[id] = n; } ~GcProfScope() { g_prof[id] += (double)(__rdtsc() - t0); } };
#define GC_PROF_SCOPE(id, name) GcProfScope _prof##id(id, name)
#else
#define GC_PROF_SCOPE(id, name)
#endif

#include "../../include/gcgpu.h"
#include "gc_host_graph.h"
#include "gc_seeder.h"
#include "gc_post.cuh"

struct GcRead
{
	std::string name;     // full FASTA/FASTQ header after '>' / '@'
	std::string sequence;
};

// one alignment of the final output: AlignmentResult::AlignmentItem with its trace already reduced to the
// mappings and edit runs of the vg::Alignment (token stream of gc_post.cuh)
struct GcAlnItem
{
	std::vector<uint32_t> tokens;
	uint32_t matches = 0, steps = 0; // identity = matches / steps
	int32_t traceScore = 0;       // OnewayTrace::score (what AddAlignment serialises)
	size_t alignmentStart = 0, alignmentEnd = 0;
	size_t alignmentScore = 0;
	size_t seedGoodness = 0;
};

struct GcReadResult
{
	std::vector<GcAlnItem> alignments; // final, sorted by alignmentStart
	bool usedChain = false;            // S6: the chained (CLC) alignment was strictly better
	bool dropped = false;              // assertion-class failure in the whole-read pass: the reference drops the read
	bool broke = false;                // any assertion-class failure (the run reports "Alignment broke with some reads")
	// the fields of the reference's --short-verbose line (Aligner.cpp:909-915)
	size_t anchors = 0, chained = 0, pathBp = 0, clcScore = 0, longEditDistance = 0;
	std::string gamRecord;   // the read's gzip member when the record was made on the device (gamOnDevice); its alignments then carry no tokens
	int oneNodeOverlapsNow = 0, oneNodeOverlapsAll = 0; // the reference's progress-line counters (Aligner.cpp:747,768-771,792,820)
	bool hasLong = false;
	size_t seedsFound = 0, seedsExtended = 0;
	size_t alignmentsBeforeSelection = 0; // --no-colinear-chaining: stats.allAlignmentsCount counts the whole-read alignments before GreedyLength (Aligner.cpp:924-930)
};

struct GcPipelineParams
{
	double minimizerSeedDensity = 10;
	size_t seedClusterMinSize = 1;
	long long colinearGap = 10000;
	long long colinearSplitLen = 35;
	long long colinearSplitGap = 35;
	bool tryAllSeeds = true;
	// --no-colinear-chaining (AlignerMain.cpp:108,198-199): "align as in GraphAligner" -- the whole-read pass only (align_fn,
	// Aligner.cpp:596-600), its alignments selected by GreedyLength (:927-930); no fragments, chain or NW distances
	bool colinearChaining = true;
	// The reference's progress line (--short-verbose, Aligner.cpp:909-915) prints, as "actual N bps", the LENGTH OF THE EDIT PATH of the
	// chained alignment (`longest` is swapped with the converted trace, :875) -- of every read, also where the whole-read alignment is
	// written.  The edit path is otherwise only computed where the chain wins (the decision needs the distance only); with this set it
	// is computed for every read so that the line is the reference's.
	bool exactProgressLine = false;
	// the GAM records of the reads whose whole-read alignments are written are encoded and compressed on the device
	// (gcgpu_encode_gam): GcReadResult::gamRecord.  Set by callers that write GAM only (JSON / GAF need the token streams here).
	bool gamOnDevice = false;
	size_t s1FirstRoundSeeds = 1;   // S1 speculation: seeds extended per read in the first round ...
	size_t s1LaterRoundSeeds = 8;   // ... in the second round ...
	size_t s1TailRoundSeeds = 32;  // ... and from the third round on: few reads get that far, their rounds cost one extension latency each whatever the item count
};

struct GcPipelineStats
{
	uint64_t k1Items = 0, k1Columns = 0, k1Launches = 0;
	uint64_t k3Items = 0, k3Blocks = 0;
	uint64_t k2Reads = 0, k2Anchors = 0;
	double k1Ms = 0, k2Ms = 0, k3Ms = 0, s0Ms = 0;
	double hostSeedMs = 0, hostS1Ms = 0, hostS2Ms = 0, hostConnectMs = 0;
	uint64_t s1Rounds = 0;
	uint64_t s1Wasted = 0; // speculative S1 seed extensions whose result was discarded
};

// hands the reference's minimizer index (MinimizerSeeder.h:17-29 as flat arrays) to a libgcgpu context
// GFA segment names of the original nodes, for the GAM records made on the device (gcgpu_encode_gam)
inline int gcUploadNodeNames(gcgpu_ctx* ctx, const GcHostGraph& g)
{
	std::vector<uint32_t> off(g.origNames.size() + 1, 0);
	std::string chars;
	for (size_t i = 0; i < g.origNames.size(); i++) { chars += g.origNames[i]; off[i + 1] = (uint32_t)chars.size(); }
	return gcgpu_set_node_names(ctx, off.data(), chars.data());
}

inline int gcUploadMinimizerIndex(gcgpu_ctx* ctx, const GcHostGraph& g)
{
	gcgpu_minimizer_index mi;
	mi.k = (uint32_t)g.mzLength; mi.window = (uint32_t)g.mzWindow; mi.max_count = g.mzMaxCount;
	mi.num_kmers = g.mzKmers.size(); mi.kmers = g.mzKmers.data(); mi.kmer_start = g.mzKmerStart.data();
	return gcgpu_set_minimizer_index(ctx, &mi);
}

namespace gcpipe {

inline char complementChar(char c)
{
	switch (c)
	{
		case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'T': case 't': return 'A'; case 'G': case 'g': return 'C';
		case 'N': case 'n': return 'N'; case 'U': case 'u': return 'A'; case 'R': case 'r': return 'Y'; case 'Y': case 'y': return 'R';
		case 'K': case 'k': return 'M'; case 'M': case 'm': return 'K'; case 'S': case 's': return 'S'; case 'W': case 'w': return 'W';
		case 'B': case 'b': return 'V'; case 'V': case 'v': return 'B'; case 'D': case 'd': return 'H'; case 'H': case 'h': return 'D';
	}
	return 0; // the reference asserts (CommonUtils.cpp:131)
}
inline uint8_t complementMask(uint8_t m) { return (uint8_t)(((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3)); }

// Common::characterMatch (GraphAlignerCommon.h:193-217) restricted to graphs without ambiguous bases
inline bool characterMatch(char sequenceCharacter, char graphCharacter)
{
	if (sequenceCharacter == graphCharacter) return true;
	uint8_t m = gcEncodeBase(sequenceCharacter);
	int b = graphCharacter == 'A' ? 0 : graphCharacter == 'C' ? 1 : graphCharacter == 'G' ? 2 : graphCharacter == 'T' ? 3 : -1;
	if (sequenceCharacter == '-' || b < 0) return false;
	return (m >> b) & 1;
}

// AlignmentSelection::alignmentIncompatible (AlignmentSelection.cpp:13-31)
template <typename Aln>
inline bool alignmentIncompatible(const Aln& left, const Aln& right)
{
	const float OverlapIncompatibleFractionCutoff = 0.05;
	auto minOverlapLen = std::min((left.alignmentEnd - left.alignmentStart), (right.alignmentEnd - right.alignmentStart)) * OverlapIncompatibleFractionCutoff;
	size_t leftStart = left.alignmentStart, leftEnd = left.alignmentEnd, rightStart = right.alignmentStart, rightEnd = right.alignmentEnd;
	if (leftStart > rightStart) { std::swap(leftStart, rightStart); std::swap(leftEnd, rightEnd); }
	int overlap = 0;
	if (leftEnd > rightStart) overlap = leftEnd - rightStart;
	return overlap > minOverlapLen;
}
// GreedySelectAlignments with alignmentLengthCompare (AlignmentSelection.h:36-55, .cpp:45-51)
template <typename Aln>
inline std::vector<Aln> selectGreedyLength(const std::vector<Aln>& alignments)
{
	std::vector<size_t> items;
	for (size_t i = 0; i < alignments.size(); i++) items.push_back(i);
	std::sort(items.begin(), items.end(), [&alignments](size_t l, size_t r)
	{
		const Aln& left = alignments[l]; const Aln& right = alignments[r];
		if ((left.alignmentEnd - left.alignmentStart) > (right.alignmentEnd - right.alignmentStart)) return true;
		if ((right.alignmentEnd - right.alignmentStart) > (left.alignmentEnd - left.alignmentStart)) return false;
		if (left.alignmentScore < right.alignmentScore) return true;
		return false;
	});
	std::vector<Aln> result;
	for (auto i : items)
	{
		if (!std::any_of(result.begin(), result.end(), [&alignments, i](const Aln& existing) { return alignmentIncompatible(existing, alignments[i]); }))
			result.push_back(alignments[i]);
	}
	return result;
}

struct MatrixPos { size_t node; size_t nodeOffset; size_t seqPos; };

// pathToTrace (Aligner.cpp:409-424), including its single-node quirk
inline std::vector<MatrixPos> pathToTrace(const GcHostGraph& g, const std::vector<size_t>& path, size_t firstNodeOffset, size_t lastNodeOffset)
{
	std::vector<MatrixPos> ret;
	for (size_t node : path)
	{
		size_t S = 0, L = g.nodeLength[node];
		if (node == path[0]) S = firstNodeOffset;
		else if (node == path.back()) L = lastNodeOffset + 1;
		MatrixPos p { node, S, 0 };
		while (p.nodeOffset < L) { ret.push_back(p); p.nodeOffset++; }
	}
	return ret;
}

// AlignmentGraph::getChainPath (AlignmentGraph.cpp:1866-1916): FIFO BFS with the unsigned distance prune
struct ChainPathScratch { std::vector<size_t> vis, dis, Q, pre; size_t flag = 1; };
inline std::vector<size_t> getChainPath(const GcHostGraph& g, ChainPathScratch& s, size_t S, size_t T, long long sep_limit)
{
	size_t N = g.numNodes();
	if (s.vis.size() < N) { s.vis.resize(N, 0); s.pre.resize(N); s.dis.resize(N); s.Q.reserve(1024); }
	s.Q.clear();
	s.Q.push_back(S);
	s.vis[S] = ++s.flag;
	s.dis[S] = 0;
	for (size_t i = 0; s.vis[T] != s.flag && i < s.Q.size(); )
	{
		size_t v = s.Q[i++];
		if (s.dis[v] > (size_t)sep_limit) continue; // size_t vs long long comparison, AlignmentGraph.cpp:1897
		for (uint32_t e = g.outStart[v]; e < g.outStart[v + 1]; e++)
		{
			size_t t = g.outNbr[e];
			if (s.vis[t] != s.flag)
			{
				s.Q.push_back(t);
				s.vis[t] = s.flag;
				s.dis[t] = s.dis[v] + g.nodeLength[t];
				s.pre[t] = v;
			}
		}
	}
	std::vector<size_t> tmp;
	if (s.vis[T] != s.flag) return tmp;
	for (size_t i = T; i != S; i = s.pre[i]) tmp.push_back(i);
	tmp.push_back(S);
	std::reverse(tmp.begin(), tmp.end());
	return tmp;
}

}

class GcPipeline
{
public:
	GcPipeline(const GcHostGraph& graph, gcgpu_ctx* ctx, const GcPipelineParams& params) : g(graph), ctx(ctx), params(params) {}
	GcPipelineStats stats;

	void alignBatch(const std::vector<GcRead>& reads, std::vector<GcReadResult>& out);
	void setGamOnDevice(bool on) { params.gamOnDevice = on; }

private:
	const GcHostGraph& g;
	gcgpu_ctx* ctx;
	GcPipelineParams params;

	// page-locked host buffers for what crosses PCIe, grow-only, reused across batches
	struct Pinned
	{
		void* p = nullptr; size_t cap = 0;
		void ensure(size_t bytes)
		{
			if (bytes <= cap) return;
			if (p) gcgpu_host_free(p);
			cap = bytes + bytes / 4 + (1 << 20);
			p = gcgpu_host_alloc(cap);
			if (!p) { cap = 0; throw std::runtime_error("gcgpu_host_alloc failed"); }
		}
		~Pinned() { if (p) gcgpu_host_free(p); }
		Pinned() = default;
		Pinned(const Pinned&) = delete;
		Pinned& operator=(const Pinned&) = delete;
	};
	Pinned charsBuf, seedBuf, cellBuf, extBuf, tokenBuf, opsBuf, gamBuf;
	std::vector<std::unique_ptr<Pinned>> coverPool; // seed-coverage masks of the S1 rounds of a batch

	void stats_s1Wasted_add(size_t n) { if (n) { _Pragma("omp atomic") stats.s1Wasted += n; } }
	void check(int rc, const char* what)
	{
		if (rc != GCGPU_OK) throw std::runtime_error(std::string(what) + " failed: " + gcgpu_last_error());
	}
};

// ------------------------------------------------------------------------------------------
inline void GcPipeline::alignBatch(const std::vector<GcRead>& reads, std::vector<GcReadResult>& out)
{
	size_t R = reads.size();
	out.assign(R, GcReadResult());
	if (R == 0) return;
	const bool traceOn = getenv("GC_TRACE") != nullptr;
	auto wallNow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double tPhase = wallNow();
	double devMs = 0; // wall time spent inside libgcgpu calls during the current phase
	auto phase = [&](const char* name) { if (traceOn) { double n = wallNow(); fprintf(stderr, "[gc] phase %-10s %.2f ms (libgcgpu calls %.2f ms, host %.2f ms)\n", name, n - tPhase, devMs, n - tPhase - devMs); tPhase = n; devMs = 0; } };
	// ---- the reads' characters go to the device once; codes (forward + reverse complement) are derived there
	std::vector<gcgpu_read> rd(R);
	uint64_t totalChars = 0;
	for (size_t r = 0; r < R; r++) { rd[r].char_offset = totalChars; rd[r].len = (int32_t)reads[r].sequence.size(); rd[r].first_cell = 0; rd[r].num_cells = 0; rd[r].reserved = 0; totalChars += reads[r].sequence.size(); }
	charsBuf.ensure(totalChars + 16);
	// A character outside the IUPAC alphabet makes the reference's CommonUtils::Complement assert (CommonUtils.cpp:131-133) when
	// AlignOneWay reverse-complements the read (GraphAligner.h:124): the read is dropped (Aligner.cpp:585-592).
	std::vector<uint8_t> invalidChars(R, 0);
	#pragma omp parallel for schedule(dynamic, 16)
	for (size_t r = 0; r < R; r++)
	{
		const std::string& sq = reads[r].sequence;
		memcpy((char*)charsBuf.p + rd[r].char_offset, sq.data(), sq.size());
		uint8_t all = 0xFF;
		for (char ch : sq) all &= (uint8_t)(gcEncodeBase(ch) ? 0xFF : 0);
		invalidChars[r] = all ? 0 : 1;
	}
	{ double tDev = wallNow(); check(gcgpu_load_reads(ctx, (const char*)charsBuf.p, totalChars, rd.data(), (uint32_t)R), "gcgpu_load_reads"); devMs += wallNow() - tDev; stats.s0Ms += gcgpu_last_kernel_ms(ctx); }
	// ---- S0: seeds (the reference calls getSeeds + OrderSeeds twice per read with identical results).
	// k-mer walk + index probes on the device; the count sort, density cut, seed-hit expansion and clustering per read on the host
	std::vector<std::vector<GcSeedHit>> seedsOrdered(R), seedsByPos(R);
	{
		std::vector<gcgpu_seed_read> sr(R);
		for (size_t r = 0; r < R; r++) { sr[r].seq_offset = 2 * rd[r].char_offset; sr[r].seq_len = rd[r].len; sr[r].reserved = 0; }
		std::vector<uint64_t> matchOff(R + 1, 0);
		uint64_t used = 0;
		double tDev = wallNow();
		check(gcgpu_seed(ctx, nullptr, 2 * totalChars, sr.data(), (uint32_t)R, matchOff.data(), nullptr, 0, &used), "gcgpu_seed");
		seedBuf.ensure((used + 1) * sizeof(gcgpu_seed_match));
		check(gcgpu_fetch_seed_matches(ctx, (gcgpu_seed_match*)seedBuf.p, 0, used), "gcgpu_fetch_seed_matches");
		devMs += wallNow() - tDev;
		stats.s0Ms += gcgpu_last_kernel_ms(ctx);
		const gcgpu_seed_match* matches = (const gcgpu_seed_match*)seedBuf.p;
		#pragma omp parallel
		{
			gcseed::Scratch scratch;
			#pragma omp for schedule(dynamic, 4)
			for (size_t r = 0; r < R; r++)
			{
				GC_PROF_SCOPE(0, "seed.seedRead");
				// count sort + density cut + expansion, clustering, goodness order and the split pass's position order (the device
				// holds the seeds in position order: "cells")
				gcseed::seedRead(g, matches + matchOff[r], (size_t)(matchOff[r + 1] - matchOff[r]), reads[r].sequence.size(), params.minimizerSeedDensity, scratch, seedsOrdered[r], seedsByPos[r]);
				out[r].seedsFound = (params.colinearChaining ? 2 : 1) * seedsOrdered[r].size(); // counted in align_fn and again for the split pass (Aligner.cpp:553,663)
			}
		}
	}
	{
		uint64_t totalCells = 0;
		for (size_t r = 0; r < R; r++) { rd[r].first_cell = (uint32_t)totalCells; rd[r].num_cells = (uint32_t)seedsByPos[r].size(); totalCells += seedsByPos[r].size(); }
		if (totalCells >= 0xFFFFFFFFull) throw std::runtime_error("too many seeds in one batch");
		cellBuf.ensure((totalCells + 1) * sizeof(gcgpu_seed_cell));
		gcgpu_seed_cell* cells = (gcgpu_seed_cell*)cellBuf.p;
		#pragma omp parallel for schedule(dynamic, 16)
		for (size_t r = 0; r < R; r++)
		{
			const std::vector<GcSeedHit>& byPos = seedsByPos[r];
			for (size_t i = 0; i < byPos.size(); i++)
			{
				gcgpu_seed_cell& c = cells[rd[r].first_cell + i];
				c.seq_pos = (int32_t)byPos[i].seqPos; c.node = (uint32_t)byPos[i].alignmentGraphNodeId; c.read = (uint32_t)r; c.offset = (uint8_t)byPos[i].alignmentGraphNodeOffset;
				c.flags = byPos[i].seedClusterSize >= params.seedClusterMinSize ? 1 : 0; c.reserved = 0;
			}
		}
		double tDev = wallNow();
		check(gcgpu_set_seed_cells(ctx, cells, totalCells, rd.data(), (uint32_t)R), "gcgpu_set_seed_cells");
		devMs += wallNow() - tDev;
	}
	phase("seed");

	// ---- S1: whole-read alignment, AlignOneWay(seeds, sloppy=true) (GraphAligner.h:114-203).
	// The reference walks the seeds in goodness order and decides, from the alignments kept so far,
	// whether to extend each one.  An extension is a pure function of (read, seed), so every ROUND
	// extends, for every read, the next few seeds that pass the skip rules under the current state;
	// the results are then consumed strictly in seed order with the rules re-evaluated exactly as
	// the reference does -- a speculative result whose seed turns out to be skipped is discarded.
	// The host sees an extension as (start, end, score) + one bit per seed of the read: "this seed's
	// cell lies on the trace" (exactAlignmentPart, evaluated on the device for all seeds at once).
	struct S1Aln
	{
		uint32_t pair = 0;               // index in trace set 0
		const uint32_t* cover = nullptr; // bit i: seed i (position order) lies on the trace
		int32_t traceScore = 0;
		size_t alignmentStart = 0, alignmentEnd = 0, alignmentScore = 0, seedGoodness = 0;
		bool degenerate() const { return !(alignmentEnd - 1 > alignmentStart); } // exactAlignmentPart asserts trace.back().seqPos > trace[0].seqPos
		bool covers(const GcSeedHit& seed) const { return (cover[seed.byPosIdx >> 5] >> (seed.byPosIdx & 31)) & 1; }
	};
	struct S1Cand { size_t seedIdx; uint32_t ext; };
	struct S1State
	{
		size_t i = 0; std::vector<S1Aln> alns; size_t seedsExtended = 0; size_t seedScoreForEndToEndAln = 0; bool done = false; std::vector<S1Cand> cands; size_t round = 0;
		size_t extBase = 0;
		// memo of the skip rules: alignments are only ever added and never change, so "this seed is skipped" is permanent and
		// "no alignment so far contains this seed cell" only needs the alignments added since it was last asked
		std::vector<S1Aln> alnsAdded;         // insertion order (alns is kept sorted by alignmentStart like the reference's vector)
		std::vector<uint8_t> skip;            // per seed: 1 = a skip rule fired
		std::vector<uint32_t> checked;        // per seed: alnsAdded[0..checked) do not contain its cell
		bool degenerate = false;              // an alignment on which exactAlignmentPart asserts exists: evaluate in reference order
	};
	std::vector<S1State> s1(R);
	for (size_t r = 0; r < R; r++) { if (seedsOrdered[r].empty()) s1[r].done = true; else if (invalidChars[r]) { s1[r].done = true; out[r].dropped = true; } else { s1[r].skip.assign(seedsOrdered[r].size(), 0); s1[r].checked.assign(seedsOrdered[r].size(), 0); } }
	// 0 = extend, 1 = skip, 2 = stop the seed loop, 3 = assertion (read dropped)
	auto seedRule = [&](S1State& st, const GcSeedHit& seed, size_t idx) -> int
	{
		if (seed.seedGoodness < st.seedScoreForEndToEndAln) return 2;
		if (seed.seedClusterSize < params.seedClusterMinSize) return 1;
		if (!st.degenerate)
		{
			if (st.skip[idx]) return 1;
			for (const auto& aln : st.alns)
				if (aln.alignmentStart <= seed.seqPos && aln.alignmentEnd >= seed.seqPos && aln.seedGoodness > seed.seedGoodness) { st.skip[idx] = 1; return 1; }
			for (uint32_t k = st.checked[idx]; k < st.alnsAdded.size(); k++)
				if (st.alnsAdded[k].covers(seed)) { st.skip[idx] = 1; return 1; }
			st.checked[idx] = (uint32_t)st.alnsAdded.size();
			return 0;
		}
		for (const auto& aln : st.alns)
			if (aln.alignmentStart <= seed.seqPos && aln.alignmentEnd >= seed.seqPos && aln.seedGoodness > seed.seedGoodness) return 1;
		for (const auto& aln : st.alns) { if (aln.degenerate()) return 3; if (aln.covers(seed)) return 1; }
		return 0;
	};
	std::vector<gcgpu_seed_ext> exts;
	std::vector<gcgpu_pair_brief> brief;
	std::vector<uint64_t> coverOff;
	size_t s1Calls = 0;
	while (true)
	{
		exts.clear();
		std::vector<size_t> active;
		// candidates are collected per read in parallel, then concatenated
		#pragma omp parallel for schedule(dynamic, 8)
		for (size_t r = 0; r < R; r++)
		{
			S1State& st = s1[r];
			if (st.done) continue;
			GC_PROF_SCOPE(3, "s1.collect");
			const std::vector<GcSeedHit>& seedHits = seedsOrdered[r];
			size_t want = st.round == 0 ? params.s1FirstRoundSeeds : (st.round == 1 ? params.s1LaterRoundSeeds : params.s1TailRoundSeeds);
			st.cands.clear();
			for (size_t i = st.i; i < seedHits.size() && st.cands.size() < want; i++)
			{
				int rule = seedRule(st, seedHits[i], i);
				if (rule >= 2) break; // decided again, in order, when the results are consumed
				if (rule == 1) continue;
				S1Cand c; c.seedIdx = i; c.ext = 0;
				st.cands.push_back(c);
			}
			st.round++;
			if (st.cands.empty())
			{
				// nothing left to extend: finish the seed walk for the assertion-class exit only
				for (; st.i < seedHits.size(); st.i++) { int rule = seedRule(st, seedHits[st.i], st.i); if (rule == 3) out[r].dropped = true; if (rule >= 2) break; }
				st.done = true;
			}
		}
		coverOff.clear();
		uint64_t coverWords = 0;
		for (size_t r = 0; r < R; r++)
		{
			S1State& st = s1[r];
			if (st.done || st.cands.empty()) continue;
			st.extBase = exts.size();
			for (S1Cand& c : st.cands)
			{
				c.ext = (uint32_t)exts.size();
				gcgpu_seed_ext e; e.cell = rd[r].first_cell + seedsOrdered[r][c.seedIdx].byPosIdx; e.frag_start = -1;
				exts.push_back(e);
				coverOff.push_back(coverWords);
				coverWords += (rd[r].num_cells + 31) / 32;
			}
			active.push_back(r);
		}
		if (active.empty()) break;
		coverOff.push_back(coverWords);
		stats.s1Rounds++;
		brief.resize(exts.size());
		if (coverPool.size() <= s1Calls) coverPool.emplace_back(new Pinned());
		Pinned& coverBuf = *coverPool[s1Calls];
		coverBuf.ensure((coverWords + 1) * 4);
		const uint32_t* cover = (const uint32_t*)coverBuf.p;
		uint32_t firstPair = 0; uint64_t cols = 0;
		{
			double tDev = wallNow();
			check(gcgpu_extend_seeds(ctx, 0, s1Calls > 0 ? 1 : 0, 0, exts.data(), (uint32_t)exts.size(), brief.data(), (uint32_t*)coverBuf.p, coverOff.data(), &firstPair, &cols), "gcgpu_extend_seeds");
			devMs += wallNow() - tDev;
			s1Calls++;
			stats.k1Items += 2 * exts.size(); stats.k1Ms += gcgpu_last_kernel_ms(ctx); stats.k1Launches++; stats.k1Columns += cols;
			if (traceOn) fprintf(stderr, "[gc] extend seeds=%zu kernel_ms=%.3f columns=%llu cover_KB=%.1f\n", exts.size(), (double)gcgpu_last_kernel_ms(ctx), (unsigned long long)cols, coverWords * 4 / 1e3);
		}
		#pragma omp parallel for schedule(dynamic, 4)
		for (size_t k = 0; k < active.size(); k++)
		{
			size_t r = active[k];
			GC_PROF_SCOPE(4, "s1.consume");
			S1State& st = s1[r];
			const std::vector<GcSeedHit>& seedHits = seedsOrdered[r];
			size_t next = 0; // next unconsumed speculative result
			for (; st.i < seedHits.size(); st.i++)
			{
				const GcSeedHit& seed = seedHits[st.i];
				int rule = seedRule(st, seed, st.i);
				if (rule == 3) { out[r].dropped = true; st.i = seedHits.size(); break; }
				if (rule == 2) { st.i = seedHits.size(); break; }
				if (rule == 1) continue;
				while (next < st.cands.size() && st.cands[next].seedIdx < st.i) next++;
				if (next >= st.cands.size() || st.cands[next].seedIdx != st.i) break; // not extended yet: first seed of the next round
				const S1Cand& c = st.cands[next++];
				st.seedsExtended += 1;
				const gcgpu_pair_brief& b = brief[c.ext];
				if (b.flags & GCGPU_PAIR_INTERNAL) out[r].dropped = true;
				if (!(b.flags & (GCGPU_PAIR_BWD | GCGPU_PAIR_FWD)) || b.end == b.start) continue;
				S1Aln item;
				item.pair = firstPair + c.ext; item.cover = cover + coverOff[c.ext];
				item.traceScore = b.score; item.alignmentScore = (size_t)b.score; item.alignmentStart = (size_t)b.start; item.alignmentEnd = (size_t)b.end; item.seedGoodness = seed.seedGoodness;
				st.alns.emplace_back(item);
				st.alnsAdded.emplace_back(item);
				if (item.degenerate()) st.degenerate = true;
				std::sort(st.alns.begin(), st.alns.end(), [](const S1Aln& left, const S1Aln& right) { return left.alignmentStart < right.alignmentStart; });
				if (st.alns[0].alignmentStart == 0)
				{
					size_t minSeedGoodness = st.alns[0].seedGoodness;
					size_t contiguousEnd = st.alns[0].alignmentEnd;
					for (size_t i = 1; i < st.alns.size(); i++)
					{
						if (st.alns[i].alignmentStart <= contiguousEnd)
						{
							minSeedGoodness = std::min(minSeedGoodness, st.alns[i].seedGoodness);
							contiguousEnd = std::max(contiguousEnd, st.alns[i].alignmentEnd);
						}
					}
					if (contiguousEnd == reads[r].sequence.size()) st.seedScoreForEndToEndAln = minSeedGoodness;
				}
			}
			if (st.i >= seedHits.size()) st.done = true;
			stats_s1Wasted_add(st.cands.size() - next);
		}
	}
	phase("s1");
	// GreedyLength selection of the GA alignments (Aligner.cpp:637-641)
	std::vector<std::vector<S1Aln>> longAlns(R);
	std::vector<size_t> longSeedsExtended(R, 0);
	#pragma omp parallel for schedule(dynamic, 16)
	for (size_t r = 0; r < R; r++)
	{
		longSeedsExtended[r] = s1[r].seedsExtended;
		if (out[r].dropped) { out[r].broke = true; s1[r].alns.clear(); continue; }
		out[r].alignmentsBeforeSelection = s1[r].alns.size();
		if (!s1[r].alns.empty()) longAlns[r] = gcpipe::selectGreedyLength(s1[r].alns);
		s1[r].alns.clear();
	}

	// ---- S2: fragment anchoring (Aligner.cpp:656-730): every window seed of every fragment is extended speculatively, the
	// in-order seed loop of each fragment and the anchors are evaluated on the device
	const size_t len = (size_t)params.colinearSplitLen, sep = (size_t)params.colinearSplitGap;
	std::vector<gcgpu_read_anchors> perRead(R);
	if (R) memset(perRead.data(), 0, R * sizeof(gcgpu_read_anchors));
	if (params.colinearChaining)
	{
		std::vector<std::vector<gcgpu_frag>> localFrags(R);
		std::vector<std::vector<gcgpu_seed_ext>> localExts(R);
		#pragma omp parallel for schedule(dynamic, 8)
		for (size_t r = 0; r < R; r++)
		{
			if (seedsByPos[r].empty() || out[r].dropped) continue; // `cont` stays true after an assertion in the whole-read pass: every fragment is skipped (Aligner.cpp:700-703)
			GC_PROF_SCOPE(5, "s2.items");
			const std::vector<GcSeedHit>& seeds = seedsByPos[r];
			size_t L = reads[r].sequence.length();
			size_t sl = 0, sr = 0;
			for (size_t l = 0; l + len <= L; l += sep)
			{
				while (sr < seeds.size() && seeds[sr].seqPos + seeds[sr].matchLen <= l + len) sr++;
				while (sl < sr && seeds[sl].seqPos < l) sl++;
				if (sl >= sr) continue;
				gcgpu_frag f; f.read = (uint32_t)r; f.start = (int32_t)l; f.first_ext = (uint32_t)localExts[r].size(); f.num_exts = (uint32_t)(sr - sl);
				for (size_t i = sl; i < sr; i++) { gcgpu_seed_ext e; e.cell = rd[r].first_cell + (uint32_t)i; e.frag_start = (int32_t)l; localExts[r].push_back(e); }
				localFrags[r].push_back(f);
			}
		}
		std::vector<size_t> fragBase(R + 1, 0), extBase(R + 1, 0);
		for (size_t r = 0; r < R; r++) { fragBase[r + 1] = fragBase[r] + localFrags[r].size(); extBase[r + 1] = extBase[r] + localExts[r].size(); }
		std::vector<gcgpu_frag> frags(fragBase[R]);
		extBuf.ensure((extBase[R] + 1) * sizeof(gcgpu_seed_ext));
		gcgpu_seed_ext* fexts = (gcgpu_seed_ext*)extBuf.p;
		#pragma omp parallel for schedule(dynamic, 16)
		for (size_t r = 0; r < R; r++)
		{
			for (size_t k = 0; k < localFrags[r].size(); k++) { gcgpu_frag f = localFrags[r][k]; f.first_ext += (uint32_t)extBase[r]; frags[fragBase[r] + k] = f; }
			if (!localExts[r].empty()) memcpy(fexts + extBase[r], localExts[r].data(), localExts[r].size() * sizeof(gcgpu_seed_ext));
		}
		uint64_t cols = 0;
		double tDev = wallNow();
		check(gcgpu_fragment_anchors(ctx, 1, (int32_t)len, fexts, (uint32_t)extBase[R], frags.data(), (uint32_t)frags.size(), (uint32_t)R, perRead.data(), &cols), "gcgpu_fragment_anchors");
		devMs += wallNow() - tDev;
		stats.k1Items += 2 * extBase[R]; stats.k1Ms += gcgpu_last_kernel_ms(ctx); stats.k1Launches++; stats.k1Columns += cols;
		if (traceOn) fprintf(stderr, "[gc] fragments=%zu seeds=%zu kernel_ms=%.3f columns=%llu\n", frags.size(), extBase[R], (double)gcgpu_last_kernel_ms(ctx), (unsigned long long)cols);
	}
	phase("s2");
	// ---- S3: chaining (K2) on the anchors the device holds; only the chained anchors come back
	std::vector<uint32_t> chainLen(R);
	std::vector<int64_t> chainScore(R);
	std::vector<gcgpu_chained_anchor> chained;
	std::vector<uint32_t> chainedPaths;
	std::vector<uint64_t> chainedOff(R + 1, 0);
	if (params.colinearChaining)
	{
		uint64_t nChained = 0, nPathNodes = 0, nAnchors = 0;
		double tDev = wallNow();
		check(gcgpu_chain_resident(ctx, (uint32_t)R, chainLen.data(), chainScore.data(), &nChained, &nPathNodes), "gcgpu_chain_resident");
		stats.k2Ms += gcgpu_last_kernel_ms(ctx);
		chained.resize(nChained + 1); chainedPaths.resize(nPathNodes + 1);
		check(gcgpu_fetch_chained(ctx, chained.data(), chainedPaths.data()), "gcgpu_fetch_chained");
		devMs += wallNow() - tDev;
		for (size_t r = 0; r < R; r++) { chainedOff[r + 1] = chainedOff[r] + chainLen[r]; nAnchors += perRead[r].anchors; if (perRead[r].dropped) out[r].broke = true; }
		stats.k2Reads += R; stats.k2Anchors += nAnchors;
	}
	phase("s3");
	// ---- S4: chain -> node path (Aligner.cpp:738-831).  A candidate segment is kept as (node path, first offset,
	// last offset); its per-base position list (pathToTrace) is only materialised for the reads where the chain wins --
	// here only its LENGTH (the `longest` comparison, `<` => first wins ties) is needed; the bases are expanded on the device.
	struct PathSeg { std::vector<size_t> path; size_t firstOffset = 0, lastOffset = 0, size = 0; };
	std::vector<PathSeg> longest(R);
	// number of positions pathToTrace emits (same per-node tests by VALUE as Aligner.cpp:409-424)
	auto pathTraceSize = [&](const std::vector<size_t>& path, size_t firstNodeOffset, size_t lastNodeOffset)
	{
		size_t total = 0;
		for (size_t node : path)
		{
			size_t S = 0, L = g.nodeLength[node];
			if (node == path[0]) S = firstNodeOffset;
			else if (node == path.back()) L = lastNodeOffset + 1;
			if (L > S) total += L - S;
		}
		return total;
	};
	#pragma omp parallel
	{
		gcpipe::ChainPathScratch scratch;
		// the reference's std::unordered_set<size_t> nodes, as an epoch-stamped array that lives as long as the thread
		static thread_local std::vector<uint32_t> inPath;
		static thread_local uint32_t epoch = 0;
		if (inPath.size() != g.numNodes() || epoch > 0xF0000000u) { inPath.assign(g.numNodes(), 0); epoch = 0; }
		#pragma omp for schedule(dynamic, 4)
		for (size_t r = 0; r < R; r++)
		{
			GC_PROF_SCOPE(8, "s4.all");
			std::vector<size_t> pos_path;
			size_t firstNodeOffset = 0, lastNodeOffset = 0;
			epoch++;
			int overlapsTmp = 0;
			auto closeSegment = [&]()
			{
				size_t sz = pathTraceSize(pos_path, firstNodeOffset, lastNodeOffset);
				if (longest[r].size < sz) { longest[r].path = pos_path; longest[r].firstOffset = firstNodeOffset; longest[r].lastOffset = lastNodeOffset; longest[r].size = sz; out[r].oneNodeOverlapsNow = overlapsTmp; }
			};
			for (uint32_t ci = 0; ci < chainLen[r]; ci++)
			{
				const gcgpu_chained_anchor& anchor = chained[chainedOff[r] + ci];
				const uint32_t* apath = chainedPaths.data() + anchor.path_first;
				if (pos_path.empty())
				{
					pos_path.assign(apath, apath + anchor.path_len);
					firstNodeOffset = anchor.first_offset;
					lastNodeOffset = anchor.last_offset;
					for (size_t j : pos_path) inPath[j] = epoch;
				}
				else
				{
					if (apath[0] == pos_path.back()) { overlapsTmp++; out[r].oneNodeOverlapsAll++; }
					bool gap = apath[0] == pos_path.back() && params.colinearGap != -1 && (long long)anchor.first_offset - (long long)lastNodeOffset > params.colinearGap + 1;
					std::vector<size_t> path;
					if (inPath[apath[0]] != epoch && pos_path.back() != apath[0])
					{
						long long gapLimit = params.colinearGap;
						if (gapLimit != -1) gapLimit -= (long long)anchor.first_offset + (long long)((long long)g.nodeLength[pos_path.back()] - (long long)lastNodeOffset - 1);
						path = gcpipe::getChainPath(g, scratch, pos_path.back(), apath[0], gapLimit);
						if (path.empty()) gap = true;
					}
					if (gap)
					{
						closeSegment();
						epoch++; // nodes.clear()
						pos_path.clear();
						firstNodeOffset = anchor.first_offset;
					}
					else
						for (size_t j : path) if (inPath[j] != epoch) { inPath[j] = epoch; pos_path.push_back(j); }
					for (uint32_t k = 0; k < anchor.path_len; k++) { size_t j = apath[k]; if (inPath[j] != epoch) { inPath[j] = epoch; pos_path.push_back(j); } }
					lastNodeOffset = anchor.last_offset;
				}
			}
			if (!pos_path.empty()) closeSegment();
		}
	}
	phase("s4");
	// ---- S1b + S5: NW distances (K3), then the edit path only where the chain wins (S6).  The sequence buffer is put
	// together on the device: the read, the padded path of the first whole-read alignment, the bases of the chained path.
	std::vector<gcgpu_nw_item> nwItems;
	std::vector<int> gaItem(R, -1), clcItem(R, -1);
	uint64_t nwBufSize = 0;
	{
		std::vector<gcgpu_nw_piece> pieces;
		std::vector<uint32_t> pathNodes;
		std::vector<int> readPiece(R, -1), gaPiece(R, -1), clcPiece(R, -1);
		for (size_t r = 0; r < R; r++)
		{
			bool needGa = params.colinearChaining && !longAlns[r].empty();
			bool needClc = params.colinearChaining && !seedsOrdered[r].empty() && !out[r].dropped;
			if (!needGa && !needClc) continue;
			gcgpu_nw_piece pc; memset(&pc, 0, sizeof(pc));
			pc.kind = GCGPU_PIECE_READ; pc.index = (uint32_t)r;
			readPiece[r] = (int)pieces.size(); pieces.push_back(pc);
			if (needGa)
			{
				memset(&pc, 0, sizeof(pc));
				pc.kind = GCGPU_PIECE_PAIR_PATH; pc.set = 0; pc.index = longAlns[r][0].pair;
				gaPiece[r] = (int)pieces.size(); pieces.push_back(pc);
			}
			if (needClc && !longest[r].path.empty())
			{
				memset(&pc, 0, sizeof(pc));
				pc.kind = GCGPU_PIECE_NODE_PATH; pc.first_node = pathNodes.size(); pc.num_nodes = (uint32_t)longest[r].path.size();
				pc.first_offset = (uint32_t)longest[r].firstOffset; pc.last_offset = (uint32_t)longest[r].lastOffset;
				for (size_t node : longest[r].path) pathNodes.push_back((uint32_t)node);
				clcPiece[r] = (int)pieces.size(); pieces.push_back(pc);
			}
		}
		std::vector<uint64_t> pieceOff(pieces.size() + 1, 0);
		if (!pieces.empty())
		{
			double tDev = wallNow();
			check(gcgpu_nw_compose(ctx, pieces.data(), (uint32_t)pieces.size(), pathNodes.data(), pathNodes.size(), pieceOff.data()), "gcgpu_nw_compose");
			devMs += wallNow() - tDev;
			stats.k3Ms += gcgpu_last_kernel_ms(ctx);
		}
		nwBufSize = pieceOff[pieces.size()];
		for (size_t r = 0; r < R; r++)
		{
			if (readPiece[r] < 0) continue;
			bool needGa = gaPiece[r] >= 0;
			bool needClc = params.colinearChaining && !seedsOrdered[r].empty() && !out[r].dropped;
			size_t readLen = reads[r].sequence.size();
			// first cutoff of the NW passes: the whole-read alignment bounds its own distance from above
			// (score + unaligned read ends); the chained path normally lies at the same locus.  Only a
			// starting point -- the kernel doubles the cutoff until the pass succeeds, like edlib does from 64.
			int32_t kHint = 0, clcHint = (int32_t)(readLen / 5);
			if (needGa)
			{
				// the whole-read alignment IS a global alignment of (its padded path string, read): its score + the unaligned read
				// ends + the padding of the first and last node (< 2 x 64 graph characters, traceToPoses) bounds the distance from
				// above, so one pass at that cutoff always succeeds
				const S1Aln& a0 = longAlns[r][0];
				size_t ub = a0.alignmentScore + a0.alignmentStart + (readLen - a0.alignmentEnd);
				kHint = (int32_t)std::min<size_t>(ub + 130, (size_t)1 << 30);
				clcHint = (int32_t)std::min<size_t>(ub + ub * 3 / 10, readLen / 5);
				gcgpu_nw_item it; it.query_offset = pieceOff[gaPiece[r]]; it.target_offset = pieceOff[readPiece[r]]; it.query_len = (int32_t)(pieceOff[gaPiece[r] + 1] - pieceOff[gaPiece[r]]); it.target_len = (int32_t)readLen;
				it.k_hint = kHint; it.want_path = 0;
				gaItem[r] = (int)nwItems.size(); nwItems.push_back(it);
			}
			if (needClc)
			{
				gcgpu_nw_item it; it.target_offset = pieceOff[readPiece[r]]; it.target_len = (int32_t)readLen;
				if (clcPiece[r] >= 0) { it.query_offset = pieceOff[clcPiece[r]]; it.query_len = (int32_t)(pieceOff[clcPiece[r] + 1] - pieceOff[clcPiece[r]]); }
				else { it.query_offset = 0; it.query_len = 0; } // no chain: the empty path string
				// the chained path usually costs 5-40 % more than the whole-read alignment; the first guess is capped at a fifth of
				// the read (the widest band the two-blocks-per-lane class holds for 10 kb reads) and is also what a read without a
				// whole-read alignment starts from (edlib's 64 would cost five doubling passes before the answer fits)
				it.k_hint = clcHint; it.want_path = 0;
				clcItem[r] = (int)nwItems.size(); nwItems.push_back(it);
			}
		}
	}
	std::vector<gcgpu_nw_result> nwRes(nwItems.size());
	uint64_t opsUsed = 0;
	if (!nwItems.empty())
	{
		double tDev = wallNow();
		int rc = gcgpu_nw(ctx, nullptr, nwBufSize, nwItems.data(), (uint32_t)nwItems.size(), nwRes.data(), nullptr, 0, &opsUsed);
		if (rc != GCGPU_OK && rc != GCGPU_ERR_INTERNAL) check(rc, "gcgpu_nw");
		devMs += wallNow() - tDev;
		stats.k3Items += nwItems.size(); stats.k3Ms += gcgpu_last_kernel_ms(ctx);
		for (const auto& x : nwRes) stats.k3Blocks += x.blocks;
		if (traceOn)
		{
			size_t zero = 0, ok = 0, fail = 0; double sumHint = 0, sumD = 0;
			for (size_t i = 0; i < nwItems.size(); i++)
			{
				if (nwItems[i].k_hint <= 0) zero++;
				else if (nwRes[i].distance <= nwItems[i].k_hint) { ok++; sumHint += nwItems[i].k_hint; sumD += nwRes[i].distance; }
				else { fail++; if (fail <= 12) fprintf(stderr, "[gc]   hint %d distance %d\n", nwItems[i].k_hint, nwRes[i].distance); }
			}
			fprintf(stderr, "[gc] nw hints: none=%zu ok=%zu (mean hint %.0f, mean distance %.0f) too_small=%zu\n", zero, ok, ok ? sumHint / ok : 0.0, ok ? sumD / ok : 0.0, fail);
		}
	}
	// decision (Aligner.cpp:901-920): better = no GA alignment, or long_edit_distance > CLC score (strict)
	std::vector<gcgpu_nw_item> pathItems;
	std::vector<size_t> pathRead;
	for (size_t r = 0; r < R; r++)
	{
		GcReadResult& res = out[r];
		res.anchors = perRead[r].anchors;
		res.chained = chainLen[r];
		res.pathBp = longest[r].size;
		res.hasLong = gaItem[r] >= 0;
		// edlib's status != OK leaves long_edit_distance at the read length (Aligner.cpp:646-648)
		if (gaItem[r] >= 0) res.longEditDistance = nwRes[gaItem[r]].status == 0 ? (size_t)nwRes[gaItem[r]].distance : reads[r].sequence.size();
		if (clcItem[r] >= 0) res.clcScore = (size_t)nwRes[clcItem[r]].distance;
		res.seedsExtended = 0;
		bool haveClc = clcItem[r] >= 0 && longest[r].size != 0 && nwRes[clcItem[r]].status == 0;
		bool better = haveClc && (longAlns[r].empty() || res.longEditDistance > res.clcScore);
		res.usedChain = better;
		if (better || (params.exactProgressLine && haveClc))
		{
			gcgpu_nw_item it = nwItems[clcItem[r]];
			it.want_path = 2; it.k_hint = nwRes[clcItem[r]].distance; // the distance is known: only the edit path is computed
			pathItems.push_back(it); pathRead.push_back(r);
		}
	}
	std::vector<gcgpu_nw_result> pathRes(pathItems.size());
	const uint8_t* ops = nullptr;
	if (!pathItems.empty())
	{
		uint64_t cap = 0;
		for (const auto& it : pathItems) cap += (uint64_t)it.query_len + it.target_len + 8;
		opsBuf.ensure(cap + 16);
		double tDev = wallNow();
		// an alignment whose edit path cannot be reconstructed degrades that read only: the reference clears `longest` when
		// edlib's status is not OK and keeps the whole-read alignment (Aligner.cpp:846-848)
		int rc = gcgpu_nw(ctx, nullptr, nwBufSize, pathItems.data(), (uint32_t)pathItems.size(), pathRes.data(), (uint8_t*)opsBuf.p, cap, &opsUsed);
		if (rc != GCGPU_OK && rc != GCGPU_ERR_INTERNAL) check(rc, "gcgpu_nw(path)");
		devMs += wallNow() - tDev;
		ops = (const uint8_t*)opsBuf.p;
		stats.k3Items += pathItems.size(); stats.k3Ms += gcgpu_last_kernel_ms(ctx);
		for (const auto& x : pathRes) stats.k3Blocks += x.blocks;
	}
	phase("nw");
	// ---- S5 trace conversion (Aligner.cpp:851-897) for the reads where the chain wins
	#pragma omp parallel for schedule(dynamic, 4)
	for (size_t k = 0; k < pathRead.size(); k++)
	{
		size_t r = pathRead[k];
		if (params.exactProgressLine) out[r].pathBp = pathRes[k].status == 0 ? pathRes[k].ops_len : 0; // longest.swap(trace) / longest.clear() (Aligner.cpp:846-875)
		if (!out[r].usedChain) continue; // edit path computed for the progress line only
		if (pathRes[k].status != 0 || pathRes[k].ops_len == 0) { out[r].usedChain = false; continue; }
		const std::string& sequence = reads[r].sequence;
		const std::vector<gcpipe::MatrixPos> lg = gcpipe::pathToTrace(g, longest[r].path, longest[r].firstOffset, longest[r].lastOffset);
		const uint8_t* op = ops + pathRes[k].ops_offset;
		size_t n = pathRes[k].ops_len;
		std::vector<GcTokenStep> steps(n);
		std::vector<size_t> splitNode(n);
		size_t pos_i = 0, seq_i = 0;
		for (size_t j = 0; j < n; j++)
		{
			GcTokenStep& t = steps[j];
			size_t node = lg[pos_i].node, off = lg[pos_i].nodeOffset;
			splitNode[j] = node;
			t.seqPos = (int32_t)seq_i;
			char sequenceCharacter = seq_i < sequence.size() ? sequence[seq_i] : '-';
			t.match = gcpipe::characterMatch(sequenceCharacter, g.nodeChar((uint32_t)node, (uint32_t)off));
			t.node = g.nodeIDs[node];
			t.nodeOffset = (uint32_t)(off + g.nodeOffset[node]);
			t.nodeSwitch = false;
			unsigned char c = op[j];
			if (c == 0 || c == 3) { pos_i++; seq_i++; }
			else if (c == 1) pos_i++;
			else if (c == 2) seq_i++;
			seq_i = std::min(seq_i, sequence.length() - 1);
			pos_i = std::min(pos_i, lg.size() - 1);
		}
		// nodeSwitch compares the SPLIT nodes of consecutive entries (Aligner.cpp:880-884)
		for (size_t j = 0; j + 1 < n; j++) steps[j].nodeSwitch = splitNode[j] != splitNode[j + 1];
		GcAlnItem item;
		struct VecSrc { const std::vector<GcTokenStep>* v; GcTokenStep operator()(uint32_t i) const { return (*v)[i]; } } src { &steps };
		GcTokenCounts cnt = gc_tokenize(src, (uint32_t)n, (uint32_t*)nullptr);
		item.tokens.resize(cnt.tokens);
		gc_tokenize(src, (uint32_t)n, item.tokens.data());
		item.matches = cnt.matches; item.steps = cnt.matches + cnt.mismatches + cnt.insertions + cnt.deletions;
		item.traceScore = 0; // the reference leaves trace.score at 0 for the chained alignment
		item.alignmentScore = out[r].clcScore;
		item.alignmentStart = (size_t)steps[0].seqPos;
		item.alignmentEnd = (size_t)steps.back().seqPos + 1;
		out[r].alignments.clear();
		out[r].alignments.push_back(std::move(item));
	}
	// ---- S7 for the whole-read alignments that are written.  GAM only: the records themselves come from the device
	const bool recordsOnDevice = params.gamOnDevice;
	std::vector<uint8_t> wantTokens(R, 0); // reads whose token streams are fetched to the host
	bool tokensForSome = false;
	if (recordsOnDevice)
	{
		std::vector<gcgpu_gam_read> gr;
		std::vector<gcgpu_gam_aln> ga;
		std::vector<size_t> grRead;
		std::string names;
		for (size_t r = 0; r < R; r++)
		{
			GcReadResult& res = out[r];
			if (res.usedChain || longAlns[r].empty()) continue;
			// the record lists the alignments in the order of the final sort by alignmentStart (Aligner.cpp:1004, the same std::sort call)
			std::vector<S1Aln> sorted = longAlns[r];
			std::sort(sorted.begin(), sorted.end(), [](const S1Aln& left, const S1Aln& right) { return left.alignmentStart < right.alignmentStart; });
			gcgpu_gam_read g1; g1.read = (uint32_t)r; g1.first_aln = (uint32_t)ga.size(); g1.num_alns = (uint32_t)sorted.size(); g1.name_len = (uint32_t)reads[r].name.size(); g1.name_offset = names.size();
			names += reads[r].name;
			res.alignments.clear();
			for (const S1Aln& a : sorted)
			{
				gcgpu_gam_aln x; x.pair = a.pair; x.start = (int32_t)a.alignmentStart; x.end = (int32_t)a.alignmentEnd; x.trace_score = a.traceScore;
				ga.push_back(x);
				GcAlnItem item; item.traceScore = a.traceScore; item.alignmentScore = a.alignmentScore; item.alignmentStart = a.alignmentStart; item.alignmentEnd = a.alignmentEnd; item.seedGoodness = a.seedGoodness;
				res.alignments.push_back(std::move(item));
			}
			gr.push_back(g1); grRead.push_back(r);
		}
		std::vector<uint64_t> memberOff(gr.size() + 1, 0);
		uint64_t used = 0;
		if (!gr.empty())
		{
			double tDev = wallNow();
			check(gcgpu_encode_gam(ctx, 0, gr.data(), (uint32_t)gr.size(), ga.data(), (uint32_t)ga.size(), names.data(), names.size(), memberOff.data(), &used), "gcgpu_encode_gam");
			stats.k1Ms += gcgpu_last_kernel_ms(ctx);
			gamBuf.ensure(used + 16);
			check(gcgpu_fetch_gam(ctx, (uint8_t*)gamBuf.p, 0, used), "gcgpu_fetch_gam");
			devMs += wallNow() - tDev;
		}
		std::vector<uint8_t> needHost(R, 0);
		bool anyHost = false;
		#pragma omp parallel for schedule(dynamic, 16)
		for (size_t k = 0; k < gr.size(); k++)
		{
			size_t r = grRead[k];
			if (memberOff[k + 1] > memberOff[k]) out[r].gamRecord.assign((const char*)gamBuf.p + memberOff[k], memberOff[k + 1] - memberOff[k]);
			else { needHost[r] = 1; _Pragma("omp atomic write") anyHost = true; }
		}
		for (size_t r = 0; r < R; r++)
		{
			GcReadResult& res = out[r];
			if (!res.usedChain) res.seedsExtended = res.alignments.empty() ? 0 : longSeedsExtended[r];
			else res.seedsExtended = perRead[r].last_frag_extended;
			res.seedsExtended += perRead[r].seeds_extended;
		}
		// a record the device could not encode (no complete Huffman code for its statistics): the token path below, for those reads only
		if (anyHost) { wantTokens = needHost; tokensForSome = true; }
	}
	else { tokensForSome = true; for (size_t r = 0; r < R; r++) wantTokens[r] = out[r].usedChain ? 0 : 1; }
	if (tokensForSome)
	{
		std::vector<uint32_t> pairs;
		std::vector<size_t> firstOfRead(R + 1, 0);
		for (size_t r = 0; r < R; r++)
		{
			firstOfRead[r] = pairs.size();
			if (wantTokens[r]) for (const S1Aln& a : longAlns[r]) pairs.push_back(a.pair);
		}
		firstOfRead[R] = pairs.size();
		std::vector<gcgpu_aln_tokens> meta(pairs.size());
		uint64_t used = 0;
		if (!pairs.empty())
		{
			double tDev = wallNow();
			check(gcgpu_encode_alignments(ctx, 0, pairs.data(), (uint32_t)pairs.size(), meta.data(), &used), "gcgpu_encode_alignments");
			stats.k1Ms += gcgpu_last_kernel_ms(ctx);
			tokenBuf.ensure((used + 1) * 4);
			check(gcgpu_fetch_tokens(ctx, (uint32_t*)tokenBuf.p, 0, used), "gcgpu_fetch_tokens");
			devMs += wallNow() - tDev;
		}
		const uint32_t* tok = (const uint32_t*)tokenBuf.p;
		#pragma omp parallel for schedule(dynamic, 16)
		for (size_t r = 0; r < R; r++)
		{
			GcReadResult& res = out[r];
			if (recordsOnDevice && !wantTokens[r]) continue; // record made on the device (or chained alignment): nothing to fetch
			if (!res.usedChain)
			{
				res.alignments.clear();
				res.alignments.resize(longAlns[r].size());
				for (size_t k = 0; k < longAlns[r].size(); k++)
				{
					const S1Aln& a = longAlns[r][k];
					const gcgpu_aln_tokens& m = meta[firstOfRead[r] + k];
					GcAlnItem& item = res.alignments[k];
					item.tokens.assign(tok + m.token_offset, tok + m.token_offset + m.num_tokens);
					item.matches = m.matches; item.steps = m.steps;
					item.traceScore = a.traceScore; item.alignmentScore = a.alignmentScore; item.alignmentStart = a.alignmentStart; item.alignmentEnd = a.alignmentEnd; item.seedGoodness = a.seedGoodness;
				}
				res.seedsExtended = res.alignments.empty() ? 0 : longSeedsExtended[r]; // stats.seedsExtended += alignments.seedsExtended (Aligner.cpp:995)
			}
			else res.seedsExtended = perRead[r].last_frag_extended; // `alignments` still carries the last fragment's seedsExtended (Aligner.cpp:691,995)
			res.seedsExtended += perRead[r].seeds_extended;
			std::sort(res.alignments.begin(), res.alignments.end(), [](const GcAlnItem& left, const GcAlnItem& right) { return left.alignmentStart < right.alignmentStart; });
		}
	}
	phase("final");
#ifdef GC_PROF
	if (traceOn) for (int i = 0; i < 32; i++) if (g_profName[i]) fprintf(stderr, "[prof] %-32s %.2f ms\n", g_profName[i], g_prof[i] / 2.0e6);
#endif
	if (traceOn) fprintf(stderr, "[gc] batch reads=%zu s1_rounds=%llu s1_wasted=%llu\n", R, (unsigned long long)stats.s1Rounds, (unsigned long long)stats.s1Wasted);
}
