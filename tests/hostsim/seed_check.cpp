// seed_check <index.gcidx> <reads.fa>: S0 parity tool.  For every read it compares the device form of
// the k-mer walk + index probes (gcgpu_seed through the C ABI -- libgcgpu.so on a GPU box, the test
// double on the build box) with the sequential restatement of the reference (seed_ref.h), match by
// match, and the seed hits built from both.  Prints "reads R matches M seeds S mismatches X".
#include <cstdio>
#include <fstream>
#include <iostream>
#include "../../include/gcgpu.h"
#include "../../graphchainer_b200/csrc/gc_pipeline.h"
#include "seed_ref.h"

int main(int argc, char** argv)
{
	if (argc < 3) { fprintf(stderr, "usage: seed_check index.gcidx reads.fa [stages]\n"); return 2; }
	GcIndexFile idx; idx.load(argv[1]);
	GcHostGraph g; g.fromIndex(idx);
	std::vector<std::string> reads;
	{
		std::ifstream in(argv[2]); std::string line, cur; bool have = false;
		while (std::getline(in, line)) { if (!line.empty() && line[0] == '>') { if (have) reads.push_back(cur); cur.clear(); have = true; } else cur += line; }
		if (have) reads.push_back(cur);
	}
	gcgpu_graph gg; memset(&gg, 0, sizeof(gg));
	gg.num_nodes = (uint32_t)g.numNodes(); gg.node_length = g.nodeLength.data(); gg.node_seq = g.nodeSeq.data();
	gg.in_start = g.inStart.data(); gg.in_nbr = g.inNbr.data(); gg.out_start = g.outStart.data(); gg.out_nbr = g.outNbr.data();
	gg.component_number = g.componentNumber.data(); gg.linearizable = g.linearizable.data();
	gcgpu_params gp; gp.initial_bandwidth = 10;
	gcgpu_ctx* ctx = nullptr;
	if (gcgpu_create(0, &gg, &gp, &ctx) != GCGPU_OK) { fprintf(stderr, "gcgpu_create: %s\n", gcgpu_last_error()); return 1; }
	if (gcUploadMinimizerIndex(ctx, g) != GCGPU_OK) { fprintf(stderr, "gcgpu_set_minimizer_index: %s\n", gcgpu_last_error()); return 1; }
	std::vector<uint8_t> codes; std::vector<gcgpu_seed_read> sr(reads.size());
	for (size_t r = 0; r < reads.size(); r++)
	{
		sr[r].seq_offset = codes.size(); sr[r].seq_len = (int32_t)reads[r].size(); sr[r].reserved = 0;
		for (char c : reads[r]) codes.push_back(gcEncodeSeedBase(c));
	}
	codes.resize(codes.size() + 8, 0);
	std::vector<uint64_t> off(reads.size() + 1); uint64_t used = 0;
	if (gcgpu_seed(ctx, codes.data(), codes.size(), sr.data(), (uint32_t)sr.size(), off.data(), nullptr, 0, &used) != GCGPU_OK) { fprintf(stderr, "gcgpu_seed: %s\n", gcgpu_last_error()); return 1; }
	std::vector<gcgpu_seed_match> m(used + 1);
	if (gcgpu_fetch_seed_matches(ctx, m.data(), 0, used) != GCGPU_OK) { fprintf(stderr, "fetch: %s\n", gcgpu_last_error()); return 1; }
	uint64_t realMatches = used;
	SeedRefIndex ref(g);
	size_t mismatches = 0, seeds = 0;
	for (size_t r = 0; r < reads.size(); r++)
	{
		auto want = seedRefMatches(g, ref, reads[r]);
		std::vector<std::tuple<size_t, size_t, size_t, size_t>> got;
		for (uint64_t i = off[r]; i < off[r + 1]; i++) got.emplace_back((size_t)m[i].pos, (size_t)0, (size_t)m[i].start, (size_t)m[i].count);
		if (got != want) { mismatches++; if (mismatches <= 5) fprintf(stderr, "read %zu: %zu matches vs %zu in the restatement\n", r, got.size(), want.size()); continue; }
		seeds += (size_t)(off[r + 1] - off[r]);
	}
	// ---- against the reference's own records (oracle/_ref/gc_refdump --gc-stages): its seed vector after OrderSeeds and after the
	// split pass's sort by position, field by field -- this pins the k-mer walk, the index answers, the density cut, the
	// clustering and all three tie orders (std::sort on partial keys) to the live reference
	if (argc > 3)
	{
		std::ifstream st(argv[3]);
		std::string line;
		struct Rec { uint32_t seqPos, node, off, good, cluster, matchLen; };
		auto parseList = [&](std::vector<Rec>& out)
		{
			out.clear();
			std::streampos p = st.tellg();
			while (std::getline(st, line))
			{
				if (line.compare(0, 2, "S ") != 0) { st.seekg(p); break; }
				long nodeID, nodeOffset, seqPos, matchLen, rev, agNode, agOff, raw, good, cluster;
				sscanf(line.c_str(), "S %ld %ld %ld %ld %ld %ld %ld %ld %ld %ld", &nodeID, &nodeOffset, &seqPos, &matchLen, &rev, &agNode, &agOff, &raw, &good, &cluster);
				out.push_back(Rec { (uint32_t)seqPos, (uint32_t)agNode, (uint32_t)agOff, (uint32_t)good, (uint32_t)cluster, (uint32_t)matchLen });
				p = st.tellg();
			}
		};
		std::string seq; std::vector<Rec> wantOrdered, wantByPos; bool haveRead = false;
		size_t refReads = 0, refSeeds = 0;
		gcseed::Scratch scratch;
		auto check = [&]()
		{
			if (!haveRead) return;
			refReads++;
			std::vector<uint8_t> c; for (char ch : seq) c.push_back(gcEncodeSeedBase(ch));
			size_t len = c.size(); c.resize(c.size() + 8, 0);
			gcgpu_seed_read one; one.seq_offset = 0; one.seq_len = (int32_t)len; one.reserved = 0;
			uint64_t o2[2] = { 0, 0 }, u2 = 0;
			if (gcgpu_seed(ctx, c.data(), c.size(), &one, 1, o2, nullptr, 0, &u2) != GCGPU_OK) { mismatches++; return; }
			std::vector<gcgpu_seed_match> mm(u2 + 1);
			if (gcgpu_fetch_seed_matches(ctx, mm.data(), 0, u2) != GCGPU_OK) { mismatches++; return; }
			std::vector<GcSeedHit> ordered, byPos;
			gcseed::seedRead(g, mm.data(), u2, len, 10, scratch, ordered, byPos);
			auto same = [](const std::vector<GcSeedHit>& a, const std::vector<Rec>& b)
			{
				if (a.size() != b.size()) return false;
				for (size_t i = 0; i < a.size(); i++)
					if (a[i].seqPos != b[i].seqPos || a[i].alignmentGraphNodeId != b[i].node || a[i].alignmentGraphNodeOffset != b[i].off || a[i].seedGoodness != b[i].good || a[i].seedClusterSize != b[i].cluster || a[i].matchLen != b[i].matchLen) return false;
				return true;
			};
			refSeeds += ordered.size();
			bool ok = same(ordered, wantOrdered) && same(byPos, wantByPos);
			for (size_t i = 0; ok && i < ordered.size(); i++) ok = byPos[ordered[i].byPosIdx].orderedIdx == i && ordered[i].orderedIdx == i;
			if (!ok) { mismatches++; if (mismatches <= 5) fprintf(stderr, "read %zu of the stage records: seeds differ from the reference's (%zu vs %zu ordered)\n", refReads, ordered.size(), wantOrdered.size()); }
		};
		while (std::getline(st, line))
		{
			if (line.compare(0, 5, "READ ") == 0)
			{
				check();
				size_t sp = line.find(' ', 5);
				seq = line.substr(sp + 1); haveRead = true; wantOrdered.clear(); wantByPos.clear();
			}
			else if (line.compare(0, 13, "SEEDS_ORDERED") == 0) parseList(wantOrdered);
			else if (line.compare(0, 11, "SEEDS_BYPOS") == 0) parseList(wantByPos);
		}
		check();
		printf("ref_reads %zu ref_seeds %zu ", refReads, refSeeds);
	}
	// second pass with an index made of EVERY k-mer of the reads (count 1 each): the matches are then exactly the
	// k-mers iterateKmers emits, which pins the emission rule itself (homopolymer re-emission, restarts after N/U)
	size_t emitted = 0;
	{
		GcHostGraph fake;
		fake.mzLength = g.mzLength; fake.mzWindow = g.mzWindow; fake.mzMaxCount = 1000;
		std::unordered_map<uint64_t, uint32_t> seen;
		for (const auto& rd : reads)
			gcseed::iterateKmers(rd, g.mzLength, g.mzLength /* window == k: every position is emitted */, [&](size_t, size_t kmer) { if (!seen.count(kmer)) { seen[kmer] = (uint32_t)fake.mzKmers.size(); fake.mzKmers.push_back(kmer); } });
		for (size_t i = 0; i <= fake.mzKmers.size(); i++) fake.mzKmerStart.push_back((uint32_t)i);
		if (gcUploadMinimizerIndex(ctx, fake) != GCGPU_OK) { fprintf(stderr, "gcgpu_set_minimizer_index: %s\n", gcgpu_last_error()); return 1; }
		if (gcgpu_seed(ctx, codes.data(), codes.size(), sr.data(), (uint32_t)sr.size(), off.data(), nullptr, 0, &used) != GCGPU_OK) { fprintf(stderr, "gcgpu_seed: %s\n", gcgpu_last_error()); return 1; }
		m.resize(used + 1);
		if (gcgpu_fetch_seed_matches(ctx, m.data(), 0, used) != GCGPU_OK) { fprintf(stderr, "fetch: %s\n", gcgpu_last_error()); return 1; }
		SeedRefIndex fref(fake);
		for (size_t r = 0; r < reads.size(); r++)
		{
			auto want = seedRefMatches(fake, fref, reads[r]);
			emitted += want.size();
			bool same = want.size() == off[r + 1] - off[r];
			for (size_t i = 0; same && i < want.size(); i++) { const auto& x = m[off[r] + i]; same = x.pos == std::get<0>(want[i]) && x.start == std::get<2>(want[i]) && x.count == std::get<3>(want[i]); }
			if (!same) { mismatches++; if (mismatches <= 5) fprintf(stderr, "read %zu: emission differs (%llu vs %zu k-mers)\n", r, (unsigned long long)(off[r + 1] - off[r]), want.size()); }
		}
	}
	printf("reads %zu matches %llu seeds %zu emitted %zu mismatches %zu\n", reads.size(), (unsigned long long)realMatches, seeds, emitted, mismatches);
	gcgpu_destroy(ctx);
	return mismatches ? 1 : 0;
}
