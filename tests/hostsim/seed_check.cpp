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
	if (argc < 3) { fprintf(stderr, "usage: seed_check index.gcidx reads.fa\n"); return 2; }
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
		auto a = gcseed::seedsFromMatches(g, got, reads[r].size(), 10), b = gcseed::seedsFromMatches(g, want, reads[r].size(), 10);
		seeds += a.size();
		if (a.size() != b.size()) { mismatches++; continue; }
		for (size_t i = 0; i < a.size(); i++) if (a[i].nodeID != b[i].nodeID || a[i].nodeOffset != b[i].nodeOffset || a[i].seqPos != b[i].seqPos || a[i].reverse != b[i].reverse || a[i].rawSeedGoodness != b[i].rawSeedGoodness) { mismatches++; break; }
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
		if (gcgpu_seed(ctx, nullptr, codes.size(), sr.data(), (uint32_t)sr.size(), off.data(), nullptr, 0, &used) != GCGPU_OK) { fprintf(stderr, "gcgpu_seed: %s\n", gcgpu_last_error()); return 1; }
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
