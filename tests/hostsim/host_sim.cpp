// CPU-side check of the device algorithms (test infrastructure for `-m "not gpu"`).
//
// Compiles the GC_HD functions of graphchainer_b200/csrc with g++ and replays the
// stage records written by oracle/_ref/gc_refdump (EXT/RES = one K1 work item and
// the reference's answer).  This validates the kernel LOGIC on the GPU-less build
// box; the GPU parity tests call the same functions through libgcgpu's C ABI.
// Nothing in the shipped library uses this program.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "../../graphchainer_b200/csrc/gc_host_graph.h"

struct Ext { int bigraphNode; size_t offset; std::string seq; bool failed; int score; std::vector<uint64_t> trace; };

static bool parseTrace(std::istringstream& ss, std::vector<uint64_t>& out)
{
	size_t n; ss >> n;
	out.clear();
	for (size_t i = 0; i < n; i++)
	{
		std::string tok; ss >> tok;
		unsigned long node, off; long sp; int sw; char c1, c2;
		if (sscanf(tok.c_str(), "%lu,%lu,%ld,%d,%c%c", &node, &off, &sp, &sw, &c1, &c2) != 6) return false;
		out.push_back(gc_pack_trace((uint32_t)node, (uint32_t)off, (int32_t)sp, sw != 0));
	}
	return true;
}

int main(int argc, char** argv)
{
	if (argc < 4) { std::cerr << "usage: host_sim k1 index.gcidx stages.txt [maxItems]" << std::endl; return 2; }
	std::string mode = argv[1];
	GcIndexFile idx; idx.load(argv[2]);
	GcHostGraph hg; hg.fromIndex(idx);
	GcGraphView g = hg.view();
	GcViterbiTables vt = gcMakeViterbiTables();
	std::ifstream in(argv[3]);
	size_t maxItems = argc > 4 ? strtoull(argv[4], nullptr, 10) : (size_t)-1;
	std::string line;
	size_t total = 0, bad = 0;
	uint64_t columns = 0;
	Ext cur; bool haveExt = false;
	while (std::getline(in, line))
	{
		if (line.compare(0, 4, "EXT ") == 0)
		{
			std::istringstream ss(line);
			std::string tag, stage, dir; size_t fragL, seedIdx;
			ss >> tag >> stage >> fragL >> seedIdx >> dir >> cur.bigraphNode >> cur.offset >> cur.seq;
			if (cur.seq == "-") cur.seq = "";
			haveExt = true;
			continue;
		}
		if (line.compare(0, 4, "RES ") != 0 || !haveExt) continue;
		haveExt = false;
		if (mode != "k1") continue;
		{
			std::istringstream ss(line);
			std::string tag, first; ss >> tag >> first;
			if (first == "F") { cur.failed = true; cur.trace.clear(); }
			else { cur.failed = false; cur.score = atoi(first.c_str()); if (!parseTrace(ss, cur.trace)) { std::cerr << "bad trace line" << std::endl; return 2; } }
		}
		if (total >= maxItems) break;
		total++;
		// ---- run the work item
		uint32_t node = hg.unitigNode(cur.bigraphNode, cur.offset);
		uint32_t off = (uint32_t)(cur.offset - hg.nodeOffset[node]);
		std::vector<uint8_t> seq(cur.seq.size());
		for (size_t i = 0; i < seq.size(); i++) seq[i] = gcEncodeBase(cur.seq[i]);
		int32_t seqLen = (int32_t)seq.size();
		int32_t numSlices = (seqLen + 63) / 64;
		uint32_t itemCap = 64 + 16 * numSlices, heapCap = 256;
		GcK1Result res;
		std::vector<uint64_t> trace;
		for (int attempt = 0; attempt < 6; attempt++)
		{
			std::vector<GcSliceMeta> slices(numSlices + 2);
			std::vector<GcNodeItem> items(itemCap);
			std::vector<uint64_t> heap(heapCap);
			trace.assign(2 * (size_t)seqLen + 256, 0);
			GcK1Workspace ws { slices.data(), items.data(), heap.data(), itemCap, heapCap };
			GcK1Params prm { 10 };
			gc_k1_extend(g, vt, prm, seq.data(), seqLen, node, off, ws, trace.data(), (uint32_t)trace.size(), res);
			if (res.status == GC_OVERFLOW_ITEMS) { itemCap *= 4; continue; }
			if (res.status == GC_OVERFLOW_HEAP) { heapCap *= 4; continue; }
			break;
		}
		columns += res.columns;
		bool ok = true;
		if (cur.failed) ok = res.status == GC_FAILED;
		else
		{
			ok = res.status == GC_OK && res.score == cur.score && res.traceLen == cur.trace.size();
			for (size_t i = 0; ok && i < cur.trace.size(); i++) ok = trace[i] == cur.trace[i];
		}
		if (!ok)
		{
			bad++;
			if (bad <= 5)
			{
				std::cerr << "MISMATCH item " << total << " node " << cur.bigraphNode << " off " << cur.offset << " len " << seqLen << " : ref " << (cur.failed ? "F" : std::to_string(cur.score)) << "/" << cur.trace.size()
					<< " got status " << res.status << " score " << res.score << " len " << res.traceLen << std::endl;
				size_t n = std::min((size_t)res.traceLen, cur.trace.size());
				for (size_t i = 0; i < n; i++) if (trace[i] != cur.trace[i])
				{
					auto pr = [](uint64_t t) { char b[96]; snprintf(b, sizeof(b), "(%u,%u,%d,%d)", (uint32_t)t, (uint32_t)((t >> 32) & 63), (int)((t >> 39) & 0x1FFFFFF) - 1, (int)((t >> 38) & 1)); return std::string(b); };
					std::cerr << "  first diff at " << i << " ref " << pr(cur.trace[i]) << " got " << pr(trace[i]) << (i ? " prev " + pr(cur.trace[i-1]) : "") << std::endl;
					break;
				}
			}
		}
	}
	std::cout << "{\"mode\":\"" << mode << "\",\"items\":" << total << ",\"mismatches\":" << bad << ",\"columns\":" << columns << "}" << std::endl;
	return bad == 0 ? 0 : 1;
}
