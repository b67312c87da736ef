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
#include "../../graphchainer_b200/csrc/gc_k3.cuh"
#include "../../graphchainer_b200/csrc/gc_k2.cuh"
#include <numeric>
#include <algorithm>

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


static uint8_t k3code(char c) { switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; } }

// returns 0 if distance (and ops when wantOps != nullptr) match
static int runK3(const std::string& query, const std::string& target, int expectDist, const std::string* wantOps, uint64_t& work)
{
	std::vector<uint8_t> q(query.size()), t(target.size());
	for (size_t i = 0; i < q.size(); i++) q[i] = k3code(query[i]);
	for (size_t i = 0; i < t.size(); i++) t[i] = k3code(target[i]);
	int32_t Q = (int32_t)q.size(), T = (int32_t)t.size();
	int32_t nb = (Q + 63) / 64; if (nb < 1) nb = 1;
	std::vector<uint64_t> peq(4 * (size_t)nb), rpeq(4 * (size_t)nb);
	gc_k3_build_peq(q.data(), Q, peq.data(), nb);
	std::vector<uint8_t> rq(q.rbegin(), q.rend());
	gc_k3_build_peq(rq.data(), Q, rpeq.data(), nb);
	std::vector<GcK3Block> ba(nb + 1), bb(nb + 1);
	int32_t d = gc_k3_distance(peq.data(), nb, Q, t.data(), T, ba.data(), 64, work);
	if (d != expectDist) { std::cerr << "K3 distance mismatch: got " << d << " want " << expectDist << " (q=" << Q << " t=" << T << ")" << std::endl; return 1; }
	if (!wantOps) return 0;
	if (Q == 0 || T == 0) return wantOps->empty() ? 0 : 1;
	GcK3PathWorkspace w;
	w.peq = peq.data(); w.rpeq = rpeq.data(); w.nbTotal = nb; w.qTotal = Q; w.tTotal = T;
	w.blocksA = ba.data(); w.blocksB = bb.data();
	std::vector<GcK3Block> store(60000); std::vector<uint32_t> colStart((size_t)T + 1); std::vector<GcK3Frame> stack(128);
	w.store = store.data(); w.storeCap = (uint32_t)store.size(); w.colStart = colStart.data(); w.colCap = (uint32_t)colStart.size(); w.stack = stack.data(); w.stackCap = (uint32_t)stack.size();
	std::vector<uint8_t> ops((size_t)Q + T + 8);
	uint32_t nOps = 0;
	if (!gc_k3_path(w, t.data(), d, ops.data(), nOps, (uint32_t)ops.size(), work)) { std::cerr << "K3 path failed internally (q=" << Q << " t=" << T << ")" << std::endl; return 1; }
	std::string got(nOps, '0');
	for (uint32_t i = 0; i < nOps; i++) got[i] = (char)('0' + ops[i]);
	if (got != *wantOps)
	{
		size_t i = 0; while (i < got.size() && i < wantOps->size() && got[i] == (*wantOps)[i]) i++;
		std::cerr << "K3 path mismatch at op " << i << " (len got " << got.size() << " want " << wantOps->size() << ", q=" << Q << " t=" << T << " d=" << d << ")" << std::endl;
		return 1;
	}
	return 0;
}

static int k3Main(const char* stagesPath)
{
	std::ifstream in(stagesPath);
	std::string line, readSeq, pathseq;
	size_t total = 0, bad = 0; uint64_t work = 0;
	while (std::getline(in, line))
	{
		if (line.compare(0, 5, "READ ") == 0) { std::istringstream ss(line); std::string tag, name; ss >> tag >> name >> readSeq; }
		else if (line.compare(0, 11, "GA_PATHSEQ ") == 0)
		{
			std::istringstream ss(line); std::string tag, ps; int d; ss >> tag >> ps >> d;
			total++; bad += runK3(ps, readSeq, d, nullptr, work);
		}
		else if (line.compare(0, 8, "PATHSEQ ") == 0) { pathseq = line.substr(8); if (pathseq == "-") pathseq = ""; }
		else if (line.compare(0, 6, "EDLIB ") == 0)
		{
			std::istringstream ss(line); std::string tag, first; ss >> tag >> first;
			if (first == "ERR") continue;
			int d = atoi(first.c_str()), len, st, en; std::string ops;
			ss >> len >> st >> en >> ops;
			total++; bad += runK3(pathseq, readSeq, d, &ops, work);
		}
	}
	std::cout << "{\"mode\":\"k3\",\"items\":" << total << ",\"mismatches\":" << bad << ",\"columns\":" << work << "}" << std::endl;
	return bad == 0 ? 0 : 1;
}

static GcMpcView mpcView(const GcHostGraph& hg)
{
	GcMpcView m;
	m.compMap = hg.compMap.data(); m.compIdx = hg.compIdx.data(); m.compStart = hg.compStart.data(); m.topoIds = hg.topoIds.data();
	m.pathsStart = hg.pathsStart.data(); m.pathsK = hg.pathsK.data(); m.backStart = hg.backStart.data(); m.backNode = hg.backNode.data(); m.backK = hg.backK.data();
	return m;
}

static int k2Main(const GcHostGraph& hg, const char* stagesPath)
{
	GcMpcView m = mpcView(hg);
	std::ifstream in(stagesPath);
	std::string line;
	std::vector<GcAnchor> anchors;
	size_t total = 0, bad = 0, nAnch = 0;
	while (std::getline(in, line))
	{
		if (line.compare(0, 8, "ANCHORS ") == 0) anchors.clear();
		else if (line.compare(0, 3, "AN ") == 0)
		{
			std::istringstream ss(line); std::string tag; long x, y, fn, fo, ln, lo; size_t np;
			ss >> tag >> x >> y >> fn >> fo >> ln >> lo >> np;
			std::vector<uint32_t> path(np); for (auto& v : path) ss >> v;
			GcAnchor a; a.startNode = path[0]; a.endNode = path.back(); a.x = (int32_t)x; a.y = (int32_t)y;
			anchors.push_back(a);
		}
		else if (line.compare(0, 6, "CHAIN ") == 0)
		{
			std::istringstream ss(line); std::string tag; size_t n; ss >> tag >> n;
			std::vector<uint32_t> want(n); for (auto& v : want) ss >> v;
			uint32_t na = (uint32_t)anchors.size();
			std::vector<uint32_t> order(na); std::iota(order.begin(), order.end(), 0);
			std::stable_sort(order.begin(), order.end(), [&](uint32_t l, uint32_t r) { return anchors[l].y < anchors[r].y; });
			std::vector<int32_t> score(na), pred(na); std::vector<uint32_t> chain(na + 1); int64_t best = 0;
			uint32_t len = gc_k2_chain_seq(m, anchors.data(), na, order.data(), score.data(), pred.data(), chain.data(), &best);
			chain.resize(len);
			total++; nAnch += na;
			if (chain != want) { bad++; if (bad <= 5) std::cerr << "K2 mismatch: chain len " << len << " want " << n << " anchors " << na << std::endl; }
		}
	}
	std::cout << "{\"mode\":\"k2\",\"items\":" << total << ",\"mismatches\":" << bad << ",\"columns\":" << nAnch << "}" << std::endl;
	return bad == 0 ? 0 : 1;
}

int main(int argc, char** argv)
{
	if (argc < 4) { std::cerr << "usage: host_sim k1 index.gcidx stages.txt [maxItems]" << std::endl; return 2; }
	std::string mode = argv[1];
	if (mode == "k3") return k3Main(argv[3]);
	GcIndexFile idx; idx.load(argv[2]);
	GcHostGraph hg; hg.fromIndex(idx);
	if (mode == "k2") return k2Main(hg, argv[3]);
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
