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
#include "../../graphchainer_b200/csrc/gc_k1s.cuh"
#include "../../graphchainer_b200/csrc/gc_k3.cuh"
#include "../../graphchainer_b200/csrc/gc_k3w.cuh"
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


// ---- warp form of the K3 pass (gc_k3w.cuh) with 32 emulated lanes: the shuffle of the device
// kernel (lane L reads what lane L-1 returned from the previous step) becomes an array copy.
template <int NB>
static uint64_t k3wPassSim(const GcK3wPass& p, GcK3Block* blocksOut)
{
	GcK3wLane<NB> lanes[32];
	GcK3wSegment segs[32];
	uint32_t send[32], recv[32];
	for (int l = 0; l < 32; l++) { gc_k3w_lane_init(p, lanes[l], l); send[l] = 0; }
	int32_t tau = 0;
	while (tau <= p.tauEnd)
	{
		int32_t ev = GC_K3W_NO_EVENT;
		for (int l = 0; l < 32; l++) ev = std::min(ev, gc_k3w_next_event(p, lanes[l], tau));
		if (ev > tau)
		{
			int32_t end = ev <= p.tauEnd ? ev : p.tauEnd + 1;
			for (int l = 0; l < 32; l++) segs[l] = gc_k3w_segment(p, lanes[l], tau);
			for (; tau < end; tau++)
			{
				for (int l = 0; l < 32; l++) recv[l] = send[(l + 31) & 31];
				for (int l = 0; l < 32; l++) send[l] = p.store ? gc_k3w_lane_fast_step<NB, true>(p, lanes[l], segs[l], tau, recv[l]) : gc_k3w_lane_fast_step<NB, false>(p, lanes[l], segs[l], tau, recv[l]);
			}
		}
		else
		{
			for (int l = 0; l < 32; l++) recv[l] = send[(l + 31) & 31];
			for (int l = 0; l < 32; l++) send[l] = gc_k3w_lane_step(p, lanes[l], tau, recv[l], blocksOut);
			tau++;
		}
	}
	uint64_t work = 0;
	for (int l = 0; l < 32; l++) work += lanes[l].work;
	return work;
}
static uint64_t k3wPassSimNB(const GcK3wPass& p, int NB, GcK3Block* blocksOut)
{
	switch (NB)
	{
		case 1: return k3wPassSim<1>(p, blocksOut);
		case 2: return k3wPassSim<2>(p, blocksOut);
		case 3: return k3wPassSim<3>(p, blocksOut);
		case 4: return k3wPassSim<4>(p, blocksOut);
		default: return k3wPassSim<8>(p, blocksOut);
	}
}

struct K3wHostExec
{
	bool leader() const { return true; }
	void sync() const {}
	uint32_t fromLeader(uint32_t v) const { return v; }
	uint64_t pass(const GcK3wPass& p, int NB, GcK3Block* out) { return k3wPassSimNB(p, NB, out); }
	int32_t firstSplitRow(const GcK3Block* A, int32_t lfb, int32_t llb, const GcK3Block* B, int32_t rfb, int32_t rlb, int32_t q, int32_t best)
	{
		const int32_t INF = 1 << 29;
		for (int32_t r = 0; r <= q - 2; r++)
		{
			int32_t b = r >> 6; int32_t ls = (b < lfb || b > llb) ? INF : gc_k3_cell(A[b], r);
			int32_t rr = q - 1 - (r + 1); int32_t rb = rr >> 6; int32_t rs = (rb < rfb || rb > rlb) ? INF : gc_k3_cell(B[rb], rr);
			if (ls + rs == best) return r;
		}
		return -1;
	}
};
// the edit path through the warp form; returns "" on internal failure
static std::string k3wPathSim(const std::vector<uint8_t>& q, const std::vector<uint8_t>& t, int32_t best, uint64_t& work)
{
	int32_t Q = (int32_t)q.size(), T = (int32_t)t.size();
	int32_t nb = (Q + 63) / 64; if (nb < 1) nb = 1;
	std::vector<uint64_t> peq(4 * (size_t)nb), rpeq(4 * (size_t)nb);
	gc_k3_build_peq(q.data(), Q, peq.data(), nb);
	std::vector<uint8_t> rq(q.rbegin(), q.rend());
	gc_k3_build_peq(rq.data(), Q, rpeq.data(), nb);
	std::vector<GcK3Block> ba(nb + 1), bb(nb + 1), store(60000);
	std::vector<GcK3Frame> stack(128);
	GcK3wPathWorkspace w;
	w.peq = peq.data(); w.rpeq = rpeq.data(); w.nbTotal = nb; w.qTotal = Q; w.tTotal = T; w.blocksA = ba.data(); w.blocksB = bb.data();
	w.store = store.data(); w.storeCap = (uint32_t)store.size(); w.stack = stack.data(); w.stackCap = (uint32_t)stack.size(); w.maxNB = 8;
	std::vector<uint8_t> ops((size_t)Q + T + 8);
	uint32_t nOps = 0;
	K3wHostExec ex;
	if (!gc_k3w_path(ex, w, t.data(), best, ops.data(), nOps, (uint32_t)ops.size(), work)) return "!";
	std::string got(nOps, '0');
	for (uint32_t i = 0; i < nOps; i++) got[i] = (char)('0' + ops[i]);
	return got;
}
static int roundNB(int nb) { return nb <= 4 ? nb : 8; }
// edit distance through the warp form; extraNB > 0 forces a larger group size than necessary
static int32_t k3wDistanceSim(const uint64_t* peq, int32_t nb, int32_t Q, const uint8_t* t, int32_t T, int32_t kStart, int extraNB, uint64_t& work)
{
	if (Q == 0 || T == 0) return Q > T ? Q : T;
	std::vector<GcK3Block> blocks(nb + 1);
	int32_t k = kStart < 64 ? 64 : kStart;
	int32_t diff = Q > T ? Q - T : T - Q, mx = Q > T ? Q : T;
	while (true)
	{
		if (k >= diff)
		{
			int32_t kk = k > mx ? mx : k;
			int NB = roundNB(gc_k3w_blocks_per_lane(Q, T, kk) + extraNB);
			GcK3wPass p = gc_k3w_make_pass(peq, nb, 0, Q, t, 0, 1, T, kk, T - 1, NB);
			work += k3wPassSimNB(p, NB, blocks.data());
			int32_t v = gc_k3_cell(blocks[(Q - 1) >> 6], Q - 1);
			if (v <= kk) return v;
		}
		k *= 2;
	}
}

// random pairs (substitutions + indels) checked against the thread form, plus sub-query /
// reversed / stop-column passes as the Hirschberg recursion issues them
static int k3wMain(const char* stagesPath)
{
	size_t total = 0, bad = 0; uint64_t work = 0, workRef = 0;
	auto check = [&](const std::vector<uint8_t>& q, const std::vector<uint8_t>& t, int extraNB, int kStart)
	{
		int32_t Q = (int32_t)q.size(), T = (int32_t)t.size();
		int32_t nb = (Q + 63) / 64; if (nb < 1) nb = 1;
		std::vector<uint64_t> peq(4 * (size_t)nb);
		gc_k3_build_peq(q.data(), Q, peq.data(), nb);
		std::vector<GcK3Block> ba(nb + 1);
		int32_t want = gc_k3_distance(peq.data(), nb, Q, t.data(), T, ba.data(), 64, workRef);
		int32_t got = k3wDistanceSim(peq.data(), nb, Q, t.data(), T, kStart, extraNB, work);
		total++;
		if (got != want) { bad++; if (bad <= 5) std::cerr << "K3W mismatch q=" << Q << " t=" << T << " extraNB=" << extraNB << " kStart=" << kStart << ": got " << got << " want " << want << std::endl; }
		if (Q > 0 && T > 0 && extraNB == 0 && kStart == 0)
		{
			// edit path: warp form vs thread form (the latter is pinned against edlib's op strings by mode k3)
			std::vector<uint8_t> rq(q.rbegin(), q.rend());
			std::vector<uint64_t> rpeq(4 * (size_t)nb);
			gc_k3_build_peq(rq.data(), Q, rpeq.data(), nb);
			std::vector<GcK3Block> bb(nb + 1), store(60000); std::vector<uint32_t> colStart((size_t)T + 1); std::vector<GcK3Frame> stack(128);
			GcK3PathWorkspace w;
			w.peq = peq.data(); w.rpeq = rpeq.data(); w.nbTotal = nb; w.qTotal = Q; w.tTotal = T; w.blocksA = ba.data(); w.blocksB = bb.data();
			w.store = store.data(); w.storeCap = (uint32_t)store.size(); w.colStart = colStart.data(); w.colCap = (uint32_t)colStart.size(); w.stack = stack.data(); w.stackCap = (uint32_t)stack.size();
			std::vector<uint8_t> ops((size_t)Q + T + 8); uint32_t nOps = 0; uint64_t wk = 0;
			bool okRef = gc_k3_path(w, t.data(), want, ops.data(), nOps, (uint32_t)ops.size(), wk);
			std::string ref(nOps, '0'); for (uint32_t i = 0; i < nOps; i++) ref[i] = (char)('0' + ops[i]);
			std::string gotPath = k3wPathSim(q, t, want, work);
			total++;
			if (!okRef || gotPath != ref) { bad++; if (bad <= 5) std::cerr << "K3W path mismatch q=" << Q << " t=" << T << " d=" << want << " (ref ok " << okRef << ", got len " << gotPath.size() << " ref len " << ref.size() << ")" << std::endl; }
		}
		// a Hirschberg-style half pass on a sub-query with a non-aligned offset, forward and reversed
		if (Q > 200 && T > 200 && want > 0)
		{
			int32_t qOff = 37 + (Q / 7), q2 = Q - qOff - 11, tOff = T / 9, t2 = T - tOff - 5;
			for (int rev = 0; rev < 2; rev++)
			{
				std::vector<uint8_t> rq(q.rbegin(), q.rend());
				std::vector<uint64_t> rpeq(4 * (size_t)nb);
				gc_k3_build_peq(rq.data(), Q, rpeq.data(), nb);
				const uint64_t* pq = rev ? rpeq.data() : peq.data();
				int32_t qo = rev ? Q - qOff - q2 : qOff;
				int64_t tBase = rev ? (int64_t)tOff + t2 - 1 : tOff; int32_t tStep = rev ? -1 : 1;
				int32_t k2 = want + 40; int32_t mx2 = q2 > t2 ? q2 : t2; if (k2 > mx2) k2 = mx2;
				int32_t d2 = q2 > t2 ? q2 - t2 : t2 - q2; if (k2 < d2) k2 = d2;
				int32_t stop = t2 / 2;
				std::vector<GcK3Block> refBlocks(nb + 1), gotBlocks(nb + 1);
				gc_k3_pass(pq, nb, qo, q2, t.data(), tBase, tStep, t2, k2, stop, refBlocks.data(), nullptr, nullptr);
				int NB = roundNB(gc_k3w_blocks_per_lane(q2, t2, k2) + extraNB);
				GcK3wPass p = gc_k3w_make_pass(pq, nb, qo, q2, t.data(), tBase, tStep, t2, k2, stop, NB);
				k3wPassSimNB(p, NB, gotBlocks.data());
				GcK3Band band = gc_k3_band(q2, t2, k2);
				int32_t fb = gc_k3_first_block(band, stop), lb = gc_k3_last_block(band, q2, stop);
				int32_t f2, l2; gc_k3w_stop_blocks(p, NB, f2, l2);
				total++;
				bool ok = f2 <= fb && l2 >= lb;
				// inside the strict band the warp form may only be tighter (its band is a superset), and every
				// value that is part of a <= k2 path must agree: compare cells whose reference value is <= k2
				for (int32_t r = fb * 64; ok && r <= std::min(q2 - 1, lb * 64 + 63); r++)
				{
					int32_t a = gc_k3_cell(refBlocks[r >> 6], r), b = gc_k3_cell(gotBlocks[r >> 6], r);
					if (b > a) ok = false;
				}
				if (!ok) { bad++; if (bad <= 5) std::cerr << "K3W half-pass mismatch q2=" << q2 << " t2=" << t2 << " rev=" << rev << " NB=" << NB << std::endl; }
			}
		}
	};
	uint64_t rng = 88172645463325252ULL;
	auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
	const int lens[] = { 1, 5, 63, 64, 65, 130, 700, 2049, 4100, 9000, 12500 };
	const double errs[] = { 0.0, 0.02, 0.15, 0.35 };
	for (int len : lens) for (double e : errs) for (int extra = 0; extra < 2; extra++)
	{
		std::vector<uint8_t> q(len), t;
		for (auto& c : q) c = (uint8_t)(next() & 3);
		for (int i = 0; i < len; i++)
		{
			double u = (double)(next() % 100000) / 100000.0;
			if (u < e / 3) continue;                                   // deletion
			if (u < 2 * e / 3) { t.push_back((uint8_t)(next() & 3)); continue; } // substitution
			t.push_back(q[i]);
			if (u < e) for (int z = (int)(next() % 6); z > 0; z--) t.push_back((uint8_t)(next() & 3)); // insertion
		}
		if (len == 700 && e == 0.02) t.push_back(4); // a non-ACGT target symbol matches nothing
		check(q, t, extra, 0);
		if (extra == 0) check(q, t, 0, 3000); // a generous upper bound as first cutoff: one pass
	}
	// the golden NW items as well
	std::ifstream in(stagesPath);
	std::string line, readSeq;
	std::vector<uint8_t> lastQ, lastT;
	while (std::getline(in, line))
	{
		if (line.compare(0, 5, "READ ") == 0) { std::istringstream ss(line); std::string tag, name; ss >> tag >> name >> readSeq; }
		else if (line.compare(0, 11, "GA_PATHSEQ ") == 0 || line.compare(0, 8, "PATHSEQ ") == 0)
		{
			std::istringstream ss(line); std::string tag, ps; ss >> tag >> ps;
			if (ps == "-") continue;
			std::vector<uint8_t> q(ps.size()), t(readSeq.size());
			for (size_t i = 0; i < q.size(); i++) q[i] = k3code(ps[i]);
			for (size_t i = 0; i < t.size(); i++) t[i] = k3code(readSeq[i]);
			check(q, t, 0, 0);
			lastQ = q; lastT = t;
		}
		else if (line.compare(0, 6, "EDLIB ") == 0)
		{
			std::istringstream ss(line); std::string tag, first; ss >> tag >> first;
			if (first == "ERR") continue;
			int d = atoi(first.c_str()), len, st, en; std::string ops;
			ss >> len >> st >> en >> ops;
			std::string got = k3wPathSim(lastQ, lastT, d, work);
			total++;
			if (got != ops) { bad++; if (bad <= 5) std::cerr << "K3W golden path mismatch d=" << d << std::endl; }
		}
	}
	std::cout << "{\"mode\":\"k3w\",\"items\":" << total << ",\"mismatches\":" << bad << ",\"columns\":" << work << ",\"columns_thread_form\":" << workRef << "}" << std::endl;
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
	if (mode == "k3w") return k3wMain(argv[3]);
	GcIndexFile idx; idx.load(argv[2]);
	GcHostGraph hg; hg.fromIndex(idx);
	if (mode == "k2") return k2Main(hg, argv[3]);
	GcGraphView g = hg.view();
	std::vector<GcNodeRec> nodeRecs; std::vector<uint64_t> outKeys;
	gcBuildNodeRecs(g, nodeRecs, outKeys);
	g.nodeRec = nodeRecs.data(); g.outKey = outKeys.data();
	GcViterbiTables vt = gcMakeViterbiTables();
	std::ifstream in(argv[3]);
	size_t maxItems = argc > 4 ? strtoull(argv[4], nullptr, 10) : (size_t)-1;
	std::string line;
	size_t total = 0, bad = 0;
	uint64_t columns = 0;
	Ext cur; bool haveExt = false;
	std::vector<Ext> queued; // k1q: every item of the file goes through ONE lane, one after the other (the way a lane of the GPU kernels takes items)
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
		if (mode != "k1" && mode != "k1s" && mode != "k1q") continue;
		{
			std::istringstream ss(line);
			std::string tag, first; ss >> tag >> first;
			if (first == "F") { cur.failed = true; cur.trace.clear(); }
			else { cur.failed = false; cur.score = atoi(first.c_str()); if (!parseTrace(ss, cur.trace)) { std::cerr << "bad trace line" << std::endl; return 2; } }
		}
		if (total >= maxItems) break;
		total++;
		if (mode == "k1q") { queued.push_back(cur); continue; }
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
			GcColVV colsBuf[64];
			GcK1Workspace ws { slices.data(), items.data(), heap.data(), colsBuf, itemCap, heapCap };
			GcK1Params prm { 10 };
			if (mode == "k1s")
			{
				// lane-per-item form (gc_k1s.cuh) as a warp of one lane; bit planes built as gc_planes_kernel does
				std::vector<uint32_t> keys(itemCap), scratch(2 * (size_t)heapCap);
				std::vector<GcItemAux> aux(itemCap);
				std::vector<uint64_t> planes(4 * ((size_t)seqLen / 64 + 3), 0);
				for (int32_t i = 0; i < seqLen; i++) for (int b = 0; b < 4; b++) if ((seq[i] >> b) & 1) planes[4 * (size_t)(i >> 6) + b] |= 1ULL << (i & 63);
				GcK1SWorkspace sw; sw.slices = slices.data(); sw.items = items.data(); sw.keys = keys.data(); sw.aux = aux.data(); sw.scratch = scratch.data(); sw.scratchCap = (uint32_t)scratch.size(); sw.itemCap = itemCap;
				sw.heap.base = heap.data(); sw.heap.stride = 1; sw.heap.cap = heapCap;
				res.score = GC_INT_MAX; res.traceLen = 0; res.itemsUsed = 0;
				int32_t last = gc_k1s_forward(g, vt, prm, true, seq.data(), seqLen, node, off, (total & 1) ? planes.data() : nullptr, 0, sw, res);
				if (res.status == GC_OK && last < 1) res.status = GC_FAILED;
				if (res.status == GC_OK) gc_k1s_backtrace(g, true, seq.data(), seqLen, (total & 1) ? planes.data() : nullptr, 0, sw, last, colsBuf, trace.data(), (uint32_t)trace.size(), res);
			}
			else
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
	if (mode == "k1q")
	{
		struct Slab
		{
			std::vector<uint8_t> seq; std::vector<uint64_t> planes; std::vector<GcSliceMeta> slices; std::vector<GcNodeItem> items; std::vector<uint32_t> keys, scratch; std::vector<GcItemAux> aux;
			std::vector<uint64_t> trace; GcK1SItem item; GcK1SWorkspace ws; GcK1Result res; int32_t last = 0;
		};
		std::vector<Slab> slabs(queued.size());
		std::vector<uint64_t> heap(4096);
		for (size_t k = 0; k < queued.size(); k++)
		{
			const Ext& e = queued[k]; Slab& sl = slabs[k];
			uint32_t node = hg.unitigNode(e.bigraphNode, e.offset);
			sl.seq.resize(e.seq.size());
			for (size_t i = 0; i < sl.seq.size(); i++) sl.seq[i] = gcEncodeBase(e.seq[i]);
			int32_t seqLen = (int32_t)sl.seq.size(), numSlices = (seqLen + 63) / 64;
			uint32_t itemCap = 1024 + 256 * numSlices;
			sl.planes.assign(4 * ((size_t)seqLen / 64 + 3), 0);
			for (int32_t i = 0; i < seqLen; i++) for (int b = 0; b < 4; b++) if ((sl.seq[i] >> b) & 1) sl.planes[4 * (size_t)(i >> 6) + b] |= 1ULL << (i & 63);
			sl.slices.resize(numSlices + 2); sl.items.resize(itemCap); sl.keys.resize(itemCap); sl.aux.resize(itemCap); sl.scratch.resize(2 * heap.size()); sl.trace.assign(2 * (size_t)seqLen + 256, 0);
			sl.item.seq = sl.seq.data(); sl.item.seqLen = seqLen; sl.item.startNode = node; sl.item.startOffset = (uint32_t)(e.offset - hg.nodeOffset[node]); sl.item.planeBit = 0;
			sl.ws.slices = sl.slices.data(); sl.ws.items = sl.items.data(); sl.ws.keys = sl.keys.data(); sl.ws.aux = sl.aux.data(); sl.ws.scratch = sl.scratch.data(); sl.ws.scratchCap = (uint32_t)sl.scratch.size(); sl.ws.itemCap = itemCap;
			sl.ws.heap.base = heap.data(); sl.ws.heap.stride = 1; sl.ws.heap.cap = (uint32_t)heap.size();
			sl.res.status = GC_OK; sl.res.score = GC_INT_MAX; sl.res.traceLen = 0; sl.res.itemsUsed = 0; sl.res.columns = 0;
		}
		struct Fwd
		{
			std::vector<Slab>* slabs; size_t at = 0, cur = 0;
			bool next(GcK1SItem& it, GcK1SWorkspace& ws) { if (at >= slabs->size()) return false; cur = at++; it = (*slabs)[cur].item; ws = (*slabs)[cur].ws; return true; }
			void done(GcK1Result r, int32_t last) { if (r.status == GC_OK && last < 1) r.status = GC_FAILED; Slab& sl = (*slabs)[cur]; sl.res.status = r.status; sl.res.columns = r.columns; sl.res.itemsUsed = r.itemsUsed; sl.last = last; }
		} fwd; fwd.slabs = &slabs;
		struct Bwd
		{
			std::vector<Slab>* slabs; size_t at = 0, cur = 0;
			bool next(GcK1SItem& it, GcK1SWorkspace& ws, int32_t& last, uint64_t*& out, uint32_t& cap, GcK1Result& r)
			{
				while (at < slabs->size() && (*slabs)[at].res.status != GC_OK) at++;
				if (at >= slabs->size()) return false;
				cur = at++; Slab& sl = (*slabs)[cur]; it = sl.item; ws = sl.ws; last = sl.last; out = sl.trace.data(); cap = (uint32_t)sl.trace.size(); r = sl.res; return true;
			}
			void done(const GcK1Result& r) { (*slabs)[cur].res = r; }
		} bwd; bwd.slabs = &slabs;
		GcK1Params prm { 10 };
		GcK1SWorkspace lane; lane.slices = nullptr; lane.items = nullptr; lane.keys = nullptr; lane.aux = nullptr; lane.scratch = nullptr; lane.scratchCap = 0; lane.itemCap = 0; lane.heap.base = heap.data(); lane.heap.stride = 1; lane.heap.cap = (uint32_t)heap.size();
		GcColVV colsBuf[64];
		gc_k1s_forward_items(g, vt, prm, (const uint64_t*)nullptr, lane, fwd);
		gc_k1s_backtrace_items(g, (const uint64_t*)nullptr, lane, colsBuf, bwd);
		for (size_t k = 0; k < queued.size(); k++)
		{
			const Ext& e = queued[k]; const Slab& sl = slabs[k];
			columns += sl.res.columns;
			bool ok = true;
			if (e.failed) ok = sl.res.status == GC_FAILED;
			else
			{
				ok = sl.res.status == GC_OK && sl.res.score == e.score && sl.res.traceLen == e.trace.size();
				for (size_t i = 0; ok && i < e.trace.size(); i++) ok = sl.trace[i] == e.trace[i];
			}
			if (!ok) { bad++; if (bad <= 5) std::cerr << "MISMATCH queued item " << k << " status " << sl.res.status << " score " << sl.res.score << " len " << sl.res.traceLen << std::endl; }
		}
	}
	std::cout << "{\"mode\":\"" << mode << "\",\"items\":" << total << ",\"mismatches\":" << bad << ",\"columns\":" << columns << "}" << std::endl;
	return bad == 0 ? 0 : 1;
}
