// K3, warp form -- one banded NW Myers pass computed by the 32 lanes of a warp as a
// skewed wavefront (the per-lane logic is GC_HD so that tests/hostsim can replay it with
// 32 emulated lanes; the shuffle that links the lanes lives in gcgpu.cu).
//
// Mapping.  The query's 64-row blocks (edlib's Block, edlib.cpp:409-444) are tied in GROUPS of
// NB consecutive blocks; group g belongs to lane g % 32 and is advanced through column c at
// wavefront step tau = c + g.  The horizontal delta leaving the last block of group g-1 at
// (c, tau-1) is exactly what group g needs at (c, tau), so one shuffle per step carries the
// whole inter-lane dependency.  A group is only computed for the columns in which the
// Ukkonen band of the cutoff k touches it ( dlo = -(k-(q-t))/2 .. dhi = (k+(q-t))/2, the band
// of edlib.cpp:755 ); when it leaves the band its lane picks up group g+32 which enters the
// band further down.  NB is chosen so that this hand-over never collides:
//     (dhi - dlo) <= 1984 * NB + 32        (time of (g+32, first column) > time of (g, last column)).
// All block state (P, M, score, the four Eq words) stays in registers.
//
// Exactness: cells outside the band are replaced by upper bounds (a fresh block starts from
// "all +1" below the block above, a block whose upper neighbour left the band takes hin=+1),
// cells on any path of cost <= k keep their true values, so a result <= k is the true edit
// distance -- the same argument as gc_k3.cuh, independent of band bookkeeping.
#pragma once
#include "gc_k3.cuh"

#define GC_K3W_LANES 32
#if defined(__CUDACC__)
#define GC_UNROLL _Pragma("unroll")
#else
#define GC_UNROLL
#endif

struct GcK3wPass
{
	const uint64_t* peq;   // profile of the whole query, peq[c * nbTotal + b]
	int32_t nbTotal;
	int32_t qOff, q;       // sub-query
	const uint8_t* target; // symbols read as target[tBase + c * tStep]
	int64_t tBase;
	int32_t tStep, t;
	int32_t stopCol;       // last column to compute
	int32_t dlo, dhi;      // band of diagonals (row - column)
	int32_t nb;            // blocks of the sub-query
	int32_t numGroups;     // ceil(nb / NB)
	int32_t tauEnd;        // last wavefront step
	GcK3Block* store;      // leaf traceback: block b of column c -> store[c * storeStride + b] (null: not kept)
	int32_t storeStride;
	int32_t lanes;         // lanes of the wavefront: 32 (one warp) or 32 x the warps of a block (gc_k3b_*, hand-over through shared memory)
};

// blocks per lane so that a lane is done with group g before group g + lanes enters the band:
// (dhi - dlo) <= 64 * NB * (lanes - 1) + lanes
GC_HD int32_t gc_k3w_blocks_per_lane(int32_t q, int32_t t, int32_t k, int32_t lanes = GC_K3W_LANES)
{
	GcK3Band band = gc_k3_band(q, t, k);
	int32_t width = band.dhi - band.dlo;
	if (width <= lanes) return 1;
	int32_t per = 64 * (lanes - 1);
	return (width - lanes + per - 1) / per;
}

GC_HD GcK3wPass gc_k3w_make_pass(const uint64_t* peq, int32_t nbTotal, int32_t qOff, int32_t q, const uint8_t* target, int64_t tBase, int32_t tStep, int32_t t, int32_t k, int32_t stopCol, int32_t NB, int32_t lanes = GC_K3W_LANES)
{
	GcK3wPass p;
	p.lanes = lanes;
	p.peq = peq; p.nbTotal = nbTotal; p.qOff = qOff; p.q = q; p.target = target; p.tBase = tBase; p.tStep = tStep; p.t = t; p.stopCol = stopCol;
	GcK3Band band = gc_k3_band(q, t, k);
	p.dlo = band.dlo; p.dhi = band.dhi;
	p.nb = (q + 63) / 64;
	p.numGroups = (p.nb + NB - 1) / NB;
	int32_t gLast = (stopCol + p.dhi) / (64 * NB);
	if (gLast > p.numGroups - 1) gLast = p.numGroups - 1;
	p.tauEnd = stopCol + gLast;
	p.store = nullptr; p.storeStride = 0;
	return p;
}

template <int NB>
struct GcK3wLane
{
	uint64_t P[NB], M[NB];
	uint64_t eq[NB][4];
	int32_t score[NB];
	int32_t g;          // current group of this lane
	int32_t nbHere;     // existing blocks in the group (NB except for the last group)
	int32_t cFirst, cLast; // columns in which the group is computed; cFirst > cLast = never
	int32_t aboveLast;  // last column in which group g-1 is computed (-1: none)
	uint32_t prevRecv;  // what the lane above sent one step earlier
	uint32_t work;      // block steps done by this lane
};

template <int NB>
GC_HD void gc_k3w_set_group(const GcK3wPass& p, GcK3wLane<NB>& s, int32_t g)
{
	s.g = g;
	if (g >= p.numGroups) { s.cFirst = 0x7FFFFFFF; s.cLast = -1; s.nbHere = 0; s.aboveLast = -1; return; }
	int32_t b0 = g * NB;
	int32_t b1 = b0 + NB - 1; if (b1 > p.nb - 1) b1 = p.nb - 1;
	s.nbHere = b1 - b0 + 1;
	int32_t cf = 64 * b0 - p.dhi; if (cf < 0) cf = 0;
	int32_t cl = 64 * b1 + 63 - p.dlo; if (cl > p.stopCol) cl = p.stopCol;
	s.cFirst = cf; s.cLast = cl;
	s.aboveLast = g > 0 ? 64 * (b0 - 1) + 63 - p.dlo : -1;
	if (cf > cl) return;
	GC_UNROLL
	for (int i = 0; i < NB; i++)
		GC_UNROLL
		for (int c = 0; c < 4; c++)
			s.eq[i][c] = (i < s.nbHere) ? gc_k3_eq(p.peq, p.nbTotal, p.qOff, p.q, c, b0 + i) : 0ULL;
}

template <int NB>
GC_HD void gc_k3w_lane_init(const GcK3wPass& p, GcK3wLane<NB>& s, int32_t lane)
{
	s.prevRecv = 0; s.work = 0;
	GC_UNROLL
	for (int i = 0; i < NB; i++) { s.P[i] = ~0ULL; s.M[i] = 0; s.score[i] = 0; }
	gc_k3w_set_group(p, s, lane);
}

// one wavefront step of one lane.  `recv` = the value the lane above (lane-1, cyclically)
// returned from its previous step.  Returns the value to hand to the lane below:
// (score of the group's last block after this column) << 2 | (hout + 1).
template <int NB>
GC_HD uint32_t gc_k3w_lane_step(const GcK3wPass& p, GcK3wLane<NB>& s, int32_t tau, uint32_t recv, GcK3Block* blocksOut)
{
	uint32_t send = 0;
	int32_t c = tau - s.g;
	if (c >= s.cFirst && c <= s.cLast)
	{
		bool aboveActive = c <= s.aboveLast;
		int hin = aboveActive ? (int)(recv & 3u) - 1 : 1;
		if (c == s.cFirst)
		{
			int32_t base;
			if (c == 0) base = s.g * NB * 64;                                    // D[i][-1] = i + 1
			else if (aboveActive) base = (int32_t)(recv >> 2) - hin;             // score of the block above after column c-1
			else base = (int32_t)(s.prevRecv >> 2);
			GC_UNROLL
			for (int i = 0; i < NB; i++) { s.P[i] = ~0ULL; s.M[i] = 0; s.score[i] = base + 64 * (i + 1); }
		}
		int sym = p.target[p.tBase + (int64_t)c * p.tStep];
		int32_t lastScore = 0;
		GC_UNROLL
		for (int i = 0; i < NB; i++)
		{
			if (i < s.nbHere)
			{
				uint64_t Eq = sym == 0 ? s.eq[i][0] : sym == 1 ? s.eq[i][1] : sym == 2 ? s.eq[i][2] : sym == 3 ? s.eq[i][3] : 0ULL;
				hin = gc_k3_block(s.P[i], s.M[i], Eq, hin);
				s.score[i] += hin;
				lastScore = s.score[i];
			}
		}
		s.work += (uint32_t)s.nbHere;
		send = ((uint32_t)lastScore << 2) | (uint32_t)(hin + 1);
		if (p.store)
		{
			GcK3Block* dst = p.store + (int64_t)c * p.storeStride + s.g * NB;
			GC_UNROLL
			for (int i = 0; i < NB; i++)
				if (i < s.nbHere) { GcK3Block bl; bl.P = s.P[i]; bl.M = s.M[i]; bl.score = s.score[i]; bl.pad = 0; dst[i] = bl; }
		}
		if (c == p.stopCol)
		{
			GC_UNROLL
			for (int i = 0; i < NB; i++)
				if (i < s.nbHere) { GcK3Block bl; bl.P = s.P[i]; bl.M = s.M[i]; bl.score = s.score[i]; bl.pad = 0; blocksOut[s.g * NB + i] = bl; }
		}
		if (c == s.cLast) gc_k3w_set_group(p, s, s.g + p.lanes);
	}
	s.prevRecv = recv;
	return send;
}


// ---- fast path.  A lane's control state only changes at a few "event" steps: the first and the
// last column of its group (block initialisation, hand-over to group g+32, stop-column write) and
// the step after the group above has left the band.  Between events every lane is either
// computing with constant settings or idle, so the warp runs whole segments of steps through
// gc_k3w_lane_fast_step (no control flow besides one predicate) and executes the general
// gc_k3w_lane_step only at event steps.
#define GC_K3W_NO_EVENT 0x7FFFFFFF
template <int NB>
GC_HD int32_t gc_k3w_next_event(const GcK3wPass& p, const GcK3wLane<NB>& s, int32_t tau)
{
	if (s.cFirst > s.cLast) return GC_K3W_NO_EVENT;
	int32_t ev = GC_K3W_NO_EVENT;
	int32_t e1 = s.cFirst + s.g; if (e1 >= tau && e1 < ev) ev = e1;
	int32_t e2 = s.cLast + s.g; if (e2 >= tau && e2 < ev) ev = e2;
	int32_t e3 = s.aboveLast + 1 + s.g; if (s.aboveLast >= 0 && e3 >= tau && e3 < ev) ev = e3;
	return ev;
}

// per-segment constants of a lane
struct GcK3wSegment
{
	bool active;            // the lane computes in every step of the segment
	bool useRecv;           // the group above is in the band: take its horizontal delta
	const uint8_t* tptr;    // symbol of the lane's column at the first step of the segment
};
template <int NB>
GC_HD GcK3wSegment gc_k3w_segment(const GcK3wPass& p, const GcK3wLane<NB>& s, int32_t tau)
{
	GcK3wSegment seg;
	int32_t c = tau - s.g;
	seg.active = c > s.cFirst && c < s.cLast;
	seg.useRecv = c <= s.aboveLast;
	seg.tptr = p.target + (seg.active ? p.tBase + (int64_t)c * p.tStep : p.tBase);
	return seg;
}

template <int NB, bool STORE>
GC_HD uint32_t gc_k3w_lane_fast_step(const GcK3wPass& p, GcK3wLane<NB>& s, GcK3wSegment& seg, int32_t tau, uint32_t recv)
{
	uint32_t send = 0;
	if (seg.active)
	{
		int hin = seg.useRecv ? (int)(recv & 3u) - 1 : 1;
		int sym = *seg.tptr;
		seg.tptr += p.tStep;
		uint64_t keep = sym < 4 ? ~0ULL : 0ULL;
		int32_t lastScore = 0;
		GC_UNROLL
		for (int i = 0; i < NB; i++)
		{
			if (i < s.nbHere)
			{
				uint64_t lo = (sym & 1) ? s.eq[i][1] : s.eq[i][0];
				uint64_t hi = (sym & 1) ? s.eq[i][3] : s.eq[i][2];
				uint64_t Eq = ((sym & 2) ? hi : lo) & keep;
				hin = gc_k3_block(s.P[i], s.M[i], Eq, hin);
				s.score[i] += hin;
				lastScore = s.score[i];
			}
		}
		s.work += (uint32_t)s.nbHere;
		send = ((uint32_t)lastScore << 2) | (uint32_t)(hin + 1);
		if (STORE)
		{
			int32_t c = tau - s.g;
			GcK3Block* dst = p.store + (int64_t)c * p.storeStride + s.g * NB;
			GC_UNROLL
			for (int i = 0; i < NB; i++)
				if (i < s.nbHere) { GcK3Block bl; bl.P = s.P[i]; bl.M = s.M[i]; bl.score = s.score[i]; bl.pad = 0; dst[i] = bl; }
		}
	}
	s.prevRecv = recv;
	return send;
}

// blocks [first, last] of the stop column that the pass wrote to blocksOut (group-granular superset of the band)
GC_HD void gc_k3w_stop_blocks(const GcK3wPass& p, int32_t NB, int32_t& first, int32_t& last)
{
	int32_t lo = p.stopCol + p.dlo; if (lo < 0) lo = 0;
	int32_t hi = p.stopCol + p.dhi; if (hi > p.q - 1) hi = p.q - 1;
	int32_t gLo = (lo >> 6) / NB, gHi = (hi >> 6) / NB;
	first = gLo * NB;
	last = gHi * NB + NB - 1; if (last > p.nb - 1) last = p.nb - 1;
}

// blocks [first, last] that a pass computes in column c (group-granular superset of the band)
GC_HD void gc_k3w_column_blocks(const GcK3wPass& p, int32_t NB, int32_t c, int32_t& first, int32_t& last)
{
	int32_t lo = c + p.dlo; if (lo < 0) lo = 0;
	int32_t hi = c + p.dhi; if (hi > p.q - 1) hi = p.q - 1;
	int32_t gLo = (lo >> 6) / NB, gHi = (hi >> 6) / NB;
	first = gLo * NB;
	last = gHi * NB + NB - 1; if (last > p.nb - 1) last = p.nb - 1;
}

// ------------------------------------------------------------------ alignment path, warp form
// Same recursion as gc_k3_path (obtainAlignment / obtainAlignmentHirschberg / obtainAlignmentTraceback,
// edlib.cpp:1164-1399, 945-1144) with every banded pass run by the warp.  `Exec` supplies the
// warp primitives: on the device they are shuffles/ballots (gcgpu.cu), in tests/hostsim an
// emulation with 32 lanes.  Control flow is warp-uniform; only the leader lane writes.
struct GcK3wPathWorkspace
{
	const uint64_t* peq;    // [4 * nbTotal]
	const uint64_t* rpeq;   // [4 * nbTotal] profile of the reversed query
	int32_t nbTotal, qTotal, tTotal;
	GcK3Block* blocksA;     // [nbTotal]
	GcK3Block* blocksB;     // [nbTotal]
	GcK3Block* store;       // [storeCap] leaf columns, full layout column-major
	uint32_t storeCap;
	GcK3Frame* stack;       // [stackCap]
	uint32_t stackCap;
	int32_t maxNB;          // largest group size the executor supports
};

// leaf: canonical traceback (up, then left, then diagonal) over the stored columns (leader lane only)
GC_HD bool gc_k3w_leaf_traceback(const GcK3wPass& p, int32_t NB, const GcK3Frame& f, uint8_t* ops, uint32_t& nOps, uint32_t opsCap)
{
	int32_t q = f.q, t = f.t;
	uint32_t start = nOps;
	int32_t i = q - 1, j = t - 1;
	const int32_t INF = 1 << 29;
	int32_t cachedCol = -1, cf = 0, cl = -1;
	auto cell = [&](int32_t ii, int32_t jj) -> int32_t
	{
		if (ii < 0 && jj < 0) return 0;
		if (ii < 0) return jj + 1;
		if (jj < 0) return ii + 1;
		if (jj != cachedCol) { gc_k3w_column_blocks(p, NB, jj, cf, cl); cachedCol = jj; }
		int32_t b = ii >> 6;
		if (b < cf || b > cl) return INF;
		return gc_k3_cell(p.store[(int64_t)jj * p.storeStride + b], ii);
	};
	int32_t cur = cell(i, j);
	if (cur != f.best) return false;
	while (i >= 0 || j >= 0)
	{
		if (nOps >= opsCap) return false;
		if (j < 0) { ops[nOps++] = 1; i--; continue; }
		if (i < 0) { ops[nOps++] = 2; j--; continue; }
		int32_t u = cell(i - 1, j);
		if (u + 1 == cur) { ops[nOps++] = 1; i--; cur = u; continue; }
		int32_t l = cell(i, j - 1);
		if (l + 1 == cur) { ops[nOps++] = 2; j--; cur = l; continue; }
		int32_t ul = cell(i - 1, j - 1);
		if (ul == cur) ops[nOps++] = 0;
		else if (ul + 1 == cur) ops[nOps++] = 3;
		else return false;
		i--; j--; cur = ul;
	}
	for (uint32_t a = start, b = nOps; a + 1 < b; a++, b--) { uint8_t tmp = ops[a]; ops[a] = ops[b - 1]; ops[b - 1] = tmp; }
	return true;
}

GC_HD int32_t gc_k3w_round_nb_path(int32_t nb) { return nb <= 1 ? 1 : nb <= 2 ? 2 : nb <= 4 ? 4 : nb <= 8 ? 8 : 0; }

// One frame of the recursion: either a leaf (its edit operations are appended at ops[nOps..]) or a Hirschberg split
// (two child frames are returned: ch[0] = upper-left, to be aligned FIRST, ch[1] = lower-right).  The depth-first
// driver below and the level-parallel kernel (gcgpu.cu: one warp per frame, one launch per recursion level) share it.
template <class Exec>
GC_HD bool gc_k3w_path_frame(Exec& ex, const GcK3wPathWorkspace& w, const uint8_t* target, const GcK3Frame& f, uint8_t* ops, uint32_t& nOps, uint32_t opsCap, uint64_t& work,
	GcK3Frame ch[2], int32_t& nChildren)
{
	nChildren = 0;
	if (f.q == 0 || f.t == 0)
	{
		uint32_t n = (uint32_t)(f.q + f.t);
		if (nOps + n > opsCap) return false;
		if (ex.leader()) for (uint32_t x = 0; x < n; x++) ops[nOps + x] = f.q == 0 ? 2 : 1;
		nOps += n;
		return true;
	}
	int32_t q = f.q, t = f.t, k = f.best;
	int32_t mx = q > t ? q : t;
	if (k > mx) k = mx;
	int32_t NB = gc_k3w_round_nb_path(gc_k3w_blocks_per_lane(q, t, k));
	if (NB == 0 || NB > w.maxNB) return false;
	int64_t nb = (q + 63) / 64;
	int64_t alignmentDataSize = (2LL * 8 + 4) * nb * t + 2LL * 4 * t;
	if (alignmentDataSize < 1024 * 1024)
	{
		if ((uint64_t)nb * (uint64_t)t > w.storeCap) return false;
		GcK3wPass p = gc_k3w_make_pass(w.peq, w.nbTotal, f.qOff, q, target, f.tOff, 1, t, k, t - 1, NB);
		p.store = w.store; p.storeStride = (int32_t)nb;
		work += ex.pass(p, NB, w.blocksA);
		uint32_t n2 = nOps;
		bool ok = true;
		if (ex.leader()) ok = gc_k3w_leaf_traceback(p, NB, f, ops, n2, opsCap);
		ok = ex.fromLeader((uint32_t)ok) != 0;
		nOps = ex.fromLeader(n2);
		return ok;
	}
	// ---- Hirschberg split (edlib.cpp:1234-1399)
	int32_t leftW = t / 2, rightW = t - leftW;
	GcK3wPass pf = gc_k3w_make_pass(w.peq, w.nbTotal, f.qOff, q, target, f.tOff, 1, t, k, leftW - 1, NB);
	work += ex.pass(pf, NB, w.blocksA);
	int32_t rqOff = w.qTotal - f.qOff - q;
	GcK3wPass pr = gc_k3w_make_pass(w.rpeq, w.nbTotal, rqOff, q, target, (int64_t)f.tOff + t - 1, -1, t, k, rightW - 1, NB);
	work += ex.pass(pr, NB, w.blocksB);
	int32_t lfb, llb, rfb, rlb;
	gc_k3w_stop_blocks(pf, NB, lfb, llb);
	gc_k3w_stop_blocks(pr, NB, rfb, rlb);
	int32_t row = ex.firstSplitRow(w.blocksA, lfb, llb, w.blocksB, rfb, rlb, q, f.best);
	const int32_t INF = 1 << 29;
	int32_t leftScore = -1, rightScore = -1;
	if (row >= 0)
	{
		leftScore = gc_k3_cell(w.blocksA[row >> 6], row);
		int32_t rr = q - 1 - (row + 1);
		rightScore = gc_k3_cell(w.blocksB[rr >> 6], rr);
	}
	else
	{
		row = -2;
		{
			int32_t rr = q - 1; int32_t b = rr >> 6;
			int32_t rs = (b < rfb || b > rlb) ? INF : gc_k3_cell(w.blocksB[b], rr);
			if (leftW + rs == f.best) { row = -1; leftScore = leftW; rightScore = rs; }
		}
		if (row == -2)
		{
			int32_t b = (q - 1) >> 6;
			int32_t ls = (b < lfb || b > llb) ? INF : gc_k3_cell(w.blocksA[b], q - 1);
			if (ls + rightW == f.best) { row = q - 1; leftScore = ls; rightScore = rightW; }
		}
		if (row == -2) return false;
	}
	int32_t ulHeight = row + 1, lrHeight = q - ulHeight;
	ch[0].qOff = f.qOff; ch[0].q = ulHeight; ch[0].tOff = f.tOff; ch[0].t = leftW; ch[0].best = leftScore;
	ch[1].qOff = f.qOff + ulHeight; ch[1].q = lrHeight; ch[1].tOff = f.tOff + leftW; ch[1].t = rightW; ch[1].best = rightScore;
	nChildren = 2;
	return true;
}

// depth-first driver: one warp aligns one pair from start to end
template <class Exec>
GC_HD bool gc_k3w_path(Exec& ex, const GcK3wPathWorkspace& w, const uint8_t* target, int32_t best, uint8_t* ops, uint32_t& nOps, uint32_t opsCap, uint64_t& work)
{
	nOps = 0;
	uint32_t sp = 0;
	if (ex.leader()) { GcK3Frame f0; f0.qOff = 0; f0.q = w.qTotal; f0.tOff = 0; f0.t = w.tTotal; f0.best = best; w.stack[0] = f0; }
	sp = 1;
	ex.sync();
	while (sp > 0)
	{
		GcK3Frame f = w.stack[--sp];
		ex.sync(); // every lane has read the frame before the leader may overwrite the slot
		GcK3Frame ch[2]; int32_t nch = 0;
		if (!gc_k3w_path_frame(ex, w, target, f, ops, nOps, opsCap, work, ch, nch)) return false;
		if (nch)
		{
			if (sp + 2 > w.stackCap) return false;
			if (ex.leader())
			{
				w.stack[sp] = ch[1];     // processed after the upper-left part
				w.stack[sp + 1] = ch[0];
			}
			sp += 2;
			ex.sync();
		}
	}
	return true;
}
