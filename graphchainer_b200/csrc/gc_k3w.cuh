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
};

GC_HD int32_t gc_k3w_blocks_per_lane(int32_t q, int32_t t, int32_t k)
{
	GcK3Band band = gc_k3_band(q, t, k);
	int32_t width = band.dhi - band.dlo;
	if (width <= 32) return 1;
	return (width - 32 + 1983) / 1984;
}

GC_HD GcK3wPass gc_k3w_make_pass(const uint64_t* peq, int32_t nbTotal, int32_t qOff, int32_t q, const uint8_t* target, int64_t tBase, int32_t tStep, int32_t t, int32_t k, int32_t stopCol, int32_t NB)
{
	GcK3wPass p;
	p.peq = peq; p.nbTotal = nbTotal; p.qOff = qOff; p.q = q; p.target = target; p.tBase = tBase; p.tStep = tStep; p.t = t; p.stopCol = stopCol;
	GcK3Band band = gc_k3_band(q, t, k);
	p.dlo = band.dlo; p.dhi = band.dhi;
	p.nb = (q + 63) / 64;
	p.numGroups = (p.nb + NB - 1) / NB;
	int32_t gLast = (stopCol + p.dhi) / (64 * NB);
	if (gLast > p.numGroups - 1) gLast = p.numGroups - 1;
	p.tauEnd = stopCol + gLast;
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
		if (c == p.stopCol)
		{
			GC_UNROLL
			for (int i = 0; i < NB; i++)
				if (i < s.nbHere) { GcK3Block bl; bl.P = s.P[i]; bl.M = s.M[i]; bl.score = s.score[i]; bl.pad = 0; blocksOut[s.g * NB + i] = bl; }
		}
		if (c == s.cLast) gc_k3w_set_group(p, s, s.g + GC_K3W_LANES);
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
