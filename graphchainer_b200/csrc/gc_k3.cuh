// K3 -- global (NW) sequence-to-sequence alignment by Myers' bit-vector algorithm:
// the reference's two edlibAlign() calls per read (Aligner.cpp:645 TASK_DISTANCE on the
// whole-read path, Aligner.cpp:845 TASK_PATH on the chained path).
//
// What is reproduced and why it is exact:
//  * editDistance (edlib.cpp:193-212): edlib doubles k from 64 until the banded NW pass
//    succeeds; any successful pass returns the TRUE edit distance, so the value does not
//    depend on band bookkeeping.  We run an Ukkonen band of the diagonals a <=k solution
//    can touch ( -(k-(q-t))/2 .. (k+(q-t))/2 ) -- the same initial band as edlib.cpp:755.
//  * alignment path (edlib.cpp:1164-1399): obtainAlignment recurses by Hirschberg when the
//    traceback table would reach 1 MiB ((2*8+4)*blocks*t + 8*t, edlib.cpp:1189-1193),
//    splitting the target at t/2 and the query at the FIRST row r in 0..q-2 with
//    left[r]+right[r+1]==best, else r=-1, else r=q-1 (edlib.cpp:1322-1356); leaves are
//    solved by a traceback that prefers up (insert) over left (delete) over diagonal
//    (edlib.cpp:1023,1057,1088).  Both rules only ever act on cells of optimal paths,
//    whose banded values are exact, so they are functions of the true DP matrix and are
//    reproduced here with our own band.
// Ops: 0 match, 1 insert (consumes a query=path base), 2 delete (consumes a target=read
// base), 3 mismatch (edlib.h EDLIB_EDOP_*).
//
// Sequence codes: 0..3 = A C G T (upper case only -- edlib compares raw characters,
// edlib.cpp:1420-1459), anything else = 4 for the target (matches nothing); the query is
// a graph path and only contains A C G T.
#pragma once
#include "gc_common.cuh"

struct GcK3Block
{
	uint64_t P;
	uint64_t M;
	int32_t score; // value at the last row (bit 63) of the block
	int32_t pad;
};

// edlib.cpp:409-444 calculateBlock: one column step of one 64-row block, hin/hout in {-1,0,1}
GC_HD int gc_k3_block(uint64_t& P, uint64_t& M, uint64_t Eq, int hin)
{
	uint64_t hinIsNeg = (uint64_t)(hin >> 2) & 1ULL;
	uint64_t Xv = Eq | M;
	Eq |= hinIsNeg;
	uint64_t Xh = (((Eq & P) + P) ^ P) | Eq;
	uint64_t Ph = M | ~(Xh | P);
	uint64_t Mh = P & Xh;
	int hout = (int)(Ph >> 63) - (int)(Mh >> 63);
	Ph <<= 1;
	Mh <<= 1;
	Mh |= hinIsNeg;
	Ph |= (uint64_t)((hin + 1) >> 1);
	P = Mh | ~(Xv | Ph);
	M = Ph & Xv;
	return hout;
}

// Query profile: peq[c * nbTotal + b] bit i <=> query[64 b + i] == c, for the WHOLE query;
// sub-queries [qOff, qOff+qLen) are read through a funnel shift.
GC_HD void gc_k3_build_peq(const uint8_t* query, int32_t qLen, uint64_t* peq, int32_t nbTotal)
{
	for (int32_t b = 0; b < 4 * nbTotal; b++) peq[b] = 0;
	for (int32_t i = 0; i < qLen; i++)
	{
		uint8_t c = query[i];
		if (c < 4) peq[(int32_t)c * nbTotal + (i >> 6)] |= 1ULL << (i & 63);
	}
}
GC_HD uint64_t gc_k3_eq(const uint64_t* peq, int32_t nbTotal, int32_t qOff, int32_t qLen, int c, int32_t b)
{
	if (c > 3) return 0;
	const uint64_t* row = peq + (int64_t)c * nbTotal;
	int32_t w = (qOff >> 6) + b;
	int sh = qOff & 63;
	uint64_t v = row[w] >> sh;
	if (sh && w + 1 < nbTotal) v |= row[w + 1] << (64 - sh);
	int32_t remaining = qLen - b * 64;
	if (remaining < 64) v &= (remaining <= 0) ? 0ULL : ((1ULL << remaining) - 1);
	return v;
}

struct GcK3Band { int32_t dlo, dhi; };
GC_HD GcK3Band gc_k3_band(int32_t q, int32_t t, int32_t k)
{
	GcK3Band b;
	b.dlo = -((k - (q - t)) / 2);
	b.dhi = (k + (q - t)) / 2;
	return b;
}
GC_HD int32_t gc_k3_first_block(const GcK3Band& band, int32_t c) { int32_t lo = c + band.dlo; return lo <= 0 ? 0 : lo >> 6; }
GC_HD int32_t gc_k3_last_block(const GcK3Band& band, int32_t q, int32_t c) { int32_t hi = c + band.dhi; if (hi > q - 1) hi = q - 1; return hi >> 6; }

// value of row `i` (global row of the sub-query) from its block
GC_HD int32_t gc_k3_cell(const GcK3Block& bl, int32_t i)
{
	int r = i & 63;
	uint64_t above = r == 63 ? 0ULL : (~0ULL << (r + 1));
	return bl.score - gc_popc(bl.P & above) + gc_popc(bl.M & above);
}

// One banded NW pass of query[qOff, qOff+q) against `t` target symbols read as
// target[tBase + s * tStep], columns 0..stopCol.  `blocks` has ceil(q/64) entries.  If
// `store` is non-null the band blocks of every column are appended to it (leaf traceback):
// column c occupies store[colStart[c] .. ) for blocks firstBlock(c)..lastBlock(c).
// Returns the number of block steps (work units).
GC_HD uint64_t gc_k3_pass(const uint64_t* peq, int32_t nbTotal, int32_t qOff, int32_t q, const uint8_t* target, int64_t tBase, int32_t tStep, int32_t t, int32_t k, int32_t stopCol,
	GcK3Block* blocks, GcK3Block* store, uint32_t* colStart)
{
	GcK3Band band = gc_k3_band(q, t, k);
	int32_t lb = gc_k3_last_block(band, q, 0);
	for (int32_t b = 0; b <= lb; b++) { blocks[b].P = ~0ULL; blocks[b].M = 0; blocks[b].score = (b + 1) * 64; }
	uint64_t work = 0;
	uint32_t stored = 0;
	for (int32_t c = 0; c <= stopCol; c++)
	{
		int32_t fb = gc_k3_first_block(band, c);
		int32_t nlb = gc_k3_last_block(band, q, c);
		while (lb < nlb)
		{
			lb++;
			blocks[lb].P = ~0ULL; blocks[lb].M = 0; blocks[lb].score = blocks[lb - 1].score + 64;
		}
		int sym = target[tBase + (int64_t)c * tStep];
		int hin = 1;
		for (int32_t b = fb; b <= lb; b++)
		{
			uint64_t Eq = gc_k3_eq(peq, nbTotal, qOff, q, sym, b);
			uint64_t P = blocks[b].P, M = blocks[b].M;
			hin = gc_k3_block(P, M, Eq, hin);
			blocks[b].P = P; blocks[b].M = M; blocks[b].score += hin;
		}
		work += (uint64_t)(lb - fb + 1);
		if (store)
		{
			colStart[c] = stored;
			for (int32_t b = fb; b <= lb; b++) store[stored++] = blocks[b];
		}
	}
	return work;
}

// edit distance with cutoff k: exact value if <= k, else -1 (myersCalcEditDistanceNW semantics)
GC_HD int32_t gc_k3_distance_k(const uint64_t* peq, int32_t nbTotal, int32_t qOff, int32_t q, const uint8_t* target, int64_t tBase, int32_t tStep, int32_t t, int32_t k, GcK3Block* blocks, uint64_t& work)
{
	int32_t diff = q > t ? q - t : t - q;
	if (k < diff) return -1;
	int32_t mx = q > t ? q : t;
	if (k > mx) k = mx;
	work += gc_k3_pass(peq, nbTotal, qOff, q, target, tBase, tStep, t, k, t - 1, blocks, nullptr, nullptr);
	int32_t v = gc_k3_cell(blocks[(q - 1) >> 6], q - 1);
	return v <= k ? v : -1;
}

// edlibAlign(..., k=-1, NW, TASK_DISTANCE).editDistance  (edlib.cpp:141-212)
GC_HD int32_t gc_k3_distance(const uint64_t* peq, int32_t nbTotal, int32_t q, const uint8_t* target, int32_t t, GcK3Block* blocks, int32_t kStart, uint64_t& work)
{
	if (q == 0 || t == 0) return q > t ? q : t;
	int32_t k = kStart < 64 ? 64 : kStart;
	while (true)
	{
		int32_t d = gc_k3_distance_k(peq, nbTotal, 0, q, target, 0, 1, t, k, blocks, work);
		if (d >= 0) return d;
		k *= 2;
	}
}

// ------------------------------------------------------------------ alignment path
struct GcK3Frame { int32_t qOff, q, tOff, t, best; };

struct GcK3PathWorkspace
{
	const uint64_t* peq;    // [4 * nbTotal] profile of the query
	const uint64_t* rpeq;   // [4 * nbTotal] profile of the reversed query
	int32_t nbTotal;
	int32_t qTotal, tTotal;
	GcK3Block* blocksA;     // [nbTotal]
	GcK3Block* blocksB;     // [nbTotal]
	GcK3Block* store;       // [storeCap] leaf columns
	uint32_t* colStart;     // [leaf columns]
	uint32_t storeCap;
	uint32_t colCap;
	GcK3Frame* stack;       // [stackCap]
	uint32_t stackCap;
};

// leaf: canonical traceback (up, then left, then diagonal) over the stored band; ops are
// produced last-to-first and reversed in place.  Returns false on an internal inconsistency.
GC_HD bool gc_k3_leaf(const GcK3PathWorkspace& w, const uint8_t* target, const GcK3Frame& f, uint8_t* ops, uint32_t& nOps, uint32_t opsCap, uint64_t& work)
{
	int32_t q = f.q, t = f.t, k = f.best;
	int32_t mx = q > t ? q : t;
	if (k > mx) k = mx;
	GcK3Band band = gc_k3_band(q, t, k);
	// storage needed
	{
		uint64_t need = 0;
		for (int32_t c = 0; c < t; c++) need += (uint64_t)(gc_k3_last_block(band, q, c) - gc_k3_first_block(band, c) + 1);
		if (need > w.storeCap || (uint32_t)t > w.colCap) return false;
	}
	work += gc_k3_pass(w.peq, w.nbTotal, f.qOff, q, target, f.tOff, 1, t, k, t - 1, w.blocksA, w.store, w.colStart);
	uint32_t start = nOps;
	int32_t i = q - 1, j = t - 1;
	const int32_t INF = 1 << 29;
	// cell(i,j) with boundaries D[i][-1] = i+1, D[-1][j] = j+1, D[-1][-1] = 0
	auto cell = [&](int32_t ii, int32_t jj) -> int32_t
	{
		if (ii < 0 && jj < 0) return 0;
		if (ii < 0) return jj + 1;
		if (jj < 0) return ii + 1;
		int32_t fb = gc_k3_first_block(band, jj), lb = gc_k3_last_block(band, q, jj);
		int32_t b = ii >> 6;
		if (b < fb || b > lb) return INF;
		return gc_k3_cell(w.store[w.colStart[jj] + (uint32_t)(b - fb)], ii);
	};
	int32_t cur = cell(i, j);
	if (cur != f.best) return false;
	while (i >= 0 || j >= 0)
	{
		if (nOps >= opsCap) return false;
		if (j < 0) { ops[nOps++] = 1; i--; continue; }      // only boundary cells left: move up
		if (i < 0) { ops[nOps++] = 2; j--; continue; }      // move left
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
	// reverse this leaf's ops
	for (uint32_t a = start, b = nOps; a + 1 < b; a++, b--) { uint8_t tmp = ops[a]; ops[a] = ops[b - 1]; ops[b - 1] = tmp; }
	return true;
}

// edlibAlign(..., NW, TASK_PATH).alignment for a known distance `best` (obtainAlignment, edlib.cpp:1164-1216).
// `rtarget` is not needed: the reverse pass walks the target backwards.
GC_HD bool gc_k3_path(const GcK3PathWorkspace& w, const uint8_t* target, int32_t best, uint8_t* ops, uint32_t& nOps, uint32_t opsCap, uint64_t& work)
{
	nOps = 0;
	uint32_t sp = 0;
	GcK3Frame f0; f0.qOff = 0; f0.q = w.qTotal; f0.tOff = 0; f0.t = w.tTotal; f0.best = best;
	w.stack[sp++] = f0;
	while (sp > 0)
	{
		GcK3Frame f = w.stack[--sp];
		if (f.q == 0 || f.t == 0)
		{
			uint32_t n = (uint32_t)(f.q + f.t);
			if (nOps + n > opsCap) return false;
			for (uint32_t x = 0; x < n; x++) ops[nOps++] = f.q == 0 ? 2 : 1;
			continue;
		}
		int64_t nb = (f.q + 63) / 64;
		int64_t alignmentDataSize = (2LL * 8 + 4) * nb * f.t + 2LL * 4 * f.t;
		if (alignmentDataSize < 1024 * 1024)
		{
			if (!gc_k3_leaf(w, target, f, ops, nOps, opsCap, work)) return false;
			continue;
		}
		// ---- Hirschberg split (edlib.cpp:1234-1399)
		int32_t q = f.q, t = f.t, k = f.best;
		int32_t mx = q > t ? q : t;
		if (k > mx) k = mx;
		int32_t leftW = t / 2, rightW = t - leftW;
		GcK3Band band = gc_k3_band(q, t, k);
		// forward: columns 0..leftW-1 of query[qOff..] vs target[tOff..]
		work += gc_k3_pass(w.peq, w.nbTotal, f.qOff, q, target, f.tOff, 1, t, k, leftW - 1, w.blocksA, nullptr, nullptr);
		// reverse: reversed query vs reversed target, columns 0..rightW-1
		int32_t rqOff = w.qTotal - f.qOff - q;
		work += gc_k3_pass(w.rpeq, w.nbTotal, rqOff, q, target, (int64_t)f.tOff + t - 1, -1, t, k, rightW - 1, w.blocksB, nullptr, nullptr);
		int32_t lfb = gc_k3_first_block(band, leftW - 1), llb = gc_k3_last_block(band, q, leftW - 1);
		int32_t rfb = gc_k3_first_block(band, rightW - 1), rlb = gc_k3_last_block(band, q, rightW - 1);
		const int32_t INF = 1 << 29;
		auto leftScoreAt = [&](int32_t r) -> int32_t { int32_t b = r >> 6; if (b < lfb || b > llb) return INF; return gc_k3_cell(w.blocksA[b], r); };
		// right[r] = cost of query[r..q) vs right half = reversed row q-1-r
		auto rightScoreAt = [&](int32_t r) -> int32_t { int32_t rr = q - 1 - r; int32_t b = rr >> 6; if (b < rfb || b > rlb) return INF; return gc_k3_cell(w.blocksB[b], rr); };
		int32_t row = -2, leftScore = -1, rightScore = -1;
		for (int32_t r = 0; r <= q - 2; r++)
		{
			int32_t ls = leftScoreAt(r), rs = rightScoreAt(r + 1);
			if (ls + rs == f.best) { row = r; leftScore = ls; rightScore = rs; break; }
		}
		if (row == -2)
		{
			int32_t rs = rightScoreAt(0);
			if (leftW + rs == f.best) { row = -1; leftScore = leftW; rightScore = rs; }
		}
		if (row == -2)
		{
			int32_t ls = leftScoreAt(q - 1);
			if (ls + rightW == f.best) { row = q - 1; leftScore = ls; rightScore = rightW; }
		}
		if (row == -2) return false;
		int32_t ulHeight = row + 1, lrHeight = q - ulHeight;
		if (sp + 2 > w.stackCap) return false;
		GcK3Frame lr; lr.qOff = f.qOff + ulHeight; lr.q = lrHeight; lr.tOff = f.tOff + leftW; lr.t = rightW; lr.best = rightScore;
		GcK3Frame ul; ul.qOff = f.qOff; ul.q = ulHeight; ul.tOff = f.tOff; ul.t = leftW; ul.best = leftScore;
		w.stack[sp++] = lr; // processed after ul
		w.stack[sp++] = ul;
	}
	return true;
}
