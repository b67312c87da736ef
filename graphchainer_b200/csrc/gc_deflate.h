// Fast gzip member writer for the GAM records (RFC 1951 DEFLATE in an RFC 1952 wrapper).
//
// The reference writes every read's alignments as one gzip member through protobuf's
// GzipOutputStream (src/Aligner.cpp:261-281, zlib level 6, ~15 MB/s per thread); any valid gzip
// member decodes to the same record.  At B200 alignment rates zlib itself (60 MB/s at level 1, 3.4
// raw bytes per read base) was the largest host stage, so level 1 of this driver is a purpose-built
// encoder: greedy LZ77 with a single-probe hash of 4-byte groups (32 KB window), one DEFLATE block
// with a DYNAMIC Huffman code built from the record's own token statistics (the records are dominated
// by a few dozen distinct protobuf tag/varint bytes, where a fixed code gains nothing on literals).
// Levels >= 2 still go through zlib.  tests/test_output.py inflates the members with zlib.
#pragma once
#include <zlib.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace gcdeflate {

// LSB-first bit writer into a caller-sized buffer: one unaligned 8-byte store per put, no per-byte loop
struct BitWriter
{
	uint8_t* p;
	uint64_t acc = 0;
	int n = 0;
	explicit BitWriter(uint8_t* dst) : p(dst) {}
	inline void put(uint64_t bits, int count) // count <= 56, n < 8 on entry
	{
		acc |= bits << n;
		n += count;
		std::memcpy(p, &acc, 8);
		p += n >> 3;
		acc >>= (n & ~7);
		n &= 7;
	}
	uint8_t* flush() { if (n > 0) { *p++ = (uint8_t)(acc & 0xFF); acc = 0; n = 0; } return p; }
};

inline uint32_t reverseBits(uint32_t v, int len)
{
	uint32_t r = 0;
	for (int i = 0; i < len; i++) { r = (r << 1) | (v & 1); v >>= 1; }
	return r;
}

// Length-limited Huffman code lengths of up to 320 symbols, without a heap or an allocation (three codes are built per
// gzip member, i.e. per read): the used symbols sorted by frequency, the classic two-queue merge (leaves in one sorted
// queue, internal nodes appear in non-decreasing weight in the other), depths from the parent links, and for codes
// deeper than maxLen the leaf-moving repair on the per-length counts that keeps the Kraft sum at one; lengths are then
// handed out longest-first to the rarest symbols.
inline void buildLengths(const uint32_t* freq, int n, int maxLen, uint8_t* lens)
{
	struct Leaf { uint32_t w; uint16_t sym; };
	Leaf leaves[320];
	int m = 0;
	for (int i = 0; i < n; i++) { lens[i] = 0; if (freq[i]) { leaves[m].w = freq[i]; leaves[m].sym = (uint16_t)i; m++; } }
	if (m == 0) return;
	if (m == 1) { lens[leaves[0].sym] = 1; return; }
	std::sort(leaves, leaves + m, [](const Leaf& a, const Leaf& b) { return a.w != b.w ? a.w < b.w : a.sym < b.sym; });
	// nodes 0..m-1 = leaves in ascending weight, m..2m-2 = internal nodes in creation order
	uint64_t weight[640]; int16_t parent[640];
	for (int i = 0; i < m; i++) weight[i] = leaves[i].w;
	int nextLeaf = 0, nextInternal = m, made = m;
	auto take = [&]() -> int
	{
		if (nextLeaf < m && (nextInternal >= made || weight[nextLeaf] <= weight[nextInternal])) return nextLeaf++;
		return nextInternal++;
	};
	while (made < 2 * m - 1)
	{
		int a = take(), b = take();
		weight[made] = weight[a] + weight[b];
		parent[a] = (int16_t)made; parent[b] = (int16_t)made;
		made++;
	}
	// depths: the root is the last node; every parent has a larger index than its children
	uint8_t depth[640];
	depth[made - 1] = 0;
	int blCount[64] = { 0 };
	int overflow = 0;
	for (int i = made - 2; i >= 0; i--)
	{
		int d = depth[parent[i]] + 1;
		if (d > maxLen) { d = maxLen; overflow++; } // internal nodes too: the count below is what the repair loop needs to restore the Kraft sum
		depth[i] = (uint8_t)d;
		if (i < m) blCount[d]++;
	}
	if (overflow > 0)
	{
		// every leaf that was cut to maxLen over-subscribes the code: move a leaf one level down from the deepest level that has
		// one to spare, which frees room for a pair at the bottom
		do
		{
			int bits = maxLen - 1;
			while (blCount[bits] == 0) bits--;
			blCount[bits]--;
			blCount[bits + 1] += 2;
			blCount[maxLen]--;
			overflow -= 2;
		} while (overflow > 0);
	}
	// longest codes to the rarest symbols (leaves are in ascending weight)
	int k = 0;
	for (int bits = maxLen; bits >= 1; bits--)
		for (int c = blCount[bits]; c > 0; c--) lens[leaves[k++].sym] = (uint8_t)bits;
}

// canonical codes (RFC 1951 3.2.2), bit-reversed for the LSB-first writer
inline void buildCodes(const uint8_t* lens, int n, uint16_t* codes)
{
	int blCount[16] = { 0 };
	for (int i = 0; i < n; i++) blCount[lens[i]]++;
	blCount[0] = 0;
	uint32_t next[16]; uint32_t code = 0;
	for (int b = 1; b < 16; b++) { code = (code + blCount[b - 1]) << 1; next[b] = code; }
	for (int i = 0; i < n; i++) codes[i] = lens[i] ? (uint16_t)reverseBits(next[lens[i]]++, lens[i]) : 0;
}

struct Tables
{
	uint16_t lenSym[259]; uint8_t lenExtraBits[259]; uint16_t lenBase[259];
	uint8_t distSymLo[512]; uint8_t distSymHi[256];
	uint8_t distExtraBits[32]; uint16_t distBase[32];
	uint8_t symExtraBits[286];
	Tables()
	{
		std::memset(symExtraBits, 0, sizeof(symExtraBits)); std::memset(distExtraBits, 0, sizeof(distExtraBits)); std::memset(distBase, 0, sizeof(distBase));
		static const uint16_t lb[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
		static const uint8_t le[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
		for (int s = 0; s < 29; s++)
		{
			int hi = s == 28 ? 258 : lb[s] + (1 << le[s]) - 1;
			for (int l = lb[s]; l <= hi && l <= 258; l++) { if (s < 28 && l == 258) continue; lenSym[l] = (uint16_t)(257 + s); lenExtraBits[l] = le[s]; lenBase[l] = lb[s]; }
			symExtraBits[257 + s] = le[s];
		}
		lenSym[258] = 285; lenExtraBits[258] = 0; lenBase[258] = 258;
		static const uint16_t db[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
		static const uint8_t de[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
		for (int s = 0; s < 30; s++)
		{
			distExtraBits[s] = de[s]; distBase[s] = db[s];
			for (uint32_t d = db[s]; d < (uint32_t)db[s] + (1u << de[s]); d++)
			{
				if (d <= 512) distSymLo[d - 1] = (uint8_t)s;
				else distSymHi[(d - 1) >> 7] = (uint8_t)s; // codes above 512 cover whole 128-blocks
			}
		}
	}
	inline int distSym(uint32_t d) const { return d <= 512 ? distSymLo[d - 1] : distSymHi[(d - 1) >> 7]; }
};

inline const Tables& tables() { static const Tables t; return t; }

struct Encoder
{
	// token: bits 0-8 literal/length symbol | 9-13 distance symbol (30 = none: a literal) | 14-18 length extra value | 19-31 distance extra value
	// -- everything the emitter needs, so that it runs without a data-dependent branch
	std::vector<uint32_t> tokens;
	std::vector<int32_t> head;
	Encoder() : head(1 << 15, -1) {}

	static inline uint32_t load32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }

	void tokenize(const uint8_t* p, size_t n)
	{
		tokens.clear();
		tokens.reserve(n / 2 + 16);
		std::fill(head.begin(), head.end(), -1);
		size_t i = 0;
		while (i + 4 <= n)
		{
			uint32_t v = load32(p + i);
			uint32_t h = (v * 2654435761u) >> 17;
			int32_t cand = head[h];
			head[h] = (int32_t)i;
			if (cand >= 0 && i - (size_t)cand <= 32768 && load32(p + cand) == v)
			{
				size_t maxLen = std::min<size_t>(258, n - i);
				size_t len = 4;
				while (len + 8 <= maxLen)
				{
					uint64_t a, b; std::memcpy(&a, p + i + len, 8); std::memcpy(&b, p + cand + len, 8);
					if (a != b) { len += (size_t)(__builtin_ctzll(a ^ b) >> 3); goto done; }
					len += 8;
				}
				while (len < maxLen && p[i + len] == p[cand + len]) len++;
			done:
				if (len > maxLen) len = maxLen;
				{
					const Tables& T = tables();
					uint32_t dist = (uint32_t)(i - (size_t)cand);
					uint32_t ds = (uint32_t)T.distSym(dist);
					tokens.push_back((uint32_t)T.lenSym[len] | (ds << 9) | ((uint32_t)(len - T.lenBase[len]) << 14) | ((dist - T.distBase[ds]) << 19));
				}
				// index a couple of positions inside the match so that later data can still find it
				if (i + len + 4 <= n)
				{
					uint32_t v1 = load32(p + i + 1); head[(v1 * 2654435761u) >> 17] = (int32_t)(i + 1);
					uint32_t v2 = load32(p + i + len - 1); head[(v2 * 2654435761u) >> 17] = (int32_t)(i + len - 1);
				}
				i += len;
			}
			else
			{
				tokens.push_back((uint32_t)p[i] | (30u << 9));
				i++;
			}
		}
		for (; i < n; i++) tokens.push_back((uint32_t)p[i] | (30u << 9));
	}

	static bool complete(const uint8_t* lens, int n, int maxLen)
	{
		uint64_t kraft = 0; int used = 0;
		for (int i = 0; i < n; i++) if (lens[i]) { kraft += 1ull << (maxLen - lens[i]); used++; }
		return kraft == (1ull << maxLen) || (used == 1 && maxLen != 7);
	}
	// one final dynamic-Huffman block holding all tokens; false if a code could not be made complete (caller uses zlib)
	bool writeBlock(BitWriter& bw)
	{
		const Tables& T = tables();
		uint32_t litFreq[512] = { 0 }, distFreq[32] = { 0 };
		for (uint32_t t : tokens) { litFreq[t & 0x1FF]++; distFreq[(t >> 9) & 31]++; }
		distFreq[30] = 0;
		litFreq[256] = 1;
		uint8_t litLens[286], distLens[32] = { 0 };
		buildLengths(litFreq, 286, 15, litLens);
		buildLengths(distFreq, 30, 15, distLens);
		int usedDist = 0;
		for (int i = 0; i < 30; i++) if (distLens[i]) usedDist++;
		if (usedDist == 0) distLens[0] = 1;                   // at least one distance code must be described
		if (!complete(litLens, 286, 15) || !complete(distLens, 30, 15)) return false;
		uint16_t litCodes[286], distCodes[32] = { 0 };
		buildCodes(litLens, 286, litCodes);
		buildCodes(distLens, 30, distCodes);
		int hlit = 286; while (hlit > 257 && litLens[hlit - 1] == 0) hlit--;
		int hdist = 30; while (hdist > 1 && distLens[hdist - 1] == 0) hdist--;
		// code length alphabet over the concatenated lengths, with the run symbols 16/17/18
		uint8_t all[316]; int na = 0;
		for (int i = 0; i < hlit; i++) all[na++] = litLens[i];
		for (int i = 0; i < hdist; i++) all[na++] = distLens[i];
		struct Cl { uint8_t sym; uint8_t extra; };
		Cl cl[316]; int ncl = 0;
		for (int i = 0; i < na; )
		{
			int j = i; while (j < na && all[j] == all[i]) j++;
			int run = j - i;
			if (all[i] == 0)
			{
				while (run >= 11) { int r = std::min(run, 138); cl[ncl++] = Cl { 18, (uint8_t)(r - 11) }; run -= r; }
				if (run >= 3) { cl[ncl++] = Cl { 17, (uint8_t)(run - 3) }; run = 0; }
				while (run-- > 0) cl[ncl++] = Cl { 0, 0 };
			}
			else
			{
				cl[ncl++] = Cl { all[i], 0 }; run--;
				while (run >= 3) { int r = std::min(run, 6); cl[ncl++] = Cl { 16, (uint8_t)(r - 3) }; run -= r; }
				while (run-- > 0) cl[ncl++] = Cl { all[i], 0 };
			}
			i = j;
		}
		uint32_t clFreq[19] = { 0 };
		for (int i = 0; i < ncl; i++) clFreq[cl[i].sym]++;
		uint8_t clLens[19]; uint16_t clCodes[19];
		buildLengths(clFreq, 19, 7, clLens);
		{ int used = 0, only = 0; for (int i = 0; i < 19; i++) if (clLens[i]) { used++; only = i; } if (used == 1) clLens[only == 0 ? 1 : 0] = 1; } // the code length code must be complete
		buildCodes(clLens, 19, clCodes);
		if (!complete(clLens, 19, 7)) return false;
		static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
		int hclen = 19; while (hclen > 4 && clLens[order[hclen - 1]] == 0) hclen--;
		bw.put(1, 1);            // BFINAL
		bw.put(2, 2);            // BTYPE = dynamic
		bw.put((uint32_t)(hlit - 257), 5);
		bw.put((uint32_t)(hdist - 1), 5);
		bw.put((uint32_t)(hclen - 4), 4);
		for (int i = 0; i < hclen; i++) bw.put(clLens[order[i]], 3);
		for (int i = 0; i < ncl; i++)
		{
			bw.put(clCodes[cl[i].sym], clLens[cl[i].sym]);
			if (cl[i].sym == 16) bw.put(cl[i].extra, 2);
			else if (cl[i].sym == 17) bw.put(cl[i].extra, 3);
			else if (cl[i].sym == 18) bw.put(cl[i].extra, 7);
		}
		for (uint32_t t : tokens)
		{
			uint32_t sym = t & 0x1FF, ds = (t >> 9) & 31;
			uint64_t v = litCodes[sym]; int nb = litLens[sym];
			v |= (uint64_t)((t >> 14) & 31) << nb; nb += T.symExtraBits[sym];
			v |= (uint64_t)distCodes[ds] << nb; nb += distLens[ds];       // ds == 30 (literal): no bits
			v |= (uint64_t)(t >> 19) << nb; nb += T.distExtraBits[ds];
			bw.put(v, nb);
		}
		bw.put(litCodes[256], litLens[256]);
		return true;
	}

	std::string gzipMember(const std::string& raw)
	{
		tokenize((const uint8_t*)raw.data(), raw.size());
		// worst case: every token a 15-bit literal code + the code descriptions
		std::string out;
		out.resize(10 + tokens.size() * 6 + 1024 + 8);
		static const unsigned char header[10] = { 0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 4, 3 }; // deflate, no flags, mtime 0, XFL fastest, OS unix
		std::memcpy(&out[0], header, 10);
		BitWriter bw((uint8_t*)&out[10]);
		if (!writeBlock(bw)) return std::string();
		uint8_t* end = bw.flush();
		uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)raw.data(), (uInt)raw.size());
		uint32_t isize = (uint32_t)raw.size();
		for (int i = 0; i < 4; i++) *end++ = (uint8_t)((crc >> (8 * i)) & 0xFF);
		for (int i = 0; i < 4; i++) *end++ = (uint8_t)((isize >> (8 * i)) & 0xFF);
		out.resize((size_t)(end - (uint8_t*)&out[0]));
		return out;
	}
};

}
