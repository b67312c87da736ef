// S7 on the device: one read's GAM record -- the vg::Alignment messages of its alignments in proto3 wire format, packed
// into one gzip member -- made by ONE warp from the alignment token streams that gc_post.cuh leaves in HBM.
//   GraphAlignerVGAlignment::traceToAlignment + AddAlignment + replaceDigraphNodeIdsWithOriginalNodeIds
//       (src/GraphAlignerVGAlignment.h:37-165, src/GraphAligner.h:205-212, src/Aligner.cpp:152-165)
//   writeGAMToQueue: varint64 count, {varint32 size, message}*, one gzip member per read (src/Aligner.cpp:261-281, stream.hpp:24-51)
//
// Why on the device: encoding and compressing the records was the largest host stage left (36 of ~63 ns per read base and
// core; gc_output.h / gc_deflate.h), every record is independent, and the device idles between the DP launches of a batch.
// A record is ~3.4 bytes per read base.  Lane 0 of the warp writes the record (a chain of varint fields whose sizes depend on
// each other); the DEFLATE encoder is the host's design (greedy LZ77 with one hash probe per position, one dynamic-Huffman block
// from the record's own statistics) cut into 32 chunks for the 32 lanes -- see gc_gzip_member below.  The same functions,
// compiled for the host, run in the C-ABI test double and in tests/hostsim/gz_check.cpp, where zlib inflates their output.
#pragma once
#include <string.h>
#include "gc_common.cuh"
#include "gc_post.cuh"

// ---------------------------------------------------------------- tables (built on the host once: gcBuildGamTables)
struct GcDeflateTables
{
	uint16_t lenSym[259]; uint8_t lenExtraBits[259]; uint16_t lenBase[259];
	uint8_t distSymLo[512]; uint8_t distSymHi[256];
	uint8_t distExtraBits[32]; uint16_t distBase[32];
	uint8_t symExtraBits[288];
	uint32_t crc[256];
	uint32_t x2n[32];   // x^(2^k) modulo the CRC-32 polynomial (reflected): joins the CRCs of a record's chunks
};
inline void gcBuildGamTables(GcDeflateTables& T)
{
	for (size_t i = 0; i < sizeof(T); i++) ((uint8_t*)&T)[i] = 0;
	static const uint16_t lb[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
	static const uint8_t le[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
	for (int s = 0; s < 29; s++)
	{
		int hi = s == 28 ? 258 : lb[s] + (1 << le[s]) - 1;
		for (int l = lb[s]; l <= hi && l <= 258; l++) { if (s < 28 && l == 258) continue; T.lenSym[l] = (uint16_t)(257 + s); T.lenExtraBits[l] = le[s]; T.lenBase[l] = lb[s]; }
		T.symExtraBits[257 + s] = le[s];
	}
	T.lenSym[258] = 285; T.lenExtraBits[258] = 0; T.lenBase[258] = 258;
	static const uint16_t db[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
	static const uint8_t de[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
	for (int s = 0; s < 30; s++)
	{
		T.distExtraBits[s] = de[s]; T.distBase[s] = db[s];
		for (uint32_t d = db[s]; d < (uint32_t)db[s] + (1u << de[s]); d++)
		{
			if (d <= 512) T.distSymLo[d - 1] = (uint8_t)s;
			else T.distSymHi[(d - 1) >> 7] = (uint8_t)s; // codes above 512 cover whole 128-blocks
		}
	}
	for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; T.crc[i] = c; }
	auto mul = [](uint32_t a, uint32_t b) { uint32_t m = 1u << 31, r = 0; for (;;) { if (a & m) { r ^= b; if ((a & (m - 1)) == 0) break; } m >>= 1; b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1; } return r; };
	uint32_t x = 1u << 30; // x^1
	T.x2n[0] = x;
	for (int k = 1; k < 32; k++) { x = mul(x, x); T.x2n[k] = x; }
}

// ---------------------------------------------------------------- proto3 wire format of the record
struct GcNameTable
{
	const int32_t* origIndexOfId;  // digraph node id -> index of the original node (-1: none)
	const uint32_t* nameOff;       // [numOrig + 1] into nameChars
	const uint8_t* nameChars;      // GFA segment names
};
// one alignment of a read: its token stream (gc_tokenize) and the scalar fields of the message
struct GcGamAln
{
	uint64_t tokenOff; uint32_t numTokens;
	int32_t start, end;      // AlignmentItem::alignmentStart / alignmentEnd
	int32_t traceScore;      // vg::Alignment::score (0 is omitted)
	uint32_t matches, steps; // identity = matches / steps (GraphAlignerVGAlignment.h:150)
};

GC_HD uint32_t gc_gam_vsize(uint64_t v) { uint32_t n = 1; while (v >= 0x80) { v >>= 7; n++; } return n; }
GC_HD uint8_t* gc_gam_putv(uint8_t* p, uint64_t v) { while (v >= 0x80) { *p++ = (uint8_t)((v & 0x7F) | 0x80); v >>= 7; } *p++ = (uint8_t)v; return p; }

GC_HD uint32_t gc_gam_position_size(const GcNameTable& nt, uint32_t digraphNode, uint32_t offset, uint32_t& nameLen, uint32_t& nameStart)
{
	int64_t nodeId = digraphNode / 2;
	int32_t oi = nt.origIndexOfId[digraphNode];
	nameStart = oi >= 0 ? nt.nameOff[oi] : 0;
	nameLen = oi >= 0 ? nt.nameOff[oi + 1] - nameStart : 0;
	return (nodeId ? 1 + gc_gam_vsize((uint64_t)nodeId) : 0) + (offset ? 1 + gc_gam_vsize(offset) : 0) + ((digraphNode & 1) ? 2 : 0) + (nameLen ? 1 + gc_gam_vsize(nameLen) + nameLen : 0);
}
GC_HD uint32_t gc_gam_edit_size(uint32_t type, uint32_t len)
{
	uint32_t from = type == GC_EDIT_INSERTION ? 0 : len, to = type == GC_EDIT_DELETION ? 0 : len;
	uint32_t seqLen = (type == GC_EDIT_MISMATCH || type == GC_EDIT_INSERTION) ? len : 0;
	return (from ? 1 + gc_gam_vsize(from) : 0) + (to ? 1 + gc_gam_vsize(to) : 0) + (seqLen ? 1 + gc_gam_vsize(seqLen) + seqLen : 0);
}

// size of one alignment's message; also the size of its path submessage
GC_HD uint32_t gc_gam_message_size(const GcNameTable& nt, const GcGamAln& a, const uint32_t* tok, uint32_t nameLenRead, uint32_t& pathSize)
{
	pathSize = 0;
	uint32_t i = 0, rank = 0;
	while (i < a.numTokens)
	{
		// a mapping: marker, digraph node, offset, then its edits
		uint32_t nl, ns;
		uint32_t posSize = gc_gam_position_size(nt, tok[i + 1], tok[i + 2], nl, ns);
		i += 3;
		uint32_t mm = 1 + gc_gam_vsize(posSize) + posSize;
		while (i < a.numTokens && (tok[i] & 0x3FFFFFFFu) != 0) { uint32_t e = gc_gam_edit_size(tok[i] >> 30, tok[i] & 0x3FFFFFFFu); mm += 1 + gc_gam_vsize(e) + e; i++; }
		if (rank) mm += 1 + gc_gam_vsize(rank);
		rank++;
		pathSize += 1 + gc_gam_vsize(mm) + mm;
	}
	uint32_t alnLen = (uint32_t)(a.end - a.start);
	double identity = (double)a.matches / (double)a.steps;
	uint64_t bits; memcpy(&bits, &identity, 8);
	return (alnLen ? 1 + gc_gam_vsize(alnLen) + alnLen : 0) + 1 + gc_gam_vsize(pathSize) + pathSize + (nameLenRead ? 1 + gc_gam_vsize(nameLenRead) + nameLenRead : 0)
		+ (a.traceScore ? 1 + gc_gam_vsize((uint64_t)(int64_t)a.traceScore) : 0) + (a.start ? 1 + gc_gam_vsize((uint64_t)(int64_t)a.start) : 0) + (bits ? 10 : 0);
}
// bytes of the whole record (for sizing the output buffers)
GC_HD uint32_t gc_gam_record_size(const GcNameTable& nt, const GcGamAln* alns, uint32_t nAlns, const uint32_t* tokens, uint32_t nameLenRead)
{
	uint32_t total = gc_gam_vsize(nAlns);
	for (uint32_t k = 0; k < nAlns; k++) { uint32_t ps; uint32_t m = gc_gam_message_size(nt, alns[k], tokens + alns[k].tokenOff, nameLenRead, ps); total += gc_gam_vsize(m) + m; }
	return total;
}
// writes the record, returns its length.  readChars = the read's characters as given by the caller (the sequence field and the
// characters of mismatch / insertion edits are copied from it).
GC_HD uint32_t gc_gam_write_record(const GcNameTable& nt, const uint8_t* readChars, const uint8_t* readName, uint32_t nameLenRead, const GcGamAln* alns, uint32_t nAlns, const uint32_t* tokens, uint8_t* out)
{
	uint8_t* p = gc_gam_putv(out, nAlns);
	for (uint32_t k = 0; k < nAlns; k++)
	{
		const GcGamAln& a = alns[k];
		const uint32_t* tok = tokens + a.tokenOff;
		uint32_t pathSize;
		uint32_t msgSize = gc_gam_message_size(nt, a, tok, nameLenRead, pathSize);
		p = gc_gam_putv(p, msgSize);
		uint32_t alnLen = (uint32_t)(a.end - a.start);
		if (alnLen) { *p++ = 0x0A; p = gc_gam_putv(p, alnLen); for (uint32_t c = 0; c < alnLen; c++) p[c] = readChars[a.start + c]; p += alnLen; }
		*p++ = 0x12; p = gc_gam_putv(p, pathSize);
		uint32_t i = 0, rank = 0, nextChar = (uint32_t)a.start;
		bool firstEdit = true;
		while (i < a.numTokens)
		{
			uint32_t digraphNode = tok[i + 1], offset = tok[i + 2];
			uint32_t nl, ns;
			uint32_t posSize = gc_gam_position_size(nt, digraphNode, offset, nl, ns);
			// size of this mapping: position + its edits (+ rank)
			uint32_t j = i + 3, mm = 1 + gc_gam_vsize(posSize) + posSize;
			while (j < a.numTokens && (tok[j] & 0x3FFFFFFFu) != 0) { uint32_t e = gc_gam_edit_size(tok[j] >> 30, tok[j] & 0x3FFFFFFFu); mm += 1 + gc_gam_vsize(e) + e; j++; }
			if (rank) mm += 1 + gc_gam_vsize(rank);
			*p++ = 0x12; p = gc_gam_putv(p, mm);
			*p++ = 0x0A; p = gc_gam_putv(p, posSize);
			uint64_t nodeId = digraphNode / 2;
			if (nodeId) { *p++ = 0x08; p = gc_gam_putv(p, nodeId); }
			if (offset) { *p++ = 0x10; p = gc_gam_putv(p, offset); }
			if (digraphNode & 1) { *p++ = 0x20; *p++ = 1; }
			if (nl) { *p++ = 0x2A; p = gc_gam_putv(p, nl); for (uint32_t c = 0; c < nl; c++) p[c] = nt.nameChars[ns + c]; p += nl; }
			for (i += 3; i < j; i++)
			{
				uint32_t type = tok[i] >> 30, len = tok[i] & 0x3FFFFFFFu;
				uint32_t from = type == GC_EDIT_INSERTION ? 0 : len, to = type == GC_EDIT_DELETION ? 0 : len;
				uint32_t seqLen = (type == GC_EDIT_MISMATCH || type == GC_EDIT_INSERTION) ? len : 0;
				*p++ = 0x12; p = gc_gam_putv(p, gc_gam_edit_size(type, len));
				if (from) { *p++ = 0x08; p = gc_gam_putv(p, from); }
				if (to) { *p++ = 0x10; p = gc_gam_putv(p, to); }
				if (seqLen)
				{
					*p++ = 0x1A; p = gc_gam_putv(p, seqLen);
					for (uint32_t c = 0; c < seqLen; c++) p[c] = readChars[nextChar + c];
					if (firstEdit && type == GC_EDIT_MISMATCH) p[0] = readChars[0]; // sic, GraphAlignerVGAlignment.h:75
					p += seqLen;
				}
				nextChar += to;
				firstEdit = false;
			}
			if (rank) { *p++ = 0x28; p = gc_gam_putv(p, rank); }
			rank++;
		}
		if (nameLenRead) { *p++ = 0x1A; p = gc_gam_putv(p, nameLenRead); for (uint32_t c = 0; c < nameLenRead; c++) p[c] = readName[c]; p += nameLenRead; }
		if (a.traceScore) { *p++ = 0x30; p = gc_gam_putv(p, (uint64_t)(int64_t)a.traceScore); }
		if (a.start) { *p++ = 0x38; p = gc_gam_putv(p, (uint64_t)(int64_t)a.start); }
		double identity = (double)a.matches / (double)a.steps;
		uint64_t bits; memcpy(&bits, &identity, 8);
		if (bits) { *p++ = 0x81; *p++ = 0x01; for (int b = 0; b < 8; b++) *p++ = (uint8_t)(bits >> (8 * b)); }
	}
	return (uint32_t)(p - out);
}

// ---------------------------------------------------------------- DEFLATE (RFC 1951) in a gzip member (RFC 1952)
// workspace of one record: hash heads + token buffer
struct GcDeflateWs { uint32_t* tokens; uint32_t tokenCap; };
GC_HD size_t gc_deflate_ws_bytes(uint32_t rawBytes) { return ((size_t)rawBytes + 16) * 4; }

struct GcBitWriter
{
	uint8_t* p; uint64_t acc; int n;
	GC_HD void put(uint64_t bits, int count) // count <= 48
	{
		acc |= bits << n;
		n += count;
		while (n >= 8) { *p++ = (uint8_t)acc; acc >>= 8; n -= 8; }
	}
	GC_HD uint8_t* flush() { if (n > 0) { *p++ = (uint8_t)acc; acc = 0; n = 0; } return p; }
};

GC_HD uint32_t gc_deflate_load32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
GC_HD int gc_deflate_dist_sym(const GcDeflateTables& T, uint32_t d) { return d <= 512 ? T.distSymLo[d - 1] : T.distSymHi[(d - 1) >> 7]; }

// token: bits 0-8 literal/length symbol | 9-13 distance symbol (30 = a literal) | 14-18 length extra value | 19-31 distance extra value

// length-limited Huffman code lengths (two-queue merge over the symbols sorted by frequency; the leaf-moving repair of the per-length
// counts for codes deeper than maxLen).  n <= 288.
GC_HD void gc_deflate_lengths(const uint32_t* freq, int n, int maxLen, uint8_t* lens)
{
	uint32_t w[288]; uint16_t sym[288];
	int m = 0;
	for (int i = 0; i < n; i++) { lens[i] = 0; if (freq[i]) { w[m] = freq[i]; sym[m] = (uint16_t)i; m++; } }
	if (m == 0) return;
	if (m == 1) { lens[sym[0]] = 1; return; }
	// shell sort by (weight, symbol)
	for (int gap = m / 2; gap > 0; gap /= 2)
		for (int i = gap; i < m; i++)
		{
			uint32_t cw = w[i]; uint16_t cs = sym[i];
			int j = i;
			while (j >= gap && (w[j - gap] > cw || (w[j - gap] == cw && sym[j - gap] > cs))) { w[j] = w[j - gap]; sym[j] = sym[j - gap]; j -= gap; }
			w[j] = cw; sym[j] = cs;
		}
	uint64_t weight[576]; int16_t parent[576]; uint8_t depth[576];
	for (int i = 0; i < m; i++) weight[i] = w[i];
	int nextLeaf = 0, nextInternal = m, made = m;
	while (made < 2 * m - 1)
	{
		int pick[2];
		for (int t = 0; t < 2; t++)
		{
			if (nextLeaf < m && (nextInternal >= made || weight[nextLeaf] <= weight[nextInternal])) pick[t] = nextLeaf++;
			else pick[t] = nextInternal++;
		}
		weight[made] = weight[pick[0]] + weight[pick[1]];
		parent[pick[0]] = (int16_t)made; parent[pick[1]] = (int16_t)made;
		made++;
	}
	int blCount[17];
	for (int b = 0; b < 17; b++) blCount[b] = 0;
	depth[made - 1] = 0;
	int overflow = 0;
	for (int i = made - 2; i >= 0; i--)
	{
		int d = depth[parent[i]] + 1;
		if (d > maxLen) { d = maxLen; overflow++; }
		depth[i] = (uint8_t)d;
		if (i < m) blCount[d]++;
	}
	while (overflow > 0)
	{
		int bits = maxLen - 1;
		while (blCount[bits] == 0) bits--;
		blCount[bits]--;
		blCount[bits + 1] += 2;
		blCount[maxLen]--;
		overflow -= 2;
	}
	int k = 0;
	for (int bits = maxLen; bits >= 1; bits--)
		for (int c = blCount[bits]; c > 0; c--) lens[sym[k++]] = (uint8_t)bits;
}
GC_HD uint32_t gc_deflate_reverse(uint32_t v, int len) { uint32_t r = 0; for (int i = 0; i < len; i++) { r = (r << 1) | (v & 1); v >>= 1; } return r; }
GC_HD void gc_deflate_codes(const uint8_t* lens, int n, uint16_t* codes)
{
	int blCount[16];
	for (int b = 0; b < 16; b++) blCount[b] = 0;
	for (int i = 0; i < n; i++) blCount[lens[i]]++;
	blCount[0] = 0;
	uint32_t next[16]; uint32_t code = 0;
	next[0] = 0;
	for (int b = 1; b < 16; b++) { code = (code + blCount[b - 1]) << 1; next[b] = code; }
	for (int i = 0; i < n; i++) codes[i] = lens[i] ? (uint16_t)gc_deflate_reverse(next[lens[i]]++, lens[i]) : 0;
}
GC_HD bool gc_deflate_complete(const uint8_t* lens, int n, int maxLen)
{
	uint64_t kraft = 0; int used = 0;
	for (int i = 0; i < n; i++) if (lens[i]) { kraft += 1ull << (maxLen - lens[i]); used++; }
	return kraft == (1ull << maxLen) || (used == 1 && maxLen != 7);
}

// the dynamic-Huffman codes of a block from its symbol counts, and the block's header (BFINAL = 1, the code descriptions);
// false if a code could not be made complete
GC_HD bool gc_deflate_block_header(uint32_t* litFreq, uint32_t* distFreq, uint8_t* litLens, uint8_t* distLens, uint16_t* litCodes, uint16_t* distCodes, GcBitWriter& bw)
{
	distFreq[30] = 0; distFreq[31] = 0;
	litFreq[256] = 1;
	for (int i = 0; i < 32; i++) distLens[i] = 0;
	gc_deflate_lengths(litFreq, 286, 15, litLens);
	gc_deflate_lengths(distFreq, 30, 15, distLens);
	litLens[286] = litLens[287] = 0;
	int usedDist = 0;
	for (int i = 0; i < 30; i++) if (distLens[i]) usedDist++;
	if (usedDist == 0) distLens[0] = 1; // at least one distance code must be described
	if (!gc_deflate_complete(litLens, 286, 15) || !gc_deflate_complete(distLens, 30, 15)) return false;
	for (int i = 0; i < 32; i++) distCodes[i] = 0;
	gc_deflate_codes(litLens, 286, litCodes);
	gc_deflate_codes(distLens, 30, distCodes);
	int hlit = 286; while (hlit > 257 && litLens[hlit - 1] == 0) hlit--;
	int hdist = 30; while (hdist > 1 && distLens[hdist - 1] == 0) hdist--;
	// code length alphabet over the concatenated lengths, with the run symbols 16/17/18
	uint8_t all[320]; int na = 0;
	for (int i = 0; i < hlit; i++) all[na++] = litLens[i];
	for (int i = 0; i < hdist; i++) all[na++] = distLens[i];
	uint8_t clSym[320], clExtra[320]; int ncl = 0;
	for (int i = 0; i < na; )
	{
		int j = i; while (j < na && all[j] == all[i]) j++;
		int run = j - i;
		if (all[i] == 0)
		{
			while (run >= 11) { int r = run < 138 ? run : 138; clSym[ncl] = 18; clExtra[ncl++] = (uint8_t)(r - 11); run -= r; }
			if (run >= 3) { clSym[ncl] = 17; clExtra[ncl++] = (uint8_t)(run - 3); run = 0; }
			while (run-- > 0) { clSym[ncl] = 0; clExtra[ncl++] = 0; }
		}
		else
		{
			clSym[ncl] = all[i]; clExtra[ncl++] = 0; run--;
			while (run >= 3) { int r = run < 6 ? run : 6; clSym[ncl] = 16; clExtra[ncl++] = (uint8_t)(r - 3); run -= r; }
			while (run-- > 0) { clSym[ncl] = all[i]; clExtra[ncl++] = 0; }
		}
		i = j;
	}
	uint32_t clFreq[19];
	for (int i = 0; i < 19; i++) clFreq[i] = 0;
	for (int i = 0; i < ncl; i++) clFreq[clSym[i]]++;
	uint8_t clLens[19]; uint16_t clCodes[19];
	gc_deflate_lengths(clFreq, 19, 7, clLens);
	{ int used = 0, only = 0; for (int i = 0; i < 19; i++) if (clLens[i]) { used++; only = i; } if (used == 1) clLens[only == 0 ? 1 : 0] = 1; } // the code length code must be complete
	gc_deflate_codes(clLens, 19, clCodes);
	if (!gc_deflate_complete(clLens, 19, 7)) return false;
	const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
	int hclen = 19; while (hclen > 4 && clLens[order[hclen - 1]] == 0) hclen--;
	bw.put(1, 1);            // BFINAL
	bw.put(2, 2);            // BTYPE = dynamic
	bw.put((uint32_t)(hlit - 257), 5);
	bw.put((uint32_t)(hdist - 1), 5);
	bw.put((uint32_t)(hclen - 4), 4);
	for (int i = 0; i < hclen; i++) bw.put(clLens[order[i]], 3);
	for (int i = 0; i < ncl; i++)
	{
		bw.put(clCodes[clSym[i]], clLens[clSym[i]]);
		if (clSym[i] == 16) bw.put(clExtra[i], 2);
		else if (clSym[i] == 17) bw.put(clExtra[i], 3);
		else if (clSym[i] == 18) bw.put(clExtra[i], 7);
	}
	return true;
}

GC_HD uint32_t gc_crc32(const GcDeflateTables& T, const uint8_t* p, uint32_t n)
{
	uint32_t c = 0xFFFFFFFFu;
	for (uint32_t i = 0; i < n; i++) c = T.crc[(c ^ p[i]) & 0xFF] ^ (c >> 8);
	return c ^ 0xFFFFFFFFu;
}

GC_HD uint32_t gc_crc_multmodp(uint32_t a, uint32_t b)
{
	uint32_t m = 1u << 31, p = 0;
	for (;;)
	{
		if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
		m >>= 1;
		b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
	}
	return p;
}
// CRC of A || B from the CRCs of A and B: crc(A) times x^(8 |B|), plus crc(B) (zlib's crc32_combine)
GC_HD uint32_t gc_crc_shift(const GcDeflateTables& T, uint32_t crc, uint64_t bytesAfter)
{
	uint32_t p = 1u << 31; uint32_t k = 3;
	while (bytesAfter) { if (bytesAfter & 1) p = gc_crc_multmodp(T.x2n[k & 31], p); bytesAfter >>= 1; k++; }
	return gc_crc_multmodp(p, crc);
}

// ---------------------------------------------------------------- one gzip member made by the 32 lanes of a warp
// The record is cut into 32 chunks.  Every lane parses its chunk (greedy LZ77, one hash probe per position, matches inside the chunk:
// what repeats in a record -- field tags, node name prefixes, edit shapes -- repeats at short range) and counts its symbols into the
// record's counts; lane 0 makes ONE dynamic-Huffman code for the record and the block header; the lanes' bit counts are scanned
// and every lane writes its chunk's codes where they belong in the member (32-bit words, the two words a lane shares with its
// neighbours by atomic OR); the CRCs of the chunks are joined (gc_crc_shift).  Written as per-lane stages: the kernel runs a stage
// on the 32 lanes and meets at a warp barrier, the host-side test double runs it for lane 0..31 in turn -- same bytes.
#define GC_GZ_LANES 32
#define GC_GZ_HASH_BITS 8
struct GcGzShared // per record: shared memory in the kernel
{
	uint16_t head[GC_GZ_LANES << GC_GZ_HASH_BITS];   // hash heads of the lanes' chunks: position in the chunk, 0xFFFF = none
	uint32_t litFreq[288], distFreq[32];
	uint16_t litCodes[288], distCodes[32];
	uint8_t litLens[288], distLens[32];
	uint8_t header[704]; uint32_t headerBits;        // gzip header, block header, code descriptions: lane 0 emits them before its chunk
	uint32_t laneTokens[GC_GZ_LANES], laneBits[GC_GZ_LANES], laneStart[GC_GZ_LANES], laneCrc[GC_GZ_LANES];
	uint32_t ok;
};
GC_HD uint32_t gc_gz_chunk(uint32_t n) { return (n + GC_GZ_LANES - 1) / GC_GZ_LANES; }
GC_HD void gc_gz_count(uint32_t* counter)
{
#if defined(__CUDA_ARCH__)
	atomicAdd(counter, 1u);
#else
	(*counter)++;
#endif
}
GC_HD void gc_gz_or(uint32_t* word, uint32_t bits)
{
#if defined(__CUDA_ARCH__)
	atomicOr(word, bits);
#else
	*word |= bits;
#endif
}
// stage 0 (every lane): clear the shared counts
GC_HD void gc_gz_clear(GcGzShared& sh, uint32_t lane)
{
	for (uint32_t i = lane; i < 288; i += GC_GZ_LANES) sh.litFreq[i] = 0;
	if (lane < 32) sh.distFreq[lane] = 0;
	if (lane == 0) sh.ok = 1;
}
// stage 1 (every lane): tokens of the lane's chunk (same token format as above), symbol counts, CRC of the chunk
GC_HD void gc_gz_tokenize(const GcDeflateTables& T, const uint8_t* p, uint32_t n, uint32_t lane, uint32_t* tokens, GcGzShared& sh)
{
	const uint32_t C = gc_gz_chunk(n);
	const uint32_t begin = lane * C < n ? lane * C : n, end = begin + C < n ? begin + C : n;
	uint16_t* head = sh.head + (lane << GC_GZ_HASH_BITS);
	for (uint32_t i = 0; i < (1u << GC_GZ_HASH_BITS); i++) head[i] = 0xFFFFu;
	uint32_t* out = tokens + begin;
	uint32_t nt = 0, i = begin;
	while (i + 4 <= end)
	{
		uint32_t v = gc_deflate_load32(p + i);
		uint32_t h = (v * 2654435761u) >> (32 - GC_GZ_HASH_BITS);
		uint32_t cand = head[h];
		head[h] = (uint16_t)(i - begin);
		if (cand != 0xFFFFu && gc_deflate_load32(p + begin + cand) == v)
		{
			const uint32_t c = begin + cand;
			uint32_t maxLen = end - i < 258 ? end - i : 258;
			uint32_t len = 4;
			while (len < maxLen && p[i + len] == p[c + len]) len++;
			uint32_t dist = i - c;
			uint32_t ds = (uint32_t)gc_deflate_dist_sym(T, dist);
			uint32_t sym = T.lenSym[len];
			out[nt++] = sym | (ds << 9) | ((len - T.lenBase[len]) << 14) | ((dist - T.distBase[ds]) << 19);
			gc_gz_count(&sh.litFreq[sym]); gc_gz_count(&sh.distFreq[ds]);
			if (i + len + 4 <= end)
			{
				uint32_t v2 = gc_deflate_load32(p + i + len - 1);
				head[(v2 * 2654435761u) >> (32 - GC_GZ_HASH_BITS)] = (uint16_t)(i + len - 1 - begin);
			}
			i += len;
		}
		else { out[nt++] = (uint32_t)p[i] | (30u << 9); gc_gz_count(&sh.litFreq[p[i]]); i++; }
	}
	for (; i < end; i++) { out[nt++] = (uint32_t)p[i] | (30u << 9); gc_gz_count(&sh.litFreq[p[i]]); }
	sh.laneTokens[lane] = nt;
	sh.laneCrc[lane] = gc_crc_shift(T, gc_crc32(T, p + begin, end - begin), n - end);
}
// stage 2 (lane 0): the record's codes and everything that precedes the first chunk's codes
GC_HD void gc_gz_header(GcGzShared& sh)
{
	const uint8_t gz[10] = { 0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 4, 3 }; // deflate, no flags, mtime 0, XFL fastest, OS unix
	for (int i = 0; i < 10; i++) sh.header[i] = gz[i];
	GcBitWriter bw; bw.p = sh.header + 10; bw.acc = 0; bw.n = 0;
	if (!gc_deflate_block_header(sh.litFreq, sh.distFreq, sh.litLens, sh.distLens, sh.litCodes, sh.distCodes, bw)) { sh.ok = 0; return; }
	sh.headerBits = (uint32_t)(bw.p - sh.header) * 8 + (uint32_t)bw.n;
	bw.flush();
}
// stage 3 (every lane): bits the lane will write
GC_HD void gc_gz_bits(const GcDeflateTables& T, uint32_t n, uint32_t lane, const uint32_t* tokens, GcGzShared& sh)
{
	const uint32_t C = gc_gz_chunk(n);
	const uint32_t begin = lane * C < n ? lane * C : n;
	const uint32_t* tk = tokens + begin;
	uint32_t bits = lane == 0 ? sh.headerBits : 0;
	for (uint32_t t = 0; t < sh.laneTokens[lane]; t++)
	{
		uint32_t sym = tk[t] & 0x1FF, ds = (tk[t] >> 9) & 31;
		bits += sh.litLens[sym] + T.symExtraBits[sym];
		if (ds != 30) bits += sh.distLens[ds] + T.distExtraBits[ds];
	}
	if (lane == GC_GZ_LANES - 1) bits += sh.litLens[256]; // end of block
	sh.laneBits[lane] = bits;
}
// 32-bit words of a lane's part of the bit stream: the first and the last word may hold a neighbour's bits as well
struct GcWordWriter
{
	uint32_t* w; uint64_t acc; int n; bool first;
	GC_HD void word(uint32_t v) { if (first) { gc_gz_or(w, v); first = false; } else *w = v; w++; }
	GC_HD void put(uint32_t bits, int count) // count <= 32
	{
		acc |= (uint64_t)bits << n;
		n += count;
		if (n >= 32) { word((uint32_t)acc); acc >>= 32; n -= 32; }
	}
	GC_HD void finish() { if (n > 0) gc_gz_or(w, (uint32_t)acc); }
};
// stage 4 (every lane, after laneStart = exclusive scan of laneBits): clear the two words the lane may share
GC_HD void gc_gz_prepare(uint32_t lane, const GcGzShared& sh, uint32_t* out32)
{
	const uint32_t s = sh.laneStart[lane], e = s + sh.laneBits[lane];
	if (e == s) return;
	out32[s >> 5] = 0;
	out32[(e - 1) >> 5] = 0;
}
// stage 5 (every lane): the lane's codes
GC_HD void gc_gz_emit(const GcDeflateTables& T, uint32_t n, uint32_t lane, const uint32_t* tokens, const GcGzShared& sh, uint32_t* out32)
{
	const uint32_t C = gc_gz_chunk(n);
	const uint32_t begin = lane * C < n ? lane * C : n;
	const uint32_t* tk = tokens + begin;
	const uint32_t s = sh.laneStart[lane];
	if (sh.laneBits[lane] == 0) return;
	GcWordWriter ww; ww.w = out32 + (s >> 5); ww.acc = 0; ww.n = (int)(s & 31); ww.first = true;
	if (lane == 0)
	{
		uint32_t full = sh.headerBits >> 3, rest = sh.headerBits & 7;
		for (uint32_t b = 0; b < full; b++) ww.put(sh.header[b], 8);
		if (rest) ww.put(sh.header[full] & ((1u << rest) - 1), (int)rest);
	}
	for (uint32_t t = 0; t < sh.laneTokens[lane]; t++)
	{
		uint32_t tok = tk[t];
		uint32_t sym = tok & 0x1FF, ds = (tok >> 9) & 31;
		int nb = sh.litLens[sym];
		ww.put((uint32_t)sh.litCodes[sym] | (((tok >> 14) & 31) << nb), nb + T.symExtraBits[sym]);
		if (ds != 30)
		{
			int db = sh.distLens[ds];
			ww.put((uint32_t)sh.distCodes[ds] | ((tok >> 19) << db), db + T.distExtraBits[ds]);
		}
	}
	if (lane == GC_GZ_LANES - 1) ww.put(sh.litCodes[256], sh.litLens[256]);
	ww.finish();
}
// stage 6 (lane 0): CRC-32 and length after the last byte of the block; returns the member's size
GC_HD uint32_t gc_gz_trailer(uint32_t n, const GcGzShared& sh, uint8_t* out)
{
	uint32_t totalBits = sh.laneStart[GC_GZ_LANES - 1] + sh.laneBits[GC_GZ_LANES - 1];
	uint32_t crc = 0;
	for (int l = 0; l < GC_GZ_LANES; l++) crc ^= sh.laneCrc[l];
	uint8_t* end = out + (totalBits + 7) / 8;
	for (int i = 0; i < 4; i++) *end++ = (uint8_t)(crc >> (8 * i));
	for (int i = 0; i < 4; i++) *end++ = (uint8_t)(n >> (8 * i));
	return (uint32_t)(end - out);
}
// capacity the member needs in the worst case (15 bits per literal + the code descriptions), and the largest record the chunk
// positions (16 bits) can address; a record outside either is encoded on the host
GC_HD bool gc_gz_fits(uint32_t n, uint32_t outCap) { return (uint64_t)outCap >= (uint64_t)n * 2 + 1024 && gc_gz_chunk(n) < 0xFFFFu; }

// the stages in turn for lanes 0..31 on one thread (host-side test double; same bytes as the kernel)
inline uint32_t gc_gzip_member(const GcDeflateTables& T, const uint8_t* raw, uint32_t n, GcDeflateWs& ws, uint8_t* out, uint32_t outCap)
{
	if (!gc_gz_fits(n, outCap)) return 0;
	GcGzShared* shp = new GcGzShared; GcGzShared& sh = *shp;
	uint32_t size = 0;
	for (uint32_t l = 0; l < GC_GZ_LANES; l++) gc_gz_clear(sh, l);
	for (uint32_t l = 0; l < GC_GZ_LANES; l++) gc_gz_tokenize(T, raw, n, l, ws.tokens, sh);
	gc_gz_header(sh);
	if (sh.ok)
	{
		for (uint32_t l = 0; l < GC_GZ_LANES; l++) gc_gz_bits(T, n, l, ws.tokens, sh);
		uint32_t at = 0;
		for (uint32_t l = 0; l < GC_GZ_LANES; l++) { sh.laneStart[l] = at; at += sh.laneBits[l]; }
		for (uint32_t l = 0; l < GC_GZ_LANES; l++) gc_gz_prepare(l, sh, (uint32_t*)out);
		for (uint32_t l = 0; l < GC_GZ_LANES; l++) gc_gz_emit(T, n, l, ws.tokens, sh, (uint32_t*)out);
		size = gc_gz_trailer(n, sh, out);
	}
	delete shp;
	return size;
}
