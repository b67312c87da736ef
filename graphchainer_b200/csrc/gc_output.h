// S7: final traces -> vg::Alignment messages -> GAM / JSON records.
//   GraphAlignerVGAlignment::traceToAlignment        (src/GraphAlignerVGAlignment.h:37-165; the edit runs come from gc_post.cuh)
//   GraphAligner::AddAlignment                       (src/GraphAligner.h:205-212)
//   replaceDigraphNodeIdsWithOriginalNodeIds         (src/Aligner.cpp:152-165)
//   writeGAMToQueue / writeJSONToQueue               (src/Aligner.cpp:261-298)
// protobuf is not available in this image, so the proto3 wire format of the few messages
// involved (src/vg.proto:52-126) is written by hand; GAM framing = one gzip member per read
// holding varint64 count + {varint32 size, message}* (src/stream.hpp:24-51).
#pragma once
#include <zlib.h>
#include "gc_deflate.h"
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "gc_pipeline.h"

namespace gcout {

struct Edit { int32_t from_length = 0, to_length = 0; std::string sequence; };
struct Mapping { int64_t node_id = 0, offset = 0; bool is_reverse = false; std::string name; std::vector<Edit> edits; int64_t rank = 0; };
struct Alignment { std::string sequence, name; std::vector<Mapping> mappings; int32_t score = 0, query_position = 0; double identity = 0; };

// The final alignments arrive as token streams (gc_post.cuh): mapping = {0, digraph node id, offset in the original node},
// edit = type << 30 | run length.  The characters of mismatch and insertion runs are consecutive read characters starting
// at alignmentStart; the very first trace entry, when it is a mismatch, takes sequence[0] instead (sic, GraphAlignerVGAlignment.h:75).
struct TokenReader
{
	const uint32_t* t; size_t n, i = 0;
	size_t nextChar;      // read position of the next character an edit consumes
	bool firstEdit = true;
	TokenReader(const GcAlnItem& item) : t(item.tokens.data()), n(item.tokens.size()), nextChar(item.alignmentStart) {}
	bool atMapping() const { return i < n && (t[i] & 0x3FFFFFFFu) == 0; }
	void mapping(int& digraphNode, size_t& offset) { digraphNode = (int)t[i + 1]; offset = t[i + 2]; i += 3; }
	bool atEdit() const { return i < n && (t[i] & 0x3FFFFFFFu) != 0; }
	// from/to lengths and the edit's sequence (appended to `seq`)
	void edit(const std::string& sequence, int32_t& from, int32_t& to, std::string& seq)
	{
		uint32_t type = t[i] >> 30, len = t[i] & 0x3FFFFFFFu;
		i++;
		from = (type == GC_EDIT_INSERTION) ? 0 : (int32_t)len;
		to = (type == GC_EDIT_DELETION) ? 0 : (int32_t)len;
		if (type == GC_EDIT_MISMATCH || type == GC_EDIT_INSERTION)
		{
			seq.append(sequence, nextChar, len);
			if (firstEdit && type == GC_EDIT_MISMATCH) seq[seq.size() - len] = sequence[0];
		}
		nextChar += (size_t)to;
		firstEdit = false;
	}
};

// traceToAlignment + AddAlignment + replaceDigraphNodeIdsWithOriginalNodeIds
inline Alignment toAlignment(const GcHostGraph& g, const std::string& seq_id, const std::string& sequence, const GcAlnItem& item)
{
	Alignment result;
	result.name = seq_id;
	result.score = item.traceScore;
	TokenReader rd(item);
	int rank = 0;
	while (rd.atMapping())
	{
		int digraphNode; size_t offset;
		rd.mapping(digraphNode, offset);
		result.mappings.emplace_back();
		Mapping& m = result.mappings.back();
		m.rank = rank++;
		m.offset = (int64_t)offset;
		m.is_reverse = (digraphNode % 2) == 1;
		// replaceDigraphNodeIdsWithOriginalNodeIds (Aligner.cpp:152-165)
		m.node_id = digraphNode / 2;
		m.name = g.originalNodeName(digraphNode);
		while (rd.atEdit())
		{
			m.edits.emplace_back();
			Edit& e = m.edits.back();
			rd.edit(sequence, e.from_length, e.to_length, e.sequence);
		}
	}
	result.identity = (double)item.matches / (double)item.steps;
	// AddAlignment (GraphAligner.h:210-211)
	result.sequence = sequence.substr(item.alignmentStart, item.alignmentEnd - item.alignmentStart);
	result.query_position = (int32_t)item.alignmentStart;
	return result;
}

// ---- proto3 wire encoding (vg.proto:52-126)
inline void putVarint(std::string& out, uint64_t v) { while (v >= 0x80) { out.push_back((char)((v & 0x7F) | 0x80)); v >>= 7; } out.push_back((char)v); }
inline void putTag(std::string& out, int field, int wire) { putVarint(out, ((uint64_t)field << 3) | (uint64_t)wire); }
inline void putInt(std::string& out, int field, int64_t v) { if (v == 0) return; putTag(out, field, 0); putVarint(out, (uint64_t)v); }
inline void putStr(std::string& out, int field, const std::string& s) { if (s.empty()) return; putTag(out, field, 2); putVarint(out, s.size()); out += s; }
inline void putMsg(std::string& out, int field, const std::string& s) { putTag(out, field, 2); putVarint(out, s.size()); out += s; }

inline std::string serialize(const Alignment& a)
{
	std::string path;
	for (const Mapping& m : a.mappings)
	{
		std::string mm, pos;
		putInt(pos, 1, m.node_id); putInt(pos, 2, m.offset); if (m.is_reverse) { putTag(pos, 4, 0); putVarint(pos, 1); } putStr(pos, 5, m.name);
		putMsg(mm, 1, pos);
		for (const Edit& e : m.edits) { std::string ee; putInt(ee, 1, e.from_length); putInt(ee, 2, e.to_length); putStr(ee, 3, e.sequence); putMsg(mm, 2, ee); }
		putInt(mm, 5, m.rank);
		putMsg(path, 2, mm);
	}
	std::string out;
	putStr(out, 1, a.sequence);
	putMsg(out, 2, path);
	putStr(out, 3, a.name);
	putInt(out, 6, a.score);
	putInt(out, 7, a.query_position);
	uint64_t bits; std::memcpy(&bits, &a.identity, 8);
	if (bits != 0) { putTag(out, 16, 1); for (int i = 0; i < 8; i++) out.push_back((char)((bits >> (8 * i)) & 0xFF)); }
	return out;
}

inline std::string gzipMember(const std::string& raw)
{
	z_stream zs; std::memset(&zs, 0, sizeof(zs));
	deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY);
	std::string out;
	out.resize(deflateBound(&zs, raw.size()) + 32);
	zs.next_in = (Bytef*)raw.data(); zs.avail_in = (uInt)raw.size();
	zs.next_out = (Bytef*)&out[0]; zs.avail_out = (uInt)out.size();
	deflate(&zs, Z_FINISH);
	out.resize(out.size() - zs.avail_out);
	deflateEnd(&zs);
	return out;
}

// one read's GAM record (writeGAMToQueue, Aligner.cpp:261-281)
inline std::string gamRecord(const std::vector<Alignment>& alns)
{
	std::string raw;
	putVarint(raw, alns.size());
	for (const Alignment& a : alns) { std::string s = serialize(a); putVarint(raw, s.size()); raw += s; }
	return gzipMember(raw);
}


// ---- direct GAM encoder: the same bytes as serialize(toAlignment(...)) without building the
// message objects (one pass over the trace, two reusable buffers).  tests/ check both paths
// against the reference's records.
struct GamEncoder
{
	std::string mm, path, ee, msg;
	static void putVarintTo(std::string& out, uint64_t v) { while (v >= 0x80) { out.push_back((char)((v & 0x7F) | 0x80)); v >>= 7; } out.push_back((char)v); }
	void flushEdit(int32_t from, int32_t to, const std::string& seq)
	{
		ee.clear();
		if (from) { ee.push_back((char)0x08); putVarintTo(ee, (uint64_t)(int64_t)from); }
		if (to) { ee.push_back((char)0x10); putVarintTo(ee, (uint64_t)(int64_t)to); }
		if (!seq.empty()) { ee.push_back((char)0x1A); putVarintTo(ee, seq.size()); ee += seq; }
		mm.push_back((char)0x12); putVarintTo(mm, ee.size()); mm += ee;
	}
	void beginMapping(const GcHostGraph& g, int digraphNode, size_t offset)
	{
		mm.clear();
		ee.clear(); // position message built in ee
		int64_t nodeId = digraphNode / 2;
		if (nodeId) { ee.push_back((char)0x08); putVarintTo(ee, (uint64_t)nodeId); }
		if (offset) { ee.push_back((char)0x10); putVarintTo(ee, (uint64_t)offset); }
		if (digraphNode % 2 == 1) { ee.push_back((char)0x20); ee.push_back((char)1); }
		const std::string& name = g.originalNodeName(digraphNode);
		if (!name.empty()) { ee.push_back((char)0x2A); putVarintTo(ee, name.size()); ee += name; }
		mm.push_back((char)0x0A); putVarintTo(mm, ee.size()); mm += ee;
	}
	void endMapping(int rank)
	{
		if (rank) { mm.push_back((char)0x28); putVarintTo(mm, (uint64_t)rank); }
		path.push_back((char)0x12); putVarintTo(path, mm.size()); path += mm;
	}
	// appends varint32 size + message of one alignment to `out`
	void encode(const GcHostGraph& g, const std::string& seq_id, const std::string& sequence, const GcAlnItem& item, std::string& out)
	{
		path.clear();
		TokenReader rd(item);
		int rank = 0;
		std::string eseq;
		while (rd.atMapping())
		{
			int digraphNode; size_t offset;
			rd.mapping(digraphNode, offset);
			beginMapping(g, digraphNode, offset);
			while (rd.atEdit())
			{
				int32_t from, to;
				eseq.clear();
				rd.edit(sequence, from, to, eseq);
				flushEdit(from, to, eseq);
			}
			endMapping(rank++);
		}
		double identity = (double)item.matches / (double)item.steps;
		msg.clear();
		size_t alnLen = item.alignmentEnd - item.alignmentStart;
		if (alnLen) { msg.push_back((char)0x0A); putVarintTo(msg, alnLen); msg.append(sequence, item.alignmentStart, alnLen); }
		msg.push_back((char)0x12); putVarintTo(msg, path.size()); msg += path;
		if (!seq_id.empty()) { msg.push_back((char)0x1A); putVarintTo(msg, seq_id.size()); msg += seq_id; }
		if (item.traceScore) { msg.push_back((char)0x30); putVarintTo(msg, (uint64_t)(int64_t)item.traceScore); }
		if ((int32_t)item.alignmentStart) { msg.push_back((char)0x38); putVarintTo(msg, (uint64_t)(int64_t)(int32_t)item.alignmentStart); }
		uint64_t bits; std::memcpy(&bits, &identity, 8);
		if (bits != 0) { msg.push_back((char)0x81); msg.push_back((char)0x01); for (int i = 0; i < 8; i++) msg.push_back((char)((bits >> (8 * i)) & 0xFF)); }
		putVarintTo(out, msg.size());
		out += msg;
	}
};

// one gzip member per record.  Level 1 = the driver's own single-pass encoder (gc_deflate.h, ~2x zlib level 1 at the
// same ratio on GAM records); other levels = zlib, its deflate state (256 KB of tables) kept per thread and only reset
inline std::string gzipMemberLevel(const std::string& raw, int level)
{
	if (level == 1)
	{
		static thread_local gcdeflate::Encoder enc;
		std::string out = enc.gzipMember(raw);
		if (!out.empty()) return out;
	}
	struct State { z_stream zs; int level = -100; bool live = false; ~State() { if (live) deflateEnd(&zs); } };
	static thread_local State st;
	if (!st.live || st.level != level)
	{
		if (st.live) deflateEnd(&st.zs);
		std::memset(&st.zs, 0, sizeof(st.zs));
		deflateInit2(&st.zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY);
		st.level = level; st.live = true;
	}
	else deflateReset(&st.zs);
	z_stream& zs = st.zs;
	std::string out;
	out.resize(deflateBound(&zs, raw.size()) + 32);
	zs.next_in = (Bytef*)raw.data(); zs.avail_in = (uInt)raw.size();
	zs.next_out = (Bytef*)&out[0]; zs.avail_out = (uInt)out.size();
	deflate(&zs, Z_FINISH);
	out.resize(out.size() - zs.avail_out);
	return out;
}

// one read's GAM record straight from the final alignments (writeGAMToQueue, Aligner.cpp:261-281).
// `level` = zlib level of the gzip member; the reference's GzipOutputStream uses the zlib default (6),
// decoded records are identical at any level.
inline std::string gamRecordDirect(const GcHostGraph& g, const std::string& seq_id, const std::string& sequence, const std::vector<GcAlnItem>& alns, int level, GamEncoder& enc)
{
	std::string raw;
	raw.reserve(sequence.size() * 2 + 256);
	putVarint(raw, alns.size());
	{ GC_PROF_SCOPE(13, "gam.encode"); for (const GcAlnItem& a : alns) enc.encode(g, seq_id, sequence, a, raw); }
	GC_PROF_SCOPE(14, "gam.deflate");
	return gzipMemberLevel(raw, level);
}

// ---- GAF line of one alignment (GraphAlignerGAFAlignment::traceToAlignment, src/GraphAlignerGAFAlignment.h:37-205), from the
// same token stream as the GAM record: a token "mapping" is exactly a node of the GAF path (same inside-the-node rule,
// :103 vs GraphAlignerVGAlignment.h), the edit runs are the CIGAR once the runs of equal type on both sides of a mapping
// boundary are joined (the GAF writer does not flush its run at a node change).  Offsets: inside a mapping every
// match / mismatch / deletion step advances one base of the original node, an insertion none.
inline std::string gafLine(const GcHostGraph& g, const std::string& seq_id, const std::string& sequence, const GcAlnItem& item, bool cigarMatchMismatchMerge)
{
	std::string nodePath, cigar;
	size_t nodePathLen = 0, nodePathStart = 0;
	size_t counts[4] = { 0, 0, 0, 0 };          // by GC_EDIT_* type
	uint32_t runType = 4; size_t runLen = 0;      // CIGAR run in progress (4 = none)
	auto flushRun = [&]()
	{
		if (runLen == 0) return;
		cigar += std::to_string(runLen);
		cigar += runType == GC_EDIT_INSERTION ? 'I' : runType == GC_EDIT_DELETION ? 'D' : cigarMatchMismatchMerge ? 'M' : runType == GC_EDIT_MATCH ? '=' : 'X';
	};
	TokenReader rd(item);
	bool first = true;
	int prevNode = 0; size_t prevLastOffset = 0;
	while (rd.atMapping())
	{
		int digraphNode; size_t offset;
		rd.mapping(digraphNode, offset);
		nodePath += (digraphNode % 2) == 1 ? '<' : '>';
		const std::string& name = g.originalNodeName(digraphNode);
		nodePath += name.empty() ? std::to_string(digraphNode / 2) : name;
		size_t size = g.origSize[g.origIndexOfId[digraphNode]];
		if (first) { nodePathLen += size; nodePathStart = offset; first = false; }
		else
		{
			size_t skippedBefore = g.origSize[g.origIndexOfId[prevNode]] - 1 - prevLastOffset, skippedAfter = offset;
			nodePathLen += size - (skippedBefore + skippedAfter);
		}
		size_t advancing = 0; // steps of this mapping that move along the node
		while (rd.atEdit())
		{
			uint32_t type = rd.t[rd.i] >> 30, len = rd.t[rd.i] & 0x3FFFFFFFu;
			rd.i++;
			counts[type] += len;
			if (type != GC_EDIT_INSERTION) advancing += len;
			uint32_t cg = (cigarMatchMismatchMerge && type == GC_EDIT_MISMATCH) ? (uint32_t)GC_EDIT_MATCH : type;
			if (cg != runType) { flushRun(); runType = cg; runLen = 0; }
			runLen += len;
		}
		prevNode = digraphNode;
		prevLastOffset = offset + advancing - 1;
	}
	flushRun();
	const size_t matches = counts[GC_EDIT_MATCH], all = counts[0] + counts[1] + counts[2] + counts[3];
	size_t nodePathEnd = nodePathLen - (g.origSize[g.origIndexOfId[prevNode]] - 1 - prevLastOffset);
	auto dbl = [](double v) { char b[64]; snprintf(b, sizeof(b), "%g", v); return std::string(b); }; // operator<<(double): precision 6, %g
	std::string line = seq_id + "\t" + std::to_string(sequence.size()) + "\t" + std::to_string(item.alignmentStart) + "\t" + std::to_string(item.alignmentEnd) + "\t+\t" + nodePath + "\t"
		+ std::to_string(nodePathLen) + "\t" + std::to_string(nodePathStart) + "\t" + std::to_string(nodePathEnd) + "\t" + std::to_string(matches) + "\t" + std::to_string(all) + "\t255";
	line += "\tNM:i:" + std::to_string(all - matches);
	line += "\tdv:f:" + dbl(1.0 - ((double)matches / (double)all));
	line += "\tid:f:" + dbl((double)matches / (double)all);
	line += "\tcg:Z:" + cigar;
	return line;
}

inline std::string jsonEscape(const std::string& s)
{
	std::string r = "\"";
	for (unsigned char c : s)
	{
		if (c == '"') r += "\\\""; else if (c == '\\') r += "\\\\"; else if (c == '\n') r += "\\n"; else if (c == '\t') r += "\\t"; else if (c == '\r') r += "\\r";
		else if (c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); r += b; } else r.push_back((char)c);
	}
	return r + "\"";
}
inline std::string jsonDouble(double d)
{
	char b[64];
	for (int prec = 15; prec <= 17; prec++) { snprintf(b, sizeof(b), "%.*g", prec, d); if (strtod(b, nullptr) == d) break; }
	return b;
}
// MessageToJsonString with preserve_proto_field_names (Aligner.cpp:283-298): one line per alignment
inline std::string jsonLine(const Alignment& a)
{
	std::string o = "{";
	bool first = true;
	auto sep = [&o](bool& f) { if (!f) o += ","; f = false; };
	if (!a.sequence.empty()) { sep(first); o += "\"sequence\":" + jsonEscape(a.sequence); }
	sep(first); o += "\"path\":{";
	if (!a.mappings.empty())
	{
		o += "\"mapping\":[";
		for (size_t i = 0; i < a.mappings.size(); i++)
		{
			if (i) o += ",";
			const Mapping& m = a.mappings[i];
			o += "{\"position\":{"; bool qf = true;
			if (m.node_id) { sep(qf); o += "\"node_id\":\"" + std::to_string(m.node_id) + "\""; }
			if (m.offset) { sep(qf); o += "\"offset\":\"" + std::to_string(m.offset) + "\""; }
			if (m.is_reverse) { sep(qf); o += "\"is_reverse\":true"; }
			if (!m.name.empty()) { sep(qf); o += "\"name\":" + jsonEscape(m.name); }
			o += "}";
			if (!m.edits.empty())
			{
				o += ",\"edit\":[";
				for (size_t e = 0; e < m.edits.size(); e++)
				{
					if (e) o += ",";
					o += "{"; bool ef = true;
					if (m.edits[e].from_length) { sep(ef); o += "\"from_length\":" + std::to_string(m.edits[e].from_length); }
					if (m.edits[e].to_length) { sep(ef); o += "\"to_length\":" + std::to_string(m.edits[e].to_length); }
					if (!m.edits[e].sequence.empty()) { sep(ef); o += "\"sequence\":" + jsonEscape(m.edits[e].sequence); }
					o += "}";
				}
				o += "]";
			}
			if (m.rank) o += ",\"rank\":\"" + std::to_string(m.rank) + "\"";
			o += "}";
		}
		o += "]";
	}
	o += "}";
	if (!a.name.empty()) { sep(first); o += "\"name\":" + jsonEscape(a.name); }
	if (a.score) { sep(first); o += "\"score\":" + std::to_string(a.score); }
	if (a.query_position) { sep(first); o += "\"query_position\":" + std::to_string(a.query_position); }
	if (a.identity != 0) { sep(first); o += "\"identity\":" + jsonDouble(a.identity); }
	o += "}";
	return o;
}

}
