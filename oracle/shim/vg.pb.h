// TEST INFRASTRUCTURE ONLY (oracle build shim) -- not part of the product.
//
// Stand-in for the protoc-generated vg.pb.h, which cannot be generated in this
// image (no protoc / libprotobuf C++).  It lets the UNMODIFIED reference
// sources under /root/reference/src compile, and emits real proto3 wire bytes
// for the messages the alignment path serialises (vg.proto:15-126), so the
// oracle binary writes a genuine .gam.  Only the accessors the reference
// actually calls are provided.
#ifndef GC_ORACLE_SHIM_VG_PB_H
#define GC_ORACLE_SHIM_VG_PB_H

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <deque>
#include <map>
#include <mutex>
#include <climits>

#define GOOGLE_PROTOBUF_VERIFY_VERSION

namespace vgshim {

inline void putVarint(std::string& out, uint64_t v)
{
	while (v >= 0x80) { out.push_back((char)((v & 0x7F) | 0x80)); v >>= 7; }
	out.push_back((char)v);
}
inline void putTag(std::string& out, int field, int wire) { putVarint(out, ((uint64_t)field << 3) | (uint64_t)wire); }
inline void putInt(std::string& out, int field, int64_t v) { if (v == 0) return; putTag(out, field, 0); putVarint(out, (uint64_t)v); }
inline void putBool(std::string& out, int field, bool v) { if (!v) return; putTag(out, field, 0); putVarint(out, 1); }
inline void putStr(std::string& out, int field, const std::string& s) { if (s.empty()) return; putTag(out, field, 2); putVarint(out, s.size()); out += s; }
inline void putMsg(std::string& out, int field, const std::string& s) { putTag(out, field, 2); putVarint(out, s.size()); out += s; }
inline void putDouble(std::string& out, int field, double d)
{
	uint64_t bits; std::memcpy(&bits, &d, 8);
	if (bits == 0) return;
	putTag(out, field, 1);
	for (int i = 0; i < 8; i++) out.push_back((char)((bits >> (8 * i)) & 0xFF));
}

struct Reader
{
	const unsigned char* p; const unsigned char* end; bool ok = true;
	Reader(const std::string& s) : p((const unsigned char*)s.data()), end((const unsigned char*)s.data() + s.size()) {}
	bool done() const { return p >= end || !ok; }
	uint64_t varint()
	{
		uint64_t r = 0; int shift = 0;
		while (p < end) { unsigned char c = *p++; r |= (uint64_t)(c & 0x7F) << shift; if (!(c & 0x80)) return r; shift += 7; if (shift > 63) break; }
		ok = false; return 0;
	}
	std::string bytes()
	{
		uint64_t n = varint();
		if (!ok || (uint64_t)(end - p) < n) { ok = false; return ""; }
		std::string s((const char*)p, n); p += n; return s;
	}
	double fixed64()
	{
		if (end - p < 8) { ok = false; return 0; }
		uint64_t bits = 0; for (int i = 0; i < 8; i++) bits |= (uint64_t)p[i] << (8 * i);
		p += 8; double d; std::memcpy(&d, &bits, 8); return d;
	}
	void skip(int wire)
	{
		switch (wire) { case 0: varint(); break; case 1: if (end - p < 8) ok = false; else p += 8; break; case 2: bytes(); break; case 5: if (end - p < 4) ok = false; else p += 4; break; default: ok = false; }
	}
};

}

namespace vg {

class Edit
{
public:
	int32_t from_length() const { return from_length_; }
	int32_t to_length() const { return to_length_; }
	const std::string& sequence() const { return sequence_; }
	void set_from_length(int32_t v) { from_length_ = v; }
	void set_to_length(int32_t v) { to_length_ = v; }
	void set_sequence(const std::string& s) { sequence_ = s; }
	void Encode(std::string& out) const
	{
		vgshim::putInt(out, 1, from_length_); vgshim::putInt(out, 2, to_length_); vgshim::putStr(out, 3, sequence_);
	}
	bool Decode(const std::string& s)
	{
		vgshim::Reader r(s);
		while (!r.done()) { uint64_t t = r.varint(); int f = (int)(t >> 3), w = (int)(t & 7);
			if (f == 1 && w == 0) from_length_ = (int32_t)r.varint(); else if (f == 2 && w == 0) to_length_ = (int32_t)r.varint(); else if (f == 3 && w == 2) sequence_ = r.bytes(); else r.skip(w); }
		return r.ok;
	}
private:
	int32_t from_length_ = 0; int32_t to_length_ = 0; std::string sequence_;
};

class Position
{
public:
	int64_t node_id() const { return node_id_; }
	int64_t offset() const { return offset_; }
	bool is_reverse() const { return is_reverse_; }
	const std::string& name() const { return name_; }
	void set_node_id(int64_t v) { node_id_ = v; }
	void set_offset(int64_t v) { offset_ = v; }
	void set_is_reverse(bool v) { is_reverse_ = v; }
	void set_name(const std::string& s) { name_ = s; }
	void Encode(std::string& out) const
	{
		vgshim::putInt(out, 1, node_id_); vgshim::putInt(out, 2, offset_); vgshim::putBool(out, 4, is_reverse_); vgshim::putStr(out, 5, name_);
	}
	bool Decode(const std::string& s)
	{
		vgshim::Reader r(s);
		while (!r.done()) { uint64_t t = r.varint(); int f = (int)(t >> 3), w = (int)(t & 7);
			if (f == 1 && w == 0) node_id_ = (int64_t)r.varint(); else if (f == 2 && w == 0) offset_ = (int64_t)r.varint(); else if (f == 4 && w == 0) is_reverse_ = r.varint() != 0; else if (f == 5 && w == 2) name_ = r.bytes(); else r.skip(w); }
		return r.ok;
	}
private:
	int64_t node_id_ = 0; int64_t offset_ = 0; bool is_reverse_ = false; std::string name_;
};

class Mapping
{
public:
	Mapping() {}
	Mapping(const Mapping& o) : edits_(o.edits_), rank_(o.rank_) { if (o.position_) position_.reset(new Position(*o.position_)); }
	Mapping& operator=(const Mapping& o) { if (this != &o) { edits_ = o.edits_; rank_ = o.rank_; position_.reset(o.position_ ? new Position(*o.position_) : nullptr); } return *this; }
	const Position& position() const { static const Position empty; return position_ ? *position_ : empty; }
	Position* mutable_position() { if (!position_) position_.reset(new Position); return position_.get(); }
	void set_allocated_position(Position* p) { position_.reset(p); }
	bool has_position() const { return (bool)position_; }
	Edit* add_edit() { edits_.emplace_back(); return &edits_.back(); }
	const Edit& edit(int i) const { return edits_[i]; }
	Edit* mutable_edit(int i) { return &edits_[i]; }
	int edit_size() const { return (int)edits_.size(); }
	int64_t rank() const { return rank_; }
	void set_rank(int64_t r) { rank_ = r; }
	void Encode(std::string& out) const
	{
		if (position_) { std::string s; position_->Encode(s); vgshim::putMsg(out, 1, s); }
		for (const auto& e : edits_) { std::string s; e.Encode(s); vgshim::putMsg(out, 2, s); }
		vgshim::putInt(out, 5, rank_);
	}
	bool Decode(const std::string& s)
	{
		vgshim::Reader r(s);
		while (!r.done()) { uint64_t t = r.varint(); int f = (int)(t >> 3), w = (int)(t & 7);
			if (f == 1 && w == 2) { if (!mutable_position()->Decode(r.bytes())) return false; }
			else if (f == 2 && w == 2) { if (!add_edit()->Decode(r.bytes())) return false; }
			else if (f == 5 && w == 0) rank_ = (int64_t)r.varint(); else r.skip(w); }
		return r.ok;
	}
private:
	std::unique_ptr<Position> position_;
	std::deque<Edit> edits_; // deque: pointers returned by add_edit() stay valid
	int64_t rank_ = 0;
};

class Path
{
public:
	Path() {}
	Path(const Path& o) : name_(o.name_) { for (const auto& m : o.mappings_) mappings_.emplace_back(new Mapping(*m)); }
	Path& operator=(const Path& o) { if (this != &o) { name_ = o.name_; mappings_.clear(); for (const auto& m : o.mappings_) mappings_.emplace_back(new Mapping(*m)); } return *this; }
	Mapping* add_mapping() { mappings_.emplace_back(new Mapping); return mappings_.back().get(); }
	const Mapping& mapping(int i) const { return *mappings_[i]; }
	Mapping* mutable_mapping(int i) { return mappings_[i].get(); }
	int mapping_size() const { return (int)mappings_.size(); }
	const std::string& name() const { return name_; }
	void set_name(const std::string& s) { name_ = s; }
	void Encode(std::string& out) const
	{
		vgshim::putStr(out, 1, name_);
		for (const auto& m : mappings_) { std::string s; m->Encode(s); vgshim::putMsg(out, 2, s); }
	}
	bool Decode(const std::string& s)
	{
		vgshim::Reader r(s);
		while (!r.done()) { uint64_t t = r.varint(); int f = (int)(t >> 3), w = (int)(t & 7);
			if (f == 1 && w == 2) name_ = r.bytes(); else if (f == 2 && w == 2) { if (!add_mapping()->Decode(r.bytes())) return false; } else r.skip(w); }
		return r.ok;
	}
private:
	std::string name_;
	std::vector<std::unique_ptr<Mapping>> mappings_;
};

class Alignment
{
public:
	Alignment() {}
	Alignment(const Alignment& o) : sequence_(o.sequence_), name_(o.name_), score_(o.score_), query_position_(o.query_position_), identity_(o.identity_) { if (o.path_) path_.reset(new Path(*o.path_)); }
	Alignment& operator=(const Alignment& o) { if (this != &o) { sequence_ = o.sequence_; name_ = o.name_; score_ = o.score_; query_position_ = o.query_position_; identity_ = o.identity_; path_.reset(o.path_ ? new Path(*o.path_) : nullptr); } return *this; }
	const std::string& sequence() const { return sequence_; }
	void set_sequence(const std::string& s) { sequence_ = s; }
	const std::string& name() const { return name_; }
	void set_name(const std::string& s) { name_ = s; }
	int32_t score() const { return score_; }
	void set_score(int32_t s) { score_ = s; }
	int32_t query_position() const { return query_position_; }
	void set_query_position(int32_t q) { query_position_ = q; }
	double identity() const { return identity_; }
	void set_identity(double d) { identity_ = d; }
	const Path& path() const { static const Path empty; return path_ ? *path_ : empty; }
	Path* mutable_path() { if (!path_) path_.reset(new Path); return path_.get(); }
	void set_allocated_path(Path* p) { path_.reset(p); }
	bool has_path() const { return (bool)path_; }
	bool SerializeToString(std::string* out) const
	{
		out->clear();
		vgshim::putStr(*out, 1, sequence_);
		if (path_) { std::string s; path_->Encode(s); vgshim::putMsg(*out, 2, s); }
		vgshim::putStr(*out, 3, name_);
		vgshim::putInt(*out, 6, score_);
		vgshim::putInt(*out, 7, query_position_);
		vgshim::putDouble(*out, 16, identity_);
		return true;
	}
	bool ParseFromString(const std::string& s)
	{
		*this = Alignment();
		vgshim::Reader r(s);
		while (!r.done()) { uint64_t t = r.varint(); int f = (int)(t >> 3), w = (int)(t & 7);
			if (f == 1 && w == 2) sequence_ = r.bytes();
			else if (f == 2 && w == 2) { if (!mutable_path()->Decode(r.bytes())) return false; }
			else if (f == 3 && w == 2) name_ = r.bytes();
			else if (f == 6 && w == 0) score_ = (int32_t)r.varint();
			else if (f == 7 && w == 0) query_position_ = (int32_t)r.varint();
			else if (f == 16 && w == 1) identity_ = r.fixed64();
			else r.skip(w); }
		return r.ok;
	}
private:
	std::string sequence_; std::string name_; int32_t score_ = 0; int32_t query_position_ = 0; double identity_ = 0;
	std::unique_ptr<Path> path_;
};

class Node
{
public:
	int64_t id() const { return id_; }
	const std::string& sequence() const { return sequence_; }
	const std::string& name() const { return name_; }
	void set_id(int64_t v) { id_ = v; }
	void set_sequence(const std::string& s) { sequence_ = s; }
	void set_name(const std::string& s) { name_ = s; }
	bool Decode(const std::string& s)
	{
		vgshim::Reader r(s);
		while (!r.done()) { uint64_t t = r.varint(); int f = (int)(t >> 3), w = (int)(t & 7);
			if (f == 1 && w == 2) sequence_ = r.bytes(); else if (f == 2 && w == 2) name_ = r.bytes(); else if (f == 3 && w == 0) id_ = (int64_t)r.varint(); else r.skip(w); }
		return r.ok;
	}
private:
	int64_t id_ = 0; std::string sequence_; std::string name_;
};

class Edge
{
public:
	int64_t from() const { return from_; }
	int64_t to() const { return to_; }
	bool from_start() const { return from_start_; }
	bool to_end() const { return to_end_; }
	int32_t overlap() const { return overlap_; }
	void set_from(int64_t v) { from_ = v; }
	void set_to(int64_t v) { to_ = v; }
	void set_from_start(bool v) { from_start_ = v; }
	void set_to_end(bool v) { to_end_ = v; }
	void set_overlap(int32_t v) { overlap_ = v; }
	bool Decode(const std::string& s)
	{
		vgshim::Reader r(s);
		while (!r.done()) { uint64_t t = r.varint(); int f = (int)(t >> 3), w = (int)(t & 7);
			if (f == 1 && w == 0) from_ = (int64_t)r.varint(); else if (f == 2 && w == 0) to_ = (int64_t)r.varint(); else if (f == 3 && w == 0) from_start_ = r.varint() != 0; else if (f == 4 && w == 0) to_end_ = r.varint() != 0; else if (f == 5 && w == 0) overlap_ = (int32_t)r.varint(); else r.skip(w); }
		return r.ok;
	}
private:
	int64_t from_ = 0; int64_t to_ = 0; bool from_start_ = false; bool to_end_ = false; int32_t overlap_ = 0;
};

class Graph
{
public:
	int node_size() const { return (int)nodes_.size(); }
	const Node& node(int i) const { return nodes_[i]; }
	Node* add_node() { nodes_.emplace_back(); return &nodes_.back(); }
	int edge_size() const { return (int)edges_.size(); }
	const Edge& edge(int i) const { return edges_[i]; }
	Edge* add_edge() { edges_.emplace_back(); return &edges_.back(); }
	bool SerializeToString(std::string* out) const { out->clear(); return true; }
	bool ParseFromString(const std::string& s)
	{
		*this = Graph();
		vgshim::Reader r(s);
		while (!r.done()) { uint64_t t = r.varint(); int f = (int)(t >> 3), w = (int)(t & 7);
			if (f == 1 && w == 2) { if (!add_node()->Decode(r.bytes())) return false; } else if (f == 2 && w == 2) { if (!add_edge()->Decode(r.bytes())) return false; } else r.skip(w); }
		return r.ok;
	}
private:
	std::vector<Node> nodes_; std::vector<Edge> edges_;
};

}

#endif
