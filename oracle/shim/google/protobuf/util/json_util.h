// TEST INFRASTRUCTURE ONLY (oracle build shim) -- see oracle/shim/vg.pb.h.
// proto3-JSON printer for vg::Alignment with the conventions of
// MessageToJsonString(preserve_proto_field_names): 64-bit integers as quoted
// strings, zero/empty fields omitted, doubles in shortest round-trip form.
#ifndef GC_ORACLE_SHIM_PB_JSON_H
#define GC_ORACLE_SHIM_PB_JSON_H
#include <string>
#include <cstdio>
#include <cstdlib>
#include "vg.pb.h"
namespace google { namespace protobuf { namespace util {
struct JsonPrintOptions { bool preserve_proto_field_names = false; bool add_whitespace = false; bool always_print_primitive_fields = false; };
inline std::string gcJsonEscape(const std::string& s)
{
	std::string r = "\"";
	for (unsigned char c : s)
	{
		if (c == '"') r += "\\\""; else if (c == '\\') r += "\\\\"; else if (c == '\n') r += "\\n"; else if (c == '\t') r += "\\t"; else if (c == '\r') r += "\\r";
		else if (c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); r += b; } else r.push_back((char)c);
	}
	r += "\""; return r;
}
inline std::string gcJsonDouble(double d)
{
	char b[64];
	for (int prec = 15; prec <= 17; prec++) { snprintf(b, sizeof(b), "%.*g", prec, d); if (strtod(b, nullptr) == d) break; }
	return b;
}
inline void MessageToJsonString(const vg::Alignment& a, std::string* out, const JsonPrintOptions&)
{
	std::string& o = *out; o = "{"; bool first = true;
	auto sep = [&o](bool& f) { if (!f) o += ","; f = false; };
	if (!a.sequence().empty()) { sep(first); o += "\"sequence\":" + gcJsonEscape(a.sequence()); }
	if (a.has_path())
	{
		sep(first); o += "\"path\":{"; bool pf = true;
		if (!a.path().name().empty()) { sep(pf); o += "\"name\":" + gcJsonEscape(a.path().name()); }
		if (a.path().mapping_size() > 0)
		{
			sep(pf); o += "\"mapping\":[";
			for (int i = 0; i < a.path().mapping_size(); i++)
			{
				if (i) o += ",";
				const vg::Mapping& m = a.path().mapping(i); o += "{"; bool mf = true;
				if (m.has_position())
				{
					sep(mf); o += "\"position\":{"; bool qf = true; const vg::Position& p = m.position();
					if (p.node_id()) { sep(qf); o += "\"node_id\":\"" + std::to_string(p.node_id()) + "\""; }
					if (p.offset()) { sep(qf); o += "\"offset\":\"" + std::to_string(p.offset()) + "\""; }
					if (p.is_reverse()) { sep(qf); o += "\"is_reverse\":true"; }
					if (!p.name().empty()) { sep(qf); o += "\"name\":" + gcJsonEscape(p.name()); }
					o += "}";
				}
				if (m.edit_size() > 0)
				{
					sep(mf); o += "\"edit\":[";
					for (int e = 0; e < m.edit_size(); e++)
					{
						if (e) o += ",";
						o += "{"; bool ef = true;
						if (m.edit(e).from_length()) { sep(ef); o += "\"from_length\":" + std::to_string(m.edit(e).from_length()); }
						if (m.edit(e).to_length()) { sep(ef); o += "\"to_length\":" + std::to_string(m.edit(e).to_length()); }
						if (!m.edit(e).sequence().empty()) { sep(ef); o += "\"sequence\":" + gcJsonEscape(m.edit(e).sequence()); }
						o += "}";
					}
					o += "]";
				}
				if (m.rank()) { sep(mf); o += "\"rank\":\"" + std::to_string(m.rank()) + "\""; }
				o += "}";
			}
			o += "]";
		}
		o += "}";
	}
	if (!a.name().empty()) { sep(first); o += "\"name\":" + gcJsonEscape(a.name()); }
	if (a.score()) { sep(first); o += "\"score\":" + std::to_string(a.score()); }
	if (a.query_position()) { sep(first); o += "\"query_position\":" + std::to_string(a.query_position()); }
	if (a.identity() != 0) { sep(first); o += "\"identity\":" + gcJsonDouble(a.identity()); }
	o += "}";
}
} } }
#endif
