// TEST INFRASTRUCTURE ONLY (oracle build shim) -- see oracle/shim/vg.pb.h.
#ifndef GC_ORACLE_SHIM_PB_CODED_H
#define GC_ORACLE_SHIM_PB_CODED_H
#include <string>
#include "google/protobuf/io/zero_copy_stream.h"
namespace google { namespace protobuf { namespace io {
class CodedOutputStream
{
public:
	explicit CodedOutputStream(ZeroCopyOutputStream* sub) : sub(sub) {}
	void WriteVarint64(uint64_t v)
	{
		char buf[10]; int n = 0;
		while (v >= 0x80) { buf[n++] = (char)((v & 0x7F) | 0x80); v >>= 7; }
		buf[n++] = (char)v;
		sub->Append(buf, n);
	}
	void WriteVarint32(uint32_t v) { WriteVarint64(v); }
	void WriteRaw(const void* data, int size) { sub->Append((const char*)data, (size_t)size); }
private:
	ZeroCopyOutputStream* sub;
};
class CodedInputStream
{
public:
	explicit CodedInputStream(ZeroCopyInputStream* sub) : sub(sub) {}
	bool ReadVarint64(uint64* out)
	{
		uint64_t r = 0; int shift = 0; char c;
		while (true)
		{
			if (sub->Fetch(&c, 1) != 1) return false;
			r |= (uint64_t)((unsigned char)c & 0x7F) << shift;
			if (!((unsigned char)c & 0x80)) break;
			shift += 7;
			if (shift > 63) return false;
		}
		*out = r;
		return true;
	}
	bool ReadVarint32(uint32_t* out) { uint64 v; if (!ReadVarint64(&v)) return false; *out = (uint32_t)v; return true; }
	bool ReadString(std::string* s, int size)
	{
		s->resize((size_t)size);
		size_t got = 0;
		while (got < (size_t)size) { size_t n = sub->Fetch(&(*s)[got], (size_t)size - got); if (n == 0) return false; got += n; }
		return true;
	}
private:
	ZeroCopyInputStream* sub;
};
} } }
#endif
