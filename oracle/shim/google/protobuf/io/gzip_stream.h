// TEST INFRASTRUCTURE ONLY (oracle build shim) -- see oracle/shim/vg.pb.h.
// GzipOutputStream: collects bytes and, on destruction, writes ONE gzip member
// (zlib default level, like protobuf's GzipOutputStream defaults) to the sink.
// GzipInputStream: inflates a stream of concatenated gzip members.
#ifndef GC_ORACLE_SHIM_PB_GZIP_H
#define GC_ORACLE_SHIM_PB_GZIP_H
#include <zlib.h>
#include <string>
#include <vector>
#include <cstring>
#include "google/protobuf/io/zero_copy_stream.h"
namespace google { namespace protobuf { namespace io {
class GzipOutputStream : public ZeroCopyOutputStream
{
public:
	explicit GzipOutputStream(ZeroCopyOutputStream* sub) : sub(sub) {}
	~GzipOutputStream() override
	{
		z_stream zs; std::memset(&zs, 0, sizeof(zs));
		deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY);
		zs.next_in = (Bytef*)buffer.data(); zs.avail_in = (uInt)buffer.size();
		char out[65536];
		int ret;
		do {
			zs.next_out = (Bytef*)out; zs.avail_out = sizeof(out);
			ret = deflate(&zs, Z_FINISH);
			sub->Append(out, sizeof(out) - zs.avail_out);
		} while (ret != Z_STREAM_END && ret == Z_OK);
		deflateEnd(&zs);
	}
	void Append(const char* data, size_t n) override { buffer.append(data, n); }
private:
	ZeroCopyOutputStream* sub;
	std::string buffer;
};
class GzipInputStream : public ZeroCopyInputStream
{
public:
	explicit GzipInputStream(ZeroCopyInputStream* sub) : sub(sub), pos(0), loaded(false) {}
	size_t Fetch(char* out, size_t n) override
	{
		load();
		size_t avail = data.size() - pos;
		if (n > avail) n = avail;
		std::memcpy(out, data.data() + pos, n);
		pos += n;
		return n;
	}
private:
	void load()
	{
		if (loaded) return;
		loaded = true;
		std::string raw; char buf[65536]; size_t got;
		while ((got = sub->Fetch(buf, sizeof(buf))) > 0) raw.append(buf, got);
		size_t rawpos = 0;
		while (rawpos < raw.size())
		{
			z_stream zs; std::memset(&zs, 0, sizeof(zs));
			if (inflateInit2(&zs, 15 + 32) != Z_OK) return;
			zs.next_in = (Bytef*)raw.data() + rawpos; zs.avail_in = (uInt)(raw.size() - rawpos);
			int ret;
			do {
				zs.next_out = (Bytef*)buf; zs.avail_out = sizeof(buf);
				ret = inflate(&zs, Z_NO_FLUSH);
				if (ret != Z_OK && ret != Z_STREAM_END) { inflateEnd(&zs); return; }
				data.append(buf, sizeof(buf) - zs.avail_out);
			} while (ret != Z_STREAM_END);
			rawpos = raw.size() - zs.avail_in;
			inflateEnd(&zs);
		}
	}
	ZeroCopyInputStream* sub;
	std::string data;
	size_t pos;
	bool loaded;
};
} } }
#endif
