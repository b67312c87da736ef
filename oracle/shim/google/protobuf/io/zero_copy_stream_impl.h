// TEST INFRASTRUCTURE ONLY (oracle build shim) -- see oracle/shim/vg.pb.h.
#ifndef GC_ORACLE_SHIM_PB_ZCSI_H
#define GC_ORACLE_SHIM_PB_ZCSI_H
#include <istream>
#include <ostream>
#include "google/protobuf/io/zero_copy_stream.h"
namespace google { namespace protobuf { namespace io {
class OstreamOutputStream : public ZeroCopyOutputStream
{
public:
	explicit OstreamOutputStream(std::ostream* out) : out(out) {}
	void Append(const char* data, size_t n) override { out->write(data, (std::streamsize)n); }
private:
	std::ostream* out;
};
class IstreamInputStream : public ZeroCopyInputStream
{
public:
	explicit IstreamInputStream(std::istream* in) : in(in) {}
	size_t Fetch(char* buf, size_t n) override { in->read(buf, (std::streamsize)n); return (size_t)in->gcount(); }
private:
	std::istream* in;
};
} } }
#endif
