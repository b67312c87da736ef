// TEST INFRASTRUCTURE ONLY (oracle build shim) -- see oracle/shim/vg.pb.h.
// Minimal byte-sink / byte-source interfaces with the protobuf class names the
// reference uses (Aligner.cpp:230-240,264-279; stream.hpp:27-47,86-116).
#ifndef GC_ORACLE_SHIM_PB_ZCS_H
#define GC_ORACLE_SHIM_PB_ZCS_H
#include <cstddef>
#include <string>
#include "google/protobuf/stubs/common.h"
namespace google { namespace protobuf { namespace io {
class ZeroCopyOutputStream
{
public:
	virtual ~ZeroCopyOutputStream() {}
	virtual void Append(const char* data, size_t n) = 0;
};
class ZeroCopyInputStream
{
public:
	virtual ~ZeroCopyInputStream() {}
	// read up to n bytes, return the number read (0 at end of data)
	virtual size_t Fetch(char* out, size_t n) = 0;
};
} } }
#endif
