// TEST INFRASTRUCTURE ONLY (oracle build shim) -- see oracle/shim/vg.pb.h.
#ifndef GC_ORACLE_SHIM_PB_COMMON_H
#define GC_ORACLE_SHIM_PB_COMMON_H
#include <cstdint>
namespace google { namespace protobuf {
typedef uint64_t uint64;
typedef uint32_t uint32;
typedef int64_t int64;
typedef int32_t int32;
} }
#endif
