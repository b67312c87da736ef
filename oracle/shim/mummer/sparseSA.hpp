// TEST INFRASTRUCTURE ONLY (oracle build shim) -- see oracle/shim/vg.pb.h.
// Declarations only: MUM/MEM seeding (MummerSeeder.cpp) is a non-default mode,
// not compiled into the oracle; its public methods are stubbed in ref_stubs.cpp.
#ifndef GC_ORACLE_SHIM_MUMMER_H
#define GC_ORACLE_SHIM_MUMMER_H
#include <string>
#include <vector>
namespace mummer { namespace mummer {
struct match_t { long ref; long query; long len; };
class sparseSA { public: sparseSA() {} };
} }
#endif
