// TEST INFRASTRUCTURE ONLY (oracle build shim) -- see oracle/shim/vg.pb.h.
// sdsl::int_vector<0> is used by the reference (MinimizerSeeder.h:26-28) purely
// as a packed integer array; a plain 64-bit vector with a remembered width has
// the same observable behaviour.
#ifndef GC_ORACLE_SHIM_SDSL_INT_VECTOR_H
#define GC_ORACLE_SHIM_SDSL_INT_VECTOR_H
#include <cstdint>
#include <vector>
namespace sdsl {
template <int W>
class int_vector : public std::vector<uint64_t>
{
public:
	int_vector() : std::vector<uint64_t>(), w(64) {}
	void width(uint8_t nw) { w = nw; }
	uint8_t width() const { return w; }
private:
	uint8_t w;
};
namespace util {
template <typename V> void set_to_value(V& v, uint64_t val) { for (auto& x : v) x = val; }
}
}
#endif
