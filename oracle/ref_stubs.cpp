// TEST INFRASTRUCTURE ONLY (oracle build) -- not part of the product.
// MUM/MEM seeding (src/MummerSeeder.cpp) needs mummer4 + boost::serialization,
// which are absent here, and is a non-default seeding mode (AlignerMain.cpp:186-193
// selects minimizer seeding).  The reference's Aligner.cpp still references the
// class, so its public methods are stubbed to abort if ever reached.
#include <cstdlib>
#include <iostream>
#include "MummerSeeder.h"

static void gcNoMummer()
{
	std::cerr << "oracle build: MUM/MEM seeding is not available (mummer4 absent)" << std::endl;
	std::abort();
}
MummerSeeder::MummerSeeder(const GfaGraph&, const std::string&) { gcNoMummer(); }
MummerSeeder::MummerSeeder(const vg::Graph&, const std::string&) { gcNoMummer(); }
std::vector<SeedHit> MummerSeeder::getMemSeeds(std::string, size_t, size_t) const { gcNoMummer(); return {}; }
std::vector<SeedHit> MummerSeeder::getMumSeeds(std::string, size_t, size_t) const { gcNoMummer(); return {}; }
