// deflate_check <file.gam>...: every gzip member of the given GAM files is inflated with zlib, re-encoded with the
// driver's own DEFLATE encoder (graphchainer_b200/csrc/gc_deflate.h), inflated with zlib again and compared; then the
// same round trip on edge cases (empty, tiny, long runs, incompressible, skewed alphabets that need the length limit).
#include <cstdio>
#include <fstream>
#include <random>
#include <sstream>
#include "../../graphchainer_b200/csrc/gc_deflate.h"

static bool inflateMember(const std::string& gz, size_t& pos, std::string& out)
{
	z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, 15 + 16);
	out.assign(1 << 20, 0);
	zs.next_in = (Bytef*)gz.data() + pos; zs.avail_in = (uInt)(gz.size() - pos);
	size_t have = 0; int rc;
	do
	{
		if (have == out.size()) out.resize(out.size() * 2);
		zs.next_out = (Bytef*)&out[have]; zs.avail_out = (uInt)(out.size() - have);
		rc = inflate(&zs, Z_NO_FLUSH);
		have = out.size() - zs.avail_out;
	} while (rc == Z_OK);
	out.resize(have); pos += zs.total_in; inflateEnd(&zs);
	return rc == Z_STREAM_END;
}

int main(int argc, char** argv)
{
	std::vector<std::string> raws;
	for (int a = 1; a < argc; a++)
	{
		std::ifstream in(argv[a], std::ios::binary); std::stringstream ss; ss << in.rdbuf(); std::string data = ss.str();
		size_t pos = 0; std::string rec;
		while (pos < data.size()) { if (!inflateMember(data, pos, rec)) { fprintf(stderr, "cannot inflate %s\n", argv[a]); return 2; } raws.push_back(rec); }
	}
	size_t fromFiles = raws.size();
	std::mt19937 rng(1);
	raws.push_back(""); raws.push_back("a"); raws.push_back("abc"); raws.push_back("abcd"); raws.push_back(std::string(100000, 'x')); raws.push_back("abcdabcdabcdabcdabcdabcd");
	{ std::string r(70000, 0); for (auto& c : r) c = (char)rng(); raws.push_back(r); }
	{ std::string r(50000, 0); for (auto& c : r) c = "ACGT"[rng() & 3]; raws.push_back(r); }
	{ std::string r(300, 0); for (size_t i = 0; i < r.size(); i++) r[i] = (char)i; raws.push_back(r); }
	{ std::string r; for (int i = 0; i < 200000; i++) r.push_back((char)(rng() % 3 ? 'a' : (rng() & 255))); raws.push_back(r); }
	// Fibonacci-like frequencies: the unrestricted Huffman tree is deeper than 15 levels, the length limit must repair it
	{ std::string r; uint64_t f0 = 1, f1 = 1; for (int sym = 0; sym < 30; sym++) { for (uint64_t k = 0; k < f0 && r.size() < 3000000; k++) r.push_back((char)(sym * 7 + 1)); uint64_t t = f0 + f1; f0 = f1; f1 = t; } std::shuffle(r.begin(), r.end(), rng); raws.push_back(r); }
	{ std::string r(40000, 0); for (size_t i = 0; i < r.size(); i++) r[i] = (char)(i % 33000 < 16500 ? rng() : r[i - 16500 + 0]); raws.push_back(r); }
	gcdeflate::Encoder enc;
	size_t bad = 0, fallback = 0, rawBytes = 0, outBytes = 0;
	for (const std::string& r : raws)
	{
		std::string gz = enc.gzipMember(r);
		if (gz.empty()) { fallback++; continue; } // the caller would use zlib
		size_t pos = 0; std::string back;
		if (!inflateMember(gz, pos, back) || pos != gz.size() || back != r) bad++;
		rawBytes += r.size(); outBytes += gz.size();
	}
	printf("records %zu from_files %zu raw %zu out %zu bad %zu fallback %zu\n", raws.size(), fromFiles, rawBytes, outBytes, bad, fallback);
	return bad ? 1 : 0;
}
