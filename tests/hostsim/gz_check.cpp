// The gzip members of gc_gam.cuh (the 32-lane stages, run lane after lane on the host) must inflate with zlib to the bytes
// they were made from: all sizes around the chunking (32 chunks), incompressible, constant and record-like inputs.
#include <zlib.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include "../../graphchainer_b200/csrc/gc_gam.cuh"

static bool check(const GcDeflateTables& T, const std::vector<uint8_t>& raw, const char* what)
{
	uint32_t n = (uint32_t)raw.size();
	std::vector<uint8_t> gz((size_t)n * 2 + 1024), wsBuf(gc_deflate_ws_bytes(n));
	GcDeflateWs ws; ws.tokens = (uint32_t*)wsBuf.data(); ws.tokenCap = n + 16;
	uint32_t size = gc_gzip_member(T, raw.data(), n, ws, gz.data(), (uint32_t)gz.size());
	if (size == 0) { fprintf(stderr, "%s n=%u: no member\n", what, n); return false; }
	std::vector<uint8_t> back(n + 16);
	z_stream zs; memset(&zs, 0, sizeof(zs));
	if (inflateInit2(&zs, 15 + 16) != Z_OK) return false;
	zs.next_in = gz.data(); zs.avail_in = size; zs.next_out = back.data(); zs.avail_out = (uInt)back.size();
	int rc = inflate(&zs, Z_FINISH);
	bool ok = rc == Z_STREAM_END && zs.total_out == n && zs.avail_in == 0 && memcmp(back.data(), raw.data(), n) == 0;
	inflateEnd(&zs);
	if (!ok) fprintf(stderr, "%s n=%u: inflate rc=%d total_out=%lu avail_in=%u (member %u bytes)\n", what, n, rc, (unsigned long)zs.total_out, zs.avail_in, size);
	return ok;
}

int main()
{
	GcDeflateTables* T = new GcDeflateTables; gcBuildGamTables(*T);
	std::mt19937 rng(7);
	int bad = 0, cases = 0; uint64_t rawBytes = 0;
	const uint32_t sizes[] = { 0, 1, 3, 4, 5, 31, 32, 33, 63, 64, 65, 127, 129, 255, 1000, 4099, 34000, 70001, 300000 };
	for (uint32_t n : sizes)
	{
		std::vector<uint8_t> a(n), b(n, 0x41), c(n), d(n);
		for (uint32_t i = 0; i < n; i++) a[i] = (uint8_t)rng();                                 // incompressible
		const char* text = "\x12\x1a\x0a\x0c\x08\x95\x03\x10\x07\x2a\x06s10231\x12\x04\x08\x05\x10\x05\x12\x08\x08\x01\x10\x01\x1a\x01G";
		size_t tl = 30;
		for (uint32_t i = 0; i < n; i++) c[i] = (rng() % 13 == 0) ? (uint8_t)("ACGT"[rng() & 3]) : (uint8_t)text[i % tl]; // record-like: tags, names, edits
		for (uint32_t i = 0; i < n; i++) d[i] = (uint8_t)("ACGT"[rng() & 3]);                   // a read's characters
		for (auto* v : { &a, &b, &c, &d }) { cases++; rawBytes += n; if (!check(*T, *v, v == &a ? "random" : v == &b ? "constant" : v == &c ? "record" : "dna")) bad++; }
	}
	printf("{\"cases\":%d,\"bad\":%d,\"raw_bytes\":%llu}\n", cases, bad, (unsigned long long)rawBytes);
	delete T;
	return bad ? 1 : 0;
}
