// Flat alignment-graph index (".gcidx"): named little-endian arrays.
//
// The arrays are exactly the members of the reference's AlignmentGraph
// (AlignmentGraph.h:145-172), its MPC index (AlignmentGraph.cpp:1328-1391) and its
// minimizer index (MinimizerSeeder.h:17-29) in reference numbering; they are what
// gcgpu_create() takes (include/gcgpu.h).  The file form doubles as the on-disk
// index cache the reference left as a stub (AlignmentGraph.cpp:1490-1495).
//
// Layout: "GCIDX001", then records { u32 nameLen, name, u8 dtype, u64 count, data }
// with dtype 1=u8 4=u32 5=i32 8=u64.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

struct GcIndexArray
{
	uint8_t dtype = 0;
	uint64_t count = 0;
	std::vector<uint8_t> bytes;
	size_t elemSize() const { return dtype == 1 ? 1 : (dtype == 8 ? 8 : 4); }
	const uint8_t* u8() const { return bytes.data(); }
	const uint32_t* u32() const { return (const uint32_t*)bytes.data(); }
	const int32_t* i32() const { return (const int32_t*)bytes.data(); }
	const uint64_t* u64() const { return (const uint64_t*)bytes.data(); }
};

struct GcIndexFile
{
	std::map<std::string, GcIndexArray> arrays;
	std::vector<std::string> order;

	const GcIndexArray& get(const std::string& name) const
	{
		auto it = arrays.find(name);
		if (it == arrays.end()) throw std::runtime_error("gcidx: missing array " + name);
		return it->second;
	}
	bool has(const std::string& name) const { return arrays.count(name) != 0; }

	void put(const std::string& name, uint8_t dtype, uint64_t count, const void* data)
	{
		GcIndexArray a;
		a.dtype = dtype;
		a.count = count;
		a.bytes.resize(count * a.elemSize());
		if (count) std::memcpy(a.bytes.data(), data, a.bytes.size());
		if (!arrays.count(name)) order.push_back(name);
		arrays[name] = std::move(a);
	}
	void putU8(const std::string& n, const std::vector<uint8_t>& v) { put(n, 1, v.size(), v.data()); }
	void putU32(const std::string& n, const std::vector<uint32_t>& v) { put(n, 4, v.size(), v.data()); }
	void putI32(const std::string& n, const std::vector<int32_t>& v) { put(n, 5, v.size(), v.data()); }
	void putU64(const std::string& n, const std::vector<uint64_t>& v) { put(n, 8, v.size(), v.data()); }

	void load(const std::string& path)
	{
		FILE* f = std::fopen(path.c_str(), "rb");
		if (!f) throw std::runtime_error("gcidx: cannot open " + path);
		char magic[8];
		if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "GCIDX001", 8) != 0) { std::fclose(f); throw std::runtime_error("gcidx: bad magic in " + path); }
		while (true)
		{
			uint32_t nl;
			if (std::fread(&nl, 4, 1, f) != 1) break;
			std::string name(nl, '\0');
			GcIndexArray a;
			if (std::fread(&name[0], 1, nl, f) != nl || std::fread(&a.dtype, 1, 1, f) != 1 || std::fread(&a.count, 8, 1, f) != 1) { std::fclose(f); throw std::runtime_error("gcidx: truncated " + path); }
			a.bytes.resize(a.count * a.elemSize());
			if (a.bytes.size() && std::fread(a.bytes.data(), 1, a.bytes.size(), f) != a.bytes.size()) { std::fclose(f); throw std::runtime_error("gcidx: truncated " + path); }
			order.push_back(name);
			arrays[name] = std::move(a);
		}
		std::fclose(f);
	}
	void save(const std::string& path) const
	{
		FILE* f = std::fopen(path.c_str(), "wb");
		if (!f) throw std::runtime_error("gcidx: cannot write " + path);
		std::fwrite("GCIDX001", 1, 8, f);
		for (const auto& name : order)
		{
			const GcIndexArray& a = arrays.at(name);
			uint32_t nl = (uint32_t)name.size();
			std::fwrite(&nl, 4, 1, f); std::fwrite(name.data(), 1, nl, f); std::fwrite(&a.dtype, 1, 1, f); std::fwrite(&a.count, 8, 1, f);
			if (a.bytes.size()) std::fwrite(a.bytes.data(), 1, a.bytes.size(), f);
		}
		std::fclose(f);
	}
};
