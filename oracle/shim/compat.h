// Force-included glue so the reference's MSVC-dialect translation units build with g++.
// Test infrastructure only (oracle/_ref build); contains no reference logic.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdint>
#include <cassert>
#include <ctime>
#include <functional>
#include <unordered_map>
#include <condition_variable>
#define __forceinline inline __attribute__((always_inline))
#define __declspec(x)
// WorldOctreeNode.hpp:83 calls the MSVC-internal std::hash<uint32_t>::_Do_hash(v); libstdc++'s
// integer hash is the identity, so a functional cast to hash<>::result_type is equivalent.
#define _Do_hash(v) result_type(v)
static inline void* _aligned_malloc(size_t size, size_t align)
{
	void* p = nullptr;
	if (align < sizeof(void*)) align = sizeof(void*);
	if (posix_memalign(&p, align, size ? size : align)) return nullptr;
	return p;
}
static inline void _aligned_free(void* p) { free(p); }
