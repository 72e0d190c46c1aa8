// Build shim for the reference oracle (oracle/_ref): the reference includes
// <ankerl/unordered_dense.h> (vcpkg package, not on this image).  Its three maps are only used
// through contains()/operator[] and never iterated (/root/reference/src/paf_data.cpp:739-740,
// 756-760, 1493-1494, 1562), so std::unordered_map is semantics-preserving.
// TEST INFRASTRUCTURE ONLY — nothing in the product includes this.
#pragma once
#include <unordered_map>
namespace ankerl {
namespace unordered_dense {
template <class K, class V>
using map = std::unordered_map<K, V>;
}
}  // namespace ankerl
