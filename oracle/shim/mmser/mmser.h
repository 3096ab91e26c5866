// Stand-in for the un-vendored dependency mmser 1.0.1 (cpmpack.json) so that the reference headers
// compile without network access.  TEST INFRASTRUCTURE ONLY (oracle/_ref build).  mmser::vector is a
// mmap-able std::vector replacement; only the std::vector interface is used on the search path.
#pragma once
#include <type_traits>
#include <vector>
namespace mmser {
template <typename T> using vector = std::vector<T>;
template <typename T> struct is_trivially_copyable_t : std::is_trivially_copyable<T> {};
}
