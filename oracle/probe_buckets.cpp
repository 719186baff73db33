// Test infrastructure (oracle). Prints the bucket-count growth sequence of libstdc++'s
// std::unordered_map<size_t,...> (identity hash, max_load_factor 1.0): one line per rehash,
// "<element count at which the rehash fires> <new bucket count>".
// The reference's grid_subsampling_cpu.cpp:26,45-47 iterates such a map, so its output order depends on this.
#include <cstdio>
#include <cstdlib>
#include <unordered_map>
int main(int argc, char** argv) {
  size_t limit = argc > 1 ? (size_t)atoll(argv[1]) : 3000000;
  std::unordered_map<size_t, int> m;
  size_t b = m.bucket_count();
  for (size_t i = 0; i < limit; i++) {
    m.emplace(i, 0);
    if (m.bucket_count() != b) { b = m.bucket_count(); printf("%zu %zu\n", i + 1, b); }
  }
  return 0;
}
