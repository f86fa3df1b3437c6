// row_writer_check.cpp — compairr_b200/csrc/cli/row_writer.h writes the same bytes with any number
// of threads and any block size.  Compiled with g++ by tests/test_row_writer.py; no GPU.
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../compairr_b200/csrc/cli/row_writer.h"

static std::string render(uint64_t n, int threads, uint64_t block) {
  FILE* f = tmpfile();
  write_rows_parallel(f, n, threads, [&](uint64_t a, uint64_t b, std::string& buf) {
    for (uint64_t i = a; i < b; i++) {
      if (i % 7 == 3) continue;  // rows may be skipped (dedup prints group leaders only)
      append_u64(buf, i * 2654435761u);
      buf += '\t';
      buf.append((size_t)(i % 13), 'x');
      buf += '\n';
    }
  }, block);
  fflush(f);
  const long size = ftell(f);
  std::string out((size_t)size, '\0');
  rewind(f);
  if (size && fread(&out[0], 1, (size_t)size, f) != (size_t)size) out.clear();
  fclose(f);
  return out;
}

int main() {
  std::string u;
  append_u64(u, 0);
  append_u64(u, 18446744073709551615ull);
  if (u != "018446744073709551615") return 2;
  for (uint64_t n : {0ull, 1ull, 999ull, 100000ull}) {
    const std::string want = render(n, 1, 1u << 15);
    for (int threads : {2, 5, 16})
      for (uint64_t block : {1ull, 64ull, 4096ull, 1ull << 15})
        if (render(n, threads, block) != want) {
          fprintf(stderr, "row_writer_check: mismatch n=%llu threads=%d block=%llu\n", (unsigned long long)n, threads,
                  (unsigned long long)block);
          return 1;
        }
  }
  if (host_threads(7, true) != 7 || host_threads(1, false) < 1 || host_threads(1, true) != 1) return 3;
  printf("row_writer_check ok\n");
  return 0;
}
