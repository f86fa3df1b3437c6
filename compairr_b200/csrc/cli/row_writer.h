// row_writer.h — formats output rows on several host threads and writes them in row order.
//
// The reference prints every row with fprintf from one thread (pairs under the global mutex,
// src/overlap.cc:455-507; cluster rows, src/cluster.cc:420-444; dedup rows, src/dedup.cc:45-58): at
// 10^7 rows that is seconds, more than the GPU phases take.  Here rows are cut into blocks; each
// thread formats whole blocks into its own string and hands them to the file strictly in block
// order (a ticket: block k is written after block k-1), so the bytes are the same as a serial
// loop's.
#pragma once
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

// fn(first_row, last_row_exclusive, out): append the text of rows [first, last) to out.
template <typename Fn>
void write_rows_parallel(FILE* f, uint64_t n_rows, int threads, Fn fn, uint64_t block = 1u << 15) {
  if (n_rows == 0) return;
  const uint64_t n_blocks = (n_rows + block - 1) / block;
  const int nt = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(threads, 1), n_blocks));
  if (nt == 1) {
    std::string buf;
    for (uint64_t b = 0; b < n_blocks; b++) {
      buf.clear();
      fn(b * block, std::min(n_rows, (b + 1) * block), buf);
      fwrite(buf.data(), 1, buf.size(), f);
    }
    return;
  }
  std::atomic<uint64_t> next{0};
  std::mutex mu;
  std::condition_variable cv;
  uint64_t turn = 0;  // next block the file expects
  auto work = [&] {
    std::string buf;
    for (;;) {
      const uint64_t b = next.fetch_add(1);
      if (b >= n_blocks) return;
      buf.clear();
      fn(b * block, std::min(n_rows, (b + 1) * block), buf);
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return turn == b; });
      fwrite(buf.data(), 1, buf.size(), f);
      turn = b + 1;
      lk.unlock();
      cv.notify_all();
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
}

// Host threads for the reader and the writers: -t N if given (also -t 1), else the machine's (at
// most 32).  The reference's -t 1 default means "one worker"; here the workers are on the GPU and
// -t only sizes the host-side parsing and formatting, which never change a byte of the output.
inline int host_threads(int64_t opt_threads, bool given) {
  if (given || opt_threads > 1) return (int)std::max<int64_t>(opt_threads, 1);
  const unsigned hw = std::thread::hardware_concurrency();
  return (int)std::max(1u, std::min(hw, 32u));
}

// decimal digits of an unsigned value, appended (std::to_string allocates)
inline void append_u64(std::string& s, uint64_t v) {
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) s += tmp[--n];
}
