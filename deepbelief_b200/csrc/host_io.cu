// Host-side batch assembly for the data loader (deepbelief/layers/data.py:92-94 picks a batch with a
// fancy-indexed read `datapoints[sorted_rows]`; at 16 KB rows x 16384 rows that single-threaded gather
// costs more than the whole CD-10 step on the GPU). Rows are copied by a small pool of std::threads
// straight into the caller's (pinned) staging buffer. No device code in this file.
#include "context.h"
#include <cstring>
#include <thread>
#include <vector>

extern "C" int db_host_gather_rows(const void* src, size_t row_bytes, const int64_t* rows, size_t n_rows,
                                   size_t n_src_rows, void* dst, int threads) {
  if (!src || !rows || !dst || row_bytes == 0) return DB_ERR_BAD_ARG;
  for (size_t i = 0; i < n_rows; ++i)
    if (rows[i] < 0 || (size_t)rows[i] >= n_src_rows) return DB_ERR_BAD_ARG;
  if (threads <= 0) {
    unsigned hc = std::thread::hardware_concurrency();
    threads = hc == 0 ? 4 : (int)(hc > 16 ? 16 : hc);
  }
  const size_t total = n_rows * row_bytes;
  if (total < (size_t)4 << 20 || threads == 1) threads = 1;      // small batches: not worth a thread launch
  if ((size_t)threads > n_rows) threads = (int)(n_rows ? n_rows : 1);
  const char* s = static_cast<const char*>(src);
  char* d = static_cast<char*>(dst);
  auto work = [=](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) std::memcpy(d + i * row_bytes, s + (size_t)rows[i] * row_bytes, row_bytes);
  };
  if (threads == 1) { work(0, n_rows); return DB_OK; }
  std::vector<std::thread> pool;
  pool.reserve(threads);
  const size_t per = (n_rows + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const size_t lo = (size_t)t * per, hi = lo + per < n_rows ? lo + per : n_rows;
    if (lo >= hi) break;
    pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return DB_OK;
}
