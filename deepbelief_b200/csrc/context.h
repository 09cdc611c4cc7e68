// Internal: the opaque handle behind db_handle_t and the error-reporting helpers of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/deepbelief_b200.h"
#include "common.cuh"

struct db_context_s {
  int device;
  int num_sms;
  int cc_major, cc_minor;
  char err[512];
  // fork/join resources of the GP chain (gp.cu, db::side_lane): two auxiliary streams and a pool of timing-less events,
  // created on first use and destroyed with the handle. No device memory.
  cudaStream_t side_stream, side_stream2;
  static constexpr int kSideEvents = 96;
  cudaEvent_t side_events[kSideEvents];
  int side_ready;          // 0 not created yet, 1 usable, -1 creation failed (the chain then stays on one stream)
};

namespace db {

inline int set_err(db_handle_t h, int code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

// Creates the auxiliary stream / events on first use; false when they are not available.
inline bool side_lane(db_handle_t h, cudaStream_t caller) {
  if (h->side_ready != 0) return h->side_ready > 0;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;      // no resource creation inside somebody's capture
  if (cudaStreamIsCapturing(caller, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return false; }
  h->side_ready = -1;
  if (cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
  if (cudaStreamCreateWithFlags(&h->side_stream2, cudaStreamNonBlocking) != cudaSuccess) {
    cudaStreamDestroy(h->side_stream); cudaGetLastError(); return false;
  }
  for (int i = 0; i < db_context_s::kSideEvents; ++i)
    if (cudaEventCreateWithFlags(&h->side_events[i], cudaEventDisableTiming) != cudaSuccess) {
      for (int k = 0; k < i; ++k) cudaEventDestroy(h->side_events[k]);
      cudaStreamDestroy(h->side_stream);
      cudaStreamDestroy(h->side_stream2);
      cudaGetLastError();
      return false;
    }
  h->side_ready = 1;
  return true;
}

inline int tc_ctas(db_handle_t h) { return h->num_sms; }     // persistent tensor-core kernels: one CTA per SM

inline int check_launch(db_handle_t h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(h, DB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return DB_OK;
}

inline void to_rng(const db_rng_t* r, RngArgs& o) {
  if (r) {
    o.seed_lo = (uint32_t)r->seed; o.seed_hi = (uint32_t)(r->seed >> 32);
    o.draw_lo = (uint32_t)r->draw; o.draw_hi = (uint32_t)(r->draw >> 32);
    o.row_offset = r->row_offset;
    o.clock = (const unsigned long long*)r->clock;
  } else {
    o.seed_lo = o.seed_hi = o.draw_lo = o.draw_hi = o.row_offset = 0;
    o.clock = nullptr;
  }
}

}  // namespace db

#define DB_REQUIRE(h, cond, msg) \
  do { if (!(cond)) return db::set_err((h), DB_ERR_BAD_ARG, "%s: %s", __func__, (msg)); } while (0)
