// Context management of the C ABI (include/deepbelief_b200.h). The library is stateless apart from
// this handle (device id, SM count, last error text); it never owns device memory.
#include "context.h"
#include <new>

extern "C" int db_version(void) { return 100; }

extern "C" int db_create(db_handle_t* out, int device) {
  if (!out) return DB_ERR_BAD_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DB_ERR_CUDA;   // no CPU fallback
  if (device < 0 || device >= count) return DB_ERR_BAD_ARG;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DB_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return DB_ERR_CUDA;
  db_context_s* c = new (std::nothrow) db_context_s();
  if (!c) return DB_ERR_CUDA;
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->err[0] = 0;
  *out = c;
  return DB_OK;
}

extern "C" int db_destroy(db_handle_t h) {
  delete h;
  return DB_OK;
}

extern "C" const char* db_last_error(db_handle_t h) { return h ? h->err : "null handle"; }

extern "C" int db_num_sms(db_handle_t h) { return h ? h->num_sms : 0; }

