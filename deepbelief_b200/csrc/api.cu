// Context management of the C ABI (include/deepbelief_b200.h). The library is stateless apart from
// this handle (device id, SM count, last error text, and — once the GP chain has run — two auxiliary
// streams with their fork/join events); it never owns device memory.
#include "context.h"
#include <cstring>
#include <cstdint>
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
  c->side_ready = 0;
  *out = c;
  return DB_OK;
}

extern "C" int db_destroy(db_handle_t h) {
  if (h && h->side_ready > 0) {
    for (int i = 0; i < db_context_s::kSideEvents; ++i) cudaEventDestroy(h->side_events[i]);
    cudaStreamDestroy(h->side_stream);
    cudaStreamDestroy(h->side_stream2);
  }
  delete h;
  return DB_OK;
}

extern "C" int db_enable_peer_access(db_handle_t h, int peer_device) {
  DB_REQUIRE(h, peer_device >= 0, "bad device index");
  int cur = -1;
  cudaGetDevice(&cur);
  if (peer_device == cur) return DB_OK;
  int can = 0;
  cudaDeviceCanAccessPeer(&can, cur, peer_device);
  if (!can) return db::set_err(h, DB_ERR_UNSUPPORTED, "device %d cannot access device %d (no NVLink / PCIe peer path)", cur, peer_device);
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return DB_OK; }
  if (e != cudaSuccess) return db::set_err(h, DB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
  return DB_OK;
}

// ---- CUDA IPC: one rank's device buffer mapped into the other ranks' processes (single node) -------------------------------
// Export: the handle of the cudaMalloc block that contains ptr (a framework's caching allocator hands out sub-ranges) and
// the offset of ptr inside it. Import: opened in the CURRENT device's context with lazy peer access, so kernels of this
// device may dereference the result (a pointer imported in the exporting device's context is not reachable from here).
extern "C" int db_ipc_export(db_handle_t h, const void* ptr, void* handle_out64, unsigned long long* offset_out) {
  DB_REQUIRE(h, ptr && handle_out64 && offset_out, "null operand");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle is exchanged as 64 bytes");
  typedef int (*PFN_range)(unsigned long long*, size_t*, unsigned long long);      // cuMemGetAddressRange_v2
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return db::set_err(h, DB_ERR_CUDA, "cuMemGetAddressRange is not available");
  unsigned long long base = 0; size_t size = 0;
  if (reinterpret_cast<PFN_range>(fn)(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0)
    return db::set_err(h, DB_ERR_CUDA, "cuMemGetAddressRange failed for %p", ptr);
  cudaIpcMemHandle_t hd;
  const cudaError_t e = cudaIpcGetMemHandle(&hd, reinterpret_cast<void*>((uintptr_t)base));
  if (e != cudaSuccess) return db::set_err(h, DB_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  memcpy(handle_out64, &hd, 64);
  *offset_out = (unsigned long long)(uintptr_t)ptr - base;
  return DB_OK;
}

extern "C" int db_ipc_import(db_handle_t h, const void* handle64, unsigned long long offset, void** ptr_out) {
  DB_REQUIRE(h, handle64 && ptr_out, "null operand");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  void* base = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return db::set_err(h, DB_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  *ptr_out = static_cast<char*>(base) + offset;
  return DB_OK;
}

extern "C" const char* db_last_error(db_handle_t h) { return h ? h->err : "null handle"; }

extern "C" int db_num_sms(db_handle_t h) { return h ? h->num_sms : 0; }

