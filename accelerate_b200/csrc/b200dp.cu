// b200dp.cu — host side of the C ABI declared in include/b200dp.h: context, symmetric segments
// (cuMemCreate + POSIX-fd export, or cudaMalloc + cudaIpc), multicast binding, fd passing, and the
// launch wrappers around the kernels in kernels.cuh.  No torch types anywhere; the library links
// cudart statically and resolves driver entry points at run time, so it loads (and exports every
// symbol) on a machine without a GPU driver.
#include <cuda.h>
#include <cuda_runtime.h>
#include <errno.h>
#include <poll.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <atomic>

#include "kernels.cuh"

using namespace b200dp;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define RT(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(B200DP_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,  \
                  __LINE__);                                                                       \
  } while (0)
#define DRV(call)                                                                                  \
  do {                                                                                             \
    CUresult r_ = (call);                                                                          \
    if (r_ != CUDA_SUCCESS) {                                                                      \
      const char* s_ = nullptr;                                                                    \
      if (g_drv.GetErrorString) g_drv.GetErrorString(r_, &s_);                                     \
      return fail(B200DP_ECUDA, "%s failed: %s (%s:%d)", #call, s_ ? s_ : "?", __FILE__, __LINE__); \
    }                                                                                              \
  } while (0)
#define KERNEL_OK(ctx)                                                                             \
  do {                                                                                             \
    cudaError_t e_ = cudaGetLastError();                                                           \
    if (e_ != cudaSuccess)                                                                         \
      return fail(B200DP_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),        \
                  __FILE__, __LINE__);                                                             \
    (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                                       \
  } while (0)

// ------------------------------------------------------------------------------------------------
// driver API, resolved lazily through the runtime (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
struct DriverApi {
  bool tried = false, ok = false;
  decltype(&cuGetErrorString) GetErrorString = nullptr;
  decltype(&cuDeviceGetAttribute) DeviceGetAttribute = nullptr;
  decltype(&cuMemCreate) MemCreate = nullptr;
  decltype(&cuMemRelease) MemRelease = nullptr;
  decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
  decltype(&cuMemAddressFree) MemAddressFree = nullptr;
  decltype(&cuMemMap) MemMap = nullptr;
  decltype(&cuMemUnmap) MemUnmap = nullptr;
  decltype(&cuMemSetAccess) MemSetAccess = nullptr;
  decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
  decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle = nullptr;
  decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle = nullptr;
  decltype(&cuMulticastCreate) MulticastCreate = nullptr;
  decltype(&cuMulticastAddDevice) MulticastAddDevice = nullptr;
  decltype(&cuMulticastBindMem) MulticastBindMem = nullptr;
  decltype(&cuMulticastGetGranularity) MulticastGetGranularity = nullptr;
  decltype(&cuMulticastUnbind) MulticastUnbind = nullptr;
};
static DriverApi g_drv;

template <typename F>
static bool resolve(const char* name, F* out) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr ||
      q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  *out = reinterpret_cast<F>(p);
  return true;
}
static bool load_driver() {
  if (g_drv.tried) return g_drv.ok;
  g_drv.tried = true;
  bool ok = true;
  ok &= resolve("cuGetErrorString", &g_drv.GetErrorString);
  ok &= resolve("cuDeviceGetAttribute", &g_drv.DeviceGetAttribute);
  ok &= resolve("cuMemCreate", &g_drv.MemCreate);
  ok &= resolve("cuMemRelease", &g_drv.MemRelease);
  ok &= resolve("cuMemAddressReserve", &g_drv.MemAddressReserve);
  ok &= resolve("cuMemAddressFree", &g_drv.MemAddressFree);
  ok &= resolve("cuMemMap", &g_drv.MemMap);
  ok &= resolve("cuMemUnmap", &g_drv.MemUnmap);
  ok &= resolve("cuMemSetAccess", &g_drv.MemSetAccess);
  ok &= resolve("cuMemGetAllocationGranularity", &g_drv.MemGetAllocationGranularity);
  ok &= resolve("cuMemExportToShareableHandle", &g_drv.MemExportToShareableHandle);
  ok &= resolve("cuMemImportFromShareableHandle", &g_drv.MemImportFromShareableHandle);
  g_drv.ok = ok;
  // multicast entry points are optional
  resolve("cuMulticastCreate", &g_drv.MulticastCreate);
  resolve("cuMulticastAddDevice", &g_drv.MulticastAddDevice);
  resolve("cuMulticastBindMem", &g_drv.MulticastBindMem);
  resolve("cuMulticastGetGranularity", &g_drv.MulticastGetGranularity);
  resolve("cuMulticastUnbind", &g_drv.MulticastUnbind);
  return ok;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
constexpr int kMaxSeg = 32;
constexpr int kStatAreas = B200DP_NUM_CHANNELS + 1;  // one partial area per channel + multi_stats
constexpr int kEventPool = 64;
constexpr int kPartialMax = 1280;  // CTAs that may leave a stats partial (>= 8 x 148)
constexpr int kPipeMaxChunks = 1024;

struct Segment {
  bool used = false;
  int kind = 0;
  size_t bytes = 0;  // padded
  void* local = nullptr;
  void* peer[B200DP_MAX_WORLD] = {};
  bool peer_mapped[B200DP_MAX_WORLD] = {};
  CUmemGenericAllocationHandle handle = 0;  // POSIXFD kind
  // multicast
  CUmemGenericAllocationHandle mc_handle = 0;
  void* mc_ptr = nullptr;
  bool mc_attached = false, mc_bound = false;
};

struct b200dp_ctx {
  int rank = 0, world = 1, device = 0;
  int sm_count = 148;
  bool vmm_ok = false, mc_ok = false;
  size_t gran = 2u << 20;
  Segment seg[kMaxSeg];
  int control_seg = -1;
  uint32_t* d_ctr = nullptr;      // [NUM_CHANNELS][MAX_BLOCKS]
  Record* d_partial = nullptr;    // [kStatAreas][MAX_BLOCKS]
  uint32_t* d_ticket = nullptr;   // [kStatAreas]
  uint32_t* d_pipe = nullptr;     // [2][kPipeMaxChunks] hand-off counters of reduce_pipelined_kernel
  int* h_status = nullptr;        // mapped pinned
  int* d_status = nullptr;        // device alias of h_status
  uint32_t* h_small = nullptr;    // mapped pinned [MAX_WORLD * SMALL_MAX/4]
  uint32_t* d_small = nullptr;
  int* h_small_flag = nullptr;
  int* d_small_flag = nullptr;
  int small_seq = 0;
  unsigned long long timeout_ns = 600ull * 1000 * 1000 * 1000;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t events[kEventPool] = {};
  int next_event = 0;
  std::atomic<unsigned long long> launches{0};
  int fdx_sock = -1;
};

extern "C" int b200dp_version(void) { return B200DP_VERSION; }
extern "C" const char* b200dp_last_error(void) { return g_err; }
extern "C" size_t b200dp_control_bytes(void) { return kControlBytes; }

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

extern "C" int b200dp_init(int rank, int world, int device, b200dp_ctx** out) {
  if (!out) return fail(B200DP_EINVAL, "out is NULL");
  if (world < 1 || world > B200DP_MAX_WORLD || rank < 0 || rank >= world)
    return fail(B200DP_EINVAL, "bad rank/world %d/%d (max world %d)", rank, world, B200DP_MAX_WORLD);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(B200DP_ECUDA, "no CUDA device available (%s): the b200dp data path has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= ndev) return fail(B200DP_EINVAL, "device %d out of range (%d devices)", device, ndev);
  RT(cudaSetDevice(device));
  RT(cudaFree(0));
  b200dp_ctx* c = new b200dp_ctx();
  c->rank = rank;
  c->world = world;
  c->device = device;
  cudaDeviceProp prop;
  RT(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  if (prop.major < 10)
    fprintf(stderr, "[b200dp] warning: device %d is sm_%d%d; this library is built for sm_100a only\n", device,
            prop.major, prop.minor);

  if (load_driver()) {
    int v = 0;
    if (g_drv.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, device) == CUDA_SUCCESS && v) {
      int fdok = 0;
      g_drv.DeviceGetAttribute(&fdok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, device);
      c->vmm_ok = fdok != 0;
    }
    int mc = 0;
    if (g_drv.MulticastCreate && g_drv.MulticastBindMem &&
        g_drv.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device) == CUDA_SUCCESS)
      c->mc_ok = mc != 0 && c->vmm_ok;
    if (c->vmm_ok) {
      CUmemAllocationProp ap = {};
      ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
      ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      ap.location.id = device;
      ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      size_t g = 0;
      if (g_drv.MemGetAllocationGranularity(&g, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && g)
        c->gran = g;
    }
  }
  const char* env = getenv("B200DP_TIMEOUT_MS");
  if (env && atoll(env) > 0) c->timeout_ns = (unsigned long long)atoll(env) * 1000000ull;

  RT(cudaMalloc(&c->d_ctr, sizeof(uint32_t) * B200DP_NUM_CHANNELS * B200DP_MAX_BLOCKS));
  RT(cudaMemset(c->d_ctr, 0, sizeof(uint32_t) * B200DP_NUM_CHANNELS * B200DP_MAX_BLOCKS));
  RT(cudaMalloc(&c->d_partial, sizeof(Record) * kStatAreas * kPartialMax));
  RT(cudaMemset(c->d_partial, 0, sizeof(Record) * kStatAreas * kPartialMax));
  RT(cudaMalloc(&c->d_ticket, sizeof(uint32_t) * kStatAreas));
  RT(cudaMemset(c->d_ticket, 0, sizeof(uint32_t) * kStatAreas));
  RT(cudaMalloc(&c->d_pipe, sizeof(uint32_t) * 2 * kPipeMaxChunks));
  RT(cudaMemset(c->d_pipe, 0, sizeof(uint32_t) * 2 * kPipeMaxChunks));
  RT(cudaHostAlloc(&c->h_status, 64, cudaHostAllocMapped));
  *c->h_status = 0;
  RT(cudaHostGetDevicePointer(&c->d_status, c->h_status, 0));
  RT(cudaHostAlloc(&c->h_small, B200DP_MAX_WORLD * B200DP_SMALL_MAX + 64, cudaHostAllocMapped));
  RT(cudaHostGetDevicePointer(&c->d_small, c->h_small, 0));
  c->h_small_flag = reinterpret_cast<int*>(reinterpret_cast<char*>(c->h_small) + B200DP_MAX_WORLD * B200DP_SMALL_MAX);
  c->d_small_flag = reinterpret_cast<int*>(reinterpret_cast<char*>(c->d_small) + B200DP_MAX_WORLD * B200DP_SMALL_MAX);
  *c->h_small_flag = 0;
  int lo = 0, hi = 0;
  RT(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  RT(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
  for (int i = 0; i < kEventPool; ++i) RT(cudaEventCreateWithFlags(&c->events[i], cudaEventDisableTiming));
  RT(cudaDeviceSynchronize());
  *out = c;
  return B200DP_OK;
}

static int seg_release(b200dp_ctx* c, Segment& s);

extern "C" int b200dp_finalize(b200dp_ctx* c) {
  if (!c) return B200DP_OK;
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < kMaxSeg; ++i)
    if (c->seg[i].used) seg_release(c, c->seg[i]);
  if (c->fdx_sock >= 0) close(c->fdx_sock);
  for (int i = 0; i < kEventPool; ++i)
    if (c->events[i]) cudaEventDestroy(c->events[i]);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  cudaFree(c->d_ctr);
  cudaFree(c->d_partial);
  cudaFree(c->d_ticket);
  cudaFree(c->d_pipe);
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->h_small) cudaFreeHost(c->h_small);
  cudaGetLastError();
  delete c;
  return B200DP_OK;
}

extern "C" int b200dp_query(b200dp_ctx* c, int key, long long* out) {
  if (!c || !out) return fail(B200DP_EINVAL, "null argument");
  switch (key) {
    case B200DP_Q_LAUNCHES: *out = (long long)c->launches.load(); break;
    case B200DP_Q_MULTICAST_OK: *out = c->mc_ok; break;
    case B200DP_Q_VMM_OK: *out = c->vmm_ok; break;
    case B200DP_Q_STATUS: *out = *reinterpret_cast<volatile int*>(c->h_status); break;
    case B200DP_Q_SM_COUNT: *out = c->sm_count; break;
    case B200DP_Q_COMM_STREAM: *out = (long long)(uintptr_t)c->comm_stream; break;
    default: return fail(B200DP_EINVAL, "unknown query key %d", key);
  }
  return B200DP_OK;
}
extern "C" int b200dp_set_timeout_ms(b200dp_ctx* c, long long ms) {
  if (!c || ms <= 0) return fail(B200DP_EINVAL, "bad timeout");
  c->timeout_ns = (unsigned long long)ms * 1000000ull;
  return B200DP_OK;
}
extern "C" int b200dp_reset_status(b200dp_ctx* c) {
  if (!c) return fail(B200DP_EINVAL, "null ctx");
  *reinterpret_cast<volatile int*>(c->h_status) = 0;
  return B200DP_OK;
}

// ------------------------------------------------------------------------------------------------
// segments
// ------------------------------------------------------------------------------------------------
static size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }

static int set_rw_access(b200dp_ctx* c, void* ptr, size_t bytes) {
  CUmemAccessDesc acc = {};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = c->device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  DRV(g_drv.MemSetAccess((CUdeviceptr)ptr, bytes, &acc, 1));
  return B200DP_OK;
}

extern "C" int b200dp_seg_create(b200dp_ctx* c, size_t bytes, int prefer_kind, int* seg, void** local_ptr,
                                 size_t* padded_bytes, int* kind) {
  if (!c || !seg || !local_ptr || bytes == 0) return fail(B200DP_EINVAL, "bad arguments to seg_create");
  DeviceGuard g(c->device);
  int id = -1;
  for (int i = 0; i < kMaxSeg; ++i)
    if (!c->seg[i].used) { id = i; break; }
  if (id < 0) return fail(B200DP_ESTATE, "out of segment slots (%d)", kMaxSeg);
  Segment& s = c->seg[id];
  s = Segment();
  const bool want_fd = prefer_kind == B200DP_HANDLE_POSIXFD && c->vmm_ok;
  if (want_fd) {
    size_t gran = c->gran;
    if (c->mc_ok && g_drv.MulticastGetGranularity) {  // make every VMM segment multicast-bindable
      CUmulticastObjectProp mp = {};
      mp.numDevices = c->world > 1 ? c->world : 2;
      mp.size = round_up(bytes, gran);
      mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      size_t mg = 0;
      if (g_drv.MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > gran)
        gran = mg;
    }
    s.bytes = round_up(bytes, gran);
    CUmemAllocationProp ap = {};
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = c->device;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    DRV(g_drv.MemCreate(&s.handle, s.bytes, &ap, 0));
    CUdeviceptr va = 0;
    DRV(g_drv.MemAddressReserve(&va, s.bytes, gran, 0, 0));
    DRV(g_drv.MemMap(va, s.bytes, 0, s.handle, 0));
    s.local = (void*)va;
    int rc = set_rw_access(c, s.local, s.bytes);
    if (rc) return rc;
    s.kind = B200DP_HANDLE_POSIXFD;
  } else {
    s.bytes = round_up(bytes, 2u << 20);
    RT(cudaMalloc(&s.local, s.bytes));
    s.kind = B200DP_HANDLE_CUDAIPC;
  }
  RT(cudaMemset(s.local, 0, s.bytes));
  RT(cudaDeviceSynchronize());
  s.peer[c->rank] = s.local;
  s.peer_mapped[c->rank] = true;
  s.used = true;
  *seg = id;
  *local_ptr = s.local;
  if (padded_bytes) *padded_bytes = s.bytes;
  if (kind) *kind = s.kind;
  return B200DP_OK;
}

static int get_seg(b200dp_ctx* c, int seg, Segment** out) {
  if (!c || seg < 0 || seg >= kMaxSeg || !c->seg[seg].used) return fail(B200DP_EINVAL, "bad segment id %d", seg);
  *out = &c->seg[seg];
  return B200DP_OK;
}

extern "C" int b200dp_seg_export(b200dp_ctx* c, int seg, unsigned char handle[64], int* fd) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  DeviceGuard g(c->device);
  if (s->kind == B200DP_HANDLE_POSIXFD) {
    if (!fd) return fail(B200DP_EINVAL, "fd out-pointer required for POSIXFD segments");
    int f = -1;
    DRV(g_drv.MemExportToShareableHandle(&f, s->handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *fd = f;
    if (handle) memset(handle, 0, 64);
  } else {
    if (!handle) return fail(B200DP_EINVAL, "handle buffer required for CUDAIPC segments");
    cudaIpcMemHandle_t h;
    RT(cudaIpcGetMemHandle(&h, s->local));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle, &h, 64);
    if (fd) *fd = -1;
  }
  return B200DP_OK;
}

extern "C" int b200dp_seg_import(b200dp_ctx* c, int seg, int peer, int kind, const unsigned char handle[64], int fd) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if (peer < 0 || peer >= c->world) return fail(B200DP_EINVAL, "bad peer %d", peer);
  if (peer == c->rank) return B200DP_OK;
  if (kind != s->kind) return fail(B200DP_EINVAL, "peer %d exported kind %d, local segment is kind %d", peer, kind, s->kind);
  DeviceGuard g(c->device);
  if (kind == B200DP_HANDLE_POSIXFD) {
    CUmemGenericAllocationHandle h;
    DRV(g_drv.MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    CUdeviceptr va = 0;
    DRV(g_drv.MemAddressReserve(&va, s->bytes, c->gran, 0, 0));
    DRV(g_drv.MemMap(va, s->bytes, 0, h, 0));
    rc = set_rw_access(c, (void*)va, s->bytes);
    if (rc) return rc;
    g_drv.MemRelease(h);  // the mapping keeps the allocation alive
    close(fd);
    s->peer[peer] = (void*)va;
  } else {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    RT(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer[peer] = p;
  }
  s->peer_mapped[peer] = true;
  return B200DP_OK;
}

extern "C" int b200dp_seg_peer_ptr(b200dp_ctx* c, int seg, int peer, void** ptr) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if (peer < 0 || peer >= c->world || !s->peer_mapped[peer]) return fail(B200DP_ESTATE, "peer %d of segment %d not mapped", peer, seg);
  *ptr = s->peer[peer];
  return B200DP_OK;
}

static int seg_release(b200dp_ctx* c, Segment& s) {
  if (s.mc_ptr) {
    g_drv.MemUnmap((CUdeviceptr)s.mc_ptr, s.bytes);
    g_drv.MemAddressFree((CUdeviceptr)s.mc_ptr, s.bytes);
  }
  if (s.mc_handle) g_drv.MemRelease(s.mc_handle);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank || !s.peer_mapped[r]) continue;
    if (s.kind == B200DP_HANDLE_POSIXFD) {
      g_drv.MemUnmap((CUdeviceptr)s.peer[r], s.bytes);
      g_drv.MemAddressFree((CUdeviceptr)s.peer[r], s.bytes);
    } else {
      cudaIpcCloseMemHandle(s.peer[r]);
    }
  }
  if (s.kind == B200DP_HANDLE_POSIXFD) {
    g_drv.MemUnmap((CUdeviceptr)s.local, s.bytes);
    g_drv.MemAddressFree((CUdeviceptr)s.local, s.bytes);
    g_drv.MemRelease(s.handle);
  } else {
    cudaFree(s.local);
  }
  cudaGetLastError();
  s = Segment();
  return B200DP_OK;
}

extern "C" int b200dp_seg_destroy(b200dp_ctx* c, int seg) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  if (c->control_seg == seg) c->control_seg = -1;
  return seg_release(c, *s);
}

// ---- multicast ----
extern "C" int b200dp_mc_create(b200dp_ctx* c, int seg, int* fd) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if (!c->mc_ok) return fail(B200DP_ENOSUP, "multicast objects not supported on this device/driver");
  if (s->kind != B200DP_HANDLE_POSIXFD) return fail(B200DP_ESTATE, "multicast needs a POSIXFD (cuMemCreate) segment");
  DeviceGuard g(c->device);
  CUmulticastObjectProp mp = {};
  mp.numDevices = c->world;
  mp.size = s->bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  DRV(g_drv.MulticastCreate(&s->mc_handle, &mp));
  int f = -1;
  DRV(g_drv.MemExportToShareableHandle(&f, s->mc_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *fd = f;
  return B200DP_OK;
}
extern "C" int b200dp_mc_attach(b200dp_ctx* c, int seg, int fd) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if (!c->mc_ok) return fail(B200DP_ENOSUP, "multicast objects not supported on this device/driver");
  DeviceGuard g(c->device);
  if (!s->mc_handle) {  // non-creator ranks import the object
    DRV(g_drv.MemImportFromShareableHandle(&s->mc_handle, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  }
  if (fd >= 0) close(fd);
  DRV(g_drv.MulticastAddDevice(s->mc_handle, c->device));
  s->mc_attached = true;
  return B200DP_OK;
}
extern "C" int b200dp_mc_bind(b200dp_ctx* c, int seg) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if (!s->mc_attached) return fail(B200DP_ESTATE, "mc_attach must precede mc_bind");
  DeviceGuard g(c->device);
  DRV(g_drv.MulticastBindMem(s->mc_handle, 0, s->handle, 0, s->bytes, 0));
  CUdeviceptr va = 0;
  DRV(g_drv.MemAddressReserve(&va, s->bytes, c->gran, 0, 0));
  DRV(g_drv.MemMap(va, s->bytes, 0, s->mc_handle, 0));
  rc = set_rw_access(c, (void*)va, s->bytes);
  if (rc) return rc;
  s->mc_ptr = (void*)va;
  s->mc_bound = true;
  return B200DP_OK;
}
extern "C" int b200dp_seg_mc_ptr(b200dp_ctx* c, int seg, void** ptr) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  *ptr = s->mc_bound ? s->mc_ptr : nullptr;
  return B200DP_OK;
}

extern "C" int b200dp_set_control(b200dp_ctx* c, int seg) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if (s->bytes < kControlBytes) return fail(B200DP_EINVAL, "control segment too small: %zu < %zu", s->bytes, kControlBytes);
  for (int r = 0; r < c->world; ++r)
    if (!s->peer_mapped[r]) return fail(B200DP_ESTATE, "control segment: peer %d not mapped", r);
  c->control_seg = seg;
  return B200DP_OK;
}

// ------------------------------------------------------------------------------------------------
// fd passing (SCM_RIGHTS over abstract-namespace unix datagram sockets)
// ------------------------------------------------------------------------------------------------
static socklen_t make_addr(const char* name, sockaddr_un* a) {
  memset(a, 0, sizeof(*a));
  a->sun_family = AF_UNIX;
  size_t n = strlen(name);
  if (n > sizeof(a->sun_path) - 2) n = sizeof(a->sun_path) - 2;
  a->sun_path[0] = '\0';  // abstract namespace: no filesystem entry, vanishes with the socket
  memcpy(a->sun_path + 1, name, n);
  return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}
extern "C" int b200dp_fdx_open(b200dp_ctx* c, const char* name) {
  if (!c || !name) return fail(B200DP_EINVAL, "null argument");
  if (c->fdx_sock >= 0) close(c->fdx_sock);
  int s = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  if (s < 0) return fail(B200DP_ESYS, "socket: %s", strerror(errno));
  sockaddr_un a;
  socklen_t len = make_addr(name, &a);
  if (bind(s, (sockaddr*)&a, len) != 0) {
    int e = errno;
    close(s);
    return fail(B200DP_ESYS, "bind(@%s): %s", name, strerror(e));
  }
  c->fdx_sock = s;
  return B200DP_OK;
}
extern "C" int b200dp_fdx_send(b200dp_ctx* c, const char* peer_name, int fd, int tag) {
  if (!c || c->fdx_sock < 0) return fail(B200DP_ESTATE, "fdx socket not open");
  sockaddr_un a;
  socklen_t len = make_addr(peer_name, &a);
  int payload[2] = {tag, c->rank};
  iovec iov = {payload, sizeof(payload)};
  char cbuf[CMSG_SPACE(sizeof(int))];
  memset(cbuf, 0, sizeof(cbuf));
  msghdr msg = {};
  msg.msg_name = &a;
  msg.msg_namelen = len;
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = cbuf;
  msg.msg_controllen = sizeof(cbuf);
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  cm->cmsg_level = SOL_SOCKET;
  cm->cmsg_type = SCM_RIGHTS;
  cm->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(cm), &fd, sizeof(int));
  for (int attempt = 0; attempt < 200; ++attempt) {
    if (sendmsg(c->fdx_sock, &msg, 0) >= 0) return B200DP_OK;
    if (errno != ECONNREFUSED && errno != ENOENT && errno != EAGAIN && errno != ENOBUFS) break;
    usleep(10000);  // peer not bound yet / queue full
  }
  return fail(B200DP_ESYS, "sendmsg(@%s): %s", peer_name, strerror(errno));
}
extern "C" int b200dp_fdx_recv(b200dp_ctx* c, int* fd, int* tag, int timeout_ms) {
  if (!c || c->fdx_sock < 0 || !fd || !tag) return fail(B200DP_ESTATE, "fdx socket not open");
  pollfd p = {c->fdx_sock, POLLIN, 0};
  int pr = poll(&p, 1, timeout_ms);
  if (pr == 0) return fail(B200DP_ETIMEOUT, "fdx_recv timed out after %d ms", timeout_ms);
  if (pr < 0) return fail(B200DP_ESYS, "poll: %s", strerror(errno));
  int payload[2] = {0, 0};
  iovec iov = {payload, sizeof(payload)};
  char cbuf[CMSG_SPACE(sizeof(int))];
  msghdr msg = {};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = cbuf;
  msg.msg_controllen = sizeof(cbuf);
  if (recvmsg(c->fdx_sock, &msg, MSG_CMSG_CLOEXEC) < 0) return fail(B200DP_ESYS, "recvmsg: %s", strerror(errno));
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  if (!cm || cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS)
    return fail(B200DP_ESYS, "recvmsg: no fd in message");
  memcpy(fd, CMSG_DATA(cm), sizeof(int));
  *tag = payload[0];
  return B200DP_OK;
}
extern "C" int b200dp_fdx_close(b200dp_ctx* c) {
  if (c && c->fdx_sock >= 0) {
    close(c->fdx_sock);
    c->fdx_sock = -1;
  }
  return B200DP_OK;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
extern "C" int b200dp_stream_wait(b200dp_ctx* c, void* waiter, void* signaller) {
  if (!c) return fail(B200DP_EINVAL, "null ctx");
  if (waiter == signaller) return B200DP_OK;
  DeviceGuard g(c->device);
  cudaEvent_t ev = c->events[c->next_event];
  c->next_event = (c->next_event + 1) % kEventPool;
  RT(cudaEventRecord(ev, (cudaStream_t)signaller));
  RT(cudaStreamWaitEvent((cudaStream_t)waiter, ev, 0));
  return B200DP_OK;
}

static int make_barrier(b200dp_ctx* c, int channel, BarrierCtx* b) {
  if (c->control_seg < 0) return fail(B200DP_ESTATE, "control segment not set (b200dp_set_control)");
  if (channel < 0 || channel >= B200DP_NUM_CHANNELS) return fail(B200DP_EINVAL, "bad channel %d", channel);
  Segment& s = c->seg[c->control_seg];
  for (int r = 0; r < B200DP_MAX_WORLD; ++r) b->pads.p[r] = r < c->world ? s.peer[r] : nullptr;
  b->ctr = c->d_ctr + size_t(channel) * B200DP_MAX_BLOCKS;
  b->status = c->d_status;
  b->timeout_ns = c->timeout_ns;
  b->rank = c->rank;
  b->world = c->world;
  b->channel = channel;
  return B200DP_OK;
}
static void fill_peers(b200dp_ctx* c, const Segment& s, Peers* p) {
  for (int r = 0; r < B200DP_MAX_WORLD; ++r) p->p[r] = r < c->world ? s.peer[r] : nullptr;
}
static int all_mapped(b200dp_ctx* c, const Segment& s) {
  for (int r = 0; r < c->world; ++r)
    if (!s.peer_mapped[r]) return fail(B200DP_ESTATE, "segment peer %d not mapped", r);
  return B200DP_OK;
}
static int make_sink(b200dp_ctx* c, int area, int slot, int world, StatsSink* k) {
  if (slot >= B200DP_MAX_SLOTS) return fail(B200DP_EINVAL, "slot %d out of range", slot);
  if (slot >= 0 && c->control_seg < 0) return fail(B200DP_ESTATE, "control segment not set");
  k->partial = c->d_partial + size_t(area) * kPartialMax;
  k->ticket = c->d_ticket + area;
  if (c->control_seg >= 0) fill_peers(c, c->seg[c->control_seg], &k->control);
  k->slot = slot;
  k->rank = c->rank;
  k->world = world;
  return B200DP_OK;
}
static size_t dtype_size(int dtype) {
  switch (dtype) {
    case B200DP_F32: case B200DP_I32: return 4;
    case B200DP_BF16: case B200DP_F16: return 2;
    case B200DP_F64: case B200DP_I64: return 8;
    case B200DP_U8: return 1;
    default: return 0;
  }
}
static int clamp_blocks(b200dp_ctx* c, size_t work_items, size_t per_block, int max_blocks) {
  size_t want = (work_items + per_block - 1) / per_block;
  int cap = max_blocks > 0 ? max_blocks : c->sm_count;
  if (cap > c->sm_count) cap = c->sm_count;  // all CTAs of a flag-synchronised kernel must be co-resident
  if (cap > B200DP_MAX_BLOCKS) cap = B200DP_MAX_BLOCKS;
  if (want < 1) want = 1;
  return int(want < size_t(cap) ? want : size_t(cap));
}

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
template <typename T>
static void launch_allreduce(const AllreduceArgs& a, int world, bool nvls, int grid, cudaStream_t st) {
  if (nvls) {
    static const int unroll = [] {
      const char* e = getenv("B200DP_NVLS_UNROLL");
      return (e && atoi(e) == 8) ? 8 : 4;
    }();
    if (unroll == 8) allreduce_inplace_kernel<T, 8, true><<<grid, kThreads, 0, st>>>(a);
    else allreduce_inplace_kernel<T, 4, true><<<grid, kThreads, 0, st>>>(a);
    return;
  }
  switch (world) {  // W ranks known at compile time: all W loads of a vector are issued back to back
    case 1: allreduce_inplace_kernel<T, 1, false><<<grid, kThreads, 0, st>>>(a); break;
    case 2: allreduce_inplace_kernel<T, 2, false><<<grid, kThreads, 0, st>>>(a); break;
    case 3: allreduce_inplace_kernel<T, 3, false><<<grid, kThreads, 0, st>>>(a); break;
    case 4: allreduce_inplace_kernel<T, 4, false><<<grid, kThreads, 0, st>>>(a); break;
    case 5: allreduce_inplace_kernel<T, 5, false><<<grid, kThreads, 0, st>>>(a); break;
    case 6: allreduce_inplace_kernel<T, 6, false><<<grid, kThreads, 0, st>>>(a); break;
    case 7: allreduce_inplace_kernel<T, 7, false><<<grid, kThreads, 0, st>>>(a); break;
    case 8: allreduce_inplace_kernel<T, 8, false><<<grid, kThreads, 0, st>>>(a); break;
    default: allreduce_inplace_kernel<T, 0, false><<<grid, kThreads, 0, st>>>(a); break;  // 9..16 ranks
  }
}

extern "C" int b200dp_allreduce_bucket(b200dp_ctx* c, int seg, size_t elem_off, size_t count, int dtype,
                                       float prescale, int slot, int flags, int algo, int channel, int max_blocks,
                                       void* after_stream, void* stream) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if ((rc = all_mapped(c, *s))) return rc;
  if (dtype != B200DP_F32 && dtype != B200DP_BF16 && dtype != B200DP_F16)
    return fail(B200DP_EINVAL, "allreduce_bucket supports f32/bf16/f16, got dtype %d", dtype);
  const size_t es = dtype_size(dtype);
  const size_t byte_off = elem_off * es, nbytes = count * es;
  if ((byte_off & 15) || (nbytes & 15)) return fail(B200DP_EINVAL, "bucket offset/size must be 16-byte multiples (off %zu, bytes %zu)", byte_off, nbytes);
  if (byte_off + nbytes > s->bytes) return fail(B200DP_EINVAL, "bucket [%zu, +%zu) exceeds segment (%zu)", byte_off, nbytes, s->bytes);
  if (count == 0) return B200DP_OK;
  bool nvls = false;
  if (algo == B200DP_ALGO_NVLS) {
    if (!s->mc_bound) return fail(B200DP_ESTATE, "NVLS requested but segment %d has no multicast binding", seg);
    nvls = true;
  } else if (algo == B200DP_ALGO_AUTO) {
    nvls = s->mc_bound && c->world > 1;
  } else if (algo != B200DP_ALGO_TWOSHOT) {
    return fail(B200DP_EINVAL, "bad algo %d for in-place allreduce", algo);
  }
  DeviceGuard g(c->device);
  AllreduceArgs a;
  if ((rc = make_barrier(c, channel, &a.bar))) return rc;
  fill_peers(c, *s, &a.bufs);
  a.mc = reinterpret_cast<char*>(s->mc_ptr);
  a.byte_off = byte_off;
  a.nvec = nbytes / 16;
  a.prescale = prescale;
  a.flags = flags;
  const int want_slot = (flags & B200DP_F_STATS) ? slot : -1;
  if ((flags & B200DP_F_STATS) && slot < 0) return fail(B200DP_EINVAL, "B200DP_F_STATS needs a slot");
  if ((rc = make_sink(c, channel, want_slot, c->world, &a.sink))) return rc;
  const size_t per_rank = (a.nvec + c->world - 1) / c->world;
  const int grid = clamp_blocks(c, per_rank, size_t(kThreads) * 4, max_blocks);
  cudaStream_t st = (cudaStream_t)stream;
  if (after_stream != B200DP_NO_STREAM && after_stream != stream) {
    if ((rc = b200dp_stream_wait(c, stream, after_stream))) return rc;
  }
  switch (dtype) {
    case B200DP_F32: launch_allreduce<float>(a, c->world, nvls, grid, st); break;
    case B200DP_BF16: launch_allreduce<__nv_bfloat16>(a, c->world, nvls, grid, st); break;
    default: launch_allreduce<__half>(a, c->world, nvls, grid, st); break;
  }
  KERNEL_OK(c);
  return B200DP_OK;
}

extern "C" int b200dp_fence(b200dp_ctx* c, int channel, void* stream) {
  if (!c) return fail(B200DP_EINVAL, "null ctx");
  DeviceGuard g(c->device);
  BarrierCtx b;
  int rc = make_barrier(c, channel, &b);
  if (rc) return rc;
  fence_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(b);
  KERNEL_OK(c);
  return B200DP_OK;
}

extern "C" int b200dp_allreduce_oneshot(b200dp_ctx* c, int seg, size_t byte_off, size_t count, int dtype,
                                        double divisor, void* dst, int channel, int max_blocks, void* stream) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if ((rc = all_mapped(c, *s))) return rc;
  const size_t es = dtype_size(dtype);
  if (!es) return fail(B200DP_EINVAL, "bad dtype %d", dtype);
  if (byte_off & 15) return fail(B200DP_EINVAL, "staging offset must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(dst) & 15) return fail(B200DP_EINVAL, "dst must be 16-byte aligned");
  if (byte_off + round_up(count * es, 16) > s->bytes) return fail(B200DP_EINVAL, "staging range exceeds segment");
  if (count == 0) return B200DP_OK;
  DeviceGuard g(c->device);
  OneshotArgs a;
  if ((rc = make_barrier(c, channel, &a.bar))) return rc;
  fill_peers(c, *s, &a.bufs);
  a.byte_off = byte_off;
  a.count = count;
  a.divisor = divisor;
  a.dst = dst;
  const int grid = clamp_blocks(c, count * es / 16, kThreads, max_blocks);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case B200DP_F32: allreduce_oneshot_kernel<float><<<grid, kThreads, 0, st>>>(a); break;
    case B200DP_BF16: allreduce_oneshot_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(a); break;
    case B200DP_F16: allreduce_oneshot_kernel<__half><<<grid, kThreads, 0, st>>>(a); break;
    case B200DP_F64: allreduce_oneshot_kernel<double><<<grid, kThreads, 0, st>>>(a); break;
    case B200DP_I32: allreduce_oneshot_kernel<int><<<grid, kThreads, 0, st>>>(a); break;
    case B200DP_I64: allreduce_oneshot_kernel<long long><<<grid, kThreads, 0, st>>>(a); break;
    default: allreduce_oneshot_kernel<unsigned char><<<grid, kThreads, 0, st>>>(a); break;
  }
  KERNEL_OK(c);
  return B200DP_OK;
}

// ------------------------------------------------------------------------------------------------
// K6
// ------------------------------------------------------------------------------------------------
extern "C" int b200dp_finalize_norm(b200dp_ctx* c, int slot0, int nslots, int all_ranks, double norm_type,
                                    float max_norm, const float* scale_ptr, float* out, void* stream) {
  if (!c || !out) return fail(B200DP_EINVAL, "null argument");
  if (c->control_seg < 0) return fail(B200DP_ESTATE, "control segment not set");
  if (slot0 < 0 || nslots < 0 || slot0 + nslots > B200DP_MAX_SLOTS) return fail(B200DP_EINVAL, "bad slot range [%d,+%d)", slot0, nslots);
  const bool is_inf = norm_type > 1e300;
  if (!is_inf && norm_type != 2.0) return fail(B200DP_ENOSUP, "norm_type %g: only 2 and inf are fused", norm_type);
  DeviceGuard g(c->device);
  const char* ctl = reinterpret_cast<const char*>(c->seg[c->control_seg].local);
  finalize_norm_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ctl, slot0, nslots, all_ranks ? 0 : c->rank,
                                                          all_ranks ? c->world : 1, is_inf ? 1 : 0, max_norm,
                                                          scale_ptr, out);
  KERNEL_OK(c);
  return B200DP_OK;
}

static int stream_grid(b200dp_ctx* c, size_t bytes, int max_blocks) {
  // HBM streaming: 8 CTAs of 256 threads per SM saturate the memory system; never more CTAs than work
  size_t want = (bytes + size_t(kStreamThreads) * 64 - 1) / (size_t(kStreamThreads) * 64);
  size_t cap = max_blocks > 0 ? size_t(max_blocks) : size_t(c->sm_count) * 8;
  if (want < 1) want = 1;
  return int(want < cap ? want : cap);
}

extern "C" int b200dp_scale_flat(b200dp_ctx* c, void* ptr, size_t count, int dtype, const float* mult_ptr,
                                 int max_blocks, void* stream) {
  if (!c || !ptr || !mult_ptr) return fail(B200DP_EINVAL, "null argument");
  if (count == 0) return B200DP_OK;
  DeviceGuard g(c->device);
  const int grid = stream_grid(c, count * dtype_size(dtype), max_blocks);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case B200DP_F32: scale_flat_kernel<float><<<grid, kStreamThreads, 0, st>>>((float*)ptr, count, mult_ptr); break;
    case B200DP_BF16: scale_flat_kernel<__nv_bfloat16><<<grid, kStreamThreads, 0, st>>>((__nv_bfloat16*)ptr, count, mult_ptr); break;
    case B200DP_F16: scale_flat_kernel<__half><<<grid, kStreamThreads, 0, st>>>((__half*)ptr, count, mult_ptr); break;
    default: return fail(B200DP_EINVAL, "scale_flat supports f32/bf16/f16");
  }
  KERNEL_OK(c);
  return B200DP_OK;
}

extern "C" int b200dp_clamp_flat(b200dp_ctx* c, void* ptr, size_t count, int dtype, float v, int max_blocks,
                                 void* stream) {
  if (!c || !ptr) return fail(B200DP_EINVAL, "null argument");
  if (count == 0) return B200DP_OK;
  DeviceGuard g(c->device);
  const int grid = stream_grid(c, count * dtype_size(dtype), max_blocks);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case B200DP_F32: clamp_flat_kernel<float><<<grid, kStreamThreads, 0, st>>>((float*)ptr, count, v); break;
    case B200DP_BF16: clamp_flat_kernel<__nv_bfloat16><<<grid, kStreamThreads, 0, st>>>((__nv_bfloat16*)ptr, count, v); break;
    case B200DP_F16: clamp_flat_kernel<__half><<<grid, kStreamThreads, 0, st>>>((__half*)ptr, count, v); break;
    default: return fail(B200DP_EINVAL, "clamp_flat supports f32/bf16/f16");
  }
  KERNEL_OK(c);
  return B200DP_OK;
}

extern "C" int b200dp_cast(b200dp_ctx* c, const void* src, int src_dtype, void* dst, int dst_dtype, size_t count,
                           float divisor, int max_blocks, void* stream) {
  if (!c || !src || !dst) return fail(B200DP_EINVAL, "null argument");
  if (count == 0) return B200DP_OK;
  DeviceGuard g(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = stream_grid(c, count * 6, max_blocks);
  if (src_dtype == B200DP_F32 && dst_dtype == B200DP_BF16)
    cast_kernel<float, __nv_bfloat16><<<grid, kStreamThreads, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, count, divisor);
  else if (src_dtype == B200DP_F32 && dst_dtype == B200DP_F16)
    cast_kernel<float, __half><<<grid, kStreamThreads, 0, st>>>((const float*)src, (__half*)dst, count, divisor);
  else if (src_dtype == B200DP_BF16 && dst_dtype == B200DP_F32)
    cast_kernel<__nv_bfloat16, float><<<grid, kStreamThreads, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, count, divisor);
  else if (src_dtype == B200DP_F16 && dst_dtype == B200DP_F32)
    cast_kernel<__half, float><<<grid, kStreamThreads, 0, st>>>((const __half*)src, (float*)dst, count, divisor);
  else
    return fail(B200DP_EINVAL, "cast %d -> %d not supported (f32 <-> bf16/f16 only)", src_dtype, dst_dtype);
  KERNEL_OK(c);
  return B200DP_OK;
}

// ------------------------------------------------------------------------------------------------
// K5
// ------------------------------------------------------------------------------------------------
static int build_mt(const void* const* ptrs, const size_t* numels, int n, MTTable* t) {
  uint32_t acc = 0;
  for (int i = 0; i < n; ++i) {
    t->ptr[i] = const_cast<void*>(ptrs[i]);
    t->numel[i] = numels[i];
    t->chunk_start[i] = acc;
    acc += uint32_t((numels[i] + kMTChunk - 1) / kMTChunk);
  }
  t->chunk_start[n] = acc;
  t->n = n;
  return int(acc);
}

extern "C" int b200dp_multi_stats(b200dp_ctx* c, const void* const* ptrs, const size_t* numels, int n, int dtype,
                                  int slot, void* stream) {
  if (!c || (n > 0 && (!ptrs || !numels))) return fail(B200DP_EINVAL, "null argument");
  if (dtype != B200DP_F32 && dtype != B200DP_BF16 && dtype != B200DP_F16) return fail(B200DP_EINVAL, "multi_stats supports f32/bf16/f16");
  if (slot < 0) return fail(B200DP_EINVAL, "slot required");
  DeviceGuard g(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  static thread_local MTTable t;
  int launches = (n + kMTMax - 1) / kMTMax;
  if (launches == 0) launches = 1;  // n == 0: publish a zero record
  for (int l = 0; l < launches; ++l) {
    const int lo = l * kMTMax;
    const int cnt = (n - lo) < kMTMax ? (n - lo) : kMTMax;
    const int chunks = build_mt(ptrs + lo, numels + lo, cnt > 0 ? cnt : 0, &t);
    StatsSink local;
    int rc = make_sink(c, B200DP_NUM_CHANNELS, slot + l, 1, &local);
    if (rc) return rc;
    // world = 1 view: publish only into the LOCAL control segment, column `rank` of the slot row
    for (int r = 0; r < B200DP_MAX_WORLD; ++r) local.control.p[r] = nullptr;
    local.control.p[0] = c->seg[c->control_seg].local;
    int grid = chunks < c->sm_count * 8 ? chunks : c->sm_count * 8;
    if (grid < 1) grid = 1;
    if (grid > kPartialMax) grid = kPartialMax;
    switch (dtype) {
      case B200DP_F32: multi_stats_kernel<float><<<grid, kStreamThreads, 0, st>>>(t, local); break;
      case B200DP_BF16: multi_stats_kernel<__nv_bfloat16><<<grid, kStreamThreads, 0, st>>>(t, local); break;
      default: multi_stats_kernel<__half><<<grid, kStreamThreads, 0, st>>>(t, local); break;
    }
    KERNEL_OK(c);
  }
  return B200DP_OK;
}

extern "C" int b200dp_multi_scale(b200dp_ctx* c, void* const* ptrs, const size_t* numels, int n, int dtype,
                                  const float* mult_ptr, void* stream) {
  if (!c || !mult_ptr || (n > 0 && (!ptrs || !numels))) return fail(B200DP_EINVAL, "null argument");
  if (dtype != B200DP_F32 && dtype != B200DP_BF16 && dtype != B200DP_F16) return fail(B200DP_EINVAL, "multi_scale supports f32/bf16/f16");
  DeviceGuard g(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  static thread_local MTTable t;
  for (int lo = 0; lo < n; lo += kMTMax) {
    const int cnt = (n - lo) < kMTMax ? (n - lo) : kMTMax;
    const int chunks = build_mt(ptrs + lo, numels + lo, cnt, &t);
    if (chunks == 0) continue;
    int grid = chunks < c->sm_count * 8 ? chunks : c->sm_count * 8;
    switch (dtype) {
      case B200DP_F32: multi_scale_kernel<float><<<grid, kStreamThreads, 0, st>>>(t, mult_ptr); break;
      case B200DP_BF16: multi_scale_kernel<__nv_bfloat16><<<grid, kStreamThreads, 0, st>>>(t, mult_ptr); break;
      default: multi_scale_kernel<__half><<<grid, kStreamThreads, 0, st>>>(t, mult_ptr); break;
    }
    KERNEL_OK(c);
  }
  return B200DP_OK;
}

// ------------------------------------------------------------------------------------------------
// K9
// ------------------------------------------------------------------------------------------------
extern "C" int b200dp_multi_adam(b200dp_ctx* c, void* const* params, const void* const* grads, const size_t* numels,
                                 const size_t* offs, int n, void* exp_avg, void* exp_avg_sq, double decay_mul, double l2,
                                 double beta1, double beta2, double neg_step_size, double bc2_sqrt, double eps,
                                 const float* mult_ptr, const float* found_inf_ptr, void* stream) {
  if (!c || (n > 0 && (!params || !grads || !numels || !offs)) || !exp_avg || !exp_avg_sq)
    return fail(B200DP_EINVAL, "null argument");
  DeviceGuard g(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  AdamHyper h;
  h.decay_mul = float(decay_mul);
  h.l2 = float(l2);
  h.lerp_w = float(1.0 - beta1);  // differences in double, one rounding: what torch's Python side hands its kernels
  h.beta2 = float(beta2);
  h.one_minus_beta2 = float(1.0 - beta2);
  h.neg_step_size = float(neg_step_size);
  h.bc2_sqrt = float(bc2_sqrt);
  h.eps = float(eps);
  static thread_local AdamTable t;
  for (int lo = 0; lo < n; lo += kMTMax) {
    const int cnt = (n - lo) < kMTMax ? (n - lo) : kMTMax;
    uint32_t acc = 0;
    for (int i = 0; i < cnt; ++i) {
      t.param[i] = reinterpret_cast<float*>(params[lo + i]);
      t.grad[i] = reinterpret_cast<const float*>(grads[lo + i]);
      t.numel[i] = numels[lo + i];
      t.off[i] = offs[lo + i];
      t.chunk_start[i] = acc;
      acc += uint32_t((numels[lo + i] + kMTChunk - 1) / kMTChunk);
    }
    t.chunk_start[cnt] = acc;
    t.n = cnt;
    if (acc == 0) continue;
    const int grid = int(acc) < c->sm_count * 8 ? int(acc) : c->sm_count * 8;
    multi_adam_kernel<<<grid, kStreamThreads, 0, st>>>(t, reinterpret_cast<float*>(exp_avg),
                                                        reinterpret_cast<float*>(exp_avg_sq), h, mult_ptr, found_inf_ptr);
    KERNEL_OK(c);
  }
  return B200DP_OK;
}

// ------------------------------------------------------------------------------------------------
// K4
// ------------------------------------------------------------------------------------------------
template <typename Table>
static void pack_push(Table* t, int* n, uint32_t* acc, const char* s, char* d, size_t bytes) {
  t->src[*n] = s;
  t->dst[*n] = d;
  t->bytes[*n] = bytes;
  t->chunk_start[*n] = *acc;
  *acc += uint32_t((bytes + t->chunk_bytes - 1) / t->chunk_bytes);
  ++*n;
}

struct BulkCfg {
  int warps, stages, stage_bytes, ctas_per_sm;
};
static BulkCfg bulk_cfg() {
  static const BulkCfg cfg = [] {
    auto geti = [](const char* name, int dflt, int lo, int hi) {
      const char* e = getenv(name);
      int v = e ? atoi(e) : dflt;
      return v < lo ? lo : (v > hi ? hi : v);
    };
    BulkCfg c;
    // defaults from the single-GPU sweep (profiles/r02_pack_sweep_n1.jsonl): 4 issuing warps x 2 stages x 8 KB, 2 CTAs
    // per SM = 6.43 TB/s on 1 GB (98 % of the measured copy peak); 64 KB of shared memory per CTA
    c.warps = geti("B200DP_BULK_WARPS", 4, 1, kBulkMaxWarps);
    c.stages = geti("B200DP_BULK_STAGES", 2, 2, kBulkMaxStages);
    c.stage_bytes = geti("B200DP_BULK_STAGE_KB", 8, 1, 64) * 1024;
    c.ctas_per_sm = geti("B200DP_BULK_CTAS_PER_SM", 2, 1, 4);
    while (size_t(c.warps) * c.stages * c.stage_bytes > size_t(220) * 1024 / c.ctas_per_sm && c.stages > 2) --c.stages;
    while (size_t(c.warps) * c.stages * c.stage_bytes > size_t(220) * 1024 / c.ctas_per_sm && c.warps > 1) --c.warps;
    return c;
  }();
  return cfg;
}

template <int CAP>
static int pack_impl(b200dp_ctx* c, const void* const* srcs, void* const* dsts, const size_t* nbytes, int n, int use_bulk,
                     cudaStream_t st) {
  using Table = PackTableT<CAP>;
  static thread_local Table bulk, plain;
  const BulkCfg cfg = bulk_cfg();
  bulk.chunk_bytes = uint32_t(cfg.stage_bytes);
  plain.chunk_bytes = kPackChunk;
  int nb = 0, np = 0;
  uint32_t accb = 0, accp = 0;
  auto flush = [&](bool force) -> int {
    if (nb > 0 && (force || nb >= CAP - 1)) {
      bulk.chunk_start[nb] = accb;
      bulk.n = nb;
      const size_t smem = size_t(cfg.warps) * cfg.stages * cfg.stage_bytes;
      static bool attr_set = false;  // per instantiation
      if (!attr_set) {
        RT(cudaFuncSetAttribute(pack_bulk_kernel<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        attr_set = true;
      }
      int want = int((accb + cfg.warps - 1) / cfg.warps);
      int cap = c->sm_count * cfg.ctas_per_sm;
      int grid = want < cap ? want : cap;
      pack_bulk_kernel<CAP><<<grid, cfg.warps * 32, smem, st>>>(bulk, c->d_status, cfg.stages);
      KERNEL_OK(c);
      nb = 0;
      accb = 0;
    }
    if (np > 0 && (force || np >= CAP - 2)) {
      plain.chunk_start[np] = accp;
      plain.n = np;
      int grid = int(accp) < c->sm_count * 8 ? int(accp) : c->sm_count * 8;
      pack_kernel<CAP><<<grid, kStreamThreads, 0, st>>>(plain);
      KERNEL_OK(c);
      np = 0;
      accp = 0;
    }
    return B200DP_OK;
  };
  for (int i = 0; i < n; ++i) {
    const char* s = reinterpret_cast<const char*>(srcs[i]);
    char* d = reinterpret_cast<char*>(dsts[i]);
    size_t b = nbytes[i];
    if (b == 0) continue;
    const bool aligned = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
    if (use_bulk && aligned && b >= 16) {
      const size_t main = b & ~size_t(15);
      pack_push(&bulk, &nb, &accb, s, d, main);
      if (b > main) pack_push(&plain, &np, &accp, s + main, d + main, b - main);
    } else {
      pack_push(&plain, &np, &accp, s, d, b);
    }
    int rc = flush(false);
    if (rc) return rc;
  }
  return flush(true);
}

extern "C" int b200dp_pack(b200dp_ctx* c, const void* const* srcs, void* const* dsts, const size_t* nbytes, int n,
                           int use_bulk, void* stream) {
  if (!c || (n > 0 && (!srcs || !dsts || !nbytes))) return fail(B200DP_EINVAL, "null argument");
  DeviceGuard g(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  // a gradient bucket (~10 tensors) fits the small table: ~0.9 KB of kernel parameters instead of ~18 KB
  if (n <= kPackSmall - 2) return pack_impl<kPackSmall>(c, srcs, dsts, nbytes, n, use_bulk, st);
  return pack_impl<kMTMax>(c, srcs, dsts, nbytes, n, use_bulk, st);
}

// ------------------------------------------------------------------------------------------------
// K2/K3
// ------------------------------------------------------------------------------------------------
extern "C" int b200dp_pull(b200dp_ctx* c, int seg, size_t byte_off, size_t nbytes, void* dst, size_t dst_stride,
                           int first, int nranks, int channel, int max_blocks, void* stream) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if ((rc = all_mapped(c, *s))) return rc;
  if (byte_off & 15) return fail(B200DP_EINVAL, "staging offset must be 16-byte aligned");
  if (byte_off + nbytes > s->bytes) return fail(B200DP_EINVAL, "staging range exceeds segment");
  if (first < 0 || nranks < 1 || first + nranks > c->world) return fail(B200DP_EINVAL, "bad rank range [%d,+%d)", first, nranks);
  if (!dst && nbytes) return fail(B200DP_EINVAL, "dst is NULL");
  DeviceGuard g(c->device);
  PullArgs a;
  if ((rc = make_barrier(c, channel, &a.bar))) return rc;
  fill_peers(c, *s, &a.bufs);
  a.byte_off = byte_off;
  a.nbytes = nbytes;
  a.dst = reinterpret_cast<char*>(dst);
  a.dst_stride = dst_stride;
  a.first = first;
  a.nranks = nranks;
  const int grid = clamp_blocks(c, (nbytes / 16 + 1) * size_t(nranks), size_t(kThreads) * 4, max_blocks);
  pull_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(a);
  KERNEL_OK(c);
  return B200DP_OK;
}

// ------------------------------------------------------------------------------------------------
// K2c / K1c: fused large-message gather / reduce (stage + exchange in one launch)
// ------------------------------------------------------------------------------------------------
static int env_blocks(const char* name, int dflt) {
  const char* e = getenv(name);
  int v = e ? atoi(e) : 0;
  return v > 0 ? v : dflt;
}

extern "C" int b200dp_gather(b200dp_ctx* c, int seg, size_t byte_off, const void* src, size_t nbytes, void* dst,
                             size_t dst_stride, int channel, int max_blocks, void* stream) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if ((rc = all_mapped(c, *s))) return rc;
  if (byte_off & 15) return fail(B200DP_EINVAL, "staging offset must be 16-byte aligned");
  if (byte_off + round_up(nbytes, 16) > s->bytes) return fail(B200DP_EINVAL, "staging range exceeds segment");
  if (nbytes == 0) return B200DP_OK;
  if (!src || !dst) return fail(B200DP_EINVAL, "null src/dst");
  DeviceGuard g(c->device);
  GatherArgs a;
  if ((rc = make_barrier(c, channel, &a.bar))) return rc;
  fill_peers(c, *s, &a.bufs);
  a.byte_off = byte_off;
  a.nbytes = nbytes;
  a.src = reinterpret_cast<const char*>(src);
  a.dst = reinterpret_cast<char*>(dst);
  a.dst_stride = dst_stride;
  static const int dflt = env_blocks("B200DP_GATHER_BLOCKS", 148);  // staging + pull need the whole chip (r02 sweep)
  const int grid = clamp_blocks(c, (nbytes + 15) / 16, size_t(kThreads) * 4, max_blocks > 0 ? max_blocks : dflt);
  gather_fused_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(a);
  KERNEL_OK(c);
  return B200DP_OK;
}

template <typename T>
static void launch_pipe_reduce(const PipeReduceArgs& a, int world, bool nvls, int grid, cudaStream_t st) {
  if (nvls) {
    reduce_pipelined_kernel<T, 0, true><<<grid, kThreads, 0, st>>>(a);
    return;
  }
  switch (world) {
    case 2: reduce_pipelined_kernel<T, 2, false><<<grid, kThreads, 0, st>>>(a); break;
    case 4: reduce_pipelined_kernel<T, 4, false><<<grid, kThreads, 0, st>>>(a); break;
    case 8: reduce_pipelined_kernel<T, 8, false><<<grid, kThreads, 0, st>>>(a); break;
    default: reduce_pipelined_kernel<T, 0, false><<<grid, kThreads, 0, st>>>(a); break;
  }
}

template <typename T>
static void launch_reduce(const ReduceArgs& a, int world, bool nvls, int grid, cudaStream_t st) {
  if (nvls) {
    reduce_fused_kernel<T, 0, true><<<grid, kThreads, 0, st>>>(a);
    return;
  }
  switch (world) {
    case 2: reduce_fused_kernel<T, 2, false><<<grid, kThreads, 0, st>>>(a); break;
    case 4: reduce_fused_kernel<T, 4, false><<<grid, kThreads, 0, st>>>(a); break;
    case 8: reduce_fused_kernel<T, 8, false><<<grid, kThreads, 0, st>>>(a); break;
    default: reduce_fused_kernel<T, 0, false><<<grid, kThreads, 0, st>>>(a); break;
  }
}

extern "C" int b200dp_reduce(b200dp_ctx* c, int seg, size_t byte_off, const void* src, size_t count, int dtype,
                             double divisor, void* dst, int algo, int channel, int max_blocks, void* stream) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if ((rc = all_mapped(c, *s))) return rc;
  if (dtype != B200DP_F32 && dtype != B200DP_BF16 && dtype != B200DP_F16)
    return fail(B200DP_EINVAL, "b200dp_reduce supports f32/bf16/f16, got dtype %d", dtype);
  const size_t es = dtype_size(dtype);
  if (byte_off & 15) return fail(B200DP_EINVAL, "staging offset must be 16-byte aligned");
  if (byte_off + round_up(count * es, 16) > s->bytes) return fail(B200DP_EINVAL, "staging range exceeds segment");
  if (count == 0) return B200DP_OK;
  if (!src || !dst) return fail(B200DP_EINVAL, "null src/dst");
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % es)
    return fail(B200DP_EINVAL, "src/dst not aligned to the element size");
  bool nvls = false;
  if (algo == B200DP_ALGO_NVLS) {
    if (!s->mc_bound) return fail(B200DP_ESTATE, "NVLS requested but segment %d has no multicast binding", seg);
    nvls = true;
  } else if (algo == B200DP_ALGO_AUTO) {
    nvls = s->mc_bound && c->world >= 4;
  } else if (algo != B200DP_ALGO_TWOSHOT) {
    return fail(B200DP_EINVAL, "bad algo %d for b200dp_reduce", algo);
  }
  DeviceGuard g(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  // ---- two ranks: one-shot by push (no staging copy, no copy-out) ----
  static const int push2_on = env_blocks("B200DP_PUSH2", 1);
  static const size_t push2_min = size_t(env_blocks("B200DP_PUSH2_MIN_MB", 4)) << 20;
  if (push2_on == 1 && c->world == 2 && !nvls && count * es >= push2_min && (count * es) % 16 == 0 &&
      ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    Push2Args p;
    if ((rc = make_barrier(c, channel, &p.bar))) return rc;
    fill_peers(c, *s, &p.bufs);
    p.byte_off = byte_off;
    p.nvec = count * es / 16;
    p.src = reinterpret_cast<const char*>(src);
    p.dst = reinterpret_cast<char*>(dst);
    p.divisor = float(divisor);
    static const size_t round_kb = size_t(env_blocks("B200DP_PUSH2_ROUND_KB", 128));
    p.round_vecs = round_kb * 1024 / 16;
    const int grid = max_blocks > 0 && max_blocks < c->sm_count ? max_blocks : c->sm_count;
    switch (dtype) {
      case B200DP_F32: reduce_push2_kernel<float><<<grid, kThreads, 0, st>>>(p); break;
      case B200DP_BF16: reduce_push2_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(p); break;
      default: reduce_push2_kernel<__half><<<grid, kThreads, 0, st>>>(p); break;
    }
    KERNEL_OK(c);
    return B200DP_OK;
  }
  // ---- large messages: the pipelined kernel (comm CTAs + copy CTAs, chunk hand-off) ----
  static const size_t pipe_min = size_t(env_blocks("B200DP_PIPE_MIN_MB", 8)) << 20;
  // chunk = 1/8 of the message clamped to [4, 16] MiB unless B200DP_PIPE_CHUNK_MB pins it: per-chunk cost (one flag
  // barrier + the multimem ramp, ~4 us) against pipeline depth (profiles/r02_collective_sweep_n4_*.json)
  static const size_t pipe_chunk_env = size_t(env_blocks("B200DP_PIPE_CHUNK_MB", 0)) << 20;
  size_t pipe_chunk = pipe_chunk_env;
  if (pipe_chunk == 0) {
    pipe_chunk = (count * es / 8) >> 20 << 20;
    if (pipe_chunk < (size_t(4) << 20)) pipe_chunk = size_t(4) << 20;
    if (pipe_chunk > (size_t(16) << 20)) pipe_chunk = size_t(16) << 20;
  }
  static const int pipe_on = env_blocks("B200DP_PIPE_REDUCE", 1);
  // (two-shot worlds keep the fused kernel: with plain peer loads the per-chunk hand-off costs more than it hides)
  // (B200DP_PIPE_REDUCE=2 forces it for any algorithm: tests exercise the two-shot instantiation that way)
  if (((pipe_on == 1 && nvls) || pipe_on == 2) && count * es >= pipe_min && c->sm_count >= 64) {
    PipeReduceArgs p;
    if ((rc = make_barrier(c, channel, &p.bar))) return rc;
    fill_peers(c, *s, &p.bufs);
    p.mc = reinterpret_cast<char*>(s->mc_ptr);
    p.byte_off = byte_off;
    p.count = count;
    p.src = src;
    p.dst = dst;
    p.divisor = float(divisor);
    const size_t nvec = (count * es + 15) / 16;
    size_t chunk_vecs = pipe_chunk / 16;
    while ((nvec + chunk_vecs - 1) / chunk_vecs > size_t(kPipeMaxChunks)) chunk_vecs *= 2;
    p.chunk_vecs = chunk_vecs;
    p.nchunks = int((nvec + chunk_vecs - 1) / chunk_vecs);
    static const int ncomm_nvls = env_blocks("B200DP_PIPE_COMM_NVLS", 16), ncomm_two = env_blocks("B200DP_PIPE_COMM_TWOSHOT", 32);
    p.ncomm = nvls ? ncomm_nvls : ncomm_two;
    const int grid = max_blocks > 0 && max_blocks < c->sm_count ? max_blocks : c->sm_count;
    if (p.ncomm > grid / 2) p.ncomm = grid / 2;
    if (p.ncomm < 1) p.ncomm = 1;
    p.lag = 2;
    p.staged = c->d_pipe;
    p.reduced = c->d_pipe + kPipeMaxChunks;
    RT(cudaMemsetAsync(c->d_pipe, 0, sizeof(uint32_t) * 2 * kPipeMaxChunks, st));
    switch (dtype) {
      case B200DP_F32: launch_pipe_reduce<float>(p, c->world, nvls, grid, st); break;
      case B200DP_BF16: launch_pipe_reduce<__nv_bfloat16>(p, c->world, nvls, grid, st); break;
      default: launch_pipe_reduce<__half>(p, c->world, nvls, grid, st); break;
    }
    KERNEL_OK(c);
    return B200DP_OK;
  }
  ReduceArgs a;
  if ((rc = make_barrier(c, channel, &a.bar))) return rc;
  fill_peers(c, *s, &a.bufs);
  a.mc = reinterpret_cast<char*>(s->mc_ptr);
  a.byte_off = byte_off;
  a.count = count;
  a.src = src;
  a.dst = dst;
  a.divisor = float(divisor);
  static const int dflt = env_blocks("B200DP_REDUCE_FUSED_BLOCKS", 148);  // the local copy phases want every SM (r02 sweep)
  const int grid = clamp_blocks(c, (count * es + 15) / 16, size_t(kThreads) * 4, max_blocks > 0 ? max_blocks : dflt);
  switch (dtype) {
    case B200DP_F32: launch_reduce<float>(a, c->world, nvls, grid, st); break;
    case B200DP_BF16: launch_reduce<__nv_bfloat16>(a, c->world, nvls, grid, st); break;
    default: launch_reduce<__half>(a, c->world, nvls, grid, st); break;
  }
  KERNEL_OK(c);
  return B200DP_OK;
}

// copy-engine peer copy: the "what can this link do" yardstick the kernels are measured against
extern "C" int b200dp_peer_copy(b200dp_ctx* c, int seg, int peer, size_t byte_off, void* dst, size_t nbytes, void* stream) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if (peer < 0 || peer >= c->world || !s->peer_mapped[peer]) return fail(B200DP_ESTATE, "peer %d of segment %d not mapped", peer, seg);
  if (byte_off + nbytes > s->bytes) return fail(B200DP_EINVAL, "range exceeds segment");
  if (!dst) return fail(B200DP_EINVAL, "dst is NULL");
  DeviceGuard g(c->device);
  RT(cudaMemcpyAsync(dst, reinterpret_cast<const char*>(s->peer[peer]) + byte_off, nbytes, cudaMemcpyDeviceToDevice,
                     (cudaStream_t)stream));
  return B200DP_OK;
}

// ------------------------------------------------------------------------------------------------
// K2b
// ------------------------------------------------------------------------------------------------
extern "C" size_t b200dp_push_max_bytes(void) { return size_t(kPushCtas) * kPushWindow; }
extern "C" size_t b200dp_push_area_bytes(void) { return kPushAreaBytes; }

extern "C" int b200dp_push(b200dp_ctx* c, int seg, int mode, const void* src, size_t nbytes, int dtype, double divisor,
                           void* dst, int channel, void* stream) {
  Segment* s;
  int rc = get_seg(c, seg, &s);
  if (rc) return rc;
  if ((rc = all_mapped(c, *s))) return rc;
  if (s->bytes < kPushAreaBytes) return fail(B200DP_EINVAL, "segment too small for the push area (%zu < %zu)", s->bytes, kPushAreaBytes);
  if (nbytes == 0) return B200DP_OK;
  if (nbytes > size_t(kPushCtas) * kPushWindow) return fail(B200DP_EINVAL, "push: %zu bytes exceed the %zu-byte limit", nbytes, size_t(kPushCtas) * kPushWindow);
  if (!src || !dst) return fail(B200DP_EINVAL, "null src/dst");
  const size_t es = mode == 1 ? dtype_size(dtype) : 1;
  if (mode == 1 && (!es || nbytes % es)) return fail(B200DP_EINVAL, "push reduce: bad dtype/size");
  if (mode == 1 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % es))
    return fail(B200DP_EINVAL, "push reduce: src/dst not aligned to the element size");
  DeviceGuard g(c->device);
  PushArgs a;
  if ((rc = make_barrier(c, channel, &a.bar))) return rc;
  fill_peers(c, *s, &a.bufs);
  a.area_off = s->bytes - kPushAreaBytes;
  a.src = reinterpret_cast<const char*>(src);
  a.dst = reinterpret_cast<char*>(dst);
  a.nbytes = nbytes;
  a.divisor = divisor;
  const int grid = int((nbytes + kPushWindow - 1) / kPushWindow);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) {
    push_kernel<0, unsigned char><<<grid, kThreads, 0, st>>>(a);
  } else {
    switch (dtype) {
      case B200DP_F32: push_kernel<1, float><<<grid, kThreads, 0, st>>>(a); break;
      case B200DP_BF16: push_kernel<1, __nv_bfloat16><<<grid, kThreads, 0, st>>>(a); break;
      case B200DP_F16: push_kernel<1, __half><<<grid, kThreads, 0, st>>>(a); break;
      case B200DP_F64: push_kernel<1, double><<<grid, kThreads, 0, st>>>(a); break;
      case B200DP_I32: push_kernel<1, int><<<grid, kThreads, 0, st>>>(a); break;
      case B200DP_I64: push_kernel<1, long long><<<grid, kThreads, 0, st>>>(a); break;
      default: push_kernel<1, unsigned char><<<grid, kThreads, 0, st>>>(a); break;
    }
  }
  KERNEL_OK(c);
  return B200DP_OK;
}

// ------------------------------------------------------------------------------------------------
// K8
// ------------------------------------------------------------------------------------------------
extern "C" int b200dp_exchange_small(b200dp_ctx* c, const void* host_src, int nbytes, void* host_out, void* stream) {
  if (!c || !host_src || !host_out) return fail(B200DP_EINVAL, "null argument");
  if (nbytes <= 0 || nbytes > B200DP_SMALL_MAX) return fail(B200DP_EINVAL, "nbytes %d not in (0, %d]", nbytes, B200DP_SMALL_MAX);
  DeviceGuard g(c->device);
  BarrierCtx b;
  int rc = make_barrier(c, B200DP_CH_SMALL, &b);
  if (rc) return rc;
  SmallPayload p;
  memset(&p, 0, sizeof(p));
  memcpy(&p, host_src, nbytes);
  const int nwords = (nbytes + 3) / 4;
  const int flag = ++c->small_seq;
  exchange_small_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(b, p, nwords, c->d_small, c->d_small_flag, flag);
  KERNEL_OK(c);
  // host wait on the mapped flag: no cudaMemcpy, no stream synchronise
  timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  unsigned long spins = 0;
  while (*reinterpret_cast<volatile int*>(c->h_small_flag) != flag) {
    if ((++spins & 0xfff) == 0) {
      timespec t1;
      clock_gettime(CLOCK_MONOTONIC, &t1);
      double el = (t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec);
      if (el > double(c->timeout_ns) + 5e9) return fail(B200DP_ETIMEOUT, "exchange_small: timed out waiting for the device");
      if (cudaStreamQuery((cudaStream_t)stream) != cudaErrorNotReady) {  // finished (or failed) without raising the flag?
        if (*reinterpret_cast<volatile int*>(c->h_small_flag) == flag) break;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(B200DP_ECUDA, "exchange_small: %s", cudaGetErrorString(e));
      }
    }
  }
  if (*reinterpret_cast<volatile int*>(c->h_status) != 0) return fail(B200DP_ETIMEOUT, "exchange_small: device-side flag wait timed out (status %d)", *c->h_status);
  const char* src = reinterpret_cast<const char*>(c->h_small);
  char* out = reinterpret_cast<char*>(host_out);
  for (int r = 0; r < c->world; ++r) memcpy(out + size_t(r) * nbytes, src + size_t(r) * nwords * 4, nbytes);
  return B200DP_OK;
}

// ------------------------------------------------------------------------------------------------
// K7
// ------------------------------------------------------------------------------------------------
extern "C" int b200dp_pad_copy(b200dp_ctx* c, const void* src, void* dst, size_t outer, size_t old_dim, size_t new_dim,
                               size_t inner_bytes, int elem_size, uint64_t pad_bits, int pad_first, void* stream) {
  if (!c || !dst || (!src && old_dim * outer * inner_bytes)) return fail(B200DP_EINVAL, "null argument");
  if (new_dim < old_dim) return fail(B200DP_EINVAL, "new_dim < old_dim");
  if (elem_size != 1 && elem_size != 2 && elem_size != 4 && elem_size != 8) return fail(B200DP_EINVAL, "elem_size must be 1,2,4,8");
  if (inner_bytes % elem_size) return fail(B200DP_EINVAL, "inner_bytes not a multiple of elem_size");
  const size_t total_bytes = outer * new_dim * inner_bytes;
  if (total_bytes == 0) return B200DP_OK;
  DeviceGuard g(c->device);
  // 16-byte pattern = pad element replicated
  unsigned char pat[16];
  for (int i = 0; i < 16; ++i) pat[i] = (unsigned char)((pad_bits >> (8 * (i % elem_size))) & 0xff);
  uint4 pat16;
  memcpy(&pat16, pat, 16);
  // widest unit that divides the row and keeps both pointers aligned
  int unit = 16;
  const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst);
  while (unit > 1 && ((inner_bytes % unit) || (al % unit))) unit >>= 1;
  const size_t inner_u = inner_bytes / unit;
  const size_t shift = pad_first ? new_dim - old_dim : 0;
  const int grid = stream_grid(c, total_bytes, 0);
  cudaStream_t st = (cudaStream_t)stream;
  switch (unit) {
    case 16: pad_copy_kernel<uint4><<<grid, kStreamThreads, 0, st>>>((const uint4*)src, (uint4*)dst, outer, old_dim, new_dim, inner_u, pat16, shift); break;
    case 8: pad_copy_kernel<uint2><<<grid, kStreamThreads, 0, st>>>((const uint2*)src, (uint2*)dst, outer, old_dim, new_dim, inner_u, pat16, shift); break;
    case 4: pad_copy_kernel<uint32_t><<<grid, kStreamThreads, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, outer, old_dim, new_dim, inner_u, pat16, shift); break;
    case 2: pad_copy_kernel<uint16_t><<<grid, kStreamThreads, 0, st>>>((const uint16_t*)src, (uint16_t*)dst, outer, old_dim, new_dim, inner_u, pat16, shift); break;
    default: pad_copy_kernel<unsigned char><<<grid, kStreamThreads, 0, st>>>((const unsigned char*)src, (unsigned char*)dst, outer, old_dim, new_dim, inner_u, pat16, shift); break;
  }
  KERNEL_OK(c);
  return B200DP_OK;
}
