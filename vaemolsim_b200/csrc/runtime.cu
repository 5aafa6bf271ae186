// runtime.cu -- device memory / stream / event plumbing exported through the C ABI so that the Python host
// (ctypes + NumPy) needs no other CUDA binding.  No reference counterpart: TensorFlow's runtime did this.
#include "common.cuh"
#include <atomic>
#include <stdlib.h>
#include <string.h>
#include <nvtx3/nvToolsExt.h>

namespace vms {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

bool nvtx_enabled() {
  static const bool on = [] {
    const char* e = getenv("VMS_NVTX");
    return e && e[0] && e[0] != '0';
  }();
  return on;
}
void nvtx_push(const char* name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

static int g_sm[64];
static int g_smem[64];
static void fill_props() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  if (g_sm[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    g_sm[dev] = v > 0 ? v : 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    g_smem[dev] = v > 0 ? v : 48 * 1024;
  }
}
int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  fill_props();
  return (dev >= 0 && dev < 64 && g_sm[dev]) ? g_sm[dev] : 148;
}
int max_smem_optin() {
  int dev = 0;
  cudaGetDevice(&dev);
  fill_props();
  return (dev >= 0 && dev < 64 && g_smem[dev]) ? g_smem[dev] : 48 * 1024;
}
}  // namespace vms

using namespace vms;

extern "C" {

const char* vms_last_error(void) { return g_err; }
int vms_abi_version(void) { return 1; }
unsigned long long vms_launch_count(void) { return g_launches.load(); }

vms_status vms_device_count(int* count) {
  VMS_REQUIRE(count, VMS_ERR_INVALID_ARG, "count is NULL");
  *count = 0;
  VMS_CUDA(cudaGetDeviceCount(count));
  return VMS_OK;
}
vms_status vms_set_device(int device) {
  VMS_CUDA(cudaSetDevice(device));
  return VMS_OK;
}
vms_status vms_device_info(int device, int64_t out[5]) {
  VMS_REQUIRE(out, VMS_ERR_INVALID_ARG, "out is NULL");
  cudaDeviceProp p;
  VMS_CUDA(cudaGetDeviceProperties(&p, device));
  out[0] = p.multiProcessorCount;
  out[1] = p.major;
  out[2] = p.minor;
  out[3] = (int64_t)p.sharedMemPerBlockOptin;
  out[4] = p.l2CacheSize;
  return VMS_OK;
}
vms_status vms_malloc(void** p, size_t bytes) {
  VMS_REQUIRE(p, VMS_ERR_INVALID_ARG, "device_ptr is NULL");
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  VMS_CUDA(cudaMalloc(p, bytes));
  return VMS_OK;
}
vms_status vms_free(void* p) {
  if (p) VMS_CUDA(cudaFree(p));
  return VMS_OK;
}
vms_status vms_malloc_host(void** p, size_t bytes) {
  VMS_REQUIRE(p, VMS_ERR_INVALID_ARG, "pinned_ptr is NULL");
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  VMS_CUDA(cudaMallocHost(p, bytes));
  return VMS_OK;
}
vms_status vms_free_host(void* p) {
  if (p) VMS_CUDA(cudaFreeHost(p));
  return VMS_OK;
}
vms_status vms_memcpy_h2d(void* d, const void* s, size_t n, vms_stream st) {
  if (n) VMS_CUDA(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, as_stream(st)));
  return VMS_OK;
}
vms_status vms_memcpy_d2h(void* d, const void* s, size_t n, vms_stream st) {
  if (n) VMS_CUDA(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, as_stream(st)));
  return VMS_OK;
}
vms_status vms_memcpy_d2d(void* d, const void* s, size_t n, vms_stream st) {
  if (n) VMS_CUDA(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, as_stream(st)));
  return VMS_OK;
}
vms_status vms_memcpy2d_d2d(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t rows, vms_stream st) {
  if (w && rows) VMS_CUDA(cudaMemcpy2DAsync(d, dp, s, sp, w, rows, cudaMemcpyDeviceToDevice, as_stream(st)));
  return VMS_OK;
}
vms_status vms_memset(void* p, int v, size_t n, vms_stream st) {
  if (n) VMS_CUDA(cudaMemsetAsync(p, v, n, as_stream(st)));
  return VMS_OK;
}
vms_status vms_stream_create(vms_stream* s) {
  VMS_REQUIRE(s, VMS_ERR_INVALID_ARG, "stream is NULL");
  cudaStream_t st;
  VMS_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *s = (vms_stream)st;
  return VMS_OK;
}
vms_status vms_stream_destroy(vms_stream s) {
  if (s) VMS_CUDA(cudaStreamDestroy(as_stream(s)));
  return VMS_OK;
}
vms_status vms_stream_synchronize(vms_stream s) {
  VMS_CUDA(cudaStreamSynchronize(as_stream(s)));
  return VMS_OK;
}
vms_status vms_device_synchronize(void) {
  VMS_CUDA(cudaDeviceSynchronize());
  return VMS_OK;
}
vms_status vms_event_create(vms_event* e) {
  VMS_REQUIRE(e, VMS_ERR_INVALID_ARG, "event is NULL");
  cudaEvent_t ev;
  VMS_CUDA(cudaEventCreate(&ev));
  *e = (vms_event)ev;
  return VMS_OK;
}
vms_status vms_event_destroy(vms_event e) {
  if (e) VMS_CUDA(cudaEventDestroy((cudaEvent_t)e));
  return VMS_OK;
}
vms_status vms_event_record(vms_event e, vms_stream s) {
  VMS_CUDA(cudaEventRecord((cudaEvent_t)e, as_stream(s)));
  return VMS_OK;
}
vms_status vms_stream_wait_event(vms_stream s, vms_event e) {
  VMS_CUDA(cudaStreamWaitEvent(as_stream(s), (cudaEvent_t)e, 0));
  return VMS_OK;
}
vms_status vms_event_synchronize(vms_event e) {
  VMS_CUDA(cudaEventSynchronize((cudaEvent_t)e));
  return VMS_OK;
}
// ---- CUDA graphs for launch-bound host loops (the tape-based training step: ~100 small kernels issued from Python)
// begin: the stream starts capturing (relaxed mode: allocations are legal while capturing);  end: instantiates the captured
// work and reports how many of this library's kernel launches it holds (they did NOT run: the launch counter is rolled back
// and advanced per replay instead);  abort: ends a capture and discards it.
static thread_local unsigned long long g_capture_count0 = 0;
vms_status vms_graph_begin_capture(vms_stream s) {
  VMS_REQUIRE(s, VMS_ERR_INVALID_ARG, "graph_begin_capture: the legacy default stream cannot capture");
  g_capture_count0 = vms_launch_count();
  VMS_CUDA(cudaStreamBeginCapture(as_stream(s), cudaStreamCaptureModeRelaxed));
  return VMS_OK;
}
vms_status vms_graph_abort_capture(vms_stream s) {
  cudaGraph_t g = nullptr;
  cudaStreamEndCapture(as_stream(s), &g);
  if (g) cudaGraphDestroy(g);
  cudaGetLastError();  // an invalidated capture must not poison later launch checks
  count_launch(-(int)(vms_launch_count() - g_capture_count0));
  return VMS_OK;
}
vms_status vms_graph_end_capture(vms_stream s, void** graph_exec, int* n_kernels) {
  VMS_REQUIRE(graph_exec && n_kernels, VMS_ERR_INVALID_ARG, "graph_end_capture: NULL argument");
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(as_stream(s), &g);
  const int n = (int)(vms_launch_count() - g_capture_count0);
  count_launch(-n);
  if (e != cudaSuccess || !g) {
    cudaGetLastError();
    if (g) cudaGraphDestroy(g);
    set_error("graph capture failed: %s", cudaGetErrorString(e));
    return VMS_ERR_CUDA;
  }
  cudaGraphExec_t ex = nullptr;
  e = cudaGraphInstantiate(&ex, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("graph instantiation failed: %s", cudaGetErrorString(e));
    return VMS_ERR_CUDA;
  }
  *graph_exec = (void*)ex;
  *n_kernels = n;
  return VMS_OK;
}
vms_status vms_graph_launch(void* graph_exec, int n_kernels, vms_stream s) {
  VMS_REQUIRE(graph_exec, VMS_ERR_INVALID_ARG, "graph_launch: NULL graph");
  VMS_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, as_stream(s)));
  count_launch(n_kernels);
  return VMS_OK;
}
vms_status vms_graph_destroy(void* graph_exec) {
  if (graph_exec) VMS_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return VMS_OK;
}

vms_status vms_event_elapsed_ms(vms_event a, vms_event b, float* ms) {
  VMS_REQUIRE(ms, VMS_ERR_INVALID_ARG, "ms is NULL");
  VMS_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return VMS_OK;
}

}  // extern "C"
