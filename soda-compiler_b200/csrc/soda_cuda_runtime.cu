// Host runtime of the SODA CUDA backend: see soda_cuda_runtime.h.
#include "soda_cuda_runtime.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "soda_cuda_device.cuh"

namespace soda {
namespace {

#define SODA_CHECK(call, code)                                              \
  do {                                                                      \
    cudaError_t err_ = (call);                                              \
    if (err_ != cudaSuccess) {                                              \
      fprintf(stderr, "ERROR: %s failed: %s (%s:%d)\n", #call,              \
              cudaGetErrorString(err_), __FILE__, __LINE__);                \
      return (code);                                                        \
    }                                                                       \
  } while (0)

using EncodeTiledFn = CUresult (*)(
    CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
    CUtensorMapFloatOOBfill);

struct Device {
  bool ready = false;
  int error = kNoDeviceInterface;
  int sm_count = 0;
  EncodeTiledFn encode = nullptr;
};

std::mutex g_mutex;
Device g_device;
soda_cuda_stats_t g_stats;
cudaEvent_t g_ev[6];   // kernel start/stop, h2d start/stop, d2h start/stop
bool g_ev_ready = false;
bool g_stats_pending = false;

struct PoolEntry {
  void* ptr;
  size_t bytes;
  bool busy;
};
std::vector<PoolEntry> g_pool;

bool env_flag(const char* name) {
  const char* v = getenv(name);
  return v != nullptr && v[0] != '\0' && v[0] != '0';
}

int ensure_device() {
  if (g_device.ready) return kSuccess;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    fprintf(stderr, "ERROR: no CUDA device; the SODA CUDA backend has no CPU "
                    "fallback\n");
    return kNoDeviceInterface;
  }
  int dev = 0;
  SODA_CHECK(cudaGetDevice(&dev), kNoDeviceInterface);
  SODA_CHECK(cudaDeviceGetAttribute(&g_device.sm_count,
                                    cudaDevAttrMultiProcessorCount, dev),
             kNoDeviceInterface);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SODA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn,
                                     cudaEnableDefault, &qres),
             kNoDeviceInterface);
  g_device.encode = reinterpret_cast<EncodeTiledFn>(fn);
  for (auto& ev : g_ev) SODA_CHECK(cudaEventCreate(&ev), kNoDeviceInterface);
  g_ev_ready = true;
  g_device.ready = true;
  return kSuccess;
}

void* pool_acquire(size_t bytes) {
  PoolEntry* best = nullptr;
  for (auto& e : g_pool)
    if (!e.busy && e.bytes >= bytes && e.bytes <= 2 * bytes + 4096 &&
        (best == nullptr || e.bytes < best->bytes))
      best = &e;
  if (best != nullptr) {
    best->busy = true;
    return best->ptr;
  }
  void* ptr = nullptr;
  if (cudaMalloc(&ptr, bytes) != cudaSuccess) {
    cudaGetLastError();
    // drop idle buffers and retry once
    for (auto& e : g_pool)
      if (!e.busy && e.ptr != nullptr) {
        cudaFree(e.ptr);
        e.ptr = nullptr;
        e.bytes = 0;
      }
    if (cudaMalloc(&ptr, bytes) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
  }
  g_pool.push_back({ptr, bytes, true});
  return ptr;
}

void pool_release(void* ptr) {
  for (auto& e : g_pool)
    if (e.ptr == ptr) e.busy = false;
}

struct PoolLease {   // releases everything it handed out
  std::vector<void*> held;
  void* get(size_t bytes) {
    void* p = pool_acquire(bytes);
    if (p != nullptr) held.push_back(p);
    return p;
  }
  ~PoolLease() {
    for (void* p : held) pool_release(p);
  }
};

const KernelVariant* find_variant(const ProgramDesc& prog, int depth) {
  for (int i = 0; i < prog.n_variants; ++i)
    if (prog.variants[i].depth == depth) return &prog.variants[i];
  return nullptr;
}

CUtensorMapDataType tma_type(int elem) {
  switch (elem) {
    case 1: return CU_TENSOR_MAP_DATA_TYPE_UINT8;
    case 2: return CU_TENSOR_MAP_DATA_TYPE_UINT16;
    case 4: return CU_TENSOR_MAP_DATA_TYPE_UINT32;
    default: return CU_TENSOR_MAP_DATA_TYPE_UINT64;
  }
}

void valid_region(const ProgramDesc& prog, int iterate, const int32_t* dims,
                  int32_t* lo, int32_t* hi) {
  const int* w = prog.window + iterate * 2 * kRtMaxDim;
  for (int d = 0; d < kRtMaxDim; ++d) {
    lo[d] = 0;
    hi[d] = 1;
  }
  for (int d = 0; d < prog.dim; ++d) {
    lo[d] = std::max(0, -w[d]);
    hi[d] = dims[d] - std::max(0, w[kRtMaxDim + d]);
  }
}

// Blocks along the streamed dimension: minimise (waves x steps per block).
int pick_chunks(long long tile_blocks, long long resident, int rows,
                int overhead) {
  const int max_chunks = std::max(1, std::min(rows, 65535));
  long long best_cost = -1;
  int best = 1;
  for (int chunks = 1; chunks <= max_chunks; ++chunks) {
    const int chunk_rows = (rows + chunks - 1) / chunks;
    const int real_chunks = (rows + chunk_rows - 1) / chunk_rows;
    if (real_chunks != chunks) continue;
    const long long waves = (tile_blocks * chunks + resident - 1) / resident;
    const long long cost = waves * (chunk_rows + overhead);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = chunks;
    }
    if (chunk_rows <= overhead) break;   // finer only adds lead-in work
  }
  return best;
}

}  // namespace

int launch(const ProgramDesc& prog, int depth, const void* const* inputs,
           void* const* outputs, const int32_t* dims, int row_begin,
           int row_end, const int32_t* valid_lo, const int32_t* valid_hi,
           cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_mutex);
  int rc = ensure_device();
  if (rc != kSuccess) return rc;
  const KernelVariant* kv = find_variant(prog, depth);
  if (kv == nullptr) {
    fprintf(stderr, "ERROR: %s was not compiled with temporal depth %d\n",
            prog.app_name, depth);
    return kInternalError;
  }
  const int s = prog.dim - 1;
  if (row_end <= row_begin) return kSuccess;

  StreamArgs args;
  memset(&args, 0, sizeof(args));
  long long stride = 1, cells = 1;
  for (int d = 0; d < kRtMaxDim; ++d) {
    args.dims[d] = d < prog.dim ? dims[d] : 1;
    args.stride[d] = stride;
    stride *= args.dims[d];
    args.valid_lo[d] = d < prog.dim ? valid_lo[d] : 0;
    args.valid_hi[d] = d < prog.dim ? valid_hi[d] : 1;
    args.tiles[d] = 1;
  }
  cells = stride;
  if (cells >= (1LL << 40)) return kBufferExtentsTooLarge;
  long long tile_blocks = 1;
  for (int d = 0; d < s; ++d) {
    args.tiles[d] = (dims[d] + kv->own[d] - 1) / kv->own[d];
    tile_blocks *= args.tiles[d];
  }
  if (tile_blocks > 0x7fffffffLL) return kBufferExtentsTooLarge;
  args.row_begin = row_begin;
  args.row_end = row_end;

  bool aligned = (dims[0] % kv->vec) == 0;
  bool tma_ok = !env_flag("SODA_CUDA_NO_TMA");
  for (int k = 0; k < prog.n_in; ++k) {
    args.in_ptr[k] = inputs[k];
    const uintptr_t p = reinterpret_cast<uintptr_t>(inputs[k]);
    if (p % 16 != 0) aligned = tma_ok = false;
    if ((static_cast<long long>(dims[0]) * prog.in_elem[k]) % 16 != 0)
      tma_ok = false;
  }
  for (int k = 0; k < prog.n_out; ++k) {
    args.out_ptr[k] = outputs[k];
    if (reinterpret_cast<uintptr_t>(outputs[k]) % 16 != 0) aligned = false;
  }
  args.vec_store = aligned ? 1 : 0;

  if (tma_ok) {
    for (int k = 0; k < prog.n_in; ++k) {
      cuuint64_t gdim[kRtMaxDim];
      cuuint64_t gstride[kRtMaxDim];
      cuuint32_t box[kRtMaxDim];
      cuuint32_t estr[kRtMaxDim];
      for (int d = 0; d < prog.dim; ++d) {
        gdim[d] = static_cast<cuuint64_t>(dims[d]);
        estr[d] = 1;
        box[d] = d == 0 ? kv->box0 : (d < s ? kv->tile[d] : 1);
        if (d > 0)
          gstride[d - 1] =
              static_cast<cuuint64_t>(args.stride[d]) * prog.in_elem[k];
      }
      CUresult res = g_device.encode(
          &args.in_map[k], tma_type(prog.in_elem[k]), prog.dim,
          const_cast<void*>(inputs[k]), gdim, gstride, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (res != CUDA_SUCCESS) {
        if (env_flag("SODA_CUDA_VERBOSE"))
          fprintf(stderr, "INFO: tensor map rejected (%d); plain loads\n",
                  static_cast<int>(res));
        tma_ok = false;
        break;
      }
    }
  }
  const void* fn = tma_ok ? kv->kernel_tma : kv->kernel_plain;
  SODA_CHECK(cudaFuncSetAttribute(fn,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kv->smem_bytes),
             kDeviceRunFailed);
  int per_sm = 0;
  SODA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                 &per_sm, fn, kv->threads, kv->smem_bytes),
             kDeviceRunFailed);
  if (per_sm < 1) {
    fprintf(stderr, "ERROR: kernel of %s does not fit on an SM (%d B smem)\n",
            prog.app_name, kv->smem_bytes);
    return kDeviceRunFailed;
  }
  const long long resident =
      static_cast<long long>(per_sm) * g_device.sm_count;
  const int rows = row_end - row_begin;
  int chunks = pick_chunks(tile_blocks, resident, rows,
                           kv->lead + kv->out_delay);
  if (const char* forced = getenv("SODA_CUDA_CHUNKS"))
    chunks = std::max(1, std::min(rows, atoi(forced)));
  args.chunk_rows = (rows + chunks - 1) / chunks;
  chunks = (rows + args.chunk_rows - 1) / args.chunk_rows;

  dim3 grid(static_cast<unsigned>(tile_blocks), static_cast<unsigned>(chunks));
  dim3 block(kv->threads);
  void* params[] = {&args};
  SODA_CHECK(cudaLaunchKernel(fn, grid, block, params, kv->smem_bytes, stream),
             kDeviceRunFailed);
  g_stats.launches += 1;
  g_stats.used_tma = tma_ok ? 1 : 0;
  g_stats.blocks = static_cast<int32_t>(tile_blocks * chunks);
  g_stats.threads = kv->threads;
  g_stats.smem_bytes = kv->smem_bytes;
  if (env_flag("SODA_CUDA_VERBOSE"))
    fprintf(stderr,
            "INFO: %s depth %d: grid %lld x %d (%d blocks/SM resident), "
            "%d threads, %d B smem, %s, rows [%d, %d) in chunks of %d\n",
            prog.app_name, depth, tile_blocks, chunks, per_sm, kv->threads,
            kv->smem_bytes, tma_ok ? "TMA" : "plain loads", row_begin, row_end,
            args.chunk_rows);
  return kSuccess;
}

int run_device(const ProgramDesc& prog, const void* const* inputs,
               void* const* outputs, const int32_t* dims, int iterate,
               cudaStream_t stream) {
  int rc;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    rc = ensure_device();
  }
  if (rc != kSuccess) return rc;
  if (iterate <= 0) iterate = prog.iterate;
  if (iterate > prog.iterate) {
    fprintf(stderr, "ERROR: %s was compiled for at most %d iterations\n",
            prog.app_name, prog.iterate);
    return kInternalError;
  }
  // plan the launches: greedily the deepest compiled variant that still fits
  std::vector<int> depths;
  int forced = 0;
  if (const char* v = getenv("SODA_CUDA_DEPTH")) forced = atoi(v);
  for (int left = iterate; left > 0;) {
    int pick = 0;
    for (int i = 0; i < prog.n_variants; ++i) {
      const int d = prog.variants[i].depth;
      if (d <= left && (forced <= 0 || d <= forced) && d > pick) pick = d;
    }
    if (pick == 0) {
      fprintf(stderr, "ERROR: no compiled depth fits %d remaining iterations\n",
              left);
      return kInternalError;
    }
    depths.push_back(pick);
    left -= pick;
  }
  const int n_launch = static_cast<int>(depths.size());
  if (n_launch > 1 && prog.n_in != prog.n_out) return kInternalError;

  long long cells = 1;
  for (int d = 0; d < prog.dim; ++d) cells *= dims[d];
  // ping-pong between the caller's outputs and one scratch set, arranged so
  // the last launch lands in the outputs
  PoolLease lease;
  void* scratch[kRtMaxTensors] = {};
  if (n_launch > 1) {
    std::lock_guard<std::mutex> lock(g_mutex);
    for (int k = 0; k < prog.n_out; ++k) {
      scratch[k] = lease.get(static_cast<size_t>(cells) * prog.out_elem[k]);
      if (scratch[k] == nullptr) return kDeviceMallocFailed;
    }
  }
  int32_t full_lo[kRtMaxDim] = {0, 0, 0, 0};
  int32_t full_hi[kRtMaxDim] = {1, 1, 1, 1};
  for (int d = 0; d < prog.dim; ++d) full_hi[d] = dims[d];
  int32_t fin_lo[kRtMaxDim], fin_hi[kRtMaxDim];
  valid_region(prog, iterate, dims, fin_lo, fin_hi);

  {
    std::lock_guard<std::mutex> lock(g_mutex);
    memset(&g_stats, 0, sizeof(g_stats));
    g_stats.cells = cells;
    g_stats.iterate = iterate;
    g_stats.depth = depths[0];
    g_stats_pending = true;
  }
  SODA_CHECK(cudaEventRecord(g_ev[0], stream), kDeviceRunFailed);
  const void* src[kRtMaxTensors];
  void* dst[kRtMaxTensors];
  for (int k = 0; k < prog.n_in; ++k) src[k] = inputs[k];
  for (int l = 0; l < n_launch; ++l) {
    const bool last = l + 1 == n_launch;
    const bool to_outputs = ((n_launch - 1 - l) % 2) == 0;
    for (int k = 0; k < prog.n_out; ++k)
      dst[k] = to_outputs ? outputs[k] : scratch[k];
    rc = launch(prog, depths[l], src, dst, dims, 0, dims[prog.dim - 1],
                last ? fin_lo : full_lo, last ? fin_hi : full_hi, stream);
    if (rc != kSuccess) return rc;
    for (int k = 0; k < prog.n_out; ++k) src[k] = dst[k];
  }
  SODA_CHECK(cudaEventRecord(g_ev[1], stream), kDeviceRunFailed);
  if (n_launch > 1) {
    // scratch goes back to the pool when the stream has drained past here;
    // the pool is only reused by later calls on the same stream order
    SODA_CHECK(cudaStreamSynchronize(stream), kDeviceSyncFailed);
  }
  return kSuccess;
}

namespace {

bool is_query(const buffer_t* b) { return b->host == nullptr && b->dev == 0; }

void rewrite(buffer_t* b, int elem, int dim, const int32_t* min,
             const int32_t* extent) {
  int32_t stride = 1;
  for (int d = 0; d < 4; ++d) {
    b->min[d] = d < dim ? min[d] : 0;
    b->extent[d] = d < dim ? extent[d] : 0;
    b->stride[d] = d < dim ? stride : 0;
    if (d < dim) stride *= extent[d];
  }
  b->elem_size = elem;
}

}  // namespace

int run_buffers(const ProgramDesc& prog, buffer_t* const* inputs,
                buffer_t* const* outputs, const char* config) {
  (void)config;
  for (int k = 0; k < prog.n_in; ++k)
    if (inputs == nullptr || inputs[k] == nullptr) return kBufferArgumentIsNull;
  for (int k = 0; k < prog.n_out; ++k)
    if (outputs == nullptr || outputs[k] == nullptr)
      return kBufferArgumentIsNull;

  // bounds-query mode (reference host.py:204-252): a buffer with neither host
  // nor device memory gets its shape filled in; nothing is computed.
  bool query = false;
  for (int k = 0; k < prog.n_out; ++k)
    if (is_query(outputs[k])) {
      query = true;
      rewrite(outputs[k], prog.out_elem[k], prog.dim, outputs[k]->min,
              outputs[k]->extent);
    }
  for (int k = 0; k < prog.n_in; ++k)
    if (is_query(inputs[k])) {
      query = true;
      int32_t extent[4] = {0, 0, 0, 0};
      for (int d = 0; d < prog.dim; ++d)
        extent[d] = outputs[0]->extent[d] + prog.stencil_dim[d] - 1;
      rewrite(inputs[k], prog.in_elem[k], prog.dim, outputs[0]->min, extent);
    }
  if (query) return kSuccess;

  for (int k = 0; k < prog.n_out; ++k)
    if (outputs[k]->elem_size != prog.out_elem[k]) {
      fprintf(stderr, "ERROR: Buffer %s has type %s but elem_size of the "
                      "buffer passed in is %d instead of %d\n",
              prog.out_name[k], prog.out_type[k], outputs[k]->elem_size,
              prog.out_elem[k]);
      return kBadElemSize;
    }
  for (int k = 0; k < prog.n_in; ++k)
    if (inputs[k]->elem_size != prog.in_elem[k]) {
      fprintf(stderr, "ERROR: Buffer %s has type %s but elem_size of the "
                      "buffer passed in is %d instead of %d\n",
              prog.in_name[k], prog.in_type[k], inputs[k]->elem_size,
              prog.in_elem[k]);
      return kBadElemSize;
    }

  int32_t dims[kRtMaxDim] = {1, 1, 1, 1};
  long long cells = 1;
  for (int d = 0; d < prog.dim; ++d) {
    dims[d] = inputs[0]->extent[d];
    if (dims[d] <= 0) return kAccessOutOfBounds;
    cells *= dims[d];
  }
  auto dense = [&](const buffer_t* b) {
    int32_t stride = 1;
    for (int d = 0; d < prog.dim; ++d) {
      if (b->extent[d] != dims[d] || b->stride[d] != stride) return false;
      stride *= dims[d];
    }
    return true;
  };
  for (int k = 0; k < prog.n_in; ++k)
    if (!dense(inputs[k])) {
      fprintf(stderr, "ERROR: input %s is not a dense array of the common "
                      "extent\n", prog.in_name[k]);
      return kAccessOutOfBounds;
    }
  for (int k = 0; k < prog.n_out; ++k)
    if (!dense(outputs[k])) {
      fprintf(stderr, "ERROR: output %s is not a dense array of the common "
                      "extent\n", prog.out_name[k]);
      return kAccessOutOfBounds;
    }

  int rc;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    rc = ensure_device();
  }
  if (rc != kSuccess) return rc;

  cudaStream_t stream = nullptr;
  PoolLease lease;
  const void* in_dev[kRtMaxTensors];
  void* out_dev[kRtMaxTensors];
  SODA_CHECK(cudaEventRecord(g_ev[2], stream), kCopyToDeviceFailed);
  for (int k = 0; k < prog.n_in; ++k) {
    const size_t bytes = static_cast<size_t>(cells) * prog.in_elem[k];
    if (inputs[k]->dev != 0) {
      in_dev[k] = reinterpret_cast<const void*>(inputs[k]->dev);
      continue;
    }
    void* p;
    {
      std::lock_guard<std::mutex> lock(g_mutex);
      p = lease.get(bytes);
    }
    if (p == nullptr) return kDeviceMallocFailed;
    SODA_CHECK(cudaMemcpyAsync(p, inputs[k]->host, bytes,
                               cudaMemcpyHostToDevice, stream),
               kCopyToDeviceFailed);
    in_dev[k] = p;
  }
  SODA_CHECK(cudaEventRecord(g_ev[3], stream), kCopyToDeviceFailed);
  for (int k = 0; k < prog.n_out; ++k) {
    const size_t bytes = static_cast<size_t>(cells) * prog.out_elem[k];
    if (outputs[k]->dev != 0) {
      out_dev[k] = reinterpret_cast<void*>(outputs[k]->dev);
      continue;
    }
    std::lock_guard<std::mutex> lock(g_mutex);
    out_dev[k] = lease.get(bytes);
    if (out_dev[k] == nullptr) return kDeviceMallocFailed;
  }
  rc = run_device(prog, in_dev, out_dev, dims, prog.iterate, stream);
  if (rc != kSuccess) return rc;
  SODA_CHECK(cudaEventRecord(g_ev[4], stream), kCopyToHostFailed);
  for (int k = 0; k < prog.n_out; ++k) {
    if (outputs[k]->dev != 0) continue;
    const size_t bytes = static_cast<size_t>(cells) * prog.out_elem[k];
    SODA_CHECK(cudaMemcpyAsync(outputs[k]->host, out_dev[k], bytes,
                               cudaMemcpyDeviceToHost, stream),
               kCopyToHostFailed);
  }
  SODA_CHECK(cudaEventRecord(g_ev[5], stream), kCopyToHostFailed);
  SODA_CHECK(cudaStreamSynchronize(stream), kDeviceSyncFailed);
  float ms = 0;
  if (cudaEventElapsedTime(&ms, g_ev[2], g_ev[3]) == cudaSuccess)
    g_stats.h2d_ms = ms;
  if (cudaEventElapsedTime(&ms, g_ev[4], g_ev[5]) == cudaSuccess)
    g_stats.d2h_ms = ms;
  const soda_cuda_stats_t* st = last_stats();
  if (env_flag("SODA_CUDA_VERBOSE")) {
    // the two lines the reference host prints (host.py:796-800)
    fprintf(stderr, "INFO: Kernel execution time: %lf us\n",
            st->kernel_ms * 1e3);
    fprintf(stderr, "INFO: Kernel throughput: %lf pixel/ns\n",
            st->kernel_ms > 0 ? cells / (st->kernel_ms * 1e6) : 0.0);
  }
  return kSuccess;
}

const soda_cuda_stats_t* last_stats() {
  if (g_stats_pending && g_ev_ready) {
    if (cudaEventSynchronize(g_ev[1]) == cudaSuccess) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, g_ev[0], g_ev[1]) == cudaSuccess)
        g_stats.kernel_ms = ms;
    }
    cudaGetLastError();
    g_stats_pending = false;
  }
  return &g_stats;
}

void release_all() {
  std::lock_guard<std::mutex> lock(g_mutex);
  for (auto& e : g_pool)
    if (e.ptr != nullptr) cudaFree(e.ptr);
  g_pool.clear();
}

}  // namespace soda
