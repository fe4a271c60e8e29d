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

using StreamValueFn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t,
                                   unsigned int);
using AddressRangeFn = CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr);

struct Device {
  bool ready = false;
  int error = kNoDeviceInterface;
  int sm_count = 0;
  EncodeTiledFn encode = nullptr;
  StreamValueFn write_value = nullptr;
  StreamValueFn wait_value = nullptr;
  AddressRangeFn address_range = nullptr;
};

std::mutex g_mutex;
Device g_device;
soda_cuda_stats_t g_stats;
void* g_param_dev[kRtMaxTensors] = {};   // device copies of the param arrays
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
  SODA_CHECK(cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn,
                                     cudaEnableDefault, &qres),
             kNoDeviceInterface);
  g_device.write_value = reinterpret_cast<StreamValueFn>(fn);
  SODA_CHECK(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn,
                                     cudaEnableDefault, &qres),
             kNoDeviceInterface);
  g_device.wait_value = reinterpret_cast<StreamValueFn>(fn);
  SODA_CHECK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn,
                                     cudaEnableDefault, &qres),
             kNoDeviceInterface);
  g_device.address_range = reinterpret_cast<AddressRangeFn>(fn);
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

// One box of kRtMaxDim ints per output.
struct Boxes {
  int32_t lo[kRtMaxTensors * kRtMaxDim];
  int32_t hi[kRtMaxTensors * kRtMaxDim];
};

// Where each output is defined after `iterate` iterations: the bounds of the
// reference's golden loop for that tensor (host.py:1082-1091).
void valid_region(const ProgramDesc& prog, int iterate, const int32_t* dims,
                  Boxes* boxes) {
  for (int k = 0; k < kRtMaxTensors; ++k) {
    int32_t* lo = boxes->lo + k * kRtMaxDim;
    int32_t* hi = boxes->hi + k * kRtMaxDim;
    for (int d = 0; d < kRtMaxDim; ++d) {
      lo[d] = 0;
      hi[d] = 1;
    }
    if (k >= prog.n_out) continue;
    const int* w =
        prog.out_window + (iterate * prog.n_out + k) * 2 * kRtMaxDim;
    for (int d = 0; d < prog.dim; ++d) {
      lo[d] = std::max(0, -w[d]);
      hi[d] = dims[d] - std::max(0, w[kRtMaxDim + d]);
    }
  }
}

// Intermediate launches store every cell they own.
void full_region(const ProgramDesc& prog, const int32_t* dims, Boxes* boxes) {
  for (int k = 0; k < kRtMaxTensors; ++k)
    for (int d = 0; d < kRtMaxDim; ++d) {
      boxes->lo[k * kRtMaxDim + d] = 0;
      boxes->hi[k * kRtMaxDim + d] = d < prog.dim ? dims[d] : 1;
    }
}

// Blocks along the streamed dimension: minimise (waves x steps per block).
int pick_chunks(long long tile_blocks, long long resident, int rows,
                int overhead, int trip) {
  const int max_chunks = std::max(1, std::min(rows, 65535));
  long long best_cost = -1;
  int best = 1;
  for (int chunks = 1; chunks <= max_chunks; ++chunks) {
    const int chunk_rows = (rows + chunks - 1) / chunks;
    const int real_chunks = (rows + chunk_rows - 1) / chunk_rows;
    if (real_chunks != chunks) continue;
    const long long waves = (tile_blocks * chunks + resident - 1) / resident;
    // (a block runs whole trips of its streamed loop, up to trip - 1 surplus
    // steps; counting them here picks coarser chunks, which measured 2-7 %
    // slower on blur, sobel2d and denoise2d: the finer grid balances the tail)
    (void)trip;
    const long long cost = waves * (chunk_rows + overhead);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = chunks;
    }
    if (chunk_rows <= overhead) break;   // finer only adds lead-in work
  }
  return best;
}

// The launches of a run: greedily the deepest compiled variant that still
// fits the remaining iterations (SODA_CUDA_DEPTH caps the depth).
int plan_depths(const ProgramDesc& prog, int iterate, std::vector<int>* depths) {
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
    depths->push_back(pick);
    left -= pick;
  }
  if (depths->size() > 1 && prog.n_in != prog.n_out) return kInternalError;
  return kSuccess;
}

// Rows per block along the streamed dimension for a launch over `rows` rows.
int choose_chunk_rows(const ProgramDesc& prog, const KernelVariant* kv,
                      const void* fn, const int32_t* dims, int rows,
                      long long* grid_x_out, int* per_sm_out) {
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
  long long tile_blocks = 1;
  for (int d = 0; d + 1 < prog.dim; ++d)
    tile_blocks *= (dims[d] + kv->own[d] - 1) / kv->own[d];
  const long long resident =
      static_cast<long long>(per_sm) * g_device.sm_count;
  // 2-D register kernels pack `tiles_per_block` independent strips in a block
  const int per_block = std::max(1, kv->tiles_per_block);
  const long long grid_x = (tile_blocks + per_block - 1) / per_block;
  int chunks = pick_chunks(grid_x, resident, rows, kv->lead + kv->out_delay,
                           std::max(1, kv->trip));
  if (const char* forced = getenv("SODA_CUDA_CHUNKS"))
    chunks = std::max(1, std::min(rows, atoi(forced)));
  if (grid_x_out != nullptr) *grid_x_out = grid_x;
  if (per_sm_out != nullptr) *per_sm_out = per_sm;
  return (rows + chunks - 1) / chunks;
}

}  // namespace

int chunk_rows(const ProgramDesc& prog, int depth, const int32_t* dims,
               int rows) {
  std::lock_guard<std::mutex> lock(g_mutex);
  int rc = ensure_device();
  if (rc != kSuccess) return rc;
  const KernelVariant* kv = find_variant(prog, depth);
  if (kv == nullptr || rows < 1) return kInternalError;
  return choose_chunk_rows(prog, kv, kv->kernel_tma, dims, rows, nullptr,
                           nullptr);
}

int set_params(const ProgramDesc& prog, const void* const* host_arrays) {
  std::lock_guard<std::mutex> lock(g_mutex);
  int rc = ensure_device();
  if (rc != kSuccess) return rc;
  for (int k = 0; k < prog.n_param; ++k) {
    if (host_arrays == nullptr || host_arrays[k] == nullptr)
      return kBufferArgumentIsNull;
    size_t bytes = static_cast<size_t>(prog.param_elem[k]);
    for (int d = 0; d < prog.param_rank[k]; ++d) bytes *= prog.param_size[k][d];
    if (g_param_dev[k] == nullptr)
      SODA_CHECK(cudaMalloc(&g_param_dev[k], std::max<size_t>(bytes, 16)),
                 kDeviceMallocFailed);
    // earlier launches may still be reading the previous values
    SODA_CHECK(cudaDeviceSynchronize(), kDeviceSyncFailed);
    SODA_CHECK(cudaMemcpy(g_param_dev[k], host_arrays[k], bytes,
                          cudaMemcpyHostToDevice),
               kCopyToDeviceFailed);
  }
  return kSuccess;
}

int lead_rows(const ProgramDesc& prog, int depth) {
  const KernelVariant* kv = find_variant(prog, depth);
  if (kv == nullptr) return kInternalError;
  const int trip = std::max(1, kv->trip);
  return (kv->lead + kv->out_delay + trip - 1) / trip * trip;
}

int launch(const ProgramDesc& prog, int depth, const void* const* inputs,
           void* const* outputs, const int32_t* dims, int row_begin,
           int row_end, const int32_t* valid_lo, const int32_t* valid_hi,
           cudaStream_t stream, int forced_chunk_rows) {
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
    args.tiles[d] = 1;
  }
  for (int k = 0; k < prog.n_out; ++k)
    for (int d = 0; d < kRtMaxDim; ++d) {
      args.valid_lo[k][d] = d < prog.dim ? valid_lo[k * kRtMaxDim + d] : 0;
      args.valid_hi[k][d] = d < prog.dim ? valid_hi[k * kRtMaxDim + d] : 1;
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
  for (int k = 0; k < prog.n_param; ++k) {
    if (g_param_dev[k] == nullptr) {
      fprintf(stderr, "ERROR: param %s of %s has not been set "
                      "(soda_cuda_set_params)\n",
              prog.param_name[k], prog.app_name);
      return kBufferArgumentIsNull;
    }
    args.param_ptr[k] = g_param_dev[k];
  }

  if (tma_ok && kv->uses_tma) {
    for (int k = 0; k < prog.n_in; ++k) {
      cuuint64_t gdim[kRtMaxDim];
      cuuint64_t gstride[kRtMaxDim];
      cuuint32_t box[kRtMaxDim];
      cuuint32_t estr[kRtMaxDim];
      for (int d = 0; d < prog.dim; ++d) {
        gdim[d] = static_cast<cuuint64_t>(dims[d]);
        estr[d] = 1;
        box[d] = d == 0 ? kv->box0 : (d < s ? kv->tile[d] : kv->box_rows);
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
  const int rows = row_end - row_begin;
  long long grid_x = 0;
  int per_sm = 0;
  const int chosen = choose_chunk_rows(prog, kv, fn, dims, rows, &grid_x,
                                       &per_sm);
  if (chosen < 0) return chosen;
  args.chunk_rows = forced_chunk_rows > 0 ? std::min(forced_chunk_rows, rows)
                                          : chosen;
  const int chunks = (rows + args.chunk_rows - 1) / args.chunk_rows;
  if (chunks > 65535) return kBufferExtentsTooLarge;

  dim3 grid(static_cast<unsigned>(grid_x), static_cast<unsigned>(chunks));
  dim3 block(kv->threads);
  void* params[] = {&args};
  SODA_CHECK(cudaLaunchKernel(fn, grid, block, params, kv->smem_bytes, stream),
             kDeviceRunFailed);
  g_stats.launches += 1;
  g_stats.used_tma = tma_ok ? 1 : 0;
  g_stats.blocks = static_cast<int32_t>(grid_x * chunks);
  g_stats.threads = kv->threads;
  g_stats.smem_bytes = kv->smem_bytes;
  if (env_flag("SODA_CUDA_VERBOSE"))
    fprintf(stderr,
            "INFO: %s depth %d: grid %lld x %d (%d blocks/SM resident), "
            "%d threads, %d B smem, %s, rows [%d, %d) in chunks of %d\n",
            prog.app_name, depth, grid_x, chunks, per_sm, kv->threads,
            kv->smem_bytes,
            !tma_ok ? "plain loads" : kv->uses_tma ? "TMA" : "128-bit loads",
            row_begin, row_end,
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
  std::vector<int> depths;
  rc = plan_depths(prog, iterate, &depths);
  if (rc != kSuccess) return rc;
  const int n_launch = static_cast<int>(depths.size());

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
  Boxes full, fin;
  full_region(prog, dims, &full);
  valid_region(prog, iterate, dims, &fin);

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
                last ? fin.lo : full.lo, last ? fin.hi : full.hi, stream);
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

namespace {

cudaStream_t g_streams[3];   // h2d, compute, d2h
bool g_streams_ready = false;

// Host buffers, large problem: cut the streamed dimension into pieces and
// overlap  H2D(piece k+1) | all launches on piece k | D2H(piece k-1).
//
// Every launch j keeps a frontier f_j: its output rows [0, f_j) are done.
// When input rows [0, avail) are on the device, launch 0 can extend its
// frontier to avail - reach_hi (to N once everything is loaded), launch 1
// follows launch 0's frontier the same way, and so on; rows behind the last
// launch's frontier are final and go back to the host.  No cell is computed
// twice and every cell sees exactly the operands of the one-shot run, so the
// result is bit-identical.  The ping-pong buffers are shared between launches
// j and j-2; holding f_j back by max(reach_hi[j], reach_lo[j-1]) keeps launch
// j's writes below every row launch j-1 will still read.
int run_pipelined(const ProgramDesc& prog, buffer_t* const* inputs,
                  buffer_t* const* outputs, const int32_t* dims,
                  long long cells, int pieces) {
  const int s = prog.dim - 1;
  const int rows = dims[s];
  const long long row_cells = cells / rows;
  std::vector<int> depths;
  int rc = plan_depths(prog, prog.iterate, &depths);
  if (rc != kSuccess) return rc;
  const int n_launch = static_cast<int>(depths.size());
  if (!g_streams_ready) {
    for (auto& st : g_streams)
      SODA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking),
                 kDeviceRunFailed);
    g_streams_ready = true;
  }
  cudaStream_t s_in = g_streams[0], s_run = g_streams[1], s_out = g_streams[2];

  PoolLease lease;
  void* in_dev[kRtMaxTensors];
  void* out_dev[kRtMaxTensors];
  void* scratch[kRtMaxTensors] = {};
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    for (int k = 0; k < prog.n_in; ++k) {
      in_dev[k] = lease.get(static_cast<size_t>(cells) * prog.in_elem[k]);
      if (in_dev[k] == nullptr) return kDeviceMallocFailed;
    }
    for (int k = 0; k < prog.n_out; ++k) {
      out_dev[k] = lease.get(static_cast<size_t>(cells) * prog.out_elem[k]);
      if (out_dev[k] == nullptr) return kDeviceMallocFailed;
      if (n_launch > 1) {
        scratch[k] = lease.get(static_cast<size_t>(cells) * prog.out_elem[k]);
        if (scratch[k] == nullptr) return kDeviceMallocFailed;
      }
    }
    memset(&g_stats, 0, sizeof(g_stats));
    g_stats.cells = cells;
    g_stats.iterate = prog.iterate;
    g_stats.depth = depths[0];
    g_stats_pending = false;
  }
  // streamed reach of every launch
  std::vector<int> reach_lo(n_launch), reach_hi(n_launch), hold(n_launch);
  for (int j = 0; j < n_launch; ++j) {
    const int* w = prog.window + depths[j] * 2 * kRtMaxDim;
    reach_lo[j] = std::max(0, -w[s]);
    reach_hi[j] = std::max(0, w[kRtMaxDim + s]);
  }
  for (int j = 0; j < n_launch; ++j)
    hold[j] = std::max(reach_hi[j], j > 0 ? reach_lo[j - 1] : 0);
  Boxes full, fin;
  full_region(prog, dims, &full);
  valid_region(prog, prog.iterate, dims, &fin);

  std::vector<int> frontier(n_launch, 0);
  std::vector<cudaEvent_t> events;
  auto new_event = [&]() {
    cudaEvent_t ev;
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    events.push_back(ev);
    return ev;
  };
  SODA_CHECK(cudaEventRecord(g_ev[2], s_in), kCopyToDeviceFailed);
  SODA_CHECK(cudaEventRecord(g_ev[0], s_run), kDeviceRunFailed);
  int loaded = 0;
  for (int piece = 0; piece < pieces; ++piece) {
    const int upto = static_cast<int>(
        static_cast<long long>(rows) * (piece + 1) / pieces);
    if (upto <= loaded) continue;
    for (int k = 0; k < prog.n_in; ++k) {
      const size_t off =
          static_cast<size_t>(loaded) * row_cells * prog.in_elem[k];
      const size_t bytes =
          static_cast<size_t>(upto - loaded) * row_cells * prog.in_elem[k];
      SODA_CHECK(cudaMemcpyAsync(static_cast<char*>(in_dev[k]) + off,
                                 inputs[k]->host + off, bytes,
                                 cudaMemcpyHostToDevice, s_in),
                 kCopyToDeviceFailed);
    }
    loaded = upto;
    cudaEvent_t arrived = new_event();
    SODA_CHECK(cudaEventRecord(arrived, s_in), kCopyToDeviceFailed);
    SODA_CHECK(cudaStreamWaitEvent(s_run, arrived, 0), kDeviceRunFailed);
    if (piece + 1 == pieces)
      SODA_CHECK(cudaEventRecord(g_ev[3], s_in), kCopyToDeviceFailed);

    int avail = loaded;
    const int done_before = frontier[n_launch - 1];
    const void* src[kRtMaxTensors];
    void* dst[kRtMaxTensors];
    for (int j = 0; j < n_launch; ++j) {
      const int target =
          avail >= rows ? rows : std::max(frontier[j], avail - hold[j]);
      const bool last = j + 1 == n_launch;
      const bool to_outputs = ((n_launch - 1 - j) % 2) == 0;
      for (int k = 0; k < prog.n_in; ++k)
        src[k] = j == 0 ? in_dev[k]
                        : (to_outputs ? scratch[k] : out_dev[k]);
      for (int k = 0; k < prog.n_out; ++k)
        dst[k] = to_outputs ? out_dev[k] : scratch[k];
      if (target > frontier[j]) {
        rc = launch(prog, depths[j], src, dst, dims, frontier[j], target,
                    last ? fin.lo : full.lo, last ? fin.hi : full.hi, s_run);
        if (rc != kSuccess) return rc;
        frontier[j] = target;
      }
      avail = frontier[j];
    }
    const int done = frontier[n_launch - 1];
    if (done > done_before) {
      cudaEvent_t computed = new_event();
      SODA_CHECK(cudaEventRecord(computed, s_run), kDeviceRunFailed);
      SODA_CHECK(cudaStreamWaitEvent(s_out, computed, 0), kCopyToHostFailed);
      if (done_before == 0)
        SODA_CHECK(cudaEventRecord(g_ev[4], s_out), kCopyToHostFailed);
      for (int k = 0; k < prog.n_out; ++k) {
        const size_t off =
            static_cast<size_t>(done_before) * row_cells * prog.out_elem[k];
        const size_t bytes = static_cast<size_t>(done - done_before) *
                             row_cells * prog.out_elem[k];
        SODA_CHECK(cudaMemcpyAsync(outputs[k]->host + off,
                                   static_cast<char*>(out_dev[k]) + off, bytes,
                                   cudaMemcpyDeviceToHost, s_out),
                   kCopyToHostFailed);
      }
    }
  }
  SODA_CHECK(cudaEventRecord(g_ev[1], s_run), kDeviceRunFailed);
  SODA_CHECK(cudaEventRecord(g_ev[5], s_out), kCopyToHostFailed);
  SODA_CHECK(cudaStreamSynchronize(s_in), kDeviceSyncFailed);
  SODA_CHECK(cudaStreamSynchronize(s_run), kDeviceSyncFailed);
  SODA_CHECK(cudaStreamSynchronize(s_out), kDeviceSyncFailed);
  for (cudaEvent_t ev : events) cudaEventDestroy(ev);
  float ms = 0;
  if (cudaEventElapsedTime(&ms, g_ev[2], g_ev[3]) == cudaSuccess)
    g_stats.h2d_ms = ms;
  if (cudaEventElapsedTime(&ms, g_ev[4], g_ev[5]) == cudaSuccess)
    g_stats.d2h_ms = ms;
  if (cudaEventElapsedTime(&ms, g_ev[0], g_ev[1]) == cudaSuccess)
    g_stats.kernel_ms = ms;
  g_stats.reserved = pieces;
  return kSuccess;
}

}  // namespace

int run_buffers(const ProgramDesc& prog, buffer_t* const* inputs,
                buffer_t* const* outputs, const char* config,
                buffer_t* const* params) {
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

  if (prog.n_param > 0 && params != nullptr) {
    // param arrays: small, host memory, uploaded before the launches
    const void* host_arrays[kRtMaxTensors] = {};
    for (int k = 0; k < prog.n_param; ++k) {
      const buffer_t* b = params[k];
      if (b == nullptr || b->host == nullptr) return kBufferArgumentIsNull;
      if (b->elem_size != prog.param_elem[k]) {
        fprintf(stderr, "ERROR: Buffer %s has type %s but elem_size of the "
                        "buffer passed in is %d instead of %d\n",
                prog.param_name[k], prog.param_type[k], b->elem_size,
                prog.param_elem[k]);
        return kBadElemSize;
      }
      for (int d = 0; d < prog.param_rank[k]; ++d)
        if (b->extent[d] != prog.param_size[k][d]) return kAccessOutOfBounds;
      host_arrays[k] = b->host;
    }
    const int rc_params = set_params(prog, host_arrays);
    if (rc_params != kSuccess) return rc_params;
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

  // large all-host problems: overlap the copies with the launches
  bool all_host = true;
  size_t moved = 0;
  for (int k = 0; k < prog.n_in; ++k) {
    all_host = all_host && inputs[k]->dev == 0;
    moved += static_cast<size_t>(cells) * prog.in_elem[k];
  }
  for (int k = 0; k < prog.n_out; ++k) {
    all_host = all_host && outputs[k]->dev == 0;
    moved += static_cast<size_t>(cells) * prog.out_elem[k];
  }
  int pieces = static_cast<int>(std::min<size_t>(16, moved >> 27));  // 128 MiB
  if (const char* v = getenv("SODA_CUDA_PIECES")) pieces = atoi(v);
  pieces = std::min(pieces, dims[prog.dim - 1]);
  if (all_host && pieces > 1) {
    rc = run_pipelined(prog, inputs, outputs, dims, cells, pieces);
    if (rc == kSuccess && env_flag("SODA_CUDA_VERBOSE"))
      fprintf(stderr, "INFO: %d pieces: h2d %.3f ms, launches %.3f ms, d2h "
                      "%.3f ms (overlapped)\n", pieces, g_stats.h2d_ms,
              g_stats.kernel_ms, g_stats.d2h_ms);
    return rc;
  }

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

int flag_write(void* flag, uint32_t value, cudaStream_t stream) {
  if (int code = ensure_device()) return code;
  const CUresult res = g_device.write_value(
      stream, reinterpret_cast<CUdeviceptr>(flag), value,
      CU_STREAM_WRITE_VALUE_DEFAULT);
  if (res != CUDA_SUCCESS) {
    fprintf(stderr, "ERROR: cuStreamWriteValue32 failed (%d)\n",
            static_cast<int>(res));
    return kDeviceRunFailed;
  }
  return kSuccess;
}

int flag_wait_geq(void* flag, uint32_t value, cudaStream_t stream) {
  if (int code = ensure_device()) return code;
  const CUresult res = g_device.wait_value(
      stream, reinterpret_cast<CUdeviceptr>(flag), value,
      CU_STREAM_WAIT_VALUE_GEQ);
  if (res != CUDA_SUCCESS) {
    fprintf(stderr, "ERROR: cuStreamWaitValue32 failed (%d)\n",
            static_cast<int>(res));
    return kDeviceRunFailed;
  }
  return kSuccess;
}

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are 64 bytes");

int ipc_export(const void* ptr, unsigned char handle[64], uint64_t* offset) {
  if (int code = ensure_device()) return code;
  CUdeviceptr base = 0;
  size_t size = 0;
  if (g_device.address_range(&base, &size,
                             reinterpret_cast<CUdeviceptr>(ptr)) !=
      CUDA_SUCCESS) {
    fprintf(stderr, "ERROR: %p is not a device allocation\n", ptr);
    return kDeviceRunFailed;
  }
  cudaIpcMemHandle_t h;
  SODA_CHECK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)),
             kDeviceRunFailed);
  memcpy(handle, &h, 64);
  *offset = reinterpret_cast<CUdeviceptr>(ptr) - base;
  return kSuccess;
}

int ipc_open(const unsigned char handle[64], void** base) {
  if (int code = ensure_device()) return code;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  SODA_CHECK(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess),
             kDeviceRunFailed);
  return kSuccess;
}

int ipc_close(void* base) {
  SODA_CHECK(cudaIpcCloseMemHandle(base), kDeviceRunFailed);
  return kSuccess;
}

int copy_async(void* dst, const void* src, uint64_t bytes,
               cudaStream_t stream) {
  SODA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream),
             kDeviceRunFailed);
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
  for (void*& p : g_param_dev) {
    if (p != nullptr) cudaFree(p);
    p = nullptr;
  }
}

}  // namespace soda
