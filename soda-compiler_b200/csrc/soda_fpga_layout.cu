// FPGA wire format <-> dense arrays on the GPU (include/soda_fpga_layout.h).
//
// The two loop nests of the reference's generated host (reference
// src/soda/codegen/xilinx/host.py:629-686 pack, :823-901 unpack) as sm_100a
// kernels.  Pure data movement, HBM-bound: one block per tile row, threads
// walk dimension 0, so the dense side is read (pack) or written (unpack) in
// coalesced runs of a tile row and the bank side in runs interleaved
// `num_bank` ways.  blockIdx.y is the tile, blockIdx.x the row in the tile.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "soda_fpga_layout.h"

namespace {

enum { kNull = -12, kBadDescriptor = -4, kNoDevice = -19, kLaunchFailed = -23 };

struct Banks {
  void* ptr[4];
};

// What a block works on: one tile row (fixed in-tile coordinates in every
// dimension but 0), decoded once per block.
struct Row {
  bool inside;         // false: nothing of this row is moved
  long long original;  // dense offset of in-tile coordinate i = 0
  long long stream;    // stream offset of i = 0 (tile base + row + lag)
  int i_lo, i_hi;      // in-tile coordinates [i_lo, i_hi) are moved
};

template <bool kPack>
__device__ __forceinline__ Row decode_row(const soda_fpga_layout_t& a,
                                          long long row, int tile) {
  const int last = a.dim - 1;
  // the tile: dimension 0 of the tile index is the fastest
  int tile_index[3] = {0, 0, 0};
  int rest = tile;
  const long long tile_linear = rest;
  int extent0 = 0;
  bool inside = true;
  long long original = 0, pitch = 1, row_offset = 0, in_tile_pitch = 1;
  // (`row`: in-tile coordinates of the row in dimensions 1..last)
  for (int d = 0; d <= last; ++d) {
    int extent = a.dims[d];
    if (d < last) {
      tile_index[d] = rest % a.tile_num[d];
      rest /= a.tile_num[d];
      // the last tile of a dimension holds what is left of the grid
      extent = tile_index[d] == a.tile_num[d] - 1
                   ? a.dims[d] - a.tile_step[d] * tile_index[d]
                   : a.tile_size[d];
    }
    if (d == 0) {
      extent0 = extent;
      original += static_cast<long long>(tile_index[0]) * a.tile_step[0];
    } else {
      int c;
      if (d < last) {
        c = static_cast<int>(row % a.tile_size[d]);
        row /= a.tile_size[d];
      } else {
        c = static_cast<int>(row);
      }
      inside = inside && c >= a.lo[d] && c < extent - a.hi_margin[d];
      // Unpack: the reference's tile loops run in ascending order, so where
      // the cell ranges of neighbouring tiles overlap (programs whose first
      // input has a narrower window than the program: denoise's `f`) the
      // LATER tile's value stays.  Only that tile writes here.
      if (!kPack && d < last && tile_index[d] + 1 < a.tile_num[d] &&
          c - a.tile_step[d] >= a.lo[d])
        inside = false;
      const int coord = c + (d < last ? tile_index[d] * a.tile_step[d] : 0);
      original += coord * pitch;
      row_offset += c * in_tile_pitch;
    }
    pitch *= a.dims[d];
    if (d < last) in_tile_pitch *= a.tile_size[d];
  }
  Row r;
  r.inside = inside;
  r.original = original;
  r.i_lo = a.lo[0];
  r.i_hi = extent0 - a.hi_margin[0];
  if (!kPack && tile_index[0] + 1 < a.tile_num[0])
    r.i_hi = min(r.i_hi, a.lo[0] + a.tile_step[0]);  // the next tile owns the rest
  r.stream = tile_linear * a.tile_size_linearized + row_offset + a.stream_offset;
  return r;
}

// One block per tile row: threads walk dimension 0, four cells each per trip;
// the bank count is a compile-time constant — no per-cell division by a
// run-time value.  Element-sized accesses on both sides.
template <typename T, bool kPack, int kBanks>
__global__ void __launch_bounds__(256)
wire_kernel(const __grid_constant__ soda_fpga_layout_t a, T* dense,
            const __grid_constant__ Banks banks) {
  const Row r = decode_row<kPack>(a, blockIdx.x, blockIdx.y);
  if (!r.inside) return;
  T* const row_dense = dense + r.original;
  T* bank[kBanks];
#pragma unroll
  for (int b = 0; b < kBanks; ++b)
    bank[b] = static_cast<T*>(banks.ptr[a.bank_vec[b]]);
  for (int i0 = r.i_lo + threadIdx.x * 4; i0 < r.i_hi; i0 += 1024) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k;
      if (i >= r.i_hi) break;
      const unsigned long long o =
          static_cast<unsigned long long>(r.stream + i);
      T* const slot = bank[o % kBanks] + o / kBanks;
      if (kPack)
        *slot = row_dense[i];
      else
        row_dense[i] = *slot;
    }
  }
}

// ---- staged variant: one WARP per tile row ------------------------------------
//
// A tile row is one contiguous run of the dense array and one contiguous run
// in every bank (the stream interleaves the banks element by element), all
// at unrelated alignments.  A warp owns a shared-memory copy of the row and
// walks several rows; no block barrier.
//   dense side   16-byte global accesses aligned to the dense address, and —
//                the copy is laid out at the same phase — 16-byte shared
//                accesses (conflict-free);
//   bank side    4/8-byte global accesses, consecutive lanes on consecutive
//                words (coalesced), element-wise shared accesses with a stride
//                of `kBanks` elements (conflict-free for 16-bit elements in
//                two banks);
// with the loads of a row in flight together before the first is used.
// History (profiles/README.md): element-wise kernel 2258 / 1878 GB/s; a row per
// 256-thread block with a barrier between the sides 1981 / 1853; a row per
// warp with 16-byte accesses on both global sides but element-wise shared
// accesses 2583 / 2305 (shared-memory bank conflicts: ~770 wavefronts per
// 4 KB row).
template <typename T>
struct alignas(16) Vec16 {
  T v[16 / sizeof(T)];
};

constexpr int kWireWarps = 8;       // warps per block
constexpr int kWireUnroll = 4;      // accesses a lane has in flight

// Dense run g[0..n) <-> shared copy buf[phase + j], phase = misalignment of g
// in elements (so both sides of a full vector are 16-byte aligned).
template <typename T, bool kToShared>
__device__ __forceinline__ void move_dense(T* g, int n, T* buf, int phase,
                                           int lane) {
  constexpr int W = 16 / sizeof(T);
  // vector v covers run elements [v * W - phase, v * W - phase + W)
  const int vectors = (n + phase + W - 1) / W;
  for (int v0 = lane; v0 < vectors; v0 += 32 * kWireUnroll) {
    Vec16<T> pack[kWireUnroll];
    bool whole[kWireUnroll];
#pragma unroll
    for (int u = 0; u < kWireUnroll; ++u) {
      const int first = (v0 + 32 * u) * W - phase;
      whole[u] = v0 + 32 * u < vectors && first >= 0 && first + W <= n;
      if (whole[u])
        pack[u] = kToShared
                      ? *reinterpret_cast<const Vec16<T>*>(g + first)
                      : *reinterpret_cast<const Vec16<T>*>(buf + first + phase);
    }
#pragma unroll
    for (int u = 0; u < kWireUnroll; ++u) {
      const int v = v0 + 32 * u;
      if (v >= vectors) break;
      const int first = v * W - phase;
      if (whole[u]) {
        if (kToShared)
          *reinterpret_cast<Vec16<T>*>(buf + first + phase) = pack[u];
        else
          *reinterpret_cast<Vec16<T>*>(g + first) = pack[u];
      } else {      // head or tail of the run
        for (int k = 0; k < W; ++k) {
          const int j = first + k;
          if (j < 0 || j >= n) continue;
          if (kToShared)
            buf[j + phase] = g[j];
          else
            g[j] = buf[j + phase];
        }
      }
    }
  }
}

// Shared-memory bank rotation for the strided side.  Lane l holds word l of a
// run; its cell k lives (word bytes) * kBanks bytes after lane l - 1's cell k,
// so with every lane on the same k lanes a few apart share a bank
// (two banks of 16-bit cells: a 4-way conflict on every access, ~260 of the
// ~290 wavefronts a 4 KB row cost).  Round r therefore moves cell
// (r + group) mod N of each lane, group = the lane's index among the lanes it
// would collide with: every lane still moves each of its N cells once, and
// the lanes of one access spread over the banks.
template <int kWordBytes, int kStride>
__device__ __forceinline__ int conflict_group(int lane) {
  constexpr int kSpan = kWordBytes * kStride / 4;     // lane stride in banks
  constexpr int kPeriod = kSpan % 32 == 0  ? 1
                          : kSpan % 16 == 0 ? 2
                          : kSpan % 8 == 0  ? 4
                          : kSpan % 4 == 0  ? 8
                          : kSpan % 2 == 0  ? 16
                                            : 32;
  return lane / kPeriod;
}

template <typename T>
__device__ __forceinline__ T cell_of(uint64_t word, int k) {
  return static_cast<T>(word >> (k * 8 * static_cast<int>(sizeof(T))));
}

// Bank run g[0..count) <-> shared elements sm[e * stride], one 4- or 8-byte
// word of the run per lane and access (the head and tail that do not fill an aligned word
// move element-wise).
template <typename T, bool kToShared, int kStride>
__device__ __forceinline__ void move_strided(T* g, int count, T* sm, int lane) {
  constexpr int stride = kStride;
  // measured (capture r2j, blur 32768^2 u16, two banks): loads want 8-byte
  // words (unpack 4015 -> 4290 GB/s), stores 4-byte ones (pack 5642 vs 4938)
  constexpr int kWordBytes = (kToShared || sizeof(T) == 8) ? 8 : 4;
  constexpr int N = kWordBytes / static_cast<int>(sizeof(T));
  using Word = typename std::conditional<kWordBytes == 8, uint64_t,
                                         uint32_t>::type;
  union Cells {
    Word word;
    T cell[N];
  };
  const int misaligned = static_cast<int>(
      (reinterpret_cast<uintptr_t>(g) & (sizeof(Word) - 1)) / sizeof(T));
  const int head = min(count, (N - misaligned) % N);
  const int words = (count - head) / N;
  const int tail = count - head - words * N;
  if (lane < head) {
    if (kToShared) sm[lane * stride] = g[lane];
    else g[lane] = sm[lane * stride];
  }
  if (lane < tail) {
    const int e = head + words * N + lane;
    if (kToShared) sm[e * stride] = g[e];
    else g[e] = sm[e * stride];
  }
  Word* const body = reinterpret_cast<Word*>(g + head);
  T* const sm_body = sm + head * stride;
  // (pack reads the buffer in 4-byte words, a 2-way conflict at most, and
  // measured 2 % slower rotated: 5572 -> 5432 GB/s)
  const int group = kToShared ? conflict_group<kWordBytes, kStride>(lane) : 0;
  // loads: twice as many in flight as on the 16-byte side
  constexpr int kUnroll = kToShared ? 2 * kWireUnroll : kWireUnroll;
  for (int w0 = lane; w0 < words; w0 += 32 * kUnroll) {
    Cells cells[kUnroll];
    if (kToShared) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (w0 + 32 * u < words) cells[u].word = body[w0 + 32 * u];
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int w = w0 + 32 * u;
      if (w >= words) break;
      if (kToShared) {
#pragma unroll
        for (int r = 0; r < N; ++r) {
          const int k = (r + group) & (N - 1);
          sm_body[(w * N + k) * stride] = cell_of<T>(cells[u].word, k);
        }
      } else {
        uint64_t word = 0;
#pragma unroll
        for (int r = 0; r < N; ++r) {
          const int k = (r + group) & (N - 1);
          word |= static_cast<uint64_t>(sm_body[(w * N + k) * stride])
                  << (k * 8 * static_cast<int>(sizeof(T)));
        }
        body[w] = static_cast<Word>(word);
      }
    }
  }
}

template <typename T, bool kPack, int kBanks>
__global__ void __launch_bounds__(32 * kWireWarps, 4)
wire_kernel_staged(const __grid_constant__ soda_fpga_layout_t a, T* dense,
                   const __grid_constant__ Banks banks, long long rows) {
  extern __shared__ __align__(16) unsigned char staged_raw[];
  constexpr int W = 16 / sizeof(T);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // elements per row buffer: the row, its phase, rounded to whole vectors
  const int pitch = (a.tile_size[0] + 2 * W + W - 1) / W * W;
  T* const buf = reinterpret_cast<T*>(staged_raw) + warp * pitch;
  for (long long row = static_cast<long long>(blockIdx.x) * kWireWarps + warp;
       row < rows; row += static_cast<long long>(gridDim.x) * kWireWarps) {
    const Row r = decode_row<kPack>(a, row, blockIdx.y);
    if (!r.inside || r.i_hi <= r.i_lo) continue;       // the whole warp
    const int n = r.i_hi - r.i_lo;
    T* const row_dense = dense + r.original + r.i_lo;
    const int phase = static_cast<int>(
        (reinterpret_cast<uintptr_t>(row_dense) & 15) / sizeof(T));
    if (kPack) {
      move_dense<T, true>(row_dense, n, buf, phase, lane);
      __syncwarp();
    }
#pragma unroll
    for (int b = 0; b < kBanks; ++b) {
      // cells of this row whose stream element lands in bank slot b
      const long long first_o = r.stream + r.i_lo;
      const int skip =
          static_cast<int>(((b - first_o) % kBanks + kBanks) % kBanks);
      if (skip >= n) continue;
      const int count = (n - skip + kBanks - 1) / kBanks;
      T* const run = static_cast<T*>(banks.ptr[a.bank_vec[b]]) +
                     (first_o + skip) / kBanks;
      move_strided<T, !kPack, kBanks>(run, count, buf + phase + skip, lane);
    }
    __syncwarp();
    if (!kPack) move_dense<T, false>(row_dense, n, buf, phase, lane);
    __syncwarp();       // the buffer is reused by this warp's next row
  }
}

// ---- unpack, software-pipelined ------------------------------------------------
//
// In the staged kernel a warp alternates between its two sides: while it
// writes a row's dense run it has no bank loads in flight, and unpack — whose
// scattered side is the load side, 8 bytes per lane and access — ran at 68 %
// of the HBM rate where pack reaches 86 %.  Here a warp fetches the bank runs
// of its NEXT row into registers before it writes the current row's dense
// run: a whole row per warp stays in flight at all times.  The prefetched
// words stay in registers, so the warp still needs one row buffer.  Rows
// whose bank runs exceed the register budget (kWireFetchWords 8-byte words
// per lane over all banks: 4 KB of row) take the staged kernel.
// Measured (blur 32768^2 u16, tile 2000, two banks; capture r3j): staged
// 4464 GB/s; pipelined 4644; with the bank rotation below staged 4908,
// pipelined 5205 (79 % of the HBM rate).  Two blocks of 8 warps per SM: the
// prefetched row costs ~118 registers, and three blocks spill (3468).
constexpr int kWireFetchWords = 16;

template <typename T>
struct BankRun {
  const T* run;   // first cell of the row in this bank
  int count;      // cells of the row in this bank
  int skip;       // the row's first cell that lands in this bank
};

template <typename T, int kBanks>
__device__ __forceinline__ BankRun<T> bank_run(const soda_fpga_layout_t& a,
                                               const Banks& banks, const Row& r,
                                               int b) {
  const long long first_o = r.stream + r.i_lo;
  const int n = r.i_hi - r.i_lo;
  BankRun<T> g;
  g.skip = static_cast<int>(((b - first_o) % kBanks + kBanks) % kBanks);
  g.count = g.skip >= n ? 0 : (n - g.skip + kBanks - 1) / kBanks;
  g.run = static_cast<const T*>(banks.ptr[a.bank_vec[b]]) +
          (first_o + g.skip) / kBanks;
  return g;
}

// What a lane holds of one bank run: aligned 8-byte words `lane + 32 u` of
// the body, and one cell of the unaligned head and tail each.
template <typename T, int kWords>
struct Fetched {
  uint64_t body[kWords];
  T head, tail;
};

template <typename T>
struct RunSplit {
  int head, words, tail;
};

template <typename T>
__device__ __forceinline__ RunSplit<T> split_run(const T* g, int count) {
  constexpr int N = 8 / static_cast<int>(sizeof(T));
  const int misaligned =
      static_cast<int>((reinterpret_cast<uintptr_t>(g) & 7) / sizeof(T));
  RunSplit<T> s;
  s.head = min(count, (N - misaligned) % N);
  s.words = (count - s.head) / N;
  s.tail = count - s.head - s.words * N;
  return s;
}

template <typename T, int kWords>
__device__ __forceinline__ void fetch_run(const BankRun<T>& g, int lane,
                                          Fetched<T, kWords>* f) {
  constexpr int N = 8 / static_cast<int>(sizeof(T));
  const RunSplit<T> s = split_run(g.run, g.count);
  if (lane < s.head) f->head = g.run[lane];
  if (lane < s.tail) f->tail = g.run[s.head + s.words * N + lane];
  const uint64_t* const body =
      reinterpret_cast<const uint64_t*>(g.run + s.head);
#pragma unroll
  for (int u = 0; u < kWords; ++u)
    if (lane + 32 * u < s.words) f->body[u] = body[lane + 32 * u];
}

// The fetched cells into the row buffer: cell e of the run at sm[e * stride].
template <typename T, int kWords, int kStride>
__device__ __forceinline__ void place_run(const BankRun<T>& g, int lane,
                                          const Fetched<T, kWords>& f, T* sm) {
  constexpr int N = 8 / static_cast<int>(sizeof(T));
  constexpr int stride = kStride;
  const int group = conflict_group<8, kStride>(lane);
  const RunSplit<T> s = split_run(g.run, g.count);
  if (lane < s.head) sm[lane * stride] = f.head;
  if (lane < s.tail) sm[(s.head + s.words * N + lane) * stride] = f.tail;
  T* const sm_body = sm + s.head * stride;
#pragma unroll
  for (int u = 0; u < kWords; ++u) {
    const int w = lane + 32 * u;
    if (w < s.words) {
#pragma unroll
      for (int r = 0; r < N; ++r) {
        const int k = (r + group) & (N - 1);
        sm_body[(w * N + k) * stride] = cell_of<T>(f.body[u], k);
      }
    }
  }
}

template <typename T, int kBanks>
__global__ void __launch_bounds__(32 * kWireWarps, 2)
wire_unpack_pipelined(const __grid_constant__ soda_fpga_layout_t a, T* dense,
                      const __grid_constant__ Banks banks, long long rows) {
  extern __shared__ __align__(16) unsigned char staged_raw[];
  constexpr int W = 16 / sizeof(T);
  constexpr int kWords = kWireFetchWords / kBanks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pitch = (a.tile_size[0] + 2 * W + W - 1) / W * W;
  T* const buf = reinterpret_cast<T*>(staged_raw) + warp * pitch;
  const long long step = static_cast<long long>(gridDim.x) * kWireWarps;
  long long row = static_cast<long long>(blockIdx.x) * kWireWarps + warp;
  Row next;
  // the warp's next row with anything to move (uniform over the warp)
  auto advance = [&]() {
    for (; row < rows; row += step) {
      next = decode_row<false>(a, row, blockIdx.y);
      if (next.inside && next.i_hi > next.i_lo) return true;
    }
    return false;
  };
  if (!advance()) return;
  Fetched<T, kWords> fetched[kBanks];
#pragma unroll
  for (int b = 0; b < kBanks; ++b)
    fetch_run<T, kWords>(bank_run<T, kBanks>(a, banks, next, b), lane,
                         &fetched[b]);
  for (;;) {
    const Row r = next;
    const int n = r.i_hi - r.i_lo;
    T* const row_dense = dense + r.original + r.i_lo;
    const int phase = static_cast<int>(
        (reinterpret_cast<uintptr_t>(row_dense) & 15) / sizeof(T));
#pragma unroll
    for (int b = 0; b < kBanks; ++b) {
      const BankRun<T> g = bank_run<T, kBanks>(a, banks, r, b);
      place_run<T, kWords, kBanks>(g, lane, fetched[b], buf + phase + g.skip);
    }
    __syncwarp();
    row += step;
    const bool more = advance();
    if (more) {
#pragma unroll
      for (int b = 0; b < kBanks; ++b)
        fetch_run<T, kWords>(bank_run<T, kBanks>(a, banks, next, b), lane,
                             &fetched[b]);
    }
    move_dense<T, false>(row_dense, n, buf, phase, lane);
    __syncwarp();       // the buffer takes the next row
    if (!more) break;
  }
}

template <typename T, bool kPack>
int launch(const soda_fpga_layout_t& a, T* dense, const Banks& table,
           dim3 grid, cudaStream_t s) {
  // staged variant: one row buffer per warp.  SODA_FPGA_STAGED=0 / 1 picks
  // the element-wise / the staged kernel (default: staged where a row is long
  // enough for vectors to matter).
  constexpr size_t W = 16 / sizeof(T);
  const size_t pitch =
      (static_cast<size_t>(a.tile_size[0]) + 2 * W + W - 1) / W * W;
  const size_t smem = pitch * sizeof(T) * kWireWarps;
  const char* env = getenv("SODA_FPGA_STAGED");
  bool staged = a.tile_size[0] * sizeof(T) >= 512;
  if (env != nullptr && (env[0] == '0' || env[0] == '1')) staged = env[0] == '1';
  staged = staged && smem <= 200 * 1024;
  // unpack: the pipelined kernel where a row's bank runs fit its registers
  // (SODA_FPGA_PIPELINED=0: the staged kernel)
  const char* piped = getenv("SODA_FPGA_PIPELINED");
  const bool pipelined =
      !kPack && !(piped != nullptr && piped[0] == '0') &&
      (static_cast<size_t>(a.tile_size[0]) / a.num_bank + 2) * sizeof(T) <=
          static_cast<size_t>(kWireFetchWords / a.num_bank) * 256;
  const long long rows = grid.x;
  // enough blocks for every SM to hold its fill, each warp walking ~4 rows
  const long long wanted = (rows + kWireWarps * 4 - 1) / (kWireWarps * 4);
  dim3 sgrid(static_cast<unsigned>(std::max<long long>(1, wanted)), grid.y);
#define SODA_WIRE_LAUNCH(kBanks)                                              \
  if (staged && pipelined) {                                                  \
    auto fn = wire_unpack_pipelined<T, kBanks>;                               \
    if (smem > 48 * 1024 &&                                                   \
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             static_cast<int>(smem)) != cudaSuccess)          \
      return kLaunchFailed;                                                   \
    fn<<<sgrid, 32 * kWireWarps, smem, s>>>(a, dense, table, rows);           \
  } else if (staged) {                                                        \
    auto fn = wire_kernel_staged<T, kPack, kBanks>;                           \
    if (smem > 48 * 1024 &&                                                   \
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             static_cast<int>(smem)) != cudaSuccess)          \
      return kLaunchFailed;                                                   \
    fn<<<sgrid, 32 * kWireWarps, smem, s>>>(a, dense, table, rows);           \
  } else {                                                                    \
    wire_kernel<T, kPack, kBanks><<<grid, 256, 0, s>>>(a, dense, table);      \
  }
  switch (a.num_bank) {
    case 1: SODA_WIRE_LAUNCH(1) break;
    case 2: SODA_WIRE_LAUNCH(2) break;
    case 3: SODA_WIRE_LAUNCH(3) break;
    default: SODA_WIRE_LAUNCH(4) break;
  }
#undef SODA_WIRE_LAUNCH
  return 0;
}

template <bool kPack>
int run(const soda_fpga_layout_t* layout, void* dense,
        const void* const* banks, void* stream) {
  if (layout == nullptr || dense == nullptr || banks == nullptr) return kNull;
  const soda_fpga_layout_t& a = *layout;
  if (a.dim < 2 || a.dim > 4 || a.num_bank < 1 || a.num_bank > 4 ||
      a.tile_size_linearized <= 0)
    return kBadDescriptor;
  Banks table = {};
  for (int b = 0; b < a.num_bank; ++b) {
    const int bank = a.bank_vec[b];
    if (bank < 0 || bank > 3) return kBadDescriptor;
    if (banks[bank] == nullptr) return kNull;
    table.ptr[bank] = const_cast<void*>(banks[bank]);
  }
  long long tiles = 1, rows = a.dims[a.dim - 1];
  for (int d = 0; d < a.dim - 1; ++d) {
    if (a.tile_num[d] < 1 || a.tile_size[d] < 1 || a.tile_step[d] < 1)
      return kBadDescriptor;
    tiles *= a.tile_num[d];
    if (d > 0) rows *= a.tile_size[d];
  }
  if (tiles > 65535 || rows <= 0 || rows > 0x7fffffffLL) return kBadDescriptor;
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return kNoDevice;
  // grid: x = rows of a tile (in-tile coordinates 1..), y = tiles
  dim3 grid(static_cast<unsigned>(rows), static_cast<unsigned>(tiles));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = 0;
  switch (a.elem_size) {
    case 1: rc = launch<uint8_t, kPack>(a, static_cast<uint8_t*>(dense), table, grid, s); break;
    case 2: rc = launch<uint16_t, kPack>(a, static_cast<uint16_t*>(dense), table, grid, s); break;
    case 4: rc = launch<uint32_t, kPack>(a, static_cast<uint32_t*>(dense), table, grid, s); break;
    case 8: rc = launch<uint64_t, kPack>(a, static_cast<uint64_t*>(dense), table, grid, s); break;
    default: return kBadDescriptor;
  }
  if (rc != 0) return rc;
  return cudaGetLastError() == cudaSuccess ? 0 : kLaunchFailed;
}

}  // namespace

extern "C" int soda_fpga_pack(const soda_fpga_layout_t* layout,
                              const void* dense, void* const* banks,
                              void* stream) {
  return run<true>(layout, const_cast<void*>(dense), banks, stream);
}

extern "C" int soda_fpga_unpack(const soda_fpga_layout_t* layout, void* dense,
                                const void* const* banks, void* stream) {
  return run<false>(layout, dense, banks, stream);
}
