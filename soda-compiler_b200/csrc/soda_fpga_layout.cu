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

#include "soda_fpga_layout.h"

namespace {

enum { kNull = -12, kBadDescriptor = -4, kNoDevice = -19, kLaunchFailed = -23 };

struct Banks {
  void* ptr[4];
};

// One block per tile row (fixed in-tile coordinates in every dimension but 0):
// the row is decoded once per block, threads walk dimension 0, four cells
// each per trip, and the bank count is a compile-time constant — no per-cell
// division by a run-time value.
template <typename T, bool kPack, int kBanks>
__global__ void __launch_bounds__(256)
wire_kernel(const __grid_constant__ soda_fpga_layout_t a, T* dense,
            const __grid_constant__ Banks banks) {
  const int last = a.dim - 1;
  // this block's tile: dimension 0 of the tile index is the fastest
  int tile_index[3] = {0, 0, 0};
  int rest = blockIdx.y;
  const long long tile_linear = rest;
  int extent0 = 0;
  bool inside = true;
  long long original = 0, pitch = 1, row_offset = 0, in_tile_pitch = 1;
  // in-tile coordinates of this row in dimensions 1..last
  long long row = blockIdx.x;
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
  if (!inside) return;
  int i_lo = a.lo[0], i_hi = extent0 - a.hi_margin[0];
  if (!kPack && tile_index[0] + 1 < a.tile_num[0])
    i_hi = min(i_hi, a.lo[0] + a.tile_step[0]);   // the next tile owns the rest
  const long long stream =
      tile_linear * a.tile_size_linearized + row_offset + a.stream_offset;
  T* const row_dense = dense + original;
  T* bank[kBanks];
#pragma unroll
  for (int b = 0; b < kBanks; ++b)
    bank[b] = static_cast<T*>(banks.ptr[a.bank_vec[b]]);
  for (int i0 = i_lo + threadIdx.x * 4; i0 < i_hi; i0 += 1024) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k;
      if (i >= i_hi) break;
      const unsigned long long o = static_cast<unsigned long long>(stream + i);
      T* const slot = bank[o % kBanks] + o / kBanks;
      if (kPack)
        *slot = row_dense[i];
      else
        row_dense[i] = *slot;
    }
  }
}

template <typename T, bool kPack>
void launch(const soda_fpga_layout_t& a, T* dense, const Banks& table,
            dim3 grid, cudaStream_t s) {
  switch (a.num_bank) {
    case 1: wire_kernel<T, kPack, 1><<<grid, 256, 0, s>>>(a, dense, table); break;
    case 2: wire_kernel<T, kPack, 2><<<grid, 256, 0, s>>>(a, dense, table); break;
    case 3: wire_kernel<T, kPack, 3><<<grid, 256, 0, s>>>(a, dense, table); break;
    default: wire_kernel<T, kPack, 4><<<grid, 256, 0, s>>>(a, dense, table); break;
  }
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
  switch (a.elem_size) {
    case 1: launch<uint8_t, kPack>(a, static_cast<uint8_t*>(dense), table, grid, s); break;
    case 2: launch<uint16_t, kPack>(a, static_cast<uint16_t*>(dense), table, grid, s); break;
    case 4: launch<uint32_t, kPack>(a, static_cast<uint32_t*>(dense), table, grid, s); break;
    case 8: launch<uint64_t, kPack>(a, static_cast<uint64_t*>(dense), table, grid, s); break;
    default: return kBadDescriptor;
  }
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
