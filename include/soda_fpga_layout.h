/* FPGA wire format of SODA tensors <-> dense arrays, on the GPU (C ABI).
 *
 * The reference's generated OpenCL host keeps the tensors it exchanges with
 * the FPGA kernel in tiled, burst-aligned, bank-interleaved buffers: it packs
 * every input (reference src/soda/codegen/xilinx/host.py:629-686) and unpacks
 * every output (:823-901) on the CPU.  These two entry points are those loop
 * nests as sm_100a kernels on device-resident memory, for callers that hold
 * data in that format (an existing xclbin host, a capture of its DMA
 * buffers).  The descriptor is filled by soda/fpga_layout.py (`WireLayout`),
 * which restates the reference's layout formulas (host.py:262-264, 334-347,
 * 399-415, 868-877).
 *
 * Both calls are asynchronous on `stream` (a cudaStream_t; NULL = default
 * stream), touch only elements that correspond to a grid cell (burst padding
 * and the cut-off part of the last tile keep their content, as in the
 * reference), and return 0 or a negative Halide error code (-12 null
 * argument, -4 bad descriptor, -19 no CUDA device, -23 launch failure).
 */
#ifndef SODA_FPGA_LAYOUT_H_
#define SODA_FPGA_LAYOUT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct soda_fpga_layout_t {
  int32_t dim;              /* 2..4; the last dimension is streamed, not tiled */
  int32_t elem_size;        /* bytes: 1, 2, 4 or 8 */
  int32_t dims[4];          /* grid extents, dimension 0 first */
  int32_t tile_size[4];     /* TILE_SIZE_DIM_d, d < dim - 1 */
  int32_t tile_num[4];      /* tiles per dimension (host.py:262-264) */
  int32_t tile_step[4];     /* TILE_SIZE_DIM_d - STENCIL_DIM_d + 1 */
  int32_t lo[4];            /* first in-tile coordinate moved (outputs: the
                               window offset, host.py:838-852; inputs: 0) */
  int32_t hi_margin[4];     /* coordinates skipped at the high end */
  int32_t num_bank;         /* 1..4 */
  int32_t bank_vec[4];      /* DRAM bank of stream element o: bank_vec[o % n] */
  int64_t tile_size_linearized;  /* elements per tile incl. burst padding */
  int64_t stream_offset;    /* outputs: lag of the output stream (host.py:
                               868-877); inputs: 0 */
} soda_fpga_layout_t;

/* The function picks the direction, the descriptor the mapping.  With an
 * input's descriptor soda_fpga_pack is the reference host's input loop nest
 * (host.py:629-686) and soda_fpga_unpack its inverse (what a stand-in for the
 * FPGA kernel does first); with an output's descriptor soda_fpga_unpack is
 * the host's output loop nest (host.py:823-901) and soda_fpga_pack its
 * inverse (what the stand-in does last).
 *
 * dense (device) -> bank buffers (device, indexed by bank id 0..3; unused
 * banks may be NULL). */
int soda_fpga_pack(const soda_fpga_layout_t* layout, const void* dense,
                   void* const* banks, void* stream);

/* bank buffers -> the cells of dense the mapping covers.  Where the cell
 * ranges of neighbouring tiles overlap, the later tile's value is taken, as
 * the reference's ascending tile loops leave it. */
int soda_fpga_unpack(const soda_fpga_layout_t* layout, void* dense,
                     const void* const* banks, void* stream);

#ifdef __cplusplus
}
#endif

#endif  /* SODA_FPGA_LAYOUT_H_ */
