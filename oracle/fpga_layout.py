"""numpy restatement of the reference's FPGA wire-format loops — TEST
INFRASTRUCTURE (only tests/ may import it).

Follows the generated host code of the reference line by line
(src/soda/codegen/xilinx/host.py:629-686 pack, :823-901 unpack) on the
constants of ``soda.fpga_layout.WireLayout``.  Pinned by
tests/golden/fpga_layout/*.npz, which oracle/fpga_layout_ref.py produced by
compiling the unmodified reference's own loops.
"""
import itertools

import numpy as np


def _tiles(layout):
  return itertools.product(*[range(n) for n in reversed(layout.tile_num)])


def _tile_cells(layout, tile_index, lo, hi_margin):
  """In-tile coordinates (arrays, dim 0 first) of one tile's cells."""
  dim = layout.dim
  ranges = []
  for d in range(dim - 1):
    step = layout.tile_size[d] - layout.stencil_dim[d] + 1
    actual = (layout.dims[d] - step * tile_index[d]
              if tile_index[d] == layout.tile_num[d] - 1
              else layout.tile_size[d])                      # host.py:633-636
    ranges.append(np.arange(lo[d], actual - hi_margin[d]))
  ranges.append(np.arange(lo[dim - 1], layout.dims[dim - 1] -
                          hi_margin[dim - 1]))
  grids = np.meshgrid(*reversed(ranges), indexing='ij')[::-1]
  return [g.ravel().astype(np.int64) for g in grids]


def _offsets(layout, tile_index, cells, linearized, stream_offset):
  dim = layout.dim
  offset_in_tile, pitch = 0, 1
  for d in range(dim):                                        # host.py:650-653
    offset_in_tile = offset_in_tile + cells[d] * pitch
    if d < dim - 1:
      pitch *= layout.tile_size[d]
  tile_linear, pitch = 0, 1
  for d in range(dim - 1):                                    # host.py:668-672
    tile_linear += tile_index[d] * pitch
    pitch *= layout.tile_num[d]
  tiled = tile_linear * linearized + offset_in_tile + stream_offset
  original, pitch = 0, 1
  for d in range(dim):
    step = (layout.tile_size[d] - layout.stencil_dim[d] + 1
            if d < dim - 1 else 0)
    coord = cells[d] + (tile_index[d] * step if d < dim - 1 else 0)
    original = original + coord * pitch                       # host.py:659-666
    pitch *= layout.dims[d]
  return tiled, original


def pack(layout, name, dense, bank_arrays):
  """Scatter ``dense`` (shape = reversed dims) into ``bank_arrays`` (by bank
  id), in place; untouched elements keep their value."""
  banks = layout.banks(name)
  flat = dense.ravel()
  zero = [0] * layout.dim
  for reversed_index in _tiles(layout):
    tile_index = reversed_index[::-1]
    cells = _tile_cells(layout, tile_index, zero, zero)
    tiled, original = _offsets(layout, tile_index, cells,
                               layout.tile_size_linearized_i, 0)
    for slot, bank in enumerate(banks):                       # host.py:680-684
      mine = tiled % len(banks) == slot
      bank_arrays[bank][tiled[mine] // len(banks)] = flat[original[mine]]


def unpack(layout, name, dense, bank_arrays):
  """Gather the valid cells of every tile from ``bank_arrays`` into ``dense``,
  in place; other cells keep their value."""
  banks = layout.banks(name)
  flat = dense.reshape(-1)
  for reversed_index in _tiles(layout):
    tile_index = reversed_index[::-1]
    cells = _tile_cells(layout, tile_index, layout.window_offset,
                        layout.valid_hi_margin())
    tiled, original = _offsets(layout, tile_index, cells,
                               layout.tile_size_linearized_o,
                               layout.stream_offset[name])
    for slot, bank in enumerate(banks):                       # host.py:890-893
      mine = tiled % len(banks) == slot
      flat[original[mine]] = bank_arrays[bank][tiled[mine] // len(banks)]
