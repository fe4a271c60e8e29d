"""CPU model of csrc/soda_fpga_layout.cu's `wire_kernel` (test helper): the
same block/row decomposition and index arithmetic, cell by cell in Python, so
the kernel's logic is checked where there is no GPU."""
import numpy as np


def run(desc, dense, banks, pack):
  """desc: soda.fpga_layout.TensorLayout; dense: flat array; banks: {id: flat
  array}.  Mirrors run<kPack>() + wire_kernel<T, kPack, kBanks>()."""
  dim, last = desc.dim, desc.dim - 1
  tiles, rows = 1, desc.dims[last]
  for d in range(last):
    tiles *= desc.tile_num[d]
    if d > 0:
      rows *= desc.tile_size[d]
  for block_y in range(tiles):
    for block_x in range(rows):
      tile_index = [0, 0, 0]
      rest, row = block_y, block_x
      inside = True
      original, pitch, row_offset, in_tile_pitch = 0, 1, 0, 1
      extent0 = 0
      for d in range(dim):
        extent = desc.dims[d]
        if d < last:
          tile_index[d] = rest % desc.tile_num[d]
          rest //= desc.tile_num[d]
          extent = (desc.dims[d] - desc.tile_step[d] * tile_index[d]
                    if tile_index[d] == desc.tile_num[d] - 1
                    else desc.tile_size[d])
        if d == 0:
          extent0 = extent
          original += tile_index[0] * desc.tile_step[0]
        else:
          if d < last:
            c = row % desc.tile_size[d]
            row //= desc.tile_size[d]
          else:
            c = row
          inside = inside and desc.lo[d] <= c < extent - desc.hi_margin[d]
          if (not pack and d < last and
              tile_index[d] + 1 < desc.tile_num[d] and
              c - desc.tile_step[d] >= desc.lo[d]):
            inside = False
          coord = c + (tile_index[d] * desc.tile_step[d] if d < last else 0)
          original += coord * pitch
          row_offset += c * in_tile_pitch
        pitch *= desc.dims[d]
        if d < last:
          in_tile_pitch *= desc.tile_size[d]
      if not inside:
        continue
      i_lo, i_hi = desc.lo[0], extent0 - desc.hi_margin[0]
      if not pack and tile_index[0] + 1 < desc.tile_num[0]:
        i_hi = min(i_hi, desc.lo[0] + desc.tile_step[0])
      stream = (block_y * desc.tile_size_linearized + row_offset +
                desc.stream_offset)
      if i_hi <= i_lo:
        continue
      i = np.arange(i_lo, i_hi)
      o = stream + i
      for slot in range(desc.num_bank):
        mine = o % desc.num_bank == slot
        bank = banks[desc.bank_vec[slot]]
        if pack:
          bank[o[mine] // desc.num_bank] = dense[original + i[mine]]
        else:
          dense[original + i[mine]] = bank[o[mine] // desc.num_bank]
