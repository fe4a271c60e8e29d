#!/usr/bin/env python3
"""Device-resident timing of the FPGA wire-format kernels (not the contract
bench): blur 32768 x 32768 uint16, tile 2000, two banks each way."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'soda-compiler_b200')]

import numpy as np   # noqa: E402
import torch         # noqa: E402

from soda import core, fpga_layout   # noqa: E402


def main():
  dims = (32768, 32768)
  with open(os.path.join(ROOT, 'benchmarks', 'blur.soda')) as handle:
    stencil = core.Stencil.from_text(handle.read(), tile_size=[2000])
  stencil.input_stmts[0].dram = (0, 1)
  stencil.output_stmts[0].dram = (2, 3)
  layout = fpga_layout.WireLayout(stencil, dims)
  name_in, name_out = stencil.input_names[0], stencil.output_names[0]
  dense = torch.randint(-30000, 30000, tuple(reversed(dims)),
                        dtype=torch.int16, device='cuda')
  banks_in = {b: torch.zeros(layout.bank_elems(name_in), dtype=torch.int16,
                             device='cuda') for b in layout.banks(name_in)}
  banks_out = {b: torch.zeros(layout.bank_elems(name_out), dtype=torch.int16,
                              device='cuda') for b in layout.banks(name_out)}
  out = torch.zeros_like(dense)
  cells = float(np.prod(dims))
  for what, call in (
      ('pack', lambda: fpga_layout.pack(layout, name_in, dense, banks_in)),
      ('unpack', lambda: fpga_layout.unpack(layout, name_out, out,
                                            banks_out)),
      ('pack staged', lambda: fpga_layout.pack(layout, name_in, dense,
                                               banks_in)),
      ('unpack staged', lambda: fpga_layout.unpack(layout, name_out, out,
                                                   banks_out)),
      ('unpack piped', lambda: fpga_layout.unpack(layout, name_out, out,
                                                  banks_out))):
    os.environ['SODA_FPGA_STAGED'] = '0' if what in ('pack', 'unpack') else '1'
    os.environ['SODA_FPGA_PIPELINED'] = '1' if 'piped' in what else '0'
    for _ in range(3):
      call()
    torch.cuda.synchronize()
    times = []
    for _ in range(7):
      start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
      start.record()
      call()
      stop.record()
      torch.cuda.synchronize()
      times.append(start.elapsed_time(stop))
    ms = float(np.median(times))
    print('%-13s blur %dx%d uint16 tile 2000, 2 banks: %.3f ms, %.0f GB/s '
          '(2 B read + 2 B written per cell), tiles %s' % (
              what, dims[0], dims[1], ms, cells * 4 / ms / 1e6,
              layout.tile_num), flush=True)


if __name__ == '__main__':
  main()
