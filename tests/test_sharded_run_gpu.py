"""One run on host buffers spread over several devices behind the plugin call
(`<app>(..., "devices=0,1")` / soda_cuda_run / SODA_CUDA_DEVICES): bit-identical
to the CPU oracle, whole arrays.

The slabs never talk to each other — each holds ghost rows of the whole run's
reach — so a device list may name one device several times: "0,0,0" exercises
the slab arithmetic (ghost rows, trimmed launches, copy-back ranges, valid
boxes in local coordinates, one host thread per slab) on a single-GPU box;
with more GPUs visible the same cases also run across them.
"""
import numpy as np
import pytest

import common
import random_programs as rp
from soda import cuda as soda_cuda

pytestmark = pytest.mark.gpu

CASES = [
    # name, iterate, dims: several launches with a remainder; one-sided
    # window; two inputs; 3-D with a 4-launch chain
    ('jacobi2d', 5, (1024, 700)), ('blur', 1, (1000, 333)),
    ('denoise2d', 1, (512, 300)), ('heat3d', 4, (128, 64, 90)),
    ('seidel2d', 2, (517, 211)),
]


def _device_lists():
  import torch
  lists = ['0,0', '0,0,0']
  count = torch.cuda.device_count()
  if count > 1:
    lists.append(','.join(str(d) for d in range(min(count, 4))))
    lists.append('all')
  return lists


@pytest.mark.parametrize('name,iterate,dims', CASES)
def test_sharded_run_equals_oracle(name, iterate, dims, monkeypatch):
  orc = common.oracle(name, iterate)
  library = soda_cuda.compile_stencil(common.stencil(name, iterate))
  inputs = common.random_inputs(orc, dims, seed=11)
  want = orc.run(inputs)
  for pieces in ('1', '3'):
    monkeypatch.setenv('SODA_CUDA_PIECES', pieces)
    for devices in _device_lists():
      got = library.run(inputs, devices=devices)
      for k, (g, w) in enumerate(zip(got, want)):
        common.assert_bit_exact(g, w, '%s x%d on devices %s, %s piece(s), '
                                'output %d' % (name, iterate, devices, pieces,
                                               k), any_nan=name == 'denoise2d')
      slabs = library.slab_stats
      assert len(slabs) == len(library.shard_plan(
          dims, len(_expand(devices))))
      assert all(s['launches'] >= 1 for s in slabs)
      assert library.stats['launches'] == sum(s['launches'] for s in slabs)


def _expand(devices):
  import torch
  if devices == 'all':
    return list(range(torch.cuda.device_count()))
  return devices.split(',')


def test_sharded_multi_output_program(monkeypatch):
  """Outputs with different valid boxes, an output read by a later one."""
  import golden
  stencil = rp.extra_stencil('chain2')
  dims = rp.EXTRA['chain2'][1]
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil)
  inputs = common.random_inputs(orc, dims, seed=3)
  want = orc.run(inputs)
  monkeypatch.setenv('SODA_CUDA_PIECES', '2')
  got = library.run(inputs, devices='0,0,0')
  for k, (g, w) in enumerate(zip(got, want)):
    common.assert_bit_exact(g, w, 'chain2 sharded, output %d' % k)


def test_device_list_from_the_environment_and_errors(monkeypatch):
  library = soda_cuda.compile_stencil(common.stencil('jacobi2d', 3))
  orc = common.oracle('jacobi2d', 3)
  inputs = common.random_inputs(orc, (512, 200), seed=5)
  want = orc.run(inputs)
  monkeypatch.setenv('SODA_CUDA_DEVICES', '0,0')
  got = library.run(inputs)
  common.assert_bit_exact(got[0], want[0], 'SODA_CUDA_DEVICES=0,0')
  assert len(library.slab_stats) == 2
  monkeypatch.delenv('SODA_CUDA_DEVICES')
  got = library.run(inputs)
  common.assert_bit_exact(got[0], want[0], 'one device')
  assert library.slab_stats == []
  with pytest.raises(soda_cuda.CudaError):
    library.run(inputs, devices='0,99')


@pytest.mark.parametrize('name,iterate,dims', [
    ('jacobi2d', 3, (1536, 200)), ('heat3d', 2, (131, 35, 52)),
    ('blur', 1, (2000, 1000))])
def test_reference_harness_accepts_a_sharded_run(name, iterate, dims):
  """The reference's own generated `<app>_test` (unmodified, oracle/_ref/)
  calls the CUDA library where the FPGA would run; here that one call is
  spread over three slabs.  Its golden loop finds no mismatch."""
  import os
  import ref_harness
  lib = ref_harness.lib_path(name, iterate)
  if not os.path.exists(lib):
    pytest.skip('oracle/_ref was not built (needs /root/reference)')
  stencil = common.stencil(name, iterate)
  harness = ref_harness.RefHarness(lib, stencil)
  library = soda_cuda.compile_stencil(stencil)
  assert harness.test(
      dims, lambda inputs: library.run(inputs, devices='0,0,0')) == 0
  assert len(library.slab_stats) == len(library.shard_plan(dims, 3))
