"""`half` tensors: the oracle's semantics against a numpy restatement, and
the emitted kernels compile (CPU); GPU parity is tests/test_half_gpu.py.

Semantics (the reference cannot compile `half` on the host, so
this is a definition, not a reproduction): binary16 in memory; a read
converts to float; expressions evaluate by the C++ rules on float (double,
int) operands; a store rounds once to nearest even from the expression's type.
"""
import subprocess

import numpy as np
import pytest

import golden
import half_programs as hp
from soda import cuda as soda_cuda
from soda.codegen import cuda as codegen
from soda.codegen.cuda import plan

f32 = np.float32


def test_oracle_half_semantics_match_numpy():
  """halfmid, restated with numpy float32 arithmetic and one float16 cast."""
  stencil = hp.stencil_of('halfmid')
  orc = golden.Oracle(stencil)
  rng = np.random.default_rng(4)
  a = (rng.random((60, 96)) * 50).astype(f32)
  got = orc.run([a])[0]
  # m(x, y) = half(a(x, y) * 0.3f + a(x, y + 1) * 0.7f): float ops, no FMA
  m = (a[:-1, :] * f32(0.3) + a[1:, :] * f32(0.7)).astype(np.float16)
  mf = m.astype(f32)
  # o(x, y) = m(x, y) + m(x + 1, y) + m(x - 1, y) * 0.125f
  want = (mf[:, 1:-1] + mf[:, 2:]) + mf[:, :-2] * f32(0.125)
  np.testing.assert_array_equal(got[:-1, 1:-1].view(np.uint32),
                                want.view(np.uint32))


def test_oracle_rounds_once_from_double():
  stencil = hp.stencil_of('halfdbl')
  orc = golden.Oracle(stencil)
  rng = np.random.default_rng(5)
  a = rng.random((30, 48)) * 1000
  got = orc.run([a])[0]
  want = ((a[:-1, :-1] * 0.333 + a[:-1, 1:] * 1.0001) + a[1:, :-1]).astype(
      np.float16)
  np.testing.assert_array_equal(got[:-1, :-1].view(np.uint16),
                                want.view(np.uint16))


@pytest.mark.parametrize('name', sorted(hp.PROGRAMS))
def test_half_programs_plan_and_compile(name, tmp_path, monkeypatch):
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  stencil = hp.stencil_of(name)
  program = plan.extract_program(stencil)
  codegen.check_supported(program)
  golden.build(stencil)
  _, kernel, _ = soda_cuda.generate_sources(stencil)
  path = tmp_path / 'k.cu'
  path.write_text(kernel)
  done = subprocess.run(
      ['nvcc'] + soda_cuda.ARCH_FLAGS + [
          '-std=c++17', '-fmad=false', '-I', soda_cuda.CSRC_DIR, '-I',
          soda_cuda.INCLUDE_DIR, '-c', str(path), '-o', str(tmp_path / 'k.o')],
      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  assert done.returncode == 0, done.stdout[-3000:]


def test_library_reports_half_tensors():
  library = soda_cuda.compile_stencil(hp.stencil_of('halfblur'))
  assert library.inputs == [('a', 'half')] and library.outputs == [('o', 'half')]
  with pytest.raises(TypeError):
    library.run([np.zeros((64, 64), dtype=np.float32)])
