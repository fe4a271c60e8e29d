"""Four-dimensional programs: schedules on the CPU models, kernels through
nvcc (CPU); GPU parity is tests/test_dim4_gpu.py."""
import subprocess

import pytest

import dim4_programs as d4
import golden
import reg_schedule_sim as reg_sim
import schedule_sim as sim
from soda import cuda as soda_cuda
from soda.codegen import cuda as codegen
from soda.codegen.cuda import plan


@pytest.mark.parametrize('name,dims', d4.CASES)
def test_four_dimensional_schedules(name, dims, monkeypatch):
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  program = plan.extract_program(d4.stencil_of(name))
  assert program.dim == 4
  small = (140, 20, 9, 7)
  for style in ('reg', 'ring'):
    for sched in codegen.make_schedules(program, codegen.Options(style=style)):
      runner = reg_sim if sched.style == 'reg' else sim
      outs = runner.run_schedule(sched, small, 5)
      runner.check_outputs(sched, small, outs)


@pytest.mark.parametrize('name,dims', d4.CASES)
def test_four_dimensional_kernels_compile(name, dims, tmp_path, monkeypatch):
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  stencil = d4.stencil_of(name)
  golden.build(stencil)
  _, kernel, _ = soda_cuda.generate_sources(stencil)
  assert 'cp.async.bulk.tensor' not in kernel    # in the header, not emitted
  assert 'soda::tma_load(' in kernel
  path = tmp_path / 'k.cu'
  path.write_text(kernel)
  done = subprocess.run(
      ['nvcc'] + soda_cuda.ARCH_FLAGS + [
          '-std=c++17', '-fmad=false', '-I', soda_cuda.CSRC_DIR, '-I',
          soda_cuda.INCLUDE_DIR, '-c', str(path), '-o', str(tmp_path / 'k.o')],
      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  assert done.returncode == 0, done.stdout[-3000:]
