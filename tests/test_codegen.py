"""sodac --cuda-* : the plugin API and the emitted text (no GPU needed)."""
import os
import subprocess
import sys

import pytest

import common
from haoda import util
from soda import core
from soda.codegen import cuda as codegen
from soda.codegen.cuda import plan

SODAC = os.path.join(common.ROOT, 'soda-compiler_b200', 'sodac')


def sodac(*args, stdin=None):
  return subprocess.run([sys.executable, SODAC] + list(args), input=stdin,
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                        text=True, check=False)


def test_kernel_and_host_to_files_and_stdout(tmp_path):
  kernel, host = tmp_path / 'k.cu', tmp_path / 'h.cpp'
  done = sodac(common.bench_path('blur'), '--cuda-kernel', str(kernel),
               '--cuda-host', str(host), '--cuda-header', '-')
  assert done.returncode == 0, done.stderr
  assert 'int blur(buffer_t *var_input_buffer, buffer_t *var_blur_y_buffer, ' \
         'const char* xclbin)' in done.stdout
  text = kernel.read_text()
  assert '__global__ void __launch_bounds__' in text
  # 2-D programs stream through registers: TMA input queue, warp shuffles
  assert 'soda::tma_load(queue' in text and 'soda::shfl_down<' in text
  # the reference-lowered expressions, operands mapped to register histories
  assert '/ 3)' in text and 'r[0] = soda::store_cast<uint16_t>((' in text
  ring = tmp_path / 'ring.cu'
  done = sodac(common.bench_path('blur'), '--cuda-kernel', str(ring),
               '--cuda-style', 'ring')
  assert done.returncode == 0, done.stderr
  text = ring.read_text()
  assert 'soda::tma_load(' in text and 'soda::mbar_wait(' in text
  assert text.count('r[k] = soda::store_cast<uint16_t>((') == 2
  assert 'int blur(' in host.read_text()
  assert 'soda_cuda_run' in host.read_text()


def test_stdin_and_iterate_override():
  done = sodac('-', '--iterate', '6', '--cuda-temporal-depth', '4',
               '--cuda-kernel', '-', stdin=common.bench_text('jacobi2d'))
  assert done.returncode == 0, done.stderr
  assert 'soda_jacobi2d_d4(' in done.stdout      # main depth
  assert 'soda_jacobi2d_d2(' in done.stdout      # remainder 6 % 4


def test_fast_and_exact_switches():
  """--cuda-fast marks the kernel file for the tolerance-tested build;
  --cuda-exact (the default) does not."""
  fast = sodac(common.bench_path('denoise2d'), '--cuda-fast', '--cuda-kernel',
               '-')
  assert fast.returncode == 0, fast.stderr
  assert '#define SODA_CUDA_FAST_MATH 1' in fast.stdout
  assert '-prec-div=false' in fast.stdout
  for flags in ([], ['--cuda-exact']):
    exact = sodac(common.bench_path('denoise2d'), '--cuda-kernel', '-', *flags)
    assert exact.returncode == 0, exact.stderr
    assert 'SODA_CUDA_FAST_MATH' not in exact.stdout
    assert '-fmad=false' in exact.stdout


def test_errors_exit_1():
  assert sodac('-', '--cuda-kernel', '-', stdin='kernel: broken').returncode == 1
  done = sodac(common.bench_path('denoise2d'), '--iterate', '2',
               '--cuda-kernel', '-')
  assert done.returncode == 1
  assert 'number of input tensors must be the same as output' in done.stderr


def test_unsupported_programs_raise_semantic_error():
  text = ('kernel: k\nburst width: 64\nunroll factor: 1\niterate: 1\n'
          'input %s: a(8, *)\noutput %s: b(0, 0) = a(0, 0)\n')
  with pytest.raises(util.SemanticError):      # ap_int width
    codegen.make_schedules(plan.extract_program(
        core.Stencil.from_text(text % ('int5', 'int5'))))
  one_d = ('kernel: k\nburst width: 64\nunroll factor: 1\niterate: 1\n'
           'input float: a\noutput float: b(0) = a(0) + a(1)\n')
  with pytest.raises(util.SemanticError):
    codegen.make_schedules(plan.extract_program(core.Stencil.from_text(one_d)))


def test_calls_are_routed_through_exact_math_wrappers():
  import io
  program = plan.extract_program(common.stencil('denoise2d'))
  out = io.StringIO()
  codegen.print_kernel(program, codegen.make_schedules(program), out)
  assert 'soda_fn_sqrt(' in out.getvalue()
  assert ' sqrt(' not in out.getvalue().replace('soda_fn_sqrt(', '')


def test_print_code_does_not_modify_the_stencil():
  import argparse
  stencil = common.stencil('seidel2d', 2)
  before = [str(t) for t in stencil.tensors.values()]
  args = argparse.Namespace(cuda_kernel_file=os.devnull,
                            cuda_host_file=os.devnull)
  codegen.print_code(stencil, args)
  assert [str(t) for t in stencil.tensors.values()] == before


@pytest.mark.skipif(not common.have_reference(),
                    reason='needs /root/reference')
def test_backend_accepts_the_reference_stencil_object(tmp_path):
  """print_code on the UNMODIFIED reference's Stencil emits the same kernel
  as on this repo's Stencil (run in a subprocess: package names collide)."""
  script = tmp_path / 'via_reference.py'
  script.write_text('''
import importlib.util, os, sys, collections, collections.abc
root = %r
for n in ("Iterable", "Mapping"):
    setattr(collections, n, getattr(collections.abc, n))
sys.path[:0] = [os.path.join(root, "oracle", "refshim"), "/root/reference/src"]
sys.path.append(os.path.join(root, "oracle"))
import ref_tool
stencil = ref_tool.load_stencil(sys.argv[1], int(sys.argv[2]))
# load this repo's backend under private names next to the reference packages
def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec); sys.modules[name] = mod
    spec.loader.exec_module(mod); return mod
pkg = os.path.join(root, "soda-compiler_b200", "soda", "codegen", "cuda")
import types
sys.modules.setdefault("soda.codegen.cuda", types.ModuleType("soda.codegen.cuda"))
plan = load("soda.codegen.cuda.plan", os.path.join(pkg, "plan.py"))
kernel = load("soda.codegen.cuda.kernel", os.path.join(pkg, "kernel.py"))
host = load("soda.codegen.cuda.host", os.path.join(pkg, "host.py"))
backend = load("soda.codegen.cuda", os.path.join(pkg, "__init__.py"))
program = plan.extract_program(stencil)
backend.print_kernel(program, backend.make_schedules(program), sys.stdout)
''' % common.ROOT)
  for name, iterate in (('sobel2d', 1), ('jacobi2d', 4), ('denoise3d', 1)):
    via_reference = subprocess.run(
        [sys.executable, str(script),
         os.path.join(common.REFERENCE_DIR, 'tests', 'src', name + '.soda'),
         str(iterate)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
        text=True, check=False)
    assert via_reference.returncode == 0, via_reference.stderr[-2000:]
    ours = sodac(common.bench_path(name), '--iterate', str(iterate),
                 '--cuda-kernel', '-')
    assert ours.returncode == 0
    assert via_reference.stdout == ours.stdout


def test_random_programs_emit_code_nvcc_accepts(tmp_path, monkeypatch):
  """Beyond the benchmarks: the kernels emitted for seeded random programs
  (mixed types, 2-D/3-D, fused iterations) compile for sm_100a, and the
  oracle of each builds and runs."""
  import subprocess
  import numpy as np
  import golden
  import random_programs as rp
  from soda import cuda as soda_cuda
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  for seed in rp.SEEDS[:4]:
    stencil = rp.stencil_of(seed)
    _, kernel, _ = soda_cuda.generate_sources(stencil)
    path = tmp_path / ('k%d.cu' % seed)
    path.write_text(kernel)
    done = subprocess.run(
        ['nvcc'] + soda_cuda.ARCH_FLAGS + [
            '-std=c++17', '-fmad=false', '-I', soda_cuda.CSRC_DIR, '-I',
            soda_cuda.INCLUDE_DIR, '-c', str(path), '-o',
            str(tmp_path / 'k.o')],
        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert done.returncode == 0, rp.program_text(seed) + done.stdout[-2000:]
    orc = golden.Oracle(stencil)
    dims = (40, 30) if stencil.dim == 2 else (24, 20, 18)
    out = orc.run(common.random_inputs(orc, dims, seed=seed))
    assert out[0].shape == tuple(reversed(dims))


def test_let_right_hand_sides_keep_their_parentheses(tmp_path, monkeypatch):
  """The IR's `unparenthesize` is not bracket-matching (reference
  src/haoda/ir/__init__.py:877-881): applied to a let it turns
  `(a == b) & (c)` into `a == b) & (c`.  The golden loop does not apply it
  (host.py:1111-1114); neither may the kernels.  Seeds whose oracle compiles
  must give kernels nvcc accepts."""
  import subprocess
  import golden
  import expression_programs as ep
  from soda import cuda as soda_cuda
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  text = ('kernel: lets\nburst width: 64\nunroll factor: 1\niterate: 1\n'
          'input int16: a(8, *)\n'
          'output int16: v = (a(0, 0) == a(1, 0)) & (a(0, 1) == 3) '
          'o(0, 0) = a(0, 0) + v\n')
  _, kernel, _ = soda_cuda.generate_sources(core.Stencil.from_text(text))
  assert ') & (' in kernel and '== 3);' in kernel
  for seed in (3, 55):          # found by the random search: lets with `&`
    stencil = core.Stencil.from_text(ep.program(seed))
    golden.build(stencil, build_dir=str(tmp_path))
    _, kernel, _ = soda_cuda.generate_sources(stencil)
    path = tmp_path / ('k%d.cu' % seed)
    path.write_text(kernel)
    done = subprocess.run(
        ['nvcc'] + soda_cuda.ARCH_FLAGS + [
            '-std=c++17', '-fmad=false', '-I', soda_cuda.CSRC_DIR, '-I',
            soda_cuda.INCLUDE_DIR, '-c', str(path), '-o',
            str(tmp_path / 'k.o')],
        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert done.returncode == 0, ep.program(seed) + done.stdout[-1500:]


def test_top_level_quotients_are_batched_per_vector():
  """A float statement that is a quotient at its top is emitted as pairs,
  soda::div_try on each pair, and the plain quotients under one `if (rare)`
  per vector (kernel_reg.batches_rare_paths); the numerator of a chain
  `a * b / c` is `(a * b)`."""
  import io
  program = plan.extract_program(common.stencil('denoise3d'))
  stages = {stage.name: stage for stage in program.stages}
  assert stages['g'].top_division() and stages['r1'].top_division()
  assert stages['output'].top_division()
  assert not stages['r0'].top_division()      # u * f * (1.0f / 0.03f)
  assert not stages['diff_u'].top_division()
  _, (num, den) = stages['g'].render(lambda load: 'x', call_prefix='soda_fn_',
                                     split_top=True)
  assert num == '1.0f' and den.startswith('soda_fn_sqrt(')
  chain = core.Stencil.from_text(
      'kernel: q\nburst width: 64\nunroll factor: 1\niterate: 1\n'
      'input float: a(32, *)\n'
      'output float: o(0, 0) = a(0, 0) * a(1, 0) / a(0, 1)\n')
  stage = plan.extract_program(chain).stages[0]
  assert stage.top_division()
  _, (num, den) = stage.render(lambda load: 'a%d%d' % tuple(load.off),
                               split_top=True)
  assert (num, den) == ('(a00 * a10)', 'a01')
  out = io.StringIO()
  codegen.print_kernel(program, codegen.make_schedules(program), out)
  text = out.getvalue()
  # the shipped denoise3d schedule splices r1 into the output statement (a
  # spliced quotient keeps the compiler's division): two batched per step
  steps = text.count('// g: plane')
  assert steps > 0
  assert text.count('soda::div_try(num0, den0, rare)') == 2 * steps
  assert 'if (rare)' in text
  # integer and double statements keep the plain form
  ints = core.Stencil.from_text(
      'kernel: q\nburst width: 64\nunroll factor: 1\niterate: 1\n'
      'input int32: a(32, *)\n'
      'output int32: o(0, 0) = (a(0, 0) + a(1, 0)) / 3\n')
  program = plan.extract_program(ints)
  out = io.StringIO()
  codegen.print_kernel(program, codegen.make_schedules(program), out)
  assert 'div_try' not in out.getvalue()
