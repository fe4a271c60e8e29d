"""The grammar's whitelisted math calls (reference src/soda/grammar.py:25-32)
through the CUDA backend.

* Integer arguments: in the reference's golden loop `sqrt(a(0,0))` on an
  int32 tensor is the C `double sqrt(double)`; the device wrappers must offer
  exactly one candidate for it, in the exact and in the fast build (ADVICE
  r1: the float overloads made such calls ambiguous).
* `ldexp scalbn scalbln ilogb lround llround lrint llrint nexttoward` have
  wrappers (they parsed but had none); `frexp modf remquo nan`
  need pointer / string arguments no SODA expression can supply and raise
  SemanticError instead of an nvcc error.
* GPU: the program below — only exactly specified functions (correctly
  rounded or exact in IEEE-754 / C99; `pow` is not: CUDA's differed from
  glibc's by 1 ulp in 2 of 9933 cells, capture r2a) — equals the CPU oracle
  bit for bit.
"""
import subprocess

import numpy as np
import pytest

import common
import golden
from haoda import util
from soda import core
from soda import cuda as soda_cuda
from soda.codegen import cuda as codegen
from soda.codegen.cuda import plan

TEXT = '''kernel: calls
burst width: 64
unroll factor: 1
iterate: 1
input int32: a(32, *)
input float: f(32, *)
output double: o0(0, 0) = sqrt(a(0, 0)) + ldexp(f(0, 0), a(1, 0)) + scalbn(f(0, 1), 3) + scalbln(f(1, 0), a(0, 1)) + fabs(a(0, 0))
output int32: o1(0, 0) = ilogb(f(0, 0)) + lround(f(1, 0)) + lrint(f(0, 1)) + llround(f(1, 1)) + llrint(f(0, 0)) + ilogb(a(0, 0))
output double: o2(0, 0) = nexttoward(f(0, 0), f(1, 0)) + fmod(f(0, 0), f(0, 1)) + floor(f(0, 0)) + ceil(f(1, 0)) + trunc(f(0, 1)) + round(f(1, 1)) + rint(f(0, 0)) + nearbyint(f(1, 0)) + remainder(f(0, 0), f(1, 1)) + copysign(f(0, 0), a(0, 0)) + nextafter(f(0, 0), f(0, 1)) + fdim(f(0, 0), f(1, 0)) + fmax(f(0, 0), a(1, 0)) + fmin(f(0, 1), f(1, 1)) + fma(f(0, 0), f(1, 0), f(0, 1))
'''


def _nvcc_compiles(tmp_path, kernel, extra):
  path = tmp_path / 'k.cu'
  path.write_text(kernel)
  done = subprocess.run(
      ['nvcc'] + soda_cuda.ARCH_FLAGS + ['-std=c++17'] + extra + [
          '-I', soda_cuda.CSRC_DIR, '-I', soda_cuda.INCLUDE_DIR, '-c',
          str(path), '-o', str(tmp_path / 'k.o')],
      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  assert done.returncode == 0, done.stdout[-3000:]


@pytest.mark.parametrize('extra', [['-fmad=false'], ['-DSODA_CUDA_FAST_MATH']],
                         ids=['exact', 'fast'])
def test_calls_with_integer_arguments_compile(tmp_path, monkeypatch, extra):
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  _, kernel, _ = soda_cuda.generate_sources(core.Stencil.from_text(TEXT))
  for name in ('sqrt', 'ldexp', 'scalbn', 'scalbln', 'ilogb', 'lround',
               'llround', 'lrint', 'llrint', 'nexttoward'):
    assert 'soda_fn_%s(' % name in kernel
  _nvcc_compiles(tmp_path, kernel, extra)


def test_the_oracle_builds_the_same_program(tmp_path):
  orc = golden.Oracle(core.Stencil.from_text(TEXT), build_dir=str(tmp_path))
  outs = orc.run(_inputs((64, 40)))
  assert [o.dtype for o in outs] == [np.float64, np.int32, np.float64]


@pytest.mark.parametrize('call', ['frexp(a(0, 0))', 'modf(a(0, 0))',
                                  'remquo(a(0, 0), a(1, 0))', 'nan(a(0, 0))'])
def test_calls_the_dsl_cannot_feed_are_semantic_errors(call):
  text = ('kernel: bad\nburst width: 64\nunroll factor: 1\niterate: 1\n'
          'input float: a(32, *)\noutput float: o(0, 0) = a(0, 0) + %s\n' %
          call)
  program = plan.extract_program(core.Stencil.from_text(text))
  with pytest.raises(util.SemanticError) as info:
    codegen.check_supported(program)
  assert call.split('(')[0] in str(info.value)


def _inputs(dims, seed=0):
  rng = np.random.default_rng(seed)
  shape = tuple(reversed(dims))
  return [rng.integers(1, 40, size=shape).astype(np.int32),
          (rng.random(shape) * 99.5 + 0.5).astype(np.float32)]


@pytest.mark.gpu
def test_calls_match_the_oracle_bit_for_bit():
  stencil = core.Stencil.from_text(TEXT)
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil)
  for dims, seed in (((256, 50), 1), ((301, 33), 2)):
    inputs = _inputs(dims, seed)
    want = orc.run(inputs)
    got = library.run(inputs)
    for k, (g, w) in enumerate(zip(got, want)):
      common.assert_bit_exact(g, w, 'calls output %d' % k)
