"""IEEE float division without a branch per cell (needs a GPU):
``soda::div_try`` (csrc/soda_cuda_device.cuh) runs the quick sequence of
``div.rn.f32`` unconditionally and raises a flag where its operands leave the
range the sequence is valid in; the emitters then redo the statement as
written.  Wherever the flag stays down the result must be ``__fdiv_rn``'s bit
for bit, and the flag must go up for every operand outside [2^-40, 2^41)
(zeros, denormals, infinities and NaNs included)."""
import ctypes

import pytest

import test_rsqrt_exact as native

pytestmark = pytest.mark.gpu
U64P = native.U64P


def _lib():
  from soda import cuda as soda_cuda
  path = native._built(
      'libdiv_check.so',
      ['nvcc'] + soda_cuda.ARCH_FLAGS +
      ['-O3', '-std=c++17', '-fmad=false', '-shared', '-Xcompiler', '-fPIC',
       '-I', soda_cuda.CSRC_DIR, '-I', soda_cuda.INCLUDE_DIR], 'div_check.cu')
  lib = ctypes.CDLL(path)
  lib.div_check.argtypes = [ctypes.c_float, ctypes.c_uint32, ctypes.c_uint64,
                            ctypes.c_uint64, U64P, U64P, U64P]
  return lib


def _run(lib, a, first, count, random_count):
  bad, rare, missed = (ctypes.c_uint64(), ctypes.c_uint64(),
                       ctypes.c_uint64())
  rc = lib.div_check(a, first, count, random_count, ctypes.byref(bad),
                     ctypes.byref(rare), ctypes.byref(missed))
  assert rc == 0
  return bad.value, rare.value, missed.value


@pytest.mark.parametrize('a', [1.0, 3.0, -0.7, 1.0 / 3.0, 1234.567,
                               16777215.0, 5.9604645e-08])
def test_quick_division_is_exact_where_it_claims_to_be(a):
  lib = _lib()
  # every float of a few binades as divisor and as dividend, and 2^28 random
  # pairs with exponents on both sides of the range test
  for value in (1.0, 2.0 ** -39, 2.0 ** 39, 0.03, 7.5e5):
    bad, rare, missed = _run(lib, a, native._bits(value), 1 << 23, 0)
    assert bad == 0 and missed == 0, (a, value, bad, missed)
    if 2.0 ** -38 < value < 2.0 ** 38:
      assert rare == 0, (a, value, rare)
  bad, rare, missed = _run(lib, a, 0, 0, 1 << 28)
  assert bad == 0 and missed == 0 and 0 < rare < (1 << 28) // 4


def test_special_operands_take_the_exact_path():
  lib = _lib()
  for first, count in ((0, 1 << 16),                           # 0, denormals
                       (native._bits(float('inf')) - 8, 64),   # huge, inf, NaN
                       (native._bits(2.0 ** -41) - 64, 64),
                       (native._bits(2.0 ** 41), 64),
                       (native._bits(-0.0), 1 << 12)):
    bad, rare, missed = _run(lib, 1.5, first, count, 0)
    assert bad == 0 and missed == 0
    assert rare == 2 * count, (first, rare, count)
