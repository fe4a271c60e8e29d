"""`a / sqrt(x)` on float operands: the float-arithmetic decision of
csrc/soda_cuda_device.cuh (soda::RecipSqrtF32) against the FP64 evaluation the
reference's golden loop performs (`sqrt(float)` binds to the C
`double sqrt(double)`, the quotient is a double, the local stores RN32 of it).

CPU: a C model of the same arithmetic (tests/native/rsqrt_model.c) with the
hardware approximation replaced by a skewed correctly rounded value, over every
float of two binades and samples of the rest.
GPU: the device code itself over EVERY positive float.
"""
import ctypes
import os
import struct
import subprocess

import pytest

import common

NATIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'native')
BUILD = os.path.join(NATIVE, '_build')
U64P = ctypes.POINTER(ctypes.c_uint64)


def _bits(value):
  return struct.unpack('<I', struct.pack('<f', value))[0]


def _built(name, command, source):
  os.makedirs(BUILD, exist_ok=True)
  out = os.path.join(BUILD, name)
  src = os.path.join(NATIVE, source)
  header = os.path.join(common.ROOT, 'soda-compiler_b200', 'csrc',
                        'soda_cuda_device.cuh')
  newest = max(os.path.getmtime(src), os.path.getmtime(header))
  if not os.path.exists(out) or os.path.getmtime(out) < newest:
    tmp = '%s.%d.tmp' % (out, os.getpid())
    subprocess.run(command + ['-o', tmp, src], check=True)
    os.replace(tmp, out)
  return out


def build_model():
  return _built('librsqrt_model.so',
                ['gcc', '-O2', '-fopenmp', '-ffp-contract=off', '-shared',
                 '-fPIC', '-lm'], 'rsqrt_model.c')


def build_check():
  from soda import cuda as soda_cuda
  return _built('librsqrt_check.so',
                ['nvcc'] + soda_cuda.ARCH_FLAGS +
                ['-O3', '-std=c++17', '-fmad=false', '-shared', '-Xcompiler',
                 '-fPIC', '-I', soda_cuda.CSRC_DIR, '-I',
                 soda_cuda.INCLUDE_DIR], 'rsqrt_check.cu')


def _model():
  lib = ctypes.CDLL(build_model())
  lib.rsqrt_model.argtypes = [ctypes.c_float, ctypes.c_uint32,
                              ctypes.c_uint64, ctypes.c_int, U64P, U64P]
  return lib


@pytest.mark.parametrize('a', [1.0, 3.7, -0.3, 1e-3, 65504.0])
def test_model_never_decides_wrongly(a):
  lib = _model()
  # MUFU.RSQ is within 2^-22.4 of 1/sqrt(x): up to ~3 ulps off the rounded value
  for skew in (0, 2, -2, 4, -4):
    for first, count in ((_bits(1.0), _bits(4.0) - _bits(1.0)),
                         (_bits(3e-5), 1 << 21), (_bits(7e8), 1 << 21),
                         (_bits(1e-29), 1 << 20), (_bits(9e28), 1 << 20)):
      bad, undecided = ctypes.c_uint64(), ctypes.c_uint64()
      lib.rsqrt_model(a, first, count, skew, ctypes.byref(bad),
                      ctypes.byref(undecided))
      assert bad.value == 0, (a, skew, first)
      # the FP64 fallback must stay rare, or the point is lost
      assert undecided.value < count // 10000 + 64, (a, skew, first)


def test_model_sends_specials_to_the_exact_path():
  lib = _model()
  for first, count in ((0, 4096),                      # 0 and denormals
                       (_bits(float('inf')), 16),      # inf, NaNs
                       (_bits(-1.0), 16),              # negative
                       (_bits(1e31), 16), (_bits(1e-31), 16)):
    bad, undecided = ctypes.c_uint64(), ctypes.c_uint64()
    lib.rsqrt_model(1.0, first, count, 0, ctypes.byref(bad),
                    ctypes.byref(undecided))
    assert bad.value == 0 and undecided.value == count


@pytest.mark.gpu
@pytest.mark.parametrize('a,every', [(1.0, True), (0.37, False),
                                     (-1234.5, False)])
def test_device_decision_over_all_floats(a, every):
  lib = ctypes.CDLL(build_check())
  lib.rsqrt_check.argtypes = [ctypes.c_float, ctypes.c_uint32,
                              ctypes.c_uint64, U64P, U64P]
  if every:       # every positive float, denormals, inf and NaNs included
    spans = [(0, 1 << 31)]
  else:           # 2^26 consecutive floats in each of a few ranges
    spans = [(_bits(v), 1 << 26) for v in (1e-20, 5e-5, 1.0, 300.0, 1e20)]
  for first, count in spans:
    bad, undecided = ctypes.c_uint64(), ctypes.c_uint64()
    rc = lib.rsqrt_check(a, first, count, ctypes.byref(bad),
                         ctypes.byref(undecided))
    assert rc == 0
    assert bad.value == 0, (a, first, bad.value)
    if not every:
      assert undecided.value < count // 10000
    print('a=%g first=%#x: %d of %d undecided' % (a, first, undecided.value,
                                                  count))
