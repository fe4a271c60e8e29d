"""Build and drive the reference's UNMODIFIED ``<app>_test`` harness — TEST
INFRASTRUCTURE ONLY (see oracle/golden.py for who may import this).

``build_ref(soda_file, iterate)`` runs the reference frontend from
/root/reference (through oracle/ref_tool.py) to emit ``<app>.h`` and
``<app>_test.cpp`` — header.print_code and host.print_test verbatim (reference
src/soda/codegen/xilinx/header.py:7-64, host.py:984-1167) — into oracle/_ref/,
adds a trampoline that routes the harness' call to the device entry
``<app>(buffer_t*..., xclbin)`` (host.py:1068-1070) to a function pointer, and
compiles everything with the pinned oracle flags.  oracle/_ref/ is git-ignored
(generated from reference code) but travels to GPU machines with the repo
snapshot, where /root/reference does not exist.

``RefHarness(lib).test(dims, implementation)`` then calls the reference's
``<app>_test("", dims)``: the reference allocates and initialises the inputs,
calls ``implementation(inputs) -> outputs`` where the FPGA would run, recomputes
every stage on the CPU and returns its mismatch count (integers exact, floats
relative error <= 1e-5 or $THRESHOLD, host.py:1118-1146).
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, '_ref')
CXX_FLAGS = ['-O3', '-fopenmp', '-std=c++11']

sys.path.insert(0, _HERE)
import golden    # noqa: E402  pylint: disable=wrong-import-position


def lib_path(app_name, iterate):
  return os.path.join(REF_DIR, 'lib%s_it%d_ref.so' % (app_name, iterate))


def build_ref(soda_file, iterate=None, force=False):
  """Returns the path of the compiled reference harness (needs /root/reference)."""
  stencil = golden.stencil_from_file(soda_file, iterate)
  app, iterate = stencil.app_name, stencil.iterate
  lib = lib_path(app, iterate)
  if os.path.exists(lib) and not force:
    return lib
  work = os.path.join(REF_DIR, '%s_it%d' % (app, iterate))
  os.makedirs(work, exist_ok=True)
  subprocess.run(
      [sys.executable, os.path.join(_HERE, 'ref_tool.py'), 'harness',
       soda_file, '--iterate', str(iterate), '--outdir', work], check=True)
  names = list(stencil.input_names) + list(stencil.output_names)
  with open(os.path.join(work, '%s_hook.cpp' % app), 'w') as out:
    out.write('// Trampoline for the reference harness (ours, not reference '
              'code).\n#include "%s.h"\n\n' % app)
    out.write('typedef int (*soda_ref_hook_t)(buffer_t** buffers, int count, '
              'const char* xclbin);\nstatic soda_ref_hook_t g_hook = 0;\n')
    out.write('extern "C" void soda_ref_set_hook(soda_ref_hook_t hook) '
              '{ g_hook = hook; }\n\n')
    out.write('int %s(%sconst char* xclbin)\n{\n' % (
        app, ''.join('buffer_t *var_%s_buffer, ' % n for n in names)))
    out.write('  buffer_t* buffers[] = {%s};\n' % ', '.join(
        'var_%s_buffer' % n for n in names))
    out.write('  return g_hook ? g_hook(buffers, %d, xclbin) : -1;\n}\n\n' %
              len(names))
    out.write('int %s_test(const char* xclbin, const int dims[4]);\n' % app)
    out.write('extern "C" int soda_ref_test(const int* dims) '
              '{ return %s_test("", dims); }\n' % app)
  command = ['g++'] + CXX_FLAGS + [
      '-fPIC', '-shared', '-I', work,
      os.path.join(work, '%s_test.cpp' % app),
      os.path.join(work, '%s_hook.cpp' % app), '-o', lib]
  done = subprocess.run(command, stdout=subprocess.PIPE,
                        stderr=subprocess.STDOUT, text=True, check=False)
  if done.returncode != 0:
    raise RuntimeError('reference harness build failed:\n%s\n%s' % (
        ' '.join(command), done.stdout))
  return lib


class _BufferT(ctypes.Structure):
  _fields_ = [('dev', ctypes.c_uint64), ('host', ctypes.c_void_p),
              ('extent', ctypes.c_int32 * 4), ('stride', ctypes.c_int32 * 4),
              ('min', ctypes.c_int32 * 4), ('elem_size', ctypes.c_int32),
              ('host_dirty', ctypes.c_bool), ('dev_dirty', ctypes.c_bool),
              ('_padding', ctypes.c_uint8 * 2)]


_HOOK_T = ctypes.CFUNCTYPE(ctypes.c_int,
                           ctypes.POINTER(ctypes.POINTER(_BufferT)),
                           ctypes.c_int, ctypes.c_char_p)


class RefHarness:
  """The compiled reference harness of one (program, iterate)."""

  def __init__(self, lib, stencil):
    self._lib = ctypes.CDLL(lib)
    self._lib.soda_ref_test.restype = ctypes.c_int
    self._lib.soda_ref_test.argtypes = [ctypes.POINTER(ctypes.c_int)]
    self._lib.soda_ref_set_hook.argtypes = [_HOOK_T]
    self.stencil = stencil
    self.in_types = [golden.NUMPY_TYPES[t] for t in stencil.input_types]
    self.out_types = [golden.NUMPY_TYPES[t] for t in stencil.output_types]

  def test(self, dims, implementation):
    """Mismatch count of ``implementation`` as judged by the reference."""
    shape = tuple(reversed(dims))
    n_in = len(self.in_types)
    failure = []

    def hook(buffers, count, _xclbin):
      try:
        views = []
        for k in range(count):
          buf = buffers[k].contents
          dtype = (self.in_types + self.out_types)[k]
          assert buf.elem_size == np.dtype(dtype).itemsize
          assert tuple(buf.extent[:len(dims)]) == tuple(dims)
          raw = (ctypes.c_char * (int(np.prod(shape)) * buf.elem_size)
                 ).from_address(buf.host)
          views.append(np.frombuffer(raw, dtype=dtype).reshape(shape))
        outputs = implementation([v for v in views[:n_in]])
        for view, out in zip(views[n_in:], outputs):
          view[...] = out
        return 0
      except BaseException as e:   # pylint: disable=broad-except
        failure.append(e)
        return -1
    callback = _HOOK_T(hook)
    self._lib.soda_ref_set_hook(callback)
    errors = self._lib.soda_ref_test(
        (ctypes.c_int * 4)(*(list(dims) + [1] * (4 - len(dims)))))
    if failure:
      raise failure[0]
    return errors
