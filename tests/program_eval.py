"""Evaluate a plan.Program on the CPU with periodic boundaries (test helper).

Not the oracle (that is oracle/golden.py, the reference's golden loop): this
evaluates the PLANNER's view of a program — its stages in order, each Ref
resolved through ``Stage.render`` — so that two Programs that should compute
the same function (a program and its version with single-use locals spliced
into their readers) can be compared bit for bit on every cell, without a GPU.
Boundaries wrap around, so every cell of every tensor is defined.
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

import golden
from haoda import util

_BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..',
                      'oracle', '_build')


def source(program, dims):
  dim = program.dim
  cells = int(np.prod(dims))
  out = ['#include <cmath>', '#include <cstdint>', '#include <cstdlib>', '']
  if 'half' in [util.get_c_type(t) for t in program.types.values()]:
    out += ['struct half { _Float16 v; half() = default; template <typename T>'
            ' half(T x) : v(static_cast<_Float16>(x)) {} operator float() '
            'const { return static_cast<float>(v); } };']
  out += ['static inline long wrap(long x, long n) { x %= n; return x < 0 ? '
          'x + n : x; }', '']
  names = [n for n, _ in program.inputs]
  args = ', '.join(['const void* const* in', 'void* const* outp'])
  out.append('extern "C" void eval_program(%s) {' % args)
  for k, (name, haoda_type) in enumerate(program.inputs):
    c = util.get_c_type(haoda_type)
    out.append('  const %s* %s = static_cast<const %s*>(in[%d]);' % (
        c, name, c, k))
  for stage in program.stages:
    c = stage.c_type
    if stage.is_output:
      out.append('  %s* %s = static_cast<%s*>(outp[%d]);' % (
          c, stage.name, c, program.output_names.index(stage.name)))
    else:
      out.append('  %s* %s = static_cast<%s*>(malloc(%d * sizeof(%s)));' % (
          c, stage.name, c, cells, c))
    coords = ['x%d' % d for d in range(dim)]
    for d in reversed(range(dim)):
      out.append('  for (long x%d = 0; x%d < %d; ++x%d)' % (d, d, dims[d], d))
    out.append('  {')

    def ref_code(load):
      index, pitch = [], 1
      for d in range(dim):
        index.append('wrap(x%d + (%d), %d) * %d' % (d, load.off[d], dims[d],
                                                    pitch))
        pitch *= dims[d]
      return '%s[%s]' % (load.parent, ' + '.join(index))
    lets, expr = stage.render(
        ref_code, cast=lambda c_type, text: 'static_cast<%s>(%s)' % (c_type,
                                                                     text))
    for let in lets:
      out.append('    ' + let)
    here, pitch = [], 1
    for d in range(dim):
      here.append('x%d * %d' % (d, pitch))
      pitch *= dims[d]
    out.append('    %s[%s] = %s;' % (stage.name, ' + '.join(here), expr))
    out.append('  }')
    del coords
  for stage in program.stages:
    if not stage.is_output:
      out.append('  free(%s);' % stage.name)
  out.append('}')
  del names
  return '\n'.join(out) + '\n'


def evaluate(program, inputs):
  """Outputs of ONE iteration of ``program`` on ``inputs`` (numpy arrays of
  shape reversed(dims)), periodic boundaries."""
  assert not program.params
  dims = tuple(reversed(inputs[0].shape))
  text = source(program, dims)
  digest = hashlib.sha256(text.encode()).hexdigest()[:12]
  os.makedirs(_BUILD, exist_ok=True)
  lib = os.path.join(_BUILD, 'eval_%s_%s.so' % (program.app_name, digest))
  if not os.path.exists(lib):
    src = lib[:-3] + '.cpp'
    with open(src, 'w') as handle:
      handle.write(text)
    subprocess.run(['g++', '-O1', '-std=c++11', '-shared', '-fPIC', src, '-o',
                    lib + '.tmp'], check=True)
    os.replace(lib + '.tmp', lib)
  fn = ctypes.CDLL(lib).eval_program
  fn.restype = None
  outputs = [np.empty(inputs[0].shape, dtype=golden.NUMPY_TYPES[t])
             for _, t in program.outputs]
  pointers = lambda arrays: (ctypes.c_void_p * len(arrays))(
      *[a.ctypes.data for a in arrays])
  inputs = [np.ascontiguousarray(a) for a in inputs]
  fn(pointers(inputs), pointers(outputs))
  return outputs
