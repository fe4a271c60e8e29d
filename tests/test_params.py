"""`param` statements without a GPU: frontend, plan, oracle,
emitted code, C ABI.

The reference accepts `param` in its grammar (src/soda/grammar.py:38) and its
golden-loop emitter spells out what a param is — a C array `T name[s0][s1]..`
read as `name_img[i][j]..`, initialised to p+q.. and passed after the outputs
(src/soda/codegen/xilinx/host.py:1004-1008, 1053-1063, 1095-1097,
header.py:57) — but the reference cannot run such a program: building the
Stencil of any statement that reads a param raises KeyError at
src/soda/core.py:365 (params are not in its symbol table) and `print_test`
reads `param.type`, which ParamStmt does not have (host.py:1029).  So parity
is unpinned; the oracle restates the emitter's text and is checked here
against numpy.
"""
import ctypes
import io

import numpy as np
import pytest

import common
import golden
import param_programs as pp
from haoda import util
from soda import core, cuda as soda_cuda
from soda.codegen import cuda as codegen
from soda.codegen.cuda import plan


def test_frontend_and_plan():
  stencil = pp.stencil_of('conv3')
  assert stencil.param_names == ('w', 'bias')
  assert [tuple(s.size) for s in stencil.param_stmts] == [(3, 3), (1,)]
  program = plan.extract_program(stencil)
  assert program.param_stmts == [('w', 'float', (3, 3)),
                                 ('bias', 'float', (1,))]
  loads = {(l.parent, l.off) for l in program.stages[0].loads}
  # tensor loads are relative to the store point, param loads absolute
  assert ('in', (-1, -1)) in loads and ('w', (2, 1)) in loads
  assert program.param_flat(plan.Load('w', (2, 1))) == 7     # first index slowest
  assert program.window(1) == ((-1, -1), (1, 1))             # params add nothing
  assert plan.pairing_obstacle(program, 2) is not None


@pytest.mark.parametrize('body,message', [
    ('param float: w[2]\noutput float: o(0, 0) = a(0, 0) * w(2)\n', 'outside'),
    ('param float: w[2][2]\noutput float: o(0, 0) = a(0, 0) * w(1)\n',
     'dimension'),
])
def test_bad_param_index_is_a_semantic_error(body, message):
  stencil = core.Stencil.from_text(
      pp.HEADER % ('bad', 1) + 'input float: a(32, *)\n' + body)
  with pytest.raises(util.SemanticError, match=message):
    codegen.make_schedules(plan.extract_program(stencil))


def test_oracle_matches_numpy_restatement():
  orc = golden.Oracle(pp.stencil_of('conv3'))
  assert [p.tolist() for p in orc.reference_params()] == [
      [[0, 1, 2], [1, 2, 3], [2, 3, 4]], [0]]       # p+q (host.py:1053-1063)
  rng = np.random.default_rng(3)
  x = rng.random((30, 40), dtype=np.float32)
  w = rng.random((3, 3), dtype=np.float32)
  b = np.array([0.5], np.float32)
  out, = orc.run([x], params=[w, b])
  acc = None
  for j in range(3):            # the statement's own order of additions
    for i in range(3):
      term = x[j:28 + j, i:38 + i] * w[j, i]
      acc = term if acc is None else acc + term
  want = np.zeros_like(x)
  want[1:-1, 1:-1] = acc + b[0]
  common.assert_bit_exact(out, want)
  with pytest.raises(ValueError):
    orc.run([x])


def test_oracle_iterated_with_params():
  orc = golden.Oracle(pp.stencil_of('relax'))
  rng = np.random.default_rng(4)
  x = rng.random((25, 31), dtype=np.float32)
  c = np.array([0.6, 0.1], np.float32)
  out, = orc.run([x], params=[c])
  cur = x
  for _ in range(6):
    nxt = np.zeros_like(cur)
    nxt[1:-1, 1:-1] = cur[1:-1, 1:-1] * c[0] + (
        cur[1:-1, 2:] + cur[1:-1, :-2] + cur[2:, 1:-1] + cur[:-2, 1:-1]) * c[1]
    cur = nxt
  common.assert_bit_exact(out[6:-6, 6:-6], cur[6:-6, 6:-6])
  assert not out[:6].any() and not out[:, :6].any()


def test_emitted_code():
  program = plan.extract_program(pp.stencil_of('conv3'))
  kernel, host = io.StringIO(), io.StringIO()
  codegen.print_kernel(program, codegen.make_schedules(program), kernel)
  from soda.codegen.cuda import host as host_gen
  host_gen.print_code(program, host)
  assert 'static_cast<const float*>(a.param_ptr[0])' in kernel.getvalue()
  assert '__ldg(prm_w + 7)' in kernel.getvalue()          # w(2, 1)
  assert '__ldg(prm_bias + 0)' in kernel.getvalue()
  # the reference entry: inputs, outputs, params, xclbin (header.py:57-60)
  assert ('int conv3(buffer_t* var_in_buffer, buffer_t* var_out_buffer, '
          'buffer_t* var_w_buffer, buffer_t* var_bias_buffer, '
          'const char* xclbin)') in host.getvalue()
  header = io.StringIO()
  codegen.print_header(program, header)
  assert 'buffer_t *var_w_buffer, buffer_t *var_bias_buffer' in \
      header.getvalue()
  for style in ('ring',):
    ring = io.StringIO()
    codegen.print_kernel(program, codegen.make_schedules(
        program, codegen.Options(style=style)), ring)
    assert '__ldg(prm_w + 7)' in ring.getvalue()


def test_library_reports_params_and_checks_them():
  library = soda_cuda.compile_stencil(pp.stencil_of('scale16'))
  assert library.params == [('k', 'int16', (2, 1, 3))]
  x = np.zeros((8, 16, 32), np.int16)
  with pytest.raises(ValueError, match='param array'):
    library.run([x])
  with pytest.raises(ValueError, match='shape'):
    library.run([x], params=[np.zeros((2, 3), np.int16)])
  import torch
  if not torch.cuda.is_available():
    with pytest.raises(soda_cuda.CudaError) as info:      # no CPU fallback
      library.run([x], params={'k': np.ones((2, 1, 3), np.int16)})
    assert info.value.code == -19


def test_entry_point_takes_the_params_after_the_outputs():
  """`int conv3(buffer_t* in, buffer_t* out, buffer_t* w, buffer_t* bias,
  const char* xclbin)` with C++ linkage, as reference header.py:57-60 would
  declare it (tensors = inputs + outputs + params)."""
  import subprocess
  library = soda_cuda.compile_stencil(pp.stencil_of('conv3'))
  symbols = subprocess.run(['nm', '-D', '--defined-only', library.path],
                           stdout=subprocess.PIPE, text=True,
                           check=True).stdout
  assert ' T _Z5conv3P8buffer_tS0_S0_S0_PKc' in symbols
  for name in ('soda_cuda_run_params', 'soda_cuda_set_params',
               'soda_cuda_num_params', 'soda_cuda_param_size'):
    assert ' T %s\n' % name in symbols
