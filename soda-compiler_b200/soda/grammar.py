"""SODA statement nodes and the parse-tree -> IR builder.

The reference lets textX instantiate its node classes from a grammar string
(reference: src/soda/grammar.py:8-43, :45-185).  Here ``soda.dsl_parser`` parses
the same language into a neutral tree and ``build_program`` turns that tree
into this package's flattened IR: single-operand precedence levels,
parenthesised operands and identity unary chains never materialise, which is
the shape the reference reaches after ``arithmetic.simplify``
(src/soda/core.py:245-248).
"""
from haoda import ir
from haoda import util
from soda import dsl_parser
from soda.dsl_parser import FUNC_NAMES, SodaSyntaxError, parse_tree

assert dsl_parser.BINARY_LEVELS == ir.BINARY_LEVELS

# --- statement nodes --------------------------------------------------------

class InputStmt(ir.Node):
  """``input [dram a.b] type: name[(t0, t1, *)]``.

  ``tile_size`` always ends with 0 for the streamed last dimension
  (reference: src/soda/grammar.py:45-66).  Tile sizes are FPGA line-buffer
  parameters; the CUDA backend only takes the dimensionality from them.
  """
  FIELDS = ('haoda_type', 'name')
  LISTS = ('tile_size', 'dram')

  def __init__(self, **kwargs):
    super().__init__(**kwargs)
    self.dram = self.dram or (0,)
    self.tile_size = self.tile_size + (0,)

  def __str__(self):
    text = 'input %s: %s' % (self.haoda_type, self.name)
    if self.tile_size[:-1]:
      text += '(%s, *)' % ', '.join(map(str, self.tile_size[:-1]))
    return text


class LocalStmtOrOutputStmt(ir.Node):
  """``local|output type: [lets] name(store idx) = expr`` (reference :68-99)."""
  FIELDS = ('haoda_type', 'ref', 'expr')
  LISTS = ('let',)
  KEYWORD = None

  def __init__(self, **kwargs):
    super().__init__(**kwargs)
    # ``let`` variables used in later lets / the expression take the let's type.
    known = {}

    def typed(obj, _):
      if isinstance(obj, ir.Var) and obj.name in known:
        obj.haoda_type = known[obj.name]
      return obj
    lets = []
    for let in self.let:
      let = let.visit(typed)
      lets.append(let)
      try:
        known[let.name] = let.haoda_type
      except AttributeError:    # untyped let over not-yet-typed refs
        known[let.name] = None
    self.let = tuple(lets)
    self.expr = self.expr.visit(typed)

  @property
  def name(self):
    return self.ref.name

  def __str__(self):
    lets = ''.join('\n  %s' % let for let in self.let)
    return '%s %s:%s %s = %s' % (self.KEYWORD, self.haoda_type,
                                 lets + '\n ' if lets else '', self.ref,
                                 ir.unparenthesize(self.expr))


class LocalStmt(LocalStmtOrOutputStmt):
  KEYWORD = 'local'


class OutputStmt(LocalStmtOrOutputStmt):
  KEYWORD = 'output'
  LISTS = LocalStmtOrOutputStmt.LISTS + ('dram',)

  def __init__(self, **kwargs):
    super().__init__(**kwargs)
    self.dram = self.dram or (0,)


class ParamStmt(ir.Node):
  """``param type[, attrs]: name[size]...`` — a small constant array."""
  FIELDS = ('haoda_type', 'name')
  LISTS = ('attr', 'size', 'dram')

  def __str__(self):
    return 'param %s%s: %s%s' % (
        self.haoda_type, ''.join(', %s' % a for a in self.attr), self.name,
        ''.join('[%d]' % s for s in self.size))


class ParamAttr(ir.Node):
  """FPGA array-partitioning hints on a param; carried, not used on GPU."""
  FIELDS = ('dup', 'strategy', 'factor', 'dim')

  def __str__(self):
    if self.dup is not None:
      return 'dup %s' % self.dup
    text = 'partition %s' % self.strategy
    if self.strategy == 'cyclic':
      text += ' factor=%s' % self.factor
    if self.dim is not None:
      text += ' dim=%s' % self.dim
    return text


class SodaProgram(ir.Node):
  """A parsed ``.soda`` file (reference: src/soda/grammar.py:126-160)."""
  FIELDS = ('burst_width', 'iterate', 'app_name', 'unroll_factor')
  LISTS = ('input_stmts', 'param_stmts', 'local_stmts', 'output_stmts')

  def __init__(self, **kwargs):
    super().__init__(**kwargs)
    # All inputs that spell out a tile must agree; bare inputs inherit it.
    self.tile_size = None
    for stmt in self.input_stmts:
      if not stmt.tile_size[:-1]:
        continue
      if self.tile_size is None:
        self.tile_size = stmt.tile_size
      elif self.tile_size != stmt.tile_size:
        raise util.SemanticError(
            "tile size %s doesn't match previous one %s" %
            (stmt.tile_size, self.tile_size))
    if self.tile_size is None:    # 1-D program
      self.tile_size = self.input_stmts[-1].tile_size
    self.dim = len(self.tile_size)

  def __str__(self):
    lines = ['burst width: %d' % self.burst_width,
             'iterate: %d' % self.iterate,
             'kernel: %s' % self.app_name,
             'unroll factor: %d' % self.unroll_factor]
    for group in (self.input_stmts, self.param_stmts, self.local_stmts,
                  self.output_stmts):
      lines.extend(map(str, group))
    return '\n'.join(lines)


# --- neutral tree -> flattened IR --------------------------------------------

_BINARY_CLASSES = {name: getattr(ir, name) for name, _ in ir.BINARY_LEVELS}


def _build(tree):
  if not isinstance(tree, tuple):
    return tree
  rule, attrs, _ = tree
  if rule in _BINARY_CLASSES:
    operands = [_build(o) for o in attrs['operand']]
    if len(operands) == 1:
      return operands[0]
    return _BINARY_CLASSES[rule](operand=operands, operator=attrs['operator'])
  if rule == 'Unary':
    operand = _build(attrs['operand'])
    if ir.is_identity_unary(attrs['operator']):
      return operand
    return ir.Unary(operator=attrs['operator'], operand=operand)
  if rule == 'Operand':
    if attrs['num'] is not None:
      return ir.Num(text=attrs['num'])
    for key in ('cast', 'call', 'ref', 'var', 'expr'):
      if attrs[key] is not None:
        return _build(attrs[key])
  if rule == 'Cast':
    return ir.Cast(haoda_type=attrs['haoda_type'], expr=_build(attrs['expr']))
  if rule == 'Call':
    return ir.Call(name=attrs['name'], arg=[_build(a) for a in attrs['arg']])
  if rule == 'Var':
    return ir.Var(name=attrs['name'], idx=attrs['idx'])
  if rule == 'Ref':
    return ir.Ref(name=attrs['name'], idx=attrs['idx'], lat=attrs['lat'])
  if rule == 'Let':
    return ir.Let(declared_type=attrs['haoda_type'], name=attrs['name'],
                  expr=_build(attrs['expr']))
  if rule == 'InputStmt':
    return InputStmt(**attrs)
  if rule in ('LocalStmt', 'OutputStmt'):
    built = {key: ([_build(v) for v in val] if key == 'let' else _build(val))
             for key, val in attrs.items()}
    return (LocalStmt if rule == 'LocalStmt' else OutputStmt)(**built)
  if rule == 'ParamStmt':
    built = dict(attrs, attr=[_build(a) for a in attrs['attr']])
    return ParamStmt(**built)
  if rule == 'ParamAttr':
    part = attrs['partitioning']
    detail = part[1] if part is not None else dict(
        strategy=None, factor=None, dim=None)
    return ParamAttr(dup=attrs['dup'], **detail)
  if rule == 'SodaProgram':
    built = {key: ([_build(v) for v in val] if isinstance(val, list) else val)
             for key, val in attrs.items()}
    return SodaProgram(**built)
  raise util.InternalError('unknown grammar rule %s' % rule)


def build_program(tree):
  return _build(tree)


def parse(text):
  """SODA source text -> SodaProgram."""
  return build_program(parse_tree(text))


class _MetaModel:
  """``metamodel().model_from_str(text)``: the call shape sodac used with textX."""

  @staticmethod
  def model_from_str(text):
    return parse(text)

  @staticmethod
  def model_from_file(path):
    with open(path) as handle:
      return parse(handle.read())


def metamodel():
  return _MetaModel()
