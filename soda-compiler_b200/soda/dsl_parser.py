"""Recursive-descent PEG parser for the SODA DSL (standard library only).

The language is the one the reference defines as a textX grammar
(reference: src/soda/grammar.py:8-43 for programs and statements,
src/haoda/ir/__init__.py:13-64 for types, literals and expressions).
``parse_tree`` returns a neutral ``(rule, attrs, pos)`` tree whose rule and
attribute names are exactly those of that grammar; ``soda.grammar`` turns it
into this package's IR.  Keeping this module free of package imports lets the
test tooling under ``oracle/`` reuse it to feed the unmodified reference
classes in a separate process, where ``soda``/``haoda`` name the reference.

Known, deliberate difference: math-function names are matched longest-first,
so ``cosh``/``exp2``/``log10``/``atan2``/``erfc`` parse.  The reference's ordered
choice commits to the shorter prefix (``cos``, ``exp``, ...) and then fails.
"""
import re

# C precedence levels of binary operators, loosest first:
# (grammar rule, operators tried in this order)
BINARY_LEVELS = (
    ('Expr', ('||',)),
    ('LogicAnd', ('&&',)),
    ('BinaryOr', ('|',)),
    ('Xor', ('^',)),
    ('BinaryAnd', ('&',)),
    ('EqCmp', ('==', '!=')),
    ('LtCmp', ('<=', '>=', '<', '>')),
    ('AddSub', ('+', '-')),
    ('MulDiv', ('*', '/', '%')),
)

FUNC_NAMES = (
    'cos sin tan acos asin atan atan2 cosh sinh tanh acosh asinh atanh '
    'exp frexp ldexp log log10 modf exp2 expm1 ilogb log1p log2 logb scalbn '
    'scalbln pow sqrt cbrt hypot erf erfc tgamma lgamma ceil floor fmod trunc '
    'round lround llround rint lrint llrint nearbyint remainder remquo '
    'copysign nan nextafter nexttoward fdim fmax fmin fabs abs fma min max '
    'select').split()


class SodaSyntaxError(Exception):
  """The text is not a SODA program (sodac exits 1, like TextXSyntaxError)."""

  def __init__(self, message, line=None, col=None):
    where = '' if line is None else ' at line %d, column %d' % (line, col)
    super().__init__(message + where)
    self.line, self.col = line, col


class _Backtrack(Exception):
  def __init__(self, pos, expected):
    super().__init__(expected)
    self.pos, self.expected = pos, expected


_INT_SUFFIX = r'(?:[Uu][Ll][Ll]?|[Ll]?[Ll]?[Uu]?)'
_RE = {
    'skip': re.compile(r'(?:\s+|#[^\n]*)*'),
    'ID': re.compile(r'[^\d\W]\w*\b'),
    'INT': re.compile(r'[-+]?[0-9]+\b'),
    'Type': re.compile(r'u?int[1-9]\d*(?:_[1-9]\d*)?'
                       r'|float[1-9]\d*(?:_[1-9]\d*)?|float|double|half'),
    'Float': re.compile(r'(?:(?:\d*\.\d+|\d+\.)(?:[+-]?[Ee]\d+)?'
                        r'|\d+[+-]?[Ee]\d+)[FfLl]?'),
    'UInt': re.compile(r'0[Xx][0-9a-fA-F]+%s|0[Bb][01]+%s|0[0-7]+%s|\d+%s' %
                       ((_INT_SUFFIX,) * 4)),
    'Func': re.compile('(?:%s)(?=\\s*\\()' % '|'.join(
        sorted(FUNC_NAMES, key=len, reverse=True))),
}


class _Parser:
  """PEG-style recursive descent with explicit backtracking points."""

  def __init__(self, text):
    self.text = text
    self.pos = 0
    self.furthest = (0, 'a SODA program')

  # --- token level -------------------------------------------------------
  def _skip(self):
    self.pos = _RE['skip'].match(self.text, self.pos).end()

  def _fail(self, expected):
    if self.pos >= self.furthest[0]:
      self.furthest = (self.pos, expected)
    raise _Backtrack(self.pos, expected)

  def lit(self, token):
    self._skip()
    if not self.text.startswith(token, self.pos):
      self._fail(repr(token))
    self.pos += len(token)
    return token

  def rx(self, kind):
    self._skip()
    m = _RE[kind].match(self.text, self.pos)
    if m is None:
      self._fail(kind)
    self.pos = m.end()
    return m.group()

  def attempt(self, rule, *args):
    """Run ``rule``; on failure restore the position and return None."""
    start = self.pos
    try:
      return rule(*args)
    except _Backtrack:
      self.pos = start
      return None

  def peek(self, token):
    self._skip()
    return self.text.startswith(token, self.pos)

  def node(self, rule, start, **attrs):
    return (rule, attrs, start)

  # --- program level -----------------------------------------------------
  def program(self):
    attrs = dict(input_stmts=[], param_stmts=[], local_stmts=[],
                 output_stmts=[])
    seen = set()

    def scalar(key, *words, value='INT'):
      for word in words:
        self.lit(word)
      self.lit(':')
      got = self.rx(value)
      return key, (int(got) if value == 'INT' else got)

    def run(key, rule):
      stmts = [rule()]
      while True:
        more = self.attempt(rule)
        if more is None:
          return key, stmts
        stmts.append(more)

    items = (
        lambda: scalar('burst_width', 'burst', 'width'),
        lambda: scalar('iterate', 'iterate'),
        lambda: scalar('app_name', 'kernel', value='ID'),
        lambda: scalar('unroll_factor', 'unroll', 'factor'),
        lambda: run('input_stmts', self.input_stmt),
        lambda: run('param_stmts', self.param_stmt),
        lambda: run('local_stmts', self.local_stmt),
        lambda: run('output_stmts', self.output_stmt),
    )
    start = self.pos
    while True:
      self._skip()
      if self.pos == len(self.text):
        break
      for item in items:
        got = self.attempt(item)
        if got is not None:
          break
      else:
        self._fail('a header item or statement')
      key, value = got
      if key in seen:   # every group element matches once, as one run
        self._fail('each of the %s only once' % key.replace('_', ' '))
      seen.add(key)
      attrs[key] = value
    for key in ('burst_width', 'iterate', 'app_name', 'unroll_factor',
                'input_stmts', 'output_stmts'):
      if key not in seen:
        self._fail(key.replace('_', ' ').replace(' stmts', ' statement'))
    return self.node('SodaProgram', start, **attrs)

  def dram(self):
    banks = []
    if self.attempt(self.lit, 'dram') is not None:
      banks.append(int(self.rx('INT')))
      while self.attempt(self.lit, '.') is not None:
        banks.append(int(self.rx('INT')))
    return banks

  def input_stmt(self):
    self._skip()
    start = self.pos
    self.lit('input')
    dram = self.dram()
    haoda_type = self.rx('Type')
    self.lit(':')
    name = self.rx('ID')
    tile_size = []

    def tile():
      self.lit('(')
      while True:
        size = self.attempt(lambda: (int(self.rx('INT')), self.lit(','))[0])
        if size is None:
          break
        tile_size.append(size)
      self.lit('*')
      self.lit(')')
      return True
    if self.attempt(tile) is None:
      del tile_size[:]
    return self.node('InputStmt', start, dram=dram, haoda_type=haoda_type,
                     name=name, tile_size=tile_size)

  def _compute_stmt(self, keyword, rule, with_dram):
    self._skip()
    start = self.pos
    self.lit(keyword)
    attrs = {}
    if with_dram:
      attrs['dram'] = self.dram()
    attrs['haoda_type'] = self.rx('Type')
    self.lit(':')
    lets = []
    while True:
      let = self.attempt(self.let)
      if let is None:
        break
      lets.append(let)
    ref = self.ref()
    self.lit('=')
    return self.node(rule, start, let=lets, ref=ref, expr=self.expr(),
                     **attrs)

  def local_stmt(self):
    return self._compute_stmt('local', 'LocalStmt', False)

  def output_stmt(self):
    return self._compute_stmt('output', 'OutputStmt', True)

  def param_stmt(self):
    self._skip()
    start = self.pos
    self.lit('param')
    dram = self.dram()
    haoda_type = self.rx('Type')
    attr = []
    while self.attempt(self.lit, ',') is not None:
      attr.append(self.param_attr())
    self.lit(':')
    name = self.rx('ID')
    size = []
    while self.attempt(self.lit, '[') is not None:
      size.append(int(self.rx('INT')))
      self.lit(']')
    return self.node('ParamStmt', start, dram=dram, haoda_type=haoda_type,
                     attr=attr, name=name, size=size)

  def param_attr(self):
    self._skip()
    start = self.pos
    if self.attempt(self.lit, 'dup') is not None:
      return self.node('ParamAttr', start, dup=self.int_(), partitioning=None)
    self.lit('partition')
    part = dict(strategy=None, dim=None, factor=None)
    if self.attempt(self.lit, 'complete') is not None:
      part['strategy'] = 'complete'
    else:
      self.lit('cyclic')
      part['strategy'] = 'cyclic'
      self.lit('factor')
      self.lit('=')
      part['factor'] = self.int_()

    def dim():
      self.lit('dim')
      self.lit('=')
      return self.int_()
    part['dim'] = self.attempt(dim)
    return self.node('ParamAttr', start, dup=None,
                     partitioning=self.node('Partitioning', start, **part))

  # --- expression level --------------------------------------------------
  def int_(self):
    """Rule ``Int``: optional sign, then hex/bin/oct/dec with suffixes."""
    self._skip()
    sign = ''
    for candidate in '+-':
      if self.text.startswith(candidate, self.pos):
        sign = candidate
        self.pos += 1
        break
    return sign + self.rx('UInt')

  def num(self):
    got = self.attempt(self.rx, 'Float')
    return got if got is not None else self.int_()

  def let(self):
    self._skip()
    start = self.pos

    def typed_name():
      haoda_type = self.rx('Type')
      return haoda_type, self.rx('ID')
    got = self.attempt(typed_name)
    haoda_type, name = got if got is not None else (None, self.rx('ID'))
    self.lit('=')
    return self.node('Let', start, haoda_type=haoda_type, name=name,
                     expr=self.expr())

  def ref(self):
    self._skip()
    start = self.pos
    name = self.rx('ID')
    self.lit('(')
    idx = [int(self.rx('INT'))]
    while self.attempt(self.lit, ',') is not None:
      idx.append(int(self.rx('INT')))
    self.lit(')')
    lat = self.attempt(lambda: (self.lit('~'), self.int_())[1])
    return self.node('Ref', start, name=name, idx=idx, lat=lat)

  def _level(self, depth):
    if depth == len(BINARY_LEVELS):
      return self.unary()
    rule, operators = BINARY_LEVELS[depth]
    self._skip()
    start = self.pos
    operands, used = [self._level(depth + 1)], []

    def tail():
      for op in operators:
        if self.attempt(self.lit, op) is not None:
          return op, self._level(depth + 1)
      self._fail(' or '.join(map(repr, operators)))
    while True:
      more = self.attempt(tail)
      if more is None:
        break
      used.append(more[0])
      operands.append(more[1])
    return self.node(rule, start, operand=operands, operator=used)

  def expr(self):
    return self._level(0)

  def unary(self):
    self._skip()
    start = self.pos
    operators = []
    while True:
      self._skip()
      if self.pos < len(self.text) and self.text[self.pos] in '+-~!':
        operators.append(self.text[self.pos])
        self.pos += 1
      else:
        break
    return self.node('Unary', start, operator=operators,
                     operand=self.operand())

  def operand(self):
    self._skip()
    start = self.pos
    attrs = dict(cast=None, call=None, ref=None, num=None, var=None,
                 expr=None)

    def cast():
      haoda_type = self.rx('Type')
      self.lit('(')
      inner = self.expr()
      self.lit(')')
      return self.node('Cast', start, haoda_type=haoda_type, expr=inner)

    def call():
      name = self.rx('Func')
      self.lit('(')
      args = [self.expr()]
      while self.attempt(self.lit, ',') is not None:
        args.append(self.expr())
      self.lit(')')
      return self.node('Call', start, name=name, arg=args)

    def var():
      name = self.rx('ID')
      idx = []
      while self.attempt(self.lit, '[') is not None:
        idx.append(self.int_())
        self.lit(']')
      return self.node('Var', start, name=name, idx=idx)

    def paren():
      self.lit('(')
      inner = self.expr()
      self.lit(')')
      return inner

    for key, rule in (('cast', cast), ('call', call), ('ref', self.ref),
                      ('num', self.num), ('var', var), ('expr', paren)):
      got = self.attempt(rule)
      if got is not None:
        attrs[key] = got
        return self.node('Operand', start, **attrs)
    self._fail('an operand')


def parse_tree(text):
  """SODA source -> neutral ``(rule, attrs, pos)`` tree (textX attribute names)."""
  parser = _Parser(text)
  try:
    return parser.program()
  except _Backtrack:
    pos, expected = parser.furthest
    line = text.count('\n', 0, pos) + 1
    col = pos - (text.rfind('\n', 0, pos) + 1) + 1
    snippet = text[pos:pos + 20].split('\n')[0]
    raise SodaSyntaxError(
        'expected %s before %r' % (expected, snippet), line, col) from None
