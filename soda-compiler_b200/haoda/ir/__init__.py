"""Expression IR for SODA stencil stages and its lowering to C expressions.

This is the expression half of the reference's ``haoda.ir``
(reference: src/haoda/ir/__init__.py:13-64 grammar, :66-349 node classes,
:857-881 helpers), rebuilt for the CUDA backend.  The FPGA-only nodes of the
reference (FIFO, Module, DelayedRef, FIFORef, DRAMRef, ModuleTrait; :351-855)
model the dataflow micro-architecture and have no counterpart here.

Two properties of the reference are load-bearing for result parity and are
reproduced on purpose:

* ``c_expr`` of a compound binary node is ``parenthesize(joined operands)``
  and ``parenthesize`` first strips *every* leading ``(`` / trailing ``)``
  pair it sees, matched or not (reference :874-881).  ``x - ((a)+(b))``
  therefore lowers to ``x - (a) + (b)``.  The reference's CPU golden loop
  and its FPGA kernel both evaluate that string, so it *is* the semantics.
* literals are carried verbatim (``0.2f``, ``.125f``, ``3``); arithmetic
  types are whatever C++ deduces from the emitted string.

Trees are built already flattened: single-operand precedence levels,
parenthesised sub-expressions and identity unary chains never materialise
(what reference ``arithmetic.base.flatten`` produces, src/haoda/ir/arithmetic/
base.py:16-93).
"""
import copy

from haoda import util


def unparenthesize(expr):
  """Strip outer '(' ... ')' characters the way the reference does.

  Deliberately not bracket-matching (reference: src/haoda/ir/__init__.py:877-881).
  """
  text = str(expr)
  while text[:1] == '(' and text[-1:] == ')':
    text = text[1:-1]
  return text


def parenthesize(expr):
  return '(' + unparenthesize(expr) + ')'


def str2int(text, none_val=None):
  """C-style integer literal (with U/L suffixes, 0x/0b/0 prefixes) -> int."""
  if text is None:
    return none_val
  digits = text.rstrip('UuLl')
  sign = 1
  if digits[:1] in '+-':
    sign = -1 if digits[0] == '-' else 1
    digits = digits[1:]
  head = digits[:2].lower()
  if head == '0x':
    return sign * int(digits, 16)
  if head == '0b':
    return sign * int(digits, 2)
  if len(digits) > 1 and digits[0] == '0':
    return sign * int(digits, 8)
  return sign * int(digits)


def literal_type(text):
  """haoda type of a numeric literal (reference: src/haoda/ir/__init__.py:298-311)."""
  low = text.lower()
  if 'u' in low:
    return 'uint64' if 'll' in low else 'uint32'
  if 'll' in low:
    return 'int64'
  if 'fl' in low:
    return 'double'
  if 'f' in low or 'e' in low:
    return 'float'
  if '.' in text:
    return 'double'
  return 'int32'


class Node:
  """Base of all IR nodes: value-comparable, hashable, copy-on-visit."""
  FIELDS = ()   # attributes holding one value / child
  LISTS = ()    # attributes holding a tuple of values / children

  def __init__(self, **kwargs):
    for name in self.FIELDS:
      setattr(self, name, kwargs.pop(name, None))
    for name in self.LISTS:
      setattr(self, name, tuple(kwargs.pop(name, ())))
    if kwargs:
      raise TypeError('%s got unexpected attributes %s' %
                      (type(self).__name__, sorted(kwargs)))

  def _key(self):
    return (type(self).__name__,
            tuple(getattr(self, name) for name in self.FIELDS),
            tuple(getattr(self, name) for name in self.LISTS))

  def __eq__(self, other):
    return isinstance(other, Node) and self._key() == other._key()

  def __hash__(self):
    return hash(self._key())

  def __repr__(self):
    return '%s<%s>' % (type(self).__name__, self)

  @property
  def c_type(self):
    return util.get_c_type(self.haoda_type)

  @property
  def width_in_bits(self):
    return util.get_width_in_bits(self.haoda_type)

  def visit(self, callback, args=None):
    """Rebuild the tree top-down through ``callback(node_copy, args)``.

    The callback receives a shallow copy.  Returning a different object
    substitutes it and stops descending; returning the copy (or None) keeps
    it and descends into its children.  The receiver is never modified
    (same contract as reference src/haoda/ir/__init__.py:99-155).
    """
    mine = copy.copy(self)
    got = callback(mine, args)
    if got is not None and got is not mine:
      return got
    for name in mine.FIELDS:
      child = getattr(mine, name)
      if isinstance(child, Node):
        setattr(mine, name, child.visit(callback, args))
    for name in mine.LISTS:
      setattr(mine, name, tuple(
          item.visit(callback, args) if isinstance(item, Node) else item
          for item in getattr(mine, name)))
    return mine


class Num(Node):
  """A numeric literal, kept as written in the source."""
  FIELDS = ('text',)

  def __str__(self):
    return self.text

  @property
  def c_expr(self):
    return self.text

  @property
  def haoda_type(self):
    return literal_type(self.text)


class Var(Node):
  """A named scalar: a ``let`` variable, or raw code spliced by an emitter."""
  FIELDS = ('name', 'haoda_type')
  LISTS = ('idx',)

  def __str__(self):
    return self.name + ''.join('[%s]' % i for i in self.idx)

  c_expr = property(__str__)


def make_var(code, haoda_type=None):
  """Wrap emitter-produced code so it can stand in for a Ref in a tree."""
  return Var(name=code, haoda_type=haoda_type, idx=())


class Ref(Node):
  """``name(i, j, ...)``: one element of a tensor, relative to the store point."""
  FIELDS = ('name', 'lat', 'haoda_type')
  LISTS = ('idx',)

  def __init__(self, **kwargs):
    super().__init__(**kwargs)
    if isinstance(self.lat, str):
      self.lat = str2int(self.lat)

  def __str__(self):
    text = '%s(%s)' % (self.name, ', '.join(map(str, self.idx)))
    return text if self.lat is None else '%s ~%d' % (text, self.lat)
  # No c_expr: a Ref only has meaning once an emitter maps it to storage
  # (reference behaviour too: host.py:1093-1110 rewrites every Ref first).


class Let(Node):
  """``[type] name = expr`` local binding of a stage."""
  FIELDS = ('declared_type', 'name', 'expr')

  @property
  def haoda_type(self):
    return self.declared_type or self.expr.haoda_type

  def __str__(self):
    text = '%s = %s' % (self.name, unparenthesize(self.expr))
    return text if self.declared_type is None else (
        '%s %s' % (self.declared_type, text))

  @property
  def c_expr(self):
    return 'const %s %s = %s;' % (self.c_type, self.name,
                                  unparenthesize(self.expr.c_expr))


class BinaryOp(Node):
  """A chain ``o0 op1 o1 op2 o2 ...`` at one C precedence level (>= 2 operands)."""
  LISTS = ('operand', 'operator')

  def _join(self, render):
    parts = [render(self.operand[0])]
    for op, rhs in zip(self.operator, self.operand[1:]):
      parts += [op, render(rhs)]
    return parenthesize(' '.join(parts))

  def __str__(self):
    return self._join(str)

  @property
  def c_expr(self):
    return self._join(lambda node: node.c_expr)

  @property
  def haoda_type(self):
    return self.operand[0].haoda_type


# One class per precedence level, lowest binding first
# (reference grammar: src/haoda/ir/__init__.py:30-55).
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
for _name, _ops in BINARY_LEVELS:
  globals()[_name] = type(_name, (BinaryOp,), {'OPERATORS': _ops,
                                               '__doc__': ' '.join(_ops)})
del _name, _ops


class Unary(Node):
  """Prefix operators applied right-to-left: ``-!x``."""
  FIELDS = ('operand',)
  LISTS = ('operator',)

  def __str__(self):
    return ''.join(self.operator) + str(self.operand)

  @property
  def c_expr(self):
    return ''.join(self.operator) + self.operand.c_expr

  @property
  def haoda_type(self):
    return self.operand.haoda_type


def is_identity_unary(operators):
  """True if a prefix-operator chain is a no-op.

  Only '+'/'-' with an even number of '-', or an even number of '!' alone
  (reference: src/haoda/ir/arithmetic/base.py:78-86).
  """
  ops = tuple(operators)
  if all(op in '+-' for op in ops) and ops.count('-') % 2 == 0:
    return True
  return all(op == '!' for op in ops) and len(ops) % 2 == 0


class Cast(Node):
  """``type(expr)`` -> ``static_cast<ctype >(expr)``."""
  FIELDS = ('haoda_type', 'expr')

  def __str__(self):
    return self.haoda_type + parenthesize(self.expr)

  @property
  def c_expr(self):
    return 'static_cast<%s >%s' % (self.c_type, parenthesize(self.expr.c_expr))


class Call(Node):
  """``func(arg, ...)`` from the DSL's math-function whitelist."""
  FIELDS = ('name',)
  LISTS = ('arg',)

  def __str__(self):
    return '%s(%s)' % (self.name, ', '.join(map(str, self.arg)))

  @property
  def c_expr(self):
    return '%s(%s)' % (self.name, ', '.join(a.c_expr for a in self.arg))

  @property
  def haoda_type(self):
    return self.arg[1 if self.name == 'select' else 0].haoda_type


def collect(node, node_type):
  """All sub-nodes of ``node_type`` in evaluation (left-to-right) order."""
  found = []

  def look(obj, _):
    if isinstance(obj, node_type):
      found.append(obj)
  node.visit(look)
  return found
