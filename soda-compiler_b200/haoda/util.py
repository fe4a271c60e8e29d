"""Type map, error classes and a small source-text writer.

API mirror of the parts of the reference's ``haoda.util`` that sit on the
stencil execution path (reference: src/haoda/util.py:18-25 error classes,
:145-180 type helpers, :27-131 Printer).  Written from scratch for the CUDA
backend; the FPGA-only helpers (module/port/bundle names) are out of scope.
"""
import contextlib

# coordinate letters used by the reference's emitted host code
# (reference: src/haoda/util.py:6-8); the CUDA emitter reuses the
# "original coordinates" set so generated code reads like the golden loop.
COORDS_TILED = 'xyzw'
COORDS_IN_TILE = 'ijkl'
COORDS_IN_ORIG = 'pqrs'

_NAMED_FLOAT_BITS = {'half': 16, 'float': 32, 'double': 64}
_STD_INT_BITS = (8, 16, 32, 64)


class InternalError(Exception):
  """A bug in the compiler itself."""


class SemanticError(Exception):
  """The SODA program parsed but is not meaningful (sodac exits 1)."""


class SemanticWarn(Exception):
  """A recoverable oddity (sodac logs it and exits 0)."""


def _split_type(haoda_type):
  """'uint16' -> ('uint', 16); 'float32' -> ('float', 32); 'half' -> ('float', 16)."""
  if haoda_type in _NAMED_FLOAT_BITS:
    return 'float', _NAMED_FLOAT_BITS[haoda_type]
  for prefix in ('uint', 'int', 'float'):
    if haoda_type.startswith(prefix):
      digits = haoda_type[len(prefix):].split('_')[0]
      if digits.isdigit():
        return prefix, int(digits)
  raise InternalError('unknown haoda type: %s' % haoda_type)


def get_c_type(haoda_type):
  """haoda type name -> C type name (reference: src/haoda/util.py:145-159).

  8/16/32/64-bit integers map to <stdint.h> names, float32/float64 to
  float/double, other integer widths to the HLS ``ap_(u)int<N>`` spelling
  (which the CUDA backend rejects), anything else passes through.
  """
  if haoda_type is None:
    return None
  if haoda_type in ('float32', 'float64'):
    return 'float' if haoda_type == 'float32' else 'double'
  for prefix in ('uint', 'int'):
    if haoda_type.startswith(prefix):
      width = haoda_type[len(prefix):]
      if width.isdigit() and int(width) in _STD_INT_BITS:
        return haoda_type + '_t'
      return 'ap_%s<%s>' % (prefix, width)
  return haoda_type


def get_haoda_type(c_type):
  return c_type[:-2] if c_type.endswith('_t') else c_type


def get_width_in_bits(haoda_type):
  """Bit width of a haoda type, or of anything carrying ``.haoda_type``."""
  if not isinstance(haoda_type, str):
    if hasattr(haoda_type, 'haoda_type'):
      return get_width_in_bits(haoda_type.haoda_type)
    raise InternalError('unknown haoda type: %s' % (haoda_type,))
  return _split_type(haoda_type)[1]


def get_width_in_bytes(haoda_type):
  return (get_width_in_bits(haoda_type) + 7) // 8


def is_float(haoda_type):
  return haoda_type in ('half', 'double') or haoda_type.startswith('float')


def idx2str(idx):
  return '(%s)' % ', '.join(map(str, idx))


def lst2str(idx):
  return '[%s]' % ', '.join(map(str, idx))


class Printer:
  """Indenting line writer used by the code emitters.

  Same surface as the reference's Printer (println / do_indent / un_indent /
  do_scope / un_scope / new_var / last_var / for_ / if_) so emitters written
  against either read alike.
  """

  def __init__(self, out, tab=2):
    self._out = out
    self._tab = tab
    self._level = 0
    self._scope_notes = []
    self._var_count = 0

  def println(self, line='', indent=-1):
    if not line:
      self._out.write('\n')
      return
    level = self._level if indent < 0 else indent
    self._out.write(' ' * (level * self._tab) + line + '\n')

  def printlns(self, *lines):
    for line in lines:
      self.println(line)

  def do_indent(self):
    self._level += 1

  def un_indent(self):
    self._level -= 1

  def do_scope(self, comment=''):
    self.println('{')
    self._level += 1
    self._scope_notes.append(comment)

  def un_scope(self, comment='', suffix=''):
    self._level -= 1
    opened_with = self._scope_notes.pop()
    note = comment or opened_with
    self.println('}%s%s' % (suffix, ' // %s' % note if note else ''))

  def new_var(self):
    self._var_count += 1
    return self.last_var()

  def last_var(self, offset=-1):
    return 'assign_%d' % (self._var_count + 1 + offset)

  @contextlib.contextmanager
  def for_(self, *args):
    self.println('for (%s)' % ('; '.join(args) if len(args) == 3
                               else ' : '.join(args)))
    self.do_scope()
    yield
    self.un_scope()

  @contextlib.contextmanager
  def if_(self, cond):
    self.println('if (%s)' % cond)
    self.do_scope()
    yield
    self.un_scope()


def print_define(printer, var, val):
  printer.printlns('#ifndef %s' % var, '#define %s %d' % (var, val),
                   '#endif//%s' % var)


def print_guard(printer, var, val):
  printer.printlns('#ifdef %s' % var, '#if %s != %d' % (var, val),
                   '#error %s != %d' % (var, val),
                   '#endif//%s != %d' % (var, val), '#endif//%s' % var)
