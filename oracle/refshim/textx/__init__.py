"""Minimal stand-in for textX (reference requirements.txt:1; not installable
offline) so the UNMODIFIED reference frontend can run in this container.

TEST TOOLING ONLY — nothing in the product imports this.  The only surface the
reference uses is ``metamodel_from_str(grammar, classes=...)`` ->
``.model_from_str(text)`` and ``exceptions.TextXSyntaxError`` (src/sodac:7,80-88,
129).  The grammar string is ignored: the SODA language is parsed by the
repo's own standalone parser (soda-compiler_b200/soda/dsl_parser.py, loaded
by file path because ``soda`` names the reference package in this process),
and the resulting neutral tree is instantiated bottom-up into the classes the
reference passes in, with textX's conventions: keyword arguments named after
grammar attributes, unmatched optional attributes None, repeated attributes
lists, ``_tx_position`` set on every object.
"""
import importlib.util
import os

from textx import exceptions

_PARSER_PATH = os.path.join(
    os.path.dirname(os.path.abspath(__file__)), '..', '..', '..',
    'soda-compiler_b200', 'soda', 'dsl_parser.py')
_spec = importlib.util.spec_from_file_location('_b200_dsl_parser', _PARSER_PATH)
_dsl = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_dsl)


class _Plain:
  """What textX builds for a rule that has no user class (e.g. Partitioning)."""

  def __init__(self, **attrs):
    self.__dict__.update(attrs)


class _MetaModel:
  def __init__(self, classes):
    self._classes = {cls.__name__: cls for cls in classes}

  def _instantiate(self, tree):
    if isinstance(tree, list):
      return [self._instantiate(item) for item in tree]
    if not isinstance(tree, tuple):
      return tree
    rule, attrs, position = tree
    kwargs = {key: self._instantiate(val) for key, val in attrs.items()}
    obj = self._classes.get(rule, _Plain)(**kwargs)
    obj._tx_position = position
    return obj

  def model_from_str(self, text):
    try:
      tree = _dsl.parse_tree(text)
    except _dsl.SodaSyntaxError as e:
      raise exceptions.TextXSyntaxError(str(e), e.line, e.col) from None
    return self._instantiate(tree)

  def model_from_file(self, path):
    with open(path) as handle:
      return self.model_from_str(handle.read())


def metamodel_from_str(grammar, classes=(), **_):
  del grammar   # the language is fixed; see module docstring
  return _MetaModel(classes)
