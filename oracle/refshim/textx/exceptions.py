"""`textx.exceptions` surface used by the reference driver (src/sodac:129)."""


class TextXError(Exception):
  pass


class TextXSyntaxError(TextXError):
  def __init__(self, message, line=None, col=None):
    super().__init__(message)
    self.line, self.col = line, col


class TextXSemanticError(TextXError):
  pass
