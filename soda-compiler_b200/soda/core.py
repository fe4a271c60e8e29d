"""Stencil IR: the tensor DAG of a SODA program and its window geometry.

API mirror of the part of the reference's ``soda.core`` that defines *what* a
stencil program computes (reference: src/soda/core.py:20-156 Tensor, :179-267
Stencil validation, :329-405 per-iteration tensor replication, :407-554 stage
order, :787-835 window helpers).  The reference computes FPGA reuse-buffer
chains, FIFO delays and a dataflow module graph eagerly in the same class
(:407-777, src/soda/dataflow.py); those model FPGA line buffers and are out of
scope — on the GPU their role is played by the shared-memory plane rings the
CUDA backend sizes from the same window geometry (soda/codegen/cuda/plan.py).
"""
import collections
import functools
import itertools
import logging

from haoda import ir
from haoda import util
from soda import grammar

_logger = logging.getLogger().getChild(__name__)


class Tensor:
  """One input, local or output array and how it is computed.

  Attributes:
    name, haoda_type, c_type
    st_ref: Ref giving the store index of the defining statement (None: input)
    lets, expr: the computation (None for inputs)
    parents / children: OrderedDict name -> Tensor
    ld_refs: OrderedDict parent name -> [Ref] sorted by row-major offset
  """

  def __init__(self, stmt, dim):
    self.haoda_type = stmt.haoda_type
    self._dim = dim
    if isinstance(stmt, grammar.InputStmt):
      self._name, self.st_ref, self.lets, self.expr = stmt.name, None, (), None
    elif isinstance(stmt, grammar.LocalStmtOrOutputStmt):
      self._name, self.st_ref = None, stmt.ref
      self.lets, self.expr = stmt.let, stmt.expr
    else:
      raise util.InternalError('cannot initialize a Tensor from %s' %
                               type(stmt))
    self.parents = collections.OrderedDict()
    self.children = collections.OrderedDict()
    self.ld_refs = collections.OrderedDict()

  @property
  def name(self):
    return self._name if self.st_ref is None else self.st_ref.name

  @property
  def st_idx(self):
    return (0,) * self._dim if self.st_ref is None else self.st_ref.idx

  @property
  def c_type(self):
    return util.get_c_type(self.haoda_type)

  @property
  def ld_indices(self):
    return collections.OrderedDict(
        (name, collections.OrderedDict((ref.idx, ref) for ref in refs))
        for name, refs in self.ld_refs.items())

  def is_input(self):
    return not self.parents

  def is_output(self):
    return not self.children

  def is_producer(self):
    return not self.is_output()

  def is_consumer(self):
    return not self.is_input()

  def visit_loads(self, callback, args=None):
    for let in self.lets:
      let.visit(callback, args)
    self.expr.visit(callback, args)

  def mutate(self, callback, args=None):
    self.lets = tuple(let.visit(callback, args) for let in self.lets)
    self.expr = self.expr.visit(callback, args)
    self.st_ref = self.st_ref.visit(callback, args)

  def loads(self):
    """Every Ref read by the lets and the expression, in evaluation order."""
    if self.expr is None:
      return []
    found = []
    self.visit_loads(lambda obj, _: found.append(obj)
                     if isinstance(obj, ir.Ref) else None)
    return found

  def __str__(self):
    return 'Tensor %s: %s = %s (parents: %s; children: %s)' % (
        self.haoda_type, self.st_ref or self.name, self.expr,
        ', '.join(self.parents), ', '.join(self.children))


def _row_major(idx):
  """Sort key: last (streamed) dimension most significant."""
  return tuple(reversed(idx))


class Stencil:
  """A validated SODA program.

  Constructor keywords are those ``sodac`` passes (reference: src/sodac:109-123):
  burst_width, iterate, app_name, tile_size, unroll_factor, dim, param_stmts,
  input_stmts, local_stmts, output_stmts and optionally dram_in / dram_out.
  burst width, unroll factor, tile sizes and DRAM banks are FPGA mapping
  parameters: they are validated and carried, and ignored by the CUDA backend.
  """

  def __init__(self, **kwargs):
    self.iterate = kwargs.pop('iterate')
    if self.iterate < 1:
      raise util.SemanticError('cannot iterate %d times' % self.iterate)
    self.burst_width = kwargs.pop('burst_width')
    self.app_name = kwargs.pop('app_name')
    self.tile_size = tuple(kwargs.pop('tile_size'))
    self.unroll_factor = kwargs.pop('unroll_factor')
    self.dim = kwargs.pop('dim')
    self.param_stmts = tuple(kwargs.pop('param_stmts'))
    self.input_stmts = tuple(kwargs.pop('input_stmts'))
    self.local_stmts = tuple(kwargs.pop('local_stmts'))
    self.output_stmts = tuple(kwargs.pop('output_stmts'))
    self._apply_dram('input', self.input_stmts, kwargs.pop('dram_in', None),
                     '^')
    self._apply_dram('output', self.output_stmts,
                     kwargs.pop('dram_out', None), ',')

    if self.iterate > 1:
      # iteration i feeds output k back into input k, so they must pair up
      # (reference: src/soda/core.py:228-243)
      if len(self.input_stmts) != len(self.output_stmts):
        raise util.SemanticError(
            'number of input tensors must be the same as output if iterate > '
            '1 times, currently there are %d input(s) but %d output(s)' %
            (len(self.input_stmts), len(self.output_stmts)))
      if self.input_types != self.output_types:
        raise util.SemanticError(
            'input must have the same type(s) as output if iterate > 1 '
            'times, current input has type %s but output has type %s' %
            (util.lst2str(self.input_types),
             util.lst2str(self.output_types)))

    names = [s.name for s in itertools.chain(
        self.input_stmts, self.param_stmts, self.local_stmts,
        self.output_stmts)]
    for name, count in collections.Counter(names).items():
      if count > 1:
        raise util.SemanticError('tensor `%s` is defined %d times' %
                                 (name, count))
    for stmt in itertools.chain(self.local_stmts, self.output_stmts):
      if len(stmt.ref.idx) != self.dim:
        raise util.SemanticError(
            '`%s` is stored with %d indices in a %d-dimensional program' %
            (stmt.name, len(stmt.ref.idx), self.dim))
    self.tensors    # builds and checks the DAG  pylint: disable=pointless-statement

  @classmethod
  def from_text(cls, text, iterate=None, burst_width=None, unroll_factor=None,
                tile_size=None, dram_in=None, dram_out=None):
    """Parse SODA source and build the Stencil with sodac's overrides applied
    (what reference src/sodac:80-123 does between reading the file and
    calling the backend).  ``tile_size``: per-dimension, 0 = keep."""
    model = grammar.parse(text)
    tiles = []
    for d in range(model.dim - 1):
      forced = tile_size[d] if tile_size and d < len(tile_size) else 0
      tiles.append(forced if forced > 0 else model.tile_size[d])
    tiles.append(0)
    pick = lambda given, parsed: parsed if given is None else given
    return cls(burst_width=pick(burst_width, model.burst_width),
               iterate=pick(iterate, model.iterate), dram_in=dram_in,
               dram_out=dram_out, app_name=model.app_name,
               input_stmts=model.input_stmts, param_stmts=model.param_stmts,
               local_stmts=model.local_stmts, output_stmts=model.output_stmts,
               dim=model.dim, tile_size=tiles,
               unroll_factor=pick(unroll_factor, model.unroll_factor))

  @classmethod
  def from_file(cls, path, **overrides):
    with open(path) as handle:
      return cls.from_text(handle.read(), **overrides)

  @staticmethod
  def _apply_dram(kind, stmts, spec, separator):
    """``name:1.2<sep>name2:3`` or ``1.2`` for all (reference :198-226)."""
    if spec is None:
      return
    if ':' not in spec:
      for stmt in stmts:
        stmt.dram = tuple(map(int, spec.split('.')))
      return
    by_name = {stmt.name: stmt for stmt in stmts}
    for item in spec.split(separator):
      name, banks = item.split(':')
      if name not in by_name:
        raise util.SemanticError('no %s named `%s`' % (kind, name))
      by_name[name].dram = tuple(map(int, banks.split('.')))

  # --- names and types ----------------------------------------------------
  input_names = property(lambda self: tuple(s.name for s in self.input_stmts))
  param_names = property(lambda self: tuple(s.name for s in self.param_stmts))
  local_names = property(lambda self: tuple(s.name for s in self.local_stmts))
  output_names = property(
      lambda self: tuple(s.name for s in self.output_stmts))
  input_types = property(
      lambda self: tuple(s.haoda_type for s in self.input_stmts))
  param_types = property(
      lambda self: tuple(s.haoda_type for s in self.param_stmts))
  local_types = property(
      lambda self: tuple(s.haoda_type for s in self.local_stmts))
  output_types = property(
      lambda self: tuple(s.haoda_type for s in self.output_stmts))

  @functools.cached_property
  def symbol_table(self):
    """name -> haoda type of every input, local, output and param."""
    return {stmt.name: stmt.haoda_type for stmt in itertools.chain(
        self.input_stmts, self.local_stmts, self.output_stmts,
        self.param_stmts)}

  def name_in_iter(self, name, iteration):
    """Name of ``name``'s replica in ``iteration`` (reference :337-357).

    Inputs/locals of iteration i>0 are ``<name>_iter<i>``; an output of a
    non-final iteration *is* the next iteration's input of the same position.
    """
    if name in self.param_names:
      return name
    if name in self.output_names and iteration < self.iterate - 1:
      name = self.input_names[self.output_names.index(name)]
      iteration += 1
    elif name not in self.symbol_table:
      raise util.SemanticError('unknown tensor `%s`' % name)
    return name if iteration == 0 or name in self.output_names else (
        '%s_iter%d' % (name, iteration))

  @functools.cached_property
  def tensors(self):
    """OrderedDict name -> Tensor over all ``iterate`` replicas of the stages."""
    tensor_map = collections.OrderedDict(
        (stmt.name, Tensor(stmt, self.dim)) for stmt in self.input_stmts)
    for iteration in range(self.iterate):
      def rename(obj, _, iteration=iteration):
        if isinstance(obj, ir.Ref):
          if obj.name not in self.symbol_table:
            raise util.SemanticError('unknown tensor `%s`' % obj.name)
          if len(obj.idx) != self.dim and obj.name not in self.param_names:
            raise util.SemanticError(
                '`%s` has %d indices in a %d-dimensional program' %
                (obj, len(obj.idx), self.dim))
          obj.haoda_type = self.symbol_table[obj.name]
          obj.name = self.name_in_iter(obj.name, iteration)
        return obj
      replicas = []
      for stmt in itertools.chain(self.local_stmts, self.output_stmts):
        tensor = Tensor(stmt.visit(rename), self.dim)
        # shift so the smallest load index per dimension is 0 (reference
        # :373-379); load-minus-store offsets, the semantics, are unchanged
        loads = [ref for ref in tensor.loads()
                 if ref.name not in self.param_names]
        if loads:
          low = tuple(min(ref.idx[d] for ref in loads)
                      for d in range(self.dim))
          if any(low):
            def shift(obj, _, low=low):
              if isinstance(obj, ir.Ref) and obj.name not in self.param_names:
                obj.idx = tuple(i - s for i, s in zip(obj.idx, low))
              return obj
            tensor.mutate(shift)
        self._type_lets(tensor)
        if tensor.name in tensor_map:
          raise util.SemanticError('tensor `%s` is defined twice' %
                                   tensor.name)
        tensor_map[tensor.name] = tensor
        replicas.append(tensor)
      for tensor in replicas:
        by_parent = collections.OrderedDict()
        for ref in tensor.loads():
          if ref.name not in self.param_names:
            by_parent.setdefault(ref.name, []).append(ref)
        for parent_name, refs in by_parent.items():
          if parent_name not in tensor_map:
            raise util.SemanticError(
                '`%s` reads `%s`, which is not produced before it' %
                (tensor.name, parent_name))
          parent = tensor_map[parent_name]
          parent.children[tensor.name] = tensor
          tensor.parents[parent_name] = parent
          tensor.ld_refs[parent_name] = sorted(
              refs, key=lambda ref: _row_major(ref.idx))
    return tensor_map

  @staticmethod
  def _type_lets(tensor):
    """Give untyped let variables the type of their defining expression."""
    known = {}

    def typed(obj, _):
      if isinstance(obj, ir.Var) and obj.haoda_type is None:
        obj.haoda_type = known.get(obj.name)
      return obj
    lets = []
    for let in tensor.lets:
      let = let.visit(typed)
      known[let.name] = let.haoda_type
      lets.append(let)
    tensor.lets = tuple(lets)
    tensor.expr = tensor.expr.visit(typed)

  @functools.cached_property
  def chronological_tensors(self):
    """Tensors in the order stages run: breadth-first from the inputs.

    A tensor is emitted once all its parents are (reference :407-554; the
    FPGA delay bookkeeping done in the same pass there is not needed).
    """
    ordered = [self.tensors[name] for name in self.input_names]
    done = set(self.input_names)
    queue = collections.deque(ordered)
    while queue:
      for child in queue.popleft().children.values():
        if child.name not in done and done.issuperset(child.parents):
          done.add(child.name)
          ordered.append(child)
          queue.append(child)
    if len(ordered) != len(self.tensors):
      missing = [name for name in self.tensors if name not in done]
      raise util.SemanticError('cannot schedule %s: cyclic or unreachable '
                               'from the inputs' % ', '.join(missing))
    return ordered

  @property
  def producer_tensors(self):
    return [t for t in self.tensors.values() if t.is_producer()]

  @property
  def consumer_tensors(self):
    return [t for t in self.tensors.values() if t.is_consumer()]

  def valid_bounds(self, tensor):
    """Per dimension ``(lo, hi_margin)``: the tensor is defined on
    ``lo[d] <= x[d] < dims[d] - hi_margin[d]``, the region where every
    transitive input access is in bounds — the loop bounds of the
    reference's golden loop (src/soda/codegen/xilinx/host.py:1082-1091).
    """
    window = get_overall_stencil_window(
        tuple(self.tensors[name] for name in self.input_names), tensor)
    low = get_stencil_window_offset(window)
    extent = get_stencil_dim(window)
    return low, tuple(e - l - 1 for e, l in zip(extent, low))

  def __str__(self):
    return 'Stencil %s: %d-d, iterate %d, [%s] -> [%s]' % (
        self.app_name, self.dim, self.iterate, ', '.join(self.input_names),
        ', '.join(self.output_names))


# --- window geometry ---------------------------------------------------------

_window_cache = {}


def get_overall_stencil_window(input_tensor, output_tensor):
  """Sorted tuple of every offset at which ``output_tensor``'s store point
  transitively reads ``input_tensor`` (one tensor, or an iterable whose
  windows are united), relative to the store point (reference :793-830).
  """
  if not isinstance(input_tensor, Tensor):
    points = set()
    for tensor in input_tensor:
      points.update(get_overall_stencil_window(tensor, output_tensor))
    return tuple(sorted(points))
  key = (id(input_tensor), id(output_tensor))
  hit = _window_cache.get(key)
  if hit is not None and hit[0] is input_tensor and hit[1] is output_tensor:
    return hit[2]
  points = set()
  store = output_tensor.st_idx
  for name, refs in output_tensor.ld_refs.items():
    steps = {tuple(i - s for i, s in zip(ref.idx, store)) for ref in refs}
    if name == input_tensor.name:
      points |= steps
    else:
      inner = get_overall_stencil_window(input_tensor,
                                         output_tensor.parents[name])
      points.update(tuple(a + b for a, b in zip(p, q))
                    for p in inner for q in steps)
  result = tuple(sorted(points))
  _window_cache[key] = (input_tensor, output_tensor, result)
  return result


def get_stencil_dim(points):
  """Bounding-box extent of a window per dimension (reference :787-791)."""
  points = list(points)
  return [max(p[d] for p in points) - min(p[d] for p in points) + 1
          for d in range(len(points[0]))]


def get_stencil_window_offset(stencil_window):
  """``-min`` per dimension of a window normalised to store at 0 (:832-835)."""
  points = list(stencil_window)
  return tuple(-min(p[d] for p in points) for d in range(len(points[0])))


def get_stencil_distance(stencil_window, tile_size):
  """Linearised span of a window in a tiled layout (reference :782-785)."""
  def serialize(vec):
    offset, pitch = 0, 1
    for d, v in enumerate(vec):
      offset += v * pitch
      pitch *= tile_size[d]
    return offset
  return (max(serialize(p) for p in stencil_window) +
          serialize(get_stencil_window_offset(stencil_window)))
