"""Scheduling of a SODA stencil onto the streaming GPU kernel.

Two steps, both pure Python so they can be tested without a GPU:

``extract_program(stencil)``
    reads a ``soda.core.Stencil`` (this package's or, duck-typed, the
    reference's) and keeps what execution needs: the stages of ONE iteration
    in dependency order, each with its lowered expression and the relative
    offsets of its loads (the ``ref.idx - st_ref.idx`` of the reference's
    golden loop, src/soda/codegen/xilinx/host.py:1099-1101), plus how outputs
    feed inputs across ``iterate`` (by position, src/soda/core.py:347-351).

``Schedule(program, depth, tile, ...)``
    maps ``depth`` fused iterations onto one kernel.  The model is the
    reference's own dataflow pipeline turned sideways: the last dimension is
    streamed (never tiled: src/soda/grammar.py:34, README.md:248), the others
    are tiled.  A thread block owns one tile of the non-streamed dimensions and
    walks the streamed dimension one *plane* (a row in 2-D, an x-y tile in 3-D)
    per step.  Every tensor of the fused chain lives in a shared-memory ring
    of planes — the GPU image of the reference's reuse buffers
    (src/soda/core.py:612-777) — and every stage computes one plane per step,
    ``delay`` steps behind the newest input plane, reading only planes that
    were completed in earlier steps, so one block barrier per step suffices.

Garbage tolerance.  All stages compute the whole tile; cells whose stencil
reaches outside the tile (or outside the grid, where loads return 0) hold
garbage.  Only cells whose whole transitive window lies inside the tile are
stored, so tiles overlap by the accumulated halo, and in-plane neighbour
addresses are plain linear offsets: a read that wraps into the next row or
the next ring slot only ever feeds a garbage cell.  This removes every
boundary branch from the stage bodies.
"""
import collections

from haoda import util

Load = collections.namedtuple('Load', 'parent off')   # off: offset per dim


def _is_ref(obj):
  return type(obj).__name__ == 'Ref'


class Code:
  """Leaf standing in for a Ref once an emitter has chosen its storage."""

  def __init__(self, code):
    self.c_expr = code

  def __str__(self):
    return self.c_expr

  def visit(self, callback, args=None):
    del callback, args
    return self


def _is_call(obj):
  return type(obj).__name__ == 'Call'


class Stage:
  """One local/output statement of one iteration.

  ``inlined`` maps the names of other stages to their Stage objects: a Ref to
  one of them is not a load but that stage's own expression, spliced in at the
  Ref's offset and rounded through the stage's declared type exactly as the
  store into its array would (``inline_single_use`` builds such stages)."""

  def __init__(self, name, haoda_type, lets, expr, store_idx, is_output,
               params=(), alias=None, inlined=None):
    self.params = frozenset(params)   # names of param arrays
    # replica name -> statement name: with iterate > 1 the Stencil IR calls
    # output k of the first iteration `<input k>_iter1` (reference
    # core.py:347-351), also where a later statement reads it
    self.alias = dict(alias or {})
    self.inlined = dict(inlined or {})
    self.name = name
    self.haoda_type = haoda_type
    self.c_type = util.get_c_type(haoda_type)
    self.lets = tuple(lets)   # ir.Let-like: .c_type .name .expr
    self.expr = expr          # expression tree with Ref leaves
    self.store_idx = tuple(store_idx)
    self.is_output = is_output
    self.loads = []           # unique Load, first-use order
    for load in self.walk_loads():
      if load not in self.loads:
        self.loads.append(load)

  def load_of(self, ref, shift=None):
    """The Load a Ref of this statement stands for; ``shift`` displaces it
    (the offset at which this statement is itself spliced into another)."""
    if ref.name in self.params:
      # a param is a small constant array indexed absolutely (the golden
      # loop reads `<name>_img[i][j]`, reference host.py:1095-1097)
      return Load(ref.name, tuple(ref.idx))
    off = tuple(a - b for a, b in zip(ref.idx, self.store_idx))
    if shift is not None:
      off = tuple(a + b for a, b in zip(off, shift))
    return Load(self.alias.get(ref.name, ref.name), off)

  def walk_loads(self, shift=None):
    """Every load of the statement in first-use order, through inlined
    statements, with repeats."""
    found = []

    def note(obj, _):
      if _is_ref(obj):
        load = self.load_of(obj, shift)
        if load.parent in self.inlined:
          found.extend(self.inlined[load.parent].walk_loads(load.off))
        else:
          found.append(load)
    for node in tuple(let.expr for let in self.lets) + (self.expr,):
      node.visit(note)
    return found

  def calls(self):
    names = []

    def note(obj, _):
      if _is_call(obj):
        names.append(obj.name)
    for node in tuple(let.expr for let in self.lets) + (self.expr,):
      node.visit(note)
    return names

  def top_division(self):
    """True if the statement is a quotient at its top: `A / B` or
    `A * .. / B` (the chain's last operator is the division) and has no lets."""
    ops = getattr(self.expr, 'operator', ())
    return bool(ops) and ops[-1] == '/' and not self.lets

  def render(self, ref_code, call_prefix='', let_prefix='', shift=None,
             cast=None, split_top=False):
    """``(let lines, expression)`` as C, each Ref replaced by ``ref_code(load)``.

    The expression text is the IR's own ``c_expr`` lowering of the tree with
    Refs swapped for storage accesses — the recipe of the golden loop's
    ``mutate_load_for_host`` (host.py:1093-1117), so operator order and
    parenthesisation are the reference's, quirks included.  A Ref to an
    inlined statement becomes ``cast(its C type, its expression)`` — by
    default ``soda::store_cast<T>(..)``, the rounding its array would apply.
    ``call_prefix`` / ``let_prefix`` rename math calls and let variables.
    ``split_top`` (statements with ``top_division``): the expression comes
    back as ``(numerator, denominator)``; a chain `a * b / c` folds from the
    left as C does, so its numerator is `(a * b)`.
    """
    if cast is None:
      cast = lambda c_type, text: 'soda::store_cast<%s>(%s)' % (c_type, text)
    let_names = {let.name for let in self.lets}

    def swap(obj, _):
      if _is_ref(obj):
        load = self.load_of(obj, shift)
        if load.parent in self.inlined:
          other = self.inlined[load.parent]
          _, text = other.render(ref_code, call_prefix, let_prefix, load.off,
                                 cast)
          return Code(cast(other.c_type, text))
        return Code(ref_code(load))
      if call_prefix and _is_call(obj) and not obj.name.startswith(
          call_prefix):
        obj.name = call_prefix + obj.name
      elif (let_prefix and type(obj).__name__ == 'Var' and
            obj.name in let_names):
        obj.name = let_prefix + obj.name
      return obj
    # like the golden loop (host.py:1111-1114) the right-hand side keeps its
    # parentheses: the IR's `unparenthesize` is not bracket-matching and would
    # turn `(a == b) & (c)` into `a == b) & (c`
    lets = ['const %s %s%s = %s;' % (let.c_type, let_prefix, let.name,
                                     let.expr.visit(swap).c_expr)
            for let in self.lets]
    if split_top:
      assert self.top_division()
      texts = [node.visit(swap).c_expr for node in self.expr.operand]
      numerator = texts[0]
      for op, rhs in zip(self.expr.operator[:-1], texts[1:-1]):
        numerator = '(%s %s %s)' % (numerator, op, rhs)
      return lets, (numerator, texts[-1])
    return lets, self.expr.visit(swap).c_expr


class Program:
  """A stencil program reduced to what execution needs."""

  def __init__(self, app_name, dim, iterate, inputs, outputs, stages,
               params=()):
    # params: [(name, haoda_type, size tuple)] — small constant arrays passed
    # next to the tensors (reference header.py:57-60)
    self.param_stmts = [(n, t, tuple(size)) for n, t, size in params]
    params = [n for n, _, _ in self.param_stmts]
    self.app_name = app_name
    self.dim = dim
    self.iterate = iterate
    self.inputs = list(inputs)      # [(name, haoda_type)]
    self.outputs = list(outputs)    # [(name, haoda_type)]
    self.stages = list(stages)      # dependency order, one iteration
    self.params = list(params)
    self.input_names = [n for n, _ in self.inputs]
    self.output_names = [n for n, _ in self.outputs]
    self.types = dict(self.inputs)
    self.types.update((s.name, s.haoda_type) for s in self.stages)
    self.types.update((n, t) for n, t, _ in self.param_stmts)
    # output k of an iteration is input k of the next (by position)
    self.feedback = (dict(zip(self.input_names, self.output_names))
                     if len(self.inputs) == len(self.outputs) else {})

  def c_type(self, name):
    return util.get_c_type(self.types[name])

  def param_index(self, name):
    return self.params.index(name)

  def param_flat(self, load):
    """Flat element index of a param load: the golden loop declares
    ``T <name>_img[s0][s1]..`` and reads ``<name>_img[i][j]..`` (reference
    host.py:1004-1008, 1095-1097), i.e. the first index is the slowest."""
    size = self.param_stmts[self.param_index(load.parent)][2]
    if len(load.off) != len(size):
      raise util.SemanticError('param `%s` has %d dimension(s), indexed with '
                               '%d' % (load.parent, len(size), len(load.off)))
    flat = 0
    for index, extent in zip(load.off, size):
      if not 0 <= index < extent:
        raise util.SemanticError('index %s is outside param `%s`%s' % (
            tuple(load.off), load.parent, list(size)))
      flat = flat * extent + index
    return flat

  def check_params(self):
    for stage in self.stages:
      for load in stage.loads:
        if load.parent in self.params:
          self.param_flat(load)

  def elem_size(self, name):
    return util.get_width_in_bytes(self.types[name])

  def _reach(self, iterations):
    """Per iteration, name -> (lo, hi) box of input offsets, or None."""
    zero = (0,) * self.dim
    reach = {name: (zero, zero) for name in self.input_names}
    history = []
    for it in range(iterations):
      for stage in self.stages:
        boxes = []
        for load in stage.loads:
          box = reach.get(load.parent)
          if box is not None:   # params / input-free stages contribute nothing
            boxes.append((tuple(map(sum, zip(box[0], load.off))),
                          tuple(map(sum, zip(box[1], load.off)))))
        reach[stage.name] = (
            tuple(min(b[0][d] for b in boxes) for d in range(self.dim)),
            tuple(max(b[1][d] for b in boxes) for d in range(self.dim))
        ) if boxes else None
      history.append(dict(reach))
      if it + 1 < iterations:
        reach = {name: reach[self.feedback[name]] for name in self.input_names}
    return history

  def window(self, iterations=None):
    """``(lo, hi)`` per dimension: bounding box of every offset at which ANY
    output of ``iterations`` chained iterations reads the original inputs.
    This union sizes halos, tiles and ghost planes; which cells of an output
    are *defined* is a per-output matter, see ``window_of``."""
    iterations = self.iterate if iterations is None else iterations
    zero = (0,) * self.dim
    if iterations == 0:
      return zero, zero
    last = self._reach(iterations)[-1]
    outs = [last[name] for name in self.output_names if last[name]]
    if not outs:
      return zero, zero
    return (tuple(min(o[0][d] for o in outs) for d in range(self.dim)),
            tuple(max(o[1][d] for o in outs) for d in range(self.dim)))

  def window_of(self, output, iterations=None):
    """``(lo, hi)`` per dimension: bounding box of the offsets at which output
    number ``output`` (or the output of that name) reads the original inputs
    after ``iterations`` chained iterations — the box of the reference's
    overall stencil window from all inputs to THAT tensor
    (core.py:793-830), which is what bounds its golden loop nest
    (host.py:1082-1091).  Outputs of one program may differ."""
    iterations = self.iterate if iterations is None else iterations
    zero = (0,) * self.dim
    if iterations == 0:
      return zero, zero
    name = output if isinstance(output, str) else self.output_names[output]
    return self._reach(iterations)[-1][name] or (zero, zero)

  def check_windows(self):
    """Every stage's window must contain its own store point in every
    dimension.  Otherwise the reference's golden loop itself runs out of
    bounds (its loops start at -min and end at dims-max unclamped,
    host.py:1082-1091), so there is no defined result to reproduce."""
    for reach in self._reach(1):
      for stage in self.stages:
        box = reach.get(stage.name)
        if box is None:
          raise util.SemanticError(
              '`%s` does not depend on any input' % stage.name)
        for d in range(self.dim):
          if box[0][d] > 0 or box[1][d] < 0:
            raise util.SemanticError(
                '`%s` only reads inputs at offsets %d..%d from its store '
                'point in dimension %d: the window must include 0' %
                (stage.name, box[0][d], box[1][d], d))

  def valid_region(self, dims, iterations=None, output=0):
    """``[(lo, hi)]`` per dimension where output ``output`` is defined: the
    bounds of the reference's golden loop for that tensor
    (host.py:1082-1091)."""
    lo, hi = self.window_of(output, iterations)
    return [(max(0, -l), d - max(0, h)) for l, h, d in zip(lo, hi, dims)]

  def valid_regions(self, dims, iterations=None):
    """``valid_region`` of every output, in program order."""
    return [self.valid_region(dims, iterations, k)
            for k in range(len(self.outputs))]


def emit_param_pointers(p, program):
  """Kernel prologue: one typed pointer per param array (device copies made
  by the runtime, passed in StreamArgs::param_ptr)."""
  for k, (name, haoda_type, _) in enumerate(program.param_stmts):
    c_type = util.get_c_type(haoda_type)
    p.println('const %s* const prm_%s = static_cast<const %s*>('
              'a.param_ptr[%d]);' % (c_type, name, c_type, k))
    p.println('(void)prm_%s;' % name)


def param_code(program, load):
  """Device code of a param load: a uniform read-only load the compiler
  hoists out of the streamed loop."""
  return '__ldg(prm_%s + %d)' % (load.parent, program.param_flat(load))


def inline_single_use(program):
  """``program`` with every local statement that is read exactly once — by
  one statement, at one offset — spliced into its reader (None if there is
  nothing to splice).

  A local that is written only to be read back once costs a register history
  (or a shared plane) from the step it is computed to the step it is used;
  spliced in, it is evaluated where it is used, from operands its reader
  mostly holds anyway (denoise3d: six differences, r0 and r1 — 100 of the
  128 registers per thread were histories).  The value is the same: the
  local's expression with the same operands in the same order, rounded
  through the local's declared type as the store into its array rounds it.
  Statements with math calls or let bindings stay (their lazily evaluated
  forms, e.g. the exact `a / sqrt(x)`, are resolved by the store).
  """
  uses = collections.Counter()
  for stage in program.stages:
    for load in set(stage.walk_loads()):
      uses[load.parent] += 1
  by_name = {stage.name: stage for stage in program.stages}
  chosen = [stage.name for stage in program.stages
            if not stage.is_output and uses[stage.name] == 1 and
            not stage.lets and not stage.calls() and not stage.inlined]
  if not chosen:
    return None
  kept = []
  rebuilt = {}
  for stage in program.stages:       # dependency order: parents come first
    inlined = {name: rebuilt.get(name, by_name[name]) for name in chosen
               if any(load.parent == name for load in stage.loads)}
    new = Stage(stage.name, stage.haoda_type, stage.lets, stage.expr,
                stage.store_idx, stage.is_output, stage.params, stage.alias,
                inlined) if inlined else stage
    rebuilt[stage.name] = new
    if stage.name not in chosen:
      kept.append(new)
  result = Program(program.app_name, program.dim, program.iterate,
                   program.inputs, program.outputs, kept,
                   params=program.param_stmts)
  result.types = dict(program.types)    # the spliced locals keep their types
  return result


def extract_program(stencil):
  """soda.core.Stencil (ours or, duck-typed, the reference's) -> Program."""
  n_in = len(stencil.input_stmts)
  local_names = list(stencil.local_names)
  output_names = list(stencil.output_names)
  stage_names = local_names + output_names
  # the first replica of every statement; with iterate > 1 the reference
  # names an output of iteration 0 ``<input>_iter1`` — undo that by position
  replicas = list(stencil.tensors.values())[n_in:n_in + len(stage_names)]
  alias = {}
  if stencil.iterate > 1:
    alias = {in_name + '_iter1': out_name for in_name, out_name in
             zip(stencil.input_names, output_names)}
  by_name = {
      name: Stage(name, tensor.haoda_type, tensor.lets, tensor.expr,
                  tensor.st_ref.idx, name in output_names,
                  params=stencil.param_names, alias=alias)
      for name, tensor in zip(stage_names, replicas)}
  known = set(stencil.input_names) | set(stencil.param_names)
  placed, order, pending = set(known), [], list(stage_names)
  while pending:
    ready = [name for name in pending
             if {load.parent for load in by_name[name].loads} <= placed]
    if not ready:
      bad = {load.parent for name in pending for load in by_name[name].loads}
      bad -= placed | set(stage_names)
      raise util.SemanticError(
          'unknown tensor(s) %s' % ', '.join(sorted(bad)) if bad else
          'cyclic dependency among %s' % ', '.join(pending))
    for name in ready:
      placed.add(name)
      order.append(by_name[name])
      pending.remove(name)
  return Program(
      stencil.app_name, stencil.dim, stencil.iterate,
      list(zip(stencil.input_names, stencil.input_types)),
      list(zip(stencil.output_names, stencil.output_types)),
      order, params=[(stmt.name, stmt.haoda_type, tuple(stmt.size))
                     for stmt in stencil.param_stmts])


# --- the fused, streamed schedule --------------------------------------------

def _pow2(n):
  p = 1
  while p < n:
    p *= 2
  return p


class Node:
  """A tensor of the fused chain: an input plane stream or a stage replica."""

  def __init__(self, index, name, iteration, stage, c_type, elem_size):
    self.index = index
    self.name = name            # program tensor name
    self.iteration = iteration  # fused iteration it belongs to (inputs: -1)
    self.stage = stage          # None for inputs
    self.c_type = c_type
    self.elem_size = elem_size
    self.loads = []             # [(Node, off)] resolved, stage.loads order
    self.consumers = []         # [(Node, off)]
    self.delay = 0              # computes/receives plane (b - delay) at step b
    self.ring_depth = 0         # planes in shared memory (0: no ring)
    self.output_index = None    # position among program outputs if stored
    self.input_index = None

  @property
  def is_input(self):
    return self.stage is None

  @property
  def ident(self):
    return self.name if self.iteration <= 0 else '%s_t%d' % (self.name,
                                                             self.iteration)

  @property
  def avail(self):
    """At step b the newest plane of this tensor a stage may read is b-avail."""
    return self.delay if self.is_input else self.delay + 1


class Schedule:
  """``depth`` iterations of ``program`` fused into one streaming kernel.

  Args:
    depth: iterations fused per launch (temporal blocking depth T).
    tile: extents of the block tile in the non-streamed dims, dim 0 first.
    vec: cells per thread per vector along dim 0.
    threads: threads per block.
    prefetch: input planes requested ahead of the one being consumed.
  """

  style = 'ring'
  tiles_per_block = 1
  paired = False

  @property
  def chain(self):
    """Iterations chained node by node (half of ``depth`` when paired)."""
    return self.depth // 2 if self.paired else self.depth

  def __init__(self, program, depth, tile, vec, threads, prefetch=2):
    self.program = program
    self.depth = depth
    self.tile = tuple(tile)
    self.vec = vec
    self.threads = threads
    self.prefetch = prefetch
    dim = program.dim
    if dim < 2:
      raise util.SemanticError(
          'the streaming kernel needs at least one tiled dimension')
    if len(self.tile) != dim - 1:
      raise util.SemanticError('tile must have %d extents' % (dim - 1))
    if depth > 1 and not program.feedback:
      raise util.SemanticError(
          'cannot fuse iterations: outputs do not pair with inputs')
    self.sdim = dim - 1                      # streamed dimension
    self.plane_elems = 1
    for extent in self.tile:
      self.plane_elems *= extent
    if self.tile[0] % vec:
      raise util.SemanticError('tile width must be a multiple of vec')
    self.plane_vecs = self.plane_elems // vec
    tile_threads = threads // self.tiles_per_block
    if self.plane_vecs % tile_threads:
      raise util.SemanticError('threads must divide the vectors per plane')
    self.vecs_per_thread = self.plane_vecs // tile_threads
    self._build_nodes()
    self._assign_delays()
    self._size_rings()
    self._measure_halos()

  # pitch of each tiled dim in elements within a plane
  def plane_pitch(self, d):
    pitch = 1
    for extent in self.tile[:d]:
      pitch *= extent
    return pitch

  def plane_offset(self, off):
    """Linear in-plane offset of a load (tiled dims only)."""
    return sum(off[d] * self.plane_pitch(d) for d in range(self.sdim))

  def _build_nodes(self):
    program = self.program
    self.nodes = []
    current = {}       # program tensor name -> Node visible to this iteration
    for k, (name, haoda_type) in enumerate(program.inputs):
      node = Node(len(self.nodes), name, -1, None, util.get_c_type(haoda_type),
                  util.get_width_in_bytes(haoda_type))
      node.input_index = k
      self.nodes.append(node)
      current[name] = node
    self.inputs = list(self.nodes)
    self.stage_nodes = []
    for it in range(self.chain):
      for stage in program.stages:
        node = Node(len(self.nodes), stage.name, it, stage, stage.c_type,
                    util.get_width_in_bytes(stage.haoda_type))
        for load in stage.loads:
          if load.parent in program.params:
            continue
          parent = current[load.parent]
          node.loads.append((parent, load.off))
          parent.consumers.append((node, load.off))
        self.nodes.append(node)
        self.stage_nodes.append(node)
        current[stage.name] = node
      if it + 1 < self.chain:
        for name in program.input_names:
          current[name] = current[program.feedback[name]]
    self.outputs = []
    for k, name in enumerate(program.output_names):
      current[name].output_index = k
      self.outputs.append(current[name])

  def _assign_delays(self):
    s = self.sdim
    for node in self.stage_nodes:
      needs = [parent.avail + off[s] for parent, off in node.loads]
      node.delay = max(needs) if needs else 0
    self.out_delay = max(node.delay for node in self.outputs)

  def _size_rings(self):
    s = self.sdim
    for node in self.nodes:
      if not node.consumers:
        node.ring_depth = 0
        continue
      # oldest plane any consumer still reads at step b is b - oldest
      oldest = max(c.delay - off[s] for c, off in node.consumers)
      if node.is_input:
        span = oldest + self.prefetch + 1
      else:
        span = oldest - node.delay + 1
      node.ring_depth = _pow2(max(span, 1))

  def _measure_halos(self):
    lo, hi = self.program.window(self.depth)
    self.window = (lo, hi)
    self.halo_lo = [max(0, -l) for l in lo]
    self.halo_hi = [max(0, h) for h in hi]
    # dim 0 halos round up to whole vectors so owned cells stay vector aligned,
    # and to 16 bytes of every input: a TMA box that starts at a global
    # address not aligned to 16 bytes traps (observed on B200 with 8-byte
    # aligned tile origins: "illegal instruction" at the cp.async.bulk.tensor)
    v = self.vec
    for node in self.inputs:
      v = max(v, 16 // node.elem_size)
    self.tile_halo_lo = [(-(-self.halo_lo[0] // v)) * v] + self.halo_lo[1:-1]
    self.tile_halo_hi = [(-(-self.halo_hi[0] // v)) * v] + self.halo_hi[1:-1]
    self.own = [t - a - b for t, a, b in
                zip(self.tile, self.tile_halo_lo, self.tile_halo_hi)]
    if min(self.own) <= 0:
      raise util.SemanticError(
          'tile %s is not larger than the halo of %d fused iteration(s)' %
          (self.tile, self.depth))
    # steps before/after the owned streamed range [r0, r1):
    # b runs from r0 - lead to r1 - 1 + out_delay
    self.lead = self.halo_lo[self.sdim]
    # any single load's in-plane reach, for the guard zones around the rings
    reach = [abs(self.plane_offset(off)) for node in self.stage_nodes
             for _, off in node.loads]
    self.guard_elems = max(reach) if reach else 0

  def steps(self, rows):
    """Steps a block runs to produce ``rows`` owned streamed planes."""
    return rows + self.lead + self.out_delay

  def smem_bytes(self):
    total = 0
    for node in self.nodes:
      ring = node.ring_depth * self.plane_elems * node.elem_size
      total += -(-ring // 128) * 128
    guard = -(-self.guard_elems * 8 // 128) * 128
    barriers = 8 * max([n.ring_depth for n in self.inputs] + [1]) * \
        len(self.inputs)
    return total + 2 * guard + -(-barriers // 128) * 128

  def describe(self):
    lines = ['schedule %s: depth %d, tile %s, vec %d, %d threads, own %s, '
             'halo -%s +%s, lead %d, out delay %d, smem %d B' % (
                 self.program.app_name, self.depth, self.tile, self.vec,
                 self.threads, self.own, self.tile_halo_lo, self.tile_halo_hi,
                 self.lead, self.out_delay, self.smem_bytes())]
    for node in self.nodes:
      lines.append('  %-16s delay %3d ring %2d %s' % (
          node.ident, node.delay, node.ring_depth,
          '-> out[%d]' % node.output_index
          if node.output_index is not None else ''))
    return '\n'.join(lines)


# --- the register-streaming schedule -----------------------------------------

_PAIR_TOKEN = None
FLAT_GROUPS = 4   # boxes in the input queue of a 2-D register kernel


def arithmetic_weight(program):
  """Rough count of instructions one cell of one iteration costs: operators of
  the lowered expressions, a float division as 10, a math call as 15
  (benchmarks, estimate / measured per cell: blur 8 / 20, sobel2d 16 / 23,
  jacobi2d 5 / 5.7, denoise2d 68 / 85).  Only used to
  tell compute-bound programs from bandwidth-bound ones."""
  total = 0
  for stage in program.stages:
    lets, text = stage.render(lambda load: 'x',
                              cast=lambda c_type, inner: '(%s)' % inner)
    # an IEEE division is a dozen instructions, an integer division by a
    # literal a multiply and a shift
    division = 10 if util.is_float(stage.haoda_type) else 2
    for piece in list(lets) + [text]:
      total += sum(piece.count(op) for op in '+-*<>&|^?')
      total += division * piece.count('/')
    total += 15 * len(stage.calls())
  return total


def bytes_per_cell(program):
  """Compulsory HBM traffic of one pass: every input read once, every output
  written once."""
  return sum(util.get_width_in_bytes(t)
             for _, t in program.inputs + program.outputs)


def max_elem_size(program):
  """Bytes of the widest tensor cell of the program."""
  return max(util.get_width_in_bytes(t) for t in program.types.values())


def pairing_obstacle(program, depth):
  """Why ``depth`` fused iterations of ``program`` cannot run two per
  instruction on packed f32x2 arithmetic (None if they can).

  Pairing evaluates every stage expression on (iteration k, iteration
  k + depth/2) operand pairs with add/mul/fma.rn.f32x2, which round each half
  exactly like the scalar instruction.  That covers float32 tensors and
  expressions built from + - * on tensor cells, float literals with an `f`
  suffix and integer literals (C++ converts those to float first).
  """
  import re
  global _PAIR_TOKEN
  if _PAIR_TOKEN is None:
    _PAIR_TOKEN = re.compile(
        r'\s+|[-+*()]|(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?[fF]|\d+(?![.\deE])')
  if depth < 2 or depth % 2:
    return 'the depth must be even'
  if not program.feedback:
    return 'outputs do not pair with inputs'
  if program.params:
    return 'param tensors'
  for name, haoda_type in program.types.items():
    if haoda_type not in ('float', 'float32'):
      return '`%s` is %s, not float32' % (name, haoda_type)
  for stage in program.stages:
    lets, expr = stage.render(lambda load: '',
                              cast=lambda c_type, text: '(%s)' % text)
    if lets:
      return 'let bindings'
    pos = 0
    while pos < len(expr):
      match = _PAIR_TOKEN.match(expr, pos)
      if not match:
        return '`%s` uses more than + - * on float32 operands' % stage.name
      pos = match.end()
  return None

class RegSchedule(Schedule):
  """``depth`` iterations fused into one register-streaming kernel.

  Same streamed/tiled split as ``Schedule``, but a tensor of the fused chain
  is no longer a shared-memory ring by default.  One warp spans the tile in
  dimension 0 (``tile[0] == 32 * vec``) and every thread walks the streamed
  dimension with its own ``vec`` cells, so

  * a load at offsets (dx, 0, .., dz) is served from the thread's **register
    history** of the parent (the last few streamed planes of its own cells);
    ``dx != 0`` takes the missing cells from the neighbouring lanes with warp
    shuffles — this is the whole story for 2-D programs, which use no shared
    memory for anything but the input queue and no block barrier at all: a
    block is ``warps`` independent strips; each warp keeps a private
    shared-memory ring of ``flat_slots`` input rows that one elected lane
    fills by TMA, a box of ``flat_box`` rows per request, ``prefetch`` rows
    ahead (one mbarrier per box), so rows in flight cost no registers and the HBM latency is covered however deep the
    fused chain is;
  * only a load with an in-plane offset in a dimension other than 0 (3-D:
    dy != 0) goes through a **shared-memory plane ring** of the parent, which
    the producing stage writes next to its registers; such a parent must have
    been written in an earlier step (one block barrier per step), a register
    parent may be produced in the same step, so the pipeline is as short as
    the data dependences allow.  Inputs of 3-D programs arrive by TMA in a
    plane ring as before and are copied to the history once per step.

  Register histories are addressed statically: the streamed loop is unrolled
  ``period`` times and the row of age k (computed k steps ago) of a node
  lives in slot ``(phase - k) mod period``.
  """

  style = 'reg'

  def __init__(self, program, depth, vec, warps, tile_rest=(), prefetch=2,
               paired=False, min_blocks=1, groups=None):
    self.warps = warps
    self.min_blocks = min_blocks     # resident blocks per SM to compile for
    self.tiles_per_block = warps if program.dim == 2 else 1
    self.input_in_smem = program.dim > 2
    # 2-D: the per-warp input queue holds FLAT_GROUPS boxes of `flat_box`
    # rows; one TMA request brings a whole box (requests of a single 512-byte
    # row are bound by the TMA unit's request rate, ~1 per 50 cycles per SM).
    # A box is requested (FLAT_GROUPS - 2) boxes (about `prefetch` rows) ahead
    # of its first use, into the slots of the box consumed before the previous
    # one.
    self.flat_groups = groups or FLAT_GROUPS
    if self.flat_groups < 3:
      raise util.SemanticError('the input queue holds at least 3 boxes')
    self.paired = paired
    if paired:
      why = pairing_obstacle(program, depth)
      if why:
        raise util.SemanticError('cannot pair iterations: ' + why)
    super().__init__(program, depth, (32 * vec,) + tuple(tile_rest), vec,
                     32 * warps, prefetch)
    # A box is a whole number of history periods, so that one trip of the
    # streamed loop consumes exactly one box: which row of the box a step
    # reads and which register slot it fills are both compile-time constants.
    self.flat_box = 0
    if program.dim == 2:
      want = max(1, prefetch // (self.flat_groups - 2))
      self.flat_box = self.period * max(1, (want + self.period // 2) //
                                        self.period)
    self.flat_slots = self.flat_groups * self.flat_box

  def via_smem(self, off):
    """Does a load at offset ``off`` need the parent's shared plane?"""
    return any(off[1:self.sdim])

  def _assign_delays(self):
    s = self.sdim
    for node in self.stage_nodes:
      needs = []
      for parent, off in node.loads:
        lag = 1 if self.via_smem(off) and not parent.is_input else 0
        needs.append(parent.delay + lag + off[s])
      node.delay = max(needs) if needs else 0
    self.out_delay = max(node.delay for node in self.outputs)
    if self.paired:
      # lane B (iterations chain .. depth-1) trails lane A by `pair_lag`
      # steps: it is fed with lane A's newest output row one step later
      if len({node.delay for node in self.outputs}) != 1:
        raise util.SemanticError('cannot pair iterations: outputs are '
                                 'produced at different delays')
      self.pair_lag = self.out_delay + 1
      self.out_delay += self.pair_lag

  def _size_rings(self):
    s = self.sdim
    spans = [1]
    for node in self.nodes:
      reg_ages = [c.delay - off[s] for c, off in node.consumers
                  if not self.via_smem(off)]
      smem_ages = [c.delay - off[s] for c, off in node.consumers
                   if self.via_smem(off)]
      # registers hold the rows of age node.delay .. hist_oldest
      node.hist_oldest = max(reg_ages) if reg_ages else None
      node.hist_newest = node.delay
      # planes a shared ring must hold (rounded to a usable depth below)
      if node.is_input and self.input_in_smem:
        node.ring_depth = (max(smem_ages + [0]) + self.prefetch + 1
                           if node.consumers else 0)
      elif smem_ages:
        node.ring_depth = max(smem_ages) - node.delay + 1
      else:
        node.ring_depth = 0
      if reg_ages:
        spans.append(node.hist_oldest - node.hist_newest + 1)
    self.period = max(spans)
    # One trip of the streamed loop is `trip` steps, fully unrolled: a whole
    # number of history periods, and a multiple of every ring depth, so that
    # register slots AND shared-memory ring slots are compile-time constants
    # (no per-step address arithmetic).  All input rings share one depth (one
    # mbarrier per slot covers every input).
    needs = [n.ring_depth for n in self.nodes if n.ring_depth]
    self.trip = self.period
    while needs and self.trip < max(needs):
      self.trip += self.period
    in_need = max([n.ring_depth for n in self.inputs] + [0])
    for node in self.nodes:
      if node.ring_depth:
        need = in_need if node.is_input else node.ring_depth
        node.ring_depth = min(d for d in range(need, self.trip + 1)
                              if self.trip % d == 0)
    if in_need and self.input_in_smem:
      # slots the rounding added are used: request further ahead, keeping one
      # slot of slack (heat3d depth 2, ring of 6: 2 ahead 1122, 3 ahead 1161,
      # 4 ahead 1124 GCell/s)
      depth = max(n.ring_depth for n in self.inputs)
      self.prefetch = max(self.prefetch, self.prefetch + depth - in_need - 1)

  def _measure_halos(self):
    super()._measure_halos()
    reach = [abs(self.plane_offset(off)) for node in self.stage_nodes
             for _, off in node.loads if self.via_smem(off)]
    self.guard_elems = max(reach) if reach else 0
    # Planes of a shared ring sit `ring_pitch` elements apart: a neighbour
    # read that leaves its plane (a halo row's y - 1 or y + 1: garbage that
    # only feeds garbage) lands in the gap between two planes, which nobody
    # writes, instead of in the next slot, which another warp may be writing
    # in the same step — harmless either way, but only this way is the
    # kernel free of shared-memory races (compute-sanitizer racecheck,
    # tests/test_sanitizer_gpu.py).  Whole 128-element units keep every slot
    # aligned for TMA.
    self.ring_gap = -(-self.guard_elems // 128) * 128
    self.ring_pitch = self.plane_elems + self.ring_gap

  def describe(self):
    lines = ['%sregister-streaming schedule %s: depth %d, tile %s x %d/block, '
             'vec %d, %d threads, period %d, own %s, halo -%s +%s, lead %d, '
             'out delay %d' % (
                 'paired (f32x2: iterations k and k+%d share an instruction) '
                 % self.chain if self.paired else '',
                 self.program.app_name, self.depth, self.tile,
                 self.tiles_per_block, self.vec, self.threads, self.period,
                 self.own, self.tile_halo_lo, self.tile_halo_hi, self.lead,
                 self.out_delay)]
    for node in self.nodes:
      lines.append('  %-16s delay %3d regs %s ring %2d %s' % (
          node.ident, node.delay,
          'ages %d..%d' % (node.hist_newest, node.hist_oldest)
          if node.hist_oldest is not None else '-', node.ring_depth,
          '-> out[%d]' % node.output_index
          if node.output_index is not None else ''))
    return '\n'.join(lines)
