"""CUDA (sm_100a) backend of sodac: ``--cuda-kernel`` / ``--cuda-host``.

Same plugin shape as the reference's Xilinx OpenCL backend
(reference src/soda/codegen/xilinx/opencl.py:75-95 ``add_arguments``, :97-136
``print_code``; wired at src/sodac:12,66,127): ``add_arguments(group)`` adds the
backend's options to the sodac parser and ``print_code(stencil, args)`` writes
each requested artefact to a file or, for ``-``, to stdout.  The stencil is
not modified.  Unsupported programs raise ``haoda.util.SemanticError`` (sodac
exits 1).

Artefacts
  --cuda-kernel FILE   the streaming kernels (.cu) + their variant table
  --cuda-host FILE     program descriptor, ``<app>()`` entry, C ABI (.cpp/.cu)
  --cuda-header FILE   ``<app>.h`` as the reference emits it (buffer_t +
                       prototype), for callers of ``<app>()``

Build: ``nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -shared``
over the two files plus csrc/soda_cuda_runtime.cu (``soda.cuda.build`` does it).
"""
import math
import shutil
import sys
import tempfile

from haoda import util
from soda.codegen.cuda import host as host_mod
from soda.codegen.cuda import kernel as kernel_mod
from soda.codegen.cuda import kernel_reg as kernel_reg_mod
from soda.codegen.cuda import plan as plan_mod
from soda.codegen.cuda import tuned as tuned_mod

SUPPORTED_TYPES = {
    'uint8', 'uint16', 'uint32', 'uint64', 'int8', 'int16', 'int32', 'int64',
    'float', 'float32', 'double', 'float64', 'half'}
SMEM_LIMIT = 227 * 1024
REG_HISTORY_BUDGET = 80   # registers per thread held across steps (2-D)
REG_HISTORY_3D = 64       # 3-D: above this, one vector per thread
FLAT_SMEM_TARGET = 75 * 1024   # 2-D input queues per block: 3 blocks per SM


def add_arguments(parser):
  parser.add_argument(
      '--cuda-kernel', type=str, dest='cuda_kernel_file', metavar='file',
      help='CUDA kernel code (sm_100a) for the B200 backend')
  parser.add_argument(
      '--cuda-host', type=str, dest='cuda_host_file', metavar='file',
      help='host C++ code for the B200 backend: defines <app>() and the C ABI')
  parser.add_argument(
      '--cuda-header', type=str, dest='cuda_header_file', metavar='file',
      help='host C++ header declaring <app>() (same as --xocl-header)')
  parser.add_argument(
      '--cuda-temporal-depth', type=int, dest='cuda_depth', metavar='T',
      help='iterations fused per kernel launch (default: chosen per program)')
  parser.add_argument(
      '--cuda-tile', type=int, nargs='+', dest='cuda_tile', metavar='N',
      help='thread-block tile in the non-streamed dimensions')
  parser.add_argument(
      '--cuda-threads', type=int, dest='cuda_threads', metavar='N',
      help='threads per block')
  parser.add_argument(
      '--cuda-vec', type=int, dest='cuda_vec', metavar='N',
      help='cells per thread per vector access along dimension 0')
  parser.add_argument(
      '--cuda-prefetch', type=int, dest='cuda_prefetch', metavar='N',
      help='input planes requested ahead of the one being consumed')
  parser.add_argument(
      '--cuda-groups', type=int, dest='cuda_groups', metavar='N',
      help='2-D kernels: TMA boxes in the per-warp input queue (3, 4, 8, ..); '
      'a box holds prefetch / (N - 2) rows')
  parser.add_argument(
      '--cuda-min-blocks', type=int, dest='cuda_min_blocks', metavar='N',
      help='resident blocks per SM the kernel is compiled for (caps its '
      'registers)')
  parser.add_argument(
      '--cuda-paired', type=int, dest='cuda_paired', metavar='0|1',
      help='evaluate fused iterations k and k + depth/2 together on packed '
      'f32x2 arithmetic (default: whenever the program allows it)')
  parser.add_argument(
      '--cuda-autotune', type=int, nargs='+', dest='cuda_autotune',
      metavar='N',
      help='grid extents to tune the kernel configuration for, on the GPU of '
      'this machine (times the candidates of soda.cuda_tune and records the '
      'winner in tuned.json before emitting); the GPU counterpart of '
      'exploring --tile-size / --unroll-factor')
  parser.add_argument(
      '--cuda-inline', type=int, dest='cuda_inline', metavar='0|1',
      help='splice every local that is read exactly once into its reader '
      'instead of keeping it in registers between the step that computes it '
      'and the step that uses it (bit-identical; default 0: measured no '
      'faster)')
  parser.add_argument(
      '--cuda-fast', action='store_true', dest='cuda_fast',
      help='emit the kernels for the non-exact build (FMA contraction, '
      'approximate division, a / sqrt(x) as a refined rsqrt; within 1e-6 '
      'relative or 2 ulp of the reference, not bit-exact): the kernel file '
      'defines SODA_CUDA_FAST_MATH; compile it WITHOUT -fmad=false and with '
      '-prec-div=false -prec-sqrt=false')
  parser.add_argument(
      '--cuda-exact', action='store_false', dest='cuda_fast',
      help='emit the kernels for the bit-exact build (the default): '
      'reference operation order, no FMA contraction; compile with '
      '-fmad=false')
  parser.add_argument(
      '--cuda-style', type=str, dest='cuda_style', choices=['reg', 'ring'],
      help='kernel family: `reg` keeps the streamed window of every tensor '
      'in registers and shares dimension-0 neighbours by warp shuffle '
      '(default); `ring` keeps every tensor in a shared-memory plane ring')


class Options:
  """Tuning knobs of the backend; None means "choose for me"."""

  def __init__(self, depth=None, tile=None, threads=None, vec=None,
               prefetch=None, style=None, paired=None, min_blocks=None,
               groups=None, inline=None):
    self.depth, self.tile, self.threads = depth, tile, threads
    self.vec, self.prefetch, self.style = vec, prefetch, style
    self.paired = None if paired is None else bool(paired)
    self.min_blocks = min_blocks
    self.groups = groups
    # splice locals that are read once into their reader (default: no)
    self.inline = None if inline is None else bool(inline)

  @classmethod
  def from_args(cls, args):
    return cls(depth=getattr(args, 'cuda_depth', None),
               tile=getattr(args, 'cuda_tile', None),
               threads=getattr(args, 'cuda_threads', None),
               vec=getattr(args, 'cuda_vec', None),
               prefetch=getattr(args, 'cuda_prefetch', None),
               style=getattr(args, 'cuda_style', None),
               paired=getattr(args, 'cuda_paired', None),
               min_blocks=getattr(args, 'cuda_min_blocks', None),
               groups=getattr(args, 'cuda_groups', None),
               inline=getattr(args, 'cuda_inline', None))

  def is_default(self):
    return all(v is None for v in vars(self).values())

  def key(self):
    return 'd%s_t%s_n%s_v%s_p%s_%s_%s_%s_g%s_i%s' % (
        self.depth, 'x'.join(map(str, self.tile)) if self.tile else None,
        self.threads, self.vec, self.prefetch, self.style, self.paired,
        self.min_blocks, self.groups, self.inline)


# Whitelisted by the grammar (reference src/soda/grammar.py:25-32) but not
# callable from a SODA expression: their C prototypes take a pointer or a
# string, which the DSL has no way to write (the reference's generated code
# does not compile for them either).
UNCALLABLE = {'frexp': 'an `int*` exponent', 'modf': 'a `double*` integral '
              'part', 'remquo': 'an `int*` quotient', 'nan': 'a string'}


def _check_calls(program):
  def visit(obj, _):
    if type(obj).__name__ == 'Call' and obj.name in UNCALLABLE:
      raise util.SemanticError(
          '`%s` needs %s argument, which a SODA expression cannot supply' %
          (obj.name, UNCALLABLE[obj.name]))
    return obj
  for stage in program.stages:
    for node in tuple(let.expr for let in stage.lets) + (stage.expr,):
      node.visit(visit)


def check_supported(program):
  _check_calls(program)
  for name, haoda_type in list(program.types.items()):
    if haoda_type not in SUPPORTED_TYPES:
      raise util.SemanticError(
          'type `%s` of `%s` is not supported by the CUDA backend' %
          (haoda_type, name))
  if len(program.params) > 8:
    raise util.SemanticError('at most 8 params')
  program.check_params()
  if not 2 <= program.dim <= 4:
    raise util.SemanticError('the CUDA backend handles 2- to 4-dimensional '
                             'programs, not %d' % program.dim)
  if len(program.inputs) > 8 or len(program.outputs) > 8:
    raise util.SemanticError('at most 8 inputs and 8 outputs')
  program.check_windows()


def default_vec(program):
  widest = max(util.get_width_in_bytes(t) for t in program.types.values())
  return max(1, 16 // widest)


def _default_tiles(program, vec):
  """Candidate block tiles, preferred first."""
  if program.dim == 2:
    return [(128 * vec,), (64 * vec,), (32 * vec,)]
  if program.dim == 3:
    return [(16 * vec, 16), (16 * vec, 8), (8 * vec, 8)]
  return [(8 * vec, 8, 4), (8 * vec, 4, 4)]


def layout_of(sched):
  """Shared-memory layout of a schedule (by kernel family)."""
  if sched.style == 'reg':
    return kernel_reg_mod.Layout(sched)
  return kernel_mod.Layout(sched)


def _make_reg_schedule(program, depth, options, limit):
  """Register-streaming schedule: the caller's knobs where given, otherwise
  the candidate with the least halo overhead that fits."""
  vec = options.vec or min(default_vec(program), 8)
  if program.dim == 2:
    warps = (options.threads or 128) // 32
    paired = (options.paired if options.paired is not None else
              plan_mod.pairing_obstacle(program, depth) is None)
    # boxes in the per-warp input queue: four (two in flight) for programs
    # the HBM rate bounds; three for unpaired compute-bound ones, where the
    # shared memory saved buys resident warps and one box in flight is ample
    # (denoise2d 271 -> 318 GCell/s; sobel2d / blur / seidel2d lose 5-13 %
    # with three, paired jacobi2d x8 is neutral: profiles capture r2v)
    compute_bound = (not paired and plan_mod.arithmetic_weight(program) *
                     depth >= 5 * plan_mod.bytes_per_cell(program))
    groups = options.groups or (3 if compute_bound else plan_mod.FLAT_GROUPS)
    # rows per TMA box (measured, profiles/r1i): 6 keeps six blocks resident;
    # paired kernels hold fewer, fatter warps and want deeper queues
    prefetch = options.prefetch if options.prefetch is not None else (
        (9 if paired else 6) * (groups - 2))
    while True:
      try:
        sched = plan_mod.RegSchedule(program, depth, vec, warps, (), prefetch,
                                     paired=paired, groups=groups)
      except util.SemanticError as e:
        # pairing chosen by default but impossible for this chain (outputs
        # produced at different delays): the unpaired register kernel, not
        # the shared-memory family
        if not paired or options.paired or 'cannot pair' not in str(e):
          raise
        paired = False
        if options.prefetch is None:
          prefetch = 6 * (groups - 2)
        continue
      # keep three blocks per SM resident: shorter boxes if the queues of all
      # inputs do not fit (programs with several inputs or long periods)
      if (options.prefetch is not None or sched.flat_box <= sched.period or
          kernel_reg_mod.Layout(sched).total <= FLAT_SMEM_TARGET):
        break
      prefetch = (sched.flat_box - sched.period) * (groups - 2)
    sched.min_blocks = options.min_blocks or default_min_blocks(sched)
    return sched
  if options.tile:
    if options.tile[0] != 32 * vec:
      raise util.SemanticError('register-streaming tiles are 32 x vec = %d '
                               'cells wide' % (32 * vec))
    rests = [tuple(options.tile[1:])]
  elif program.dim == 3:
    # fused iterations widen the halo: 48-row tiles keep more of each tile
    # (capture r3h: heat3d x4 1072 -> 1109 GCell/s, 768^3 x3 881 -> 934);
    # single sweeps are HBM-bound and lose 1 % to the fewer, larger blocks
    rests = [(32,), (16,), (8,)]
    if depth > 1 and plan_mod.max_elem_size(program) <= 4:
      rests.insert(0, (48,))
  else:
    rests = [(8, 4), (4, 4)]
  prefetches = ([options.prefetch] if options.prefetch is not None
                else [2, 1])
  best, problem = None, None
  for rest in rests:
    rows = math.prod(rest)
    for prefetch in prefetches:
      # two vectors per thread halve the shuffles and barriers per cell, but
      # only while the register histories leave room for the working set
      tall = rows > 32 and not options.tile
      warp_choices = ([options.threads // 32] if options.threads else
                      [rows // 2] if tall else
                      [max(1, min(16, rows // 2)), max(1, min(16, rows))])
      try:
        for warps in warp_choices:
          sched = plan_mod.RegSchedule(program, depth, vec, warps, rest,
                                       prefetch,
                                       min_blocks=options.min_blocks or 1)
          if history_registers(sched) <= REG_HISTORY_3D:
            break
        total = kernel_reg_mod.Layout(sched).total
      except util.SemanticError as e:
        problem = problem or e
        continue
      # 24 warps leave 85 registers a thread: light histories only
      if tall and history_registers(sched) > REG_HISTORY_3D // 2:
        continue
      # one block of 512 threads per SM is the design point whenever the
      # histories are light: then the whole shared memory is the block's
      roomy = (limit if history_registers(sched) > REG_HISTORY_3D // 2
               and not options.tile else max(limit, SMEM_LIMIT - 2048))
      if total > roomy:
        problem = problem or util.SemanticError(
            'depth %d with tile %s needs %d bytes of shared memory (limit '
            '%d)' % (depth, sched.tile, total, limit))
        continue
      score = math.prod(sched.own) / math.prod(sched.tile)
      if best is None or score > best[0] + 1e-9:
        best = (score, sched)
      break
  if best is None:
    raise problem
  sched = best[1]
  if not options.min_blocks:
    # HBM-bound sweeps (one light stage) want two blocks per SM to hide the
    # per-step barrier: cap the registers at 64 when shared memory allows it
    # (heat3d depth 1: 613 -> 745 GCell/s; deeper chains spill and lose)
    smem = kernel_reg_mod.Layout(sched).total + 1024
    if (history_registers(sched) <= 16 and 2 * smem <= SMEM_LIMIT and
        2 * sched.threads <= 2048):
      sched.min_blocks = 2
  return sched


def make_schedule(program, depth, options, limit=SMEM_LIMIT):
  """The schedule for ``depth`` fused iterations: the caller's knobs where
  given, otherwise the first candidate that fits in shared memory."""
  if options.style in (None, 'reg'):
    try:
      return _make_reg_schedule(program, depth, options,
                                min(limit, SMEM_LIMIT // 2))
    except util.SemanticError:
      if options.style == 'reg':
        raise
  vec = options.vec or default_vec(program)
  tiles = [tuple(options.tile)] if options.tile else _default_tiles(program,
                                                                    vec)
  prefetches = ([options.prefetch] if options.prefetch is not None
                else [3, 2, 1])
  problem = None
  for tile in tiles:
    for prefetch in prefetches:
      threads = options.threads or min(256, math.prod(tile) // vec)
      try:
        sched = plan_mod.Schedule(program, depth, tile, vec, threads, prefetch)
        total = kernel_mod.Layout(sched).total
      except util.SemanticError as e:
        problem = problem or e
        continue
      if total <= limit:
        return sched
      problem = problem or util.SemanticError(
          'depth %d with tile %s needs %d bytes of shared memory (limit %d)' %
          (depth, tile, total, limit))
  raise problem


def default_min_blocks(sched):
  """Resident blocks per SM to compile a 2-D register kernel for: enough
  registers for the histories plus working set, as many warps as that
  leaves."""
  need = history_registers(sched) + 64
  blocks = 65536 // (sched.threads * max(need, 32))
  smem = kernel_reg_mod.Layout(sched).total + 1024   # + per-block reserve
  return max(1, min(blocks, 2048 // sched.threads, SMEM_LIMIT // smem, 16))


def history_registers(sched):
  """Registers a thread of a register-streaming kernel keeps across steps."""
  total = 0
  for node in sched.nodes:
    if node.hist_oldest is not None:
      total += (node.hist_oldest - node.delay) * sched.vec * \
          sched.vecs_per_thread * (2 if sched.paired else 1)
  return total


def make_schedules(program, options=None, fast_math=False):
  """The kernel variants to compile: the main temporal depth and, when it does
  not divide ``iterate``, the depth of the remainder.

  ``options.inline``: schedule the program with its single-use locals spliced
  into their readers (plan.inline_single_use; a schedule carries the program
  it was made for, ``sched.program``).  Off unless asked for: bit-identical
  and fewer registers held across steps, but measured no faster where it
  applies (capture r2d: denoise2d 268 vs 270 GCell/s, denoise3d 110 vs 113,
  sobel2d 1272 vs 1410 — the spliced 16-bit locals need an explicit narrowing
  that lazy truncation avoids)."""
  options = options or Options()
  check_supported(program)
  if options.is_default():
    # soda.cuda_tune's winner, if any (``fast_math``: the one of that build)
    found = tuned_mod.lookup(program, fast_math=fast_math)
    if found:
      options = Options(**found)
  spliced = plan_mod.inline_single_use(program) if options.inline else None
  return _make_schedules(spliced or program, options)


def _make_schedules(program, options):
  iterate = program.iterate
  if options.depth:
    main = max(1, min(options.depth, iterate))
  elif not program.feedback or iterate == 1:
    main = 1
  elif options.style in (None, 'reg') and program.dim == 2:
    # as deep as the register file carries comfortably: every fused level
    # keeps (rows of history - 1) x vec values live across steps
    main = 1
    for depth in (2, 4, 8, 16):
      if depth > iterate:
        break
      try:
        sched = make_schedule(program, depth, options)
      except util.SemanticError:
        break     # this depth fits no kernel family: keep the last that did
      if sched.style != 'reg' or history_registers(sched) > REG_HISTORY_BUDGET:
        break
      main = depth
  else:
    main = 1
    for depth in (2, 4, 8, 16):
      if depth > iterate:
        break
      if program.dim > 2 and depth > 2:
        break
      try:
        make_schedule(program, depth, options, SMEM_LIMIT // 2)
      except util.SemanticError:
        break
      main = depth
  depths = [main]
  if iterate % main:
    depths.append(iterate % main)
  return [make_schedule(program, depth, options) for depth in depths]


def print_kernel(program, schedules, kernel_file, fast_math=False):
  p = util.Printer(kernel_file)
  p.println('// CUDA kernels of SODA program `%s` for sm_100a.' %
            program.app_name)
  p.println('// Generated by sodac --cuda-kernel; do not edit.')
  if fast_math:
    p.println('// Build (--cuda-fast, NOT bit-exact): nvcc -gencode '
              'arch=compute_100a,code=sm_100a')
    p.println('//   -prec-div=false -prec-sqrt=false   (and no -fmad=false)')
    p.println('#ifndef SODA_CUDA_FAST_MATH')
    p.println('#define SODA_CUDA_FAST_MATH 1')
    p.println('#endif')
  else:
    p.println('// Build: nvcc -gencode arch=compute_100a,code=sm_100a '
              '-fmad=false')
  p.println('#include "soda_cuda_device.cuh"')
  p.println('#include "soda_cuda_runtime.h"')
  p.println()
  layouts = [(kernel_reg_mod if sched.style == 'reg' else kernel_mod)
             .emit_kernel(p, sched) for sched in schedules]
  host_mod.emit_variant_table(p, program.app_name, schedules, layouts)


def print_header(program, header_file):
  """``<app>.h``: what reference header.print_code emits (header.py:7-64),
  reduced to the buffer_t definition and the prototype."""
  p = util.Printer(header_file)
  guard = 'HALIDE_%s_H_' % program.app_name.upper()
  names = ([n for n, _ in program.inputs] + [n for n, _ in program.outputs] +
           list(program.params))       # reference header.py:57
  p.printlns('#ifndef %s' % guard, '#define %s' % guard, '',
             '#include "soda_cuda.h"  // buffer_t', '',
             '#ifndef HALIDE_FUNCTION_ATTRS', '#define HALIDE_FUNCTION_ATTRS',
             '#endif//HALIDE_FUNCTION_ATTRS', '')
  p.println('int %s(%sconst char* xclbin) HALIDE_FUNCTION_ATTRS;' % (
      program.app_name, ''.join('buffer_t *var_%s_buffer, ' % n
                                for n in names)))
  p.printlns('', '#endif//%s' % guard)


def _emit(path, writer):
  """Write through a temp file, then to ``path`` or stdout for ``-``
  (same pattern as reference opencl.py:99-106)."""
  with tempfile.TemporaryFile(mode='w+') as tmp:
    writer(tmp)
    tmp.seek(0)
    if path == '-':
      shutil.copyfileobj(tmp, sys.stdout)
    else:
      with open(path, 'w') as out:
        shutil.copyfileobj(tmp, out)


def print_code(stencil, args):
  kernel_file = getattr(args, 'cuda_kernel_file', None)
  host_file = getattr(args, 'cuda_host_file', None)
  header_file = getattr(args, 'cuda_header_file', None)
  if kernel_file is None and host_file is None and header_file is None:
    return
  program = plan_mod.extract_program(stencil)
  check_supported(program)
  dims = getattr(args, 'cuda_autotune', None)
  if dims:
    if len(dims) != program.dim:
      raise util.SemanticError('--cuda-autotune takes %d extents' %
                               program.dim)
    from soda import cuda_tune     # needs nvcc and a GPU
    results = cuda_tune.tune(stencil, tuple(dims))
    if results and results[0][1]:
      cuda_tune.record(program, dims, results[0][0], results[0][1])
  if kernel_file is not None:
    fast = bool(getattr(args, 'cuda_fast', False))
    schedules = make_schedules(program, Options.from_args(args), fast)
    _emit(kernel_file, lambda f: print_kernel(program, schedules, f, fast))
  if host_file is not None:
    _emit(host_file, lambda f: host_mod.print_code(program, f))
  if header_file is not None:
    _emit(header_file, lambda f: print_header(program, f))
