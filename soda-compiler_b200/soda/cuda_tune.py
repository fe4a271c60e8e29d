"""Autotuner of the CUDA backend: picks the kernel configuration of a SODA
program by timing candidates on the GPU at hand.

The reference exposes its design-space knobs on the sodac command line
(``--tile-size``, ``--unroll-factor``, ``--burst-width``, reference
src/sodac:35-46, soda/core.py:187-189) and leaves the exploration to the user;
the FPGA knobs have no meaning on a GPU, whose counterparts are the temporal
depth, the block shape, the vector width and the depth of the input queue
(``sodac --cuda-*``).  ``tune`` explores those:

1. ``candidates(program)`` enumerates option sets the planner accepts
   (``codegen.make_schedules`` succeeds and the kernel fits an SM);
2. every candidate is compiled offline (nvcc, in parallel, cached by source
   hash like any other build) — this half needs no GPU;
3. each library runs the whole program on device-resident arrays of the
   requested extents, timed with CUDA events; every candidate's outputs must
   equal the first candidate's bit for bit (they are the same arithmetic in
   the same order) or the candidate is discarded;
4. the winner is recorded in ``codegen/cuda/tuned.json`` under the program's
   signature; ``codegen.make_schedules`` uses the entry whenever the caller
   gives no options of its own (SODA_CUDA_TUNED=0 disables the table).
"""
import concurrent.futures
import contextlib
import itertools
import os

import numpy as np

from haoda import util
from soda import cuda as soda_cuda
from soda.codegen import cuda as codegen
from soda.codegen.cuda import plan as plan_mod
from soda.codegen.cuda import tuned

signature = tuned.signature
record = tuned.record
load_table = tuned.load_table


@contextlib.contextmanager
def untuned():
  """Inside, an empty option set means the planner's own choice, not the
  winner of an earlier tuning run."""
  before = os.environ.get('SODA_CUDA_TUNED')
  os.environ['SODA_CUDA_TUNED'] = '0'
  try:
    yield
  finally:
    if before is None:
      del os.environ['SODA_CUDA_TUNED']
    else:
      os.environ['SODA_CUDA_TUNED'] = before


def _accepts(program, options):
  try:
    for sched in codegen.make_schedules(program, codegen.Options(**options)):
      if codegen.layout_of(sched).total > codegen.SMEM_LIMIT:
        return False
    return True
  except (util.SemanticError, ValueError, ZeroDivisionError):
    return False


def candidates(program, limit=32):
  """Option sets (``codegen.Options`` keyword dicts) worth timing, the
  planner's own choice first."""
  with untuned():
    return _candidates(program, limit)


def _candidates(program, limit):
  # the planner's depth with every block/queue geometry, then the other
  # depths with the planner's geometry (a full product is mostly compile time)
  depths = []
  if program.feedback and program.iterate > 1:
    cap = 16 if program.dim == 2 else 4
    depths = [d for d in (1, 2, 4, 8, 16) if d <= min(program.iterate, cap)]
  grid = []
  # programs with single-use locals: also with those spliced into their
  # readers (fewer registers held across steps; it paid where the registers
  # bought a taller tile: denoise3d, profiles/README.md capture r2o)
  splices = (None, 1) if plan_mod.inline_single_use(program) else (None,)
  if program.dim == 2:
    for threads, groups, prefetch in itertools.product(
        (None, 64, 256), (None, 3, 8), (None, 12, 36)):
      grid.append({'threads': threads, 'groups': groups, 'prefetch': prefetch})
    grid += [{'inline': 1, 'threads': t} for t in (None, 64) if splices[-1]]
  else:
    vec = codegen.default_vec(program)
    rests = [None] + ([[32 * vec, r] for r in (32, 16, 8)]
                      if program.dim == 3 else [])
    for tile, prefetch in itertools.product(rests, (None, 1, 2, 3)):
      grid.append({'tile': tile, 'prefetch': prefetch})
    if program.dim == 3:
      # tall tiles with two vectors per thread
      for inline, prefetch in itertools.product(splices, (1, 2)):
        grid.append({'tile': [32 * vec, 32], 'threads': 512,
                     'prefetch': prefetch, 'inline': inline})
      grid += [{'inline': 1} for _ in splices[1:]]
  grid += [{'depth': depth} for depth in depths]
  chosen, seen = [], set()
  for options in grid:
    options = {k: v for k, v in options.items() if v is not None}
    if not _accepts(program, options):
      continue
    # different option sets may resolve to the same kernels
    key = tuple(s.describe() + str(getattr(s, 'min_blocks', ''))
                for s in codegen.make_schedules(
                    program, codegen.Options(**options)))
    if key in seen:
      continue
    seen.add(key)
    chosen.append(options)
  return chosen[:limit]


def build_all(stencil, option_sets, jobs=8, fast_math=False):
  """Compile every candidate (no GPU needed); ``[(options, path or error)]``."""
  def one(options):
    try:
      return options, soda_cuda.build(stencil, fast_math=fast_math,
                                      options=codegen.Options(**options))
    except Exception as e:   # pylint: disable=broad-except
      return options, e
  with untuned(), concurrent.futures.ThreadPoolExecutor(
      max_workers=jobs) as pool:
    return list(pool.map(one, option_sets))


def _device_arrays(library, dims, seed=7):
  import torch
  gen = torch.Generator(device='cuda')
  gen.manual_seed(seed)
  shape = tuple(reversed(dims))
  signed = {1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}

  def torch_type(haoda_type):
    dtype = np.dtype(soda_cuda.NUMPY_TYPES[haoda_type])
    return (torch.from_numpy(np.empty(0, dtype)).dtype if dtype.kind == 'f'
            else signed[dtype.itemsize])
  inputs = []
  for _, haoda_type in library.inputs:
    dtype = torch_type(haoda_type)
    if dtype.is_floating_point:
      inputs.append(torch.rand(shape, generator=gen, device='cuda',
                               dtype=torch.float32).to(dtype))
    else:
      info = torch.iinfo(dtype)
      inputs.append(torch.randint(info.min, info.max, shape, generator=gen,
                                  device='cuda', dtype=dtype))
  outputs = [torch.empty(shape, dtype=torch_type(t), device='cuda')
             for _, t in library.outputs]
  return inputs, outputs


def time_library(library, dims, reps=5, timer=None):
  """Median milliseconds of a whole run on device arrays, and the outputs."""
  import torch
  inputs, outputs = _device_arrays(library, dims)
  stream = torch.cuda.current_stream().cuda_stream
  for _ in range(2):
    library.run_device(inputs, outputs, dims, 0, stream)
  torch.cuda.synchronize()
  times = []
  for _ in range(reps):
    start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
    start.record()
    library.run_device(inputs, outputs, dims, 0, stream)
    stop.record()
    torch.cuda.synchronize()
    times.append(start.elapsed_time(stop))
  return float(np.median(times)), outputs


def tune(stencil, dims, option_sets=None, reps=5, jobs=8, log=None,
         measure=time_library):
  """Times every candidate on ``dims``; ``[(ms, options)]`` best first.

  ``measure(library, dims, reps) -> (ms, outputs)`` is replaceable for tests.
  """
  program = plan_mod.extract_program(stencil)
  if option_sets is None:
    option_sets = candidates(program)
  results, reference = [], None
  for options, built in build_all(stencil, option_sets, jobs):
    if isinstance(built, Exception):
      if log:
        log('%-60s build failed: %s' % (options, str(built)[:120]))
      continue
    library = soda_cuda.load(built)
    try:
      ms, outputs = measure(library, dims, reps)
    except Exception as e:   # pylint: disable=broad-except
      if log:
        log('%-60s run failed: %s' % (options, e))
      continue
    finally:
      library.release()
    if reference is None:
      reference = [o.clone() if hasattr(o, 'clone') else o for o in outputs]
    elif not all(_same(a, b) for a, b in zip(outputs, reference)):
      if log:
        log('%-60s DISCARDED: outputs differ from the first candidate' %
            options)
      continue
    if log:
      log('%-60s %9.3f ms' % (options, ms))
    results.append((ms, options))
  results.sort(key=lambda item: item[0])
  return results


def _same(a, b):
  if hasattr(a, 'view') and hasattr(a, 'element_size'):    # torch tensors
    import torch
    kind = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[
        a.element_size()]
    return bool(torch.equal(a.view(kind), b.view(kind)))
  return a == b
