"""The table of tuned kernel configurations (``tuned.json`` next to this file).

Written by ``soda.cuda_tune`` (``sodac --cuda-autotune`` / tools/autotune.py)
after timing candidates on a GPU; read by ``make_schedules`` whenever the
caller gives no options of its own.  An entry is keyed by the program's
signature — the lowered stage expressions, the tensor types and the iteration
count, i.e. everything the generated kernels depend on — so editing a program
silently falls back to the planner's own choice.  SODA_CUDA_TUNED=0 disables
the table.
"""
import hashlib
import json
import os

TABLE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                          'tuned.json')
_cache = {}


def signature(program):
  digest = hashlib.sha256()
  digest.update(('%s|%d|%d|' % (program.app_name, program.dim,
                                program.iterate)).encode())
  for name, haoda_type in program.inputs + program.outputs:
    digest.update(('%s:%s|' % (name, haoda_type)).encode())
  for stage in program.stages:
    lets, expr = stage.render(
        lambda load: '%s%s' % (load.parent, tuple(load.off)))
    digest.update(('%s:%s=%s;%s|' % (stage.name, stage.haoda_type, expr,
                                     lets)).encode())
  return digest.hexdigest()[:16]


def load_table(path=None):
  path = path or TABLE_PATH
  try:
    stamp = os.stat(path).st_mtime_ns
  except OSError:
    return {}
  if _cache.get(path, (None,))[0] != stamp:
    try:
      with open(path) as handle:
        _cache[path] = (stamp, json.load(handle))
    except (OSError, ValueError):
      return {}
  return _cache[path][1]


def lookup(program, path=None, fast_math=False):
  """The tuned ``Options`` keywords of ``program``, or None.  An entry may
  hold ``options_fast`` for the fast-math build, whose kernels have other
  register needs than the exact ones (denoise3d: one vector per thread in 28
  warps exact, two vectors in 16 warps fast)."""
  if os.environ.get('SODA_CUDA_TUNED', '1') == '0':
    return None
  entry = load_table(path).get(signature(program))
  if not entry:
    return None
  if fast_math and 'options_fast' in entry:
    return dict(entry['options_fast'])
  return dict(entry['options'])


def record(program, dims, ms, options, device='', path=None):
  """Store a winner (one entry per program; a later run replaces it)."""
  path = path or TABLE_PATH
  table = dict(load_table(path))
  cells = 1.0
  for n in dims:
    cells *= float(n)
  table[signature(program)] = {
      'app': program.app_name, 'iterate': program.iterate,
      'options': options, 'dims': list(dims), 'ms': round(ms, 4),
      'gcell_per_s': round(cells * program.iterate / ms / 1e6, 1),
      'device': device}
  with open(path, 'w') as handle:
    json.dump(table, handle, indent=1, sort_keys=True)
    handle.write('\n')
  return table
