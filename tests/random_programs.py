"""Seeded random SODA programs (shared by the CPU compile test, the prebuild
tool and the GPU parity test): 2-D and 3-D, 1-4 iterations, up to four
stages with up to four neighbour reads each, tensors of one of six types with
an occasional stage of another width.

Integer programs only add, subtract and scale by small literals: the golden
loop evaluates them in C++ `int`, where overflow is undefined, and operands
read back from 8/16-bit tensors keep such sums far from it.  32-bit integers
are unsigned (wrap-around is defined).
"""
import random

from soda import core

KINDS = ('float', 'float', 'int16', 'uint8', 'double', 'uint32')
OTHER = {'float': 'double', 'double': 'float', 'int16': 'int32',
         'uint8': 'uint16', 'uint32': 'uint16'}
SEEDS = tuple(range(8))


def program_text(seed):
  rng = random.Random(seed)
  dim = rng.choice((2, 2, 3))
  kind = rng.choice(KINDS)
  floating = kind in ('float', 'double')
  n_local = rng.randint(0, 3)
  names = ['a']
  lines = ['kernel: rnd%d' % seed, 'burst width: 64', 'unroll factor: 1',
           'input %s: a(%s*)' % (kind, ''.join('8, ' for _ in range(dim - 1)))]
  zero = ', '.join('0' for _ in range(dim))
  for k in range(n_local + 1):
    target = 'l%d' % k if k < n_local else 'out'
    text = ''
    for _ in range(rng.randint(1, 3)):
      parent = rng.choice(names)
      off = [rng.randint(-2, 2) for _ in range(dim)]
      ref = '%s(%s)' % (parent, ', '.join(map(str, off)))
      if floating:
        text += ref + rng.choice((' + ', ' - ', ' * '))
      else:
        scale = rng.choice(('', '', ' * 2', ' * 3'))
        text += ref + scale + rng.choice((' + ', ' - '))
    # every window must contain the store point (Program.check_windows)
    text += '%s(%s)' % (names[-1], zero)
    if rng.random() < 0.3:
      text = '(%s)%s' % (text, {'float': ' * 0.25f', 'double': ' / 3.0'}.get(
          kind, ' / 3'))
    stage_type = kind
    if k < n_local and rng.random() < 0.2:
      stage_type = OTHER[kind]
    lines.append('%s %s: %s(%s) = %s' % (
        'local' if k < n_local else 'output', stage_type, target, zero, text))
    names.append(target)
  lines.append('iterate: %d' % rng.randint(1, 4))
  return '\n'.join(lines) + '\n'


def stencil_of(seed):
  return core.Stencil.from_text(program_text(seed))


def dims_of(stencil, seed):
  rng = random.Random(1000 + seed)
  if stencil.dim == 2:
    return (rng.choice((1024, 1061, 2048)), rng.randint(90, 200))
  return (rng.choice((128, 131, 256)), rng.randint(35, 64), rng.randint(30, 48))


# Hand-written programs for shapes the random generator does not produce:
# two inputs and two outputs, the second output reading the first within the
# iteration (under iterate > 1 the Stencil IR calls that read `a_iter1`).
EXTRA = {
    'chain2': ('kernel: chain2\nburst width: 64\nunroll factor: 1\n'
               'input float: a(32, *)\ninput float: b(32, *)\n'
               'output float: o0(0, 0) = a(0, 0) * 0.5f + b(1, 0) * 0.25f\n'
               'output float: o1(0, 0) = a(0, 1) * 0.125f + o0(-1, 0) * 0.5f - '
               'b(0, 0) * 0.25f\niterate: 3\n', (1024, 150)),
    'chain3d': ('kernel: chain3d\nburst width: 64\nunroll factor: 1\n'
                'input float: a(16, 8, *)\ninput float: b(16, 8, *)\n'
                'local float: m(0, 0, 0) = a(0, 1, 0) + b(0, 0, -1)\n'
                'output float: o0(0, 0, 0) = m(0, -1, 0) * 0.5f + a(1, 0, 0) * '
                '0.25f\n'
                'output float: o1(0, 0, 0) = o0(0, 0, 1) * 0.5f + m(0, 0, 0) * '
                '0.25f + b(0, 0, 0) * 0.125f\niterate: 2\n', (128, 40, 36)),
}


def extra_stencil(name):
  return core.Stencil.from_text(EXTRA[name][0])


# --- programs with several outputs --------------------------------------------
# The outputs of one program are defined on different boxes (the reference
# bounds each tensor's golden loop by the window from all inputs to THAT
# tensor, host.py:1082-1091): one-sided reads, an output that reads an earlier
# output of the same iteration, locals shared by some outputs only.
MULTI_SEEDS = tuple(range(100, 160))


def multi_program_text(seed):
  rng = random.Random(seed)
  dim = rng.choice((2, 2, 3))
  n_out = rng.choice((2, 2, 3))
  # iterate > 1 needs as many inputs as outputs (reference core.py:228-243)
  n_in = n_out if rng.random() < 0.6 else rng.randint(1, 2)
  n_local = rng.randint(0, 2)
  inputs = ['i%d' % k for k in range(n_in)]
  lines = ['kernel: multi%d' % seed, 'burst width: 64', 'unroll factor: 1']
  for name in inputs:
    lines.append('input float: %s(%s*)' % (
        name, ''.join('8, ' for _ in range(dim - 1))))
  zero = ', '.join('0' for _ in range(dim))
  names = list(inputs)
  # the reference's dataflow graph needs every input and local consumed
  unread = list(inputs)
  for k in range(n_local + n_out):
    is_local = k < n_local
    target = 'l%d' % k if is_local else 'o%d' % (k - n_local)
    sign = rng.choice((-1, 1, 0))     # one-sided windows are the common case
    terms = []
    last = k + 1 == n_local + n_out
    parents = [rng.choice(names) for _ in range(rng.randint(1, 3))]
    parents += unread if last else unread[:1]
    for parent in parents:
      if parent in unread:
        unread.remove(parent)
      off = [rng.randint(min(0, 2 * sign) if sign else -2,
                         max(0, 2 * sign) if sign else 2) for _ in range(dim)]
      terms.append('%s(%s) * %s' % (parent, ', '.join(map(str, off)),
                                    rng.choice(('0.5f', '0.25f', '0.125f'))))
    # every window must contain the store point (Program.check_windows): a
    # read of an input at the store point itself
    terms.append('%s(%s) * 0.0625f' % (rng.choice(inputs), zero))
    lines.append('%s float: %s(%s) = %s' % (
        'local' if is_local else 'output', target, zero, ' + '.join(terms)))
    names.append(target)
    if is_local:
      unread.append(target)
  iterate = rng.randint(1, 3) if n_in == n_out else 1
  lines.append('iterate: %d' % iterate)
  return '\n'.join(lines) + '\n'


def multi_stencil(seed):
  return core.Stencil.from_text(multi_program_text(seed))


def multi_dims(stencil, seed):
  rng = random.Random(2000 + seed)
  if stencil.dim == 2:
    return (rng.choice((300, 333, 512)), rng.randint(60, 120))
  return (rng.choice((64, 67, 128)), rng.randint(24, 40), rng.randint(20, 30))
