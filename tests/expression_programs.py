"""Seeded random programs with rich expressions: every binary operator level
of the grammar (reference src/haoda/ir/__init__.py:30-52), unary operators,
casts, calls, let bindings, literal forms.  Many are not valid C++ once
lowered (`~` of a float, `%` of doubles): the golden loop would not compile
either; tests only require that this frontend lowers them exactly like the
reference and that nvcc and g++ agree on which ones compile.
"""
import random
BIN = [['||'], ['&&'], ['|'], ['^'], ['&'], ['==', '!='], ['<=', '>=', '<', '>'], ['+', '-'], ['*', '/', '%']]
UN = ['-', '+', '~', '!']
CALLS1 = ['sqrt', 'fabs', 'exp', 'log', 'floor', 'ceil', 'cos', 'abs']
CALLS2 = ['fmax', 'fmin', 'pow', 'max', 'min', 'atan2']
TYPES = ['float', 'double', 'int16', 'uint8', 'int32', 'uint16', 'float32']
LITS = ['1', '0', '3', '0x1F', '7U', '2.f', '.5f', '0.25', '1e3', '1.5e2f', '017', '42L', '0b101', '3.0']

def ref(rng, names, dim):
  return '%s(%s)' % (rng.choice(names), ', '.join(str(rng.randint(-2, 2)) for _ in range(dim)))

def operand(rng, names, dim, depth, lets):
  r = rng.random()
  if depth <= 0 or r < 0.35:
    return ref(rng, names, dim)
  if r < 0.5:
    return rng.choice(LITS)
  if r < 0.6 and lets:
    return rng.choice(lets)
  if r < 0.7:
    return '%s(%s)' % (rng.choice(TYPES), expr(rng, names, dim, depth - 1, lets))
  if r < 0.8:
    if rng.random() < 0.6:
      return '%s(%s)' % (rng.choice(CALLS1), expr(rng, names, dim, depth - 1, lets))
    return '%s(%s, %s)' % (rng.choice(CALLS2), expr(rng, names, dim, depth - 1, lets), expr(rng, names, dim, depth - 1, lets))
  return '(%s)' % expr(rng, names, dim, depth - 1, lets)

def expr(rng, names, dim, depth, lets, level=0):
  if level == len(BIN):
    text = operand(rng, names, dim, depth, lets)
    for _ in range(rng.choice([0, 0, 0, 1, 2])):
      text = rng.choice(UN) + text
    return text
  n = rng.choice([1, 1, 1, 2, 3]) if depth > 0 else 1
  parts = [expr(rng, names, dim, depth - (1 if n > 1 else 0), lets, level + 1) for _ in range(n)]
  text = parts[0]
  for p in parts[1:]:
    text += ' %s %s' % (rng.choice(BIN[level]), p)
  return text

def program(seed):
  rng = random.Random(seed)
  dim = rng.choice([2, 3])
  kind = rng.choice(['float', 'int16', 'double', 'uint8'])
  lines = ['kernel: e%d' % seed, 'burst width: 64', 'unroll factor: 1', 'iterate: 1',
           'input %s: a(%s*)' % (kind, '8, ' * (dim - 1))]
  names = ['a']
  zero = ', '.join(['0'] * dim)
  for k in range(rng.randint(1, 3)):
    last = k == 2 or rng.random() < 0.4
    target = 'out' if last else 'l%d' % k
    lets = []
    let_text = ''
    for j in range(rng.choice([0, 0, 1, 2])):
      name = 'v%d' % j
      typ = rng.choice(['', 'float ', 'int32 '])
      let_text += '%s%s = %s ' % (typ, name, expr(rng, names, dim, 2, lets))
      lets.append(name)
    body = expr(rng, names, dim, 3, lets) + ' + %s(%s)' % (names[-1], zero)
    lines.append('%s %s: %s%s(%s) = %s' % ('output' if last else 'local', kind if last else rng.choice(TYPES[:6]), let_text, target, zero, body))
    names.append(target)
    if last: break
  if not lines[-1].startswith('output'):
    lines.append('output %s: out(%s) = %s(%s)' % (kind, zero, names[-1], zero))
  return '\n'.join(lines) + '\n'
