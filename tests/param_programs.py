"""SODA programs with `param` statements, shared by the CPU and GPU tests."""
from soda import core

HEADER = 'kernel: %s\nburst width: 64\nunroll factor: 1\niterate: %d\n'

PROGRAMS = {
    # a 3 x 3 convolution with run-time weights and a bias
    'conv3': (1, '''input float: in(32, *)
param float: w[3][3]
param float, dup 2: bias[1]
output float: out(0, 0) = in(-1, -1) * w(0, 0) + in(0, -1) * w(0, 1) + in(1, -1) * w(0, 2) + in(-1, 0) * w(1, 0) + in(0, 0) * w(1, 1) + in(1, 0) * w(1, 2) + in(-1, 1) * w(2, 0) + in(0, 1) * w(2, 1) + in(1, 1) * w(2, 2) + bias(0)
'''),
    # iterated (temporal blocking) with params: a weighted 5-point relaxation
    'relax': (6, '''input float: t0(32, *)
param float: c[2]
output float: t1(0, 0) = t0(0, 0) * c(0) + (t0(1, 0) + t0(-1, 0) + t0(0, 1) + t0(0, -1)) * c(1)
'''),
    # integers, two stages, a 3-D grid, a 3-D param
    'scale16': (1, '''input int16: a(16, 8, *)
param int16: k[2][1][3]
local int16: s(0, 0, 0) = a(0, 0, 0) * k(0, 0, 0) + a(1, 0, 0) * k(0, 0, 1) + a(0, 1, 0) * k(0, 0, 2)
output int16: o(0, 0, 0) = s(0, 0, 0) - s(0, 0, -1) * k(1, 0, 0) + s(0, -1, 0) / k(1, 0, 2)
'''),
}

# program, dims, backend options
CASES = [
    ('conv3', (1061, 97), {}), ('conv3', (2048, 160), {}),
    ('conv3', (2048, 160), {'style': 'ring'}),
    ('relax', (1536, 200), {}), ('relax', (1001, 75), {'depth': 4}),
    ('relax', (2048, 90), {'depth': 1}),
    ('scale16', (256, 48, 40), {}), ('scale16', (131, 35, 29), {}),
]


def stencil_of(name):
  iterate, body = PROGRAMS[name]
  return core.Stencil.from_text(HEADER % (name, iterate) + body)
