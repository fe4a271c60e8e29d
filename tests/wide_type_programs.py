"""Programs on 32/64-bit integers and doubles, and mixed float widths."""
from soda import core

HEADER = 'kernel: %s\nburst width: 64\nunroll factor: 1\niterate: %d\n'

PROGRAMS = {
    'dbl2d': (1, '''input double: a(32, *)
local double: s(0, 0) = a(0, 0) * 0.25 + a(1, 0) * 0.5 - a(0, 1) / 3.0
output double: o(0, 0) = s(0, 0) + s(-1, 0) * s(0, -1) + sqrt(a(0, 0))
'''),
    'dbl3d': (2, '''input double: t0(16, 8, *)
output double: t1(0, 0, 0) = (t0(1, 0, 0) + t0(-1, 0, 0) + t0(0, 1, 0) + t0(0, -1, 0) + t0(0, 0, 1) + t0(0, 0, -1)) * 0.125 + t0(0, 0, 0) * 0.25
'''),
    'i64': (1, '''input int64: a(32, *)
local int64: s(0, 0) = a(0, 0) * 3 + a(1, 0) - a(0, 1) / 7
output int64: o(0, 0) = s(0, 0) - s(-1, 0) * 5 + (s(0, -1) % 11)
'''),
    'u32vol': (1, '''input uint32: a(16, 8, *)
local uint32: s(0, 0, 0) = a(0, 0, 0) * 2654435761 + a(0, 1, 0) - a(0, 0, 1)
output uint32: o(0, 0, 0) = s(0, 0, 0) + s(1, 0, 0) * s(0, -1, 0) - s(0, 0, -1) / 5
'''),
    'i64vol': (2, '''input int64: a(16, 8, *)
output int64: o(0, 0, 0) = a(0, 0, 0) * 3 + a(1, 0, 0) - a(0, -1, 0) + a(0, 0, 1)
'''),
    'mixed': (1, '''input float: a(32, *)
local double: d(0, 0) = a(0, 0) * 0.1 + a(1, 0)
output float: o(0, 0) = d(0, 0) + d(0, -1) * a(-1, 0)
'''),
}

# program, dims, backend options
CASES = [
    ('dbl2d', (1061, 97), {}), ('dbl2d', (2048, 160), {}),
    ('dbl3d', (131, 35, 29), {}), ('dbl3d', (256, 48, 40), {'depth': 2}),
    ('i64', (999, 64), {}), ('i64', (2048, 128), {}),
    ('u32vol', (256, 48, 40), {}), ('u32vol', (131, 35, 29), {}),
    ('mixed', (2048, 96), {}), ('mixed', (777, 131), {}),
    ('dbl2d', (2048, 160), {'style': 'ring'}),
    ('dbl2d', (1061, 97), {'style': 'ring'}),
    ('dbl3d', (131, 35, 29), {'style': 'ring'}),
    # 48-row tiles in 24 warps on two-cell vectors, rows not 16-byte aligned
    # (cell-by-cell stores): the configuration ptxas 12.9 miscompiled before
    # the emitter's slow-path test became one compare
    ('dbl3d', (131, 35, 29), {'tile': [64, 48], 'threads': 768}),
    ('dbl3d', (130, 50, 29), {'tile': [64, 48], 'threads': 768}),
    ('i64vol', (131, 35, 29), {}),
    ('i64vol', (131, 35, 29), {'tile': [64, 48], 'threads': 768}),
]


def stencil_of(name):
  iterate, body = PROGRAMS[name]
  return core.Stencil.from_text(HEADER % (name, iterate) + body)
