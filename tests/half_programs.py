"""SODA programs on `half` tensors (shared by the CPU and GPU tests and the
prebuild tool).  `half` is in the DSL's type grammar (reference
src/haoda/ir/__init__.py:24) but is an HLS type: the reference's golden loop
cannot be compiled for it, so parity is unpinned and the semantics are this
repository's: binary16 storage, float evaluation, one rounding to
nearest even on store."""
from soda import core

HEADER = 'kernel: %s\nburst width: 64\nunroll factor: 1\niterate: %d\n'

PROGRAMS = {
    # half in, float local, half out; an explicit half(...) cast; int divisor
    'halfblur': (HEADER % ('halfblur', 2) +
                 'input half: a(32, *)\n'
                 'local float: s(0, 0) = a(0, 0) + a(1, 0) * 0.5f + a(0, 1) * '
                 '0.25f\n'
                 'output half: o(0, 0) = (s(0, 0) + s(-1, 0) + s(0, -1)) / 3 + '
                 'half(a(0, 0) * 1.0009765625f)\n', (1024, 120), {}),
    # half local between float tensors: the local rounds to binary16
    'halfmid': (HEADER % ('halfmid', 1) +
                'input float: a(32, *)\n'
                'local half: m(0, 0) = a(0, 0) * 0.3f + a(0, 1) * 0.7f\n'
                'output float: o(0, 0) = m(0, 0) + m(1, 0) + m(-1, 0) * '
                '0.125f\n', (517, 90), {}),
    # double arithmetic stored to half: ONE rounding, from double
    'halfdbl': (HEADER % ('halfdbl', 1) +
                'input double: a(32, *)\n'
                'output half: o(0, 0) = a(0, 0) * 0.333 + a(1, 0) * 1.0001 + '
                'a(0, 1)\n', (256, 70), {}),
    # 3-D, three iterations fed back through half
    'half3d': (HEADER % ('half3d', 3) +
               'input half: a(16, 8, *)\n'
               'output half: o(0, 0, 0) = (a(0, 0, 0) + a(1, 0, 0) + '
               'a(-1, 0, 0) + a(0, 1, 0) + a(0, -1, 0) + a(0, 0, 1) + '
               'a(0, 0, -1)) * 0.142857f\n', (128, 40, 24), {}),
    # integer tensor stored to half and back
    'halfint': (HEADER % ('halfint', 1) +
                'input int32: a(32, *)\n'
                'local half: h(0, 0) = a(0, 0) + a(1, 0)\n'
                'output int32: o(0, 0) = h(0, 0) * 2 + h(0, 1)\n',
                (512, 64), {}),
}
CASES = [(name, dims, options) for name, (_, dims, options) in
         sorted(PROGRAMS.items())]


def stencil_of(name):
  return core.Stencil.from_text(PROGRAMS[name][0])
