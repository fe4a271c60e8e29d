"""Four-dimensional SODA programs (shared by the CPU and GPU tests and the
prebuild tool).  The reference's benchmarks stop at three dimensions; its
grammar and golden loop do not (`input T: a(t0, t1, t2, *)`)."""
from soda import core

HEADER = 'kernel: %s\nburst width: 64\nunroll factor: 1\niterate: %d\n'
PROGRAMS = {
    'heat4d': (HEADER % ('heat4d', 2) + 'input float: a(16, 8, 4, *)\n'
               'output float: o(0, 0, 0, 0) = (a(0, 0, 0, 0) + a(1, 0, 0, 0) + '
               'a(-1, 0, 0, 0) + a(0, 1, 0, 0) + a(0, -1, 0, 0) + '
               'a(0, 0, 1, 0) + a(0, 0, -1, 0) + a(0, 0, 0, 1) + '
               'a(0, 0, 0, -1)) * 0.111f\n', (140, 21, 10, 9)),
    # two stages, one-sided windows, 16-bit integers
    'box4d': (HEADER % ('box4d', 1) + 'input uint16: a(16, 8, 4, *)\n'
              'local uint16: s(0, 0, 0, 0) = a(0, 0, 0, 0) + a(1, 0, 0, 0) + '
              'a(0, 1, 0, 0)\n'
              'output uint16: o(0, 0, 0, 0) = (s(0, 0, 0, 0) + s(0, 0, 1, 0) + '
              's(0, 0, 0, 1) + s(0, 0, 0, 2)) / 3\n', (256, 16, 12, 11)),
}
CASES = [(name, dims) for name, (_, dims) in sorted(PROGRAMS.items())]


def stencil_of(name):
  return core.Stencil.from_text(PROGRAMS[name][0])
