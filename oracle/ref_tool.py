#!/usr/bin/env python3
"""Run the UNMODIFIED reference frontend from /root/reference — TEST TOOLING ONLY.

Nothing under soda-compiler_b200/ imports this; it exists so the oracle and the
product frontend can be pinned against the reference's own output.  It must
run in its own process (``python oracle/ref_tool.py ...``) because the
reference's packages are also called ``soda`` and ``haoda``.

What is patched, and only in this process:
  * ``textx`` and ``cached_property`` resolve to oracle/refshim/ (neither is
    installable offline);
  * ``collections.Iterable`` / ``collections.Mapping`` aliases, removed in
    Python 3.10 and used by the reference (e.g. src/soda/core.py:795).
Everything downstream of parsing — soda.core.Stencil, haoda.ir ``c_expr``,
``host.print_test``, ``header.print_code`` — is the reference's code as it
lies on disk.

Commands
  describe FILE [--iterate N]
      JSON on stdout: stage order, store indices, golden-loop bounds and each
      stage's lowered C expression with every Ref rendered as
      ``name@(relative offset)``.
  harness FILE --outdir DIR [--iterate N]
      Write ``<app>.h`` (header.print_code) and ``<app>_test.cpp``: the
      reference's verbatim ``<app>_test`` golden-loop harness
      (host.print_test, src/soda/codegen/xilinx/host.py:984-1167) preceded only
      by the includes it needs and the ``error_report`` declaration of
      host.py:46.
  host FILE [--iterate N]
      The reference's complete generated OpenCL host file on stdout
      (host.print_code): the source of the FPGA wire-format loops that
      oracle/fpga_layout_ref.py extracts.
"""
import argparse
import collections
import collections.abc
import io
import json
import os
import sys

REFERENCE_SRC = os.environ.get('SODA_REFERENCE_SRC', '/root/reference/src')
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
  if not os.path.isdir(REFERENCE_SRC):
    sys.exit('reference sources not found at %s' % REFERENCE_SRC)
  for name in ('Iterable', 'Mapping', 'Sequence', 'Callable'):
    if not hasattr(collections, name):
      setattr(collections, name, getattr(collections.abc, name))
  sys.path[:0] = [os.path.join(HERE, 'refshim'), REFERENCE_SRC]


def load_stencil(path, iterate=None, tile_size=None, burst_width=None):
  """What src/sodac:80-123 does, without the CLI (``--tile-size`` and
  ``--burst-width`` override the program's, :94-110)."""
  import textx
  from soda import core, grammar
  with open(path) as handle:
    model = textx.metamodel_from_str(
        grammar.GRAMMAR, classes=grammar.CLASSES).model_from_str(handle.read())
  override = list(tile_size or [])
  tile_size = [override[d] if d < len(override) and override[d] > 0
               else model.tile_size[d] for d in range(model.dim - 1)] + [0]
  return core.Stencil(
      burst_width=burst_width or model.burst_width,
      iterate=model.iterate if iterate is None else iterate,
      dram_in=None, dram_out=None, app_name=model.app_name,
      input_stmts=model.input_stmts, param_stmts=model.param_stmts,
      local_stmts=model.local_stmts, output_stmts=model.output_stmts,
      dim=model.dim, tile_size=tile_size, unroll_factor=model.unroll_factor)


def describe(stencil):
  from haoda import ir
  from soda import core
  inputs = tuple(stencil.tensors[name] for name in stencil.input_names)
  stages = []
  for tensor in stencil.chronological_tensors:
    if tensor.is_input():
      continue

    def render(obj, _, tensor=tensor):
      if isinstance(obj, ir.Ref):
        rel = ','.join(str(a - b) for a, b in zip(obj.idx, tensor.st_ref.idx))
        return ir.make_var('%s@(%s)' % (obj.name, rel))
      return obj
    window = core.get_overall_stencil_window(inputs, tensor)
    low = core.get_stencil_window_offset(window)
    extent = core.get_stencil_dim(window)
    stages.append(collections.OrderedDict(
        name=tensor.name, haoda_type=tensor.haoda_type, c_type=tensor.c_type,
        st_idx=list(tensor.st_ref.idx),
        lets=[[let.c_type, let.name, let.expr.visit(render).c_expr]
              for let in tensor.lets],
        c_expr=tensor.expr.visit(render).c_expr,
        parents=list(tensor.parents),
        lo=list(low), hi_margin=[e - l - 1 for e, l in zip(extent, low)]))
  return collections.OrderedDict(
      app_name=stencil.app_name, dim=stencil.dim, iterate=stencil.iterate,
      inputs=[[n, t] for n, t in zip(stencil.input_names,
                                     stencil.input_types)],
      outputs=[[n, t] for n, t in zip(stencil.output_names,
                                      stencil.output_types)],
      tensors=list(stencil.tensors), stages=stages)


def write_harness(stencil, outdir):
  from haoda import util
  from soda.codegen.xilinx import header, host
  os.makedirs(outdir, exist_ok=True)
  app = stencil.app_name
  with open(os.path.join(outdir, app + '.h'), 'w') as handle:
    header.print_code(stencil, handle)
  body = io.StringIO()
  host.print_test(util.Printer(body), stencil)
  with open(os.path.join(outdir, app + '_test.cpp'), 'w') as handle:
    handle.write('// Generated by oracle/ref_tool.py from the unmodified '
                 'reference; do not commit.\n')
    for name in ('cassert', 'cfloat', 'cmath', 'cstdbool', 'cstddef',
                 'cstdint', 'cstdio', 'cstdlib', 'cstring'):
      handle.write('#include <%s>\n' % name)   # host.py:14-17, C headers only
    handle.write('#include "%s.h"\n\n' % app)
    handle.write('FILE* const* error_report = &stderr;\n\n')  # host.py:46
    handle.write(body.getvalue())


def write_host(stencil, out):
  """The reference's complete generated OpenCL host file (host.print_code,
  src/soda/codegen/xilinx/host.py:1169-1204), verbatim."""
  from soda.codegen.xilinx import host
  host.print_code(stencil, out)


def main():
  parser = argparse.ArgumentParser(description=__doc__.split('\n')[0])
  sub = parser.add_subparsers(dest='command', required=True)
  for name in ('describe', 'harness', 'host'):
    cmd = sub.add_parser(name)
    cmd.add_argument('soda_file')
    cmd.add_argument('--iterate', type=int)
    cmd.add_argument('--tile-size', type=int, nargs='+')
    cmd.add_argument('--burst-width', type=int)
    if name == 'harness':
      cmd.add_argument('--outdir', required=True)
  args = parser.parse_args()
  _import_reference()
  stencil = load_stencil(args.soda_file, args.iterate, args.tile_size,
                         args.burst_width)
  if args.command == 'describe':
    json.dump(describe(stencil), sys.stdout, indent=1)
    sys.stdout.write('\n')
  elif args.command == 'host':
    write_host(stencil, sys.stdout)
  else:
    write_harness(stencil, args.outdir)


if __name__ == '__main__':
  main()
