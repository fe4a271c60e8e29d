"""The FPGA wire format of SODA tensors, as a GPU pack / unpack.

The reference's generated OpenCL host does not hand the FPGA kernel dense
arrays.  Every input is rewritten into a **tiled, burst-aligned,
bank-interleaved** buffer (reference src/soda/codegen/xilinx/host.py:629-686)
and every output is gathered back out of one (:823-901):

* the non-streamed dimensions are cut into tiles of ``TILE_SIZE_DIM_d`` cells
  that overlap by the stencil window, i.e. advance by ``TILE_SIZE_DIM_d -
  STENCIL_DIM_d + 1`` (:262-264); the streamed dimension is not tiled;
* a tile is linearised dimension 0 fastest with pitch ``TILE_SIZE_DIM_d``
  (``offset_in_tile``, :650-653) and padded to whole bursts of
  ``BURST_WIDTH / width * num_bank`` elements (``tile_size_linearized``,
  :334-347); tiles follow each other, dimension 0 of the tile index fastest;
* element ``o`` of that stream lives in DRAM bank ``bank_vec[o % num_bank]``
  at position ``o / num_bank`` (:680-684);
* the kernel's output stream lags its input stream by the stencil distance,
  so output cell ``x`` of a tile is found at ``x + stencil_offset`` (:878-893),
  and only the cells whose window fits the tile are valid (:838-852).

``WireLayout`` computes those constants from a ``soda.core.Stencil`` with the
reference's own formulas; ``pack`` / ``unpack`` run the CUDA kernels of
csrc/soda_fpga_layout.cu on device-resident arrays (C ABI:
include/soda_fpga_layout.h), so data in the FPGA flow's wire format — a host
written for an xclbin, a capture of its DMA buffers — can enter and leave the
GPU backend without a CPU repacking pass.  Pure byte movement: HBM-bound.
"""
import ctypes
import hashlib
import os
import shutil
import subprocess

from haoda import util
from soda import core

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(_PKG_ROOT, 'csrc', 'soda_fpga_layout.cu')
INCLUDE_DIR = os.path.join(os.path.dirname(_PKG_ROOT), 'include')
HEADER = os.path.join(INCLUDE_DIR, 'soda_fpga_layout.h')
MAX_DRAM_BANK = 4     # reference src/soda/util.py MAX_DRAM_BANK


class TensorLayout(ctypes.Structure):
  """``soda_fpga_layout_t`` of include/soda_fpga_layout.h."""
  _fields_ = [('dim', ctypes.c_int32), ('elem_size', ctypes.c_int32),
              ('dims', ctypes.c_int32 * 4), ('tile_size', ctypes.c_int32 * 4),
              ('tile_num', ctypes.c_int32 * 4),
              ('tile_step', ctypes.c_int32 * 4),
              ('lo', ctypes.c_int32 * 4),
              ('hi_margin', ctypes.c_int32 * 4),
              ('num_bank', ctypes.c_int32),
              ('bank_vec', ctypes.c_int32 * 4),
              ('tile_size_linearized', ctypes.c_int64),
              ('stream_offset', ctypes.c_int64)]


class WireLayout:
  """Layout constants of one program on one grid.

  Args:
    stencil: soda.core.Stencil (tile sizes, burst width, ``dram`` banks of
      every input and output statement, the stencil windows).
    dims: grid extents, dimension 0 first.
  """

  def __init__(self, stencil, dims):
    self.stencil = stencil
    self.dim = stencil.dim
    self.dims = tuple(dims)
    if len(self.dims) != self.dim:
      raise ValueError('expected %d extents' % self.dim)
    self.tile_size = tuple(stencil.tile_size[:self.dim - 1])
    tensors = stencil.tensors
    inputs = [tensors[name] for name in stencil.input_names]
    first_out = tensors[stencil.output_names[0]]
    window = core.get_overall_stencil_window(inputs, first_out)
    # STENCIL_DIM_d and STENCIL_DISTANCE (host.py:1180-1194)
    self.stencil_dim = tuple(core.get_stencil_dim(window))
    distance = core.get_stencil_distance(window, stencil.tile_size)
    offset = distance - _serialize(core.get_stencil_window_offset(window),
                                   stencil.tile_size)
    self.stencil_distance = max(distance, offset)
    for d in range(self.dim - 1):
      if self.tile_size[d] < self.stencil_dim[d]:
        raise util.SemanticError('tile size %d is smaller than the stencil '
                                 'window in dimension %d' % (
                                     self.tile_size[d], d))
    # tiles per dimension (host.py:262-264)
    self.tile_num = tuple(
        (self.dims[d] - self.stencil_dim[d] + 1 + self.tile_size[d] -
         self.stencil_dim[d]) // (self.tile_size[d] - self.stencil_dim[d] + 1)
        for d in range(self.dim - 1))
    self.stmts = {s.name: s for s in
                  list(stencil.input_stmts) + list(stencil.output_stmts)}
    in0, out0 = stencil.input_stmts[0], stencil.output_stmts[0]
    # host.py:334-347: every size derives from the FIRST input / output
    self.tile_pixel_num = self.dims[-1]
    for extent in self.tile_size:
      self.tile_pixel_num *= extent
    self.tile_burst_num = (self.tile_pixel_num - 1) // self._burst(in0) + 1
    self.tile_size_linearized_i = self.tile_burst_num * self._burst(in0)
    self.tile_size_linearized_o = self.tile_burst_num * self._burst(out0)
    # the valid cells of a tile on the output side (host.py:827-852): window
    # of the FIRST input to the FIRST output
    window0 = core.get_overall_stencil_window(inputs[0], first_out)
    self.window_offset = tuple(core.get_stencil_window_offset(window0))
    self.window_dim = tuple(core.get_stencil_dim(window0))
    # per output: where in the stream its cell x sits (host.py:868-877)
    self.stream_offset = {}
    for name in stencil.output_names:
      w = core.get_overall_stencil_window(inputs, tensors[name])
      self.stream_offset[name] = core.get_stencil_distance(
          w, stencil.tile_size) - _serialize(
              core.get_stencil_window_offset(w), stencil.tile_size)

  def _bits(self, stmt):
    return util.get_width_in_bits(stmt.haoda_type)

  def _burst(self, stmt):
    """Elements per burst across the statement's banks."""
    return self.stencil.burst_width // self._bits(stmt) * len(stmt.dram)

  def banks(self, name):
    return tuple(self.stmts[name].dram)

  def bank_elems(self, name):
    """Elements of each bank buffer of tensor ``name`` (host.py:399-415)."""
    stmt = self.stmts[name]
    linearized = (self.tile_size_linearized_i
                  if name in self.stencil.input_names
                  else self.tile_size_linearized_o)
    tiles = 1
    for n in self.tile_num:
      tiles *= n
    return (tiles * linearized // len(stmt.dram) +
            ((self.stencil_distance - 1) // self._burst(stmt) + 1) *
            (self.stencil.burst_width // self._bits(stmt)))

  def check(self, name):
    """The reference sizes every tile of the stream from the FIRST input's
    burst count (host.py:334-347).  When an output uses fewer banks (or
    wider elements) than that input, its tiles are shorter than their cell
    count, overlap in the stream, and the last one is read beyond the bank
    buffer the same host allocated (host.py:399-415).  Such a configuration
    has no defined content; it is refused instead of read out of bounds."""
    stmt = self.stmts[name]
    is_input = name in self.stencil.input_names
    linearized = (self.tile_size_linearized_i if is_input
                  else self.tile_size_linearized_o)
    tiles = 1
    for n in self.tile_num:
      tiles *= n
    # the highest in-tile offset that is moved
    margin = ([0] * self.dim if is_input else list(self.valid_hi_margin()))
    extents = list(self.tile_size) + [self.dims[-1]]
    reach, pitch = 0, 1
    for d in range(self.dim):
      reach += (extents[d] - margin[d] - 1) * pitch
      pitch *= extents[d]
    last = ((tiles - 1) * linearized + reach +
            (0 if is_input else self.stream_offset[name]))
    if last // len(stmt.dram) >= self.bank_elems(name):
      raise util.SemanticError(
          'wire format of `%s`: stream element %d lies beyond the bank '
          'buffer of %d elements the reference allocates (bank counts of '
          'inputs and outputs differ)' % (name, last, self.bank_elems(name)))

  def descriptor(self, name):
    """The C struct the kernels take for tensor ``name``."""
    self.check(name)
    stmt = self.stmts[name]
    is_input = name in self.stencil.input_names
    out = TensorLayout()
    out.dim = self.dim
    out.elem_size = self._bits(stmt) // 8
    pad = lambda xs, fill: list(xs) + [fill] * (4 - len(xs))
    out.dims[:] = pad(self.dims, 1)
    out.tile_size[:] = pad(self.tile_size, 1)
    out.tile_num[:] = pad(self.tile_num, 1)
    # tiles advance by TILE - STENCIL_DIM + 1 (the program's window); the
    # cells an output tile holds are those whose window (first input ->
    # first output) fits: [lo, extent - hi_margin) per dimension
    out.tile_step[:] = pad([t - s + 1 for t, s in zip(self.tile_size,
                                                      self.stencil_dim)], 1)
    if is_input:
      out.lo[:] = [0, 0, 0, 0]
      out.hi_margin[:] = [0, 0, 0, 0]
    else:
      out.lo[:] = pad(self.window_offset, 0)
      out.hi_margin[:] = pad(self.valid_hi_margin(), 0)
    out.num_bank = len(stmt.dram)
    out.bank_vec[:] = pad(stmt.dram, 0)
    out.tile_size_linearized = (self.tile_size_linearized_i if is_input
                                else self.tile_size_linearized_o)
    out.stream_offset = 0 if is_input else self.stream_offset[name]
    return out

  def valid_hi_margin(self):
    """Cells at the high end of a tile (or of the streamed dimension) that the
    output side skips: ``window_dim - 1 - window_offset`` (host.py:838-852)."""
    return tuple(n - 1 - o for n, o in zip(self.window_dim,
                                           self.window_offset))


def _serialize(vec, tile_size):
  offset, pitch = 0, 1
  for d, v in enumerate(vec):
    offset += v * pitch
    pitch *= tile_size[d]
  return offset


# --- the CUDA side ---------------------------------------------------------------

def lib_path():
  """In-tree, named by the hash of its sources (travels with the snapshot to
  GPU machines; file times do not survive that trip)."""
  digest = hashlib.sha256()
  for path in (CSRC, HEADER):
    with open(path, 'rb') as handle:
      digest.update(handle.read())
  return os.path.join(_PKG_ROOT, '_build', 'fpga_layout-%s' %
                      digest.hexdigest()[:12], 'libsoda_fpga_layout.so')


def build(force=False):
  """Compile csrc/soda_fpga_layout.cu for sm_100a (in-tree, cached)."""
  LIB_PATH = lib_path()
  if os.path.exists(LIB_PATH) and not force:
    return LIB_PATH
  if shutil.which('nvcc') is None:
    raise RuntimeError('nvcc not found: the wire-format kernels are CUDA only')
  os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
  tmp = '%s.%d.tmp' % (LIB_PATH, os.getpid())
  command = ['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3',
             '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-shared',
             '-I', INCLUDE_DIR, CSRC, '-o', tmp]
  done = subprocess.run(command, stdout=subprocess.PIPE,
                        stderr=subprocess.STDOUT, text=True, check=False)
  if done.returncode != 0:
    raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(command),
                                                 done.stdout))
  os.replace(tmp, LIB_PATH)
  return LIB_PATH


_lib = None


def library():
  global _lib
  if _lib is None:
    _lib = ctypes.CDLL(build())
    for name in ('soda_fpga_pack', 'soda_fpga_unpack'):
      fn = getattr(_lib, name)
      fn.restype = ctypes.c_int
      fn.argtypes = [ctypes.POINTER(TensorLayout), ctypes.c_void_p,
                     ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]
  return _lib


def _bank_pointers(layout, name, banks):
  """banks: {bank id: device tensor or pointer}; every bank the statement
  uses must be there."""
  values = [0] * MAX_DRAM_BANK
  for bank in layout.banks(name):
    item = banks[bank]
    values[bank] = item.data_ptr() if hasattr(item, 'data_ptr') else int(item)
  return (ctypes.c_void_p * MAX_DRAM_BANK)(*values)


def pack(layout, name, dense, banks, stream=None):
  """Dense device array of input ``name`` -> its bank buffers (asynchronous).
  Elements of the buffers that belong to no cell (burst padding, the cut-off
  part of the last tile) are left as they are, like the reference does."""
  desc = layout.descriptor(name)
  code = library().soda_fpga_pack(
      ctypes.byref(desc),
      dense.data_ptr() if hasattr(dense, 'data_ptr') else int(dense),
      _bank_pointers(layout, name, banks), stream)
  if code:
    raise RuntimeError('soda_fpga_pack failed: %d' % code)


def unpack(layout, name, dense, banks, stream=None):
  """Bank buffers of output ``name`` -> the valid cells of its dense device
  array (asynchronous); other cells are left as they are."""
  desc = layout.descriptor(name)
  code = library().soda_fpga_unpack(
      ctypes.byref(desc),
      dense.data_ptr() if hasattr(dense, 'data_ptr') else int(dense),
      _bank_pointers(layout, name, banks), stream)
  if code:
    raise RuntimeError('soda_fpga_unpack failed: %d' % code)
