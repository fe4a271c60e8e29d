"""Slab-partitioned multi-GPU execution of a compiled SODA program.

One process per GPU (``torch.distributed``, NCCL over NVLink).  The grid is cut
into ``world`` contiguous slabs along the streamed (last, never tiled:
reference src/soda/grammar.py:34, README.md:248) dimension.  Every rank keeps
its slab plus ghost planes: ``reach_lo`` below and ``reach_hi`` above, the
streamed-dimension reach of one temporally blocked launch (the window of
``depth`` iterations, reference src/soda/core.py:793-830).  Per launch:

    1. compute the planes next to the slab faces (what the neighbours need),
    2. send them to the neighbours' ghost zones (NCCL send/recv on the
       communication stream) while
    3. the interior of the slab is computed on the main stream,
    4. the next launch waits for the ghosts.

Ranks at the ends of the grid have no neighbour there: the kernel reads
outside the local array as 0, exactly as a single GPU reads outside the grid,
so the sharded result is bit-identical to the single-GPU result (tested).

The reference has no multi-device support; this is the part of
the backend that is new functionality rather than a replacement.
"""
import numpy as np
import torch
import torch.distributed as dist

from soda import cuda as soda_cuda


def partition(rows, world):
  """Balanced contiguous split of ``rows`` planes: [(begin, end)] per rank."""
  return [(rows * r // world, rows * (r + 1) // world) for r in range(world)]


class SlabRunner:
  """Runs ``library`` on this rank's slab of a ``global_dims`` grid.

  Args:
    library: soda.cuda.Library (compiled program).
    global_dims: grid extents, dimension 0 first.
    rank, world: position in the process group (``group`` or the default).
    feedback: {input index: output index} fed back between iterations; default
      by position when #inputs == #outputs (reference core.py:347-351).
      Inputs not fed back (e.g. denoise's ``f``) keep their initial ghosts.
    device: torch device of the local arrays (default: current CUDA device).
    compute: test hook replacing the kernel launch,
      ``compute(depth, inputs, outputs, local_dims, row_begin, row_end,
      valid_lo, valid_hi)``; the product path always launches the CUDA kernel
      and refuses to run without a GPU.
  """

  def __init__(self, library, global_dims, rank, world, feedback=None,
               group=None, device=None, compute=None):
    self.library = library
    self.global_dims = tuple(global_dims)
    self.rank, self.world, self.group = rank, world, group
    self.dim = library.dim
    if compute is None:
      if not torch.cuda.is_available():
        raise RuntimeError('SlabRunner launches CUDA kernels; no GPU found '
                           '(there is no CPU fallback)')
      device = device or torch.device('cuda', torch.cuda.current_device())
    self.device = device or torch.device('cpu')
    self._compute = compute or self._launch
    self.on_gpu = self.device.type == 'cuda'
    n_in, n_out = len(library.inputs), len(library.outputs)
    if feedback is None:
      feedback = {k: k for k in range(n_in)} if n_in == n_out else {}
    self.feedback = dict(feedback)
    rows = self.global_dims[-1]
    if rows < world:
      raise ValueError('fewer streamed planes than ranks')
    self.begin, self.end = partition(rows, world)[rank]
    # ghost depth: the reach of the deepest compiled launch
    self.depths = sorted(library.depths, reverse=True)
    lo, hi = library.window(self.depths[0])
    self.ghost_lo = max(0, -lo[-1]) if rank > 0 else 0
    self.ghost_hi = max(0, hi[-1]) if rank + 1 < world else 0
    self.reach_lo, self.reach_hi = max(0, -lo[-1]), max(0, hi[-1])
    if min(e - b for b, e in partition(rows, world)) < max(self.reach_lo,
                                                           self.reach_hi):
      raise ValueError('slabs are thinner than the halo of one launch')
    self.local_begin = self.begin - self.ghost_lo    # global plane of local 0
    self.local_rows = (self.end + self.ghost_hi) - self.local_begin
    self.local_dims = self.global_dims[:-1] + (self.local_rows,)
    shape = tuple(reversed(self.local_dims))
    dtype = lambda t: torch.from_numpy(
        np.empty(0, soda_cuda.NUMPY_TYPES[t])).dtype
    self.inputs = [torch.zeros(shape, dtype=dtype(t), device=self.device)
                   for _, t in library.inputs]
    # two sets of output-typed arrays to ping-pong between launches
    self.buffers = [[torch.zeros(shape, dtype=dtype(t), device=self.device)
                     for _, t in library.outputs] for _ in range(2)]
    self.current = list(self.inputs)     # what the next launch reads
    self.outputs = None
    if self.on_gpu:
      self.main = torch.cuda.current_stream(self.device)
      self.comm = torch.cuda.Stream(self.device)
    self.launch_count = 0

  # ---- data movement ----------------------------------------------------
  def owned(self, tensor):
    """View of the planes this rank owns."""
    a = self.begin - self.local_begin
    return tensor[a:a + (self.end - self.begin)]

  def load_local(self, owned_inputs):
    """Copy this rank's owned planes in and fill every input's ghosts."""
    for local, given in zip(self.inputs, owned_inputs):
      self.owned(local).copy_(given)
    self.current = list(self.inputs)
    for req in self._exchange(self.inputs):
      req.wait()

  def _exchange(self, tensors):
    """Start sending face planes to the neighbours' ghost zones."""
    ops = []
    a = self.begin - self.local_begin
    b = a + (self.end - self.begin)
    for tensor in tensors:
      if self.rank > 0:
        # my lowest planes are the upper ghosts of rank-1; its highest planes
        # are my lower ghosts
        ops.append(dist.P2POp(dist.isend, tensor[a:a + self.reach_hi],
                              self.rank - 1, self.group))
        ops.append(dist.P2POp(dist.irecv, tensor[a - self.ghost_lo:a],
                              self.rank - 1, self.group))
      if self.rank + 1 < self.world:
        ops.append(dist.P2POp(dist.isend, tensor[b - self.reach_lo:b],
                              self.rank + 1, self.group))
        ops.append(dist.P2POp(dist.irecv, tensor[b:b + self.ghost_hi],
                              self.rank + 1, self.group))
    ops = [op for op in ops if op.tensor.numel() > 0]
    for op in ops:
      # NCCL has no unsigned 16/32/64-bit types: planes travel as bytes
      # (slices of whole planes are contiguous, so the view is free)
      if op.tensor.dtype in (torch.uint16, torch.uint32, torch.uint64):
        op.tensor = op.tensor.view(torch.uint8)
    return dist.batch_isend_irecv(ops) if ops else []

  # ---- compute ----------------------------------------------------------
  def _launch(self, depth, inputs, outputs, local_dims, row_begin, row_end,
              valid_lo, valid_hi):
    self.library.launch(depth, inputs, outputs, local_dims, row_begin,
                        row_end, valid_lo, valid_hi,
                        torch.cuda.current_stream(self.device).cuda_stream)
    self.launch_count += 1

  def plan(self, iterate):
    """Depths of the launches that make up ``iterate`` iterations."""
    depths, left = [], iterate
    while left > 0:
      fits = [d for d in self.depths if d <= left]
      if not fits:
        raise ValueError('no compiled depth fits %d iterations' % left)
      depths.append(fits[0])
      left -= fits[0]
    return depths

  def launches_per_run(self, iterate):
    faces = (1 if self.rank > 0 else 0) + (1 if self.rank + 1 < self.world
                                           else 0)
    return (len(self.plan(iterate)) - 1) * (1 + faces) + 1

  def run(self, iterate=None):
    """All iterations on the current inputs; returns the owned output planes."""
    iterate = self.library.iterate if iterate is None else iterate
    depths = self.plan(iterate)
    if len(depths) > 1 and not self.feedback:
      raise ValueError('iterations need outputs that feed the inputs')
    region = self.library.valid_region(self.global_dims, iterate)
    full_lo, full_hi = [0] * self.dim, list(self.local_dims)
    fin_lo = [lo for lo, _ in region]
    fin_hi = [hi for _, hi in region]
    fin_lo[-1] = min(max(0, fin_lo[-1] - self.local_begin), self.local_rows)
    fin_hi[-1] = min(max(0, fin_hi[-1] - self.local_begin), self.local_rows)
    a = self.begin - self.local_begin
    b = a + (self.end - self.begin)
    current = list(self.inputs)
    pending = []
    for n, depth in enumerate(depths):
      last = n + 1 == len(depths)
      target = self.buffers[n % 2]
      lo, hi = (fin_lo, fin_hi) if last else (full_lo, full_hi)
      for req in pending:       # ghosts of `current` must have landed
        req.wait()
      pending = []

      def go(row_begin, row_end, current=current, target=target, lo=lo,
             hi=hi, depth=depth):
        if row_end > row_begin:
          self._compute(depth, current, target, self.local_dims, row_begin,
                        row_end, lo, hi)
      if last or self.world == 1:
        go(a, b)
      else:
        # faces first, so their transfer overlaps the interior
        low_face = min(b, a + self.reach_hi) if self.rank > 0 else a
        high_face = max(low_face, b - self.reach_lo) \
            if self.rank + 1 < self.world else b
        go(a, low_face)
        go(high_face, b)
        fed = [target[out] for out in self.feedback.values()]
        if self.on_gpu:
          faces_done = self.main.record_event()
          with torch.cuda.stream(self.comm):
            self.comm.wait_event(faces_done)
            pending = self._exchange(fed)
        else:
          pending = self._exchange(fed)
        go(low_face, high_face)
      if not last:
        nxt = list(current)
        for inp, out in self.feedback.items():
          nxt[inp] = target[out]
        current = nxt
    self.outputs = [self.owned(t) for t in target]
    return self.outputs
