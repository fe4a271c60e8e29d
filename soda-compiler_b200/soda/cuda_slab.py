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

Two ways to move the face planes.

``p2p`` (default on GPUs)  The interior launch is sized to fill every SM for
    its whole duration, so an NCCL send/recv *kernel* queued next to it only
    gets an SM when the interior drains: the exchange ends up serialised
    after the compute it was meant to hide behind (measured: 2 GPUs at 86 %
    of 2 x one GPU).  Here the neighbours' arrays are mapped into this
    process (CUDA IPC) and the face planes are written straight into the
    neighbour's ghost rows by the **copy engine** (peer-to-peer over
    NVLink), followed by a stream-ordered 32-bit flag
    (cuStreamWriteValue32); the neighbour's next launch waits on the flag
    (cuStreamWaitValue32).  Nothing on this path needs an SM.  Output arrays
    rotate through three buffers, which is what makes writing into a
    neighbour's ghost rows safe without a credit message: the buffer written
    in launch n was last read in launch n - 2, and a rank cannot be two
    launches ahead of the neighbour it waits on.  Across ``run()`` calls a
    second flag says "previous run finished".
``collective``  torch.distributed send/recv (NCCL on GPUs, gloo in the CPU
    tests), ping-pong buffers.  Set SODA_CUDA_SLAB_EXCHANGE=collective.

The reference has no multi-device support; this is the part of
the backend that is new functionality rather than a replacement.
"""
import os

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
      valid_lo, valid_hi)`` (one box per output); the product path always
      launches the CUDA kernel
      and refuses to run without a GPU.
  """

  def __init__(self, library, global_dims, rank, world, feedback=None,
               group=None, device=None, compute=None, exchange=None):
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
    if exchange is None:
      exchange = os.environ.get('SODA_CUDA_SLAB_EXCHANGE') or (
          'p2p' if self.on_gpu and compute is None and world > 1
          else 'collective')
    if exchange not in ('p2p', 'collective'):
      raise ValueError('exchange must be `p2p` or `collective`')
    self.exchange = exchange
    # sets of output-typed arrays the launches rotate through
    self.buffers = [[torch.zeros(shape, dtype=dtype(t), device=self.device)
                     for _, t in library.outputs]
                    for _ in range(3 if exchange == 'p2p' else 2)]
    self.current = list(self.inputs)     # what the next launch reads
    self.outputs = None
    if self.on_gpu:
      self.main = torch.cuda.current_stream(self.device)
      self.comm = torch.cuda.Stream(self.device)
      # the face launches: high priority, so their blocks are dispatched
      # before those of the interior launch queued next to them
      self.face_streams = [torch.cuda.Stream(self.device, priority=-1)
                           for _ in range(2)]
    self.launch_count = 0
    self._chunk_cache = {}
    if exchange == 'p2p':
      self._setup_p2p()

  # ---- peer-to-peer exchange ------------------------------------------------
  def _setup_p2p(self):
    """Map the neighbours' output arrays and flags into this process: CUDA
    IPC handles opened in MY context on MY device (soda_cuda_ipc_open).  The
    process never touches the neighbour's GPU through a context of its own —
    a second context there time-slices with the neighbour's kernels (measured
    with torch's tensor sharing, which does create one: 2.3x slower)."""
    lib = self.library
    self.fed_outputs = sorted(set(self.feedback.values()))
    # flags: 0 ghosts from below landed, 1 from above, 2/3 the neighbour
    # below/above finished its previous run, 7 scratch
    self.flags = torch.zeros(8, dtype=torch.int32, device=self.device)
    torch.cuda.synchronize(self.device)
    mine = {'flags': lib.ipc_export(self.flags.data_ptr()),
            'buffers': [[lib.ipc_export(bufs[k].data_ptr())
                         for k in self.fed_outputs] for bufs in self.buffers]}
    everyone = [None] * self.world
    dist.all_gather_object(everyone, mine, group=self.group)
    self.peers = {}
    opened = {}

    def address(exported):
      handle, offset = exported
      if handle not in opened:       # an allocation is opened once
        opened[handle] = lib.ipc_open(handle)
      return opened[handle] + offset
    rows = self.global_dims[-1]
    self.plane_bytes = [self.buffers[0][k][0].numel() *
                        self.buffers[0][k].element_size()
                        for k in self.fed_outputs]
    for side, q in ((0, self.rank - 1), (1, self.rank + 1)):
      if not 0 <= q < self.world:
        continue
      begin, end = partition(rows, self.world)[q]
      first = self.reach_lo if q > 0 else 0       # its local index of `begin`
      self.peers[side] = {
          'flags': address(everyone[q]['flags']),
          'buffers': [[address(t) for t in bufs]
                      for bufs in everyone[q]['buffers']],
          # its ghost rows that mirror my face: above its slab if it is below
          # me, below its slab if it is above me
          'ghost': ((first + end - begin, first + end - begin + self.reach_hi)
                    if side == 0 else (first - self.reach_lo, first)),
      }
    # a flag in a neighbour's memory is written by cuStreamWriteValue32 where
    # the driver accepts a peer-mapped address, else by a 4-byte peer copy of
    # a locally written value (both stream-ordered, neither needs an SM)
    self.stage = torch.zeros(2, dtype=torch.int32, device=self.device)
    self.flag_by_copy = False
    try:
      for peer in self.peers.values():
        lib.flag_write(peer['flags'] + 4 * 7, 1, self.comm.cuda_stream)
      self.comm.synchronize()
    except Exception:   # pylint: disable=broad-except
      self.flag_by_copy = True
    dist.barrier(group=self.group)
    self.tick = 0          # exchanges so far (same on every rank)
    self.run_index = 0
    self.push_done = None

  def _raise_flag(self, side, index, value, stream):
    """Stream-ordered ``peer.flags[index] = value``."""
    lib, target = self.library, self.peers[side]['flags'] + 4 * index
    if not self.flag_by_copy:
      lib.flag_write(target, value, stream.cuda_stream)
      return
    source = self.stage.data_ptr() + 4 * side
    lib.flag_write(source, value, stream.cuda_stream)
    lib.copy_async(target, source, 4, stream.cuda_stream)

  def _push_faces(self, target, bset, first_pass):
    """On the communication stream: copy my face planes into the neighbours'
    ghost rows of the same buffer set, then raise their `landed` flag."""
    lib = self.library
    a = self.begin - self.local_begin
    b = a + (self.end - self.begin)
    stream = self.comm.cuda_stream
    self.tick += 1
    for side, peer in self.peers.items():
      if first_pass and self.run_index:
        # the neighbour may still be reading this buffer set in its last run
        lib.flag_wait_geq(self.flags.data_ptr() + 4 * (2 + side),
                          self.run_index, stream)
      face = (a, a + self.reach_hi) if side == 0 else (b - self.reach_lo, b)
      g0, g1 = peer['ghost']
      if g1 > g0:
        for slot, out in enumerate(self.fed_outputs):
          plane = self.plane_bytes[slot]
          lib.copy_async(peer['buffers'][bset][slot] + g0 * plane,
                         target[out].data_ptr() + face[0] * plane,
                         (g1 - g0) * plane, stream)
      # I am `above` my lower neighbour (its flag 1) and `below` my upper one
      self._raise_flag(side, 1 - side, self.tick, self.comm)
    self.push_done = self.comm.record_event()

  def _await_ghosts(self):
    """The next launches on the main stream need the ghosts of exchange
    number ``self.tick``."""
    stream = self.main.cuda_stream
    for side in self.peers:
      self.library.flag_wait_geq(self.flags.data_ptr() + 4 * side, self.tick,
                                 stream)
    if self.push_done is not None:
      self.main.wait_event(self.push_done)   # my faces left before I rewrite

  def _finish_run(self):
    """Tell the neighbours this run no longer reads any buffer."""
    self.run_index += 1
    for side in self.peers:
      self._raise_flag(side, 3 - side, self.run_index, self.main)

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
              valid_lo, valid_hi, chunk_rows=0):
    self.library.launch(depth, inputs, outputs, local_dims, row_begin,
                        row_end, valid_lo, valid_hi,
                        torch.cuda.current_stream(self.device).cuda_stream,
                        chunk_rows)
    self.launch_count += 1

  def _split(self, depth, a, b):
    """``(low_face, high_face, chunk_rows)``: [a, low_face) and
    [high_face, b) are computed first and sent to the neighbours.

    A block pays a lead-in of ``lead_rows`` rows besides the rows it owns, so
    a face of just the ``reach`` rows a neighbour needs streams most of its
    rows twice.  Where that is a visible share of the slab (3-D programs:
    heat3d on 4 GPUs, 2 x 6 extra plane-steps per launch on 256 planes) the
    faces are instead whole blocks of the decomposition the library picks for
    the slab (``chunk_rows`` rows each): cut at block boundaries, the three
    launches together are exactly the blocks of the one-launch decomposition.
    Where it is not (2-D programs: 25 extra rows on 16384), minimal faces
    measured 2 % faster (profiles/README.md, capture r1p).
    SODA_CUDA_SLAB_FACES=minimal|chunk overrides the choice."""
    need_lo = self.reach_hi if self.rank > 0 else 0         # rows from a up
    need_hi = self.reach_lo if self.rank + 1 < self.world else 0
    chunk = 0
    if self.on_gpu and self._compute == self._launch:
      mode = os.environ.get('SODA_CUDA_SLAB_FACES')
      if mode not in ('minimal', 'chunk'):
        faces = (1 if need_lo else 0) + (1 if need_hi else 0)
        extra = faces * self.library.lead_rows(depth)
        mode = 'chunk' if extra > 0.01 * (b - a) else 'minimal'
      key = (depth, b - a)
      if mode == 'chunk' and key not in self._chunk_cache:
        self._chunk_cache[key] = self.library.chunk_rows(
            depth, self.local_dims, b - a)
      chunk = self._chunk_cache[key] if mode == 'chunk' else 0
    if chunk:
      round_up = lambda n: -(-n // chunk) * chunk
      # Thin slab, coarse blocks (1024^3 on 8 GPUs: 128 planes in 2 blocks of
      # 64): whole-block faces would leave little or no interior to hide the
      # exchange behind, and the ranks at the ends of the grid, with one face,
      # would run two 64-plane blocks back to back.  No block longer than a
      # quarter of the slab: faces of one block each, at least half the slab
      # as interior (heat3d, 128 planes per rank: 4.69 -> 4.51 ms per 32
      # iterations, jacobi3d 4.19 -> 4.01; 3, 5, 6 or 8 blocks are slower:
      # captures r3a-r3d).
      chunk = min(chunk, -(-(b - a) // 4))
      if (chunk >= max(need_lo, need_hi, self.library.lead_rows(depth), 1) and
          round_up(need_lo) + round_up(need_hi) < b - a):
        need_lo, need_hi = round_up(need_lo), round_up(need_hi)
      else:
        chunk = 0       # thin slab: minimal faces, chunks chosen per launch
    low_face = min(b, a + need_lo)
    high_face = max(low_face, b - need_hi)
    return low_face, high_face, chunk

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

  def valid_regions(self, iterate):
    """Per output, ``[(lo, hi)]`` per dimension of the global grid after
    ``iterate`` iterations: each output is defined on the box of its own
    window (reference host.py:1082-1091).  More iterations than the program
    was compiled for means repeated application with the fed-back tensors
    (``u <- output``; the reference refuses to iterate programs whose inputs
    and outputs differ, src/soda/core.py:228-233, so its harness would be
    called once per application): an application reads arrays that are
    defined where every output of the previous one is, and shrinks that box
    by each output's own window, as the golden loop bounds do."""
    n_out = len(self.library.outputs)
    if iterate <= self.library.iterate:
      return self.library.valid_regions(self.global_dims, iterate)
    common_lo, common_hi = [0] * self.dim, [0] * self.dim
    margins = []
    for depth in self.plan(iterate):
      margins = []
      for k in range(n_out):
        step_lo, step_hi = self.library.window_of(k, depth)
        margins.append(([a + max(0, -b) for a, b in zip(common_lo, step_lo)],
                        [a + max(0, b) for a, b in zip(common_hi, step_hi)]))
      common_lo = [max(m[0][d] for m in margins) for d in range(self.dim)]
      common_hi = [max(m[1][d] for m in margins) for d in range(self.dim)]
    return [[(l, max(l, n - h)) for l, h, n in zip(lo, hi, self.global_dims)]
            for lo, hi in margins]

  def valid_region(self, iterate, output=0):
    return self.valid_regions(iterate)[output]

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
    n_out = len(self.library.outputs)
    full_lo = [[0] * self.dim for _ in range(n_out)]
    full_hi = [list(self.local_dims) for _ in range(n_out)]
    fin_lo, fin_hi = [], []      # one box per output, in local coordinates
    for region in self.valid_regions(iterate):
      lo = [a for a, _ in region]
      hi = [b for _, b in region]
      lo[-1] = min(max(0, lo[-1] - self.local_begin), self.local_rows)
      hi[-1] = min(max(0, hi[-1] - self.local_begin), self.local_rows)
      fin_lo.append(lo)
      fin_hi.append(hi)
    a = self.begin - self.local_begin
    b = a + (self.end - self.begin)
    current = list(self.inputs)
    pending = []
    p2p = self.exchange == 'p2p'
    for n, depth in enumerate(depths):
      last = n + 1 == len(depths)
      target = self.buffers[n % len(self.buffers)]
      lo, hi = (fin_lo, fin_hi) if last else (full_lo, full_hi)
      for req in pending:       # ghosts of `current` must have landed
        req.wait()
      pending = []
      if p2p and n:
        self._await_ghosts()

      def go(row_begin, row_end, chunk=0, current=current, target=target,
             lo=lo, hi=hi, depth=depth):
        if row_end > row_begin:
          if chunk:
            self._compute(depth, current, target, self.local_dims, row_begin,
                          row_end, lo, hi, chunk)
          else:
            self._compute(depth, current, target, self.local_dims, row_begin,
                          row_end, lo, hi)
      if last or self.world == 1:
        go(a, b)
      else:
        # faces first, so their transfer overlaps the interior
        low_face, high_face, chunk = self._split(depth, a, b)
        faces = [(a, low_face), (high_face, b)]
        face_events = []
        if self.on_gpu:
          # each face on its own high-priority stream, the interior next to
          # them on the main stream: the faces' blocks go first, the interior
          # fills the SMs as they drain
          begun = self.main.record_event()
          for stream, (r0, r1) in zip(self.face_streams, faces):
            if r1 > r0:
              with torch.cuda.stream(stream):
                stream.wait_event(begun)
                go(r0, r1, chunk)
                face_events.append(stream.record_event())
        else:
          for r0, r1 in faces:
            go(r0, r1)
        fed = [target[out] for out in self.feedback.values()]
        if p2p:
          with torch.cuda.stream(self.comm):
            for event in face_events:
              self.comm.wait_event(event)
            self._push_faces(target, n % len(self.buffers), n == 0)
        elif self.on_gpu:
          with torch.cuda.stream(self.comm):
            for event in face_events:
              self.comm.wait_event(event)
            pending = self._exchange(fed)
        else:
          pending = self._exchange(fed)
        go(low_face, high_face, chunk)
        for event in face_events:      # the next launch reads the faces too
          self.main.wait_event(event)
      if not last:
        nxt = list(current)
        for inp, out in self.feedback.items():
          nxt[inp] = target[out]
        current = nxt
    if p2p and self.world > 1:
      self._finish_run()
    self.outputs = [self.owned(t) for t in target]
    return self.outputs
