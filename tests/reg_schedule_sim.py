"""CPU model of the register-streaming kernels' schedule (test helper).

The counterpart of schedule_sim.py for ``plan.RegSchedule`` — the kernel
family that runs by default (soda/codegen/cuda/kernel_reg.py).  A block (2-D: a
warp's strip) is executed step by step with the storage and index arithmetic
the emitter generates:

* per node, a register history of ``period`` slots: the plane of age k sits in
  slot ``(step - k) mod period`` (ages node.delay .. node.hist_oldest exist);
* dimension-0 neighbours come from the thread's own vector or, by shuffle,
  from adjacent lanes: a neighbour outside the 32 x vec strip is garbage;
* 3-D: loads with an in-plane offset in a dimension other than 0 read the
  parent's shared ring at slot ``(step - age) mod depth`` with a linear offset
  (wrapping across rows and slots like the real address arithmetic); stages
  write their ring slot AFTER every stage of the step has read (one barrier
  per step); input planes arrive ``prefetch`` steps ahead in the input ring;
* paired kernels carry two lanes per cell: lane A chains iterations
  0..chain-1 from the input, lane B chains iterations chain..depth-1 from the
  plane lane A's output stage produced one step earlier; lane B's output is
  what is stored;
* chunks with a lead-in, trips of whole periods, overlapping tiles, ownership
  and the valid region as in the kernel.

Values are identities (node, iteration lane, linear global index) or -1 for
garbage; a cell keeps its identity only if every operand was the prescribed
one.
"""
import numpy as np

GARBAGE = -1
UNWRITTEN = -7


def _code(tag, coords, dims):
  inside = np.ones(coords[0].shape, dtype=bool)
  lin = np.zeros(coords[0].shape, dtype=np.int64)
  pitch = 1
  for c, n in zip(coords, dims):
    inside &= (c >= 0) & (c < n)
    lin += c.astype(np.int64) * pitch
    pitch *= n
  return np.where(inside, tag * (1 << 40) + lin, GARBAGE)


def run_schedule(sched, dims, chunk_rows, final=True):
  dims = tuple(dims)
  s = sched.sdim
  tile = sched.tile
  plane = sched.plane_elems
  lanes = 2 if sched.paired else 1
  n_nodes = len(sched.nodes)
  outs = [np.full(dims[::-1], UNWRITTEN, dtype=np.int64)
          for _ in sched.outputs]
  valids = sched.program.valid_regions(dims, sched.depth)
  n_tiles = [-(-dims[d] // sched.own[d]) for d in range(s)]
  n_chunks = -(-dims[s] // chunk_rows)
  guard = sched.guard_elems
  pitch = sched.ring_pitch      # slot stride: a plane plus an unwritten gap
  assert pitch >= plane + guard or not guard
  trip = (sched.flat_box if s == 1 else sched.trip) or 1
  pos = np.arange(plane)
  cell, rest = [], pos
  for extent in tile:
    cell.append(rest % extent)
    rest = rest // extent
  feeds = {}      # input node index -> output node feeding lane B
  if sched.paired:
    for node in sched.outputs:
      feeds[sched.inputs[node.output_index].index] = node

  def tag(node, lane):
    """Identity of the tensor a (node, lane) pair stands for."""
    return node.index + lane * n_nodes

  for tile_index in np.ndindex(*n_tiles[::-1]):
    tile_index = tile_index[::-1]
    origin = [tile_index[d] * sched.own[d] - sched.tile_halo_lo[d]
              for d in range(s)]
    gcoord = [origin[d] + cell[d] for d in range(s)]
    owned = np.ones(plane, dtype=bool)
    for d in range(s):
      owned &= cell[d] >= sched.tile_halo_lo[d]
      owned &= cell[d] < tile[d] - sched.tile_halo_hi[d]
      owned &= gcoord[d] < dims[d]
    for chunk in range(n_chunks):
      r0 = chunk * chunk_rows
      r1 = min(dims[s], r0 + chunk_rows)
      base = r0 - sched.lead
      steps = -(-((r1 - r0) + sched.lead + sched.out_delay) // trip) * trip
      # register histories: [node][lane][slot] -> plane of identities
      hist = {n.index: [[np.full(plane, GARBAGE, dtype=np.int64)
                         for _ in range(sched.period)] for _ in range(lanes)]
              for n in sched.nodes if n.hist_oldest is not None}
      hist_rel = {n.index: [None] * sched.period for n in sched.nodes
                  if n.hist_oldest is not None}     # plane a slot holds
      rings = {n.index: [np.full(
          guard + max(n.ring_depth, max([m.ring_depth for m in sched.inputs] +
                                        [1]) if n.is_input else 0) * pitch +
          guard, GARBAGE, dtype=np.int64) for _ in range(lanes)]
               for n in sched.nodes if n.ring_depth and s > 1}
      fb = {k: np.full(plane, GARBAGE, dtype=np.int64) for k in feeds}
      issued = {n.index: -1 for n in sched.inputs}
      # every input ring has the depth of the deepest (one mbarrier per slot
      # covers all inputs)
      in_depth = max([n.ring_depth for n in sched.inputs] + [1])

      def input_plane(node, rel, lane_b_source=None):
        row = base + rel
        return _code(tag(node, 0), gcoord + [np.full(plane, row)], dims)

      for i in range(steps):
        phase = i % trip
        # ---- inputs: the plane of age 0 enters the histories (and, 3-D, the
        # input ring `prefetch` planes ahead)
        for node in sched.inputs:
          if s > 1 and node.ring_depth:
            while issued[node.index] < i + sched.prefetch:
              issued[node.index] += 1
              slot = issued[node.index] % in_depth
              start = guard + slot * pitch
              rings[node.index][0][start:start + plane] = input_plane(
                  node, issued[node.index])
          if node.hist_oldest is None:
            continue
          slot = i % sched.period
          hist[node.index][0][slot] = input_plane(node, i)
          if sched.paired:
            hist[node.index][1][slot] = fb[node.index].copy()
          hist_rel[node.index][slot] = i
        ring_writes = []
        new_fb = {}
        for node in sched.stage_nodes:
          rel = i - node.delay
          row = base + rel
          results = []
          for lane in range(lanes):
            ok = np.ones(plane, dtype=bool)
            for parent, off in node.loads:
              age = node.delay - off[s]
              want_row = row + off[s] - (sched.pair_lag if lane else 0)
              if sched.via_smem(off):
                depth = in_depth if parent.is_input else parent.ring_depth
                if parent.is_input:
                  assert 0 <= age and age + sched.prefetch < depth, (
                      node.ident, parent.ident, 'input ring too shallow')
                  assert not sched.paired, 'paired 3-D is not emitted'
                else:
                  assert parent.delay + 1 <= age <= parent.delay + depth - 1, (
                      node.ident, parent.ident, age, 'ring timing')
                slot = (i - age) % depth
                addr = guard + slot * pitch + pos + sched.plane_offset(off)
                # a read that leaves the plane lands in the gap, never in a
                # slot another warp may be writing in this step
                rel = addr - guard - slot * pitch
                assert ((rel >= plane - pitch) & (rel < pitch)).all(), (
                    node.ident, parent.ident, 'read reaches another slot')
                got = rings[parent.index][lane][addr]
              else:
                assert parent.hist_oldest is not None
                assert parent.delay <= age <= parent.hist_oldest, (
                    node.ident, parent.ident, age, 'register history')
                slot = (i - age) % sched.period
                # planes before the chunk's first step were never produced
                # (registers start as 0: garbage that is never stored)
                if hist_rel[parent.index][slot] is None or i - age < 0:
                  src = np.full(plane, GARBAGE, dtype=np.int64)
                else:
                  assert hist_rel[parent.index][slot] == i - age, (
                      node.ident, parent.ident, age,
                      'history slot overwritten')
                  src = hist[parent.index][lane][slot]
                x = cell[0] + off[0]
                inside = (x >= 0) & (x < tile[0])
                got = np.where(inside, src[np.clip(pos + off[0], 0,
                                                   plane - 1)], GARBAGE)
              if lane == 0 or not parent.is_input:
                expect_tag = tag(parent, lane)
              else:
                expect_tag = tag(feeds[parent.index], 0) if (
                    len(feeds) and sched.chain > 0) else tag(parent, 0)
              want = _code(expect_tag,
                           [g + o for g, o in zip(gcoord, off)] +
                           [np.full(plane, want_row)], dims)
              ok &= (got == want) & (want != GARBAGE)
            mine = _code(tag(node, lane), gcoord + [np.full(
                plane, row - (sched.pair_lag if lane else 0))], dims)
            results.append(np.where(ok, mine, GARBAGE))
          if node.hist_oldest is not None:
            slot = (i - node.delay) % sched.period
            for lane in range(lanes):
              hist[node.index][lane][slot] = results[lane]
            hist_rel[node.index][slot] = i - node.delay
          if node.ring_depth and s > 1:
            ring_writes.append((node, results))
          if node.output_index is not None:
            if sched.paired:
              new_fb[sched.inputs[node.output_index].index] = results[0]
            value = results[-1]
            out_row = row - (sched.pair_lag if sched.paired else 0)
            if r0 <= out_row < r1:
              inside = np.ones(plane, dtype=bool)
              if final:
                valid = valids[node.output_index]
                inside &= valid[s][0] <= out_row < valid[s][1]
                for d in range(s):
                  inside &= (gcoord[d] >= valid[d][0]) & (
                      gcoord[d] < valid[d][1])
              value = np.where(inside, value, 0)
              out = outs[node.output_index]
              index = tuple([out_row] + [gcoord[d][owned]
                                         for d in range(s - 1, -1, -1)])
              assert (out[index] == UNWRITTEN).all(), 'cell stored twice'
              out[index] = value[owned]
        # one barrier per step: ring planes become visible to the next step
        for node, results in ring_writes:
          slot = (i - node.delay) % node.ring_depth
          start = guard + slot * pitch
          for lane in range(lanes):
            rings[node.index][lane][start:start + plane] = results[lane]
        fb.update(new_fb)
  return outs


def check_outputs(sched, dims, outs, final=True):
  dims = tuple(dims)
  valids = sched.program.valid_regions(dims, sched.depth)
  grids = np.meshgrid(*[np.arange(n) for n in dims[::-1]], indexing='ij')
  coords = grids[::-1]
  lanes = 2 if sched.paired else 1
  for node, out, valid in zip(sched.outputs, outs, valids):
    inside = np.ones(dims[::-1], dtype=bool)
    for c, (lo, hi) in zip(coords, valid):
      inside &= (c >= lo) & (c < hi)
    assert (out != UNWRITTEN).all(), 'unwritten cells'
    want = _code(node.index + (lanes - 1) * len(sched.nodes), coords, dims)
    assert (out[inside] == want[inside]).all(), 'wrong cells in valid region'
    if final:
      assert (out[~inside] == 0).all(), 'border not zeroed'
