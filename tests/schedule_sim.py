"""CPU model of the streaming kernel's schedule (test helper, no arithmetic).

Executes a ``plan.Schedule`` block by block with exactly the index arithmetic
the generated CUDA kernel uses — step-relative ring slots with power-of-two
masks, linear in-plane offsets that wrap across rows and slots, tiles that
overlap by the rounded halo, chunks with a lead-in — but on *identities*
instead of numbers: a cell holds ``node * M + linear global index`` if it was
computed from exactly the operands the program prescribes, else -1 (garbage).
A correct schedule leaves the right identity in every cell of the reference's
valid region of every output.
"""
import numpy as np

GARBAGE = -1


def _code(node, coords, dims):
  """Identity of tensor ``node`` at global ``coords`` (arrays), -1 outside."""
  inside = np.ones(coords[0].shape, dtype=bool)
  lin = np.zeros(coords[0].shape, dtype=np.int64)
  pitch = 1
  for c, n in zip(coords, dims):
    inside &= (c >= 0) & (c < n)
    lin += c.astype(np.int64) * pitch
    pitch *= n
  return np.where(inside, node.index * (1 << 40) + lin, GARBAGE)


def run_schedule(sched, dims, chunk_rows, final=True):
  """Returns ``[array per output]`` of identities, shape dims reversed."""
  dims = tuple(dims)
  s = sched.sdim
  tile = sched.tile
  plane = sched.plane_elems
  outs = [np.full(dims[::-1], -7, dtype=np.int64) for _ in sched.outputs]
  valids = sched.program.valid_regions(dims, sched.depth)
  n_tiles = [-(-dims[d] // sched.own[d]) for d in range(s)]
  n_chunks = -(-dims[s] // chunk_rows)
  guard = sched.guard_elems

  # in-plane coordinates of every linear position
  pos = np.arange(plane)
  cell = []
  rest = pos
  for extent in tile:
    cell.append(rest % extent)
    rest = rest // extent

  for tile_index in np.ndindex(*n_tiles[::-1]):
    tile_index = tile_index[::-1]
    origin = [tile_index[d] * sched.own[d] - sched.tile_halo_lo[d]
              for d in range(s)]
    gcoord = [origin[d] + cell[d] for d in range(s)]
    owned = np.ones(plane, dtype=bool)
    for d in range(s):
      owned &= (cell[d] >= sched.tile_halo_lo[d])
      owned &= (cell[d] < tile[d] - sched.tile_halo_hi[d])
      owned &= (gcoord[d] < dims[d])
    for chunk in range(n_chunks):
      r0 = chunk * chunk_rows
      r1 = min(dims[s], r0 + chunk_rows)
      base = r0 - sched.lead
      rings = {}
      for node in sched.nodes:
        if node.ring_depth:
          rings[node.index] = np.full(
              guard + node.ring_depth * plane + guard, GARBAGE, dtype=np.int64)
      issued = {node.index: -1 for node in sched.inputs}  # newest rel loaded

      def load_input(node, rel):
        row = base + rel
        coords = gcoord + [np.full(plane, row)]
        slot = rel & (node.ring_depth - 1)
        start = guard + slot * plane
        rings[node.index][start:start + plane] = _code(node, coords, dims)

      for i in range(sched.steps(r1 - r0)):
        # producer: keep `prefetch` planes ahead of the consumer
        for node in sched.inputs:
          if not node.ring_depth:
            continue
          while issued[node.index] < i + sched.prefetch:
            issued[node.index] += 1
            load_input(node, issued[node.index])
        results = []
        for node in sched.stage_nodes:
          rel = i - node.delay
          row = base + rel
          ok = np.ones(plane, dtype=bool)
          for parent, off in node.loads:
            prel = rel + off[s]
            # the plane must already be complete: written in an earlier step
            # (inputs: requested at or before this step)
            if parent.is_input:
              assert prel <= i, (node.ident, parent.ident)
              assert prel > issued[parent.index] - parent.ring_depth
            else:
              assert prel <= i - 1 - parent.delay, (node.ident, parent.ident)
              assert prel > i - parent.delay - parent.ring_depth, (
                  node.ident, parent.ident, 'ring too shallow')
            slot = prel & (parent.ring_depth - 1)
            addr = guard + slot * plane + pos + sched.plane_offset(off)
            got = rings[parent.index][addr]
            want = _code(parent, [g + o for g, o in zip(gcoord, off)] +
                         [np.full(plane, row + off[s])], dims)
            ok &= (got == want) & (want != GARBAGE)
          mine = _code(node, gcoord + [np.full(plane, row)], dims)
          results.append((node, rel, row, np.where(ok, mine, GARBAGE)))
        # all stages read before any writes becomes visible: one barrier/step
        for node, rel, row, value in results:
          if node.ring_depth:
            slot = rel & (node.ring_depth - 1)
            start = guard + slot * plane
            rings[node.index][start:start + plane] = value
          if node.output_index is not None and r0 <= row < r1:
            store = owned.copy()
            inside = np.ones(plane, dtype=bool)
            if final:
              valid = valids[node.output_index]
              inside &= valid[s][0] <= row < valid[s][1]
              for d in range(s):
                inside &= (gcoord[d] >= valid[d][0]) & (gcoord[d] < valid[d][1])
            value = np.where(inside, value, 0)
            out = outs[node.output_index]
            index = tuple([row] + [gcoord[d][store] for d in range(s - 1, -1, -1)])
            assert (out[index] == -7).all(), 'cell stored twice'
            out[index] = value[store]
  return outs


def check_outputs(sched, dims, outs, final=True):
  """Every cell of the valid region holds the right identity; with ``final``
  every other cell holds 0; nothing is left unwritten."""
  dims = tuple(dims)
  valids = sched.program.valid_regions(dims, sched.depth)
  grids = np.meshgrid(*[np.arange(n) for n in dims[::-1]], indexing='ij')
  coords = grids[::-1]
  for node, out, valid in zip(sched.outputs, outs, valids):
    inside = np.ones(dims[::-1], dtype=bool)
    for c, (lo, hi) in zip(coords, valid):
      inside &= (c >= lo) & (c < hi)
    assert (out != -7).all(), 'unwritten cells'
    want = _code(node, coords, dims)
    assert (out[inside] == want[inside]).all(), 'wrong cells in valid region'
    if final:
      assert (out[~inside] == 0).all(), 'border not zeroed'
