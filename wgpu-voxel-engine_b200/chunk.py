"""Host mirror of the reference's chunk.rs interface, forwarding to the C ABI.

  Chunk::new(position)                          chunk.rs:21
  Chunk::create_mesh(&self, device, adjacent)   chunk.rs:48-57
  ChunkMesh { instance_buffer, instances_len }  chunk.rs:248-251

`device` (an Arc<wgpu::Device> in the reference) is the vox Context here; the
instance buffer is the byte image ChunkMesh::new hands to create_buffer_init
(chunk.rs:255-260), one 68-byte InstanceRaw per quad in the reference's order.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import cube
from .context import Context, default_context, INSTANCE_DTYPE, VOXELS
from .coordinate import ChunkCoord3D


class ChunkMesh:
    def __init__(self, instance_buffer: np.ndarray):
        self.instance_buffer = instance_buffer          # INSTANCE_DTYPE[instances_len]
        self.instances_len = int(instance_buffer.shape[0])

    def as_bytes(self) -> bytes:
        return self.instance_buffer.tobytes()


class Chunk:
    def __init__(self, position: ChunkCoord3D, cubes: np.ndarray):
        self.position = ChunkCoord3D(*position)
        self.cubes = np.ascontiguousarray(cubes, dtype=np.uint8).reshape(VOXELS)

    @staticmethod
    def new(position, device: Optional[Context] = None, params=None) -> "Chunk":
        ctx = device or default_context()
        pos = ChunkCoord3D(*position)
        blocks = ctx.fill([pos], params=params)["blocks"][0]
        return Chunk(pos, blocks)

    def create_mesh(self, device: Optional[Context],
                    adjacent_chunks: Sequence[Optional["Chunk"]]) -> ChunkMesh:
        """adjacent_chunks: [-x, +x, -z, +z], None where not loaded (chunk_builder.rs:134-169)."""
        ctx = device or default_context()
        if len(adjacent_chunks) != 4:
            raise IndexError("adjacent_chunks must have 4 entries")     # the reference would panic on [k]
        coords, blocks, idx = [self.position], [self.cubes], []
        for a in adjacent_chunks:
            if a is None:
                idx.append(-1)
            else:
                idx.append(len(coords))
                coords.append(a.position)
                blocks.append(a.cubes)
        n = len(coords)
        nbr = np.full((n, 4), -1, np.int32)
        nbr[0] = idx
        mask = np.zeros(n, np.uint8)
        mask[0] = 1
        res = ctx.mesh(np.array(coords, np.int32), np.stack(blocks), nbr_idx=nbr, mesh_mask=mask,
                       capacity=98304)
        return ChunkMesh(res.instances.copy())


def build_chunks(positions, device: Optional[Context] = None, params=None, nbr_mask=None):
    """Batched form of the worker closure chunk_builder.rs:80-84: Chunk::new +
    create_mesh for a whole coordinate list in one call (the fast path)."""
    ctx = device or default_context()
    res = ctx.gen_mesh(np.array(positions, np.int32).reshape(-1, 3), params=params, nbr_mask=nbr_mask,
                       blocks=True)
    chunks = [Chunk(ChunkCoord3D(*map(int, p)), res.blocks[i]) for i, p in enumerate(np.asarray(positions).reshape(-1, 3))]
    meshes = [ChunkMesh(res.chunk(i)) for i in range(len(chunks))]
    return chunks, meshes
