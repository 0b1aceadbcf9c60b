"""B200-native chunk production for wgpu-voxel-engine.

Host mirror of the reference's interface for the hot path (Chunk::new,
Chunk::create_mesh, Frustum::contains) over the C ABI of libvoxb200.so
(include/voxb200.h).  The directory name carries a hyphen, so import it through
the top-level alias:  `import voxb200`.
"""
from . import _lib, cube
from ._lib import NoiseParams, VoxError, build, SO_PATH
from .chunk import Chunk, ChunkMesh, build_chunks
from . import chunk_builder
from .chunk_builder import ChunkGenerator, Player, World
from .context import (Context, MeshBatch, PinnedBuffer, INSTANCE_DTYPE, VOXELS, ROWS, default_context,
                      default_noise_params, planes_from_matrix)
from .coordinate import ChunkCoord3D, world_to_chunk_coord
from .frustum_culling import Frustum

__all__ = ["Chunk", "ChunkMesh", "build_chunks", "Context", "MeshBatch", "PinnedBuffer", "INSTANCE_DTYPE",
           "VOXELS", "ROWS", "default_context", "default_noise_params", "planes_from_matrix", "ChunkCoord3D",
           "world_to_chunk_coord", "Frustum", "ChunkGenerator", "Player", "World", "chunk_builder", "NoiseParams", "VoxError", "build", "SO_PATH", "cube"]
