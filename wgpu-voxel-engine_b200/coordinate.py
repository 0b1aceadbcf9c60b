"""coordinate.rs: ChunkCoord3D and the world <-> chunk conversions."""
from __future__ import annotations

from typing import NamedTuple


def _wrap_i32(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


class ChunkCoord3D(NamedTuple):
    """coordinate.rs:41-46; tuple ordering = the derived lexicographic Ord (x, y, z)."""
    x: int
    y: int
    z: int

    def to_world_position_i32(self):
        """coordinate.rs:60-65 (wrapping i32 multiply, as a release build)."""
        return (_wrap_i32(self.x * 32), _wrap_i32(self.y * 32), _wrap_i32(self.z * 32))

    def to_world_position_f32(self):
        """coordinate.rs:53-58"""
        import numpy as np
        return tuple(np.float32(v) for v in self.to_world_position_i32())


def world_to_chunk_coord(x: int, y: int, z: int) -> ChunkCoord3D:
    """coordinate.rs:13-18: arithmetic >> 5 (floor for negatives)."""
    return ChunkCoord3D(x >> 5, y >> 5, z >> 5)
