"""Host mirror of frustum_culling.rs: Frustum::new / update / contains."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .context import Context, default_context, planes_from_matrix


class Frustum:
    def __init__(self, planes: np.ndarray, device: Optional[Context] = None):
        self.planes = np.ascontiguousarray(planes, dtype=np.float32).reshape(6, 4)
        self._ctx = device

    @staticmethod
    def new(global_matrix, device: Optional[Context] = None) -> "Frustum":
        """frustum_culling.rs:10-14; global_matrix = camera.global_matrix, column-major."""
        return Frustum(planes_from_matrix(global_matrix), device)

    def update(self, global_matrix) -> None:
        """frustum_culling.rs:42-45"""
        self.planes = planes_from_matrix(global_matrix)

    def contains(self, pos) -> bool:
        """frustum_culling.rs:16-40 for one ChunkCoord3D."""
        return bool(self.contains_many([tuple(pos)])[0])

    def contains_many(self, coords) -> np.ndarray:
        """The batched form: one launch for a whole coordinate list."""
        ctx = self._ctx or default_context()
        return ctx.frustum_cull(self.planes.reshape(1, 6, 4), coords)[0].astype(bool)
