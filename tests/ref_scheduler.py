"""Tick-by-tick CPU restatement of the reference's chunk scheduler (chunk_builder.rs:45-352) with the
oracle doing the per-chunk work.  TEST INFRASTRUCTURE (imports oracle/).

Worker closures are asynchronous in the reference; `latency` = ticks between dispatch and the result
showing up on the channel.  latency = 0 (worker done before update_world of the same tick) gives the
race-free state: every mesh built against all loaded neighbours.  With latency >= 1 the reference
has a race: a chunk dispatched while an adjacent chunk is still in flight snapshots its neighbours
before that arrival (:77-80 vs :318-321), the arrival cannot queue it for a rebuild because it is not
in World.chunks yet (:271-277), and it keeps extra -- hidden -- border faces for good."""
from collections import deque

import numpy as np

from oracle import oracle

RENDER_DISTANCE = 5


def neighbours(c):
    x, y, z = c
    return [(x - 1, y, z), (x + 1, y, z), (x, y, z - 1), (x, y, z + 1)]


class RefScheduler:
    def __init__(self, latency=0):
        self.latency = latency
        self.chunks, self.meshes = {}, {}                # World (world.rs:14-17)
        self.load_queue, self.rebuild_queue, self.in_process = [], [], []
        self.data_chan, self.mesh_chan = deque(), deque()        # flume channels (:29-30)
        self.ticks = 0
        self.dispatched_loads = self.dispatched_rebuilds = 0

    # ---- helpers --------------------------------------------------------------------------
    def _adjacent(self, pos):                             # :134-169 (snapshot at dispatch)
        return [self.chunks.get(n) for n in neighbours(pos)]

    def _is_loaded(self, pos):                            # :257-266
        return pos in self.chunks or pos in self.load_queue or pos in self.in_process

    def _spawn_load(self, pos):                           # worker closure :80-84 / :125-129
        adj = self._adjacent(pos)
        cubes = oracle.generate_terrain(pos)
        mesh = oracle.create_mesh(pos, cubes, adj)
        self.data_chan.append((self.ticks + self.latency, pos, cubes, mesh))
        self.in_process.append(pos)
        self.dispatched_loads += 1

    # ---- one ChunkGenerator::build_chunks call (:45-67) -----------------------------------------
    def tick(self, player, frustum_planes):
        self.ticks += 1
        # load_primary_chunks :171-215
        if player.is_in_new_chunk_pos():
            player.update_chunk_pos()
            px, _, pz = player.chunk
            order = [(px, 0, pz)]
            for z in range(-1, 2):
                order += [(px + 1, 0, pz + z), (px - 1, 0, pz + z)]
            order += [(px, 0, pz + 1), (px, 0, pz - 1)]
            for pos in order:
                if not self._is_loaded(pos):              # load_chunk_directly :114-132
                    self._spawn_load(pos)
        # enqueue_chunks_in_frustum :217-255
        px, _, pz = player.chunk
        for r in range(2, RENDER_DISTANCE):
            ring = []
            for z in range(-r, r + 1):
                ring += [(px + r, 0, pz + z), (px - r, 0, pz + z)]
            for x in range(-r + 1, r):
                ring += [(px + x, 0, pz + r), (px + x, 0, pz - r)]
            for pos in ring:
                if not self._is_loaded(pos) and oracle.frustum_contains(frustum_planes, pos):
                    self.load_queue.append(pos)
        # process_chunk_loading_queue :69-87
        if self.load_queue:
            self._spawn_load(self.load_queue.pop(0))
        # process_rebuild_queue :89-112
        if not self.in_process and self.rebuild_queue:
            pos = self.rebuild_queue.pop(0)
            if pos in self.chunks:
                mesh = oracle.create_mesh(pos, self.chunks[pos], self._adjacent(pos))
                self.mesh_chan.append((self.ticks + self.latency, pos, mesh))
                self.dispatched_rebuilds += 1
        # update_world :310-332 (one of each per tick)
        if self.data_chan and self.data_chan[0][0] <= self.ticks:
            _, pos, cubes, mesh = self.data_chan.popleft()
            self.chunks[pos] = cubes
            self.meshes[pos] = mesh
            self.in_process = [p for p in self.in_process if p != pos]
            x, y, z = pos                                  # rebuild_adjacent_chunks :269-300
            for n in [(x + 1, y, z), (x - 1, y, z), (x, y, z + 1), (x, y, z - 1)]:
                if n in self.chunks:
                    self.rebuild_queue.append(n)
        if self.mesh_chan and self.mesh_chan[0][0] <= self.ticks:
            _, pos, mesh = self.mesh_chan.popleft()
            self.meshes[pos] = mesh
        # clean_up_queues :302-308
        self.load_queue.clear()
        self.rebuild_queue = sorted(set(self.rebuild_queue))
        # filter_unseen_chunks :334-352
        cx, cy, cz = player.chunk
        for p in [p for p in self.chunks if not (abs(p[0] - cx) <= RENDER_DISTANCE and
                                                 abs(p[2] - cz) <= RENDER_DISTANCE and
                                                 abs(p[1] - cy) <= RENDER_DISTANCE)]:
            del self.chunks[p]
            self.meshes.pop(p, None)
            if p in self.rebuild_queue:
                self.rebuild_queue.remove(p)

    def quiescent(self, player, frustum_planes):
        if self.in_process or self.rebuild_queue or self.data_chan or self.mesh_chan:
            return False
        if player.is_in_new_chunk_pos():
            return False
        px, _, pz = player.chunk
        for r in range(2, RENDER_DISTANCE):
            for z in range(-r, r + 1):
                for pos in [(px + r, 0, pz + z), (px - r, 0, pz + z)]:
                    if pos not in self.chunks and oracle.frustum_contains(frustum_planes, pos):
                        return False
            for x in range(-r + 1, r):
                for pos in [(px + x, 0, pz + r), (px + x, 0, pz - r)]:
                    if pos not in self.chunks and oracle.frustum_contains(frustum_planes, pos):
                        return False
        return True

    def run_until_quiescent(self, player, frustum_planes, max_ticks=5000):
        start = self.ticks
        while not self.quiescent(player, frustum_planes):
            self.tick(player, frustum_planes)
            assert self.ticks - start < max_ticks, "reference scheduler did not settle"
        return self.ticks - start
