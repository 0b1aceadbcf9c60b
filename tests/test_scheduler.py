"""the batched scheduler mirror against a tick-by-tick restatement of the
reference's scheduler.  CPU part: the restatement's quiescent state is 'every loaded chunk meshed
against all loaded neighbours'.  GPU part: one batched call per camera pose reaches the same world."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_scheduler import RefScheduler, neighbours  # noqa: E402


def _planes(eye, yaw_deg):
    """Harness-side camera (camera.rs:41-51 composition) -> planes via the oracle."""
    from oracle import oracle
    f32 = np.float32
    gl = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, .5, 0], [0, 0, .5, 1]], dtype=f32)
    t = f32(1.0) / np.tan(f32(np.deg2rad(70.0)) / f32(2))
    near, far, aspect = f32(0.1), f32(400.0), f32(1.6)
    proj = np.zeros((4, 4), f32)
    proj[0, 0], proj[1, 1] = t / aspect, t
    proj[2, 2], proj[2, 3], proj[3, 2] = (far + near) / (near - far), 2 * far * near / (near - far), -1
    a = f32(np.deg2rad(yaw_deg))
    eye = np.array(eye, f32)
    fwd = np.array([np.cos(a), 0, np.sin(a)], f32)
    s = np.cross(fwd, [0, 1, 0]).astype(f32)
    s /= np.linalg.norm(s)
    u = np.cross(s, fwd).astype(f32)
    view = np.eye(4, dtype=f32)
    view[0, :3], view[1, :3], view[2, :3] = s, u, -fwd
    view[:3, 3] = [-(s @ eye), -(u @ eye), fwd @ eye]
    m = (gl @ proj @ view).astype(f32)
    return oracle.matrix_to_planes(m.T.reshape(-1))


POSES = [((5.0, 16.0, 9.0), 30.0), ((40.0, 16.0, 9.0), 200.0), ((150.0, 16.0, -70.0), 95.0)]


def _canon(inst):
    out = inst.copy()
    out["t_matrix"] = out["t_matrix"] + np.float32(0)
    return out


def test_reference_scheduler_quiescent_state(vox, oracle):
    """CPU only.  From an empty world and a fixed camera the reference (workers finishing within the
    tick) settles with every wanted chunk loaded and every mesh built against all loaded neighbours;
    it needs at least one 7-frame tick per chunk for it."""
    player = vox.Player(POSES[0][0])
    sched = RefScheduler(latency=0)
    planes = _planes(*POSES[0])
    ticks = sched.run_until_quiescent(player, planes)
    assert len(sched.chunks) > 10 and ticks > len(sched.chunks)          # >= one tick per chunk
    # startup quirk (player.rs:10-14, chunk_builder.rs:178): no primary 3x3 until the player moves
    assert player.chunk not in sched.chunks
    for pos, cubes in sched.chunks.items():
        adj = [sched.chunks.get(n) for n in neighbours(pos)]
        want = oracle.create_mesh(pos, cubes, adj)
        assert sched.meshes[pos].tobytes() == want.tobytes()
    # entering a new chunk loads the 3x3 around it (chunk_builder.rs:178-214)
    player.update_pos(POSES[1][0])
    sched.run_until_quiescent(player, _planes(*POSES[1]))
    px, _, pz = player.chunk
    assert all((px + dx, 0, pz + dz) in sched.chunks for dx in (-1, 0, 1) for dz in (-1, 0, 1))


def test_reference_scheduler_race_with_slow_workers(vox, oracle):
    """CPU only.  With a one-tick worker latency the reference loads the same chunks but leaves stale
    meshes behind (extra border faces towards neighbours that arrived while the chunk was in flight).
    The batched path targets the race-free state above."""
    p0, p1 = vox.Player(POSES[0][0]), vox.Player(POSES[0][0])
    planes = _planes(*POSES[0])
    fast, slow = RefScheduler(latency=0), RefScheduler(latency=1)
    fast.run_until_quiescent(p0, planes)
    slow.run_until_quiescent(p1, planes)
    assert set(fast.chunks) == set(slow.chunks)
    stale = [p for p in fast.chunks if len(slow.meshes[p]) != len(fast.meshes[p])]
    assert stale and all(len(slow.meshes[p]) > len(fast.meshes[p]) for p in stale)


@pytest.mark.gpu
def test_batched_scheduler_reaches_the_reference_state(vox, oracle):
    ctx = vox.Context()
    player_ref, player_gpu = vox.Player(POSES[0][0]), vox.Player(POSES[0][0])
    ref = RefScheduler(latency=0)
    gen, world = vox.ChunkGenerator(ctx), vox.World()
    total_ticks = 0
    for eye, yaw in POSES:
        planes = _planes(eye, yaw)
        player_ref.update_pos(eye)
        player_gpu.update_pos(eye)
        total_ticks += ref.run_until_quiescent(player_ref, planes)
        calls_before = gen.gpu_calls
        gen.build_chunks(player_gpu, world, vox.Frustum(planes, ctx))
        assert gen.gpu_calls - calls_before <= 1                         # one vx_gen_mesh per pose
        assert gen.replayed_ticks >= 1
        gen.build_chunks(player_gpu, world, vox.Frustum(planes, ctx))    # idempotent once settled
        assert gen.last_batch == 0
        assert set(world.chunks) == set(ref.chunks)
        assert set(world.meshes) == set(ref.meshes)
        for pos in ref.chunks:
            assert np.array_equal(world.chunks[pos].cubes, ref.chunks[pos])
            assert _canon(world.meshes[pos].instance_buffer).tobytes() == _canon(ref.meshes[pos]).tobytes(), pos
        # draw gate (world.rs:34-37), one launch for all meshes == per-mesh Frustum::contains
        drawn = world.visible_meshes(vox.Frustum(planes, ctx), ctx)
        assert drawn == [p for p in world.meshes if oracle.frustum_contains(planes, p)]
    print("reference: %d ticks (x7 frames) to settle 3 poses; batched: %d GPU calls" % (total_ticks, gen.gpu_calls))
    # the artefact-free mode: every touched chunk meshed against all loaded neighbours
    gen2, world2, player2 = vox.ChunkGenerator(ctx, mode="converged"), vox.World(), vox.Player(POSES[0][0])
    for eye, yaw in POSES[:2]:
        player2.update_pos(eye)
        gen2.build_chunks(player2, world2, vox.Frustum(_planes(eye, yaw), ctx))
    for pos, ch in world2.chunks.items():
        adj = [world2.chunks[n].cubes if n in world2.chunks else None for n in neighbours(pos)]
        assert _canon(world2.meshes[pos].instance_buffer).tobytes() == _canon(oracle.create_mesh(pos, ch.cubes, adj)).tobytes()
    ctx.close()


def _build_cpp_sched_demo(vox, tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "wgpu-voxel-engine_b200", "host")
    exe = str(tmp_path / "host_sched_demo")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(host, "host_sched_demo.cpp"),
                           "-L" + os.path.dirname(vox.SO_PATH), "-lvoxb200",
                           "-Wl,-rpath," + os.path.dirname(vox.SO_PATH), "-o", exe])
    return exe


def test_cpp_scheduler_replay_equals_the_python_specification(vox, tmp_path):
    """CPU only.  host/chunk_builder.hpp (C++ mirror of the batched scheduler) replays the reference's queues
    to the same per-chunk neighbour snapshots, in the same order, as chunk_builder.py -- over random loaded
    sets, ring subsets and both 'entered a new chunk' cases."""
    import subprocess
    from importlib import import_module
    cb = import_module(vox.__name__ + ".chunk_builder") if hasattr(vox, "__name__") else None
    exe = _build_cpp_sched_demo(vox, tmp_path)
    rng = np.random.default_rng(11)
    for case in range(12):
        pc = tuple(int(v) for v in (rng.integers(-6, 7), 0, rng.integers(-6, 7)))
        entered = bool(case % 2)
        around = [(pc[0] + dx, 0, pc[2] + dz) for dx in range(-5, 6) for dz in range(-5, 6)]
        loaded = {c for c in around if rng.random() < (0.0 if case < 2 else 0.35)}
        ring = [c for c in cb.ring_coords(pc) if c not in loaded and rng.random() < 0.6]
        gen = vox.ChunkGenerator(device=None)
        player = vox.Player((pc[0] * 32 + 5.0, 16.0, pc[2] * 32 + 5.0))
        snap, _ = gen._replay(entered, player, set(loaded), ring)
        text = "%d %d %d %d\n%d\n" % (entered, pc[0], pc[1], pc[2], len(loaded))
        text += "".join("%d %d %d\n" % c for c in sorted(loaded)) + "%d\n" % len(ring) + "".join("%d %d %d\n" % c for c in ring)
        out = subprocess.check_output([exe, "replay"], input=text, text=True).splitlines()
        assert out[0] == "ticks %d" % gen.replayed_ticks
        got = [tuple(int(v) for v in line.split()[1:]) for line in out[1:]]
        assert got == [c + (m,) for c, m in snap.items()], case


@pytest.mark.gpu
def test_cpp_scheduler_reaches_the_python_world(vox, tmp_path):
    """The C++ ChunkGenerator driven through the same poses ends every pose with the world of the Python one
    (chunks, cube ids, InstanceRaw streams, draw gate), in one GPU batch per pose."""
    import subprocess
    exe = _build_cpp_sched_demo(vox, tmp_path)

    def fnv(b, h):
        for v in np.frombuffer(b, np.uint8).tolist():
            h = ((h ^ v) * 16777619) & 0xFFFFFFFF
        return h
    for mode in ("reference", "converged"):
        ctx = vox.Context()
        gen, world, player = vox.ChunkGenerator(ctx, mode=mode), vox.World(), vox.Player(POSES[0][0])
        text, want = "", []
        for eye, yaw in POSES:
            planes = _planes(eye, yaw)
            text += " ".join("%.9g" % v for v in list(eye) + planes.reshape(-1).tolist()) + "\n"
            player.update_pos(eye)
            fr = vox.Frustum(planes, ctx)
            gen.build_chunks(player, world, fr)
            batch, ticks = gen.last_batch, gen.replayed_ticks
            h, quads = 2166136261, 0
            for pos in sorted(world.chunks):
                h = fnv(np.array(pos, np.int32).tobytes(), h)
                h = fnv(world.chunks[pos].cubes.tobytes(), h)
                inst = _canon(world.meshes[pos].instance_buffer)
                quads += len(inst)
                h = fnv(inst.tobytes(), h)
            drawn = len(world.visible_meshes(fr, ctx))
            want.append("pose %d chunks %d meshes %d batch %d second_batch 0 ticks %d quads %d drawn %d digest %08x" % (
                len(want), len(world.chunks), len(world.meshes), batch, ticks, quads, drawn, h))
        ctx.close()
        out = subprocess.check_output([exe, "run", mode], input=text, text=True).splitlines()
        assert out == want, (mode, out, want)
