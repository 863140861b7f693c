"""The particle_io wire format and its C API (include/particle_io.h), CPU only.

Follows the idea of the reference's (disabled) round-trip tests: frames of different sizes are
concatenated into a byte stream and a Reader must hand back the same frames
(particle_io/src/lib.rs:13-94, reader.rs:114-149, writer.rs:30-67).
"""
import ctypes
import os
import socket
import threading
import time

import numpy as np
import pytest

from particle_simulator_b200 import FrameBuffer, default_metadata, io, packet_size
from particle_simulator_b200.frame import HEADER_DTYPE, METADATA_DTYPE, PARTICLE_DTYPE, SIGNATURE_END, SIGNATURE_START


def make_frame(n: int, seed: int, nulls: bool = False) -> FrameBuffer:
    rng = np.random.default_rng(seed)
    fb = FrameBuffer(max(n, 1))
    p = np.zeros(n, dtype=PARTICLE_DTYPE)
    p["x"] = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    p["y"] = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    p["vx"] = rng.normal(0, 100, n).astype(np.float32)
    p["vy"] = rng.normal(0, 100, n).astype(np.float32)
    p["ty"] = rng.integers(0, 2, n)
    if nulls:
        p["ty"][rng.random(n) < 0.3] = -1
    fb.set_particles(p)
    fb.metadata["steps_per_frame"] = 7 + seed
    return fb


def wait_for(fn, timeout=5.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        r = fn()
        if r is not None:
            return r
        time.sleep(0.002)
    raise AssertionError("timed out")


def test_layout_sizes_and_offsets():
    # sizeof / offsetof of the #[repr(C)] structs
    assert PARTICLE_DTYPE.itemsize == 20 and METADATA_DTYPE.itemsize == 80 and HEADER_DTYPE.itemsize == 96
    off = {k: HEADER_DTYPE.fields[k][1] for k in HEADER_DTYPE.names}
    assert off == {"signature_start": 0, "particle_count": 4, "metadata": 8, "signature_end": 88, "_padding": 92}
    moff = {k: METADATA_DTYPE.fields[k][1] for k in METADATA_DTYPE.names}
    assert moff["cursor_pos"] == 32 and moff["cursor_size"] == 40 and moff["step_dt"] == 44
    assert moff["steps_per_frame"] == 48 and moff["box_width"] == 52 and moff["box_height"] == 56
    assert moff["data_structure"] == 60 and moff["device"] == 64 and moff["gpu_threads_per_block_log2"] == 68
    assert moff["_padding"] == 72
    assert ctypes.sizeof(io.CFrame) == 24 and ctypes.sizeof(io.CHandle) == 16


def test_packet_size_and_header_init():
    L = io.lib()
    for n in (0, 1, 5, 65536, 10_000_000):
        assert L.packet_size(n) == 96 + 20 * n == packet_size(n)
    h = io.frame_header_init()
    assert bytes(h["signature_start"]) == SIGNATURE_START and bytes(h["signature_end"]) == SIGNATURE_END
    assert h["particle_count"] == 0 and h["_padding"] == 0
    # FrameMetadata::default(), particle.rs:132-165 -- and the Python mirror of it
    assert h["metadata"].tobytes() == default_metadata().tobytes()
    m = h["metadata"]
    assert m["steps_per_frame"] == 100 and m["gpu_threads_per_block_log2"] == 7
    assert m["data_structure"] == 1 and m["device"] == 0
    assert np.float32(m["step_dt"]) == np.float32(50e-15) and np.float32(m["box_width"]) == np.float32(50e-9)
    assert tuple(m["cursor_pos"]) == (-1.0, -1.0)
    assert np.float32(m["particles"][0]["epsilon"]) == np.float32(105.79) * np.float32(1.380649e-23)
    assert np.float32(m["particles"][1]["n"]) == np.float32(12.085)


def test_particle_is_null():
    L = io.lib()
    assert L.particle_is_null(io.CParticle(1, 2, 0.0, 0.0, -1))
    assert not L.particle_is_null(io.CParticle(1, 2, 0.0, 0.0, 0))
    assert not L.particle_is_null(io.CParticle(1, 2, 0.0, 0.0, 1))


@pytest.mark.parametrize("n", [0, 1, 17, 1000])
def test_frame_compact_matches_numpy(n):
    fb = make_frame(n, seed=n, nulls=True)
    want = fb.particles[fb.particles["ty"] >= 0].copy()
    dst = FrameBuffer(max(n, 1))
    io.frame_compact_into(fb, dst)
    assert dst.count == len(want) and dst.particles.tobytes() == want.tobytes()
    assert dst.metadata.tobytes() == fb.metadata.tobytes()
    io.frame_compact(fb)
    assert fb.count == len(want) and fb.particles.tobytes() == want.tobytes()


def test_frame_compact_without_nulls_is_identity():
    fb = make_frame(100, seed=3)
    before = fb.tobytes()
    io.frame_compact(fb)
    assert fb.tobytes() == before


def test_frame_destroy_is_idempotent():
    f = io.CFrame(None, 0, 0)
    io.lib().frame_destroy(ctypes.byref(f))  # cap == 0: no-op (c_api/src/particle.rs:66)
    assert not f.ptr


def test_writer_then_reader_roundtrip(tmp_path):
    path = str(tmp_path / "frames.bin")
    open(path, "wb").close()  # Writer::open_file appends and does not create (writer.rs:17)
    frames = [make_frame(n, seed=i) for i, n in enumerate((3, 0, 250, 1))]
    w = io.Writer.open_file(path)
    for f in frames:
        assert w.write(f)
    w.close()
    assert os.path.getsize(path) == sum(packet_size(f.count) for f in frames)
    r = io.Reader.open_file(path)
    got = [wait_for(r.read) for _ in frames]
    for f, g in zip(frames, got):
        assert g.tobytes() == f.tobytes()
    assert r.read() is None  # Ok(None): nothing queued, file is tailed (reader.rs:64-73)
    # the file is tailed: a frame appended later still arrives
    w = io.Writer.open_file(path)
    extra = make_frame(9, seed=99)
    assert w.write(extra)
    w.close()
    assert wait_for(r.read).tobytes() == extra.tobytes()
    r.close()


def test_reader_read_last_keeps_newest(tmp_path):
    path = str(tmp_path / "frames.bin")
    frames = [make_frame(5 + i, seed=10 + i) for i in range(6)]
    with open(path, "wb") as f:
        for fr in frames:
            f.write(fr.tobytes())
    r = io.Reader.open_file(path)
    time.sleep(0.2)
    ok, last = r.read_last()
    assert ok and last is not None and last.tobytes() == frames[-1].tobytes()
    ok, last = r.read_last()
    assert ok and last is None
    r.close()


def test_reader_skips_invalid_signature(tmp_path, capfd):
    # reader.rs:34-37: an invalid header is dropped (96 bytes) and reading continues
    path = str(tmp_path / "frames.bin")
    good = make_frame(4, seed=5)
    with open(path, "wb") as f:
        f.write(b"\x00" * 96)
        f.write(good.tobytes())
    r = io.Reader.open_file(path)
    got = wait_for(r.read)
    assert got.tobytes() == good.tobytes()
    r.close()
    assert "invalid signature" in capfd.readouterr().err


def test_writer_open_missing_file_aborts(tmp_path):
    # unwrap() in the reference (c_api/src/writer.rs:25-27): the process dies; check in a child
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r); from particle_simulator_b200 import io; "
            "io.Writer.open_file(%r)") % (os.path.dirname(os.path.dirname(__file__)), str(tmp_path / "nope.bin"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True)
    assert p.returncode != 0 and b"writer_open_file" in p.stderr


def test_tcp_client_roundtrip_and_disconnect():
    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]
    received = bytearray()
    to_send = [make_frame(6, seed=21), make_frame(0, seed=22)]
    upstream = make_frame(12, seed=23)

    def serve():
        conn, _ = srv.accept()
        for f in to_send:
            conn.sendall(f.tobytes())
        want = packet_size(upstream.count)
        while len(received) < want:
            chunk = conn.recv(65536)
            if not chunk:
                break
            received.extend(chunk)
        conn.close()

    th = threading.Thread(target=serve)
    th.start()
    pair = io.new_tcp_client(f"127.0.0.1:{port}")
    assert pair is not None
    r, w = pair
    for f in to_send:
        assert wait_for(r.read).tobytes() == f.tobytes()
    assert w.write(upstream)
    th.join(5)
    assert bytes(received) == upstream.tobytes()
    # the peer closed: reader_read_last reports the disconnect (c_api/src/reader.rs:56-62)
    t0 = time.time()
    ok = True
    while ok and time.time() - t0 < 5:
        ok, _ = r.read_last()
        time.sleep(0.005)
    assert not ok
    r.close()
    w.close()
    srv.close()


def test_tcp_connect_failure_returns_false(capfd):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    assert io.new_tcp_client(f"127.0.0.1:{port}") is None
    assert "particle_io_c::TCP" in capfd.readouterr().err


def test_frame_print_format(capfd):
    fb = make_frame(7, seed=1)
    io.lib().frame_print(fb.ptr)
    out = capfd.readouterr().out
    assert out.startswith("--- Frame ---\n") and out.endswith("-------------\n")
    assert "  step dt = 0.00000000000005\n" in out  # Rust `{}` never prints exponents
    assert "  box size = (0.00000005, 0.00000005)\n" in out
    assert "  particles[7] = {\n" in out and "    ...\n" in out
    assert out.count("    [") == 5


def test_scene_hex_square_geometry():
    fb = FrameBuffer(100)
    io.scene_hex_square(fb, 10, 10, (25e-9, 25e-9), 1.0, 5.0, 5.0, 0, seed=1)
    assert fb.count == 100
    r0 = io.force0_r(fb.metadata)
    assert abs(r0 - 4.01084e-10) < 1e-15
    x = fb.particles["x"].astype(np.float64) / (2**32 - 1) * 50e-9
    y = fb.particles["y"].astype(np.float64) / (2**32 - 1) * 50e-9
    # idx_x outer, idx_y inner; odd rows shifted by r0/2; row pitch sin(60 deg) r0 (presets.rs:24-43)
    assert abs((x[10] - x[0]) - r0) < 1e-16
    assert abs((y[1] - y[0]) - np.sin(np.pi / 3) * r0) < 1e-16
    assert abs((x[1] - x[0]) - r0 / 2) < 1e-16
    assert abs(x.mean() - 25e-9 - r0 / 4) < 1e-15 and abs(y.mean() - 25e-9) < 1e-15
    speed = np.hypot(fb.particles["vx"], fb.particles["vy"])
    assert np.allclose(speed, 5.0, rtol=1e-6)
    # seeded: same arguments, same scene; different seed, different velocities
    fb2 = FrameBuffer(100)
    io.scene_hex_square(fb2, 10, 10, (25e-9, 25e-9), 1.0, 5.0, 5.0, 0, seed=1)
    assert fb2.tobytes() == fb.tobytes()
    fb3 = FrameBuffer(100)
    io.scene_hex_square(fb3, 10, 10, (25e-9, 25e-9), 1.0, 5.0, 5.0, 0, seed=2)
    assert fb3.tobytes() != fb.tobytes()
    with pytest.raises(ValueError):
        io.scene_hex_square(fb, 1, 1, (25e-9, 25e-9))  # no room left


def test_scene_gas_respects_distances():
    fb = FrameBuffer(2000)
    sigma = 3.609e-10
    io.scene_gas(fb, 2000, margin=2 * sigma, min_dist=1.1 * sigma, v_min=100, v_max=200, seed=3)
    x = fb.particles["x"].astype(np.float64) / (2**32 - 1) * 50e-9
    y = fb.particles["y"].astype(np.float64) / (2**32 - 1) * 50e-9
    assert x.min() >= 2 * sigma * 0.999 and x.max() <= 50e-9 - 2 * sigma * 0.999
    from scipy.spatial import cKDTree

    d, _ = cKDTree(np.c_[x, y]).query(np.c_[x, y], k=2)
    assert d[:, 1].min() >= 1.1 * sigma * 0.999
    speed = np.hypot(fb.particles["vx"], fb.particles["vy"])
    assert speed.min() >= 100 * 0.999 and speed.max() <= 200 * 1.001
