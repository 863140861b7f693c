"""The simulator process (particle_simulator_b200/psim_simulator, csrc/simulator_main.cpp) against a fake editor.

The reference's simulator is a TCP client of the editor (cuda_simulator/src/lib/frontend.hpp:22-25,
particle_editor/src/backend.rs:37): it waits for a scene, echoes the ingested scene, then streams one compacted
snapshot per frame; a header-only frame updates the metadata, a frame with particles replaces the scene
(cuda_simulator/src/cuda_simulator.cu:7-54). The same conversation is held here over a loopback socket and over
the file transport (frontend.hpp:16-20), and every frame received is compared with the in-process stepper.
"""
import os
import socket
import subprocess
import time

import numpy as np
import pytest

from conftest import frame_from
from particle_simulator_b200 import FrameBuffer, _build, io
from particle_simulator_b200.frame import HEADER_DTYPE, packet_size

pytestmark = pytest.mark.gpu


def recv_exact(conn: socket.socket, n: int) -> bytes:
    chunks = []
    while n:
        b = conn.recv(min(n, 1 << 20))
        if not b:
            raise ConnectionError("the simulator closed the connection")
        chunks.append(b)
        n -= len(b)
    return b"".join(chunks)


def recv_frame(conn: socket.socket) -> FrameBuffer:
    head = recv_exact(conn, HEADER_DTYPE.itemsize)
    count = int(np.frombuffer(head, dtype=HEADER_DTYPE)[0]["particle_count"])
    return FrameBuffer.from_bytes(head + recv_exact(conn, packet_size(count) - len(head)))


def reference_frames(fb: FrameBuffer, grid, frames: int) -> list[bytes]:
    """[ingested scene, frame 1, frame 2, ...] from the stepper in this process."""
    from particle_simulator_b200.stepper import Stepper

    out = []
    with Stepper(grid, max(fb.count, 65536)) as st:
        st.upload(fb)
        out.append(st.download().tobytes())
        for _ in range(frames):
            st.run_frame_async()
            st.sync()
            out.append(st.download().tobytes())
    return out


@pytest.fixture()
def simulator():
    _build.build_all()
    procs = []

    def start(*args):
        p = subprocess.Popen([_build.SIMULATOR, *args], stderr=subprocess.PIPE, text=True)
        procs.append(p)
        return p

    yield start
    for p in procs:
        if p.poll() is None:
            p.kill()
        p.wait()


def test_tcp_conversation_with_a_fake_editor(golden, simulator):
    g = golden("hex2500")
    meta = g["meta"][0].copy()
    meta["steps_per_frame"] = 18
    scene = frame_from(g["input"], meta)
    want = reference_frames(scene, (6, 6), 3)

    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    srv.settimeout(60)
    proc = simulator("--connect", f"127.0.0.1:{srv.getsockname()[1]}", "--verbose")
    conn, _ = srv.accept()
    conn.settimeout(60)
    try:
        time.sleep(0.05)  # the simulator polls for its first scene
        conn.sendall(scene.tobytes())
        echo = recv_frame(conn)
        assert echo.tobytes() == want[0]                                  # the ingested scene, binned order
        assert echo.particles.tobytes() == g["binned"].tobytes()          # = the reference's own binning
        for k in (1, 2, 3):
            assert recv_frame(conn).tobytes() == want[k], f"frame {k}"   # bit-identical to the in-process stepper

        # interactive mode: a header-only frame changes the metadata a frame or two later, the particles live on
        upd = FrameBuffer(1, meta)
        upd.metadata["steps_per_frame"] = 5
        conn.sendall(upd.tobytes())
        seen = None
        for _ in range(12):
            f = recv_frame(conn)
            assert f.count == scene.count
            if int(f.metadata["steps_per_frame"]) == 5:
                seen = f
                break
        assert seen is not None, "the metadata update never showed up in a snapshot"

        # a new scene on another box: 128 x 128 cells, more particles; echoed in binned order, then stepped
        big = FrameBuffer(120 * 120)
        big.metadata["box_width"] = 100e-9
        big.metadata["box_height"] = 100e-9
        big.metadata["steps_per_frame"] = 18
        io.scene_hex_square(big, 120, 120, (50e-9, 50e-9), 1.0, 5.0, 5.0, 1, seed=9)
        want_big = reference_frames(big, (7, 7), 2)
        conn.sendall(big.tobytes())
        while True:
            f = recv_frame(conn)
            if f.count == big.count:
                break
            assert f.count == scene.count  # frames of the old scene still in flight
        assert f.tobytes() == want_big[0]
        assert recv_frame(conn).tobytes() == want_big[1]
        assert recv_frame(conn).tobytes() == want_big[2]
    finally:
        conn.close()
        srv.close()
    assert proc.wait(timeout=30) == 0, proc.stderr.read()[-2000:]


def test_file_transport(golden, simulator, tmp_path):
    g = golden("liquid4k")
    meta = g["meta"][0].copy()
    meta["steps_per_frame"] = 35
    scene = frame_from(g["input"], meta)
    want = reference_frames(scene, (6, 6), 2)
    fin, fout = tmp_path / "backend_in.bin", tmp_path / "backend_out.bin"
    fin.write_bytes(scene.tobytes())
    fout.write_bytes(b"")  # the writer appends and does not create (particle_io/src/writer.rs:17)
    proc = simulator("--files", str(fin), str(fout), "--frames", "2")
    assert proc.wait(timeout=60) == 0, proc.stderr.read()[-2000:]
    assert fout.read_bytes() == b"".join(want)


def test_previous_snapshot_can_be_downloaded_while_the_next_frame_is_enqueued(golden):
    """PsimConfig.snapshot_buffers = 2: the double buffering of the reference's main loop (cuda_simulator.cu:28-37)."""
    from particle_simulator_b200.stepper import PsimError, Stepper

    g = golden("hex2500")
    meta = g["meta"][0].copy()
    meta["steps_per_frame"] = 18
    fb = frame_from(g["input"], meta)
    want = reference_frames(fb, (6, 6), 3)
    with Stepper((6, 6), 4096, snapshot_buffers=2) as st:
        st.upload(fb)
        st.run_frame_async()          # frame 1
        st.run_frame_async()          # frame 2 enqueued before frame 1 is read
        assert st.download(age=1).tobytes() == want[1]
        st.run_frame_async()          # frame 3
        assert st.download(age=1).tobytes() == want[2]
        assert st.download(age=0).tobytes() == want[3]
    with Stepper((6, 6), 4096) as st:  # one buffer: only the latest snapshot exists
        st.upload(fb)
        st.run_frame_async()
        with pytest.raises(PsimError, match="no snapshot of age 1"):
            st.download(age=1)


def test_decimated_snapshots(golden):
    """psim_set_snapshot_stride: every k-th particle of the cell-sorted state, the state itself untouched."""
    from particle_simulator_b200.stepper import PsimError, Stepper

    g = golden("liquid4k")
    meta = g["meta"][0].copy()
    meta["steps_per_frame"] = 18
    fb = frame_from(g["input"], meta)
    with Stepper((6, 6), 8192) as st:
        st.upload(fb)
        st.run_frame_async()
        full = st.download().particles.copy()
        for stride in (3, 7, 5000):
            st.set_snapshot_stride(stride)
            st.snapshot_async()
            assert st.download().particles.tobytes() == full[::stride].tobytes()
        st.set_snapshot_stride(1)
        st.snapshot_async()
        assert st.download().particles.tobytes() == full.tobytes()
        with pytest.raises(PsimError, match="stride"):
            st.set_snapshot_stride(0)


def test_pipelined_frames_equal_synchronous_ones(golden):
    """psim_stage_frame_async / psim_upload_staged / psim_download_frame_begin / _end: the loop bench.py's e2e leg runs.
    Different scenes go in back to back; every result equals the one the synchronous calls give."""
    from particle_simulator_b200.stepper import PsimError, Stepper

    scenes = []
    for name, steps in (("hex2500", 18), ("liquid4k", 35), ("gas10k", 18), ("hex2500", 5)):
        g = golden(name)
        meta = g["meta"][0].copy()
        meta["steps_per_frame"] = steps
        scenes.append(frame_from(g["input"], meta))
    want = [reference_frames(fb, (6, 6), 1)[1] for fb in scenes]
    got = []
    outs = [FrameBuffer(16384), FrameBuffer(16384)]
    with Stepper((6, 6), 16384, snapshot_buffers=2) as st:
        with pytest.raises(PsimError, match="no frame has been staged"):
            st.upload_staged()
        st.stage_async(scenes[0])
        for k in range(len(scenes)):
            st.upload_staged()
            if k + 1 < len(scenes):
                st.stage_async(scenes[k + 1])
            st.run_frame_async()
            if k:
                st.download_end()
                got.append(outs[(k - 1) & 1].tobytes())
            st.download_begin(outs[k & 1])
            with pytest.raises(PsimError, match="already in flight"):
                st.download_begin(FrameBuffer(16384))
        st.download_end()
        got.append(outs[(len(scenes) - 1) & 1].tobytes())
    assert got == want


def test_frames_replayed_as_cuda_graphs_are_bit_identical(golden):
    """PsimConfig.use_graph: the reference's own scenes are launch-bound; a frame's launches are captured once per
    buffer parity and replayed. Results, counters and metadata updates behave exactly as without graphs."""
    import time

    from particle_simulator_b200.stepper import Stepper

    g = golden("gas10k")
    meta = g["meta"][0].copy()
    meta["steps_per_frame"] = 100
    fb = frame_from(g["input"], meta)
    outs, times = {}, {}
    for use_graph in (False, True):
        with Stepper((6, 6), 16384, use_graph=use_graph) as st:
            st.upload(fb)
            frames = []
            for k in range(6):
                if k == 3:  # interactive mode: new metadata drops the captured frames
                    m2 = meta.copy()
                    m2["step_dt"] = np.float32(5e-15)
                    st.set_metadata(m2)
                if k == 4:  # an odd number of extra steps flips the buffer parity a frame starts with
                    st.step_async(3)
                st.run_frame_async()
                frames.append(st.download().tobytes())
            assert (st.steps_executed, st.rebins_executed) == (6 * 101 + 3, 6 * 6)
            outs[use_graph] = frames
            st.sync()
            t0 = time.perf_counter()
            for _ in range(20):
                st.run_frame_async()
            st.sync()
            times[use_graph] = (time.perf_counter() - t0) / 20
    assert outs[True] == outs[False]
    print(f"10k-particle frame of 101 steps + 6 re-bins: {1e3 * times[False]:.2f} ms launched one by one, "
          f"{1e3 * times[True]:.2f} ms as a CUDA graph")
    assert times[True] < times[False]
