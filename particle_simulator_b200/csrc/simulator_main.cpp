// simulator_main.cpp -- the simulator process: a drop-in for the reference's `cuda_simulator` binary
// (reference: cuda_simulator/src/cuda_simulator.cu:7-54 main loop, cuda_simulator/src/lib/frontend.hpp:10-57).
//
// It speaks the editor's protocol through the particle_io C API (include/particle_io.h) and steps on the B200
// through the stepper's C API (include/psim_b200.h); nothing else is linked (no CUDA runtime calls here).
//
//   * connect as a TCP client to the editor (default 0.0.0.0:53123, frontend.hpp:24) or tail a pair of files
//     (frontend.hpp:16-20); wait for the first frame that carries particles (cuda_simulator.cu:44-49);
//   * upload it, start the first frame, send the ingested scene back (cuda_simulator.cu:28-31);
//   * then, per frame (compute_frame, cuda_simulator.cu:7-26): wait for the running frame, start the next one,
//     and while it runs look at the input: a frame with particles replaces the scene (upload, restart, echo);
//     a header-only frame updates the metadata from the next frame on; otherwise download the finished frame
//     and send it. The download waits only for the finished frame's snapshot, so the copy and the TCP write
//     overlap the next frame's kernels (what the reference's report wanted from its second stream,
//     doc/project.typ:710-719).
//
// Beyond the reference: the grid follows the scene (the reference is compiled for 64 x 64 cells of
// box / 64; here cells keep that width and their number follows the box), capacity follows the scene, and
// the snapshot arrives compacted (no null slots to strip).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "particle_io.h"
#include "psim_b200.h"

namespace {

struct Options {
    std::string connect = "0.0.0.0:53123";  // frontend.hpp:24
    std::string file_in, file_out;          // file mode (frontend.hpp:16-20)
    int device = -1;
    int grid_x_log2 = -1, grid_y_log2 = -1;  // -1: from the scene's box, keeping the reference's cell width
    long max_frames = -1;                    // stop after this many simulated frames (tests); -1: until disconnected
    uint32_t min_capacity = 1u << 16;        // kernel.cuh:20
    bool native_schedule = false;
    bool verbose = false;
    uint32_t snapshot_stride = 1;            // send every k-th particle only (decimated frames for display)
};

void usage(const char* argv0) {
    std::fprintf(stderr,
                 "usage: %s [--connect host:port | --files in.bin out.bin] [--device N] [--grid LX LY]\n"
                 "          [--frames N] [--capacity N] [--snapshot-stride K] [--native-schedule] [--verbose]\n"
                 "Steps particle_io scenes on a B200 and streams snapshots back (drop-in for cuda_simulator).\n",
                 argv0);
}

constexpr double kReferenceCellWidth = 50e-9 / 64;  // particle.rs:141-142 box over kernel.cuh:15-18 cells

int grid_log2_for(float box, int forced) {
    if (forced >= 0) return forced;
    int l = (int)std::lround(std::log2((double)box / kReferenceCellWidth));
    return l < 3 ? 3 : (l > 15 ? 15 : l);
}

struct Frontend {  // frontend.hpp:10-57
    Reader reader{};
    Writer writer{};
    bool is_connected = false;

    ~Frontend() {
        if (is_connected) {
            reader_destroy(&reader);
            writer_destroy(&writer);
        }
    }
    // newest frame the editor sent, or {nullptr} (reader_read_last drains the queue)
    Frame read() {
        Frame f{nullptr, 0, 0};
        if (!is_connected) return f;
        is_connected = reader_read_last(&reader, &f);
        return f;
    }
    void write(FrameHeader* frame) {
        if (!is_connected) return;
        is_connected = writer_write(&writer, frame);
    }
};

struct Simulator {
    Options opt;
    PsimStepper* stepper = nullptr;
    PsimConfig cfg{};
    FrameHeader* out = nullptr;  // page-locked frame the snapshots are downloaded into
    uint32_t out_capacity = 0;
    long frames_done = 0;

    ~Simulator() {
        if (stepper) psim_destroy(stepper);
        if (out) psim_host_free(out);
    }

    bool check(int rc, const char* what) {
        if (rc == PSIM_OK) return true;
        std::fprintf(stderr, "psim_simulator: %s failed (%d): %s\n", what, rc, psim_last_error(stepper));
        return false;
    }

    // (Re)create the stepper when the scene needs another grid or more room.
    bool fit_to(const FrameHeader* scene) {
        const uint32_t lx = (uint32_t)grid_log2_for(scene->metadata.box_width, opt.grid_x_log2);
        const uint32_t ly = (uint32_t)grid_log2_for(scene->metadata.box_height, opt.grid_y_log2);
        uint32_t want = scene->particle_count + scene->particle_count / 4;
        if (want < opt.min_capacity) want = opt.min_capacity;
        if (stepper && cfg.grid_x_log2 == lx && cfg.grid_y_log2 == ly && cfg.max_particles >= scene->particle_count) return true;
        if (stepper) {
            psim_destroy(stepper);
            stepper = nullptr;
        }
        cfg = psim_default_config();
        cfg.grid_x_log2 = lx;
        cfg.grid_y_log2 = ly;
        cfg.max_particles = want;
        cfg.device = opt.device;
        cfg.schedule = opt.native_schedule ? PSIM_SCHEDULE_NATIVE : PSIM_SCHEDULE_REFERENCE;
        cfg.snapshot_buffers = 2;  // frame k is copied out and sent while frame k+1 runs
        cfg.use_graph = 1;         // small scenes are launch-bound: replay frames as CUDA graphs
        int rc = psim_create(&cfg, &stepper);
        if (rc != PSIM_OK) {
            std::fprintf(stderr, "psim_simulator: cannot create the stepper (%d): %s\n", rc, psim_last_error(nullptr));
            stepper = nullptr;
            return false;
        }
        if (out_capacity < want) {
            if (out) psim_host_free(out);
            out = static_cast<FrameHeader*>(psim_host_alloc(packet_size(want)));
            out_capacity = out ? want : 0;
            if (!out) {
                std::fprintf(stderr, "psim_simulator: cannot allocate %zu bytes of page-locked memory\n", packet_size(want));
                return false;
            }
        }
        if (opt.verbose)
            std::fprintf(stderr, "psim_simulator: %u x %u cells, room for %u particles\n", 1u << lx, 1u << ly, want);
        return true;
    }

    // upload + start the first frame + echo the ingested scene (cuda_simulator.cu:28-31 and :16-20)
    bool start_scene(Frontend& fe, const FrameHeader* scene) {
        if (!fit_to(scene)) return false;
        if (!check(psim_set_snapshot_stride(stepper, 1), "psim_set_snapshot_stride")) return false;  // the echo is complete
        if (!check(psim_upload_frame(stepper, scene), "psim_upload_frame")) return false;
        if (!check(psim_set_snapshot_stride(stepper, opt.snapshot_stride ? opt.snapshot_stride : 1), "psim_set_snapshot_stride"))
            return false;
        out->particle_count = out_capacity;
        if (!check(psim_download_frame(stepper, out), "psim_download_frame")) return false;  // the binned scene
        if (!check(psim_run_frame_async(stepper), "psim_run_frame_async")) return false;
        fe.write(out);
        return true;
    }

    // compute_frame (cuda_simulator.cu:7-26)
    bool compute_frame(Frontend& fe) {
        if (!check(psim_sync(stepper), "psim_sync")) return false;             // frame k is done, its snapshot packed
        if (!check(psim_run_frame_async(stepper), "psim_run_frame_async")) return false;  // frame k+1 starts
        Frame in = fe.read();
        if (in.ptr && in.ptr->particle_count > 0) {  // a new scene replaces everything (the frame just started is lost,
            bool ok = start_scene(fe, in.ptr);       // as in the reference)
            frame_destroy(&in);
            return ok;
        }
        if (in.ptr) {  // interactive mode: only the metadata changes, from the frame after the running one on
            if (psim_set_metadata(stepper, &in.ptr->metadata) != PSIM_OK)  // e.g. a data_structure switch without a scene
                std::fprintf(stderr, "psim_simulator: metadata update ignored: %s\n", psim_last_error(stepper));
            frame_destroy(&in);
        }
        // frame k's snapshot (age 1: frame k+1's is enqueued already) is copied out and sent while frame k+1 runs
        out->particle_count = out_capacity;
        if (!check(psim_download_frame_ex(stepper, 1, out), "psim_download_frame_ex")) return false;
        fe.write(out);
        frames_done += 1;
        return true;
    }
};

}  // namespace

int main(int argc, char** argv) {
    Options opt;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&](int n) {
            if (i + n >= argc) {
                usage(argv[0]);
                std::exit(2);
            }
        };
        if (a == "--connect") { need(1); opt.connect = argv[++i]; }
        else if (a == "--files") { need(2); opt.file_in = argv[++i]; opt.file_out = argv[++i]; }
        else if (a == "--device") { need(1); opt.device = std::atoi(argv[++i]); }
        else if (a == "--grid") { need(2); opt.grid_x_log2 = std::atoi(argv[++i]); opt.grid_y_log2 = std::atoi(argv[++i]); }
        else if (a == "--frames") { need(1); opt.max_frames = std::atol(argv[++i]); }
        else if (a == "--capacity") { need(1); opt.min_capacity = (uint32_t)std::atol(argv[++i]); }
        else if (a == "--snapshot-stride") { need(1); opt.snapshot_stride = (uint32_t)std::atol(argv[++i]); }
        else if (a == "--native-schedule") opt.native_schedule = true;
        else if (a == "--verbose") opt.verbose = true;
        else { usage(argv[0]); return a == "--help" || a == "-h" ? 0 : 2; }
    }

    // Fail before touching the network when there is no B200: this simulator has no CPU path.
    {
        PsimConfig probe = psim_default_config();
        probe.device = opt.device;
        probe.max_particles = 1024;
        PsimStepper* s = nullptr;
        int rc = psim_create(&probe, &s);
        if (rc != PSIM_OK) {
            std::fprintf(stderr, "psim_simulator: %s\n", psim_last_error(nullptr));
            return 3;
        }
        psim_destroy(s);
    }

    Frontend fe;
    if (!opt.file_in.empty()) {
        reader_open_file(&fe.reader, opt.file_in.c_str());
        writer_open_file(&fe.writer, opt.file_out.c_str());
        fe.is_connected = true;
    } else {
        fe.is_connected = new_tcp_client(&fe.reader, &fe.writer, opt.connect.c_str());
        if (!fe.is_connected) return 4;
    }

    Simulator sim;
    sim.opt = opt;
    // wait for the first scene (cuda_simulator.cu:44-49)
    Frame first{nullptr, 0, 0};
    while (fe.is_connected) {
        first = fe.read();
        if (first.ptr && first.ptr->particle_count > 0) break;
        if (first.ptr) frame_destroy(&first);
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    if (!fe.is_connected) return 0;
    bool ok = sim.start_scene(fe, first.ptr);
    frame_destroy(&first);
    while (ok && fe.is_connected && (opt.max_frames < 0 || sim.frames_done < opt.max_frames)) ok = sim.compute_frame(fe);
    if (sim.stepper) psim_sync(sim.stepper);
    if (opt.verbose) std::fprintf(stderr, "psim_simulator: %ld frames sent\n", sim.frames_done);
    return ok ? 0 : 1;
}
