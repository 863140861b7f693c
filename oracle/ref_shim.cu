// ref_shim.cu -- TEST INFRASTRUCTURE ONLY (never linked into the product libraries).
//
// Compiles the reference's own step code, unmodified, straight from /root/reference
// (`#include "kernel.cuh"` pulls particle.cuh, kernel_bucket.cuh, kernel_compact.cuh,
// lib/thread_pool.hpp, lib/log.hpp) and exposes it behind a small C ABI so that tests and
// bench.py can drive the REAL reference implementation of the hot path:
//   kernel_prepare_frame (kernel.cuh:200-250), bucket_move / bucket_step
//   (kernel_bucket.cuh:112-180), Kernel::run_async (kernel.cuh:139-151).
// `particle_io.h` is this repo's include/particle_io.h; the three C-API functions the
// reference code references (packet_size, frame_header_init, frame_compact_into) come from
// this repo's particle_io.cpp, compiled into the same shared object.
// Built by oracle/Makefile into oracle/_ref/ (git-ignored). For grids other than the
// reference's fixed 64x64 the Makefile compiles a temporary copy of the reference sources in
// which only the two `#define BUCKETS_{X,Y}_LOG2` lines are rewritten.
#include <pthread.h>

#include <chrono>
#include <cstring>
#include <vector>

#include "kernel.cuh"

namespace {

DeviceBufferId g_cur = D_BUFFER_0;
std::vector<uint8_t> g_src_copy;

struct PrepareArgs {
    FrameHeader* src;
};

void* prepare_thread(void* p) {
    // kernel_prepare_frame keeps `bucket_len[BUCKETS_X * BUCKETS_Y]` on the stack
    // (kernel.cuh:217); run it on a thread whose stack is large enough for scaled grids.
    kernel_prepare_frame(static_cast<PrepareArgs*>(p)->src, kernel.h_frame);
    return nullptr;
}

}  // namespace

extern "C" {

uint32_t ref_grid_x_log2(void) { return BUCKETS_X_LOG2; }
uint32_t ref_grid_y_log2(void) { return BUCKETS_Y_LOG2; }
uint32_t ref_bucket_capacity(void) { return BUCKET_CAPACITY; }
uint32_t ref_slot_count(void) { return MAX_PARTICLE_COUNT; }
int ref_gpu_count(void) { return kernel.gpus_count; }
uint32_t ref_hardware_threads(void) { return std::thread::hardware_concurrency(); }

// Ingest a compact frame: kernel_prepare_frame + Kernel::write(D_BUFFER_0)
// (cuda_simulator.cu:28-29). Returns the device the reference will actually use
// (Gpu is demoted to CpuThreadPool when there is no GPU, kernel.cuh:203-205).
uint32_t ref_prepare(const FrameHeader* src) {
    size_t size = packet_size(src->particle_count);
    g_src_copy.assign(reinterpret_cast<const uint8_t*>(src), reinterpret_cast<const uint8_t*>(src) + size);
    PrepareArgs args{reinterpret_cast<FrameHeader*>(g_src_copy.data())};

    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, (size_t)64 << 20 | ((size_t)BUCKETS_COUNT * 8));
    pthread_t th;
    pthread_create(&th, &attr, prepare_thread, &args);
    pthread_join(th, nullptr);
    pthread_attr_destroy(&attr);

    kernel.write(D_BUFFER_0);
    g_cur = D_BUFFER_0;
    return kernel.h_frame->metadata.device;
}

// Replace the metadata of the current buffer (Kernel::write_metadata semantics, kernel.cuh:96-101).
void ref_set_metadata(const FrameMetadata* meta) {
    kernel.buffer[g_cur].frame.metadata = *meta;
    kernel.h_frame->metadata = *meta;
}

// Copy the whole current slot array (incl. null slots, ty < 0) to `out` (ref_slot_count() records).
void ref_read_slots(Particle* out) {
    kernel.sync();
    kernel.read(g_cur);
    std::memcpy(out, kernel.h_frame->particles, sizeof(Particle) * (size_t)MAX_PARTICLE_COUNT);
}

// Compact the current slot array into `dst` (capacity in dst->particle_count), the frame the
// reference would put on the wire (frontend.hpp:50-56).
void ref_read_compact(FrameHeader* dst) {
    kernel.sync();
    kernel.read(g_cur);
    frame_compact_into(kernel.h_frame, dst);
}

static DeviceBufferId other_of(DeviceBufferId id) { return id == D_BUFFER_0 ? D_BUFFER_1 : D_BUFFER_0; }

// One bucket_move (kernel_bucket.cuh:146-180) cur -> other.
void ref_move(void) {
    DeviceBufferId dst = other_of(g_cur);
    kernel.buffer[dst].frame = kernel.buffer[g_cur].frame;
    bucket_move(kernel.buffer[g_cur].frame, g_cur, dst);
    kernel.sync();
    g_cur = dst;
}

// One bucket_step (kernel_bucket.cuh:112-144) cur -> other.
void ref_step(void) {
    DeviceBufferId dst = other_of(g_cur);
    kernel.buffer[dst].frame = kernel.buffer[g_cur].frame;
    bucket_step(kernel.buffer[g_cur].frame, g_cur, dst);
    kernel.sync();
    g_cur = dst;
}

// One frame: Kernel::run_async(cur, other) + sync (cuda_simulator.cu:7-9). Returns seconds spent
// between enqueue and completion (what BASELINE.md section 3 calls the timed region).
double ref_run_frame(void) {
    DeviceBufferId dst = other_of(g_cur);
    auto t0 = std::chrono::steady_clock::now();
    kernel.run_async(g_cur, dst);
    kernel.sync();
    // CpuMainThread has nothing to wait for; Gpu/CpuThreadPool were waited on by sync().
    auto t1 = std::chrono::steady_clock::now();
    g_cur = dst;
    return std::chrono::duration<double>(t1 - t0).count();
}

// All-pairs cross-check on a compact array (kernel_compact.cuh:4-34), single thread.
void ref_compact_step(const Particle* src, Particle* dst, const FrameMetadata* meta, uint32_t count) {
    for (uint32_t i = 0; i < count; ++i) compact_step_kernel(src, dst, *meta, count, i);
}

// Scalar physics primitives, for spot checks of the restatement.
float ref_params_C(MiePotentialParams p) { return ParticleParams(p).C; }
float ref_f_force(MiePotentialParams p, float r) { return ParticleParams(p).f_force(r); }

}  // extern "C"
