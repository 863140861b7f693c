/*
 * particle_io.h -- C ABI of the `particle_io` frame format and I/O layer.
 *
 * This header restates, by hand, what cbindgen 0.29.2 would generate from the
 * reference crate `particle_io/c_api` (reference: particle_io/c_api/build.rs:10-24;
 * the generated header is NOT in the reference tree, see its .gitignore).  The
 * implementation behind it is C++ (particle_simulator_b200/csrc/particle_io.cpp),
 * built into libparticle_io_c.so, symbol for symbol what the reference's Rust
 * static library exports, so that code written against the reference (e.g.
 * cuda_simulator/src/lib/frontend.hpp) links against it unchanged.
 *
 * Layouts (checked by static_assert in the implementation and by tests):
 *   Particle 20 B, MiePotentialParams 16 B, FrameMetadata 80 B, FrameHeader 96 B,
 *   Frame 24 B, Reader/Writer 16 B.
 */
#pragma once

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

/* reference: particle_io/src/particle.rs:52-57 */
typedef enum DataStructure {
    CompactArray,
    MatrixBuckets,
} DataStructure;

/* reference: particle_io/src/particle.rs:80-86 */
typedef enum Device {
    Gpu,
    CpuThreadPool,
    CpuMainThread,
} Device;

/* reference: particle_io/src/particle.rs:10-18.
 * x, y are fixed-point fractions of the box: round(u32::MAX * pos / box)
 * (particle.rs:172-173); ty < 0 marks a null slot (particle.rs:21-23). */
typedef struct Particle {
    uint32_t x;
    uint32_t y;
    float vx;
    float vy;
    int32_t ty;
} Particle;

/* reference: particle_io/src/particle.rs:33-41 */
typedef struct MiePotentialParams {
    float sigma;   /* distance (m) at which V = 0 */
    float epsilon; /* dispersion energy (J)      */
    float n;
    float m;
} MiePotentialParams;

/* reference: particle_io/src/particle.rs:111-130 (defaults :132-165) */
typedef struct FrameMetadata {
    MiePotentialParams particles[2];
    float cursor_pos[2];
    float cursor_size;
    float step_dt;
    uint32_t steps_per_frame;
    float box_width;
    float box_height;
    uint32_t data_structure; /* DataStructure stored as u32 */
    uint32_t device;         /* Device stored as u32        */
    uint32_t gpu_threads_per_block_log2;
    uint32_t _padding[2];
} FrameMetadata;

/* reference: particle_io/src/particle.rs:192-208.
 * signature_start = 36 bc e9 bd, signature_end = ac c4 12 ec. */
typedef struct FrameHeader {
    uint8_t signature_start[4];
    uint32_t particle_count;
    FrameMetadata metadata;
    uint8_t signature_end[4];
    uint32_t _padding;
#ifdef __cplusplus
    /* the particle records follow the 96-byte header in memory */
    Particle particles[0];
#else
    Particle particles[];
#endif
} FrameHeader;

/* reference: particle_io/c_api/src/particle.rs:4-10 (a leaked Vec<u8>) */
typedef struct Frame {
    FrameHeader* ptr;
    size_t cap;
    size_t len;
} Frame;

/* reference: particle_io/c_api/src/reader.rs:7-12 -- opaque, caller-allocated */
typedef struct Reader {
    uint64_t _raw[2];
} Reader;

/* reference: particle_io/c_api/src/writer.rs:10-15 -- opaque, caller-allocated */
typedef struct Writer {
    uint64_t _raw[2];
} Writer;

#ifdef __cplusplus
extern "C" {
#endif

/* reference: particle_io/c_api/src/particle.rs:64-72. Idempotent; no-op if cap == 0. */
void frame_destroy(Frame* frame);
/* reference: particle_io/c_api/src/particle.rs:74-79 (Display impl particle.rs:246-287) */
void frame_print(FrameHeader* frame);
/* reference: particle_io/c_api/src/particle.rs:81-90 (Frame::compact particle.rs:349-368) */
void frame_compact(FrameHeader* frame);
/* reference: particle_io/c_api/src/particle.rs:92-102 (Frame::compact_into particle.rs:371-379).
 * dst->particle_count must hold dst's CAPACITY on entry; on return it is the live count. */
void frame_compact_into(FrameHeader* frame, FrameHeader* dst);
/* reference: particle_io/c_api/src/particle.rs:104-107 */
size_t packet_size(uint32_t particle_count);
/* reference: particle_io/c_api/src/particle.rs:109-112 */
FrameHeader frame_header_init(void);
/* reference: particle_io/c_api/src/particle.rs:114-117 */
bool particle_is_null(Particle particle);

/* reference: particle_io/c_api/src/reader.rs:18-27. Aborts if the file cannot be opened. */
void reader_open_file(Reader* reader, const char* path);
/* reference: particle_io/c_api/src/reader.rs:29-34 */
void reader_destroy(Reader* reader);
/* reference: particle_io/c_api/src/reader.rs:36-44. Non-blocking; {NULL,0,0} if nothing queued.
 * Aborts if the stream is disconnected (the reference unwrap()s). */
Frame reader_read(Reader* reader);
/* reference: particle_io/c_api/src/reader.rs:46-63. Drains the queue, keeps the newest frame
 * ({NULL,0,0} if none). Returns false once the stream has disconnected. */
bool reader_read_last(Reader* reader, Frame* frame);

/* reference: particle_io/c_api/src/writer.rs:21-30. Append-only, does not create; aborts on error. */
void writer_open_file(Writer* writer, const char* path);
/* reference: particle_io/c_api/src/writer.rs:32-37 */
void writer_destroy(Writer* writer);
/* reference: particle_io/c_api/src/writer.rs:39-59. Writes packet_size(particle_count) bytes. */
bool writer_write(Writer* writer, FrameHeader* frame);

/* reference: particle_io/c_api/src/tcp.rs:10-34. addr is "host:port". */
bool new_tcp_client(Reader* reader, Writer* writer, const char* addr);

#ifdef __cplusplus
} /* extern "C" */
#endif
