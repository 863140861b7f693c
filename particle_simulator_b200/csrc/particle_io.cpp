// particle_io.cpp -- C++ implementation of the reference's `particle_io` C API.
//
// Exports, symbol for symbol, the 15 `extern "C"` functions of the reference crate
// particle_io/c_api (Rust, no toolchain in this image) so that anything written against
// the reference's generated `particle_io.h` links against libparticle_io_c.so unchanged.
// Behaviour follows particle_io/src/{particle,reader,writer,tcp}.rs; each function cites
// the lines it restates.  Nothing here touches the GPU.
#include "particle_io.h"

#include <arpa/inet.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>

static_assert(sizeof(Particle) == 20, "Particle layout (particle.rs:10-18)");
static_assert(sizeof(MiePotentialParams) == 16, "MiePotentialParams layout");
static_assert(sizeof(FrameMetadata) == 80, "FrameMetadata layout (particle.rs:111-130)");
static_assert(sizeof(FrameHeader) == 96, "FrameHeader layout (particle.rs:192-204)");
static_assert(offsetof(FrameHeader, particle_count) == 4, "");
static_assert(offsetof(FrameHeader, metadata) == 8, "");
static_assert(offsetof(FrameHeader, signature_end) == 88, "");
static_assert(offsetof(FrameMetadata, cursor_pos) == 32, "");
static_assert(offsetof(FrameMetadata, step_dt) == 44, "");
static_assert(offsetof(FrameMetadata, steps_per_frame) == 48, "");
static_assert(offsetof(FrameMetadata, box_width) == 52, "");
static_assert(offsetof(FrameMetadata, data_structure) == 60, "");
static_assert(offsetof(FrameMetadata, gpu_threads_per_block_log2) == 68, "");
static_assert(sizeof(Frame) == 24, "Frame layout (c_api/src/particle.rs:4-10)");
static_assert(sizeof(Reader) == 16 && sizeof(Writer) == 16, "opaque handles are 2 x u64");

namespace {

// particle.rs:207-208
constexpr uint8_t kSignatureStart[4] = {0x36, 0xbc, 0xe9, 0xbd};
constexpr uint8_t kSignatureEnd[4] = {0xac, 0xc4, 0x12, 0xec};

bool header_is_valid(const FrameHeader& h) {  // particle.rs:210-212
    return std::memcmp(h.signature_start, kSignatureStart, 4) == 0 &&
           std::memcmp(h.signature_end, kSignatureEnd, 4) == 0;
}

[[noreturn]] void die(const char* what, const char* detail) {
    // The reference unwrap()s / expect()s here, which aborts the process through a Rust panic
    // across the FFI boundary.  We abort with the same kind of message on stderr.
    std::fprintf(stderr, "[particle_io_c] fatal: %s: %s\n", what, detail);
    std::abort();
}

// ---------------------------------------------------------------------------------------------
// Reader (particle_io/src/reader.rs:12-112): a background thread doing blocking reads into a
// bounded queue of 2048 frames; the consumer side never blocks.
// ---------------------------------------------------------------------------------------------
struct ReaderImpl {
    static constexpr size_t kMaxEnqueuedFrames = 2048;  // reader.rs:17

    int fd = -1;
    bool is_socket = false;
    std::thread worker;
    std::mutex mtx;
    std::condition_variable cv_space;
    std::deque<Frame> queue;          // frames owned by the library until handed out
    bool disconnected = false;        // producer has exited (stream error / closed)
    std::atomic<bool> closing{false}; // consumer dropped the reader

    // reader.rs:76-112 read_blocking: EOF => sleep 1 ms and retry (files are tailed);
    // a closed TCP stream is an error (tcp.rs:9-20).
    bool read_blocking(uint8_t* buf, size_t len) {
        while (len > 0) {
            if (closing.load(std::memory_order_relaxed)) return false;
            ssize_t n = ::read(fd, buf, len);
            if (n > 0) {
                buf += n;
                len -= (size_t)n;
            } else if (n == 0) {
                if (is_socket) return false;  // "Tcp connection closed"
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
            } else if (errno == EINTR) {
                continue;
            } else {
                return false;
            }
        }
        return true;
    }

    void run() {
        for (;;) {
            FrameHeader header;
            std::memset(&header, 0, sizeof header);  // set_invalid_signature (reader.rs:27)
            if (!read_blocking(reinterpret_cast<uint8_t*>(&header), sizeof header)) break;

            if (!header_is_valid(header)) {  // reader.rs:34-37: skip these 96 bytes, no resync
                std::fprintf(stderr, "Read frame with invalid signature\n");
                continue;
            }

            size_t size = packet_size(header.particle_count);
            uint8_t* bytes = static_cast<uint8_t*>(std::malloc(size));
            if (!bytes) break;
            std::memcpy(bytes, &header, sizeof header);
            if (!read_blocking(bytes + sizeof header, size - sizeof header)) {
                std::free(bytes);
                break;
            }

            std::unique_lock<std::mutex> lk(mtx);
            cv_space.wait(lk, [&] { return queue.size() < kMaxEnqueuedFrames || closing.load(); });
            if (closing.load()) {
                std::free(bytes);
                break;
            }
            queue.push_back(Frame{reinterpret_cast<FrameHeader*>(bytes), size, size});
        }
        std::lock_guard<std::mutex> lk(mtx);
        disconnected = true;
    }

    // reader.rs:64-73: Ok(Some) / Ok(None) / Err(()) -> 1 / 0 / -1
    int try_read(Frame* out) {
        std::lock_guard<std::mutex> lk(mtx);
        if (!queue.empty()) {
            *out = queue.front();
            queue.pop_front();
            cv_space.notify_one();
            return 1;
        }
        return disconnected ? -1 : 0;
    }

    ~ReaderImpl() {
        closing.store(true);
        if (is_socket && fd >= 0) ::shutdown(fd, SHUT_RDWR);  // tcp.rs:35-38, unblocks read()
        cv_space.notify_all();
        if (worker.joinable()) worker.join();
        if (fd >= 0) ::close(fd);
        for (Frame& f : queue) std::free(f.ptr);
    }
};

// Writer (particle_io/src/writer.rs:4-28): synchronous write_all on the caller's thread.
struct WriterImpl {
    int fd = -1;
    bool is_socket = false;

    bool write_all(const uint8_t* buf, size_t len, std::string* err) {
        while (len > 0) {
            ssize_t n = is_socket ? ::send(fd, buf, len, MSG_NOSIGNAL) : ::write(fd, buf, len);
            if (n > 0) {
                buf += n;
                len -= (size_t)n;
            } else if (n < 0 && errno == EINTR) {
                continue;
            } else {
                *err = n == 0 ? "failed to write whole buffer" : std::strerror(errno);
                return false;
            }
        }
        return true;
    }

    ~WriterImpl() {
        if (fd >= 0) {
            if (is_socket) ::shutdown(fd, SHUT_RDWR);
            ::close(fd);
        }
    }
};

ReaderImpl*& impl(Reader* r) { return *reinterpret_cast<ReaderImpl**>(&r->_raw[0]); }
WriterImpl*& impl(Writer* w) { return *reinterpret_cast<WriterImpl**>(&w->_raw[0]); }

void start_reader(Reader* reader, int fd, bool is_socket) {
    ReaderImpl* r = new ReaderImpl;
    r->fd = fd;
    r->is_socket = is_socket;
    r->worker = std::thread([r] { r->run(); });
    reader->_raw[0] = reader->_raw[1] = 0;
    impl(reader) = r;
}

// Rust's `{}` for f32: shortest decimal that round-trips, never scientific notation.
std::string rust_display_f32(float v) {
    char buf[512];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, res.ptr);
}

}  // namespace

extern "C" {

size_t packet_size(uint32_t particle_count) {  // particle.rs:225-227
    return sizeof(FrameHeader) + sizeof(Particle) * (size_t)particle_count;
}

FrameHeader frame_header_init(void) {  // FrameHeader::new(FrameMetadata::default(), 0)
    const float k_b = 1.380649e-23f;   // particle.rs:134 (f32 after inference from the fields)
    FrameHeader h;
    std::memset(&h, 0, sizeof h);
    std::memcpy(h.signature_start, kSignatureStart, 4);
    std::memcpy(h.signature_end, kSignatureEnd, 4);
    h.particle_count = 0;
    FrameMetadata& m = h.metadata;  // particle.rs:136-163
    m.cursor_pos[0] = -1.f;
    m.cursor_pos[1] = -1.f;
    m.cursor_size = 0.05f;
    m.step_dt = 50e-15f;
    m.steps_per_frame = 100;
    m.box_width = 50e-9f;
    m.box_height = 50e-9f;
    m.data_structure = (uint32_t)MatrixBuckets;
    m.device = (uint32_t)Gpu;
    m.gpu_threads_per_block_log2 = 7;
    m.particles[0] = MiePotentialParams{3.609e-10f, 105.79f * k_b, 14.08f, 6.f};   // nitrogen
    m.particles[1] = MiePotentialParams{3.404e-10f, 117.84f * k_b, 12.085f, 6.f};  // argon
    return h;
}

bool particle_is_null(Particle particle) { return particle.ty < 0; }  // particle.rs:21-23

void frame_destroy(Frame* frame) {  // c_api/src/particle.rs:64-72
    if (frame->ptr != nullptr && frame->cap > 0) {
        std::free(frame->ptr);
        frame->ptr = nullptr;
    }
}

void frame_print(FrameHeader* frame) {  // Display for Frame, particle.rs:246-287
    const FrameMetadata& m = frame->metadata;
    std::string out = "--- Frame ---\n";
    if (!header_is_valid(*frame)) out += "  signature error\n";
    out += "  step dt = " + rust_display_f32(m.step_dt) + "\n";
    out += "  steps per frame = " + std::to_string(m.steps_per_frame) + "\n";
    out += "  box size = (" + rust_display_f32(m.box_width) + ", " + rust_display_f32(m.box_height) + ")\n";
    out += "  ...\n";
    uint32_t n = frame->particle_count;
    if (n == 0) {
        out += "  particles[0] = {}\n";
    } else {
        out += "  particles[" + std::to_string(n) + "] = {\n";
        for (uint32_t i = 0; i < n && i < 5; ++i) {
            const Particle& p = frame->particles[i];
            char line[256];
            // The reference divides by u64::MAX here (particle.rs:272-273), so the percentages
            // read 0.00%; kept as is because it is what a user of the reference sees.
            std::snprintf(line, sizeof line, "    [%u] = { x=%.2f%%, y=%.2f%%, vx=%s, vy=%s, ty=%d }\n", i,
                          100. * (double)p.x / 18446744073709551615.0, 100. * (double)p.y / 18446744073709551615.0,
                          rust_display_f32(p.vx).c_str(), rust_display_f32(p.vy).c_str(), p.ty);
            out += line;
        }
        if (n > 5) out += "    ...\n";
        out += "  }\n";
    }
    out += "-------------\n";
    std::fputs(out.c_str(), stdout);
    std::fflush(stdout);
}

void frame_compact(FrameHeader* frame) {  // Frame::compact, particle.rs:349-368
    Particle* p = frame->particles;
    uint32_t n = frame->particle_count;
    uint32_t first_null = 0;
    while (first_null < n && p[first_null].ty >= 0) ++first_null;
    if (first_null == n) return;
    uint32_t count = first_null;
    for (uint32_t idx = first_null + 1; idx < n; ++idx) {
        if (p[idx].ty >= 0) p[count++] = p[idx];
    }
    frame->particle_count = count;
}

void frame_compact_into(FrameHeader* frame, FrameHeader* dst) {  // particle.rs:371-379
    dst->metadata = frame->metadata;
    uint32_t count = 0;
    for (uint32_t i = 0; i < frame->particle_count; ++i) {
        if (frame->particles[i].ty >= 0) dst->particles[count++] = frame->particles[i];
    }
    dst->particle_count = count;
}

void reader_open_file(Reader* reader, const char* path) {  // c_api/src/reader.rs:18-27
    int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) die("reader_open_file", std::strerror(errno));
    start_reader(reader, fd, false);
}

void reader_destroy(Reader* reader) {  // c_api/src/reader.rs:29-34
    delete impl(reader);
    impl(reader) = nullptr;
}

Frame reader_read(Reader* reader) {  // c_api/src/reader.rs:36-44
    Frame f{nullptr, 0, 0};
    if (impl(reader)->try_read(&f) < 0) die("reader_read", "stream disconnected");
    return f;
}

bool reader_read_last(Reader* reader, Frame* frame) {  // c_api/src/reader.rs:46-63
    Frame last{nullptr, 0, 0};
    bool succeed = true;
    for (;;) {
        Frame f{nullptr, 0, 0};
        int r = impl(reader)->try_read(&f);
        if (r < 0) succeed = false;
        if (r <= 0) break;
        if (last.ptr) std::free(last.ptr);  // older frames are dropped (iterator .last())
        last = f;
    }
    *frame = last;
    return succeed;
}

void writer_open_file(Writer* writer, const char* path) {  // c_api/src/writer.rs:21-30
    int fd = ::open(path, O_WRONLY | O_APPEND | O_CLOEXEC);  // append, no create (writer.rs:17)
    if (fd < 0) die("writer_open_file", std::strerror(errno));
    WriterImpl* w = new WriterImpl;
    w->fd = fd;
    writer->_raw[0] = writer->_raw[1] = 0;
    impl(writer) = w;
}

void writer_destroy(Writer* writer) {  // c_api/src/writer.rs:32-37
    delete impl(writer);
    impl(writer) = nullptr;
}

bool writer_write(Writer* writer, FrameHeader* frame) {  // c_api/src/writer.rs:39-59
    std::string err;
    size_t size = packet_size(frame->particle_count);
    if (!impl(writer)->write_all(reinterpret_cast<const uint8_t*>(frame), size, &err)) {
        std::fprintf(stderr, "[particle_io_c::Writer] %s\n", err.c_str());
        return false;
    }
    return true;
}

bool new_tcp_client(Reader* reader, Writer* writer, const char* addr) {  // c_api/src/tcp.rs:10-34
    std::string s(addr);
    size_t colon = s.rfind(':');
    if (colon == std::string::npos) {
        std::fprintf(stderr, "[particle_io_c::TCP] invalid socket address\n");
        return false;
    }
    std::string host = s.substr(0, colon), port = s.substr(colon + 1);
    if (host.size() >= 2 && host.front() == '[' && host.back() == ']') host = host.substr(1, host.size() - 2);

    addrinfo hints;
    std::memset(&hints, 0, sizeof hints);
    hints.ai_family = AF_UNSPEC;
    hints.ai_socktype = SOCK_STREAM;
    addrinfo* res = nullptr;
    int rc = ::getaddrinfo(host.c_str(), port.c_str(), &hints, &res);
    if (rc != 0) {
        std::fprintf(stderr, "[particle_io_c::TCP] %s\n", ::gai_strerror(rc));
        return false;
    }
    int fd = -1;
    int last_errno = ECONNREFUSED;
    for (addrinfo* ai = res; ai; ai = ai->ai_next) {
        fd = ::socket(ai->ai_family, ai->ai_socktype | SOCK_CLOEXEC, ai->ai_protocol);
        if (fd < 0) {
            last_errno = errno;
            continue;
        }
        if (::connect(fd, ai->ai_addr, ai->ai_addrlen) == 0) break;
        last_errno = errno;
        ::close(fd);
        fd = -1;
    }
    ::freeaddrinfo(res);
    if (fd < 0) {
        std::fprintf(stderr, "[particle_io_c::TCP] %s\n", std::strerror(last_errno));
        return false;
    }
    int fd2 = ::fcntl(fd, F_DUPFD_CLOEXEC, 0);  // try_clone (tcp.rs:42)
    if (fd2 < 0) {
        std::fprintf(stderr, "[particle_io_c::TCP] %s\n", std::strerror(errno));
        ::close(fd);
        return false;
    }
    start_reader(reader, fd, true);
    WriterImpl* w = new WriterImpl;
    w->fd = fd2;
    w->is_socket = true;
    writer->_raw[0] = writer->_raw[1] = 0;
    impl(writer) = w;
    return true;
}

}  // extern "C"
