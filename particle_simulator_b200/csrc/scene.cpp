// scene.cpp -- seeded scene generators (C ABI in include/psim_scene.h).
//
// Restates ParticleLattice::{hex_square, square, random_vel} of the reference
// (particle_io/src/presets.rs:16-82) and FrameMetadata::new_particle (particle.rs:168-178).
// The reference draws from the unseeded thread RNG; here a splitmix64 stream keyed by `seed`
// is used so that every scene is reproducible from its arguments.
#include "psim_scene.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct Rng {  // splitmix64
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0, 1)
};

// particle.rs:168-178; Rust's `as u32` saturates.
Particle new_particle(const FrameMetadata& meta, double px, double py, double vx, double vy, int32_t ty) {
    auto to_u32 = [](double v) -> uint32_t {
        double r = std::round(v);
        if (!(r > 0.0)) return 0u;
        if (r >= 4294967295.0) return 4294967295u;
        return (uint32_t)r;
    };
    Particle p;
    p.x = to_u32(4294967295.0 * px / (double)meta.box_width);
    p.y = to_u32(4294967295.0 * py / (double)meta.box_height);
    p.vx = (float)vx;
    p.vy = (float)vy;
    p.ty = ty;
    return p;
}

// presets.rs:76-81: speed uniform in [v_min, v_max], angle uniform in [0, 2pi), dir = (sin, cos).
void random_vel(Rng& rng, float v_min, float v_max, double* vx, double* vy) {
    float v = v_min + (float)rng.uniform() * (v_max - v_min);
    float angle = (float)(rng.uniform() * 2.0 * 3.14159265358979323846);
    *vx = (double)std::sin(angle) * (double)v;
    *vy = (double)std::cos(angle) * (double)v;
}

}  // namespace

extern "C" {

double psim_force0_r(MiePotentialParams p) {  // particle.rs:44-49
    double n = p.n, m = p.m, sigma = p.sigma;
    return sigma * std::pow(n / m, 1. / (n - m));
}

int psim_scene_hex_square(FrameHeader* frame, uint32_t capacity, uint32_t nx, uint32_t ny, double center_x,
                          double center_y, float distance_factor, float v_min, float v_max, int32_t ty,
                          uint64_t seed) {  // presets.rs:16-46
    uint64_t total = (uint64_t)nx * ny;
    if (total == 0) return 0;
    if ((uint64_t)frame->particle_count + total > capacity) return -1;
    const FrameMetadata meta = frame->metadata;
    int species = ty > 0 && ty < 2 ? ty : 0;
    double rx = psim_force0_r(meta.particles[species]) * (double)distance_factor;
    double ry = std::sin(3.14159265358979323846 / 3.) * rx;
    double start_x = center_x - rx * (double)(nx - 1) / 2.;
    double start_y = center_y - ry * (double)(ny - 1) / 2.;
    Rng rng(seed);
    Particle* out = frame->particles + frame->particle_count;
    for (uint32_t ix = 0; ix < nx; ++ix) {
        for (uint32_t iy = 0; iy < ny; ++iy) {
            double offset = iy % 2 == 0 ? 0. : rx / 2.;
            double vx, vy;
            random_vel(rng, v_min, v_max, &vx, &vy);
            *out++ = new_particle(meta, start_x + rx * (double)ix + offset, start_y + ry * (double)iy, vx, vy, ty);
        }
    }
    frame->particle_count += (uint32_t)total;
    return 0;
}

int psim_scene_hex_rows(FrameHeader* frame, uint32_t capacity, uint32_t nx, uint32_t ny, uint32_t row_begin,
                        uint32_t row_end, double center_x, double center_y, float distance_factor, float v_min,
                        float v_max, int32_t ty, uint64_t seed) {
    if (row_end > ny || row_begin > row_end) return -1;
    uint64_t total = (uint64_t)nx * (row_end - row_begin);
    if (total == 0) return 0;
    if ((uint64_t)frame->particle_count + total > capacity) return -1;
    const FrameMetadata meta = frame->metadata;
    int species = ty > 0 && ty < 2 ? ty : 0;
    double rx = psim_force0_r(meta.particles[species]) * (double)distance_factor;
    double ry = std::sin(3.14159265358979323846 / 3.) * rx;
    double start_x = center_x - rx * (double)(nx - 1) / 2.;
    double start_y = center_y - ry * (double)(ny - 1) / 2.;
    Particle* out = frame->particles + frame->particle_count;
    for (uint32_t iy = row_begin; iy < row_end; ++iy) {
        Rng rng(seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(iy + 1)));  // one stream per lattice row
        double offset = iy % 2 == 0 ? 0. : rx / 2.;
        for (uint32_t ix = 0; ix < nx; ++ix) {
            double vx, vy;
            random_vel(rng, v_min, v_max, &vx, &vy);
            *out++ = new_particle(meta, start_x + rx * (double)ix + offset, start_y + ry * (double)iy, vx, vy, ty);
        }
    }
    frame->particle_count += (uint32_t)total;
    return 0;
}

int psim_scene_square(FrameHeader* frame, uint32_t capacity, uint32_t nx, uint32_t ny, double center_x,
                      double center_y, float distance_factor, float v_min, float v_max, int32_t ty,
                      uint64_t seed) {  // presets.rs:48-74
    uint64_t total = (uint64_t)nx * ny;
    if (total == 0) return 0;
    if ((uint64_t)frame->particle_count + total > capacity) return -1;
    const FrameMetadata meta = frame->metadata;
    int species = ty > 0 && ty < 2 ? ty : 0;
    double r = psim_force0_r(meta.particles[species]) * (double)distance_factor;
    double start_x = center_x - (double)(nx - 1) / 2. * r;
    double start_y = center_y - (double)(ny - 1) / 2. * r;
    Rng rng(seed);
    Particle* out = frame->particles + frame->particle_count;
    for (uint32_t ix = 0; ix < nx; ++ix) {
        for (uint32_t iy = 0; iy < ny; ++iy) {
            double vx, vy;
            random_vel(rng, v_min, v_max, &vx, &vy);
            *out++ = new_particle(meta, start_x + (double)ix * r, start_y + (double)iy * r, vx, vy, ty);
        }
    }
    frame->particle_count += (uint32_t)total;
    return 0;
}

int psim_scene_gas(FrameHeader* frame, uint32_t capacity, uint32_t count, double margin, double min_dist,
                   float v_min, float v_max, int32_t ty, uint64_t seed) {
    if ((uint64_t)frame->particle_count + count > capacity) return -1;
    const FrameMetadata meta = frame->metadata;
    double w = meta.box_width, h = meta.box_height;
    if (!(w > 2 * margin) || !(h > 2 * margin)) return -1;
    // hash grid of side >= min_dist over the box for the rejection test
    double cell = min_dist > 0 ? min_dist : w;
    uint32_t gx = (uint32_t)std::fmin(4096., std::fmax(1., std::floor(w / cell)));
    uint32_t gy = (uint32_t)std::fmin(4096., std::fmax(1., std::floor(h / cell)));
    std::vector<std::vector<std::pair<double, double>>> grid((size_t)gx * gy);
    auto cell_index = [&](double px, double py) {
        int cx = (int)std::fmin((double)gx - 1, std::fmax(0., px / w * gx));
        int cy = (int)std::fmin((double)gy - 1, std::fmax(0., py / h * gy));
        return std::pair<int, int>(cx, cy);
    };
    // particles already in the frame take part in the distance test
    for (uint32_t i = 0; i < frame->particle_count; ++i) {
        double px = (double)frame->particles[i].x / 4294967295.0 * w;
        double py = (double)frame->particles[i].y / 4294967295.0 * h;
        auto c = cell_index(px, py);
        grid[(size_t)c.second * gx + c.first].push_back({px, py});
    }
    Rng rng(seed);
    Particle* out = frame->particles + frame->particle_count;
    uint32_t placed = 0;
    uint64_t attempts = 0, max_attempts = (uint64_t)count * 1000 + 1000;
    while (placed < count) {
        if (++attempts > max_attempts) return -1;  // box too crowded for min_dist
        double px = margin + rng.uniform() * (w - 2 * margin);
        double py = margin + rng.uniform() * (h - 2 * margin);
        auto c = cell_index(px, py);
        int cx = c.first, cy = c.second;
        bool ok = true;
        for (int dy = -1; dy <= 1 && ok; ++dy) {
            for (int dx = -1; dx <= 1 && ok; ++dx) {
                int x = cx + dx, y = cy + dy;
                if (x < 0 || y < 0 || x >= (int)gx || y >= (int)gy) continue;
                for (auto& q : grid[(size_t)y * gx + x]) {
                    double ddx = q.first - px, ddy = q.second - py;
                    if (ddx * ddx + ddy * ddy < min_dist * min_dist) {
                        ok = false;
                        break;
                    }
                }
            }
        }
        if (!ok) continue;
        grid[(size_t)cy * gx + cx].push_back({px, py});
        double vx, vy;
        random_vel(rng, v_min, v_max, &vx, &vy);
        *out++ = new_particle(meta, px, py, vx, vy, ty);
        ++placed;
    }
    frame->particle_count += count;
    return 0;
}

}  // extern "C"

// ---- presets (particle_io/src/presets.rs:84-154) -----------------------------------------------------------------

namespace {
struct Preset {  // presets.rs:84-90
    std::string name;
    float box_w = 0.f, box_h = 0.f;
    MiePotentialParams particles[2]{};
    std::vector<Particle> particles_list;
};

Preset preset_from_frame(const char* name, const FrameHeader* frame) {  // presets.rs:107-119
    Preset p;
    p.name = name ? name : "";
    p.box_w = frame->metadata.box_width;  // box_size() widens to f64 and from_frame narrows it back: the same f32
    p.box_h = frame->metadata.box_height;
    p.particles[0] = frame->metadata.particles[0];
    p.particles[1] = frame->metadata.particles[1];
    p.particles_list.assign(frame->particles, frame->particles + frame->particle_count);
    return p;
}
}  // namespace

struct PsimPresets {  // presets.rs:122-124
    std::vector<Preset> presets;
};

extern "C" {

PsimPresets* psim_presets_new(void) { return new PsimPresets; }

void psim_presets_destroy(PsimPresets* presets) { delete presets; }

size_t psim_presets_len(const PsimPresets* presets) { return presets ? presets->presets.size() : 0; }

long psim_presets_add_from_frame(PsimPresets* presets, const char* name, const FrameHeader* frame) {
    if (!presets || !frame) return -1;
    presets->presets.push_back(preset_from_frame(name, frame));
    return (long)presets->presets.size() - 1;
}

int psim_presets_change_from_frame(PsimPresets* presets, size_t index, const char* name, const FrameHeader* frame) {
    if (!presets || !frame) return -1;
    if (index >= presets->presets.size()) return 0;  // "why not": presets.rs:148-151 ignores it
    presets->presets[index] = preset_from_frame(name, frame);
    return 0;
}

long psim_presets_duplicate(PsimPresets* presets, size_t index, const char* new_name) {
    if (!presets || index >= presets->presets.size()) return -1;
    Preset copy = presets->presets[index];
    if (new_name) copy.name = new_name;
    presets->presets.push_back(std::move(copy));
    return (long)presets->presets.size() - 1;
}

int psim_presets_delete(PsimPresets* presets, size_t index) {
    if (!presets || index >= presets->presets.size()) return -1;
    presets->presets.erase(presets->presets.begin() + (long)index);
    return 0;
}

const char* psim_preset_name(const PsimPresets* presets, size_t index) {
    if (!presets || index >= presets->presets.size()) return nullptr;
    return presets->presets[index].name.c_str();
}

uint32_t psim_preset_particle_count(const PsimPresets* presets, size_t index) {
    if (!presets || index >= presets->presets.size()) return 0;
    return (uint32_t)presets->presets[index].particles_list.size();
}

int psim_preset_to_frame(const PsimPresets* presets, size_t index, FrameHeader* dst, uint32_t capacity) {
    if (!presets || !dst || index >= presets->presets.size()) return -1;
    const Preset& p = presets->presets[index];
    if (p.particles_list.size() > capacity) return -1;
    *dst = frame_header_init();  // Frame::new(): default metadata, no particles
    dst->metadata.box_width = p.box_w;
    dst->metadata.box_height = p.box_h;
    dst->metadata.particles[0] = p.particles[0];
    dst->metadata.particles[1] = p.particles[1];
    if (!p.particles_list.empty())
        std::memcpy(dst->particles, p.particles_list.data(), sizeof(Particle) * p.particles_list.size());
    dst->particle_count = (uint32_t)p.particles_list.size();
    return 0;
}

}  // extern "C"
