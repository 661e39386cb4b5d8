// Deterministic synthetic world / LiDAR generator (host only, std-only C++17).
//
// The reference ships no data (its PCDs/bags are off-repo), so every config in
// BASELINE.json is defined on this generator: a piecewise-planar "city block" (ground plane,
// axis-aligned buildings on a jittered 40 m grid, perimeter wall), a map cloud sampled from its
// surfaces and analytic ray-cast N-beam scans.  All randomness is counter-based
// splitmix64 so results do not depend on thread count or evaluation order.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/locreg_synth.h"

namespace {

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
struct Rng {  // counter-based stream: value i of stream (seed, lane)
    uint64_t key;
    uint64_t ctr = 0;
    Rng(uint64_t seed, uint64_t lane) : key(splitmix64(seed ^ splitmix64(lane * 0xD1B54A32D192ED03ull + 1))) {}
    uint64_t next() { return splitmix64(key + (ctr++) * 0x9E3779B97F4A7C15ull); }
    double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
    double uniform(double a, double b) { return a + (b - a) * uniform(); }
    double normal() {  // Box-Muller, one value per call
        double u1 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        const double u2 = uniform();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

struct Box {
    double lo[3], hi[3];
};

struct World {
    double W;
    std::vector<Box> boxes;  // buildings first, then the 4 perimeter walls
    size_t n_buildings = 0;
};

void quat_rotate(const double* q, const double* p, double* out) {  // q = [x y z w]
    const double vx = q[0], vy = q[1], vz = q[2], w = q[3];
    double ux = vy * p[2] - vz * p[1], uy = vz * p[0] - vx * p[2], uz = vx * p[1] - vy * p[0];
    ux += ux; uy += uy; uz += uz;
    out[0] = p[0] + w * ux + (vy * uz - vz * uy);
    out[1] = p[1] + w * uy + (vz * ux - vx * uz);
    out[2] = p[2] + w * uz + (vx * uy - vy * ux);
}
void quat_mul(const double* a, const double* b, double* o) {
    const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    const double n = std::sqrt(w * w + x * x + y * y + z * z);
    o[0] = x / n; o[1] = y / n; o[2] = z / n; o[3] = w / n;
}
void quat_exp(const double* w, double* q) {
    const double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double s = th < 1e-12 ? 0.5 : std::sin(0.5 * th) / th;
    q[0] = s * w[0]; q[1] = s * w[1]; q[2] = s * w[2]; q[3] = std::cos(0.5 * th);
}

bool inside_any(const World& w, double x, double y, double margin) {
    for (size_t i = 0; i < w.n_buildings; ++i) {
        const Box& b = w.boxes[i];
        if (x > b.lo[0] - margin && x < b.hi[0] + margin && y > b.lo[1] - margin && y < b.hi[1] + margin) return true;
    }
    return false;
}

// nearest hit of ray o + t*d (t > 1e-6) with ground plane and all boxes; returns t or -1
double raycast(const World& w, const double* o, const double* d) {
    double best = 1e300;
    if (d[2] < 0) {
        const double t = -o[2] / d[2];
        if (t > 1e-6) {
            const double x = o[0] + t * d[0], y = o[1] + t * d[1];
            if (std::fabs(x) <= 0.5 * w.W && std::fabs(y) <= 0.5 * w.W) best = t;
        }
    }
    for (const Box& b : w.boxes) {
        double t0 = 1e-6, t1 = best;
        bool hit = true;
        for (int a = 0; a < 3 && hit; ++a) {
            if (std::fabs(d[a]) < 1e-300) {
                if (o[a] < b.lo[a] || o[a] > b.hi[a]) hit = false;
            } else {
                double ta = (b.lo[a] - o[a]) / d[a], tb = (b.hi[a] - o[a]) / d[a];
                if (ta > tb) std::swap(ta, tb);
                if (ta > t0) t0 = ta;
                if (tb < t1) t1 = tb;
                if (t0 > t1) hit = false;
            }
        }
        if (hit && t0 < best) best = t0;
    }
    return best < 1e299 ? best : -1.0;
}

size_t scan_one(const World& w, const double* pose7, int beams, int azimuth, uint64_t seed, float* out4) {
    size_t cnt = 0;
    const double o[3] = {pose7[4], pose7[5], pose7[6]};
    for (int b = 0; b < beams; ++b) {
        const double el = (-25.0 + (beams > 1 ? 40.0 * b / (beams - 1) : 0.0)) * 0.017453292519943295;
        const double ce = std::cos(el), se = std::sin(el);
        for (int a = 0; a < azimuth; ++a) {
            const double az = 6.283185307179586 * a / azimuth;
            const double ds[3] = {ce * std::cos(az), ce * std::sin(az), se};
            double dw[3];
            quat_rotate(pose7, ds, dw);
            const double t = raycast(w, o, dw);
            if (t < 0) continue;
            Rng r(seed, static_cast<uint64_t>(b) * azimuth + a);
            const double tn = t + 0.02 * r.normal();
            if (tn > 100.0 || tn < 0.5) continue;
            out4[cnt * 4 + 0] = static_cast<float>(tn * ds[0]);
            out4[cnt * 4 + 1] = static_cast<float>(tn * ds[1]);
            out4[cnt * 4 + 2] = static_cast<float>(tn * ds[2]);
            out4[cnt * 4 + 3] = static_cast<float>(b);
            ++cnt;
        }
    }
    return cnt;
}

}  // namespace

struct synth_world { World w; };

extern "C" {

synth_world* synth_world_create(double W, uint64_t seed) {
    auto* h = new synth_world;
    World& w = h->w;
    w.W = W;
    const int cells = static_cast<int>(W / 40.0);
    for (int iy = 0; iy < cells; ++iy)
        for (int ix = 0; ix < cells; ++ix) {
            Rng r(seed, static_cast<uint64_t>(iy) * cells + ix);
            const double cx = -0.5 * W + (ix + 0.5) * 40.0 + r.uniform(-4.0, 4.0);
            const double cy = -0.5 * W + (iy + 0.5) * 40.0 + r.uniform(-4.0, 4.0);
            const double sx = r.uniform(10.0, 30.0), sy = r.uniform(10.0, 30.0), hz = r.uniform(5.0, 20.0);
            Box b;
            b.lo[0] = cx - 0.5 * sx; b.hi[0] = cx + 0.5 * sx;
            b.lo[1] = cy - 0.5 * sy; b.hi[1] = cy + 0.5 * sy;
            b.lo[2] = 0.0; b.hi[2] = hz;
            w.boxes.push_back(b);
        }
    w.n_buildings = w.boxes.size();
    const double hw = 6.0, th = 0.5, half = 0.5 * W;
    w.boxes.push_back({{-half - th, -half - th, 0}, {-half, half + th, hw}});
    w.boxes.push_back({{half, -half - th, 0}, {half + th, half + th, hw}});
    w.boxes.push_back({{-half, -half - th, 0}, {half, -half, hw}});
    w.boxes.push_back({{-half, half, 0}, {half, half + th, hw}});
    return h;
}
void synth_world_destroy(synth_world* h) { delete h; }
size_t synth_world_num_boxes(const synth_world* h) { return h->w.boxes.size(); }

// Surface samples on a `pitch` lattice (+ N(0, sigma) jitter per axis), shuffled and truncated to n_map.
// Returns the number of points written (= n_map, or fewer if the world has fewer lattice sites).
size_t synth_world_sample_map(const synth_world* h, size_t n_map, double pitch, double sigma, uint64_t seed,
                              float* out4) {
    const World& w = h->w;
    std::vector<float> cand;  // x y z id
    auto push = [&](double x, double y, double z, float tag) {
        cand.push_back(static_cast<float>(x)); cand.push_back(static_cast<float>(y));
        cand.push_back(static_cast<float>(z)); cand.push_back(tag);
    };
    const double half = 0.5 * w.W;
    const long ng = static_cast<long>(std::floor(w.W / pitch));
    for (long iy = 0; iy <= ng; ++iy)
        for (long ix = 0; ix <= ng; ++ix) {
            const double x = -half + ix * pitch, y = -half + iy * pitch;
            if (inside_any(w, x, y, 0.0)) continue;
            push(x, y, 0.0, 0.f);
        }
    for (size_t bi = 0; bi < w.boxes.size(); ++bi) {
        const Box& b = w.boxes[bi];
        const bool wall = bi >= w.n_buildings;
        const float tag = wall ? 2.f : 1.f;
        const long nx = static_cast<long>(std::floor((b.hi[0] - b.lo[0]) / pitch));
        const long ny = static_cast<long>(std::floor((b.hi[1] - b.lo[1]) / pitch));
        const long nz = static_cast<long>(std::floor((b.hi[2] - b.lo[2]) / pitch));
        for (long iz = 1; iz <= nz; ++iz) {
            const double z = b.lo[2] + iz * pitch;
            for (long ix = 0; ix <= nx; ++ix) {
                const double x = b.lo[0] + ix * pitch;
                push(x, b.lo[1], z, tag); push(x, b.hi[1], z, tag);
            }
            for (long iy = 1; iy < ny; ++iy) {
                const double y = b.lo[1] + iy * pitch;
                push(b.lo[0], y, z, tag); push(b.hi[0], y, z, tag);
            }
        }
        if (!wall)
            for (long iy = 1; iy < ny; ++iy)
                for (long ix = 1; ix < nx; ++ix) push(b.lo[0] + ix * pitch, b.lo[1] + iy * pitch, b.hi[2], 3.f);
    }
    const size_t M = cand.size() / 4;
    std::vector<uint32_t> perm(M);
    for (size_t i = 0; i < M; ++i) perm[i] = static_cast<uint32_t>(i);
    Rng rs(seed, 0x5AFEull);
    const size_t take = std::min(n_map, M);
    for (size_t i = 0; i < take; ++i) {  // partial Fisher-Yates
        const size_t j = i + static_cast<size_t>(rs.next() % (M - i));
        std::swap(perm[i], perm[j]);
    }
    for (size_t i = 0; i < take; ++i) {
        const size_t c = perm[i];
        Rng r(seed, 0x100000000ull + c);
        out4[i * 4 + 0] = static_cast<float>(cand[c * 4 + 0] + sigma * r.normal());
        out4[i * 4 + 1] = static_cast<float>(cand[c * 4 + 1] + sigma * r.normal());
        out4[i * 4 + 2] = static_cast<float>(cand[c * 4 + 2] + sigma * r.normal());
        out4[i * 4 + 3] = cand[c * 4 + 3];
    }
    return take;
}

size_t synth_world_scan(const synth_world* h, const double* pose7, int beams, int azimuth, uint64_t seed, float* out4) {
    return scan_one(h->w, pose7, beams, azimuth, seed, out4);
}

// S scans; scan s is written at out4 + s*beams*azimuth*4 floats, counts[s] points valid. threads<=0: hardware.
void synth_world_scan_batch(const synth_world* h, const double* poses7, size_t S, int beams, int azimuth, uint64_t seed,
                            float* out4, int32_t* counts, int threads) {
    if (threads <= 0) threads = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
    threads = static_cast<int>(std::min<size_t>(threads, std::max<size_t>(S, 1)));
    const size_t cap = static_cast<size_t>(beams) * azimuth;
    auto work = [&](int tid) {
        for (size_t s = tid; s < S; s += threads)
            counts[s] = static_cast<int32_t>(
                scan_one(h->w, poses7 + s * 7, beams, azimuth, splitmix64(seed + s), out4 + s * cap * 4));
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
}

// Seeded random walk through free space at z = 1.8 m, yaw following the heading.
void synth_world_poses(const synth_world* h, size_t n, uint64_t seed, double* poses7) {
    const World& w = h->w;
    Rng r(seed, 0xB0B0ull);
    double x = 0, y = 0;
    for (int tries = 0; tries < 100000 && inside_any(w, x, y, 2.0); ++tries) {
        x = r.uniform(-0.3 * w.W, 0.3 * w.W); y = r.uniform(-0.3 * w.W, 0.3 * w.W);
    }
    double heading = r.uniform(0, 6.283185307179586);
    const double lim = 0.5 * w.W - 8.0;
    for (size_t i = 0; i < n; ++i) {
        const double rot[3] = {r.uniform(-0.01, 0.01), r.uniform(-0.01, 0.01), heading};
        double q[4];
        quat_exp(rot, q);  // small roll/pitch + yaw as one rotation vector
        double* p = poses7 + i * 7;
        p[0] = q[0]; p[1] = q[1]; p[2] = q[2]; p[3] = q[3];
        p[4] = x; p[5] = y; p[6] = 1.8;
        for (int tries = 0; tries < 1000; ++tries) {
            const double hd = heading + r.uniform(-0.4, 0.4);
            const double step = r.uniform(0.5, 1.5);
            const double nx = x + step * std::cos(hd), ny = y + step * std::sin(hd);
            if (std::fabs(nx) < lim && std::fabs(ny) < lim && !inside_any(w, nx, ny, 2.0)) {
                x = nx; y = ny; heading = hd;
                break;
            }
            heading += r.uniform(-1.5, 1.5);
        }
    }
}

// out = gt ∘ δ with δ.t ~ U[-max_trans, max_trans]^3 and δ.rot = exp(U[-max_rot_rad, max_rot_rad]^3)
void synth_perturb_pose(const double* gt7, uint64_t seed, double max_trans, double max_rot_rad, double* out7) {
    Rng r(seed, 0xDE17Aull);
    const double dt[3] = {r.uniform(-max_trans, max_trans), r.uniform(-max_trans, max_trans),
                          r.uniform(-max_trans, max_trans)};
    const double dw[3] = {r.uniform(-max_rot_rad, max_rot_rad), r.uniform(-max_rot_rad, max_rot_rad),
                          r.uniform(-max_rot_rad, max_rot_rad)};
    double dq[4], q[4], rt[3];
    quat_exp(dw, dq);
    quat_mul(gt7, dq, q);
    quat_rotate(gt7, dt, rt);
    out7[0] = q[0]; out7[1] = q[1]; out7[2] = q[2]; out7[3] = q[3];
    out7[4] = gt7[4] + rt[0]; out7[5] = gt7[5] + rt[1]; out7[6] = gt7[6] + rt[2];
}

}  // extern "C"
