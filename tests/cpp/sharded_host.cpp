// TEST-ONLY: the multi-GPU entry points of the C ABI driven from a plain C++ host, the way a slam_demo consumer
// would use eight GPUs: one host thread per GPU, one handle each, NCCL inside liblocreg.so (no Python, no torch).
//   sharded_host [world]      world = number of GPUs to use (default: all visible, at most 8)
// Checks, against ONE handle doing the whole job on GPU 0:
//   locreg_relocalise_sharded  -> the same winner (global index, score, pose) on every rank
//   locreg_align_batch_sharded -> the same poses for ALL scans on every rank (ragged blocks: S % world != 0)
// Exit code 0: ok; 3: no CUDA device; 4: NCCL not available; anything else: failure.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <dlfcn.h>

#include "../../include/locreg.h"
#include "../../include/locreg_synth.h"

#define CHECK(expr)                                                                                    \
    do {                                                                                               \
        const int rc_ = (expr);                                                                        \
        if (rc_ != LOCREG_OK) {                                                                        \
            std::fprintf(stderr, "%s:%d %s -> %d (%s)\n", __FILE__, __LINE__, #expr, rc_, locreg_last_error()); \
            std::exit(rc_ == LOCREG_E_UNSUPPORTED ? 4 : 1);                                            \
        }                                                                                              \
    } while (0)

static int device_count() {
    void* cu = dlopen("libcuda.so.1", RTLD_NOW);
    if (!cu) return 0;
    auto init = reinterpret_cast<int (*)(unsigned)>(dlsym(cu, "cuInit"));
    auto count = reinterpret_cast<int (*)(int*)>(dlsym(cu, "cuDeviceGetCount"));
    int n = 0;
    if (!init || !count || init(0) != 0 || count(&n) != 0) return 0;
    return n;
}

struct Job {
    std::vector<float> map;            // float4
    std::vector<float> scans;          // float4, S scans concatenated
    std::vector<int64_t> offsets;      // S + 1
    std::vector<double> init;          // S * 7
    std::vector<double> hyp;           // n_hyp * 7 (hypotheses for scan 0)
    size_t n_map = 0, S = 0, n_hyp = 0;
};

static Job make_job(size_t S, size_t n_hyp) {
    Job j;
    synth_world* w = synth_world_create(80.0, 0x5EED0001ull);
    j.map.resize(4 * 120000);
    j.n_map = synth_world_sample_map(w, 120000, 0.2, 0.01, 0x5EED0001ull, j.map.data());
    std::vector<double> gt(S * 7);
    synth_world_poses(w, S, 0x5EED0003ull, gt.data());
    j.S = S;
    j.offsets.assign(1, 0);
    j.init.resize(S * 7);
    std::vector<float> one(4 * 16 * 500);
    for (size_t s = 0; s < S; ++s) {
        const size_t n = synth_world_scan(w, gt.data() + 7 * s, 16, 500, 0x5EED0002ull + s, one.data());
        j.scans.insert(j.scans.end(), one.begin(), one.begin() + 4 * n);
        j.offsets.push_back(j.offsets.back() + static_cast<int64_t>(n));
        synth_perturb_pose(gt.data() + 7 * s, 0x5EED0003ull + 31 * s, 0.3, 0.035, j.init.data() + 7 * s);
    }
    // hypotheses around scan 0's ground truth: an xy grid of 0.5 m pitch, the perturbed pose's rotation
    j.n_hyp = n_hyp;
    j.hyp.resize(n_hyp * 7);
    const int side = static_cast<int>(std::ceil(std::sqrt(static_cast<double>(n_hyp))));
    for (size_t i = 0; i < n_hyp; ++i) {
        std::memcpy(j.hyp.data() + 7 * i, j.init.data(), 7 * sizeof(double));
        j.hyp[7 * i + 4] = gt[4] + 0.5 * (static_cast<int>(i % side) - side / 2);
        j.hyp[7 * i + 5] = gt[5] + 0.5 * (static_cast<int>(i / side) - side / 2);
        j.hyp[7 * i + 6] = gt[6];
    }
    synth_world_destroy(w);
    return j;
}

static locreg_handle* make_handle(const Job& j, int device) {
    locreg_options o;
    CHECK(locreg_default_options(&o, LOCREG_ICP_P2PLANE));
    o.max_iteration = 6;
    o.eps = 0.0;
    locreg_handle* h = nullptr;
    CHECK(locreg_create(&o, device, &h));
    CHECK(locreg_set_target(h, j.map.data(), j.n_map, 16));
    return h;
}

struct RankOut {
    double best_pose[7];
    int64_t best_idx = -2;
    double best_score = 0;
    std::vector<double> poses;
    std::vector<locreg_result> results;
};

int main(int argc, char** argv) {
    const int n_dev = device_count();
    if (n_dev == 0) { std::printf("no CUDA device\n"); return 3; }
    int world = argc > 1 ? std::atoi(argv[1]) : (n_dev < 8 ? n_dev : 8);
    if (world < 1 || world > n_dev) { std::printf("world %d does not fit %d device(s)\n", world, n_dev); return 2; }
    const Job job = make_job(4 * static_cast<size_t>(world) + 1, 97);  // ragged on purpose

    // ---- the reference: one handle, the whole job
    locreg_handle* h0 = make_handle(job, 0);
    RankOut ref;
    CHECK(locreg_relocalise(h0, job.scans.data(), static_cast<size_t>(job.offsets[1]), 16, job.hyp.data(), job.n_hyp, ref.best_pose,
                            &ref.best_idx, &ref.best_score, nullptr, nullptr));
    ref.poses = job.init;
    ref.results.resize(job.S);
    CHECK(locreg_align_batch(h0, job.scans.data(), job.offsets.data(), 16, job.init.data(), job.S, ref.poses.data(), ref.results.data()));
    CHECK(locreg_destroy(h0));

    // ---- one thread per GPU
    unsigned char id[LOCREG_UNIQUE_ID_BYTES];
    CHECK(locreg_comm_unique_id(id));
    std::vector<RankOut> out(world);
    std::vector<std::thread> th;
    for (int r = 0; r < world; ++r) {
        th.emplace_back([&, r]() {
            locreg_handle* h = make_handle(job, r);
            CHECK(locreg_comm_init(h, id, r, world));
            int32_t rr = -1, ww = -1, ver = 0;
            CHECK(locreg_comm_info(h, &rr, &ww, &ver));
            if (rr != r || ww != world || ver <= 0) { std::fprintf(stderr, "comm_info mismatch\n"); std::exit(1); }
            RankOut& o = out[r];
            CHECK(locreg_relocalise_sharded(h, job.scans.data(), static_cast<size_t>(job.offsets[1]), 16, job.hyp.data(), job.n_hyp,
                                            o.best_pose, &o.best_idx, &o.best_score));
            size_t lo = 0, hi = 0;
            CHECK(locreg_shard_range(job.S, r, world, &lo, &hi));
            // this rank passes only its block: points, offsets relative to the block's first point, poses
            std::vector<int64_t> rel(hi - lo + 1);
            for (size_t s = lo; s <= hi; ++s) rel[s - lo] = job.offsets[s] - job.offsets[lo];
            o.poses.assign(job.S * 7, 0.0);
            std::memcpy(o.poses.data() + 7 * lo, job.init.data() + 7 * lo, (hi - lo) * 7 * sizeof(double));
            o.results.resize(job.S);
            CHECK(locreg_align_batch_sharded(h, job.scans.data() + 4 * job.offsets[lo], rel.data(), 16, job.init.data() + 7 * lo, hi - lo,
                                             job.S, o.poses.data(), o.results.data()));
            CHECK(locreg_comm_destroy(h));
            CHECK(locreg_destroy(h));
        });
    }
    for (auto& t : th) t.join();

    int bad = 0;
    for (int r = 0; r < world; ++r) {
        const RankOut& o = out[r];
        if (o.best_idx != ref.best_idx || o.best_score != ref.best_score || std::memcmp(o.best_pose, ref.best_pose, sizeof(ref.best_pose)) != 0) {
            std::fprintf(stderr, "rank %d: relocalisation winner %lld (%.9g) != %lld (%.9g)\n", r, static_cast<long long>(o.best_idx), o.best_score,
                         static_cast<long long>(ref.best_idx), ref.best_score);
            ++bad;
        }
        if (std::memcmp(o.poses.data(), ref.poses.data(), job.S * 7 * sizeof(double)) != 0) {
            std::fprintf(stderr, "rank %d: batch poses differ from the single-handle batch\n", r);
            ++bad;
        }
        for (size_t s = 0; s < job.S; ++s)
            if (o.results[s].iters != ref.results[s].iters || o.results[s].n_inlier != ref.results[s].n_inlier) {
                std::fprintf(stderr, "rank %d: result of scan %zu differs\n", r, s);
                ++bad;
                break;
            }
    }
    std::printf("sharded_host: world %d, %zu scans, %zu hypotheses, winner %lld score %.6g: %s\n", world, job.S, job.n_hyp,
                static_cast<long long>(ref.best_idx), ref.best_score, bad ? "MISMATCH" : "ok");
    return bad ? 1 : 0;
}
