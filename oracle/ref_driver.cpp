// TEST INFRASTRUCTURE (oracle).  C entry points around the reference's own paf_processor
// (include/openpose-plus.hpp:42-64, implemented by src/paf.cpp compiled unmodified from
// /root/reference).  Also exposes the real libstdc++ std::sort so the C emulation in opp_oracle.c
// can be checked against it.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include <openpose-plus.h>

#include "opp_oracle.h"

static_assert(sizeof(human_t) == sizeof(orc_human_t), "human_t layout");
static_assert(sizeof(ConnectionCandidate) == sizeof(orc_cand_t), "candidate layout");

namespace
{
// The reference prints three lines per frame (src/post-process.h:200-201, src/paf.cpp:251,290).
struct stdout_silencer {
    int saved;
    stdout_silencer()
    {
        std::fflush(stdout);
        saved = dup(1);
        const int nul = open("/dev/null", O_WRONLY);
        dup2(nul, 1);
        close(nul);
    }
    ~stdout_silencer()
    {
        std::fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};
}  // namespace

extern "C" {

void *ref_create(int feat_h, int feat_w, int out_h, int out_w, int ksize)
{
    return create_paf_processor(feat_h, feat_w, out_h, out_w, n_joins, n_connections, ksize);
}

void ref_destroy(void *p) { delete static_cast<paf_processor *>(p); }

// Runs one frame; copies up to cap humans into out; returns the number of humans found.
int ref_run(void *p, const float *conf, const float *paf, orc_human_t *out, int cap)
{
    stdout_silencer quiet;
    const std::vector<human_t> humans = (*static_cast<paf_processor *>(p))(conf, paf, false);
    const int n = (int)humans.size();
    for (int i = 0; i < n && i < cap; ++i) std::memcpy(&out[i], &humans[i], sizeof(human_t));
    return n;
}

// Times the reference on n_frames frames (conf [n,19,h,w], paf [n,38,h,w]) with n_threads
// independent paf_processor instances, each frame processed `repeat` times round-robin.
// Returns wall seconds; *humans_total receives the total count (keeps the work observable).
double ref_time_frames(int feat_h, int feat_w, int out_h, int out_w, int ksize, const float *conf,
                       const float *paf, int n_frames, int repeat, int n_threads, long *humans_total)
{
    stdout_silencer quiet;
    const size_t cs = (size_t)19 * feat_h * feat_w, ps = (size_t)38 * feat_h * feat_w;
    std::vector<std::unique_ptr<paf_processor>> procs;
    for (int t = 0; t < n_threads; ++t)
        procs.emplace_back(create_paf_processor(feat_h, feat_w, out_h, out_w, n_joins, n_connections, ksize));
    std::atomic<long> next(0), total(0);
    const long jobs = (long)n_frames * repeat;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> ths;
    for (int t = 0; t < n_threads; ++t)
        ths.emplace_back([&, t] {
            long local = 0;
            for (;;) {
                const long j = next.fetch_add(1);
                if (j >= jobs) break;
                const int f = (int)(j % n_frames);
                local += (long)(*procs[t])(conf + f * cs, paf + f * ps, false).size();
            }
            total += local;
        });
    for (auto &th : ths) th.join();
    const auto t1 = std::chrono::steady_clock::now();
    if (humans_total) *humans_total = total.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// One processor, one thread, one frame per call: the latency a caller of the reference's paf_processor sees
// (single-thread ms per frame, p50).  out_ms[iters] receives every call's wall time.
void ref_latency_frames(void *p, const float *conf, const float *paf, int n_frames, int feat_h, int feat_w, int iters, double *out_ms)
{
    stdout_silencer quiet;
    const size_t cs = (size_t)19 * feat_h * feat_w, ps = (size_t)38 * feat_h * feat_w;
    paf_processor *proc = static_cast<paf_processor *>(p);
    for (int i = 0; i < iters; ++i) {
        const int f = i % n_frames;
        const auto t0 = std::chrono::steady_clock::now();
        const size_t n = (*proc)(conf + f * cs, paf + f * ps, false).size();
        const auto t1 = std::chrono::steady_clock::now();
        out_ms[i] = std::chrono::duration<double, std::milli>(t1 - t0).count() + 0.0 * (double)n;
    }
}

// A pool of processors kept across calls (constructing one allocates ~100 MB of scratch, which
// is setup, not per-frame work): ref_pool_run times n_frames on the pool's threads.
struct ref_pool {
    int feat_h, feat_w;
    std::vector<std::unique_ptr<paf_processor>> procs;
};

void *ref_pool_create(int feat_h, int feat_w, int out_h, int out_w, int ksize, int n_threads)
{
    stdout_silencer quiet;
    ref_pool *p = new ref_pool{feat_h, feat_w, {}};
    for (int t = 0; t < n_threads; ++t)
        p->procs.emplace_back(create_paf_processor(feat_h, feat_w, out_h, out_w, n_joins, n_connections, ksize));
    return p;
}

void ref_pool_destroy(void *pp) { delete static_cast<ref_pool *>(pp); }

double ref_pool_run(void *pp, const float *conf, const float *paf, int n_frames, long *humans_total)
{
    stdout_silencer quiet;
    ref_pool *p = static_cast<ref_pool *>(pp);
    const size_t cs = (size_t)19 * p->feat_h * p->feat_w, ps = (size_t)38 * p->feat_h * p->feat_w;
    std::atomic<long> next(0), total(0);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> ths;
    for (size_t t = 0; t < p->procs.size(); ++t)
        ths.emplace_back([&, t] {
            long local = 0;
            for (;;) {
                const long j = next.fetch_add(1);
                if (j >= n_frames) break;
                local += (long)(*p->procs[t])(conf + j * cs, paf + j * ps, false).size();
            }
            total += local;
        });
    for (auto &th : ths) th.join();
    const auto t1 = std::chrono::steady_clock::now();
    if (humans_total) *humans_total = total.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// The real thing the reference calls at src/paf.cpp:151-152.
void ref_std_sort_desc(orc_cand_t *v, int n)
{
    ConnectionCandidate *c = reinterpret_cast<ConnectionCandidate *>(v);
    std::sort(c, c + n, std::greater<ConnectionCandidate>());
}
}
