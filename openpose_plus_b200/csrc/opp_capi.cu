// Host runtime behind the C-ABI of include/opp_b200.h: per-GPU handle, pipeline slots (own stream,
// device arena, pinned result buffers), and the stage sequencing that replaces
// paf_processor_impl::operator() (/root/reference src/paf.cpp:38-57).
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "opp_kernels.cuh"

static_assert(sizeof(opp_batch_t) == 88 && sizeof(opp_config_t) == 64, "C-ABI struct layout is part of the contract (ctypes mirrors it)");

namespace
{
thread_local std::string g_err;

void set_err(std::string *dst, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    if (dst) *dst = buf;
}

struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr; // materialising resize runs beside peak finding / grouping
    cudaEvent_t ev_start = nullptr, ev_done = nullptr, ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_in = nullptr; // recorded on the producer's stream (OPP_SYNC_STREAM)
    cudaEvent_t ev_h2d0 = nullptr, ev_h2d1 = nullptr; // OPP_H2D_SPLIT: the PAF copy runs on the side stream
    cudaEvent_t tr[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // OPP_TRACE: k1 start/end, k2 start/end, k3 end, copies end
    // device arena
    float *d_conf = nullptr, *d_paf = nullptr; // staged feature maps [B,19,h,w] / [B,38,h,w]
    float *d_hwc = nullptr;                    // staging for channels-last input
    float *d_conf_up = nullptr;                // only for the generic peak kernel
    int *d_counters = nullptr;                 // pk_cnt[B*18] | k2_done[B] | k3_done[B] | stats[B*4] | flags[B]
    int *d_pk_key = nullptr;
    opp_peak_t *d_peaks = nullptr;
    int *d_part_ofs = nullptr;
    opp_conn_t *d_conns = nullptr;
    int *d_n_conns = nullptr;
    float *d_cand = nullptr;
    opp_human_t *d_humans = nullptr;
    int *d_n_humans = nullptr;
    int *d_href_parts = nullptr;
    unsigned long long *d_times = nullptr; // OPP_TRACE: per (frame, limb) phase stamps of the limb kernel
    // pinned results
    opp_human_t *h_humans = nullptr;
    int *h_n_humans = nullptr, *h_flags = nullptr;
    float *h_in_conf = nullptr, *h_in_paf = nullptr; // pinned staging for a few pageable input frames (allocated on first use)
    float *d_in_conf = nullptr, *d_in_paf = nullptr; // their device-visible aliases
    int *h_done = nullptr;      // pinned completion word written by the assembly kernel (latency path)
    int done_tag = 0;           // value that word takes when the batch in flight is complete; 0 = wait on the event
    bool ms_pending = false;    // last_ms not read from the events yet
    // in-flight batch
    bool busy = false;
    bool direct_out = false; // results were written straight into the caller's pinned buffers
    int n_frames = 0;
    opp_batch_t batch{};
    float last_ms = 0.f;
    int ticket = -1;
};

} // namespace

struct opp_handle_s {
    opp_config_t cfg{};
    OppGeom g{};
    int device = 0;
    int max_smem = 0, sm_count = 0;
    bool fast_k2 = false;
    float taps[OPP_MAX_KSIZE + 1]{};
    int *d_xofs = nullptr, *d_yofs = nullptr;
    float *d_alpha = nullptr, *d_beta = nullptr;
    std::vector<Slot> slots;
    int next_slot = 0;
    int next_ticket = 0;
    int64_t launches = 0;
    std::string err;
    size_t counters_ints = 0;
    // K3 shared-memory plan
    K3Params k3_plan{};
    size_t k3_smem = 0;
    int force_tw = 0, force_th = 0;
    cudaStream_t timer_stream = nullptr;
    cudaEvent_t timer_t0 = nullptr, timer_t1 = nullptr;
    bool trace = false;
    bool fuse_resize = true;
    bool k2_skip = true;
    bool zero_copy_out = true;
    bool paf_early = true; // latency path: the limb kernel fetches its PAF tiles from pinned memory itself (OPP_NO_PAF_EARLY=1 disables)
    bool done_flag = true; // completion word in pinned memory on the latency path (OPP_NO_DONE_FLAG=1 disables)
    int tag_seq = 0;
    bool k2_store_low = true;     // full batches: fused peaks + resize kernel on the low-priority side stream, limb kernel on the high-priority one
    bool k2_skel_low = true;      // the same for the skeleton-only peak kernel
    bool generic_via_map = false; // non-integer scales: materialise the heat map first and let the generic kernel read it
    bool generic_rep = true; // integer scales outside the fast kernel's range: replication-aware generic kernel (OPP_NO_GENERIC_REP=1: via the materialised map)
    bool stage_pageable = true; // latency path for pageable inputs through pinned staging (OPP_NO_STAGE_PAGEABLE=1: cudaMemcpyAsync from pageable memory)
    bool pdl = true; // programmatic dependent launch on the latency path (OPP_NO_PDL=1 disables)
    int zero_copy_in_max = 0; // kernels reading pinned host maps in place: measured slower than staging them (kept for experiments)
    bool h2d_split = false;   // OPP_H2D_SPLIT=1: the two input tensors of a batch are copied on two streams (measured: no gain, see DESIGN)
    int ingest_max = 4;       // up to this many frames, pinned host maps are pulled in by one kernel instead of memset + 2 DMA copies
    cudaEvent_t trace_base = nullptr;
};

// Entry points run on the handle's device and give the caller its own current device back.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            set_err(&h->err, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return OPP_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

namespace
{
// cv::getGaussianKernel(k, sigma, CV_32F): taps in double, normalised, rounded to float.
void gauss_taps(int k, double sigma, float *taps)
{
    double t[OPP_MAX_KSIZE + 1], sum = 0;
    const double scale2x = -0.5 / (sigma * sigma);
    for (int i = 0; i < k; ++i) {
        const double x = i - (k - 1) * 0.5;
        t[i] = std::exp(scale2x * x * x);
        sum += t[i];
    }
    const double inv = 1.0 / sum;
    for (int i = 0; i < k; ++i) taps[i] = (float)(t[i] * inv);
}

// Python-path variant: _gauss_kernel(ksize, nsig) of openpose_plus/inference/post_process.py:13-17 is
// sqrt(outer(y, y)) / sum with y = diff(norm.cdf(linspace(-nsig - i/2, nsig + i/2, ksize + 1))), i = (2 nsig + 1) / ksize,
// i.e. the outer product of g = sqrt(y) / sum(sqrt(y)) with itself: a separable filter.  Taps in double, rounded to float.
void cdf_taps(int k, double nsig, float *taps)
{
    double e[OPP_MAX_KSIZE + 2], t[OPP_MAX_KSIZE + 1], sum = 0;
    const double interval = (2 * nsig + 1.) / k, lo = -nsig - interval / 2., hi = nsig + interval / 2.;
    for (int i = 0; i <= k; ++i) {
        const double x = i == k ? hi : lo + (hi - lo) / k * i; // numpy.linspace: start + step * i, end point exact
        e[i] = 0.5 * std::erfc(-x / std::sqrt(2.0));            // scipy.stats.norm.cdf
    }
    for (int i = 0; i < k; ++i) {
        t[i] = std::sqrt(e[i + 1] - e[i]);
        sum += t[i];
    }
    for (int i = 0; i < k; ++i) taps[i] = (float)(t[i] / sum);
}

// Area-mode coefficients of cv::resize(INTER_AREA) when up-sampling: s = floor(d*scale),
// f = (float)((d+1) - (s+1)*inv_scale), f = f <= 0 ? 0 : f - floor(f); x additionally clamps at the
// right edge.  Same statement as oracle/opp_oracle.c (validated against cv2 there).
int area_coeffs(int ssize, int dsize, bool clamp_edge, std::vector<int> &ofs, std::vector<float> &co)
{
    const double inv_scale = (double)dsize / ssize, scale = 1. / inv_scale;
    int dmax = dsize;
    ofs.resize(dsize), co.resize(2 * (size_t)dsize);
    for (int d = 0; d < dsize; ++d) {
        int s = (int)std::floor(d * scale);
        float f = (float)((d + 1) - (s + 1) * inv_scale);
        f = f <= 0 ? 0.f : f - (float)(int)std::floor(f);
        if (clamp_edge && s + 1 >= ssize) {
            if (d < dmax) dmax = d;
            if (s >= ssize - 1) f = 0, s = ssize - 1;
        }
        ofs[d] = s, co[2 * d] = 1.f - f, co[2 * d + 1] = f;
    }
    return dmax;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield" ::: "memory");
#endif
}

void plan_k3_smem(opp_handle_s *h)
{
    K3Params &p = h->k3_plan;
    const OppGeom &g = h->g;
    const int capP = h->cfg.max_peaks_per_part, capC = h->cfg.max_cands_per_limb, capH = h->cfg.max_humans;
    const size_t budget = (size_t)h->max_smem;
    // scoring / matching phase
    size_t off = 0;
    const size_t paf_bytes = 2 * (size_t)g.h * g.w * sizeof(float);
    const size_t fixed = 2 * (size_t)capP * sizeof(int2) + align_up(2 * (size_t)capP, 16) + 64;
    const size_t cand_bytes = 2 * (size_t)capC * 12;
    p.cand_in_smem = (fixed + cand_bytes) <= budget / 2;
    // the second candidate buffer and the PAF tile share their bytes (below): the tile costs only what it exceeds that buffer by
    {
        const size_t c1 = p.cand_in_smem ? cand_bytes / 2 : 0, extra = paf_bytes > c1 ? paf_bytes - c1 : 0;
        p.paf_in_smem = (fixed + (p.cand_in_smem ? cand_bytes : 0) + extra + 12 * 1024) <= budget / 2; // + survivors, steps, weak map
    }
    // The PAF tile is dead once the limb's pairs are scored and the second candidate buffer (the sort's output) is
    // not written before that: the two share their bytes.
    p.off_paf = (int)off;
    { // cand0 [capC], cand1 [capC]: cand1 starts where the PAF tile starts when both are in shared memory
        const size_t c0 = p.cand_in_smem ? align_up((size_t)capC * 12, 16) : 0, c1 = c0;
        const size_t tile = p.paf_in_smem ? align_up(paf_bytes, 16) : 0;
        p.off_cand1 = (int)off;                      // cand1 first (aliases the tile) ...
        off += c1 > tile ? c1 : tile;
        p.off_cand = (int)off, off += c0;            // ... then cand0, which scoring fills while the tile is read
    }
    p.off_pk = (int)off, off += 2 * (size_t)capP * sizeof(int2);
    p.off_used = (int)off, off += align_up(2 * (size_t)capP, 16);
    p.off_misc = (int)off, off += 64;
    p.off_keys = (int)off, off += align_up(2 * (size_t)capP * sizeof(int), 16); // unordered keys of the limb's two parts
    // survivors of the quick PAF test, later the matching's per-candidate state (one byte each): at least two rounds of
    // the CTA's threads, and room for capC state bytes where that is affordable
    size_t surv_ints = 1024;
    if (p.cand_in_smem && (size_t)capC > surv_ints * 4) surv_ints = ((size_t)capC + 3) / 4;
    p.off_surv = (int)off, p.surv_cap = (int)surv_ints, off += align_up(surv_ints * sizeof(int), 16);
    p.cand_unordered = p.cand_in_smem && capC <= 4096;
    const size_t n_steps = (size_t)std::max(g.H, g.W);
    p.steps_in_smem = n_steps <= 1024;
    p.off_steps = (int)off;
    if (p.steps_in_smem) off += align_up(n_steps * sizeof(float), 16);
    p.weak_in_smem = p.cand_unordered && ((size_t)g.h * g.w + 31) / 32 * sizeof(unsigned) <= 16 * 1024; // built from wherever the PAF planes are
    p.off_weak = (int)off;
    if (p.weak_in_smem) off += align_up((((size_t)g.h * g.w + 31) / 32) * sizeof(unsigned), 16);
    const size_t phase1 = off;
    // assembly phase (re-uses the same bytes): partial humans, survivors, connections, peak x/y/score
    off = 0;
    p.off_href = 0, off += align_up((size_t)capH * 21 * sizeof(int), 16);
    p.off_keep = (int)off, off += align_up((size_t)capH * sizeof(int), 16);
    // Staging areas are sized for crowded REAL frames (40 people: ~800 peaks, ~850 connections), not for the capacities'
    // worst case (19 x capP connections): a frame beyond them takes the limb-by-limb / global-memory forms of the same
    // code (decided per frame in assemble_frame).  The footprint of every limb CTA is what limits CTAs per SM.
    const size_t conn_cap = std::max<size_t>(capP, std::min<size_t>((size_t)OPP_N_PAIRS * capP, 1024));
    const size_t pk_cap = std::min<size_t>((size_t)OPP_N_PARTS * capP, 1024);
    const size_t conn_all = conn_cap * sizeof(opp_conn_t), conn_one = (size_t)capP * sizeof(opp_conn_t);
    const size_t pk_bytes = pk_cap * sizeof(int2); // x | y << 16, score
    p.conns_in_smem = off + conn_all + pk_bytes <= (size_t)(budget * 0.45);
    p.conn_cap = p.conns_in_smem ? (int)conn_cap : capP;
    p.off_conn = (int)off, off += align_up(p.conns_in_smem ? conn_all : conn_one, 16);
    p.score_in_smem = off + pk_bytes <= budget / 2;
    p.pk_cap = p.score_in_smem ? (int)pk_cap : 0;
    p.off_score = (int)off;
    if (p.score_in_smem) off += pk_bytes;
    // forest tables: s_c1 [17][capP] u16, s_in bitmap over 18 capP peak ids, s_lmb [conn_cap] u8
    const size_t owner_bytes = align_up((size_t)17 * capP * 2, 16) + align_up(((size_t)OPP_N_PARTS * capP + 31) / 32 * 4, 16) + align_up(conn_cap, 16);
    p.owner_in_smem = p.conns_in_smem && p.score_in_smem && off + owner_bytes <= budget / 2;
    p.off_owner = (int)off;
    if (p.owner_in_smem) off += owner_bytes;
    h->k3_smem = phase1 > off ? phase1 : off;
    if (const char *e = getenv("OPP_K3_SMEM_MIN")) { // experiments: pad the footprint (fewer CTAs per SM)
        const size_t m = (size_t)atoi(e);
        if (m > h->k3_smem && m <= budget) h->k3_smem = m;
    }
}

int free_slot(opp_handle_s *h, Slot &s)
{
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.d_conf), cudaFree(s.d_paf), cudaFree(s.d_hwc), cudaFree(s.d_conf_up), cudaFree(s.d_counters);
    cudaFree(s.d_pk_key), cudaFree(s.d_peaks), cudaFree(s.d_part_ofs), cudaFree(s.d_conns), cudaFree(s.d_n_conns);
    cudaFree(s.d_cand), cudaFree(s.d_humans), cudaFree(s.d_n_humans), cudaFree(s.d_href_parts), cudaFree(s.d_times);
    cudaFreeHost(s.h_humans), cudaFreeHost(s.h_n_humans), cudaFreeHost(s.h_flags), cudaFreeHost(s.h_done);
    cudaFreeHost(s.h_in_conf), cudaFreeHost(s.h_in_paf);
    if (s.ev_start) cudaEventDestroy(s.ev_start);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.ev_fork) cudaEventDestroy(s.ev_fork);
    if (s.ev_join) cudaEventDestroy(s.ev_join);
    if (s.ev_in) cudaEventDestroy(s.ev_in);
    if (s.ev_h2d0) cudaEventDestroy(s.ev_h2d0);
    if (s.ev_h2d1) cudaEventDestroy(s.ev_h2d1);
    if (s.side) cudaStreamDestroy(s.side);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();
    (void)h;
    return 0;
}

int alloc_slot(opp_handle_s *h, Slot &s)
{
    const opp_config_t &c = h->cfg;
    const size_t B = c.max_batch, hw = (size_t)c.feat_h * c.feat_w;
    const int capP = c.max_peaks_per_part, capC = c.max_cands_per_limb, capH = c.max_humans;
    // The store-bound resize runs on the low-priority side stream: its tens of thousands of short CTAs
    // would otherwise occupy every SM ahead of the compute-bound peak / limb kernels it overlaps with.
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (getenv("OPP_NO_PRIORITY")) prio_hi = prio_lo;
    CU(cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, prio_hi));
    CU(cudaStreamCreateWithPriority(&s.side, cudaStreamNonBlocking, prio_lo));
    CU(cudaEventCreate(&s.ev_start));
    CU(cudaEventCreate(&s.ev_done));
    CU(cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s.ev_h2d0, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s.ev_h2d1, cudaEventDisableTiming));
    if (h->trace) {
        for (auto &e : s.tr) CU(cudaEventCreate(&e));
        CU(cudaMalloc(&s.d_times, (B * OPP_N_PAIRS * 12 + 4096 * 8) * sizeof(unsigned long long)));
        CU(cudaMemset(s.d_times, 0, (B * OPP_N_PAIRS * 12 + 4096 * 8) * sizeof(unsigned long long)));
    }
    CU(cudaMalloc(&s.d_conf, B * OPP_N_HEAT * hw * sizeof(float)));
    CU(cudaMalloc(&s.d_paf, B * OPP_N_PAF * hw * sizeof(float)));
    CU(cudaMalloc(&s.d_counters, h->counters_ints * sizeof(int)));
    CU(cudaMemset(s.d_counters, 0, h->counters_ints * sizeof(int)));
    CU(cudaMalloc(&s.d_pk_key, B * OPP_N_PARTS * capP * sizeof(int)));
    CU(cudaMalloc(&s.d_peaks, B * OPP_N_PARTS * capP * sizeof(opp_peak_t)));
    CU(cudaMalloc(&s.d_part_ofs, B * (OPP_N_PARTS + 1) * sizeof(int)));
    CU(cudaMalloc(&s.d_conns, B * OPP_N_PAIRS * capP * sizeof(opp_conn_t)));
    CU(cudaMalloc(&s.d_n_conns, B * OPP_N_PAIRS * sizeof(int)));
    if (!h->k3_plan.cand_in_smem) CU(cudaMalloc(&s.d_cand, B * OPP_N_PAIRS * 2 * (size_t)capC * 12));
    CU(cudaMalloc(&s.d_humans, B * capH * sizeof(opp_human_t)));
    CU(cudaMalloc(&s.d_n_humans, B * sizeof(int)));
    CU(cudaMalloc(&s.d_href_parts, B * capH * OPP_N_PARTS * sizeof(int)));
    CU(cudaMemset(s.d_n_conns, 0, B * OPP_N_PAIRS * sizeof(int)));
    CU(cudaMemset(s.d_part_ofs, 0, B * (OPP_N_PARTS + 1) * sizeof(int)));
    CU(cudaMemset(s.d_n_humans, 0, B * sizeof(int)));
    CU(cudaMallocHost(&s.h_humans, B * capH * sizeof(opp_human_t)));
    CU(cudaMallocHost(&s.h_n_humans, B * sizeof(int)));
    CU(cudaMallocHost(&s.h_flags, B * sizeof(int)));
    CU(cudaMallocHost(&s.h_done, sizeof(int)));
    *s.h_done = 0;
    return OPP_OK;
}

int *cnt_pk(opp_handle_s *, Slot &s) { return s.d_counters; }
int *cnt_k2(opp_handle_s *h, Slot &s) { return s.d_counters + (size_t)h->cfg.max_batch * OPP_N_PARTS; }
int *cnt_k3(opp_handle_s *h, Slot &s) { return cnt_k2(h, s) + h->cfg.max_batch; }
int *cnt_stats(opp_handle_s *h, Slot &s) { return cnt_k3(h, s) + h->cfg.max_batch; }
int *cnt_flags(opp_handle_s *h, Slot &s) { return cnt_stats(h, s) + (size_t)h->cfg.max_batch * 4; }
int *cnt_batch_done(opp_handle_s *h, Slot &s) { return cnt_flags(h, s) + h->cfg.max_batch; }

// Pinned ranges handed out by opp_host_alloc: looked up without a driver call (cudaPointerGetAttributes costs about a
// microsecond per pointer, and a batch carries up to five of them ahead of its first kernel launch).
struct PinnedRange {
    uintptr_t base, size, dev;
};
std::mutex g_pin_mu;
std::vector<PinnedRange> g_pins;

// Device-visible alias of a pinned (cudaMallocHost / cudaHostRegister) host pointer, or null.
void *mapped_host(const void *p)
{
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        std::lock_guard<std::mutex> lk(g_pin_mu);
        for (const PinnedRange &r : g_pins)
            if (a >= r.base && a - r.base < r.size) return reinterpret_cast<void *>(r.dev + (a - r.base));
    }
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// Tile plan of the integer-scale peak kernel.  Limits of the kernel itself (k2_peaks_fast / launch_k2_fast_t), with nb =
// 1 (R <= S) or 2 (S < R <= 2S) cells of filter reach: a column strip holds at most 64 - 2 nb feature columns (+2 nb halo
// columns in the 64-bit activity masks) and at most 7 warps x 62 decided image columns; a row tile at most 62 - 2 nb
// feature rows (+2 + 2 nb halo rows).  Returns false when no valid plan exists for this
// geometry (opp_create then selects the replication-aware generic kernel instead of failing at launch time).
bool choose_k2_tiles(opp_handle_s *h, int n_frames, bool store, int &tw, int &th)
{
    const OppGeom &g = h->g;
    const int S = g.S > 0 ? g.S : 1;
    const int nb = g.R > S ? 2 : 1; // neighbour depth of the filter window (cells)
    int tw_max = (62 * K2_FAST_MAX_GROUPS) / S;
    if (tw_max > 64 - 2 * nb) tw_max = 64 - 2 * nb;
    if (tw_max < 1) return false;
    const int nxs = (g.w + tw_max - 1) / tw_max;
    tw = (g.w + nxs - 1) / nxs;
    // three resident CTAs per SM (the kernel is compiled for that): each must stay under ~72 KB, and
    // row tiles of about 16 feature rows keep the last wave short
    // (measured at 368x432: ~8 rows when the kernel also streams the up-sampled maps - short tiles
    // keep the store flow even - and ~23 rows in skeleton-only mode, where per-CTA fixed costs dominate)
    const int target = store ? 8 : 23;
    const int nys = (g.h + target - 1) / target;
    th = (g.h + nys - 1) / nys;
    while ((k2_fast_smem_bytes(g, tw, th) > (size_t)72 * 1024 || th > 62 - 2 * nb) && th > 4) th = (th + 1) / 2;
    // small batches: split rows too until the grid covers the chip about twice
    const long want = 2L * h->sm_count;
    while ((long)n_frames * OPP_N_PARTS * ((g.w + tw - 1) / tw) * ((g.h + th - 1) / th) < want && th > 6) th = (th + 1) / 2;
    if (h->force_tw > 0) tw = h->force_tw;
    if (h->force_th > 0) th = h->force_th;
    const int groups = (S * tw + 61) / 62;
    return tw >= 1 && th >= 1 && tw + 2 * nb <= 64 && th + 2 + 2 * nb <= 64 && groups >= 1 && groups <= K2_FAST_MAX_GROUPS &&
           k2_fast_smem_bytes(g, tw, th) <= (size_t)h->max_smem - 1024;
}

} // namespace

extern "C" {

const char *opp_version(void) { return "openpose-plus-b200 0.1 (sm_100a)"; }

const char *opp_last_error(opp_handle_t h) { return h ? h->err.c_str() : g_err.c_str(); }

void opp_config_default(opp_config_t *cfg, int feat_h, int feat_w, int out_h, int out_w, int gauss_kernel_size)
{
    std::memset(cfg, 0, sizeof *cfg);
    cfg->feat_h = feat_h, cfg->feat_w = feat_w, cfg->out_h = out_h, cfg->out_w = out_w;
    cfg->n_joins = OPP_N_HEAT, cfg->n_connections = OPP_N_PAIRS;
    cfg->gauss_kernel_size = gauss_kernel_size;
    cfg->max_batch = 64;
    cfg->device = -1;
    cfg->max_peaks_per_part = 128;
    cfg->max_cands_per_limb = 1024;
    cfg->max_humans = 128;
    cfg->n_slots = 3;
}

static void remember_pinned(void *p, size_t bytes)
{
    void *d = nullptr;
    if (cudaHostGetDevicePointer(&d, p, 0) == cudaSuccess && d) {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        g_pins.push_back({reinterpret_cast<uintptr_t>(p), bytes, reinterpret_cast<uintptr_t>(d)});
    } else {
        cudaGetLastError();
    }
}

static void forget_pinned(void *p)
{
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (size_t i = 0; i < g_pins.size(); ++i)
        if (g_pins[i].base == reinterpret_cast<uintptr_t>(p)) {
            g_pins.erase(g_pins.begin() + i);
            break;
        }
}

void *opp_host_alloc_ex(size_t bytes, int flags)
{
    void *p = nullptr;
    // portable: usable from every device of the process (one handle per GPU, process_stream_multi)
    unsigned f = cudaHostAllocPortable | cudaHostAllocMapped;
    if (flags & OPP_HOST_WRITE_COMBINED) f |= cudaHostAllocWriteCombined;
    if (cudaHostAlloc(&p, bytes, f) != cudaSuccess) {
        set_err(nullptr, "cudaHostAlloc(%zu, %u) failed", bytes, f);
        cudaGetLastError();
        return nullptr;
    }
    remember_pinned(p, bytes);
    return p;
}

void *opp_host_alloc(size_t bytes) { return opp_host_alloc_ex(bytes, OPP_HOST_DEFAULT); }

int opp_host_register(void *p, size_t bytes)
{
    if (!p || !bytes) return OPP_ERR_INVALID;
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
        set_err(nullptr, "cudaHostRegister(%p, %zu) failed: %s", p, bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return OPP_ERR_CUDA;
    }
    remember_pinned(p, bytes);
    return OPP_OK;
}

int opp_host_unregister(void *p)
{
    if (!p) return OPP_ERR_INVALID;
    forget_pinned(p);
    if (cudaHostUnregister(p) != cudaSuccess) {
        cudaGetLastError();
        return OPP_ERR_CUDA;
    }
    return OPP_OK;
}

void opp_host_free(void *p)
{
    if (!p) return;
    forget_pinned(p);
    cudaFreeHost(p);
}

void opp_destroy(opp_handle_t h)
{
    if (!h) return;
    DeviceGuard guard_(h->device);
    for (auto &s : h->slots) free_slot(h, s);
    if (h->timer_t0) cudaEventDestroy(h->timer_t0);
    if (h->timer_t1) cudaEventDestroy(h->timer_t1);
    if (h->timer_stream) cudaStreamDestroy(h->timer_stream);
    cudaFree(h->d_xofs), cudaFree(h->d_yofs), cudaFree(h->d_alpha), cudaFree(h->d_beta);
    delete h;
}

int opp_create(const opp_config_t *cfg, opp_handle_t *out)
{
    if (!cfg || !out) {
        set_err(nullptr, "opp_create: null argument");
        return OPP_ERR_INVALID;
    }
    *out = nullptr;
    opp_config_t c = *cfg;
    if (c.max_batch <= 0) c.max_batch = 64;
    if (c.max_peaks_per_part <= 0) c.max_peaks_per_part = 128;
    if (c.max_cands_per_limb <= 0) c.max_cands_per_limb = 1024;
    if (c.max_humans <= 0) c.max_humans = 128;
    if (c.n_slots <= 0) c.n_slots = 3;
    const int k = c.gauss_kernel_size;
    if (c.variant != OPP_VARIANT_CPP && c.variant != OPP_VARIANT_PYTHON) {
        set_err(nullptr, "opp_create: variant must be OPP_VARIANT_CPP or OPP_VARIANT_PYTHON");
        return OPP_ERR_INVALID;
    }
    if (c.n_joins != OPP_N_HEAT || c.n_connections != OPP_N_PAIRS) {
        set_err(nullptr, "opp_create: n_joins and n_connections must be 19 (include/openpose-plus.hpp:63)");
        return OPP_ERR_INVALID;
    }
    if (c.out_h > 32767 || c.out_w > 32767) {
        set_err(nullptr, "opp_create: output size is limited to 32767 x 32767");
        return OPP_ERR_INVALID;
    }
    if (c.feat_h < 2 || c.feat_w < 2 || c.out_h < c.feat_h || c.out_w < c.feat_w) {
        set_err(nullptr, "opp_create: output size must be >= feature size >= 2 (INTER_AREA up-sampling only)");
        return OPP_ERR_INVALID;
    }
    if (k < 1 || (k & 1) == 0 || k > OPP_MAX_KSIZE || k / 2 >= c.out_h - 1 || k / 2 >= c.out_w - 1) {
        set_err(nullptr, "opp_create: gauss_kernel_size must be odd, 1..%d and smaller than the image", OPP_MAX_KSIZE);
        return OPP_ERR_INVALID;
    }
    if ((long)OPP_N_PARTS * c.max_peaks_per_part * 4 > 200 * 1024 || c.max_humans > 8192) {
        set_err(nullptr, "opp_create: capacities too large");
        return OPP_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_err(nullptr, "opp_create: no CUDA device (this path has no CPU fallback)");
        return OPP_ERR_NO_DEVICE;
    }
    opp_handle_s *h = new opp_handle_s();
    h->cfg = c;
    int rc = OPP_OK;
    int caller_dev = -1;
    cudaGetDevice(&caller_dev);
    auto body = [&]() -> int {
        if (c.device >= 0)
            CU(cudaSetDevice(c.device));
        CU(cudaGetDevice(&h->device));
        h->cfg.device = h->device;
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, h->device));
        if (prop.major < 10) {
            set_err(&h->err, "opp_create: device %d is sm_%d%d; this library is built for sm_100a only", h->device, prop.major, prop.minor);
            return OPP_ERR_NO_DEVICE;
        }
        h->max_smem = (int)prop.sharedMemPerBlockOptin;
        h->sm_count = prop.multiProcessorCount;
        CU(opp_kernels_init(h->max_smem));
        OppGeom &g = h->g;
        g.h = c.feat_h, g.w = c.feat_w, g.H = c.out_h, g.W = c.out_w, g.K = k, g.R = k / 2;
        g.S = (c.out_h % c.feat_h == 0 && c.out_w % c.feat_w == 0 && c.out_h / c.feat_h == c.out_w / c.feat_w) ? c.out_h / c.feat_h : 0;
        if (c.variant == OPP_VARIANT_PYTHON)
            cdf_taps(k, 3.0, h->taps); // nsig fixed by the reference, post_process.py:22
        else
            gauss_taps(k, 3.0, h->taps); // sigma fixed by the reference, src/post-process.h:54
        std::vector<int> xo, yo;
        std::vector<float> al, be;
        g.xmax = area_coeffs(g.w, g.W, true, xo, al);
        area_coeffs(g.h, g.H, false, yo, be);
        CU(cudaMalloc(&h->d_xofs, xo.size() * sizeof(int)));
        CU(cudaMalloc(&h->d_yofs, yo.size() * sizeof(int)));
        CU(cudaMalloc(&h->d_alpha, al.size() * sizeof(float)));
        CU(cudaMalloc(&h->d_beta, be.size() * sizeof(float)));
        CU(cudaMemcpy(h->d_xofs, xo.data(), xo.size() * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->d_yofs, yo.data(), yo.size() * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->d_alpha, al.data(), al.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->d_beta, be.data(), be.size() * sizeof(float), cudaMemcpyHostToDevice));
        g.xofs = h->d_xofs, g.yofs = h->d_yofs, g.alpha = h->d_alpha, g.beta = h->d_beta;
        // cv::GaussianBlur's REFLECT_101 border for every k the fast kernel covers; the Python graph's zero border at its k = 25
        h->fast_k2 = k2_fast_supported(g, c.variant == OPP_VARIANT_PYTHON);
        if (const char *e = getenv("OPP_FORCE_GENERIC")) {
            if (atoi(e)) h->fast_k2 = false;
        }
        if (const char *e = getenv("OPP_K2_TW")) h->force_tw = atoi(e);
        if (const char *e = getenv("OPP_K2_TH")) h->force_th = atoi(e);
        if (h->fast_k2) { // every batch size and both modes must have a launchable tile plan, else the generic kernel takes over
            int tw_ = 0, th_ = 0;
            for (int store = 0; store < 2 && h->fast_k2; ++store)
                for (int nf = 1; nf <= c.max_batch && h->fast_k2; nf = nf < c.max_batch && 2 * nf > c.max_batch ? c.max_batch : 2 * nf)
                    if (!choose_k2_tiles(h, nf, store != 0, tw_, th_)) h->fast_k2 = false;
        }
        plan_k3_smem(h);
        if (h->k3_smem > (size_t)h->max_smem) {
            set_err(&h->err, "opp_create: capacities need %zu bytes of shared memory (> %d)", h->k3_smem, h->max_smem);
            return OPP_ERR_INVALID;
        }
        h->counters_ints = (size_t)c.max_batch * (OPP_N_PARTS + 1 + 1 + 4 + 1) + 1; // + frames assembled in this batch
        h->trace = getenv("OPP_TRACE") != nullptr;
        h->fuse_resize = getenv("OPP_NO_FUSE") == nullptr;
        h->k2_skip = getenv("OPP_K2_NOSKIP") == nullptr;
        h->zero_copy_out = getenv("OPP_NO_ZEROCOPY_OUT") == nullptr;
        h->pdl = getenv("OPP_NO_PDL") == nullptr;
        h->stage_pageable = getenv("OPP_NO_STAGE_PAGEABLE") == nullptr;
        h->generic_rep = getenv("OPP_NO_GENERIC_REP") == nullptr;
        if (const char *e = getenv("OPP_K2_LOW")) h->k2_store_low = h->k2_skel_low = atoi(e) != 0;
        if (const char *e = getenv("OPP_K2_STORE_LOW")) h->k2_store_low = atoi(e) != 0;
        if (const char *e = getenv("OPP_K2_SKEL_LOW")) h->k2_skel_low = atoi(e) != 0;
        h->generic_via_map = getenv("OPP_GENERIC_VIA_MAP") != nullptr && atoi(getenv("OPP_GENERIC_VIA_MAP")) != 0;
        h->done_flag = getenv("OPP_NO_DONE_FLAG") == nullptr;
        h->paf_early = getenv("OPP_NO_PAF_EARLY") == nullptr;
        if (const char *e = getenv("OPP_ZC_IN_MAX")) h->zero_copy_in_max = atoi(e);
        if (const char *e = getenv("OPP_INGEST_MAX")) h->ingest_max = atoi(e);
        if (h->ingest_max > c.max_batch) h->ingest_max = c.max_batch;
        h->h2d_split = getenv("OPP_H2D_SPLIT") != nullptr;
        CU(cudaStreamCreateWithFlags(&h->timer_stream, cudaStreamNonBlocking));
        if (h->trace) {
            CU(cudaEventCreate(&h->trace_base));
            CU(cudaEventRecord(h->trace_base, h->timer_stream));
        }
        CU(cudaEventCreate(&h->timer_t0));
        CU(cudaEventCreate(&h->timer_t1));
        h->slots.resize(c.n_slots);
        for (auto &s : h->slots) {
            int r = alloc_slot(h, s);
            if (r != OPP_OK) return r;
        }
        return OPP_OK;
    };
    rc = body();
    if (rc != OPP_OK) {
        g_err = h->err;
        opp_destroy(h);
        if (caller_dev >= 0) cudaSetDevice(caller_dev);
        return rc;
    }
    if (caller_dev >= 0) cudaSetDevice(caller_dev);
    *out = h;
    return OPP_OK;
}

static int fill_k2(opp_handle_s *h, Slot &s, const float *conf, const float *conf_up, int n, bool store, K2Params &k2);

static int enqueue(opp_handle_s *h, Slot &s, const opp_batch_t &b)
{
    const opp_config_t &c = h->cfg;
    const OppGeom &g = h->g;
    const int n = b.n_frames;
    const size_t hw = (size_t)g.h * g.w, HW = (size_t)g.H * g.W;
    cudaStream_t st = s.stream;
    // Device-resident maps written by a producer on its own stream (the CNN runner, src/uff-runner.cpp:199-205): the
    // slot's stream waits for it on the device; the host is never synchronised.
    if (b.in_sync == OPP_SYNC_STREAM) {
        CU(cudaEventRecord(s.ev_in, (cudaStream_t)b.in_sync_obj));
        CU(cudaStreamWaitEvent(st, s.ev_in, 0));
    } else if (b.in_sync == OPP_SYNC_EVENT) {
        CU(cudaStreamWaitEvent(st, (cudaEvent_t)b.in_sync_obj, 0));
    }
    CU(cudaEventRecord(s.ev_start, st));
    // A few frames in host memory take the latency path.  Pinned buffers are read in place; PAGEABLE ones (what a
    // caller of the reference's paf_processor passes) are first copied by this thread into the slot's pinned staging:
    // the heat maps now, the PAFs - which only the limb kernel reads - after ingest and peak kernel have been launched,
    // so that two thirds of the copy overlap with GPU work instead of preceding it.
    const bool few_host = b.in_mem == OPP_MEM_HOST && b.in_layout == OPP_LAYOUT_CHW && n <= h->ingest_max;
    const float *m_conf = few_host ? (const float *)mapped_host(b.conf) : nullptr, *m_paf = few_host ? (const float *)mapped_host(b.paf) : nullptr;
    bool late_paf_copy = false;
    if (few_host && !(m_conf && m_paf) && h->stage_pageable) {
        if (!s.h_in_conf) {
            CU(cudaMallocHost(&s.h_in_conf, (size_t)h->ingest_max * OPP_N_HEAT * hw * sizeof(float)));
            CU(cudaMallocHost(&s.h_in_paf, (size_t)h->ingest_max * OPP_N_PAF * hw * sizeof(float)));
            CU(cudaHostGetDevicePointer((void **)&s.d_in_conf, s.h_in_conf, 0));
            CU(cudaHostGetDevicePointer((void **)&s.d_in_paf, s.h_in_paf, 0));
        }
        std::memcpy(s.h_in_conf, b.conf, (size_t)n * OPP_N_HEAT * hw * sizeof(float));
        m_conf = s.d_in_conf, m_paf = s.d_in_paf;
        late_paf_copy = true;
    }
    const bool ingest = few_host && m_conf && m_paf;
    // (the limb kernel can only be scheduled early when no up-sampled maps are written on a side stream in between)
    const bool paf_early = ingest && h->pdl && h->paf_early && h->fast_k2 && h->k3_plan.paf_in_smem && !h->trace && !b.conf_up && !b.paf_up;
    if (!ingest) CU(cudaMemsetAsync(s.d_counters, 0, h->counters_ints * sizeof(int), st));

    // ---- inputs
    const float *conf = nullptr, *paf = nullptr;
    const cudaMemcpyKind in_kind = b.in_mem == OPP_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (b.in_layout == OPP_LAYOUT_HWC) {
        if (!s.d_hwc) CU(cudaMalloc(&s.d_hwc, (size_t)c.max_batch * OPP_N_PAF * hw * sizeof(float)));
        const float *src = b.conf;
        if (b.in_mem == OPP_MEM_HOST) {
            CU(cudaMemcpyAsync(s.d_hwc, b.conf, n * OPP_N_HEAT * hw * sizeof(float), in_kind, st));
            src = s.d_hwc;
        }
        CU(launch_hwc_to_chw(src, s.d_conf, n, OPP_N_HEAT, g.h, g.w, st));
        src = b.paf;
        if (b.in_mem == OPP_MEM_HOST) {
            CU(cudaMemcpyAsync(s.d_hwc, b.paf, n * OPP_N_PAF * hw * sizeof(float), in_kind, st));
            src = s.d_hwc;
        }
        CU(launch_hwc_to_chw(src, s.d_paf, n, OPP_N_PAF, g.h, g.w, st));
        h->launches += 2;
        conf = s.d_conf, paf = s.d_paf;
    } else if (ingest && paf_early) {
        // (pageable PAFs: copied into the staging buffer just before the limb kernel is launched, see below)
        // only the heat maps (and the counter reset) go through the ingest kernel; the limb kernel, scheduled early by
        // programmatic dependent launch, pulls each limb's PAF tile from the pinned buffer while the peak kernel runs
        CU(launch_ingest(m_conf, s.d_conf, (size_t)n * OPP_N_HEAT * hw, nullptr, nullptr, 0, s.d_counters, (int)h->counters_ints, st));
        h->launches += 1;
        conf = s.d_conf, paf = m_paf;
    } else if (ingest) {
        if (late_paf_copy) { // the PAFs go through the ingest kernel here: they must be staged first
            std::memcpy(s.h_in_paf, b.paf, (size_t)n * OPP_N_PAF * hw * sizeof(float));
            late_paf_copy = false;
        }
        CU(launch_ingest(m_conf, s.d_conf, (size_t)n * OPP_N_HEAT * hw, m_paf, s.d_paf, (size_t)n * OPP_N_PAF * hw, s.d_counters,
                         (int)h->counters_ints, st));
        h->launches += 1;
        conf = s.d_conf, paf = s.d_paf;
    } else if (b.in_mem == OPP_MEM_HOST && n <= h->zero_copy_in_max && mapped_host(b.conf) && mapped_host(b.paf)) {
        // latency mode: a few frames in pinned memory are read by the kernels straight over PCIe
        // (each feature row is staged into shared memory once per tile anyway); no copy is enqueued
        conf = (const float *)mapped_host(b.conf), paf = (const float *)mapped_host(b.paf);
    } else if (b.in_mem == OPP_MEM_HOST) {
        if (h->h2d_split) { // experiment: heat maps and PAFs on two copy streams at once
            CU(cudaEventRecord(s.ev_h2d0, st));
            CU(cudaStreamWaitEvent(s.side, s.ev_h2d0, 0));
            CU(cudaMemcpyAsync(s.d_paf, b.paf, n * OPP_N_PAF * hw * sizeof(float), in_kind, s.side));
            CU(cudaEventRecord(s.ev_h2d1, s.side));
            CU(cudaMemcpyAsync(s.d_conf, b.conf, n * OPP_N_HEAT * hw * sizeof(float), in_kind, st));
            CU(cudaStreamWaitEvent(st, s.ev_h2d1, 0));
        } else {
            CU(cudaMemcpyAsync(s.d_conf, b.conf, n * OPP_N_HEAT * hw * sizeof(float), in_kind, st));
            CU(cudaMemcpyAsync(s.d_paf, b.paf, n * OPP_N_PAF * hw * sizeof(float), in_kind, st));
        }
        conf = s.d_conf, paf = s.d_paf;
    } else {
        conf = b.conf, paf = b.paf; // device-resident feature maps are used in place
    }

    // ---- optional materialised up-sampled maps, beside the rest (they are outputs only)
    const bool want_up = b.conf_up || b.paf_up;
    // the generic peak kernel reads a materialised heat map only at non-integer scales; at an integer scale its
    // replication-aware form works from the feature maps like the fast kernel
    const bool generic_rep = !h->fast_k2 && g.S > 0 && h->generic_rep;
    // non-integer scales: the generic peak kernel builds its tiles from the feature maps (OPP_GENERIC_VIA_MAP=1: from a
    // materialised heat map, the first form of this path)
    const bool generic_needs_conf_up = !h->fast_k2 && !generic_rep && h->generic_via_map;
    bool forked = false;
    auto resize_on = [&](cudaStream_t rs, const float *src, float *dst, int C, int layout) -> int {
        K1Params k1{};
        k1.g = g, k1.src = src, k1.dst = dst, k1.C = C, k1.n = n, k1.layout = layout;
        CU(launch_k1(k1, rs));
        h->launches += 1;
        return OPP_OK;
    };
    const float *conf_up_for_k2 = nullptr;
    if (generic_needs_conf_up) {
        // the generic peak kernel reads the materialised heat map: CHW, on the main stream
        if (b.conf_up && b.up_layout == OPP_LAYOUT_CHW) {
            int r = resize_on(st, conf, b.conf_up, OPP_N_HEAT, OPP_LAYOUT_CHW);
            if (r) return r;
            conf_up_for_k2 = b.conf_up;
        } else {
            if (!s.d_conf_up) CU(cudaMalloc(&s.d_conf_up, (size_t)c.max_batch * OPP_N_HEAT * HW * sizeof(float)));
            int r = resize_on(st, conf, s.d_conf_up, OPP_N_HEAT, OPP_LAYOUT_CHW);
            if (r) return r;
            conf_up_for_k2 = s.d_conf_up;
        }
    }
    // Integer-scale fast path with both tensors requested channels-first: the peak kernel writes them itself.
    const bool fuse_up = h->fast_k2 && h->fuse_resize && b.conf_up && b.paf_up && b.up_layout == OPP_LAYOUT_CHW &&
                         (((uintptr_t)b.conf_up | (uintptr_t)b.paf_up) & 15) == 0 && (g.W & 3) == 0;
    if (want_up && !fuse_up) {
        CU(cudaEventRecord(s.ev_fork, st));
        CU(cudaStreamWaitEvent(s.side, s.ev_fork, 0));
        forked = true;
        const bool conf_pending = b.conf_up && !(generic_needs_conf_up && b.up_layout == OPP_LAYOUT_CHW);
        if (h->trace) CU(cudaEventRecord(s.tr[0], s.side));
        if (conf_pending && b.paf_up) { // both tensors in one launch
            K1Params k1{};
            k1.g = g, k1.n = n, k1.layout = b.up_layout;
            k1.src = conf, k1.dst = b.conf_up, k1.C = OPP_N_HEAT;
            k1.src2 = paf, k1.dst2 = b.paf_up, k1.C2 = OPP_N_PAF;
            CU(launch_k1(k1, s.side));
            h->launches += 1;
        } else {
            if (conf_pending) {
                int r = resize_on(s.side, conf, b.conf_up, OPP_N_HEAT, b.up_layout);
                if (r) return r;
            }
            if (b.paf_up) {
                int r = resize_on(s.side, paf, b.paf_up, OPP_N_PAF, b.up_layout);
                if (r) return r;
            }
        }
    }

    if (forked && h->trace) CU(cudaEventRecord(s.tr[1], s.side));
    // ---- peaks
    if (h->trace) CU(cudaEventRecord(s.tr[2], st));
    K2Params k2{};
    if (int r = fill_k2(h, s, conf, conf_up_for_k2, n, fuse_up, k2)) return r;
    if (fuse_up) k2.paf = paf, k2.up_conf = b.conf_up, k2.up_paf = b.paf_up;
    if (s.d_times && n == 1) k2.times = s.d_times + (size_t)c.max_batch * OPP_N_PAIRS * 12;
    int *d_flags = cnt_flags(h, s);
    // Latency path (a few frames, nothing else recorded on the stream between the kernels): programmatic dependent
    // launch lets the peak kernel be scheduled behind the ingest kernel and the limb kernel behind the peak kernel
    // while their predecessor still runs; each waits (griddepcontrol.wait) before touching its results.
    const bool few = n <= h->ingest_max && !h->trace && !forked;
    const bool pdl = h->pdl && few;
    if (h->fast_k2 && !few && !forked && (fuse_up ? h->k2_store_low : h->k2_skel_low)) {
        // Full batches: the peak kernel (which wants every SM for a long time) runs on the slot's LOW-priority stream, the
        // limb kernel (little work, latency bound) stays on the high-priority one - the limb kernels of the batches ahead
        // get the CTA slots the peak kernel's CTAs keep freeing instead of queueing behind them.  Materialised 368x432:
        // 163 k -> 175 k frames/s (the three slots' kernels then overlap tail to head: 0.98 of the copy peak over whole
        // batches, above the 0.956 of one fused kernel timed alone); crowded 152 k -> 161 k; every block active, skeleton-only:
        // 391 k -> 432 k.  OPP_K2_LOW=0 keeps both kernels on one stream.
        CU(cudaEventRecord(s.ev_fork, st));
        CU(cudaStreamWaitEvent(s.side, s.ev_fork, 0));
        CU(launch_k2_fast(k2, n, s.side, false));
        CU(cudaEventRecord(s.ev_join, s.side));
        CU(cudaStreamWaitEvent(st, s.ev_join, 0));
    } else if (h->fast_k2) {
        CU(launch_k2_fast(k2, n, st, pdl && ingest));
    } else {
        CU(generic_rep ? launch_k2_generic_rep(k2, n, st) : launch_k2_generic(k2, n, st));
    }
    h->launches += 1;

    if (h->trace) CU(cudaEventRecord(s.tr[3], st));
    // ---- limbs, matching, assembly
    K3Params k3 = h->k3_plan;
    k3.g = g, k3.paf = paf, k3.peaks = s.d_peaks, k3.part_ofs = s.d_part_ofs;
    k3.pk_key = s.d_pk_key, k3.conf = conf, k3.conf_up = conf_up_for_k2;
    k3.capP = c.max_peaks_per_part, k3.capC = c.max_cands_per_limb, k3.capH = c.max_humans;
    k3.cnt = k2.cnt;
    k3.cand_scratch = s.d_cand, k3.conns = s.d_conns, k3.n_conns = s.d_n_conns;
    const bool dev_out = b.out_mem == OPP_MEM_DEVICE;
    // Host results: the assembly writes them over PCIe itself (mapped pinned memory) instead of three
    // device-to-host copies behind the kernel - into the caller's buffers when those are pinned, else
    // into the slot's pinned buffers (copied out in opp_wait).
    s.direct_out = false;
    opp_human_t *k3_humans = s.d_humans;
    int *k3_counts = s.d_n_humans, *k3_flags_out = nullptr;
    if (dev_out) {
        k3_humans = b.humans, k3_counts = b.n_humans;
    } else if (h->zero_copy_out) {
        void *dh = mapped_host(b.humans), *dc = mapped_host(b.n_humans), *df = b.frame_flags ? mapped_host(b.frame_flags) : nullptr;
        if (dh && dc && (df || !b.frame_flags)) {
            k3_humans = (opp_human_t *)dh, k3_counts = (int *)dc, k3_flags_out = df ? (int *)df : s.h_flags;
            s.direct_out = true;
        } else {
            k3_humans = s.h_humans, k3_counts = s.h_n_humans, k3_flags_out = s.h_flags;
        }
    }
    k3.humans = k3_humans;
    k3.n_humans = k3_counts;
    k3.flags = d_flags;
    k3.flags_out = k3_flags_out;
    k3.href_parts = s.d_href_parts, k3.stats = cnt_stats(h, s);
    k3.times = s.d_times;
    // Latency path with host-visible results: completion is announced through a pinned word (see K3Params::host_done)
    s.done_tag = 0;
    if (few && !dev_out && h->zero_copy_out && h->done_flag) {
        s.done_tag = ++h->tag_seq;
        if (s.done_tag == 0) s.done_tag = ++h->tag_seq;
        k3.batch_done = cnt_batch_done(h, s), k3.host_done = (int *)mapped_host(s.h_done);
        k3.done_tag = s.done_tag, k3.done_frames = n;
        if (!k3.host_done) s.done_tag = 0;
    }
    k3.paf_early = paf_early;
    k3.true_index = c.variant == OPP_VARIANT_PYTHON;
    k3.thr_vec = 0.05f, k3.thr_human = 0.4f; // THRESH_VECTOR_SCORE, THRESH_HUMAN_SCORE, src/paf.cpp:61,64
    if (late_paf_copy) std::memcpy(s.h_in_paf, b.paf, (size_t)n * OPP_N_PAF * hw * sizeof(float)); // while ingest + peak kernel run
    CU(launch_k3(k3, n, h->k3_smem, st, pdl && h->fast_k2));
    h->launches += 1;

    if (h->trace) CU(cudaEventRecord(s.tr[4], st));
    // ---- results
    if (dev_out) {
        if (b.frame_flags) CU(cudaMemcpyAsync(b.frame_flags, d_flags, n * sizeof(int), cudaMemcpyDeviceToDevice, st));
    } else if (!h->zero_copy_out) {
        CU(cudaMemcpyAsync(s.h_n_humans, s.d_n_humans, n * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(s.h_flags, d_flags, n * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(s.h_humans, s.d_humans, (size_t)n * c.max_humans * sizeof(opp_human_t), cudaMemcpyDeviceToHost, st));
    }
    if (h->trace) CU(cudaEventRecord(s.tr[5], st));
    if (forked) {
        CU(cudaEventRecord(s.ev_join, s.side));
        CU(cudaStreamWaitEvent(st, s.ev_join, 0));
    }
    CU(cudaEventRecord(s.ev_done, st));
    return OPP_OK;
}

int opp_submit(opp_handle_t h, const opp_batch_t *b, int *ticket)
{
    if (!h || !b || !ticket) return OPP_ERR_INVALID;
    if (b->n_frames < 1 || b->n_frames > h->cfg.max_batch || !b->conf || !b->paf || !b->humans || !b->n_humans) {
        set_err(&h->err, "opp_submit: bad batch (n_frames=%d, max_batch=%d)", b->n_frames, h->cfg.max_batch);
        return OPP_ERR_INVALID;
    }
    if ((b->in_mem != OPP_MEM_HOST && b->in_mem != OPP_MEM_DEVICE) || (b->out_mem != OPP_MEM_HOST && b->out_mem != OPP_MEM_DEVICE) ||
        (b->in_layout != OPP_LAYOUT_CHW && b->in_layout != OPP_LAYOUT_HWC) || (b->up_layout != OPP_LAYOUT_CHW && b->up_layout != OPP_LAYOUT_HWC)) {
        set_err(&h->err, "opp_submit: bad memory kind or layout");
        return OPP_ERR_INVALID;
    }
    if (b->in_sync != OPP_SYNC_NONE && b->in_sync != OPP_SYNC_STREAM && b->in_sync != OPP_SYNC_EVENT) {
        set_err(&h->err, "opp_submit: in_sync must be OPP_SYNC_NONE, OPP_SYNC_STREAM or OPP_SYNC_EVENT");
        return OPP_ERR_INVALID;
    }
    if (b->in_sync == OPP_SYNC_EVENT && !b->in_sync_obj) {
        set_err(&h->err, "opp_submit: OPP_SYNC_EVENT needs a cudaEvent_t in in_sync_obj");
        return OPP_ERR_INVALID;
    }
    DeviceGuard guard_(h->device);
    // first free slot at or after next_slot: tickets may be waited in any order
    const int ns = (int)h->slots.size();
    int si = -1;
    for (int k = 0; k < ns; ++k)
        if (!h->slots[(h->next_slot + k) % ns].busy) {
            si = (h->next_slot + k) % ns;
            break;
        }
    if (si < 0) {
        set_err(&h->err, "opp_submit: all %d slots in flight; call opp_wait first", ns);
        return OPP_ERR_BUSY;
    }
    Slot &s = h->slots[si];
    s.batch = *b, s.n_frames = b->n_frames;
    const int rc = enqueue(h, s, *b);
    if (rc != OPP_OK) {
        cudaStreamSynchronize(s.stream);
        cudaStreamSynchronize(s.side);
        cudaGetLastError();
        return rc;
    }
    s.busy = true;
    s.ticket = h->next_ticket++;
    *ticket = s.ticket;
    h->next_slot = (si + 1) % (int)h->slots.size();
    return OPP_OK;
}

static Slot *slot_of(opp_handle_s *h, int ticket)
{
    for (auto &s : h->slots)
        if (s.ticket == ticket) return &s;
    return nullptr;
}

int opp_wait(opp_handle_t h, int ticket)
{
    if (!h) return OPP_ERR_INVALID;
    Slot *sp = slot_of(h, ticket);
    if (!sp || !sp->busy) {
        set_err(&h->err, "opp_wait: unknown ticket %d", ticket);
        return OPP_ERR_INVALID;
    }
    Slot &s = *sp;
    DeviceGuard guard_(h->device);
    cudaError_t e = cudaSuccess;
    bool retired = true;
    if (s.done_tag) {
        // spin on the completion word; look at the event now and then so that a failed launch cannot hang the caller
        // (acquire: the results in pinned memory are read after the word; on weakly ordered hosts - aarch64 - a plain
        // load would let those reads be satisfied first)
        retired = false;
        for (unsigned spins = 1; __atomic_load_n(s.h_done, __ATOMIC_ACQUIRE) != s.done_tag; ++spins) {
            cpu_relax();
            if ((spins & 0x3fff) == 0) {
                const cudaError_t q = cudaEventQuery(s.ev_done);
                if (q != cudaErrorNotReady) {
                    e = q, retired = true;
                    break;
                }
            }
        }
    } else {
        e = cudaEventSynchronize(s.ev_done);
    }
    s.busy = false;
    if (e != cudaSuccess) {
        set_err(&h->err, "opp_wait: %s", cudaGetErrorString(e));
        return OPP_ERR_CUDA;
    }
    s.ms_pending = !retired; // the kernel may still be retiring: opp_last_batch_ms reads the events when asked
    if (retired) cudaEventElapsedTime(&s.last_ms, s.ev_start, s.ev_done);
    if (h->trace) {
        float t[8] = {0};
        cudaEventElapsedTime(&t[6], h->trace_base, s.ev_start);
        cudaEventElapsedTime(&t[7], h->trace_base, s.ev_done);
        const bool up = s.batch.conf_up || s.batch.paf_up;
        for (int i = 0; i < 6; ++i)
            if (i >= 2 || up) cudaEventElapsedTime(&t[i], h->trace_base, s.tr[i]);
        if (const char *dump = s.d_times ? getenv("OPP_TRACE_DUMP") : nullptr) { // raw stamps of every (frame, limb) of the batch
            std::vector<unsigned long long> all((size_t)s.n_frames * OPP_N_PAIRS * 12);
            cudaMemcpy(all.data(), s.d_times, all.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            if (FILE *f = fopen(dump, "wb")) {
                fwrite(all.data(), sizeof(unsigned long long), all.size(), f);
                fclose(f);
            }
        }
        if (s.d_times && s.n_frames == 1) { // phase stamps of the limb kernel, relative to the earliest CTA start
            unsigned long long t[OPP_N_PAIRS * 12];
            cudaMemcpy(t, s.d_times, sizeof t, cudaMemcpyDeviceToHost);
            unsigned long long t0 = ~0ull;
            for (int l = 0; l < OPP_N_PAIRS; ++l)
                if (t[l * 12] && t[l * 12] < t0) t0 = t[l * 12];
            for (int l = 0; l < OPP_N_PAIRS; ++l) {
                fprintf(stderr, "[opp k3] limb %2d:", l);
                for (int k = 0; k < 11; ++k) fprintf(stderr, " %6.1f", t[l * 12 + k] ? (double)(t[l * 12 + k] - t0) * 1e-3 : -1.0);
                fprintf(stderr, "  (us: start staged scored sorted matched | last: entered staged assembled filtered done tree-limbs-done)\n");
            }
            cudaMemset(s.d_times, 0, sizeof t);
            // peak kernel: earliest start, latest of each phase over the CTAs
            static unsigned long long t2[4096 * 8];
            unsigned long long *d2 = s.d_times + (size_t)h->cfg.max_batch * OPP_N_PAIRS * 12;
            cudaMemcpy(t2, d2, sizeof t2, cudaMemcpyDeviceToHost);
            unsigned long long b0 = ~0ull, mx[8] = {0};
            int nct = 0;
            for (int cta = 0; cta < 4096; ++cta) {
                if (!t2[cta * 8]) continue;
                ++nct;
                if (t2[cta * 8] < b0) b0 = t2[cta * 8];
                for (int k = 0; k < 8; ++k)
                    if (t2[cta * 8 + k] > mx[k]) mx[k] = t2[cta * 8 + k];
            }
            fprintf(stderr, "[opp k2] %d CTAs, latest stamp per phase (us after first CTA start):", nct);
            for (int k = 0; k < 5; ++k) fprintf(stderr, " %6.1f", mx[k] ? (double)(mx[k] - b0) * 1e-3 : -1.0);
            fprintf(stderr, "  (start staged analysed row-pass columns)\n");
            cudaMemset(d2, 0, sizeof t2);
        }
        fprintf(stderr, "[opp trace] ticket %d n=%d start %.3f | k1 %.3f-%.3f | k2 %.3f-%.3f | k3 -%.3f | copies -%.3f | done %.3f (ms)\n", s.ticket,
                s.n_frames, t[6], t[0], t[1], t[2], t[3], t[4], t[5], t[7]);
    }
    const opp_batch_t &b = s.batch;
    if (b.out_mem == OPP_MEM_HOST && !s.direct_out) {
        const int capH = h->cfg.max_humans;
        std::memcpy(b.n_humans, s.h_n_humans, s.n_frames * sizeof(int));
        if (b.frame_flags) std::memcpy(b.frame_flags, s.h_flags, s.n_frames * sizeof(int));
        for (int f = 0; f < s.n_frames; ++f) {
            const int nh = s.h_n_humans[f] < capH ? s.h_n_humans[f] : capH;
            if (nh > 0) std::memcpy(b.humans + (size_t)f * capH, s.h_humans + (size_t)f * capH, nh * sizeof(opp_human_t));
        }
    }
    return OPP_OK;
}

int opp_stream_wait_ticket(opp_handle_t h, int ticket, void *stream)
{
    if (!h) return OPP_ERR_INVALID;
    Slot *sp = slot_of(h, ticket);
    if (!sp || !sp->busy) {
        set_err(&h->err, "opp_stream_wait_ticket: unknown ticket %d", ticket);
        return OPP_ERR_INVALID;
    }
    DeviceGuard guard_(h->device);
    CU(cudaStreamWaitEvent((cudaStream_t)stream, sp->ev_done, 0));
    return OPP_OK;
}

int opp_process(opp_handle_t h, const opp_batch_t *b)
{
    int t = -1;
    int rc = opp_submit(h, b, &t);
    if (rc != OPP_OK) return rc;
    return opp_wait(h, t);
}

int opp_bench_latency(opp_handle_t h, const opp_batch_t *batch, int iters, float *out_us)
{
    if (!h || !batch || !out_us || iters < 1) return OPP_ERR_INVALID;
    for (int i = 0; i < iters; ++i) {
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = opp_process(h, batch);
        const auto t1 = std::chrono::steady_clock::now();
        if (rc != OPP_OK) return rc;
        out_us[i] = std::chrono::duration<float, std::micro>(t1 - t0).count();
    }
    return OPP_OK;
}

float opp_last_batch_ms(opp_handle_t h, int ticket)
{
    if (!h) return -1.f;
    Slot *s = slot_of(h, ticket);
    if (!s) return -1.f;
    if (s->ms_pending && !s->busy) {
        DeviceGuard guard_(h->device);
        if (cudaEventSynchronize(s->ev_done) == cudaSuccess) cudaEventElapsedTime(&s->last_ms, s->ev_start, s->ev_done);
        s->ms_pending = false;
    }
    return s->last_ms;
}

int opp_bench_h2d(opp_handle_t h, const float *conf, const float *paf, int n_frames, int n_batches, int iters, float *out_ms)
{
    if (!h || !conf || !paf || !out_ms || iters < 1 || n_batches < 1 || n_frames < 1 || n_frames > h->cfg.max_batch) return OPP_ERR_INVALID;
    DeviceGuard guard_(h->device);
    for (auto &s : h->slots)
        if (s.busy) return OPP_ERR_BUSY;
    const size_t hw = (size_t)h->g.h * h->g.w;
    int rc = opp_timer_start(h);
    if (rc != OPP_OK) return rc;
    for (int i = 0; i < iters; ++i) {
        Slot &s = h->slots[i % h->slots.size()];
        const size_t k = (size_t)(i % n_batches) * n_frames; // distinct host batches: the host side must not be served from its caches
        CU(cudaMemcpyAsync(s.d_conf, conf + k * OPP_N_HEAT * hw, n_frames * OPP_N_HEAT * hw * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        CU(cudaMemcpyAsync(s.d_paf, paf + k * OPP_N_PAF * hw, n_frames * OPP_N_PAF * hw * sizeof(float), cudaMemcpyHostToDevice, s.stream));
    }
    *out_ms = opp_timer_stop(h);
    return *out_ms < 0.f ? OPP_ERR_CUDA : OPP_OK;
}

int64_t opp_launch_count(opp_handle_t h) { return h ? h->launches : 0; }

int opp_debug_bounds_report(opp_handle_t h, int32_t out[4])
{
    if (!h || !out) return OPP_ERR_INVALID;
    DeviceGuard guard_(h->device);
    int rec[4] = {0, 0, 0, 0};
    const cudaError_t e = opp_kernels_bounds_report(rec, true);
    if (e == cudaErrorNotSupported) {
        set_err(&h->err, "opp_debug_bounds_report: this library was built without -DOPP_DEBUG_BOUNDS");
        return OPP_ERR_INVALID;
    }
    if (e != cudaSuccess) {
        set_err(&h->err, "opp_debug_bounds_report: %s", cudaGetErrorString(e));
        return OPP_ERR_CUDA;
    }
    for (int i = 0; i < 4; ++i) out[i] = rec[i];
    return OPP_OK;
}

int opp_debug_sort(opp_handle_t h, opp_conn_t *cands, int n, int mode, int threads)
{
    if (!h || !cands || n < 0) return OPP_ERR_INVALID;
    DeviceGuard guard_(h->device);
    void *d = nullptr;
    CU(cudaMalloc(&d, (size_t)(n > 0 ? n : 1) * sizeof(opp_conn_t)));
    cudaError_t e = cudaMemcpy(d, cands, (size_t)n * sizeof(opp_conn_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_debug_sort(d, n, mode, threads, nullptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(cands, d, (size_t)n * sizeof(opp_conn_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) {
        set_err(&h->err, "opp_debug_sort: %s", cudaGetErrorString(e));
        return e == cudaErrorInvalidValue ? OPP_ERR_INVALID : OPP_ERR_CUDA;
    }
    return OPP_OK;
}

const char *opp_peak_kernel(opp_handle_t h)
{
    if (!h) return "";
    if (h->fast_k2) return "fast";
    return h->g.S > 0 && h->generic_rep ? "generic_rep" : "generic";
}

int opp_device(opp_handle_t h) { return h ? h->device : -1; }

int opp_debug_fetch(opp_handle_t h, int ticket, int what, int frame, int index, void *dst, int cap)
{
    if (!h || !dst) return -1;
    Slot *sp = slot_of(h, ticket);
    if (!sp || sp->busy || frame < 0 || frame >= sp->n_frames) return -1;
    Slot &s = *sp;
    DeviceGuard guard_(h->device);
    const int capP = h->cfg.max_peaks_per_part, capH = h->cfg.max_humans;
    cudaStreamSynchronize(s.stream); // opp_wait may have returned on the completion word while the kernel was still retiring
    auto d2h = [&](void *d, const void *src, size_t bytes) { return cudaMemcpy(d, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess; };
    switch (what) {
    case OPP_DBG_PEAKS: {
        int ofs[OPP_N_PARTS + 1];
        if (!d2h(ofs, s.d_part_ofs + frame * (OPP_N_PARTS + 1), sizeof ofs)) return -1;
        const int n = ofs[OPP_N_PARTS] < cap ? ofs[OPP_N_PARTS] : cap;
        if (n > 0 && !d2h(dst, s.d_peaks + (size_t)frame * OPP_N_PARTS * capP, n * sizeof(opp_peak_t))) return -1;
        return ofs[OPP_N_PARTS];
    }
    case OPP_DBG_CONNS: {
        if (index < 0 || index >= OPP_N_PAIRS) return -1;
        int n = 0;
        if (!d2h(&n, s.d_n_conns + frame * OPP_N_PAIRS + index, sizeof n)) return -1;
        const int m = n < cap ? n : cap;
        if (m > 0 && !d2h(dst, s.d_conns + ((size_t)frame * OPP_N_PAIRS + index) * capP, m * sizeof(opp_conn_t))) return -1;
        return n;
    }
    case OPP_DBG_PARTS: {
        if (index < 0 || index >= capH || cap < OPP_N_PARTS) return -1;
        if (!d2h(dst, s.d_href_parts + ((size_t)frame * capH + index) * OPP_N_PARTS, OPP_N_PARTS * sizeof(int))) return -1;
        return OPP_N_PARTS;
    }
    case OPP_DBG_COUNTS: { // partial humans, merges, accepted candidates, reserved
        if (cap < 4) return -1;
        if (!d2h(dst, cnt_stats(h, s) + frame * 4, 4 * sizeof(int))) return -1;
        return 4;
    }
    }
    return -1;
}

int opp_resize_device(opp_handle_t h, const float *src, int channels, int n_frames, float *dst, int dst_layout, void *stream)
{
    if (!h || !src || !dst || channels < 1 || n_frames < 1) return OPP_ERR_INVALID;
    DeviceGuard guard_(h->device);
    K1Params k1{};
    k1.g = h->g, k1.src = src, k1.dst = dst, k1.C = channels, k1.n = n_frames, k1.layout = dst_layout;
    CU(launch_k1(k1, stream ? (cudaStream_t)stream : h->slots[0].stream));
    h->launches += 1;
    return OPP_OK;
}

static int fill_k2(opp_handle_s *h, Slot &s, const float *conf, const float *conf_up, int n, bool store, K2Params &k2)
{
    const opp_config_t &c = h->cfg;
    k2.g = h->g, k2.conf = conf, k2.conf_up = conf_up;
    k2.capP = c.max_peaks_per_part;
    k2.cnt.pk_cnt = cnt_pk(h, s), k2.cnt.k2_done = cnt_k2(h, s), k2.cnt.k3_done = cnt_k3(h, s);
    k2.pk_key = s.d_pk_key, k2.peaks = s.d_peaks, k2.part_ofs = s.d_part_ofs;
    k2.flags = cnt_flags(h, s);
    std::memcpy(k2.taps, h->taps, sizeof k2.taps);
    k2.thresh = 0.05f; // THRESH_HEAT, src/paf.cpp:60
    k2.border_zero = c.variant == OPP_VARIANT_PYTHON;
    k2.skip_thresh = h->k2_skip ? k2.thresh * (1.f - 1.f / 8192.f) : -INFINITY;
    if (h->fast_k2) {
        if (!choose_k2_tiles(h, n, store, k2.tw, k2.th)) {
            set_err(&h->err, "no tile plan for the integer-scale peak kernel (tw=%d th=%d)", k2.tw, k2.th);
            return OPP_ERR_INVALID;
        }
        k2.nxs = (h->g.w + k2.tw - 1) / k2.tw, k2.nys = (h->g.h + k2.th - 1) / k2.th;
    }
    return OPP_OK;
}

int opp_resize_pair_device(opp_handle_t h, const float *conf, const float *paf, int n_frames, float *conf_up, float *paf_up, int dst_layout,
                           void *stream)
{
    if (!h || !conf || !paf || !conf_up || !paf_up || n_frames < 1) return OPP_ERR_INVALID;
    DeviceGuard guard_(h->device);
    K1Params k1{};
    k1.g = h->g, k1.n = n_frames, k1.layout = dst_layout;
    k1.src = conf, k1.dst = conf_up, k1.C = OPP_N_HEAT;
    k1.src2 = paf, k1.dst2 = paf_up, k1.C2 = OPP_N_PAF;
    CU(launch_k1(k1, stream ? (cudaStream_t)stream : h->slots[0].stream));
    h->launches += 1;
    return OPP_OK;
}

int opp_peaks_device(opp_handle_t h, const float *conf, const float *paf, int n_frames, float *conf_up, float *paf_up, void *stream)
{
    if (!h || !conf || n_frames < 1 || n_frames > h->cfg.max_batch) return OPP_ERR_INVALID;
    if (!h->fast_k2) {
        set_err(&h->err, "opp_peaks_device: only the integer-scale peak kernel runs stand-alone");
        return OPP_ERR_INVALID;
    }
    if ((conf_up || paf_up) && !(conf_up && paf_up && paf)) return OPP_ERR_INVALID;
    DeviceGuard guard_(h->device);
    Slot &s = h->slots[0];
    if (s.busy) return OPP_ERR_BUSY;
    cudaStream_t st = stream ? (cudaStream_t)stream : s.stream;
    CU(cudaMemsetAsync(s.d_counters, 0, h->counters_ints * sizeof(int), st));
    K2Params k2{};
    if (int r = fill_k2(h, s, conf, nullptr, n_frames, conf_up != nullptr, k2)) return r;
    if (conf_up) k2.paf = paf, k2.up_conf = conf_up, k2.up_paf = paf_up;
    CU(launch_k2_fast(k2, n_frames, st));
    h->launches += 1;
    return OPP_OK;
}

int opp_timer_start(opp_handle_t h)
{
    if (!h) return OPP_ERR_INVALID;
    DeviceGuard guard_(h->device);
    for (auto &s : h->slots) {
        CU(cudaEventRecord(s.ev_join, s.stream));
        CU(cudaStreamWaitEvent(h->timer_stream, s.ev_join, 0));
    }
    CU(cudaEventRecord(h->timer_t0, h->timer_stream));
    // work submitted from now on must not start before t0
    for (auto &s : h->slots) CU(cudaStreamWaitEvent(s.stream, h->timer_t0, 0));
    return OPP_OK;
}

float opp_timer_stop(opp_handle_t h)
{
    if (!h) return -1.f;
    DeviceGuard guard_(h->device);
    for (auto &s : h->slots) {
        if (cudaEventRecord(s.ev_join, s.stream) != cudaSuccess) return -1.f;
        if (cudaStreamWaitEvent(h->timer_stream, s.ev_join, 0) != cudaSuccess) return -1.f;
    }
    if (cudaEventRecord(h->timer_t1, h->timer_stream) != cudaSuccess) return -1.f;
    if (cudaEventSynchronize(h->timer_t1) != cudaSuccess) return -1.f;
    float ms = -1.f;
    cudaEventElapsedTime(&ms, h->timer_t0, h->timer_t1);
    return ms;
}

void process_conf_paf(int height, int width, int n_joins_, int n_connections_, const float *peaks_, const float *pafmap_)
{
    opp_config_t cfg;
    opp_config_default(&cfg, height, width, 8 * height, 8 * width, 17);
    cfg.n_joins = n_joins_, cfg.n_connections = n_connections_, cfg.max_batch = 1, cfg.n_slots = 1;
    opp_handle_t h = nullptr;
    if (opp_create(&cfg, &h) != OPP_OK) {
        fprintf(stderr, "process_conf_paf: %s\n", opp_last_error(nullptr));
        return;
    }
    std::vector<opp_human_t> humans(cfg.max_humans);
    int n = 0, flags = 0;
    opp_batch_t b{};
    b.conf = peaks_, b.paf = pafmap_, b.n_frames = 1, b.in_mem = OPP_MEM_HOST, b.out_mem = OPP_MEM_HOST;
    b.humans = humans.data(), b.n_humans = &n, b.frame_flags = &flags;
    if (opp_process(h, &b) != OPP_OK) {
        fprintf(stderr, "process_conf_paf: %s\n", opp_last_error(h));
    } else {
        for (int i = 0; i < n; ++i) { // human_t::print, include/openpose-plus/human.h:21-31
            for (int j = 0; j < OPP_N_PARTS; ++j) {
                const opp_body_part_t &bp = humans[i].parts[j];
                if (bp.has_value) printf("BodyPart:%d-(%.2f, %.2f) score=%.2f ", j, bp.x, bp.y, bp.score);
            }
            printf("score=%.2f\n", humans[i].score);
        }
    }
    opp_destroy(h);
}

} // extern "C"
